// bc_partition.cu — deferred, partitioned counting: the K3 step (info.rs:735-808) done once per job instead of read
// by read.
//
// Updating a (key, UMI) set and a key -> count map of several GB read by read costs two random DRAM sectors per matched
// read (with their read-modify-write at the memory controller ~190 B of traffic for an 8-byte record): on the DEL
// workload that was 40 % of k_decode.  Here k_decode only stores the packed key of every matched read into a flat
// record buffer (8 or 16 bytes per read, coalesced), and the flush
//   stage A  hash-partitions the records by (key, UMI) into pieces that fit a shared-memory table, one CTA per piece
//            drops the repeats (info.rs:780-791: only the first insert of a pair counts) and combines the pairs of
//            one key into (key, pairs) items;
//   stage B  hash-partitions those items by key and sums them per key in shared memory: the final (key, count) rows
//            (output.rs:265-270).
// Schemes without a random barcode skip the partitioning of stage A: contiguous chunks of records are pre-aggregated
// into (key, reads) items.  Combining before stage B also bounds the work a hot key (Zipf-distributed lineage
// barcodes) puts on the one CTA that owns it.  Every pass streams its input once with full-width transactions; the
// random accesses stay in shared memory.  A partition that holds more distinct keys than the table (cannot happen
// with uniform hashing: the fill target is 30 standard deviations below the capacity) raises `overflow` and the host
// falls back to the global-memory tables of bc_device.cuh.
#include "../../include/bc_b200.h"
#include <algorithm>

#include "bc_kernels.h"

namespace bc {

namespace {

constexpr uint32_t kEmpty32 = 0xFFFFFFFFu;
constexpr int kRedThreads = 256;
// (the BC_* macros exist so that tools/build_variants.py can compile alternatives for A/B timing; the defaults ship)
#ifndef BC_TABLE_SLOTS
#define BC_TABLE_SLOTS 4096
#endif
#ifndef BC_CHAIN_LIMIT
#define BC_CHAIN_LIMIT 40
#endif
#ifndef BC_SCATTER_IPT
#define BC_SCATTER_IPT 16
#endif
#ifndef BC_REDUCE_BLOCKS
#define BC_REDUCE_BLOCKS 4
#endif
constexpr uint32_t kTableSlots = BC_TABLE_SLOTS;  // u32 slots holding key-store indices
constexpr uint32_t kKeyCap = 2048;      // distinct keys a CTA can hold (narrow 48 KB, wide 64 KB of shared memory: 4 / 3 CTAs per SM)
constexpr int kRedLoads = kKeyCap / kRedThreads;  // items per thread of a partition that fits the key store

constexpr int kSplitThreads = 512;      // histogram kernel
constexpr int kScatterThreads = 256;    // scatter kernel: ~80 registers per thread, three CTAs per SM to overlap its phases
constexpr uint32_t kSplitMaxBits = 11;    // bins per level <= 2048 (shared-memory histogram)

__device__ __forceinline__ bool item_valid(unsigned long long lo, unsigned long long hi, bool wide) {
    return wide ? hi != kEmpty : lo != kEmpty;
}

// One CTA per segment of F bins: starts[s * F + b] = seg_base[s] + exclusive prefix of the segment's histogram
// (seg_base == nullptr: a single segment starting at 0), cursor = copy, starts[n_seg * F] = grand total.
__global__ void __launch_bounds__(512) k_seg_scan(const uint32_t* __restrict__ hist, const uint32_t F, const uint32_t* __restrict__ seg_base,
                                                  uint32_t* __restrict__ starts, uint32_t* __restrict__ cursor, FlushStats* stats,
                                                  const uint32_t limit) {
    __shared__ uint32_t s_warp[16];
    __shared__ uint32_t s_carry;
    uint32_t biggest = 0;
    unsigned long long big_items = 0;
    const uint32_t seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned long long off = (unsigned long long)seg * F;
    if (tid == 0) s_carry = seg_base ? seg_base[seg] : 0u;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < F; b0 += 512) {
        const uint32_t b = b0 + tid;
        const uint32_t v = b < F ? hist[off + b] : 0u;
        biggest = max(biggest, v);
        if (v > limit) big_items += v;
        uint32_t x = v;  // inclusive warp scan
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        uint32_t before = s_carry;
        for (uint32_t w = 0; w < wid; w++) before += s_warp[w];
        if (b < F) {
            starts[off + b] = before + x - v;
            cursor[off + b] = before + x - v;
        }
        __syncthreads();
        if (tid == 511) s_carry = before + x;
        __syncthreads();
    }
    if (seg == gridDim.x - 1 && tid == 0) starts[off + F] = s_carry;
    if (stats) {  // largest bin: the host checks that every partition fits a CTA's key store
        biggest = __reduce_max_sync(0xFFFFFFFFu, biggest);
        if (lane == 0 && biggest) atomicMax(&stats->max_bin, (unsigned long long)biggest);
        if (__any_sync(0xFFFFFFFFu, big_items != 0)) {
            for (int o = 16; o; o >>= 1) big_items += __shfl_xor_sync(0xFFFFFFFFu, big_items, o);
            if (lane == 0) atomicAdd(&stats->big_items, big_items);
        }
    }
}

// ---- hash partitioning, one radix level per launch --------------------------------------------------------------
// The partition of an item is p = mulhi(hash(key), P).  With more than 2048 partitions P = F1 * 2^l2: level 1 splits
// the whole input into F1 segments (digit p >> l2), level 2 splits every segment into 2^l2 partitions (digit
// p & (2^l2 - 1)); a level's digit is (p >> shift) & mask.
// The scatter kernel takes a tile of kSplitThreads * IPT consecutive items of one segment into registers (all loads
// of a thread in flight at once) and
//   1  counts the tile per bin in shared memory, turns the counts into offsets of a bin-sorted copy of the tile and
//      reserves one output run per non-empty bin with a single global atomic;
//   2  puts every item, with its output position, at its place in the sorted copy (shared memory);
//   3  writes the copy out: consecutive lanes hold consecutive items of a run, so the stores leave the SM as full
//      32-byte sectors.
// Versions measured and dropped: one global atomic per item (55 G atomics/s, ten times below the streaming rate of
// the same data); per-item stores straight to the output (partial-sector writes: the L2 reads each sector before
// merging 8 bytes into it — 2x the DRAM reads, 60 G stores/s); re-reading the tile from L2 for every sweep instead of
// keeping it in registers (latency-bound at 25-50 % occupancy: 280-550 us for 33 M items).
#ifndef BC_PART_HASH32
#define BC_PART_HASH32 0
#endif
// 64 (or 128) key bits -> 32 well-mixed bits in a dozen 32-bit instructions: one multiply per key word, then the two-round
// xorshift-multiply finaliser.  The partition hash is evaluated four times per record by the flush (histogram and scatter
// of two levels) and was a third of those kernels' instructions as a full 64-bit mix.
[[maybe_unused]] __device__ __forceinline__ uint32_t part_hash32(Key k) {
    uint32_t x = (uint32_t)k.lo * 0x85EBCA6Bu ^ (uint32_t)(k.lo >> 32) * 0xC2B2AE35u ^ (uint32_t)k.hi * 0x27D4EB2Fu ^
                 (uint32_t)(k.hi >> 32) * 0x165667B1u;
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

__device__ __forceinline__ uint32_t split_digit(unsigned long long lo, unsigned long long hi, const SplitLevel& lv) {
    // drop_bits > 0: partition by the record without its random barcode, so that a partition holds whole keys
    Key k = lv.drop_bits ? key_shr(Key{lo, hi}, lv.drop_bits) : Key{lo, hi};
    k.lo ^= lv.salt;  // the hot partitions set aside by a by-key pass share their hash bits: re-partition them with another hash
#if BC_PART_HASH32
    const uint32_t p = __umulhi(part_hash32(k), (uint32_t)lv.P);  // P < 2^32 (partition_items)
#else
    const uint32_t p = (uint32_t)__umul64hi(hash_key(k), lv.P);
#endif
    return (p >> lv.shift) & lv.mask;
}

template <bool WIDE>
__global__ void __launch_bounds__(kSplitThreads) k_split_hist(const ItemView in, const uint32_t* __restrict__ seg_starts,
                                                              const uint32_t workers, const unsigned long long n_total,
                                                              const SplitLevel lv, uint32_t* __restrict__ bins, FlushStats* stats,
                                                              const bool count_valid) {
    __shared__ uint32_t s_hist[1u << kSplitMaxBits];
    __shared__ unsigned long long s_valid;
    constexpr int U = 4;
    const uint32_t seg = blockIdx.x / workers, worker = blockIdx.x % workers;
    const unsigned long long a = seg_starts ? (unsigned long long)seg_starts[seg] : 0ULL;
    const unsigned long long e = seg_starts ? (unsigned long long)seg_starts[seg + 1] : n_total;
    uint32_t* my_bins = bins + (unsigned long long)seg * lv.F;
    const uint32_t tid = threadIdx.x;
    unsigned long long valid = 0;
    if (tid == 0) s_valid = 0;
    for (uint32_t b = tid; b < lv.F; b += kSplitThreads) s_hist[b] = 0;
    __syncthreads();
    const unsigned long long span = (unsigned long long)U * kSplitThreads;
    for (unsigned long long t0 = a + worker * span; t0 < e; t0 += workers * span) {
        unsigned long long lo[U], hi[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            const unsigned long long i = t0 + (unsigned long long)k * kSplitThreads + tid;
            lo[k] = i < e ? in.lo[i] : kEmpty;
            hi[k] = WIDE ? (i < e ? in.hi[i] : kEmpty) : 0ULL;
        }
#pragma unroll
        for (int k = 0; k < U; k++) {
            if (!item_valid(lo[k], hi[k], WIDE)) continue;
            valid++;
            atomicAdd(&s_hist[split_digit(lo[k], hi[k], lv)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t b = tid; b < lv.F; b += kSplitThreads) {
        const uint32_t c = s_hist[b];
        if (c) atomicAdd(&my_bins[b], c);
    }
    if (count_valid) {
        for (int o = 16; o; o >>= 1) valid += __shfl_xor_sync(0xFFFFFFFFu, valid, o);
        if ((tid & 31) == 0 && valid) atomicAdd(&s_valid, valid);
        __syncthreads();
        if (tid == 0 && s_valid) atomicAdd(&stats->valid, s_valid);
    }
}

template <bool WIDE, bool WEIGHTED, int IPT>
__global__ void __launch_bounds__(kScatterThreads, 3) k_split_scatter(const ItemView in, const ItemView out,
                                                                    const uint32_t* __restrict__ seg_starts, const uint32_t workers,
                                                                    const unsigned long long n_total, const SplitLevel lv,
                                                                    uint32_t* __restrict__ bins) {
    constexpr uint32_t T = kScatterThreads * IPT;
    extern __shared__ __align__(16) unsigned long long split_stage[];  // [lo | hi | w] x T u64, then T u32 positions
    __shared__ uint32_t s_hist[1u << kSplitMaxBits], s_delta[1u << kSplitMaxBits];
    __shared__ uint32_t s_warp[kScatterThreads / 32];
    const uint32_t seg = blockIdx.x / workers, worker = blockIdx.x % workers;
    const unsigned long long a = seg_starts ? (unsigned long long)seg_starts[seg] : 0ULL;
    const unsigned long long e = seg_starts ? (unsigned long long)seg_starts[seg + 1] : n_total;
    const uint32_t F = lv.F;
    uint32_t* my_bins = bins + (unsigned long long)seg * F;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned long long* st_lo = split_stage;
    unsigned long long* st_hi = st_lo + (WIDE ? T : 0);
    unsigned long long* st_w = st_hi + (WEIGHTED ? T : 0);
    uint32_t* st_pos = reinterpret_cast<uint32_t*>(st_w + T);
    for (unsigned long long t0 = a + (unsigned long long)worker * T; t0 < e; t0 += (unsigned long long)workers * T) {
        unsigned long long lo[IPT], hi[IPT], w[IPT];
        uint32_t bin2[IPT / 2];  // two 16-bit bins per register (F <= 2048; 0xFFFF = hole)
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const unsigned long long i = t0 + (unsigned long long)k * kScatterThreads + tid;
            lo[k] = i < e ? in.lo[i] : kEmpty;
            hi[k] = WIDE ? (i < e ? in.hi[i] : kEmpty) : 0ULL;
            w[k] = WEIGHTED ? (i < e ? in.w[i] : 0ULL) : 0ULL;
        }
        for (uint32_t b = tid; b < F; b += kScatterThreads) s_hist[b] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            uint32_t bn = 0xFFFFu;
            if (item_valid(lo[k], hi[k], WIDE)) {
                bn = split_digit(lo[k], hi[k], lv);
                atomicAdd(&s_hist[bn], 1u);
            }
            if (k & 1) bin2[k >> 1] |= bn << 16;
            else bin2[k >> 1] = bn;
        }
        __syncthreads();
        // exclusive prefix of the tile's histogram (thread t owns bins [t * per, t * per + per)); run reservation
        const uint32_t per = (F + kScatterThreads - 1) / kScatterThreads;  // <= 8
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t b = tid * per + k;
            c[k] = (k < (int)per && b < F) ? s_hist[b] : 0u;
            sum += c[k];
        }
        uint32_t x = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        uint32_t off = x - sum, n_tile = 0;
        for (uint32_t ww = 0; ww < kScatterThreads / 32; ww++) {
            if (ww < wid) off += s_warp[ww];
            n_tile += s_warp[ww];
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t b = tid * per + k;
            if (k < (int)per && b < F) {
                const uint32_t g = c[k] ? atomicAdd(&my_bins[b], c[k]) : 0u;
                s_delta[b] = g - off;  // output position = s_delta[bin] + position in the sorted tile
                s_hist[b] = off;       // becomes the bin's cursor inside the sorted tile
                off += c[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const uint32_t bn = (k & 1) ? bin2[k >> 1] >> 16 : bin2[k >> 1] & 0xFFFFu;
            if (bn == 0xFFFFu) continue;
            const uint32_t at = atomicAdd(&s_hist[bn], 1u);
            st_lo[at] = lo[k];
            if (WIDE) st_hi[at] = hi[k];
            if (WEIGHTED) st_w[at] = w[k];
            st_pos[at] = s_delta[bn] + at;
        }
        __syncthreads();
        for (uint32_t j = tid; j < n_tile; j += kScatterThreads) {
            const uint32_t pos = st_pos[j];
            out.lo[pos] = st_lo[j];
            if (WIDE) out.hi[pos] = st_hi[j];
            if (WEIGHTED) out.w[pos] = st_w[j];
        }
        __syncthreads();
    }
}

// ---- the exchange's own scatter: at most kMaxRanks (8) bins ----------------------------------------------------------
// With eight bins the tile histogram of k_split_scatter is eight shared-memory words that 256 threads hammer twice per
// item.  Here a thread counts its items per owner in registers (8 x 16-bit fields in two words), a warp scan and one
// word pair per warp in shared memory turn the counts into every thread's own offsets inside the owner-sorted tile, and
// a thread places its items with register arithmetic only: no shared-memory atomics, a quarter of the instructions.
// Runs are reserved like k_split_scatter's, per tile and bin: streamed, with one system-scope atomic on the owner's receive cursor
// (peers.cursor[b], in the owner's memory); bulk, on the local cursors that the host primed with the plan's positions.
__device__ __forceinline__ unsigned long long shfl_up64(unsigned long long v, int d) {
    const uint32_t lo = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)v, d), hi = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), d);
    return ((unsigned long long)hi << 32) | lo;
}

template <bool WIDE, int IPT>
__global__ void __launch_bounds__(kScatterThreads, 3) k_owner_scatter(const ItemView in, const PeerOut peers, const uint32_t workers,
                                                                    const unsigned long long n_total, const SplitLevel lv,
                                                                    uint32_t* __restrict__ bins) {
    constexpr uint32_t T = kScatterThreads * IPT;
    constexpr int NW = kScatterThreads / 32;
    extern __shared__ __align__(16) unsigned long long split_stage[];  // [lo | hi] x T u64, then T u32 positions
    __shared__ unsigned long long s_wtot[NW][2];
    __shared__ unsigned long long s_delta[kMaxRanks];  // owner position of the sorted tile's slot 0 of that bin's run (mod 2^64)
    __shared__ unsigned long long* s_lo[kMaxRanks];
    __shared__ unsigned long long* s_hi[kMaxRanks];
    __shared__ uint32_t s_fits;
    const uint32_t F = lv.F, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned long long* st_lo = split_stage;
    unsigned long long* st_hi = st_lo + (WIDE ? T : 0);
    uint32_t* st_pos = reinterpret_cast<uint32_t*>(st_hi + T);
    if (tid < kMaxRanks) {
        s_lo[tid] = peers.lo[tid];
        s_hi[tid] = peers.hi[tid];
    }
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * T; t0 < n_total; t0 += (unsigned long long)workers * T) {
        unsigned long long lo[IPT], hi[IPT];
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const unsigned long long i = t0 + (unsigned long long)k * kScatterThreads + tid;
            lo[k] = i < n_total ? in.lo[i] : kEmpty;
            hi[k] = WIDE ? (i < n_total ? in.hi[i] : kEmpty) : 0ULL;
        }
        // owner of every item (4 bits each; 15 = hole) and this thread's count per owner: fields of 16 bits, owners 0-3 | 4-7
        unsigned long long own = 0, c0 = 0, c1 = 0;
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            uint32_t b = 15u;
            if (item_valid(lo[k], hi[k], WIDE)) {
                b = split_digit(lo[k], hi[k], lv);
                const unsigned long long one = 1ULL << (16 * (b & 3u));
                if (b & 4u) c1 += one;
                else c0 += one;
            }
            own |= (unsigned long long)b << (4 * k);
        }
        unsigned long long x0 = c0, x1 = c1;  // inclusive warp scan of both words (a field never passes 32 x IPT)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y0 = shfl_up64(x0, o), y1 = shfl_up64(x1, o);
            if (lane >= (unsigned)o) {
                x0 += y0;
                x1 += y1;
            }
        }
        if (lane == 31) {
            s_wtot[wid][0] = x0;
            s_wtot[wid][1] = x1;
        }
        if (tid == 0) s_fits = 0xFFFFFFFFu;
        __syncthreads();
        unsigned long long before0 = 0, before1 = 0, tot0 = 0, tot1 = 0;  // items of the warps before this one / of the tile, per owner
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const unsigned long long a0 = s_wtot[w][0], a1 = s_wtot[w][1];
            if (w < (int)wid) {
                before0 += a0;
                before1 += a1;
            }
            tot0 += a0;
            tot1 += a1;
        }
        // start of every owner's run inside the sorted tile: exclusive prefix over the eight totals, again as packed fields
        // (a tile holds at most 4096 items: 16 bits are enough)
        unsigned long long start0 = (tot0 << 16) + (tot0 << 32) + (tot0 << 48);  // fields: 0, t0, t0+t1, t0+t1+t2
        const uint32_t sum0 = (uint32_t)(((tot0 & 0xFFFFULL) + ((tot0 >> 16) & 0xFFFFULL) + ((tot0 >> 32) & 0xFFFFULL) + (tot0 >> 48)));
        unsigned long long start1 = (tot1 << 16) + (tot1 << 32) + (tot1 << 48);
        start1 += (unsigned long long)sum0 * 0x0001000100010001ULL;
        const uint32_t n_tile = sum0 + (uint32_t)(((tot1 & 0xFFFFULL) + ((tot1 >> 16) & 0xFFFFULL) + ((tot1 >> 32) & 0xFFFFULL) + (tot1 >> 48)));
        if (tid < F) {  // one run per owner and tile
            const uint32_t b = tid;
            const uint32_t cnt = (uint32_t)(((b & 4u) ? tot1 : tot0) >> (16 * (b & 3u))) & 0xFFFFu;
            const uint32_t st = (uint32_t)(((b & 4u) ? start1 : start0) >> (16 * (b & 3u))) & 0xFFFFu;
            unsigned long long g = 0;
            if (cnt) {
                if (peers.cursor[b]) {
                    g = atomicAdd_system(peers.cursor[b], (unsigned long long)cnt);
                    atomicAdd(&peers.sent[b], (unsigned long long)cnt);
                    if (g + cnt > peers.cap) {
                        atomicAnd(&s_fits, ~(1u << b));
                        atomicExch(peers.overflow, 1u);
                    }
                } else {
                    g = atomicAdd(&bins[b], cnt);
                }
            }
            s_delta[b] = g - st;
        }
        // this thread's first slot per owner = run start + items of earlier warps + items of earlier lanes
        unsigned long long off0 = start0 + before0 + (x0 - c0), off1 = start1 + before1 + (x1 - c1);
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const uint32_t b = (uint32_t)(own >> (4 * k)) & 15u;
            if (b == 15u) continue;
            const uint32_t sh = 16 * (b & 3u);
            const uint32_t at = (uint32_t)(((b & 4u) ? off1 : off0) >> sh) & 0xFFFFu;
            if (b & 4u) off1 += 1ULL << sh;
            else off0 += 1ULL << sh;
            st_lo[at] = lo[k];
            if (WIDE) st_hi[at] = hi[k];
            st_pos[at] = b;  // the owner for now: its run's position is only known after the barrier
        }
        __syncthreads();
        const uint32_t fits = s_fits;
        for (uint32_t j = tid; j < n_tile; j += kScatterThreads) {
            const uint32_t b = st_pos[j];
            if (!((fits >> b) & 1u)) continue;
            const unsigned long long pos = s_delta[b] + j;
            s_lo[b][pos] = st_lo[j];
            if (WIDE) s_hi[b][pos] = st_hi[j];
        }
        __syncthreads();
    }
}

// ---- the shared-memory table -----------------------------------------------------------------------------------
// table[] holds indices into a key store (klo/khi/kcnt); an index is published with a 32-bit CAS on the table slot
// only after the key behind it is in place, so wide keys need no 128-bit shared-memory CAS.
//   partition fits the key store (the normal case): the items themselves are the key store — item j of the partition
//     is staged at index j, no allocation at all;
//   larger partition (a (key, UMI) pair repeated thousands of times, a hot key's partial counts): items stream
//     through, a thread reserves a key-store index for each new key from a shared counter, writes the key, fences and
//     publishes.  A reservation whose CAS lost the race is kept for the thread's next new key; what is left over at
//     the end has count 0 and is skipped.
template <bool WIDE>
struct SmemTable {
    uint32_t* table;
    unsigned long long* klo;
    unsigned long long* khi;
    uint32_t* kcnt;
    uint32_t* n_perm;  // streaming path: shared counter of reserved key-store entries (may run past kKeyCap: overflow)
};

template <bool WIDE>
__device__ __forceinline__ uint32_t smem_find_or_claim(const SmemTable<WIDE>& t, unsigned long long lo, unsigned long long hi,
                                                       uint32_t h, uint32_t& reserve) {
    uint32_t s = h & (kTableSlots - 1);
    for (;;) {
        uint32_t v = *reinterpret_cast<volatile uint32_t*>(&t.table[s]);
        if (v == kEmpty32) {
            if (reserve == kEmpty32) {
                reserve = atomicAdd(t.n_perm, 1u);
                if (reserve >= kKeyCap) {  // key store full: the CTA reports overflow after the pass
                    reserve = kEmpty32;
                    return kEmpty32;
                }
                t.kcnt[reserve] = 0u;
            }
            t.klo[reserve] = lo;
            if (WIDE) t.khi[reserve] = hi;
            __threadfence_block();
            const uint32_t old = atomicCAS(&t.table[s], kEmpty32, reserve);
            if (old == kEmpty32) {
                const uint32_t mine = reserve;
                reserve = kEmpty32;
                return mine;
            }
            v = old;
        }
        const unsigned long long a = *reinterpret_cast<volatile unsigned long long*>(&t.klo[v]);
        if (a == lo && (!WIDE || *reinterpret_cast<volatile unsigned long long*>(&t.khi[v]) == hi)) return v;
        s = (s + 1) & (kTableSlots - 1);
    }
}

// staged path: the key of item j is already at klo/khi[j]; returns the index that owns the key (j itself when new)
template <bool WIDE>
__device__ __forceinline__ uint32_t smem_find_or_claim_staged(const SmemTable<WIDE>& t, unsigned long long lo, unsigned long long hi,
                                                              uint32_t h, uint32_t j) {
    uint32_t s = h & (kTableSlots - 1);
    for (;;) {
        uint32_t v = *reinterpret_cast<volatile uint32_t*>(&t.table[s]);
        if (v == kEmpty32) {
            const uint32_t old = atomicCAS(&t.table[s], kEmpty32, j);
            if (old == kEmpty32) return j;
            v = old;
        }
        if (t.klo[v] == lo && (!WIDE || t.khi[v] == hi)) return v;
        s = (s + 1) & (kTableSlots - 1);
    }
}

// slot hash inside a CTA: independent of the 64-bit hash that chose the partition, and a handful of instructions
template <bool WIDE>
__device__ __forceinline__ uint32_t slot_hash(unsigned long long lo, unsigned long long hi) {
    uint32_t x = (uint32_t)lo * 0x9E3779B1u ^ (uint32_t)(lo >> 32) * 0x85EBCA77u;
    if (WIDE) x ^= (uint32_t)hi * 0xC2B2AE3Du ^ (uint32_t)(hi >> 32) * 0x27D4EB2Fu;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    x ^= x >> 13;
    return x;
}
template <bool WIDE>
__device__ __forceinline__ Key drop_umi(unsigned long long lo, unsigned long long hi, uint32_t umi_bits) {
    if (!WIDE) return Key{lo >> umi_bits, 0ULL};  // umi_bits < 64 whenever the record is narrow
    return key_shr(Key{lo, hi}, umi_bits);
}

// Weights travel as u64 but are summed in 32-bit shared-memory counters: a weight or a per-key sum of 2^32 or more raises
// `overflow`, and the host redoes the flush through the global tables, which add in 64 bits (the reference counts in usize).
__device__ __forceinline__ uint32_t narrow_weight(unsigned long long w, FlushStats* stats) {
    if (w >> 32) atomicExch(&stats->overflow, 16ULL);
    return (uint32_t)w;
}
__device__ __forceinline__ void add_weight(uint32_t* c, uint32_t w, FlushStats* stats) {
    const uint32_t old = atomicAdd(c, w);
    if (old + w < old) atomicExch(&stats->overflow, 16ULL);
}

__device__ __forceinline__ void clear_u32(uint32_t* p, uint32_t n, uint32_t value) {  // n multiple of 4, p 16-byte aligned
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint4 v = make_uint4(value, value, value, value);
    for (uint32_t i = threadIdx.x; i < n / 4; i += kRedThreads) q[i] = v;
}

// MODE RED_DEDUPE      : in = records (key incl. random barcode), out = (record >> umi_bits, distinct records with that key)
// MODE RED_COUNT       : in = (key, weight) items,                out = (key, sum of weights)
// MODE RED_DEDUPE_KEYED: RED_DEDUPE for partitions that hold whole keys (partitioned by the record WITHOUT its random
//   barcode) and fit the key store — one pass instead of two.  A record's home slot is the hash of its KEY, so all records
//   of a key walk the same probe sequence; slots of a sequence fill in order and never empty, hence the first entry
//   of that key a record meets on its way (or the record itself when it meets none) is the same for all of them: that
//   entry carries the key's count of distinct records.  The price is that a key's records chain behind one home slot:
//   fine for the usual one to a few random barcodes per key, quadratic for a key with hundreds.  A lane that has walked
//   kChainLimit slots for one record gives the partition up: nothing is emitted, its number goes on `hot_list`, and a
//   second launch puts the listed partitions through the two-pass RED_DEDUPE, whose cost does not depend on the keys.
constexpr uint32_t kChainLimit = BC_CHAIN_LIMIT;

template <bool WIDE, int MODE>
__device__ __forceinline__ void reduce_one(const unsigned long long p, const ItemView& in, const uint32_t* __restrict__ starts,
                                           const unsigned long long n_items, const uint32_t chunk, const uint32_t umi_bits,
                                           const ItemView& out, const unsigned long long out_cap, FlushStats* stats,
                                           const uint32_t skip_over, uint32_t* hot_list, uint32_t* hot_n) {
    extern __shared__ __align__(16) unsigned char red_smem[];
    __shared__ uint32_t s_nperm, s_warp[kRedThreads / 32], s_abort;
    __shared__ unsigned long long s_base;

    SmemTable<WIDE> t;
    t.klo = reinterpret_cast<unsigned long long*>(red_smem);
    t.khi = WIDE ? t.klo + kKeyCap : nullptr;
    t.table = reinterpret_cast<uint32_t*>(t.klo + (WIDE ? 2 : 1) * kKeyCap);
    t.kcnt = t.table + kTableSlots;
    uint32_t* kw = MODE == RED_DEDUPE_KEYED ? t.kcnt : t.kcnt + kKeyCap;  // RED_DEDUPE*: distinct records per key
    t.n_perm = &s_nperm;

    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned long long a = starts ? (unsigned long long)starts[p] : p * chunk;
    const unsigned long long e = starts ? (unsigned long long)starts[p + 1] : min(n_items, a + chunk);
    if (e <= a) return;  // empty partition: nothing to emit
    if (skip_over && e - a > skip_over) return;  // hot keys: k_gather_big hands this partition to the two-stage path

    clear_u32(t.table, kTableSlots, kEmpty32);
    if (MODE != RED_COUNT) clear_u32(kw, kKeyCap, 0u);
    if (tid == 0) {
        s_nperm = 0;
        s_abort = 0;
    }

    uint32_t n_perm;
    uint32_t* val = t.kcnt;
    if (MODE == RED_DEDUPE_KEYED) {
        if (e - a > kKeyCap) {  // the host only uses this mode when every partition fits (or is skipped above)
            if (tid == 0) atomicExch(&stats->overflow, 32ULL);
            return;
        }
        const uint32_t n = (uint32_t)(e - a);
        unsigned long long lo[kRedLoads], hi[kRedLoads];
#pragma unroll
        for (int k = 0; k < kRedLoads; k++) {
            const uint32_t j = k * kRedThreads + tid;
            const bool ok = j < n;
            lo[k] = ok ? in.lo[a + j] : kEmpty;
            hi[k] = WIDE ? (ok ? in.hi[a + j] : kEmpty) : 0ULL;
        }
#pragma unroll
        for (int k = 0; k < kRedLoads; k++) {
            const uint32_t j = k * kRedThreads + tid;
            if (j < n) {
                t.klo[j] = lo[k];
                if (WIDE) t.khi[j] = hi[k];
            }
        }
        __syncthreads();
        // One loop over probe STEPS, not over items: a lane that settles an item moves on to its next one at once, so a
        // warp runs for the longest lane total (about items x 1.5 steps) instead of the sum of per-item maxima.
        uint32_t uniq = 0;
        {
            uint32_t j = tid, s = 0, rep = kEmpty32, walked = 0;
            unsigned long long clo = 0, chi = 0;
            Key K{0, 0};
            bool fresh = true;
            while (j < n) {
                if (fresh) {
                    clo = t.klo[j];
                    chi = WIDE ? t.khi[j] : 0ULL;
                    if (!item_valid(clo, chi, WIDE)) {
                        j += kRedThreads;
                        continue;
                    }
                    K = drop_umi<WIDE>(clo, chi, umi_bits);
                    s = slot_hash<WIDE>(K.lo, K.hi) & (kTableSlots - 1);
                    rep = kEmpty32;  // first entry of this key met so far
                    walked = 0;
                    fresh = false;
                }
                uint32_t v = *reinterpret_cast<volatile uint32_t*>(&t.table[s]);
                if (v == kEmpty32) {
                    const uint32_t old = atomicCAS(&t.table[s], kEmpty32, j);
                    if (old == kEmpty32) {  // a record not seen before
                        atomicAdd(&kw[rep == kEmpty32 ? j : rep], 1u);
                        uniq++;
                        j += kRedThreads;
                        fresh = true;
                        continue;
                    }
                    v = old;
                }
                const unsigned long long vlo = t.klo[v], vhi = WIDE ? t.khi[v] : 0ULL;
                if (vlo == clo && (!WIDE || vhi == chi)) {  // a repeat (info.rs:780-791: only the first insert counts)
                    j += kRedThreads;
                    fresh = true;
                    continue;
                }
                if (rep == kEmpty32) {
                    bool same;
                    if (!WIDE) {
                        same = ((vlo ^ clo) >> umi_bits) == 0ULL;
                    } else {
                        const Key o = drop_umi<true>(vlo, vhi, umi_bits);
                        same = o.lo == K.lo && o.hi == K.hi;
                    }
                    if (same) rep = v;
                }
                s = (s + 1) & (kTableSlots - 1);
                if (++walked >= kChainLimit) {  // a hot key (or an unlucky cluster): hand the partition to the two-pass kernel
                    if (hot_list) {
                        s_abort = 1;
                        break;
                    }
                    walked = 0;
                }
                if ((walked & 7u) == 7u && *reinterpret_cast<volatile uint32_t*>(&s_abort)) break;  // another lane gave up
            }
        }
        __syncthreads();
        if (s_abort) {  // uniform: nothing was emitted, no statistic was touched
            if (tid == 0) hot_list[atomicAdd(hot_n, 1u)] = (uint32_t)p;
            return;
        }
        for (int o = 16; o; o >>= 1) uniq += __shfl_xor_sync(0xFFFFFFFFu, uniq, o);
        if (lane == 0) s_warp[wid] = uniq;
        __syncthreads();
        if (tid == 0) {
            unsigned long long u = 0;
            for (int i = 0; i < kRedThreads / 32; i++) u += s_warp[i];
            if (u) atomicAdd(&stats->unique, u);
        }
        __syncthreads();
        n_perm = n;
        val = kw;
    } else {
    if (e - a <= kKeyCap) {
        // ---- pass 1, staged: item j lives at key-store index j
        const uint32_t n = (uint32_t)(e - a);
        unsigned long long lo[kRedLoads], hi[kRedLoads];
        uint32_t w[kRedLoads];
#pragma unroll
        for (int k = 0; k < kRedLoads; k++) {
            const uint32_t j = k * kRedThreads + tid;
            const bool ok = j < n;
            lo[k] = ok ? in.lo[a + j] : kEmpty;
            hi[k] = WIDE ? (ok ? in.hi[a + j] : kEmpty) : 0ULL;
            w[k] = (MODE == RED_COUNT && in.w && ok) ? narrow_weight(in.w[a + j], stats) : 1u;
        }
#pragma unroll
        for (int k = 0; k < kRedLoads; k++) {
            const uint32_t j = k * kRedThreads + tid;
            if (j < n) {
                t.klo[j] = lo[k];
                if (WIDE) t.khi[j] = hi[k];
                t.kcnt[j] = 0u;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kRedLoads; k++) {
            const uint32_t j = k * kRedThreads + tid;
            if (j >= n || !item_valid(lo[k], hi[k], WIDE)) continue;
            const uint32_t at = smem_find_or_claim_staged<WIDE>(t, lo[k], hi[k], slot_hash<WIDE>(lo[k], hi[k]), j);
            if (MODE == RED_COUNT) add_weight(&t.kcnt[at], w[k], stats);
            else if (at == j) t.kcnt[j] = 1u;  // first of its kind; repeats keep 0
        }
        __syncthreads();
        n_perm = n;
    } else {
        // ---- pass 1, streaming: more items than the key store holds (few distinct keys, or overflow)
        __syncthreads();
        uint32_t reserve = kEmpty32;
        for (unsigned long long b = a; b < e; b += (unsigned long long)kRedLoads * kRedThreads) {
            unsigned long long lo[kRedLoads], hi[kRedLoads];
            uint32_t w[kRedLoads];
#pragma unroll
            for (int k = 0; k < kRedLoads; k++) {
                const unsigned long long i = b + (unsigned long long)k * kRedThreads + tid;
                const bool ok = i < e;
                lo[k] = ok ? in.lo[i] : kEmpty;
                hi[k] = WIDE ? (ok ? in.hi[i] : kEmpty) : 0ULL;
                w[k] = (MODE == RED_COUNT && in.w && ok) ? narrow_weight(in.w[i], stats) : 1u;
            }
#pragma unroll
            for (int k = 0; k < kRedLoads; k++) {
                if (!item_valid(lo[k], hi[k], WIDE)) continue;
                const uint32_t at = smem_find_or_claim<WIDE>(t, lo[k], hi[k], slot_hash<WIDE>(lo[k], hi[k]), reserve);
                if (at != kEmpty32) {
                    if (MODE == RED_COUNT) add_weight(&t.kcnt[at], w[k], stats);
                    else atomicAdd(&t.kcnt[at], 1u);
                }
            }
        }
        __syncthreads();
        if (s_nperm > kKeyCap) {  // uniform: more distinct keys than the key store holds
            if (tid == 0) atomicExch(&stats->overflow, 1ULL);
            return;
        }
        n_perm = s_nperm;
    }

    if (MODE == RED_DEDUPE) {
        // ---- pass 2: the distinct records, keyed by the record without its random barcode.  The key store stays as
        // it is; the table is rebuilt over indices of entries, an entry's key being (record >> umi_bits).
        clear_u32(t.table, kTableSlots, kEmpty32);
        __syncthreads();
        uint32_t uniq = 0;
        for (uint32_t i = tid; i < n_perm; i += kRedThreads) {
            if (t.kcnt[i] == 0u) continue;
            uniq++;
            const Key k = drop_umi<WIDE>(t.klo[i], WIDE ? t.khi[i] : 0ULL, umi_bits);
            uint32_t s = slot_hash<WIDE>(k.lo, k.hi) & (kTableSlots - 1);
            for (;;) {
                uint32_t v = *reinterpret_cast<volatile uint32_t*>(&t.table[s]);
                if (v == kEmpty32) {
                    const uint32_t old = atomicCAS(&t.table[s], kEmpty32, i);
                    if (old == kEmpty32) {
                        atomicAdd(&kw[i], 1u);
                        break;
                    }
                    v = old;
                }
                const Key o = drop_umi<WIDE>(t.klo[v], WIDE ? t.khi[v] : 0ULL, umi_bits);
                if (o.lo == k.lo && o.hi == k.hi) {
                    atomicAdd(&kw[v], 1u);
                    break;
                }
                s = (s + 1) & (kTableSlots - 1);
            }
        }
        for (int o = 16; o; o >>= 1) uniq += __shfl_xor_sync(0xFFFFFFFFu, uniq, o);
        if (lane == 0) s_warp[wid] = uniq;
        __syncthreads();
        if (tid == 0) {
            unsigned long long u = 0;
            for (int i = 0; i < kRedThreads / 32; i++) u += s_warp[i];
            if (u) atomicAdd(&stats->unique, u);
        }
        __syncthreads();
        val = kw;
    }

    }  // two-pass modes

    // ---- emit: entries with a non-zero value, one global reservation per CTA, each warp writes a contiguous run
    uint32_t mine = 0;
    for (uint32_t i = tid; i < n_perm; i += kRedThreads) mine += val[i] != 0u;
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    if (lane == 0) s_warp[wid] = mine;
    __syncthreads();
    if (tid == 0) {
        uint32_t total = 0;
        for (int i = 0; i < kRedThreads / 32; i++) {
            const uint32_t c = s_warp[i];
            s_warp[i] = total;
            total += c;
        }
        unsigned long long base = total ? atomicAdd(&stats->n_out, (unsigned long long)total) : 0ULL;
        if (base + total > out_cap) {  // cannot happen: the host sizes the output for one item per input item
            atomicExch(&stats->overflow, 2ULL);
            base = ~0ULL;
        }
        s_base = base;
    }
    __syncthreads();
    if (s_base == ~0ULL) return;
    // a warp's entries are those with (i / 32) % n_warps == wid, in increasing i: the same order the count used
    unsigned long long at = s_base + s_warp[wid];
    for (uint32_t i0 = wid * 32; i0 < n_perm; i0 += kRedThreads) {  // warp-uniform trip count
        const uint32_t i = i0 + lane;
        const uint32_t c = i < n_perm ? val[i] : 0u;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, c != 0u);
        if (c != 0u) {
            Key k{t.klo[i], WIDE ? t.khi[i] : 0ULL};
            if (MODE != RED_COUNT) k = drop_umi<WIDE>(k.lo, k.hi, umi_bits);
            const unsigned long long pos = at + __popc(bal & ((1u << lane) - 1u));
            out.lo[pos] = k.lo;
            if (out.hi) out.hi[pos] = k.hi;
            out.w[pos] = c;
        }
        at += __popc(bal);
    }
}

// One CTA per partition; or, with `list`, persistent CTAs over the partitions listed on the device (the hot partitions the
// keyed pass gave up on: their number is only known there).
template <bool WIDE, int MODE>
__global__ void __launch_bounds__(kRedThreads, WIDE ? 3 : BC_REDUCE_BLOCKS)
    k_reduce(const ItemView in, const uint32_t* __restrict__ starts, const unsigned long long n_items, const uint32_t chunk,
             const uint32_t umi_bits, const ItemView out, const unsigned long long out_cap, FlushStats* stats,
             const uint32_t skip_over, const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_n, uint32_t* hot_list,
             uint32_t* hot_n) {
    if (!list) {
        reduce_one<WIDE, MODE>(blockIdx.x, in, starts, n_items, chunk, umi_bits, out, out_cap, stats, skip_over, hot_list, hot_n);
        return;
    }
    const uint32_t n = *list_n;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        reduce_one<WIDE, MODE>(list[i], in, starts, n_items, chunk, umi_bits, out, out_cap, stats, skip_over, nullptr, nullptr);
        __syncthreads();  // the next partition reuses the shared memory
    }
}

template <bool WIDE, int MODE>
cudaError_t launch_reduce_t(const ItemView& in, const uint32_t* starts, unsigned long long n_items, unsigned long long n_ranges,
                            uint32_t chunk, uint32_t umi_bits, const ItemView& out, unsigned long long out_cap, FlushStats* stats,
                            uint32_t skip_over, const HotList& hot, cudaStream_t stream) {
    const size_t smem = (size_t)kKeyCap * 8 * (WIDE ? 2 : 1) + (size_t)kTableSlots * 4 + (size_t)kKeyCap * 4 * (MODE == RED_DEDUPE ? 2 : 1);
    cudaError_t e = cudaFuncSetAttribute(k_reduce<WIDE, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_reduce<WIDE, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (hot.consume) {  // persistent CTAs over the listed partitions
        const unsigned grid = (unsigned)std::min<unsigned long long>(n_ranges, 148ULL * (WIDE ? 3 : 4));
        k_reduce<WIDE, MODE><<<grid, kRedThreads, smem, stream>>>(in, starts, n_items, chunk, umi_bits, out, out_cap, stats, skip_over, hot.list,
                                                                  hot.n, nullptr, nullptr);
    } else {
        k_reduce<WIDE, MODE><<<(unsigned)n_ranges, kRedThreads, smem, stream>>>(in, starts, n_items, chunk, umi_bits, out, out_cap, stats,
                                                                                skip_over, nullptr, nullptr, hot.list, hot.n);
    }
    return cudaGetLastError();
}

unsigned stream_grid(unsigned long long n, unsigned block) {
    unsigned long long g = (n + block - 1) / block;
    const unsigned long long cap = 148ULL * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// global-table path over the record buffer: what k_insert does for routed records, minus the outcome counters
template <bool WIDE>
__global__ void k_insert_items(const Tables tables, const ItemView in, const unsigned long long n, FlushStats* stats) {
    unsigned long long valid = 0, fresh = 0, pairs = 0, uniq = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long lo = in.lo[i], hi = WIDE ? in.hi[i] : 0ULL;
        if (!item_valid(lo, hi, WIDE)) continue;
        valid++;
        bool new_key = false, new_pair = false;
        if (count_read(tables, Key{lo, hi}, &new_key, &new_pair)) uniq++;
        fresh += new_key;
        pairs += new_pair;
    }
    for (int o = 16; o; o >>= 1) {
        valid += __shfl_xor_sync(0xFFFFFFFFu, valid, o);
        fresh += __shfl_xor_sync(0xFFFFFFFFu, fresh, o);
        pairs += __shfl_xor_sync(0xFFFFFFFFu, pairs, o);
        uniq += __shfl_xor_sync(0xFFFFFFFFu, uniq, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (valid) atomicAdd(&stats->valid, valid);
        if (uniq) atomicAdd(&stats->unique, uniq);
        if (fresh && tables.map.n_entries) atomicAdd(tables.map.n_entries, fresh);
        if (pairs && tables.set.n_entries) atomicAdd(tables.set.n_entries, pairs);
    }
}

}  // namespace

uint32_t reduce_fill(bool) { return kKeyCap * 3 / 4; }  // mean 1536, sigma 39: the key store is 13 sigma away
uint32_t reduce_chunk(bool) { return kKeyCap; }
uint32_t reduce_capacity(bool) { return kKeyCap; }
uint32_t split_max_bits() { return kSplitMaxBits; }

cudaError_t launch_seg_scan(const uint32_t* hist, uint32_t n_seg, uint32_t bins_per_seg, const uint32_t* seg_base, uint32_t* starts,
                            uint32_t* cursor, FlushStats* stats, uint32_t limit, cudaStream_t stream) {
    k_seg_scan<<<n_seg, 512, 0, stream>>>(hist, bins_per_seg, seg_base, starts, cursor, stats, limit);
    return cudaGetLastError();
}

template <bool WIDE, bool WEIGHTED, int IPT>
static cudaError_t launch_scatter_t(unsigned grid, const ItemView& in, const ItemView& out, const uint32_t* seg_starts,
                                    uint32_t workers, unsigned long long n_total, const SplitLevel& lv, uint32_t* bins, cudaStream_t stream) {
    constexpr size_t T = (size_t)kScatterThreads * IPT;
    const size_t smem = T * 8 * (1 + (WIDE ? 1 : 0) + (WEIGHTED ? 1 : 0)) + T * 4;
    cudaError_t e = cudaFuncSetAttribute(k_split_scatter<WIDE, WEIGHTED, IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_split_scatter<WIDE, WEIGHTED, IPT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    k_split_scatter<WIDE, WEIGHTED, IPT><<<grid, kScatterThreads, smem, stream>>>(in, out, seg_starts, workers, n_total, lv, bins);
    return cudaGetLastError();
}

cudaError_t launch_split(bool scatter, bool wide, const ItemView& in, const ItemView& out, const uint32_t* seg_starts, uint32_t n_seg,
                         unsigned long long n_total, const SplitLevel& lv, uint32_t* bins, FlushStats* stats, bool count_valid,
                         cudaStream_t stream) {
    if (n_total == 0) return cudaSuccess;
    if (lv.F > (1u << kSplitMaxBits) || lv.F == 0) return cudaErrorInvalidValue;
    const bool weighted = in.w != nullptr;
    const unsigned long long tile = scatter ? (unsigned long long)kScatterThreads * ((wide || weighted) ? 8 : BC_SCATTER_IPT) : (unsigned long long)kSplitThreads * 16;
    // single segment: one worker per tile, at most a few waves; many segments: a few workers each
    unsigned long long workers;
    if (!seg_starts) {
        workers = (n_total + tile - 1) / tile;
        if (workers > 148ULL * 16) workers = 148ULL * 16;
        n_seg = 1;
    } else {
        workers = (148ULL * 16 + n_seg - 1) / n_seg;
        const unsigned long long per_seg = (n_total / n_seg + tile - 1) / tile;
        if (workers > per_seg) workers = per_seg;
    }
    if (workers < 1) workers = 1;
    const unsigned grid = (unsigned)(workers * n_seg);
    const uint32_t w = (uint32_t)workers;
    if (!scatter) {
        if (wide) k_split_hist<true><<<grid, kSplitThreads, 0, stream>>>(in, seg_starts, w, n_total, lv, bins, stats, count_valid);
        else k_split_hist<false><<<grid, kSplitThreads, 0, stream>>>(in, seg_starts, w, n_total, lv, bins, stats, count_valid);
        return cudaGetLastError();
    }
    if (wide) return weighted ? launch_scatter_t<true, true, 8>(grid, in, out, seg_starts, w, n_total, lv, bins, stream)
                              : launch_scatter_t<true, false, 8>(grid, in, out, seg_starts, w, n_total, lv, bins, stream);
    return weighted ? launch_scatter_t<false, true, 8>(grid, in, out, seg_starts, w, n_total, lv, bins, stream)
                    : launch_scatter_t<false, false, BC_SCATTER_IPT>(grid, in, out, seg_starts, w, n_total, lv, bins, stream);
}

// Exchange step of a multi-GPU job: this rank's records -> their owners' receive buffers.  owner = digit of `lv`
// (F = number of ranks, the key without its random barcode hashed with the level's salt); cursors[b] holds, on entry, the
// position in owner b's buffer where this rank's run starts.
cudaError_t launch_owner_scatter(bool wide, const ItemView& in, const PeerOut& peers, unsigned long long n_total, const SplitLevel& lv,
                                 uint32_t* cursors, cudaStream_t stream) {
    if (n_total == 0) return cudaSuccess;
    if (lv.F == 0 || lv.F > (uint32_t)kMaxRanks) return cudaErrorInvalidValue;
    const unsigned long long tile = (unsigned long long)kScatterThreads * (wide ? 8 : 16);
    unsigned long long workers = (n_total + tile - 1) / tile;
    if (workers > 148ULL * 16) workers = 148ULL * 16;
    const size_t smem = (size_t)tile * 8 * (wide ? 2 : 1) + (size_t)tile * 4;
    cudaError_t e;
    if (wide) {
        e = cudaFuncSetAttribute(k_owner_scatter<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_owner_scatter<true, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        k_owner_scatter<true, 8><<<(unsigned)workers, kScatterThreads, smem, stream>>>(in, peers, (uint32_t)workers, n_total, lv, cursors);
    } else {
        e = cudaFuncSetAttribute(k_owner_scatter<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_owner_scatter<false, 16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        k_owner_scatter<false, 16><<<(unsigned)workers, kScatterThreads, smem, stream>>>(in, peers, (uint32_t)workers, n_total, lv, cursors);
    }
    return cudaGetLastError();
}

cudaError_t launch_reduce(int mode, bool wide, const ItemView& in, const uint32_t* starts, unsigned long long n_items,
                          unsigned long long n_ranges, uint32_t chunk, uint32_t umi_bits, const ItemView& out,
                          unsigned long long out_cap, FlushStats* stats, uint32_t skip_over, const HotList& hot, cudaStream_t stream) {
    if (n_ranges == 0) return cudaSuccess;
    if (n_ranges > 0x7FFFFFFFULL) return cudaErrorInvalidValue;
    if (mode == RED_DEDUPE_KEYED)
        return wide ? launch_reduce_t<true, RED_DEDUPE_KEYED>(in, starts, n_items, n_ranges, chunk, umi_bits, out, out_cap, stats, skip_over, hot, stream)
                    : launch_reduce_t<false, RED_DEDUPE_KEYED>(in, starts, n_items, n_ranges, chunk, umi_bits, out, out_cap, stats, skip_over, hot, stream);
    if (mode == RED_DEDUPE)
        return wide ? launch_reduce_t<true, RED_DEDUPE>(in, starts, n_items, n_ranges, chunk, umi_bits, out, out_cap, stats, skip_over, hot, stream)
                    : launch_reduce_t<false, RED_DEDUPE>(in, starts, n_items, n_ranges, chunk, umi_bits, out, out_cap, stats, skip_over, hot, stream);
    return wide ? launch_reduce_t<true, RED_COUNT>(in, starts, n_items, n_ranges, chunk, umi_bits, out, out_cap, stats, skip_over, hot, stream)
                : launch_reduce_t<false, RED_COUNT>(in, starts, n_items, n_ranges, chunk, umi_bits, out, out_cap, stats, skip_over, hot, stream);
}

// partitions larger than `limit` copied to a contiguous buffer (their total is known from the scan).  Two steps, because a
// hot key's partition can hold millions of records: first every big partition reserves its place (one thread per
// partition, list[i] = {partition, place}), then all CTAs copy the listed partitions chunk by chunk — a single warp per
// partition took 2.5 ms for 21 MB of Zipf-distributed lineage barcodes.
__global__ void k_list_big(const uint32_t* __restrict__ starts, const unsigned long long n_parts, const uint32_t limit, uint4* __restrict__ list,
                           uint32_t* __restrict__ list_n, FlushStats* stats) {
    for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n_parts; p += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t a = starts[p], e = starts[p + 1];
        if (e - a <= limit) continue;
        const unsigned long long base = atomicAdd(&stats->n_out, (unsigned long long)(e - a));
        list[atomicAdd(list_n, 1u)] = make_uint4(a, e - a, (uint32_t)base, (uint32_t)(base >> 32));
    }
}
__global__ void __launch_bounds__(256) k_copy_big(const ItemView in, const uint4* __restrict__ list, const uint32_t* __restrict__ list_n, const ItemView out) {
    constexpr uint32_t kChunk = 4096;
    const uint32_t n = *list_n;
    for (uint32_t i = 0; i < n; i++) {
        const uint4 ent = list[i];
        const unsigned long long base = ((unsigned long long)ent.w << 32) | ent.z;
        // chunk c of entry i belongs to CTA (i + c) mod grid: thousands of partitions a little over the limit spread over all
        // CTAs just like the chunks of one huge partition
        const uint32_t first = (blockIdx.x + gridDim.x - i % gridDim.x) % gridDim.x;
        for (unsigned long long c0 = (unsigned long long)first * kChunk; c0 < ent.y; c0 += (unsigned long long)gridDim.x * kChunk) {
            const uint32_t c1 = (uint32_t)min((unsigned long long)ent.y, c0 + kChunk);
            for (uint32_t j = (uint32_t)c0 + threadIdx.x; j < c1; j += 256) {
                out.lo[base + j] = in.lo[ent.x + j];
                if (in.hi) out.hi[base + j] = in.hi[ent.x + j];
            }
        }
    }
}

cudaError_t launch_gather_big(const ItemView& in, const uint32_t* starts, unsigned long long n_parts, uint32_t limit, const ItemView& out,
                              FlushStats* stats, uint32_t* scratch /* 4 x n_parts u32, 16-byte aligned */, uint32_t* scratch_n, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(scratch_n, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    k_list_big<<<stream_grid(n_parts, 256), 256, 0, stream>>>(starts, n_parts, limit, reinterpret_cast<uint4*>(scratch), scratch_n, stats);
    k_copy_big<<<148 * 8, 256, 0, stream>>>(in, reinterpret_cast<const uint4*>(scratch), scratch_n, out);
    return cudaGetLastError();
}

cudaError_t launch_insert_items(const Tables& tables, const ItemView& in, bool wide, unsigned long long n, FlushStats* stats,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (wide) k_insert_items<true><<<stream_grid(n, 256), 256, 0, stream>>>(tables, in, n, stats);
    else k_insert_items<false><<<stream_grid(n, 256), 256, 0, stream>>>(tables, in, n, stats);
    return cudaGetLastError();
}

}  // namespace bc
