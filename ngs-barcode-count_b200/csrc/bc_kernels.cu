// bc_kernels.cu — hand-written sm_100a kernels of the decode-and-count path.
//
//  k_decode   fused K1 (locate: parse.rs:89-96,151-163,287-313) + K2 (quality parse.rs:331-375, correction
//             parse.rs:439-524,553-593) + K3 (count info.rs:735-808, counters info.rs:60-127)
//  k_build_table / k_insert / k_group / k_compact / k_marginal / k_rehash: table set-up, finish and enrichment
//             (output.rs:265-270, info.rs:840-904)
//
// Integer bit-plane string matching: XOR/LOP3 + POPC on packed planes; no tensor-core work exists on this path.
#include "../../include/bc_b200.h"
#include "bc_kernels.h"
#include "bc_decode.cuh"

namespace bc {

// the generic kernel: run constants as a kernel parameter (any scheme); bc_jit.cu builds the per-run specialisation
template <int TW>
__global__ void __launch_bounds__(kTile) k_decode(const __grid_constant__ DevCfg cfg, const BatchView batch, const DevAux aux,
                                                  const Tables tables, unsigned long long* __restrict__ counters, const DecodeOut out,
                                                  const RecOut rec, const Deferred deferred, const int flags) {
    decode_body<TW>(cfg, batch, aux, tables, counters, out, rec, deferred, flags);
}

__global__ void k_fold_counters(unsigned long long* __restrict__ stripes, unsigned long long* __restrict__ counters) {
    const uint32_t t = threadIdx.x;
    if (t >= BC_N_COUNTERS + 2) return;
    unsigned long long sum = 0;
    for (uint32_t s = 0; s < kCounterStripes; s++) {
        sum += stripes[s * kCounterStride + t];
        stripes[s * kCounterStride + t] = 0;
    }
    if (sum) counters[t] += sum;
}

cudaError_t launch_fold_counters(unsigned long long* stripes, unsigned long long* counters, cudaStream_t stream) {
    k_fold_counters<<<1, 32, 0, stream>>>(stripes, counters);
    return cudaGetLastError();
}

// ---- k_resolve: one WARP per deferred read.  Redoes the barcode step of that read (parse.rs:439-524) with the
// searches done cooperatively: lanes share the candidates of the block index (or, when a slot has none, the whole
// reference set) and reduce to (minimum distance, how many reach it, which one) — fix_error's unique-minimum rule.
__device__ __forceinline__ uint32_t warp_best(Best b, uint32_t max_err) {
    uint32_t dmin = b.d;
#pragma unroll
    for (int o = 16; o; o >>= 1) dmin = min(dmin, __shfl_xor_sync(0xFFFFFFFFu, dmin, o));
    uint32_t cnt = b.d == dmin ? b.cnt : 0u;
    uint32_t arg = b.d == dmin && b.cnt ? b.arg : kFail;
    uint32_t exact = b.exact;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        arg = min(arg, __shfl_xor_sync(0xFFFFFFFFu, arg, o));
        exact = min(exact, __shfl_xor_sync(0xFFFFFFFFu, exact, o));
    }
    if (exact != kFail) return exact;
    return (cnt == 1 && dmin <= max_err) ? arg : kFail;
}

constexpr uint32_t kMaxQueryN = 4;  // N per query the block index expands (4^N completions per block at most)

__device__ __forceinline__ uint32_t warp_scan_all(const DevAux& aux, const DevSlot& S, const SlotBits& q, int lane) {
    Best b{S.max_err + 1u, 0, kFail, kFail};
    const uint32_t lm = lenmask(S.len);
    const uint4* refs = aux.refs + S.ref_off;
    uint32_t i = lane;
    for (; i + 96 < S.n_ref; i += 128) {  // four independent loads in flight per lane
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; u++) r[u] = __ldg(&refs[i + 32 * u]);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (ref_same(r[u], q.lo, q.hi, q.nm, S.len)) b.exact = i + 32 * u;
            best_add(b, ref_dist(r[u], q.lo, q.hi, q.nm, S.len, lm), i + 32 * u);
        }
    }
    for (; i < S.n_ref; i += 32) {
        const uint4 r = __ldg(&refs[i]);
        if (ref_same(r, q.lo, q.hi, q.nm, S.len)) b.exact = i;
        best_add(b, ref_dist(r, q.lo, q.hi, q.nm, S.len, lm), i);
    }
    return warp_best(b, S.max_err);
}

// Block index, one level: every reference within D.cap of the query (N positions of the query never count) agrees
// with it on all non-N bases of at least one of the D.cap+1 blocks, hence on that block's key bases; it is then
// found in the bucket of one completion of the key's N positions.  A reference is counted at the FIRST block it
// agrees on.  The result is final when the smallest distance found is <= D.cap (the level is complete up to there).
__device__ __forceinline__ void warp_scan_level(const DevAux& aux, const DevDeep& D, const SlotBits& q, uint32_t lm, int lane,
                                                Best& b) {
    for (uint32_t p = 0; p < D.n_blocks; p++) {
        const uint32_t kpos = D.key_pos[p], kl = D.key_len[p], km = lenmask(kl);
        const uint32_t qlo = (q.lo >> kpos) & km, qhi = (q.hi >> kpos) & km, qn = (q.nm >> kpos) & km;
        const uint32_t t = __popc(qn);  // <= kMaxQueryN (checked by the caller)
        const uint32_t* start = aux.csr + D.start_off[p];
        const uint4* bucket_refs = aux.bref + D.ids_off[p];  // {lo, hi, id, 0} in bucket order: no second indirection
        for (uint32_t comp = 0; comp < (1u << (2 * t)); comp++) {
            uint32_t vlo = qlo, vhi = qhi;
            for (uint32_t m = qn, c = comp; m; m &= m - 1, c >>= 2) {  // deposit the completion at the N positions
                const uint32_t pos = (uint32_t)__ffs(m) - 1u;
                vlo |= (c & 1u) << pos;
                vhi |= ((c >> 1) & 1u) << pos;
            }
            const uint32_t bucket = vlo | (vhi << kl);
            const uint32_t a = __ldg(&start[bucket]), e = __ldg(&start[bucket + 1]);
            for (uint32_t j = a + lane; j < e; j += 32) {
                const uint4 r = __ldg(&bucket_refs[j]);
                const uint32_t diff = ((q.lo ^ r.x) | (q.hi ^ r.y)) & ~q.nm & lm;
                bool earlier = false;
                for (uint32_t pp = 0; pp < p; pp++)
                    earlier |= (diff & (lenmask(D.key_len[pp]) << D.key_pos[pp])) == 0;
                if (!earlier) best_add(b, __popc(diff), r.z);
            }
        }
    }
}

__device__ __forceinline__ uint32_t warp_scan_blocks(const DevAux& aux, const DevSlot& S, const SlotBits& q, int lane) {
    const uint32_t lm = lenmask(S.len);
    for (uint32_t lv = 0; lv < S.n_levels; lv++) {
        const DevDeep& D = aux.deep[S.deep_off + lv];
        Best b{S.max_err + 1u, 0, kFail, kFail};
        warp_scan_level(aux, D, q, lm, lane, b);
        uint32_t dmin = b.d;
#pragma unroll
        for (int o = 16; o; o >>= 1) dmin = min(dmin, __shfl_xor_sync(0xFFFFFFFFu, dmin, o));
        // complete up to D.cap: a minimum within it is the true minimum with its true multiplicity; the last level
        // (cap == max_err) also settles "nothing within the cap"
        if (dmin <= D.cap || lv + 1 == S.n_levels) return warp_best(b, S.max_err);
    }
    return kFail;
}

__global__ void __launch_bounds__(128) k_resolve(const __grid_constant__ DevCfg cfg, const BatchView batch, const DevAux aux,
                                                 const Tables tables, unsigned long long* __restrict__ counters,
                                                 const DecodeOut out, const RecOut rec, const Deferred deferred, const int flags) {
    const int lane = threadIdx.x & 31;
    const uint32_t n = *deferred.count;
    if (blockIdx.x == 0 && threadIdx.x == 0) *deferred.next_count = 0;  // nobody touches it before the next batch's k_decode
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t W = batch.W;
    unsigned long long c_matched = 0, c_dup = 0, c_sample = 0, c_counted = 0, c_new = 0, c_pair = 0;  // lane 0 only
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += n_warps) {
        const uint2 item = deferred.items[w];
        const unsigned long long ri = item.x;
        const int off = (int)(item.y & 0xFFFFu);
        const uint32_t* lo = batch.planes + ri * batch.plane_stride;
        const uint32_t* hi = lo + W;
        const uint32_t* nm = hi + W;
        int status = BC_ST_MATCHED;
        Key key{0, 0};
        for (uint32_t oi = 0; oi < cfg.n_slots; oi++) {
            const uint32_t si = cfg.order[oi];
            const DevSlot& S = cfg.slots[si];
            const SlotBits b = slot_bits<true>(lo, hi, nm, W, off + S.offset, S.len);  // lane-uniform
            if (S.mode == MODE_RAW) {
                key_raw(key, S, b, cfg.wide);
                continue;
            }
            uint32_t idx = kFail;
            bool search = true;
            if (b.nm == 0 && S.mode == MODE_TABLE) {
                idx = table_pick(__ldg(&aux.tables[S.aux_off + (b.lo | (b.hi << S.len))]), S.max_err);
                search = false;  // the table already holds the result of the full search
            } else if (b.nm == 0 && S.mode == MODE_HASH) {
                idx = hash_exact(aux, S, b.lo, b.hi);
                search = idx == kFail;
            }
            if (search) {
                if (S.n_levels && __popc(b.nm) <= kMaxQueryN) idx = warp_scan_blocks(aux, S, b, lane);
                else idx = warp_scan_all(aux, S, b, lane);
            }
            if (lane == 0 && out.slot_index) out.slot_index[ri * cfg.n_slots + si] = (int32_t)idx;
            if (idx == kFail) {
                status = S.kind == 'S' ? BC_ST_SAMPLE : BC_ST_COUNTED;
                break;
            }
            key_or(key, idx, S.key_shift);
        }
        if (lane == 0) {
            bool new_key = false, new_pair = false;
            if (status == BC_ST_MATCHED) status = count_or_append(tables, rec, ri, flags, key, &new_key, &new_pair);
            c_matched += status == BC_ST_MATCHED;
            c_dup += status == BC_ST_DUPLICATE;
            c_sample += status == BC_ST_SAMPLE;
            c_counted += status == BC_ST_COUNTED;
            c_new += new_key;
            c_pair += new_pair;
            if (flags & F_EMIT) {
                if (out.status) out.status[ri] = (uint8_t)status;
                if (out.offset) out.offset[ri] = (int16_t)off;
                if (out.repaired) out.repaired[ri] = (item.y >> 16) & 1u;
                if (out.key_lo) out.key_lo[ri] = key.lo;
                if (out.key_hi) out.key_hi[ri] = key.hi;
            }
        }
    }
    if (lane == 0 && counters) {
        if (c_matched) atomicAdd(&counters[BC_CNT_MATCHED], c_matched);
        if (c_dup) atomicAdd(&counters[BC_CNT_DUPLICATES], c_dup);
        if (c_sample) atomicAdd(&counters[BC_CNT_SAMPLE], c_sample);
        if (c_counted) atomicAdd(&counters[BC_CNT_COUNTED], c_counted);
        if (c_new && tables.map.n_entries) atomicAdd(tables.map.n_entries, c_new);
        if (c_pair && tables.set.n_entries) atomicAdd(tables.set.n_entries, c_pair);
    }
}

size_t decode_smem_bytes(const BatchView& b) { return (size_t)kTile * (b.plane_stride * 4u + 2u + 4u * b.rep_chunks); }

template <int TW>
static cudaError_t launch_decode_tw(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const Tables& tables,
                                    unsigned long long* counters, const DecodeOut& out, const RecOut& rec, const Deferred& deferred,
                                    int flags, cudaStream_t stream) {
    const size_t smem = decode_smem_bytes(batch);
    if (smem > 48 * 1024) {  // per device and per instantiation: set on every such launch (a cheap driver call, reads above ~380 nt only)
        cudaError_t e = cudaFuncSetAttribute(k_decode<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (batch.n_reads + kTile - 1) / kTile;
    k_decode<TW><<<grid, kTile, smem, stream>>>(cfg, batch, aux, tables, counters, out, rec, deferred, flags);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const Tables& tables,
                          unsigned long long* counters, const DecodeOut& out, const RecOut& rec, const Deferred& deferred, int flags,
                          cudaStream_t stream) {
    if (batch.n_reads == 0) return cudaSuccess;
    switch (cfg.TW) {
        case 1: return launch_decode_tw<1>(cfg, batch, aux, tables, counters, out, rec, deferred, flags, stream);
        case 2: return launch_decode_tw<2>(cfg, batch, aux, tables, counters, out, rec, deferred, flags, stream);
        case 3: return launch_decode_tw<3>(cfg, batch, aux, tables, counters, out, rec, deferred, flags, stream);
        case 4: return launch_decode_tw<4>(cfg, batch, aux, tables, counters, out, rec, deferred, flags, stream);
        case 5: return launch_decode_tw<5>(cfg, batch, aux, tables, counters, out, rec, deferred, flags, stream);
        case 6: return launch_decode_tw<6>(cfg, batch, aux, tables, counters, out, rec, deferred, flags, stream);
        case 7: return launch_decode_tw<7>(cfg, batch, aux, tables, counters, out, rec, deferred, flags, stream);
        case 8: return launch_decode_tw<8>(cfg, batch, aux, tables, counters, out, rec, deferred, flags, stream);
        default: return cudaErrorInvalidValue;
    }
}

// the per-run specialisation of the same kernel (bc_jit.cu), compiled for this batch geometry
cudaError_t launch_decode_jit(const void* kernel, const BatchView& batch, const DevAux& aux, const Tables& tables, unsigned long long* counters,
                              const DecodeOut& out, const RecOut& rec, const Deferred& deferred, int flags, cudaStream_t stream) {
    if (batch.n_reads == 0) return cudaSuccess;
    const unsigned grid = (batch.n_reads + kTile - 1) / kTile;
    void* args[] = {(void*)&batch, (void*)&aux, (void*)&tables, (void*)&counters, (void*)&out, (void*)&rec, (void*)&deferred, (void*)&flags};
    return cudaLaunchKernel(kernel, dim3(grid), dim3(kTile), args, decode_smem_bytes(batch), stream);
}

// deferred reads of the batch just decoded (their number is read on the device: no host round trip)
cudaError_t launch_resolve(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const Tables& tables,
                           unsigned long long* counters, const DecodeOut& out, const RecOut& rec, const Deferred& deferred, int flags,
                           cudaStream_t stream) {
    if (batch.n_reads == 0) return cudaSuccess;
    unsigned long long warps = batch.n_reads;  // at most one warp per read of the batch
    unsigned grid = (unsigned)((warps + 3) / 4);
    const unsigned cap = 148u * 16u;  // persistent warps stride over the list
    if (grid > cap) grid = cap;
    k_resolve<<<grid, 128, 0, stream>>>(cfg, batch, aux, tables, counters, out, rec, deferred, flags);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// MODE_TABLE: result of the correction for every N-free barcode value, so the hot kernel does one lookup.
__global__ void k_build_table(const DevSlot slot, const uint4* __restrict__ refs, uint32_t* __restrict__ table) {
    const uint32_t n = 1u << (2 * slot.len);
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t m = lenmask(slot.len);
    const uint32_t blo = v & m, bhi = (v >> slot.len) & m;
    Best b{256u, 0, kFail, kFail};
    for (uint32_t i = 0; i < slot.n_ref; i++) {
        const uint4 r = __ldg(&refs[slot.ref_off + i]);
        if (ref_same(r, blo, bhi, 0u, slot.len)) b.exact = i;
        best_add(b, ref_dist(r, blo, bhi, 0u, slot.len, m), i);
    }
    uint32_t e;
    if (b.exact != kFail) e = b.exact;                                               // distance 0, no tie
    else if (b.cnt == 0) e = 0xFFFFu | (0xFFu << 16);                                // empty set
    else e = (b.arg & 0xFFFFu) | (min(b.d, 255u) << 16) | (b.cnt > 1 ? 0x1000000u : 0u);
    table[v] = e;
}

cudaError_t launch_build_table(const DevSlot& slot, const DevAux& aux, uint32_t* table, cudaStream_t stream) {
    const uint32_t n = 1u << (2 * slot.len);
    k_build_table<<<(n + 255) / 256, 256, 0, stream>>>(slot, aux.refs, table);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// (key, count) rows merged from other ranks' tables (keys without the random barcode)
__global__ void k_insert(const Tables tables, const unsigned long long* __restrict__ key_lo,
                         const unsigned long long* __restrict__ key_hi, const unsigned long long* __restrict__ counts,
                         const unsigned long long n) {
    unsigned long long fresh = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        bool new_key = false;
        Key k{key_lo[i], key_hi ? key_hi[i] : 0ULL};
        map_add(tables.map, k, counts[i], &new_key);
        fresh += new_key;
    }
    for (int o = 16; o; o >>= 1) fresh += __shfl_xor_sync(0xFFFFFFFFu, fresh, o);
    if ((threadIdx.x & 31) == 0 && fresh && tables.map.n_entries) atomicAdd(tables.map.n_entries, fresh);
}

static unsigned grid_for(unsigned long long n, unsigned block) {
    unsigned long long g = (n + block - 1) / block;
    const unsigned long long cap = 148ULL * 16;  // grid-stride: a few waves over the 148 SMs
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

cudaError_t launch_insert(const Tables& tables, const unsigned long long* key_lo, const unsigned long long* key_hi,
                          const unsigned long long* counts, unsigned long long n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_insert<<<grid_for(n, 256), 256, 0, stream>>>(tables, key_lo, key_hi, counts, n);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool table_entry(const DevTable& t, unsigned long long i, Key* k, unsigned long long* c) {
    const unsigned long long* slot = t.data + i * table_stride(t.kind, t.wide);
    if (t.kind == 0) {
        k->lo = i;
        k->hi = 0;
        *c = slot[0];
        return *c != 0;
    }
    if (t.wide) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(slot);
        k->lo = v.x;
        k->hi = v.y;
        *c = t.kind == 1 ? slot[2] : 1ULL;
        return !(v.x == kEmpty && v.y == kEmpty);
    }
    if (t.kind == 1) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(slot);
        k->lo = v.x;
        k->hi = 0;
        *c = v.y;
        return v.x != kEmpty;
    }
    k->lo = slot[0];
    k->hi = 0;
    *c = 1ULL;
    return k->lo != kEmpty;
}

// occupied entries -> dense row arrays (key_hi may be nullptr for narrow keys).  A warp looks at kCompactSub x 32
// consecutive slots per round and reserves their rows with ONE atomic, so the row counter is not the bottleneck.
constexpr int kCompactSub = 8;
__global__ void k_compact(const DevTable t, unsigned long long* __restrict__ key_lo, unsigned long long* __restrict__ key_hi,
                          unsigned long long* __restrict__ count, unsigned long long* __restrict__ n_rows) {
    const unsigned long long cap = t.cap;
    const int lane = threadIdx.x & 31;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    const unsigned long long span = 32ull * kCompactSub;
    for (unsigned long long base = warp * span; base < cap; base += n_warps * span) {  // warp-uniform trip count
        Key k[kCompactSub];
        unsigned long long c[kCompactSub];
        unsigned have = 0, total = 0;
#pragma unroll
        for (int j = 0; j < kCompactSub; j++) {
            const unsigned long long i = base + (unsigned long long)j * 32 + lane;
            k[j] = Key{0, 0};
            c[j] = 0;
            const bool h = i < cap && table_entry(t, i, &k[j], &c[j]);
            if (h) have |= 1u << j;
            total += __popc(__ballot_sync(0xFFFFFFFFu, h));
        }
        if (!total) continue;
        unsigned long long basepos = 0;
        if (lane == 0) basepos = atomicAdd(n_rows, (unsigned long long)total);
        basepos = __shfl_sync(0xFFFFFFFFu, basepos, 0);
        unsigned run = 0;
#pragma unroll
        for (int j = 0; j < kCompactSub; j++) {
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, (have >> j) & 1u);
            if ((have >> j) & 1u) {
                const unsigned long long p = basepos + run + __popc(bal & ((1u << lane) - 1u));
                key_lo[p] = k[j].lo;
                if (key_hi) key_hi[p] = k[j].hi;
                count[p] = c[j];
            }
            run += __popc(bal);
        }
    }
}

cudaError_t launch_compact(const DevTable& t, unsigned long long* key_lo, unsigned long long* key_hi,
                           unsigned long long* count, unsigned long long* n_rows, cudaStream_t stream) {
    k_compact<<<grid_for(t.cap, 256), 256, 0, stream>>>(t, key_lo, key_hi, count, n_rows);
    return cudaGetLastError();
}

__global__ void k_marginal(const unsigned long long* __restrict__ key_lo, const unsigned long long* __restrict__ key_hi,
                           const unsigned long long* __restrict__ count, const unsigned long long n_rows, const Key mask,
                           const DevTable dst) {
    unsigned long long fresh = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_rows;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        Key k{key_lo[i] & mask.lo, (key_hi ? key_hi[i] : 0ULL) & mask.hi};
        bool is_new;
        map_add(dst, k, count[i], &is_new);
        if (is_new) fresh++;
    }
    for (int o = 16; o; o >>= 1) fresh += __shfl_xor_sync(0xFFFFFFFFu, fresh, o);
    if ((threadIdx.x & 31) == 0 && fresh && dst.n_entries) atomicAdd(dst.n_entries, fresh);
}

cudaError_t launch_marginal(const unsigned long long* key_lo, const unsigned long long* key_hi,
                            const unsigned long long* count, unsigned long long n_rows, Key mask, const DevTable& dst,
                            cudaStream_t stream) {
    if (n_rows == 0) return cudaSuccess;
    k_marginal<<<grid_for(n_rows, 256), 256, 0, stream>>>(key_lo, key_hi, count, n_rows, mask, dst);
    return cudaGetLastError();
}

// ---- K4: enrichment marginals, dense (info.rs:840-904).  One pass over the final rows fills every single- and
// double-barcode marginal at once.  Singles (a few thousand counters that every row hits) are privatised in shared memory and
// flushed once per CTA; doubles (2^20 counters each for three 1,024-barcode slots: L2-resident) take one fire-and-forget
// RED per row and pair.
__device__ __forceinline__ uint32_t key_field(unsigned long long lo, unsigned long long hi, uint32_t shift, uint32_t bits) {
    if (bits == 0) return 0u;
    unsigned long long v;
    if (shift >= 64) v = hi >> (shift - 64);
    else v = (lo >> shift) | (shift ? hi << (64 - shift) : 0ULL);
    return (uint32_t)v & (0xFFFFFFFFu >> (32 - bits));  // index fields are at most 24 bits wide (plan_marginals)
}

__global__ void __launch_bounds__(256) k_marginals(const unsigned long long* __restrict__ key_lo, const unsigned long long* __restrict__ key_hi,
                                                   const unsigned long long* __restrict__ count, const unsigned long long n_rows,
                                                   const __grid_constant__ MargPlan plan, unsigned long long* __restrict__ dense,
                                                   const int priv) {
    extern __shared__ unsigned long long s_single[];  // plan.n_single counters when priv
    if (priv) {
        for (uint32_t i = threadIdx.x; i < plan.n_single; i += blockDim.x) s_single[i] = 0ULL;
        __syncthreads();
    }
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_rows;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long lo = key_lo[i], hi = key_hi ? key_hi[i] : 0ULL, c = count[i];
        const unsigned long long smp = key_field(lo, hi, plan.s_shift, plan.s_bits);
        for (uint32_t a = 0; a < plan.k; a++) {
            const unsigned long long idx = plan.s_off[a] + ((smp << plan.f_bits[a]) | key_field(lo, hi, plan.f_shift[a], plan.f_bits[a]));
            if (priv) atomicAdd(&s_single[idx], c);
            else atomicAdd(&dense[idx], c);
        }
        for (uint32_t p = 0; p < plan.n_pairs; p++) {
            const uint32_t a = plan.pa[p], b = plan.pb[p];
            const unsigned long long fa = key_field(lo, hi, plan.f_shift[a], plan.f_bits[a]);
            const unsigned long long fb = key_field(lo, hi, plan.f_shift[b], plan.f_bits[b]);
            atomicAdd(&dense[plan.d_off[p] + ((((smp << plan.f_bits[b]) | fb) << plan.f_bits[a]) | fa)], c);
        }
    }
    if (priv) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < plan.n_single; i += blockDim.x) {
            const unsigned long long v = s_single[i];
            if (v) atomicAdd(&dense[i], v);  // the singles come first in `dense`
        }
    }
}

size_t marginal_smem_limit() { return 48 * 1024; }

cudaError_t launch_marginals_dense(const unsigned long long* key_lo, const unsigned long long* key_hi, const unsigned long long* count,
                                   unsigned long long n_rows, const MargPlan& plan, unsigned long long* dense, cudaStream_t stream) {
    if (n_rows == 0) return cudaSuccess;
    const bool priv = plan.n_single * 8 <= marginal_smem_limit();
    unsigned long long g = (n_rows + 255) / 256;
    if (g > 148ULL * 8) g = 148ULL * 8;  // persistent CTAs: the private singles are flushed once per CTA
    k_marginals<<<(unsigned)g, 256, priv ? plan.n_single * 8 : 0, stream>>>(key_lo, key_hi, count, n_rows, plan, dense, priv ? 1 : 0);
    return cudaGetLastError();
}

// blockIdx.y = marginal (singles first, then pairs); non-zero counters of it -> rows, or just their number
__global__ void k_marginal_rows(const __grid_constant__ MargPlan plan, const uint32_t m_first, const unsigned long long* __restrict__ dense,
                                unsigned long long* __restrict__ key_lo, unsigned long long* __restrict__ key_hi,
                                unsigned long long* __restrict__ count, uint32_t* __restrict__ mask, unsigned long long* __restrict__ n_out) {
    const uint32_t m = m_first + blockIdx.y;
    const bool single = m < plan.k;
    const uint32_t a = single ? m : plan.pa[m - plan.k], b = single ? 0u : plan.pb[m - plan.k];
    const unsigned long long off = single ? plan.s_off[a] : plan.d_off[m - plan.k];
    const uint32_t bits = plan.s_bits + plan.f_bits[a] + (single ? 0u : plan.f_bits[b]);
    const unsigned long long n = 1ULL << bits;
    const int lane = threadIdx.x & 31;
    const unsigned long long span = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long first = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    for (unsigned long long i0 = first - lane; i0 < n; i0 += span) {  // warp-uniform trip count
        const unsigned long long i = i0 + lane;
        const unsigned long long c = i < n ? dense[off + i] : 0ULL;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, c != 0ULL);
        if (!bal) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(n_out, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (c != 0ULL && key_lo) {
            const unsigned long long pos = base + __popc(bal & ((1u << lane) - 1u));
            const unsigned long long fa = i & ((1ULL << plan.f_bits[a]) - 1ULL);
            unsigned long long rest = i >> plan.f_bits[a], fb = 0;
            if (!single) {
                fb = rest & ((1ULL << plan.f_bits[b]) - 1ULL);
                rest >>= plan.f_bits[b];
            }
            Key k{0, 0};
            key_or(k, fa, plan.f_shift[a]);
            if (!single) key_or(k, fb, plan.f_shift[b]);
            if (plan.s_bits) key_or(k, rest, plan.s_shift);
            key_lo[pos] = k.lo;
            key_hi[pos] = k.hi;
            count[pos] = c;
            mask[pos] = single ? 1u << a : (1u << a) | (1u << b);
        }
    }
}

cudaError_t launch_marginals_rows(const MargPlan& plan, uint32_t m_first, uint32_t m_count, const unsigned long long* dense,
                                  unsigned long long* key_lo, unsigned long long* key_hi, unsigned long long* count, uint32_t* mask,
                                  unsigned long long* n_out, cudaStream_t stream) {
    if (m_count == 0) return cudaSuccess;
    if (m_first + m_count > plan.k + plan.n_pairs) return cudaErrorInvalidValue;
    k_marginal_rows<<<dim3(148 * 2, m_count), 256, 0, stream>>>(plan, m_first, dense, key_lo, key_hi, count, mask, n_out);
    return cudaGetLastError();
}

__global__ void k_add_u64(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, const unsigned long long n) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] += src[i];
}

cudaError_t launch_add_u64(unsigned long long* dst, const unsigned long long* src, unsigned long long n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_add_u64<<<grid_for(n, 256), 256, 0, stream>>>(dst, src, n);
    return cudaGetLastError();
}

// ---- transfer form of a host batch -> the bc_batch layout the decode kernel reads (include/bc_b200.h: bc_wire_batch) ----
// What crosses PCIe is the lo / hi planes, the lengths, the N calls (a dense plane, or a list when they are rare) and the
// quality characters as 8-, 6-, 4- or 2-bit codes.  Three small streaming kernels rebuild the fixed-stride records in HBM;
// at 8.4 M reads per batch they take well under a millisecond against tens of milliseconds for the copy they shorten.
__global__ void k_wire_planes(const WireView w, uint32_t* __restrict__ planes, uint16_t* __restrict__ read_len, const uint32_t plane_stride) {
    const unsigned long long total = (unsigned long long)w.n_reads * w.W;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long r = i / w.W;
        const uint32_t k = (uint32_t)(i - r * w.W);
        uint32_t* rec = planes + r * plane_stride;
        const uint32_t nm = w.nmask ? w.nmask[i] : 0u;
        rec[k] = w.lohi[r * 2 * w.W + k] & ~nm;  // lo = hi = 0 where the read has N: the convention of bc_batch
        rec[w.W + k] = w.lohi[r * 2 * w.W + w.W + k] & ~nm;
        rec[2 * w.W + k] = nm;
        if (k == 0) {
            if (plane_stride > 3 * w.W) rec[3 * w.W] = 0u;  // pad word of an even record
            read_len[r] = w.read_len[r];
        }
    }
}

// the N calls of the list: one bit set in the N plane, the same bit cleared in lo and hi (entries out of range are ignored)
__global__ void k_wire_ncalls(const WireView w, uint32_t* __restrict__ planes, const uint32_t plane_stride) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < w.n_calls;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t r = w.n_read[i], pos = w.n_pos[i];
        if (r >= w.n_reads || pos >= 32 * w.W) continue;
        uint32_t* rec = planes + (unsigned long long)r * plane_stride;
        const uint32_t bit = 1u << (pos & 31);
        atomicOr(&rec[2 * w.W + (pos >> 5)], bit);
        atomicAnd(&rec[pos >> 5], ~bit);
        atomicAnd(&rec[w.W + (pos >> 5)], ~bit);
    }
}

// one thread per four quality characters of a read: code i of a record sits at bits [i * BITS, (i + 1) * BITS) of its
// little-endian bit stream
template <int BITS>
__global__ void k_wire_qual(const WireView w, uint8_t* __restrict__ qual, const uint32_t qual_stride) {
    const uint32_t words = qual_stride / 4;  // output words per read
    const unsigned long long total = (unsigned long long)w.n_reads * words;
    const uint32_t* dict = reinterpret_cast<const uint32_t*>(w.dict.c);
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long r = i / words;
        const uint32_t j = (uint32_t)(i - r * words);
        uint32_t out = 0x21212121u;  // '!' beyond the packed codes
        if (4 * j < w.n_codes) {
            const uint8_t* rec = w.qual + r * w.qual_stride;
            if (BITS == 8) {
                out = reinterpret_cast<const uint32_t*>(rec)[j];
            } else {
                const uint32_t* rw = reinterpret_cast<const uint32_t*>(rec);
                const uint32_t bit = 4 * BITS * j, wi = bit >> 5, sh = bit & 31;
                const uint32_t a = rw[wi], b = (sh + 4 * BITS > 32) ? rw[wi + 1] : 0u;
                const uint32_t v = __funnelshift_r(a, b, sh);
                out = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t c = (v >> (BITS * k)) & ((1u << BITS) - 1u);
                    uint32_t ch;
                    if (BITS == 6) ch = c == 63u ? 0xFFu : c + 33u;
                    else ch = (dict[c >> 2] >> (8 * (c & 3u))) & 0xFFu;
                    out |= ch << (8 * k);
                }
            }
        }
        reinterpret_cast<uint32_t*>(qual + r * qual_stride)[j] = out;
    }
}

cudaError_t launch_wire_expand(const WireView& w, uint32_t* planes, uint16_t* read_len, uint8_t* qual, uint32_t plane_stride,
                               uint32_t qual_stride, cudaStream_t stream) {
    if (w.n_reads == 0) return cudaSuccess;
    k_wire_planes<<<grid_for((unsigned long long)w.n_reads * w.W, 256), 256, 0, stream>>>(w, planes, read_len, plane_stride);
    if (!w.nmask && w.n_calls) k_wire_ncalls<<<grid_for(w.n_calls, 256), 256, 0, stream>>>(w, planes, plane_stride);
    if (qual) {
        const unsigned g = grid_for((unsigned long long)w.n_reads * (qual_stride / 4), 256);
        switch (w.qual_bits) {
            case 8: k_wire_qual<8><<<g, 256, 0, stream>>>(w, qual, qual_stride); break;
            case 6: k_wire_qual<6><<<g, 256, 0, stream>>>(w, qual, qual_stride); break;
            case 4: k_wire_qual<4><<<g, 256, 0, stream>>>(w, qual, qual_stride); break;
            case 2: k_wire_qual<2><<<g, 256, 0, stream>>>(w, qual, qual_stride); break;
            default: return cudaErrorInvalidValue;
        }
    }
    return cudaGetLastError();
}

// empty map: every slot {key = kEmpty, count = 0}
__global__ void k_clear_map(ulonglong2* __restrict__ slots, const unsigned long long n16, const int wide) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n16;
         i += (unsigned long long)gridDim.x * blockDim.x)
        slots[i] = wide ? ((i & 1) ? make_ulonglong2(0ULL, 0ULL) : make_ulonglong2(kEmpty, kEmpty)) : make_ulonglong2(kEmpty, 0ULL);
}

cudaError_t launch_clear_map(const DevTable& t, cudaStream_t stream) {
    const unsigned long long n16 = t.cap * (t.wide ? 2ull : 1ull);
    k_clear_map<<<grid_for(n16, 256), 256, 0, stream>>>(reinterpret_cast<ulonglong2*>(t.data), n16, t.wide);
    return cudaGetLastError();
}

// move every entry of `src` (hash kinds) into the larger `dst` of the same kind
__global__ void k_rehash(const DevTable src, const DevTable dst) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < src.cap;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        Key k;
        unsigned long long c;
        if (!table_entry(src, i, &k, &c)) continue;
        bool is_new;
        if (dst.kind == 1) map_add(dst, k, c, &is_new);
        else table_find_or_insert(dst, k, &is_new);
    }
}

cudaError_t launch_rehash(const DevTable& src, const DevTable& dst, cudaStream_t stream) {
    k_rehash<<<grid_for(src.cap, 256), 256, 0, stream>>>(src, dst);
    return cudaGetLastError();
}

}  // namespace bc
