// bc_device.cuh — device-side layout of the run configuration, packed keys and the open-addressing tables.
// sm_100a only.  Semantics follow SURVEY.md §3.3 (P1-P9, Q1-Q22); citations are file:line into the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bc {

constexpr int kTile = 128;          // reads per CTA tile == threads per CTA
constexpr int kMaxTW = 8;           // template words (BC_MAX_TEMPLATE / 32)
constexpr int kMaxSlots = 16;
constexpr int kMaxQRuns = 32;
constexpr uint32_t kFail = 0xFFFFFFFFu;

enum SlotMode : uint8_t { MODE_RAW = 0, MODE_TABLE = 1, MODE_HASH = 2, MODE_SCAN = 3 };

constexpr int kMaxBlocks = 8;       // pigeonhole blocks of the deep index (max_err + 1 <= kMaxBlocks)
constexpr int kMaxBlockKey = 8;     // bases of a block that form its bucket key (4^8 buckets)
constexpr uint32_t kHalfProbeCap = 24;

struct DevSlot {
    uint16_t offset, len, max_err;
    uint8_t kind;        // 'S' 'B' 'R'
    uint8_t mode;        // SlotMode
    uint32_t n_ref;
    uint32_t ref_off;    // first uint4 {lo,hi,nm,len} of this slot in the reference array
    uint32_t aux_off;    // TABLE: first u16 of the 4^len lookup; HASH: first entry of the exact-match hash
    uint32_t aux_mask;   // HASH: capacity-1
    uint16_t key_shift;  // first key bit of this slot's field
    uint16_t key_bits;   // index bits, or 3*len for raw fields ([lo:len][hi:len][nm:len])
    // Exact pruned search for long barcodes (every reference as long as the slot and N-free):
    //  half index  — two hashes keyed by the first / second half of the barcode; holds every reference within
    //                Hamming distance 1 of a query (one of the halves is then error-free)
    //  block index — levels for distance <= 2, 3, .., max_err: level k cuts the barcode into k+1 blocks; a reference
//                within k of the query agrees with it on a whole block.  Levels are tried in turn: most misses are
//                settled by the first one, whose buckets are small.
    uint8_t has_half, n_levels;  // block index levels: level i is complete for distance <= 2 + i
    uint8_t n_inline;            // TABLE: N-free references of one length -> queries with 1-2 N resolve from the table
    uint16_t half_len0;          // bases in the first half
    uint32_t half_off, half_mask;  // two tables of half_mask+1 u64 entries {key32, id32} each, at half_off and half_off+cap
    uint32_t deep_off;           // index of this slot's first DevDeep descriptor (one per level)
};

struct DevDeep {  // one level of a slot's block index (global memory, read by k_resolve only): n_blocks = cap + 1 blocks
    uint32_t n_blocks, cap;
    uint8_t key_pos[kMaxBlocks], key_len[kMaxBlocks];  // bucket key = bases [key_pos, key_pos+key_len) of the barcode
    uint32_t start_off[kMaxBlocks];                    // first of 4^key_len + 1 u32 bucket starts (CSR) in `csr`
    uint32_t ids_off[kMaxBlocks];                      // first of n_ref {lo, hi, id, 0} references, bucket order, in `bref`
};

struct DevQRun {
    uint16_t off, len;   // position in the region walk (relative to the quality start) and length
    uint16_t n_words;    // ceil(len / 4): re-aligned words that hold the run
    uint32_t tail_mask;  // bytes of the last word that belong to the run
    uint32_t thresh;     // low quality iff the sum of the run's raw Phred+33 bytes < thresh  (parse.rs:352-355, Q12)
};

struct DevCfg {
    uint32_t L, TW, n_slots, n_qruns, max_const_err, has_fn;
    uint32_t pivot;  // template word with the most constant bases: the locate prefilter looks at it alone
    uint32_t has_umi, umi_bits, key_bits, wide;
    uint32_t t_lo[kMaxTW], t_hi[kMaxTW], t_cm[kMaxTW], t_fn[kMaxTW];
    // bit-sliced pivot prefilter (max_const_err <= 15): the pivot word's constant positions grouped by their base
    // (0 A, 1 C, 2 G, 3 T); pv_sh4[b][i] packs the positions 4 i .. 4 i + 3 of base b, one per byte
    uint32_t bs_ok, bs_k;  // bs_k = 15 - max_const_err: counters start there, overflow into bit 4 <=> too many mismatches
    uint32_t pv_n[4];
    uint32_t pv_sh4[4][8];
    // static-block variant (bs_two = 1): constant positions of the template words pivot and pivot + 1, per base rounded
    // DOWN to whole blocks of four (a subset of the positions is still a necessary condition), at most two blocks per
    // (word, base): bs2_n = blocks, bs2_sh = one shift per u32 so that a funnel shift takes it as a constant operand
    uint32_t bs_two;
    uint32_t bs2_n[2][4];
    uint32_t bs2_sh[2][4][8];
    // exact-match prefilter (phase A of the locate step): template word `xpivot`, four of its constant positions per base
    // (a base with fewer than four repeats one: testing a position twice is harmless; xs_has bit b = base b has any).
    // A window survives iff the read equals the template at all of them — necessary for an exact match.
    uint32_t xpivot, xs_has;
    uint32_t xs_sh[4][4];
    DevSlot slots[kMaxSlots];
    uint8_t order[kMaxSlots];  // sample first, then counted barcodes in order, then the random barcode
    DevQRun qruns[kMaxQRuns];
};

struct Key {
    unsigned long long lo, hi;
};

struct BatchView {
    const uint32_t* planes;
    const uint16_t* read_len;
    const uint8_t* qual;  // nullptr when the quality filter is off
    uint32_t n_reads, plane_stride, qual_stride, W;
    uint32_t rep_chunks;  // 32-offset chunks the repair range of a read of this batch can have: ceil((32 W - L) / 32)
};

// Open-addressing tables.  kind 0: dense counts (index = key); kind 1: hash map key -> count; kind 2: hash set.
// A map slot keeps key and count side by side so that one read touches one 32-byte sector:
//   map narrow {key, count}   map wide {lo, hi, count, 0}   set narrow {key}   set wide {lo, hi}
struct DevTable {
    unsigned long long* data;
    unsigned long long cap;  // slots (any size: the home slot is mulhi(hash, cap))
    unsigned long long* n_entries;
    int kind;
    int wide;
};
__host__ __device__ __forceinline__ uint32_t table_stride(int kind, int wide) {  // u64 words per slot
    return kind == 0 ? 1u : kind == 1 ? (wide ? 4u : 2u) : (wide ? 2u : 1u);
}

// What a matched read is counted into (info.rs:735-808).  Without a random barcode: map[key] += 1.  With one the
// (key, UMI) pair goes into `set` and only a pair seen for the first time bumps map[key] (info.rs:780-791), so the
// map always holds the final counts (output.rs:265-270) and nothing has to be regrouped at the end.
struct Tables {
    DevTable map;
    DevTable set;
    uint32_t umi_bits;
    int has_set;
};

// ---- what the decode kernel is handed beside the run constants (host side: bc_kernels.h) ---------------------------
struct DecodeOut {  // all optional (test hooks / decode_only)
    uint8_t* status;
    int16_t* offset;
    uint8_t* repaired;
    int32_t* slot_index;
    unsigned long long* key_lo;
    unsigned long long* key_hi;
};

struct DevAux {  // reference sets and their accelerators
    const uint4* refs;                    // {lo, hi, nm, len} per reference barcode
    const uint32_t* tables;               // 4^len direct lookups (MODE_TABLE): idx | dist << 16 | tie << 24
    const unsigned long long* hash_keys;  // exact-match hash (MODE_HASH): lo | hi << 32
    const uint32_t* hash_idx;
    const unsigned long long* half;       // half index: {key32, id32} entries, kEmpty = free
    const DevDeep* deep;                  // block index descriptors
    const uint32_t* csr;                  // block index: bucket starts
    const uint4* bref;                    // block index: references in bucket order, {lo, hi, id, 0}
};

struct Deferred {  // reads whose barcode step needs a search: {read index, offset | repaired << 16}
    uint2* items;
    uint32_t* count;       // filled by k_decode, drained by k_resolve
    uint32_t* next_count;  // the counter of the NEXT batch (two alternate): k_resolve zeroes it, so no memset per batch
};

// Deferred counting (bc_partition.cu): instead of updating the tables read by read, a matched read's packed key goes
// to slot (base + read index) of a flat record buffer — kEmpty marks the reads that did not match — and the whole
// job is de-duplicated and counted at flush time, partition by partition, in shared memory.
struct RecOut {
    unsigned long long* lo;
    unsigned long long* hi;                // nullptr when the full key (random barcode included) fits 63 bits
    unsigned long long base;               // records appended before this batch (every batch appends exactly n_reads slots)
};

enum DecodeFlags { F_INSERT = 1, F_EMIT = 2, F_LOCATE_ONLY = 8, F_APPEND = 16 };

// k_decode adds its outcome counters to striped copies (kCounterStripes x kCounterStride u64: the BC_N_COUNTERS
// outcomes, then new map / set entries); launch_fold_counters sums them into the BC_N_COUNTERS + 2 counters of the ctx.
constexpr uint32_t kCounterStripes = 64, kCounterStride = 16;

// ---------------------------------------------------------------------------------------------------------

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
__device__ __forceinline__ unsigned long long hash_key(Key k) { return mix64(k.lo ^ mix64(k.hi + 0x9e3779b97f4a7c15ULL)); }

__device__ __forceinline__ void key_or(Key& k, unsigned long long v, uint32_t shift) {
    if (shift < 64) {
        k.lo |= v << shift;
        if (shift) k.hi |= v >> (64 - shift);
    } else {
        k.hi |= v << (shift - 64);
    }
}
__device__ __forceinline__ Key key_shr(Key k, uint32_t s) {
    if (s == 0) return k;
    Key r;
    if (s < 64) {
        r.lo = (k.lo >> s) | (k.hi << (64 - s));
        r.hi = k.hi >> s;
    } else {
        r.lo = k.hi >> (s - 64);
        r.hi = 0;
    }
    return r;
}

__device__ __forceinline__ ulonglong2 cas128(ulonglong2* addr, ulonglong2 cmp, ulonglong2 val) {
    ulonglong2 old;
    asm volatile(
        "{\n"
        ".reg .b128 c, v, o;\n"
        "mov.b128 c, {%2, %3};\n"
        "mov.b128 v, {%4, %5};\n"
        "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n"
        "mov.b128 {%0, %1}, o;\n"
        "}\n"
        : "=l"(old.x), "=l"(old.y)
        : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr)
        : "memory");
    return old;
}

constexpr unsigned long long kEmpty = ~0ULL;

// Find-or-claim the slot of `key`.  Returns the slot index; *is_new tells whether this call claimed it.
// The host keeps the load factor <= 0.6, so linear probing stays short.
__device__ __forceinline__ unsigned long long table_find_or_insert(const DevTable& t, Key key, bool* is_new) {
    unsigned long long h = __umul64hi(hash_key(key), t.cap);
    const uint32_t stride = table_stride(t.kind, t.wide);
    if (!t.wide) {
        for (;;) {
            unsigned long long* slot = t.data + h * stride;
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(slot);
            if (cur == key.lo) { *is_new = false; return h; }
            if (cur == kEmpty) {
                unsigned long long old = atomicCAS(slot, kEmpty, key.lo);
                if (old == kEmpty) { *is_new = true; return h; }
                if (old == key.lo) { *is_new = false; return h; }
            }
            if (++h == t.cap) h = 0;
        }
    } else {
        const ulonglong2 empty = make_ulonglong2(kEmpty, kEmpty);
        const ulonglong2 mine = make_ulonglong2(key.lo, key.hi);
        for (;;) {
            ulonglong2* slot = reinterpret_cast<ulonglong2*>(t.data + h * stride);
            // read first: a key that is already there (hot keys!) costs a load, not a contended 128-bit CAS.  A torn
            // read can only mix "empty" and the final value, so it never equals `mine` unless the slot holds it.
            ulonglong2 cur;
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(cur.x), "=l"(cur.y) : "l"(slot));
            if (cur.x == key.lo && cur.y == key.hi) { *is_new = false; return h; }
            if (cur.x == kEmpty || cur.y == kEmpty) {
                const ulonglong2 old = cas128(slot, empty, mine);
                if (old.x == kEmpty && old.y == kEmpty) { *is_new = true; return h; }
                if (old.x == key.lo && old.y == key.hi) { *is_new = false; return h; }
            }
            if (++h == t.cap) h = 0;
        }
    }
}

// map[key] += add (dense: counts[key] += add)
__device__ __forceinline__ void map_add(const DevTable& t, Key key, unsigned long long add, bool* is_new) {
    *is_new = false;
    if (t.kind == 0) {
        atomicAdd(&t.data[key.lo], add);
        return;
    }
    const unsigned long long h = table_find_or_insert(t, key, is_new);
    atomicAdd(t.data + h * (t.wide ? 4u : 2u) + (t.wide ? 2u : 1u), add);
}

// One matched read (key includes the random barcode when the scheme has one).  Returns true when the read counts as
// "matched", false when it is a duplicate (parse.rs:65-69).
__device__ __forceinline__ bool count_read(const Tables& T, Key key, bool* new_key, bool* new_pair) {
    *new_key = false;
    *new_pair = false;
    if (!T.has_set) {
        map_add(T.map, key, 1ULL, new_key);
        return true;
    }
    table_find_or_insert(T.set, key, new_pair);
    if (!*new_pair) return false;
    map_add(T.map, key_shr(key, T.umi_bits), 1ULL, new_key);
    return true;
}

}  // namespace bc
