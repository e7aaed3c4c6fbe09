// bc_decode.cuh — device code of k_decode: the fused locate / quality / barcode-correction / record step of one read
// (parse.rs:89-163, 270-375, 439-524, 553-593; info.rs:60-127).  Compiled twice from this one text: ahead of time as the
// generic kernel k_decode<TW> (run constants in a __grid_constant__ DevCfg, any scheme), and at run time by NVRTC with
// BC_JIT defined, where the run constants are a constexpr object: every loop over template words, pivot positions,
// quality runs and barcodes unrolls, shifts and masks become immediates and the branches on a barcode's look-up mode fold
// away (bc_jit.cu).  Same statements either way, so both give the same results.
#pragma once
#include "../../include/bc_b200.h"
#include "bc_device.cuh"

#ifdef BC_JIT
#define BC_UNROLL _Pragma("unroll")
#else
#define BC_UNROLL
#endif

namespace bc {

__device__ __forceinline__ uint32_t lenmask(uint32_t len) { return len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u); }

// BC_DECODE_CHECKED = true builds k_decode with the bounds-checked plane reads (a verification build: tools/check_plane_reads.py
// runs it against the shipped one; compute-sanitizer is closed on the pool this was developed on)
#ifndef BC_DECODE_CHECKED
#define BC_DECODE_CHECKED false
#endif

// bits [pos, pos+32) of a W-word bit plane.  CHECK = false (k_decode, planes staged in shared memory): every caller
// has pos < 32 W, so word j exists, and word j + 1 is at worst the first word of the next plane / record / array of
// the tile — readable, and its bits land beyond the read's last base, where every consumer masks (template constant
// mask, slot length mask).  CHECK = true (k_resolve reads the batch in global memory) never reads past the plane.
template <bool CHECK>
__device__ __forceinline__ uint32_t plane_bits(const uint32_t* p, uint32_t W, uint32_t pos) {
    const uint32_t j = pos >> 5, s = pos & 31;
    if (!CHECK) return __funnelshift_r(p[j], p[j + 1], s);
    const uint32_t a = j < W ? p[j] : 0u;
    const uint32_t b = (j + 1) < W ? p[j + 1] : 0u;
    return __funnelshift_r(a, b, s);
}

// ---- TMA (bulk async copy) + mbarrier, raw PTX ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- fix_error (parse.rs:553-593) pieces -------------------------------------------------------------------
struct Best {  // unique-minimum tracker: smallest distance, how many references reach it, one of them
    uint32_t d, cnt, arg, exact;
};
__device__ __forceinline__ void best_add(Best& b, uint32_t d, uint32_t i) {
    if (d < b.d) {
        b.d = d;
        b.cnt = 1;
        b.arg = i;
    } else if (d == b.d) {
        b.cnt++;
    }
}
// distance of a query to one reference: compared over the shorter of the two, N on either side never counts (Q10)
__device__ __forceinline__ uint32_t ref_dist(uint4 r, uint32_t blo, uint32_t bhi, uint32_t bnm, uint32_t len, uint32_t lm) {
    const uint32_t m = r.w < len ? lenmask(r.w) : lm;
    return __popc(((blo ^ r.x) | (bhi ^ r.y)) & ~bnm & ~r.z & m);
}
__device__ __forceinline__ bool ref_same(uint4 r, uint32_t blo, uint32_t bhi, uint32_t bnm, uint32_t len) {
    return r.w == len && r.x == blo && r.y == bhi && r.z == bnm;
}

// One thread, whole reference set (table building only): exact membership wins (parse.rs:457,489), otherwise the
// unique minimum within max_err (Q5).
__device__ __forceinline__ uint32_t scan_refs(const uint4* __restrict__ refs, uint32_t n_ref, uint32_t blo, uint32_t bhi,
                                              uint32_t bnm, uint32_t len, uint32_t max_err) {
    Best b{max_err + 1, 0, kFail, kFail};
    const uint32_t lm = lenmask(len);
    for (uint32_t i = 0; i < n_ref; i++) {
        const uint4 r = __ldg(&refs[i]);
        if (ref_same(r, blo, bhi, bnm, len)) b.exact = i;
        best_add(b, ref_dist(r, blo, bhi, bnm, len, lm), i);
    }
    if (b.exact != kFail) return b.exact;
    return (b.cnt == 1 && b.d <= max_err) ? b.arg : kFail;
}

__device__ __forceinline__ uint32_t hash_exact(const DevAux& aux, const DevSlot& S, uint32_t blo, uint32_t bhi) {
    unsigned long long k = (unsigned long long)blo | ((unsigned long long)bhi << 32);
    unsigned long long h = mix64(k) & S.aux_mask;
    for (;;) {
        uint32_t idx = __ldg(&aux.hash_idx[S.aux_off + h]);
        if (idx == kFail) return kFail;
        if (__ldg(&aux.hash_keys[S.aux_off + h]) == k) return idx;
        h = (h + 1) & S.aux_mask;
    }
}

// Direct table entry (MODE_TABLE): [15:0] reference index, [23:16] smallest distance over the set, [24] several
// references reach it.  An identical reference is entered as distance 0 without a tie (parse.rs:457,489).
__device__ __forceinline__ uint32_t table_pick(uint32_t e, uint32_t max_err) {
    return (!(e & 0x1000000u) && ((e >> 16) & 0xFFu) <= max_err) ? (e & 0xFFFFu) : kFail;
}
// A query with one N: d(query, c) ignoring the N position equals the minimum over the four completions x' of that
// position of d(x', c), so the unique-minimum rule can be read off the completions' entries: the overall minimum is
// the smallest entry distance, and it is unique iff every completion reaching it names one and the same reference
// without a tie.  Valid for N-free reference sets of one length (the host sets DevSlot::n_inline only then).
// Queries with more N go to k_resolve.
// One N (the common case of the rare case): the four completions, unrolled.
__device__ __forceinline__ uint32_t table_lookup_1n(const uint32_t* __restrict__ tab, const DevSlot& S, uint32_t lo, uint32_t hi,
                                                    uint32_t nm) {
    const uint32_t base = lo | (hi << S.len);
    const uint32_t e0 = __ldg(&tab[base]), e1 = __ldg(&tab[base | nm]), e2 = __ldg(&tab[base | (nm << S.len)]),
                   e3 = __ldg(&tab[base | nm | (nm << S.len)]);
    // key = distance << 17 | tie << 16 | id: the smallest key is the smallest distance, ties flagged first
    const uint32_t k0 = ((e0 >> 16) & 0xFFu) << 17 | ((e0 >> 24) & 1u) << 16 | (e0 & 0xFFFFu);
    const uint32_t k1 = ((e1 >> 16) & 0xFFu) << 17 | ((e1 >> 24) & 1u) << 16 | (e1 & 0xFFFFu);
    const uint32_t k2 = ((e2 >> 16) & 0xFFu) << 17 | ((e2 >> 24) & 1u) << 16 | (e2 & 0xFFFFu);
    const uint32_t k3 = ((e3 >> 16) & 0xFFu) << 17 | ((e3 >> 24) & 1u) << 16 | (e3 & 0xFFFFu);
    const uint32_t dmin = min(min(k0 >> 17, k1 >> 17), min(k2 >> 17, k3 >> 17));
    // unique iff every completion at the minimum distance names the same reference and none of them is a tie
    uint32_t cand = kFail;
    bool multi = false;
    const uint32_t ks[4] = {k0, k1, k2, k3};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if ((ks[i] >> 17) != dmin) continue;
        const uint32_t id = ks[i] & 0xFFFFu;
        if ((ks[i] >> 16) & 1u) multi = true;
        if (cand == kFail) cand = id;
        else if (cand != id) multi = true;
    }
    return (!multi && dmin <= S.max_err) ? cand : kFail;
}

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

// Half index probe for a query that missed the exact lookup (or holds one N).  Every reference within distance 1
// (N positions of the query never count) agrees with the query on all non-N bases of its first or of its second
// half, so it sits in the probe chain of one completion of that half.  A reference is counted at the first half
// it agrees on.  Returns
//   HALF_RESOLVED: *idx is the unique reference at distance <= 1, or kFail (a tie at the minimum, Q5, or cap too small)
//   HALF_DEEPER:   nothing within distance 1 — the block index has to look at distance 2..max_err -> defer
enum { HALF_RESOLVED = 0, HALF_DEEPER = 1 };
__device__ __forceinline__ int half_probe(const DevAux& aux, const DevSlot& S, uint32_t blo, uint32_t bhi, uint32_t bnm,
                                          uint32_t* idx) {
    const uint32_t lm = lenmask(S.len);
    const uint32_t cap = S.half_mask + 1;
    const uint32_t h0m = lenmask(S.half_len0);
    Best b{2, 0, kFail, kFail};
    bool overflow = false;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t pos0 = h ? S.half_len0 : 0u;
        const uint32_t hl = h ? (uint32_t)S.len - S.half_len0 : (uint32_t)S.half_len0;
        const uint32_t hm = lenmask(hl);
        const uint32_t klo = (blo >> pos0) & hm, khi = (bhi >> pos0) & hm, kn = (bnm >> pos0) & hm;
        const uint32_t np = kn ? (uint32_t)__ffs(kn) - 1u : 0u;
        const unsigned long long* tab = aux.half + S.half_off + (h ? cap : 0u);
        for (uint32_t comp = 0; comp < (kn ? 4u : 1u); comp++) {
            const uint32_t key = (klo | ((comp & 1u) << np)) | ((khi | (((comp >> 1) & 1u) << np)) << 16);
            uint32_t p = mix32(key) & S.half_mask;
            for (uint32_t probes = 0;; probes++) {
                const unsigned long long e = __ldg(&tab[p]);
                if (e == kEmpty) break;
                if (probes == kHalfProbeCap) {
                    overflow = true;
                    break;
                }
                if ((uint32_t)e == key) {
                    const uint32_t id = (uint32_t)(e >> 32);
                    const uint4 r = __ldg(&aux.refs[S.ref_off + id]);
                    const uint32_t diff = ((blo ^ r.x) | (bhi ^ r.y)) & ~bnm & lm;
                    if (h == 0 || (diff & h0m) != 0) best_add(b, __popc(diff), id);
                }
                p = (p + 1) & S.half_mask;
            }
        }
    }
    if (overflow) return HALF_DEEPER;  // a crowded chain was cut short: let the block index decide
    if (b.d <= 1) {
        *idx = (b.cnt == 1 && b.d <= S.max_err) ? b.arg : kFail;
        return HALF_RESOLVED;
    }
    if (S.max_err <= 1) {
        *idx = kFail;
        return HALF_RESOLVED;
    }
    return HALF_DEEPER;
}

// ---- bit-sliced mismatch counters for the pivot prefilter ---------------------------------------------------------
// Five bit planes hold, for 32 window offsets at once, a 5-bit counter per offset (plane k = bit k of every counter).
// bs_add4 adds four 1-bit inputs per offset with carry-save adders: 11 LOP3 for 4 x 32 additions.  The top plane
// only ever ORs carries in: it says "the counter passed 15".
struct BsPlanes {
    uint32_t p1, p2, p4, p8, p16;
};
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }
__device__ __forceinline__ void bs_add4(BsPlanes& P, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4) {
    const uint32_t a1 = P.p1 ^ x1 ^ x2, c1a = maj3(P.p1, x1, x2);
    const uint32_t a2 = a1 ^ x3 ^ x4, c1b = maj3(a1, x3, x4);
    P.p1 = a2;
    const uint32_t c2 = maj3(P.p2, c1a, c1b);
    P.p2 = P.p2 ^ c1a ^ c1b;
    const uint32_t c4 = P.p4 & c2;
    P.p4 ^= c2;
    const uint32_t c8 = P.p8 & c4;
    P.p8 ^= c4;
    P.p16 |= c8;
}
// all pivot positions of one base: M0/M1 = mismatch plane of the read against that base, words j and j + 1
__device__ __forceinline__ void bs_base(BsPlanes& P, uint32_t M0, uint32_t M1, uint32_t n, const uint32_t* sh4) {
    uint32_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const uint32_t sh = sh4[i >> 2];
        bs_add4(P, __funnelshift_r(M0, M1, sh), __funnelshift_r(M0, M1, sh >> 8), __funnelshift_r(M0, M1, sh >> 16),
                __funnelshift_r(M0, M1, sh >> 24));
    }
    const uint32_t r = n - i;
    if (r) {
        const uint32_t sh = sh4[i >> 2];
        bs_add4(P, __funnelshift_r(M0, M1, sh), r > 1 ? __funnelshift_r(M0, M1, sh >> 8) : 0u,
                r > 2 ? __funnelshift_r(M0, M1, sh >> 16) : 0u, 0u);
    }
}

// static-block variant: up to two whole blocks of four positions of one (template word, base) group; the shifts sit at
// fixed constant-bank addresses, so a block is four funnel shifts and the carry-save adds, nothing else
__device__ __forceinline__ void bs_blocks(BsPlanes& P, uint32_t M0, uint32_t M1, uint32_t n_blocks, const uint32_t* sh) {
    if (n_blocks > 0)
        bs_add4(P, __funnelshift_r(M0, M1, sh[0]), __funnelshift_r(M0, M1, sh[1]), __funnelshift_r(M0, M1, sh[2]), __funnelshift_r(M0, M1, sh[3]));
    if (n_blocks > 1)
        bs_add4(P, __funnelshift_r(M0, M1, sh[4]), __funnelshift_r(M0, M1, sh[5]), __funnelshift_r(M0, M1, sh[6]), __funnelshift_r(M0, M1, sh[7]));
}

// ---- K1: locate (parse.rs:89-96, 151-163, 287-313) ----------------------------------------------------------
// Two predicates per window (Q1): the regex's exact test (a read N in a constant fails, format-N needs ACGT) and
// the repair's masked Hamming distance (N on either side is a wildcard).  Leftmost exact window wins (P1);
// otherwise the unique minimum over offsets [0, R-L) within the cap (P2, Q3, Q5).
//
// The two run as separate phases of k_decode.  Phase A (every read): the exact test only.  Bit-sliced over 32 offsets:
// for constant position q of the pivot word, funnelshift(M, q) — M the read's "differs from that position's base, or
// is N" plane — says which of the 32 offsets fail there; OR-ing sixteen of them (four per base) leaves the offsets
// that agree with the template on all sixteen, and only those (typically the one true offset) get the full-width
// test.  16 funnel shifts + 8 LOP3 per 32 windows.  Phase B (only the reads phase A could not place, about a quarter
// of a typical run): the masked Hamming scan, cut into (read, 32-offset chunk) work items that the CTA's threads take
// in turn, so that the expensive counting prefilter below runs in full warps instead of in the few lanes of every
// warp whose read needs it.
template <int TW>
__device__ __forceinline__ int locate_exact(const DevCfg& cfg, const uint32_t* lo, const uint32_t* hi, const uint32_t* nm,
                                            const uint32_t W, const int nwin) {
    const uint32_t kp = cfg.xpivot;
    // word kp + c exists for every chunk c that holds a window; word kp + c + 1 may be the next plane's first word: see
    // plane_bits
    uint32_t al = lo[kp], ah = hi[kp], an = nm[kp];
    for (int c = 0; (c << 5) < nwin; c++) {
        const uint32_t j = kp + c + 1;
        const uint32_t bl = lo[j], bh = hi[j], bn = nm[j];
        uint32_t bad = 0;
        if (cfg.xs_has & 1u) {  // template A = (0, 0)
            const uint32_t M0 = al | ah | an, M1 = bl | bh | bn;
            bad |= __funnelshift_r(M0, M1, cfg.xs_sh[0][0]) | __funnelshift_r(M0, M1, cfg.xs_sh[0][1]);
            bad |= __funnelshift_r(M0, M1, cfg.xs_sh[0][2]) | __funnelshift_r(M0, M1, cfg.xs_sh[0][3]);
        }
        if (cfg.xs_has & 2u) {  // C = (1, 0)
            const uint32_t M0 = ~al | ah | an, M1 = ~bl | bh | bn;
            bad |= __funnelshift_r(M0, M1, cfg.xs_sh[1][0]) | __funnelshift_r(M0, M1, cfg.xs_sh[1][1]);
            bad |= __funnelshift_r(M0, M1, cfg.xs_sh[1][2]) | __funnelshift_r(M0, M1, cfg.xs_sh[1][3]);
        }
        if (cfg.xs_has & 4u) {  // G = (0, 1)
            const uint32_t M0 = al | ~ah | an, M1 = bl | ~bh | bn;
            bad |= __funnelshift_r(M0, M1, cfg.xs_sh[2][0]) | __funnelshift_r(M0, M1, cfg.xs_sh[2][1]);
            bad |= __funnelshift_r(M0, M1, cfg.xs_sh[2][2]) | __funnelshift_r(M0, M1, cfg.xs_sh[2][3]);
        }
        if (cfg.xs_has & 8u) {  // T = (1, 1)
            const uint32_t M0 = ~(al & ah) | an, M1 = ~(bl & bh) | bn;
            bad |= __funnelshift_r(M0, M1, cfg.xs_sh[3][0]) | __funnelshift_r(M0, M1, cfg.xs_sh[3][1]);
            bad |= __funnelshift_r(M0, M1, cfg.xs_sh[3][2]) | __funnelshift_r(M0, M1, cfg.xs_sh[3][3]);
        }
        uint32_t cand = ~bad;
        const int rem = nwin - (c << 5);
        if (rem < 32) cand &= (1u << rem) - 1u;
        while (cand) {
            const int s = __ffs(cand) - 1;
            cand &= cand - 1;
            const int o = (c << 5) + s;
            uint32_t e = 0;
#pragma unroll
            for (int k = 0; k < TW; k++) {
                const uint32_t wl = plane_bits<BC_DECODE_CHECKED>(lo, W, o + (k << 5));
                const uint32_t wh = plane_bits<BC_DECODE_CHECKED>(hi, W, o + (k << 5));
                const uint32_t wn = plane_bits<BC_DECODE_CHECKED>(nm, W, o + (k << 5));
                e |= (((wl ^ cfg.t_lo[k]) | (wh ^ cfg.t_hi[k])) & cfg.t_cm[k]) | (wn & (cfg.t_cm[k] | cfg.t_fn[k]));
            }
            if (e == 0) return o;  // offsets are visited in increasing order: the leftmost exact window
        }
        al = bl;
        ah = bh;
        an = bn;
    }
    return -1;
}

// Result of one repair chunk, packed: [29:20] smallest distance, [17:16] how many offsets reach it (saturating at 2),
// [15:0] one of them.
__device__ __forceinline__ uint32_t rep_pack(uint32_t d, uint32_t cnt, uint32_t arg) { return d << 20 | min(cnt, 2u) << 16 | arg; }

// Phase B, one work item: offsets [32 c, 32 c + 32) of the repair range [0, nwin - 1) of one read (Q3: the last offset
// is never scanned).  Every window is first looked at through ONE template word (the pivot: the word with the most
// constant bases): a window whose pivot word alone already has more than max_const_err mismatches cannot be within the
// cap, so only the survivors get the full-width count.  The pivot test is bit-sliced over the 32 offsets: for each
// constant position q of the pivot word, funnelshift(M, q) says which offsets mismatch there (N never mismatches), and
// carry-save adders accumulate five counter planes that start at 15 - max_const_err, so the top plane is "over the cap".
template <int TW>
__device__ __forceinline__ uint32_t repair_chunk(const DevCfg& cfg, const uint32_t* lo, const uint32_t* hi, const uint32_t* nm,
                                                 const uint32_t W, const int nwin, const int c) {
    const uint32_t maxc = cfg.max_const_err;
    const int rem = nwin - 1 - (c << 5);
    if (rem <= 0) return rep_pack(maxc + 1, 0, 0);
    const uint32_t kp = cfg.pivot;
    const uint32_t p_lo = cfg.t_lo[kp], p_hi = cfg.t_hi[kp], p_cm = cfg.t_cm[kp];
    const uint32_t j = c + kp;
    const uint32_t al = lo[j], bl = lo[j + 1];
    const uint32_t ah = hi[j], bh = hi[j + 1];
    const uint32_t an = nm[j], bn = nm[j + 1];
    uint32_t cand = 0;
    if (cfg.bs_two && rem > 8) {
        // over the constant positions of TWO template words (pivot, pivot + 1), whole blocks of four only: the third plane
        // word is at worst the next plane's first word (see plane_bits), its bits beyond every window
        const uint32_t cl = lo[j + 2], ch = hi[j + 2], cn = nm[j + 2];
        BsPlanes P{(cfg.bs_k & 1u) ? ~0u : 0u, (cfg.bs_k & 2u) ? ~0u : 0u, (cfg.bs_k & 4u) ? ~0u : 0u, (cfg.bs_k & 8u) ? ~0u : 0u, 0u};
        {
            const uint32_t A0 = (al | ah) & ~an, A1 = (bl | bh) & ~bn, A2 = (cl | ch) & ~cn;
            bs_blocks(P, A0, A1, cfg.bs2_n[0][0], cfg.bs2_sh[0][0]);
            bs_blocks(P, A1, A2, cfg.bs2_n[1][0], cfg.bs2_sh[1][0]);
        }
        {
            const uint32_t C0 = (~al | ah) & ~an, C1 = (~bl | bh) & ~bn, C2 = (~cl | ch) & ~cn;
            bs_blocks(P, C0, C1, cfg.bs2_n[0][1], cfg.bs2_sh[0][1]);
            bs_blocks(P, C1, C2, cfg.bs2_n[1][1], cfg.bs2_sh[1][1]);
        }
        {
            const uint32_t G0 = (al | ~ah) & ~an, G1 = (bl | ~bh) & ~bn, G2 = (cl | ~ch) & ~cn;
            bs_blocks(P, G0, G1, cfg.bs2_n[0][2], cfg.bs2_sh[0][2]);
            bs_blocks(P, G1, G2, cfg.bs2_n[1][2], cfg.bs2_sh[1][2]);
        }
        {
            const uint32_t T0 = ~(al & ah) & ~an, T1 = ~(bl & bh) & ~bn, T2 = ~(cl & ch) & ~cn;
            bs_blocks(P, T0, T1, cfg.bs2_n[0][3], cfg.bs2_sh[0][3]);
            bs_blocks(P, T1, T2, cfg.bs2_n[1][3], cfg.bs2_sh[1][3]);
        }
        cand = ~P.p16;
    } else if (cfg.bs_ok && rem > 8) {
        BsPlanes P{(cfg.bs_k & 1u) ? ~0u : 0u, (cfg.bs_k & 2u) ? ~0u : 0u, (cfg.bs_k & 4u) ? ~0u : 0u, (cfg.bs_k & 8u) ? ~0u : 0u, 0u};
        bs_base(P, (al | ah) & ~an, (bl | bh) & ~bn, cfg.pv_n[0], cfg.pv_sh4[0]);
        bs_base(P, (~al | ah) & ~an, (~bl | bh) & ~bn, cfg.pv_n[1], cfg.pv_sh4[1]);
        bs_base(P, (al | ~ah) & ~an, (bl | ~bh) & ~bn, cfg.pv_n[2], cfg.pv_sh4[2]);
        bs_base(P, ~(al & ah) & ~an, ~(bl & bh) & ~bn, cfg.pv_n[3], cfg.pv_sh4[3]);
        cand = ~P.p16;
    } else {  // a tail of at most 8 windows, or a cap above 15: window by window
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
            if (g < rem) {
#pragma unroll
                for (int s = g; s < g + 8; s++) {
                    const uint32_t wl = __funnelshift_r(al, bl, s);
                    const uint32_t wh = __funnelshift_r(ah, bh, s);
                    const uint32_t wn = __funnelshift_r(an, bn, s);
                    const uint32_t x = (((wl ^ p_lo) | (wh ^ p_hi)) & p_cm) & ~wn;
                    if ((uint32_t)__popc(x) <= maxc) cand |= 1u << s;
                }
            }
        }
    }
    if (rem < 32) cand &= (1u << rem) - 1u;
    uint32_t best = maxc + 1, cnt = 0, arg = 0;
    while (cand) {
        const int s = __ffs(cand) - 1;
        cand &= cand - 1;
        const int o = (c << 5) + s;
        uint32_t d = 0;
#pragma unroll
        for (int k = 0; k < TW; k++) {
            const uint32_t wl = plane_bits<BC_DECODE_CHECKED>(lo, W, o + (k << 5));
            const uint32_t wh = plane_bits<BC_DECODE_CHECKED>(hi, W, o + (k << 5));
            const uint32_t wn = plane_bits<BC_DECODE_CHECKED>(nm, W, o + (k << 5));
            d += __popc(((wl ^ cfg.t_lo[k]) | (wh ^ cfg.t_hi[k])) & cfg.t_cm[k] & ~wn);
        }
        if (d < best) {
            best = d;
            cnt = 1;
            arg = (uint32_t)o;
        } else if (d == best) {
            cnt++;
        }
    }
    return rep_pack(best, cnt, arg);
}

struct SlotBits {
    uint32_t lo, hi, nm;
};
template <bool CHECK>
__device__ __forceinline__ SlotBits slot_bits(const uint32_t* lo, const uint32_t* hi, const uint32_t* nm, uint32_t W,
                                              uint32_t pos, uint32_t len) {
    const uint32_t m = lenmask(len);
    SlotBits b;
    b.nm = plane_bits<CHECK>(nm, W, pos) & m;
    b.lo = plane_bits<CHECK>(lo, W, pos) & m & ~b.nm;
    b.hi = plane_bits<CHECK>(hi, W, pos) & m & ~b.nm;
    return b;
}
// reference index into its key field: schemes whose whole key fits 63 bits never touch the high word
__device__ __forceinline__ void key_or_index(Key& key, uint32_t idx, uint32_t shift, uint32_t wide) {
    if (!wide) key.lo |= (unsigned long long)idx << shift;
    else key_or(key, idx, shift);
}
__device__ __forceinline__ void key_raw(Key& key, const DevSlot& S, const SlotBits& b, uint32_t wide) {
    // raw key (N kept as its own symbol, Q14): field = [lo:len][hi:len][nm:len]
    if (!wide && 3u * S.len <= 64u) {  // the whole field in one 64-bit word, one shift
        const unsigned long long f = (unsigned long long)b.lo | (unsigned long long)b.hi << S.len | (unsigned long long)b.nm << (2u * S.len);
        key.lo |= f << S.key_shift;
        return;
    }
    key_or(key, b.lo, S.key_shift);
    key_or(key, b.hi, S.key_shift + S.len);
    key_or(key, b.nm, S.key_shift + 2 * S.len);
}

// K3 for one matched read that k_resolve finished: count it (info.rs:735-808) or fill its record slot
__device__ __forceinline__ int count_or_append(const Tables& tables, const RecOut& rec, unsigned long long read_index, int flags,
                                               Key key, bool* new_key, bool* new_pair) {
    if (flags & F_INSERT) return count_read(tables, key, new_key, new_pair) ? BC_ST_MATCHED : BC_ST_DUPLICATE;
    if (flags & F_APPEND) {  // deferred counting: the record slot of this read (k_decode left it empty)
        const unsigned long long pos = rec.base + read_index;
        rec.lo[pos] = key.lo;
        if (rec.hi) rec.hi[pos] = key.hi;
    }
    return BC_ST_MATCHED;
}

constexpr int kDeferred = -3;  // thread-local status: the read went to the deferred list (k_resolve finishes it)

// k_decode: one thread per read, kTile reads per CTA.  The tile's packed planes and lengths are contiguous in global
// memory and land in shared memory through two TMA bulk copies signalled on one mbarrier.  The quality bytes stay in
// global memory: a read needs only the ~34 bytes under its barcodes, and without the 152-byte rows the tile is 66 bytes
// per read instead of 218 (12 CTAs per SM instead of 8); every thread prefetches its row into L2 before the locate step.
// Shared memory: [planes kTile x plane_stride u32][read_len kTile u16][repair results batch.rep_chunks x kTile u32]
template <int TW>
__device__ __forceinline__ void decode_body(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const Tables& tables,
                                            unsigned long long* __restrict__ counters, const DecodeOut& out, const RecOut& rec,
                                            const Deferred& deferred, const int flags) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_nlist, s_maxwin;
    __shared__ uint16_t s_list[kTile];

    const uint32_t tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned long long base = (unsigned long long)blockIdx.x * kTile;
    const uint32_t n_tile = (uint32_t)min((unsigned long long)kTile, batch.n_reads - base);
    const uint32_t W = batch.W;
    uint32_t* s_pl = smem;
    uint16_t* s_len = reinterpret_cast<uint16_t*>(smem + kTile * batch.plane_stride);
    uint32_t* s_res = reinterpret_cast<uint32_t*>(s_len + kTile);  // kTile * 2 bytes keeps it 4-byte aligned

    const uint32_t* qrow = nullptr;
    if (batch.qual && tid < n_tile) {  // this read's quality row: on its way into L2 while the planes are staged and searched
        qrow = reinterpret_cast<const uint32_t*>(batch.qual + (base + tid) * batch.qual_stride);
        const char* p = reinterpret_cast<const char*>(qrow);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        if (batch.qual_stride > 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128));
    }
    {
        const uint32_t* g_pl = batch.planes + base * batch.plane_stride;
        const uint16_t* g_len = batch.read_len + base;
        const uint32_t b_pl = n_tile * batch.plane_stride * 4u, b_len = n_tile * 2u;
        const bool bulk = (((b_pl | b_len) & 15u) == 0) && ((((unsigned long long)g_pl | (unsigned long long)g_len) & 15ull) == 0);
        if (tid == 0) {
            s_nlist = 0;
            s_maxwin = 0;
        }
        if (bulk) {  // uniform per CTA
            if (tid == 0) mbar_init(&s_bar, 1);
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(&s_bar, b_pl + b_len);
                bulk_g2s(s_pl, g_pl, b_pl, &s_bar);
                bulk_g2s(s_len, g_len, b_len, &s_bar);
            }
            mbar_wait(&s_bar, 0);
        } else {  // ragged last tile / unaligned caller buffers: plain cooperative copy
            for (uint32_t i = tid; i < n_tile * batch.plane_stride; i += kTile) s_pl[i] = __ldg(g_pl + i);
            if (tid < n_tile) s_len[tid] = g_len[tid];
            __syncthreads();
        }
    }

    int status = -1;  // -1: thread has no read
    bool new_key = false, new_pair = false;
    int off = -1;
    bool repaired = false;
    Key key{0, 0};
    const uint32_t* lo = s_pl + tid * batch.plane_stride;
    const uint32_t* hi = lo + W;
    const uint32_t* nm = hi + W;

    // ---- K1, phase A: leftmost exact window of this thread's read
    bool need_repair = false;
    if (tid < n_tile) {
        const uint32_t rl = s_len[tid];
        if (rl & BC_READ_UNSUPPORTED) {
            status = BC_ST_UNSUPPORTED;
        } else {
            const int nwin = (int)(rl & 0x7FFF) - (int)cfg.L + 1;  // <= 0: read shorter than the scheme (Q4) -> constant-region error
            status = BC_ST_CONSTANT;
            if (nwin > 0) {
                off = locate_exact<TW>(cfg, lo, hi, nm, W, nwin);
                if (off >= 0) status = BC_ST_MATCHED;
                else need_repair = nwin > 1;  // the repair range [0, nwin - 1) is not empty
            }
        }
    }
    // ---- K1, phase B: the reads without an exact window, as (read, chunk) items over all threads of the CTA
    uint32_t my_li = 0;
    {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, need_repair);
        if (m) {
            const int leader = __ffs(m) - 1;
            uint32_t at = 0;
            if (lane == leader) at = atomicAdd(&s_nlist, (uint32_t)__popc(m));
            at = __shfl_sync(0xFFFFFFFFu, at, leader);
            if (need_repair) {
                my_li = at + __popc(m & ((1u << lane) - 1u));
                s_list[my_li] = (uint16_t)tid;
                atomicMax(&s_maxwin, (uint32_t)((int)(s_len[tid] & 0x7FFF) - (int)cfg.L));  // size of this read's repair range
            }
        }
    }
    __syncthreads();
    const uint32_t n_list = s_nlist;
    if (n_list) {  // uniform per CTA
        const uint32_t n_ch = min((s_maxwin + 31u) >> 5, batch.rep_chunks);  // equal for well-formed batches (read_len <= 32 W)
        for (uint32_t it = tid; it < n_list * n_ch; it += kTile) {
            const uint32_t ch = it / n_list, li = it - ch * n_list;
            const uint32_t r = s_list[li];
            const uint32_t* rlo = s_pl + r * batch.plane_stride;
            const int nwin = (int)(s_len[r] & 0x7FFF) - (int)cfg.L + 1;
            s_res[ch * kTile + li] = repair_chunk<TW>(cfg, rlo, rlo + W, rlo + 2 * W, W, nwin, (int)ch);
        }
        __syncthreads();
        if (need_repair) {  // unique minimum over the chunks, within the cap (Q5)
            uint32_t best = cfg.max_const_err + 1, cnt = 0, arg = 0;
            for (uint32_t ch = 0; ch < n_ch; ch++) {
                const uint32_t p = s_res[ch * kTile + my_li];
                const uint32_t d = p >> 20, c = (p >> 16) & 3u;
                if (d < best) {
                    best = d;
                    cnt = c;
                    arg = p & 0xFFFFu;
                } else if (d == best) {
                    cnt += c;
                }
            }
            if (cnt == 1 && best <= cfg.max_const_err) {
                off = (int)arg;
                repaired = true;
                status = BC_ST_MATCHED;
                if (cfg.has_fn) {  // the regex is re-run on the repaired window: format-N still needs ACGT
                    uint32_t bad = 0;
#pragma unroll
                    for (int k = 0; k < TW; k++) bad |= plane_bits<BC_DECODE_CHECKED>(nm, W, off + (k << 5)) & cfg.t_fn[k];
                    if (bad) {
                        off = -1;
                        repaired = false;
                        status = BC_ST_CONSTANT;
                    }
                }
            }
        }
    }

    if (status == BC_ST_MATCHED && !(flags & F_LOCATE_ONLY)) {
        // ---- K2a: per-barcode average quality (parse.rs:331-375); runs and thresholds precomputed (Q8, Q12).
        // Q6: after a repair the quality string is read from 0, not from the repaired offset.
        // Byte sums with whole-word loads: the first and last word of a run are masked down to the bytes that belong to
        // it, dp4a against 0x01010101 adds the four bytes of a word.  The packer guarantees every byte >= 33 ('!'); the
        // threshold already includes that offset.  No early exit between runs: their loads overlap.
        bool lowq = false;
        if (cfg.n_qruns) {
            const uint32_t q0 = repaired ? 0u : (uint32_t)off;
            BC_UNROLL
            for (uint32_t r = 0; r < cfg.n_qruns; r++) {
                const uint32_t a = q0 + cfg.qruns[r].off, e1 = a + cfg.qruns[r].len - 1u;
                const uint32_t wa = a >> 2, wb = e1 >> 2;
                const uint32_t ma = 0xFFFFFFFFu << ((a & 3u) << 3), mb = 0xFFFFFFFFu >> ((3u - (e1 & 3u)) << 3);
                uint32_t sum;
                if (wa == wb) {
                    sum = __dp4a(__ldg(qrow + wa) & ma & mb, 0x01010101u, 0u);
                } else {
                    const uint32_t first = __ldg(qrow + wa), last = __ldg(qrow + wb);
                    sum = __dp4a(first & ma, 0x01010101u, 0u);
#pragma unroll 1
                    for (uint32_t k = wa + 1; k < wb; k++) sum = __dp4a(__ldg(qrow + k), 0x01010101u, sum);
                    sum = __dp4a(last & mb, 0x01010101u, sum);
                }
                lowq |= sum < cfg.qruns[r].thresh;
            }
        }
        // ---- K2b: barcode correction, sample first then counted barcodes in order (parse.rs:448-507).
        // Fast paths only: anything that needs a search over the reference set goes to k_resolve.
        if (lowq) {
            status = BC_ST_LOW_QUALITY;
        } else {
            BC_UNROLL
            for (uint32_t oi = 0; oi < cfg.n_slots; oi++) {
                const uint32_t si = cfg.order[oi];
                const DevSlot& S = cfg.slots[si];
                const SlotBits b = slot_bits<BC_DECODE_CHECKED>(lo, hi, nm, W, off + S.offset, S.len);
                if (S.mode == MODE_RAW) {
                    key_raw(key, S, b, cfg.wide);
                    continue;
                }
                uint32_t idx = kFail;
                bool defer = false;
                if (S.mode == MODE_TABLE) {
                    if (b.nm == 0) idx = table_pick(__ldg(&aux.tables[S.aux_off + (b.lo | (b.hi << S.len))]), S.max_err);
                    else if (S.n_inline && (b.nm & (b.nm - 1)) == 0) idx = table_lookup_1n(aux.tables + S.aux_off, S, b.lo, b.hi, b.nm);
                    else defer = true;  // two or more N in one barcode: k_resolve
                } else if (S.mode == MODE_HASH) {
                    if (b.nm == 0) idx = hash_exact(aux, S, b.lo, b.hi);
                    if (idx == kFail) {
                        if (!S.has_half || __popc(b.nm) > 1) defer = true;
                        else defer = half_probe(aux, S, b.lo, b.hi, b.nm, &idx) == HALF_DEEPER;
                    }
                } else {
                    defer = true;
                }
                if (defer) {
                    status = kDeferred;
                    break;
                }
                if (out.slot_index) out.slot_index[(base + tid) * cfg.n_slots + si] = (int32_t)idx;
                if (idx == kFail) {
                    status = S.kind == 'S' ? BC_ST_SAMPLE : BC_ST_COUNTED;
                    break;
                }
                key_or_index(key, idx, S.key_shift, cfg.wide);
            }
        }
        if (status == BC_ST_MATCHED && (flags & F_INSERT))
            status = count_read(tables, key, &new_key, &new_pair) ? BC_ST_MATCHED : BC_ST_DUPLICATE;
    }
    if (tid < n_tile) {
        if (flags & F_APPEND) {
            // deferred counting: every read owns one record slot; unmatched reads (and reads handed to k_resolve, which
            // fills the slot itself when the read matches) leave a hole.  Consecutive lanes, consecutive slots.
            const unsigned long long pos = rec.base + base + tid;
            const bool m = status == BC_ST_MATCHED;
            rec.lo[pos] = m ? key.lo : kEmpty;
            if (rec.hi) rec.hi[pos] = m ? key.hi : kEmpty;
        }
        if (status != kDeferred && (flags & F_EMIT)) {
            const unsigned long long i = base + tid;
            if (out.status) out.status[i] = (uint8_t)status;
            if (out.offset) out.offset[i] = (int16_t)off;
            if (out.repaired) out.repaired[i] = repaired ? 1 : 0;
            if (out.key_lo) out.key_lo[i] = key.lo;
            if (out.key_hi) out.key_hi[i] = key.hi;
        }
    }

    // ---- deferred reads: one warp-aggregated append per warp
    {
        const unsigned dm = __ballot_sync(0xFFFFFFFFu, status == kDeferred);
        if (dm) {
            const int leader = __ffs(dm) - 1;
            uint32_t at = 0;
            if (lane == leader) at = atomicAdd(deferred.count, (uint32_t)__popc(dm));
            at = __shfl_sync(0xFFFFFFFFu, at, leader);
            if (status == kDeferred)
                deferred.items[at + __popc(dm & ((1u << lane) - 1u))] =
                    make_uint2((uint32_t)(base + tid), (uint32_t)off | (repaired ? 0x10000u : 0u));
        }
    }
    // ---- outcome counters (info.rs:60-127): every lane contributes a 1 in its outcome's 8-bit field, two warp-wide
    // REDUX sums, then lanes 0..8 each add one counter with a single fire-and-forget RED — no shared memory and no CTA
    // barrier at the end.  The adds go to one of kCounterStripes copies (by CTA) so that no address sees more than a
    // few thousand of them per launch; k_fold_counters sums the copies into the context's counters right after.
    if (counters) {
        const uint32_t fa = (status >= 0 && status < 4) ? 1u << (8 * status) : 0u;
        const uint32_t fb = (status >= 4 && status < 7 ? 1u << (8 * (status - 4)) : 0u) | (new_key ? 1u << 24 : 0u);
        const uint32_t sa = __reduce_add_sync(0xFFFFFFFFu, fa), sb = __reduce_add_sync(0xFFFFFFFFu, fb);
        const bool inline_set = (flags & F_INSERT) && tables.has_set;
        const uint32_t sc = inline_set ? __reduce_add_sync(0xFFFFFFFFu, new_pair ? 1u : 0u) : 0u;
        // lane f < 7: outcome f (status order -> counter order, one nibble each); lane 7 / 8: new map / set entries
        const uint32_t f = lane;
        const uint32_t v = f < 4 ? (sa >> (8 * f)) & 0xFFu : f < 7 ? (sb >> (8 * (f - 4))) & 0xFFu : f == 7 ? sb >> 24 : f == 8 ? sc : 0u;
        const uint32_t dst = f < 7 ? (0x6325140u >> (4 * f)) & 0xFu : BC_N_COUNTERS + (f - 7);
        static_assert(BC_CNT_MATCHED == 0 && BC_CNT_DUPLICATES == 4 && BC_CNT_CONSTANT == 1 && BC_CNT_LOW_QUALITY == 5 &&
                          BC_CNT_SAMPLE == 2 && BC_CNT_COUNTED == 3 && BC_CNT_UNSUPPORTED == 6, "counter order");
        if (f < 9 && v) atomicAdd(&counters[(blockIdx.x % kCounterStripes) * kCounterStride + dst], (unsigned long long)v);
    }
}

}  // namespace bc
