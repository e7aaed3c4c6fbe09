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

namespace bc {

__device__ __forceinline__ uint32_t lenmask(uint32_t len) { return len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u); }

// bits [pos, pos+32) of a W-word bit plane
__device__ __forceinline__ uint32_t plane_bits(const uint32_t* p, uint32_t W, uint32_t pos) {
    uint32_t j = pos >> 5, s = pos & 31;
    uint32_t a = j < W ? p[j] : 0u;
    uint32_t b = (j + 1) < W ? p[j + 1] : 0u;
    return __funnelshift_r(a, b, s);
}

// parse.rs:553-593 over a slot's reference set, with the exact-membership short cut of parse.rs:457,489 folded
// in: an identical reference wins outright; otherwise the unique minimum within max_err, compared over the
// shorter of the two lengths, N on either side never counting (Q5, Q10).
__device__ __forceinline__ uint32_t scan_refs(const uint4* __restrict__ refs, uint32_t n_ref, uint32_t blo, uint32_t bhi,
                                              uint32_t bnm, uint32_t len, uint32_t max_err) {
    uint32_t best = max_err + 1, cnt = 0, arg = kFail, exact = kFail;
    const uint32_t lm = lenmask(len);
    for (uint32_t i = 0; i < n_ref; i++) {
        uint4 r = __ldg(&refs[i]);
        uint32_t m = r.w < len ? lenmask(r.w) : lm;
        uint32_t d = __popc(((blo ^ r.x) | (bhi ^ r.y)) & ~bnm & ~r.z & m);
        if (r.w == len && r.x == blo && r.y == bhi && r.z == bnm) exact = i;
        if (d < best) {
            best = d;
            cnt = 1;
            arg = i;
        } else if (d == best) {
            cnt++;
        }
    }
    if (exact != kFail) return exact;
    return (cnt == 1 && best <= max_err) ? arg : kFail;
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

template <int TW>
__global__ void __launch_bounds__(kTile) k_decode(const __grid_constant__ DevCfg cfg, const BatchView batch,
                                                  const DevAux aux, const DevTable table,
                                                  unsigned long long* __restrict__ counters, const DecodeOut out,
                                                  const RouteOut route, const int flags) {
    extern __shared__ uint32_t smem[];
    __shared__ unsigned int s_cnt[BC_N_COUNTERS + 1];

    const uint32_t tid = threadIdx.x;
    const unsigned long long base = (unsigned long long)blockIdx.x * kTile;
    const uint32_t n_tile = min((unsigned long long)kTile, batch.n_reads - base);
    const uint32_t W = batch.W;
    uint32_t* s_pl = smem;
    uint8_t* s_q = reinterpret_cast<uint8_t*>(smem + kTile * batch.plane_stride);

    if (tid < BC_N_COUNTERS + 1) s_cnt[tid] = 0;
    // stage the tile: both arrays are contiguous per tile, so this is a straight coalesced copy
    {
        const uint32_t* g = batch.planes + base * batch.plane_stride;
        const uint32_t nw = n_tile * batch.plane_stride;
        for (uint32_t i = tid; i < nw; i += kTile) s_pl[i] = __ldg(g + i);
        if (batch.qual) {
            const uint32_t* gq = reinterpret_cast<const uint32_t*>(batch.qual + base * batch.qual_stride);
            uint32_t* sq = reinterpret_cast<uint32_t*>(s_q);
            const uint32_t nq = n_tile * (batch.qual_stride >> 2);
            for (uint32_t i = tid; i < nq; i += kTile) sq[i] = __ldg(gq + i);
        }
    }
    __syncthreads();

    int status = -1;  // -1: thread has no read
    bool is_new = false;
    if (tid < n_tile) {
        const uint32_t* lo = s_pl + tid * batch.plane_stride;
        const uint32_t* hi = lo + W;
        const uint32_t* nm = hi + W;
        const uint32_t rl = batch.read_len[base + tid];
        const int len = rl & 0x7FFF;
        const int L = cfg.L;
        int off = -1;
        bool repaired = false;
        Key key{0, 0};

        if (rl & BC_READ_UNSUPPORTED) {
            status = BC_ST_UNSUPPORTED;
        } else {
            // ---- K1: locate.  One pass over the windows computes both predicates (Q1): the exact test of the
            // regex (a read N in a constant fails, format-N needs ACGT) and the masked Hamming distance of the
            // repair (N on either side is a wildcard).  Leftmost exact window wins (P1); otherwise the unique
            // minimum over offsets [0, R-L) within the cap (P2, Q3, Q5).
            const int nwin = len - L + 1;  // <= 0: read shorter than the scheme (Q4) -> constant-region error
            uint32_t best = cfg.max_const_err + 1, cnt = 0;
            int arg = -1, first_exact = -1;
            const int nchunks = (nwin + 31) >> 5;
            for (int c = 0; c < nchunks && first_exact < 0; c++) {
                uint32_t pl[TW + 1], ph[TW + 1], pn[TW + 1];
#pragma unroll
                for (int k = 0; k <= TW; k++) {
                    const bool in = (uint32_t)(c + k) < W;
                    pl[k] = in ? lo[c + k] : 0u;
                    ph[k] = in ? hi[c + k] : 0u;
                    pn[k] = in ? nm[c + k] : 0u;
                }
                const int smax = min(32, nwin - (c << 5));
                for (int s = 0; s < smax; s++) {
                    uint32_t d = 0, e = 0;
#pragma unroll
                    for (int k = 0; k < TW; k++) {
                        const uint32_t wl = __funnelshift_r(pl[k], pl[k + 1], s);
                        const uint32_t wh = __funnelshift_r(ph[k], ph[k + 1], s);
                        const uint32_t wn = __funnelshift_r(pn[k], pn[k + 1], s);
                        const uint32_t x = ((wl ^ cfg.t_lo[k]) | (wh ^ cfg.t_hi[k])) & cfg.t_cm[k];
                        e |= x | (wn & (cfg.t_cm[k] | cfg.t_fn[k]));
                        d += __popc(x & ~wn);
                    }
                    const int o = (c << 5) + s;
                    if (e == 0) {
                        first_exact = o;
                        break;
                    }
                    if (o < nwin - 1) {
                        if (d < best) {
                            best = d;
                            cnt = 1;
                            arg = o;
                        } else if (d == best) {
                            cnt++;
                        }
                    }
                }
            }
            int qstart = 0;
            if (first_exact >= 0) {
                off = first_exact;
                qstart = off;
            } else if (cnt == 1 && best <= cfg.max_const_err) {
                off = arg;
                repaired = true;
                qstart = 0;  // Q6: the repaired sequence starts at 0, the quality string is not re-aligned
                if (cfg.has_fn) {  // the regex is re-run on the repaired window: format-N still needs ACGT
                    uint32_t bad = 0;
#pragma unroll
                    for (int k = 0; k < TW; k++) bad |= plane_bits(nm, W, off + (k << 5)) & cfg.t_fn[k];
                    if (bad) off = -1;
                }
            }
            if (off < 0) {
                status = BC_ST_CONSTANT;
                repaired = false;
            } else if (flags & F_LOCATE_ONLY) {
                status = BC_ST_MATCHED;
            } else {
                status = BC_ST_MATCHED;
                // ---- K2a: per-barcode average quality (parse.rs:331-375); runs and thresholds precomputed (Q8, Q12)
                if (cfg.n_qruns) {
                    const uint8_t* q = s_q + tid * batch.qual_stride + qstart;
                    for (uint32_t r = 0; r < cfg.n_qruns; r++) {
                        const DevQRun run = cfg.qruns[r];
                        uint32_t sum = 0;
                        for (uint32_t i = 0; i < run.len; i++) sum += (uint8_t)(q[run.off + i] - 33);
                        if (sum < run.thresh) {
                            status = BC_ST_LOW_QUALITY;
                            break;
                        }
                    }
                }
                // ---- K2b: barcode correction, sample first then counted barcodes in order (parse.rs:448-507)
                if (status == BC_ST_MATCHED) {
                    for (uint32_t oi = 0; oi < cfg.n_slots; oi++) {
                        const uint32_t si = cfg.order[oi];
                        const DevSlot& S = cfg.slots[si];
                        const uint32_t pos = off + S.offset;
                        const uint32_t m = lenmask(S.len);
                        const uint32_t bnm = plane_bits(nm, W, pos) & m;
                        const uint32_t blo = plane_bits(lo, W, pos) & m & ~bnm;
                        const uint32_t bhi = plane_bits(hi, W, pos) & m & ~bnm;
                        if (S.mode == MODE_RAW) {
                            // raw key (N kept as its own symbol, Q14): field = [lo:len][hi:len][nm:len]
                            key_or(key, blo, S.key_shift);
                            key_or(key, bhi, S.key_shift + S.len);
                            key_or(key, bnm, S.key_shift + 2 * S.len);
                            if (out.slot_index) out.slot_index[(base + tid) * cfg.n_slots + si] = -1;
                            continue;
                        }
                        uint32_t idx = kFail;
                        if (bnm == 0 && S.mode == MODE_TABLE) {
                            const uint32_t v = __ldg(&aux.tables[S.aux_off + (blo | (bhi << S.len))]);
                            idx = v == 0xFFFFu ? kFail : v;
                        } else {
                            if (bnm == 0 && S.mode == MODE_HASH) idx = hash_exact(aux, S, blo, bhi);
                            if (idx == kFail) idx = scan_refs(aux.refs + S.ref_off, S.n_ref, blo, bhi, bnm, S.len, S.max_err);
                        }
                        if (out.slot_index) out.slot_index[(base + tid) * cfg.n_slots + si] = (int32_t)idx;
                        if (idx == kFail) {
                            status = S.kind == 'S' ? BC_ST_SAMPLE : BC_ST_COUNTED;
                            break;
                        }
                        key_or(key, idx, S.key_shift);
                    }
                }
                // ---- K3: count (info.rs:735-808)
                if (status == BC_ST_MATCHED) {
                    if (flags & F_INSERT) {
                        if (!table_count(table, key, 1ULL, &is_new)) status = BC_ST_DUPLICATE;
                    } else if (flags & F_ROUTE) {
                        const uint32_t owner = (uint32_t)(hash_key(key_shr(key, cfg.umi_bits)) % route.n_ranks);
                        const uint32_t slot = atomicAdd(&route.counts[owner], 1u);
                        if (slot < route.capacity) route.buckets[owner * route.capacity + slot] = key;
                        status = -2;  // outcome is decided by the owner rank
                    }
                }
            }
        }
        if (flags & F_EMIT) {
            const unsigned long long i = base + tid;
            if (out.status) out.status[i] = (uint8_t)status;
            if (out.offset) out.offset[i] = (int16_t)off;
            if (out.repaired) out.repaired[i] = repaired ? 1 : 0;
            if (out.key_lo) out.key_lo[i] = key.lo;
            if (out.key_hi) out.key_hi[i] = key.hi;
        }
    }

    // ---- outcome counters (info.rs:60-127): warp-aggregated, one global atomic per counter per CTA
    if (counters) {
        const int lane = tid & 31;
#pragma unroll
        for (int st = 0; st < BC_N_COUNTERS; st++) {
            const unsigned b = __ballot_sync(0xFFFFFFFFu, status == st);
            if (lane == 0 && b) atomicAdd(&s_cnt[st], __popc(b));
        }
        const unsigned bn = __ballot_sync(0xFFFFFFFFu, is_new);
        if (lane == 0 && bn) atomicAdd(&s_cnt[BC_N_COUNTERS], __popc(bn));
        __syncthreads();
        // status order -> counter order
        if (tid < BC_N_COUNTERS) {
            const int map[BC_N_COUNTERS] = {BC_CNT_MATCHED, BC_CNT_DUPLICATES, BC_CNT_CONSTANT, BC_CNT_LOW_QUALITY,
                                            BC_CNT_SAMPLE,  BC_CNT_COUNTED,    BC_CNT_UNSUPPORTED};
            if (s_cnt[tid]) atomicAdd(&counters[map[tid]], (unsigned long long)s_cnt[tid]);
        }
        if (tid == BC_N_COUNTERS && s_cnt[BC_N_COUNTERS] && table.n_entries)
            atomicAdd(table.n_entries, (unsigned long long)s_cnt[BC_N_COUNTERS]);
    }
}

size_t decode_smem_bytes(const BatchView& b) { return (size_t)kTile * (b.plane_stride * 4u + (b.qual ? b.qual_stride : 0u)); }

template <int TW>
static cudaError_t launch_decode_tw(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const DevTable& table,
                                    unsigned long long* counters, const DecodeOut& out, const RouteOut& route, int flags,
                                    cudaStream_t stream) {
    const size_t smem = decode_smem_bytes(batch);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_decode<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const unsigned grid = (batch.n_reads + kTile - 1) / kTile;
    k_decode<TW><<<grid, kTile, smem, stream>>>(cfg, batch, aux, table, counters, out, route, flags);
    return cudaGetLastError();
}

cudaError_t launch_decode(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const DevTable& table,
                          unsigned long long* counters, const DecodeOut& out, const RouteOut& route, int flags,
                          cudaStream_t stream) {
    if (batch.n_reads == 0) return cudaSuccess;
    switch (cfg.TW) {
        case 1: return launch_decode_tw<1>(cfg, batch, aux, table, counters, out, route, flags, stream);
        case 2: return launch_decode_tw<2>(cfg, batch, aux, table, counters, out, route, flags, stream);
        case 3: return launch_decode_tw<3>(cfg, batch, aux, table, counters, out, route, flags, stream);
        case 4: return launch_decode_tw<4>(cfg, batch, aux, table, counters, out, route, flags, stream);
        case 5: return launch_decode_tw<5>(cfg, batch, aux, table, counters, out, route, flags, stream);
        case 6: return launch_decode_tw<6>(cfg, batch, aux, table, counters, out, route, flags, stream);
        case 7: return launch_decode_tw<7>(cfg, batch, aux, table, counters, out, route, flags, stream);
        case 8: return launch_decode_tw<8>(cfg, batch, aux, table, counters, out, route, flags, stream);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------------------------------------
// MODE_TABLE: result of the correction for every N-free barcode value, so the hot kernel does one lookup.
__global__ void k_build_table(const DevSlot slot, const uint4* __restrict__ refs, uint16_t* __restrict__ table) {
    const uint32_t n = 1u << (2 * slot.len);
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t m = lenmask(slot.len);
    const uint32_t idx = scan_refs(refs + slot.ref_off, slot.n_ref, v & m, (v >> slot.len) & m, 0u, slot.len, slot.max_err);
    table[v] = idx == kFail ? (uint16_t)0xFFFFu : (uint16_t)idx;
}

cudaError_t launch_build_table(const DevSlot& slot, const DevAux& aux, uint16_t* table, cudaStream_t stream) {
    const uint32_t n = 1u << (2 * slot.len);
    k_build_table<<<(n + 255) / 256, 256, 0, stream>>>(slot, aux.refs, table);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
__global__ void k_insert(const DevTable table, const unsigned long long* __restrict__ key_lo,
                         const unsigned long long* __restrict__ key_hi, const Key* __restrict__ records,
                         const unsigned long long* __restrict__ counts, const unsigned long long n,
                         unsigned long long* __restrict__ counters) {
    unsigned long long matched = 0, dup = 0, fresh = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        Key k;
        if (records) {
            k = records[i];
        } else {
            k.lo = key_lo[i];
            k.hi = key_hi ? key_hi[i] : 0ULL;
        }
        bool is_new;
        if (table_count(table, k, counts ? counts[i] : 1ULL, &is_new)) matched++;
        else dup++;
        if (is_new) fresh++;
    }
    // warp-reduce, then one atomic per warp
    for (int o = 16; o; o >>= 1) {
        matched += __shfl_xor_sync(0xFFFFFFFFu, matched, o);
        dup += __shfl_xor_sync(0xFFFFFFFFu, dup, o);
        fresh += __shfl_xor_sync(0xFFFFFFFFu, fresh, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (counters) {
            if (matched) atomicAdd(&counters[BC_CNT_MATCHED], matched);
            if (dup) atomicAdd(&counters[BC_CNT_DUPLICATES], dup);
        }
        if (fresh && table.n_entries) atomicAdd(table.n_entries, fresh);
    }
}

static unsigned grid_for(unsigned long long n, unsigned block) {
    unsigned long long g = (n + block - 1) / block;
    const unsigned long long cap = 148ULL * 16;  // grid-stride: a few waves over the 148 SMs
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

cudaError_t launch_insert(const DevTable& table, const unsigned long long* key_lo, const unsigned long long* key_hi,
                          const Key* records, const unsigned long long* counts, unsigned long long n,
                          unsigned long long* counters, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_insert<<<grid_for(n, 256), 256, 0, stream>>>(table, key_lo, key_hi, records, counts, n, counters);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool table_entry(const DevTable& t, unsigned long long i, Key* k) {
    if (t.wide) {
        ulonglong2 v = t.keys128[i];
        k->lo = v.x;
        k->hi = v.y;
        return !(v.x == kEmpty && v.y == kEmpty);
    }
    k->lo = t.keys64[i];
    k->hi = 0;
    return k->lo != kEmpty;
}

__global__ void k_group(const DevTable set, const uint32_t umi_bits, const DevTable dst) {
    const unsigned long long cap = set.cap_mask + 1;
    unsigned long long fresh = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cap;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        Key k;
        if (!table_entry(set, i, &k)) continue;
        bool is_new;
        table_count(dst, key_shr(k, umi_bits), 1ULL, &is_new);
        if (is_new) fresh++;
    }
    for (int o = 16; o; o >>= 1) fresh += __shfl_xor_sync(0xFFFFFFFFu, fresh, o);
    if ((threadIdx.x & 31) == 0 && fresh && dst.n_entries) atomicAdd(dst.n_entries, fresh);
}

cudaError_t launch_group(const DevTable& set, uint32_t umi_bits, const DevTable& dst, cudaStream_t stream) {
    k_group<<<grid_for(set.cap_mask + 1, 256), 256, 0, stream>>>(set, umi_bits, dst);
    return cudaGetLastError();
}

__global__ void k_compact(const DevTable t, unsigned long long* __restrict__ key_lo, unsigned long long* __restrict__ key_hi,
                          unsigned long long* __restrict__ count, unsigned long long* __restrict__ n_rows) {
    const unsigned long long cap = t.kind == 0 ? t.cap_mask : t.cap_mask + 1;
    const int lane = threadIdx.x & 31;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long start = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    // every lane of a warp runs the same number of iterations so the ballot below is convergent
    const unsigned long long iters = (cap + stride - 1) / stride;
    for (unsigned long long it = 0; it < iters; it++) {
        const unsigned long long i = start + it * stride;
        Key k{0, 0};
        unsigned long long c = 0;
        bool have = false;
        if (i < cap) {
            if (t.kind == 0) {
                c = t.counts[i];
                k.lo = i;
                have = c != 0;
            } else {
                have = table_entry(t, i, &k);
                if (have) c = t.counts ? t.counts[i] : 1ULL;
            }
        }
        const unsigned b = __ballot_sync(0xFFFFFFFFu, have);
        if (b) {
            unsigned long long basepos = 0;
            if (lane == 0) basepos = atomicAdd(n_rows, (unsigned long long)__popc(b));
            basepos = __shfl_sync(0xFFFFFFFFu, basepos, 0);
            if (have) {
                const unsigned long long p = basepos + __popc(b & ((1u << lane) - 1u));
                key_lo[p] = k.lo;
                key_hi[p] = k.hi;
                count[p] = c;
            }
        }
    }
}

cudaError_t launch_compact(const DevTable& t, unsigned long long* key_lo, unsigned long long* key_hi,
                           unsigned long long* count, unsigned long long* n_rows, cudaStream_t stream) {
    const unsigned long long cap = t.kind == 0 ? t.cap_mask : t.cap_mask + 1;
    k_compact<<<grid_for(cap, 256), 256, 0, stream>>>(t, key_lo, key_hi, count, n_rows);
    return cudaGetLastError();
}

__global__ void k_marginal(const unsigned long long* __restrict__ key_lo, const unsigned long long* __restrict__ key_hi,
                           const unsigned long long* __restrict__ count, const unsigned long long n_rows, const Key mask,
                           const DevTable dst) {
    unsigned long long fresh = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_rows;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        Key k{key_lo[i] & mask.lo, key_hi[i] & mask.hi};
        bool is_new;
        table_count(dst, k, count[i], &is_new);
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

__global__ void k_rehash(const DevTable src, const DevTable dst) {
    const unsigned long long cap = src.cap_mask + 1;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cap;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        Key k;
        if (!table_entry(src, i, &k)) continue;
        bool is_new;
        table_count(dst, k, src.counts ? src.counts[i] : 1ULL, &is_new);
    }
}

cudaError_t launch_rehash(const DevTable& src, const DevTable& dst, cudaStream_t stream) {
    k_rehash<<<grid_for(src.cap_mask + 1, 256), 256, 0, stream>>>(src, dst);
    return cudaGetLastError();
}

}  // namespace bc
