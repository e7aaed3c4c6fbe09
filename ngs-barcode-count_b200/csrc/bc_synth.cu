// bc_synth.cu — synthetic-read generator (bench / test tool; see bc_synth.h).  One function, gen_read(), compiled
// for host and device, integer arithmetic only, so both sides produce identical reads.
#include "bc_synth.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

#define HD __host__ __device__ __forceinline__

HD uint64_t splitmix(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ULL;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
HD uint64_t hash64(uint64_t x) {
    uint64_t s = x;
    return splitmix(s);
}
HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
// skewed fraction in [0, 2^32): u, u^2 or u^3 (mass piles up near 0 => a few very abundant ids)
HD uint32_t skewed(uint32_t u, uint8_t skew) {
    uint64_t v = u;
    if (skew >= 1) v = (v * u) >> 32;
    if (skew >= 2) v = (v * u) >> 32;
    return (uint32_t)v;
}

// codes: 0..3 = A C G T, 4 = N.  qual: Phred+33 characters.
HD void gen_read(const bcs_config& c, const uint8_t* refs, uint64_t i, uint8_t* codes, uint8_t* qual) {
    uint64_t s = c.seed ^ (i * 0xD1B54A32D192ED03ULL + 0x2545F4914F6CDD1DULL);
    const uint32_t R = c.read_len, L = c.template_len;
    // random background
    for (uint32_t p = 0; p < R; p += 32) {
        uint64_t r = splitmix(s);
        for (uint32_t k = 0; k < 32 && p + k < R; k++, r >>= 2) codes[p + k] = (uint8_t)(r & 3);
    }
    const uint64_t r0 = splitmix(s);
    const bool junk = (uint32_t)r0 < c.p_junk;
    const bool lowq = (uint32_t)(r0 >> 32) < c.p_lowq;
    const uint64_t r1 = splitmix(s);
    if (!junk) {
        const uint32_t start = (uint32_t)(((r1 & 0xFFFFFFFFu) * (uint64_t)(c.max_start + 1)) >> 32);
        // which slots are barcodes, to skip them when laying down the constants
        for (uint32_t p = 0; p < L; p++) codes[start + p] = c.template_codes[p] < 4 ? c.template_codes[p] : codes[start + p];
        uint64_t mol = 0;
        bool use_mol = c.molecule_pool != 0;
        uint64_t compound = 0;
        if (use_mol) {
            mol = mulhi64(splitmix(s), c.molecule_pool);
            compound = hash64(mol ^ 0xA5A5A5A55A5A5A5AULL);
            const uint64_t r2 = splitmix(s);
            if ((uint32_t)r2 < c.p_enriched && c.n_enriched)
                compound = hash64(0xE7E7E7E7ULL + (((r2 >> 32) * (uint64_t)c.n_enriched) >> 32));
        }
        for (uint32_t k = 0; k < c.n_slots; k++) {
            const bcs_slot& S = c.slots[k];
            const uint64_t rs = splitmix(s);
            uint8_t* dst = codes + start + S.offset;
            if (S.n_ref) {
                uint32_t u = use_mol && S.kind != 'R' ? (uint32_t)hash64(compound + 0x1000193ULL * (k + 1)) : (uint32_t)rs;
                const uint32_t idx = (uint32_t)(((uint64_t)skewed(u, S.skew) * S.n_ref) >> 32);
                const uint8_t* src = refs + S.ref_off + (size_t)idx * S.ref_len;
                const uint32_t n = S.len < S.ref_len ? S.len : S.ref_len;
                for (uint32_t p = 0; p < n; p++) dst[p] = src[p];
            } else if (S.kind == 'R' && use_mol) {
                uint64_t h = hash64(mol ^ 0x0123456789ABCDEFULL);
                for (uint32_t p = 0; p < S.len; p++) {
                    if ((p & 31) == 0 && p) h = hash64(h);
                    dst[p] = (uint8_t)((h >> (2 * (p & 31))) & 3);
                }
            } else if (S.pool) {
                const uint64_t id = ((uint64_t)skewed((uint32_t)rs, S.skew) * S.pool) >> 32;
                uint64_t h = hash64(id * 0x9E3779B97F4A7C15ULL + 77);
                for (uint32_t p = 0; p < S.len; p++) {
                    if ((p & 31) == 0 && p) h = hash64(h);
                    dst[p] = (uint8_t)((h >> (2 * (p & 31))) & 3);
                }
            }  // else: the random background stays
        }
    }
    // sequencing errors over the whole read: substitutions (always to another base) and N calls
    for (uint32_t p = 0; p < R; p += 4) {
        uint64_t r = splitmix(s);
        for (uint32_t k = 0; k < 4 && p + k < R; k++, r >>= 16) {
            const uint32_t v = (uint32_t)(r & 0xFFFF);
            if (v < c.p_n16) codes[p + k] = 4;
            else if (v < (uint32_t)c.p_n16 + c.p_sub16) codes[p + k] = (uint8_t)((codes[p + k] + 1 + (v % 3)) & 3);
        }
    }
    // qualities: per-read mean, per-base jitter
    if (qual) {
        const uint32_t nib = (uint32_t)((r1 >> 32) & 0xF) + (uint32_t)((r1 >> 36) & 0xF) + (uint32_t)((r1 >> 40) & 0xF) +
                             (uint32_t)((r1 >> 44) & 0xF);  // 0..60, bell shaped around 30
        int mean = lowq ? 8 + (int)((r1 >> 48) & 7) : (int)c.q_mean + ((int)nib - 30) * (int)c.q_spread / 15;
        if (mean < 2) mean = 2;
        if (mean > 40) mean = 40;
        for (uint32_t p = 0; p < R; p += 16) {
            uint64_t r = splitmix(s);
            for (uint32_t k = 0; k < 16 && p + k < R; k++, r >>= 4) {
                int q = mean + (int)((r & 0xF) % 13) - 6;
                if (q < 2) q = 2;
                if (q > 41) q = 41;
                qual[p + k] = (uint8_t)(33 + q);
            }
        }
    }
}

__global__ void k_generate(const bcs_config cfg, const uint8_t* __restrict__ refs, const uint64_t first, const uint64_t n,
                           const uint32_t W, const uint32_t plane_stride, const uint32_t qual_stride,
                           uint32_t* __restrict__ planes, uint16_t* __restrict__ read_len, uint8_t* __restrict__ qual) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint8_t codes[BCS_MAX_READ];
    uint8_t q[BCS_MAX_READ];
    gen_read(cfg, refs, first + t, codes, qual ? q : nullptr);
    uint32_t* dst = planes + t * plane_stride;
    const uint32_t R = cfg.read_len;
    for (uint32_t w = 0; w < W; w++) {
        uint32_t lo = 0, hi = 0, nm = 0;
        for (uint32_t b = 0; b < 32 && w * 32 + b < R; b++) {
            const uint8_t cde = codes[w * 32 + b];
            if (cde == 4) nm |= 1u << b;
            else {
                lo |= (uint32_t)(cde & 1) << b;
                hi |= (uint32_t)(cde >> 1) << b;
            }
        }
        dst[w] = lo;
        dst[W + w] = hi;
        dst[2 * W + w] = nm;
    }
    for (uint32_t w = 3 * W; w < plane_stride; w++) dst[w] = 0;
    read_len[t] = (uint16_t)R;
    if (qual) {
        uint8_t* qd = qual + t * qual_stride;
        for (uint32_t p = 0; p < qual_stride; p++) qd[p] = p < R ? q[p] : (uint8_t)'!';
    }
}

// ---- INT-pipe peak microbenchmarks (register only): the denominators of the INT roofline --------
template <int OP>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t* out, const uint32_t seed, const int iters) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u + 1u, a2 = a0 * 5u + 7u, a3 = a0 * 7u + 11u;
    uint32_t b0 = ~a0, b1 = ~a1, b2 = ~a2, b3 = ~a3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            // inline PTX so that exactly eight instructions of the kind under test are issued per step
            if (OP == 0) {
#define L3(d, x, y, z) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(x), "r"(y), "r"(z))
                L3(a0, a0, b0, a1); L3(a1, a1, b1, a2); L3(a2, a2, b2, a3); L3(a3, a3, b3, a0);
                L3(b0, b0, a1, a2); L3(b1, b1, a2, a3); L3(b2, b2, a3, a0); L3(b3, b3, a0, a1);
#undef L3
            } else if (OP == 1) {
#define PC(d, x) asm volatile("popc.b32 %0, %1;" : "=r"(d) : "r"(x))
                PC(a0, b0); PC(a1, b1); PC(a2, b2); PC(a3, b3);
                PC(b0, a1); PC(b1, a2); PC(b2, a3); PC(b3, a0);
#undef PC
            } else {
#define SF(d, x, y, n) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "r"(n))
                SF(a0, a0, b0, 3); SF(a1, a1, b1, 5); SF(a2, a2, b2, 7); SF(a3, a3, b3, 9);
                SF(b0, b0, a1, 11); SF(b1, b1, a2, 13); SF(b2, b2, a3, 15); SF(b3, b3, a0, 17);
#undef SF
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
}

template <int OP>
double int_peak_once(uint32_t* d_out, int grid, int iters, double ops_per_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_int_peak<OP><<<grid, 256>>>(d_out, 12345u, iters / 8);
    cudaEventRecord(e0);
    k_int_peak<OP><<<grid, 256>>>(d_out, 12345u, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return (double)grid * 256.0 * iters * ops_per_iter / (ms * 1e-3) / 1e12;
}

size_t name_bytes_upto(uint64_t n) {  // total decimal digits of 0..n-1
    size_t total = 0;
    uint64_t lo = 0, pow = 10;
    size_t d = 1;
    while (lo < n) {
        const uint64_t hi = pow < n ? pow : n;
        total += (hi - lo) * d;
        lo = hi;
        pow *= 10;
        d++;
    }
    return total;
}

}  // namespace

extern "C" {

// Measured INT throughput of this GPU in 10^12 lane-operations per second: LOP3 (3-input logic), POPC, SHF (funnel
// shift).  out[0..2].  Returns a cudaError_t as int.
int bcs_measure_int_peaks(double* out) {
    uint32_t* d_out = nullptr;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8;
    cudaError_t e = cudaMalloc(&d_out, (size_t)grid * 256 * sizeof(uint32_t));
    if (e != cudaSuccess) return (int)e;
    const int iters = 4096;
    out[0] = int_peak_once<0>(d_out, grid, iters, 16.0 * 8.0);        // 8 LOP3 per unrolled step (each line fuses to one)
    out[1] = int_peak_once<1>(d_out, grid, iters, 16.0 * 8.0);        // 8 POPC per unrolled step
    out[2] = int_peak_once<2>(d_out, grid, iters, 16.0 * 8.0);        // 8 SHF per unrolled step
    e = cudaDeviceSynchronize();
    cudaFree(d_out);
    return (int)e;
}

int bcs_generate_device(const bcs_config* cfg, const uint8_t* refs_dev, uint64_t first_read, uint64_t n_reads,
                        uint32_t max_read_len, uint32_t* planes, uint16_t* read_len, uint8_t* qual, void* cuda_stream) {
    if (!cfg || !planes || !read_len || cfg->read_len > BCS_MAX_READ || cfg->read_len > max_read_len) return (int)cudaErrorInvalidValue;
    if (n_reads == 0) return 0;
    const uint32_t W = (max_read_len + 31) / 32;
    const uint32_t ps = (3 * W) | 1u;
    const uint32_t qs = (((max_read_len + 3) / 4) | 1u) * 4;
    const unsigned block = 128;
    const uint64_t grid = (n_reads + block - 1) / block;
    if (grid > 0x7FFFFFFFull) return (int)cudaErrorInvalidValue;
    k_generate<<<(unsigned)grid, block, 0, static_cast<cudaStream_t>(cuda_stream)>>>(*cfg, refs_dev, first_read, n_reads, W, ps, qs,
                                                                                     planes, read_len, qual);
    return (int)cudaGetLastError();
}

size_t bcs_fastq_bytes(const bcs_config* cfg, uint64_t first_read, uint64_t n_reads) {
    // "@r" + digits + "\n" + R + "\n+\n" + R + "\n"
    const size_t fixed = 2 + 1 + cfg->read_len + 3 + cfg->read_len + 1;
    return n_reads * fixed + (name_bytes_upto(first_read + n_reads) - name_bytes_upto(first_read));
}

size_t bcs_generate_fastq(const bcs_config* cfg, const uint8_t* refs_host, uint64_t first_read, uint64_t n_reads, char* out,
                          size_t cap, unsigned threads) {
    const size_t need = bcs_fastq_bytes(cfg, first_read, n_reads);
    if (!out || cap < need || cfg->read_len > BCS_MAX_READ) return 0;
    if (threads == 0) threads = 1;
    const size_t fixed = 2 + 1 + cfg->read_len + 3 + cfg->read_len + 1;
    auto work = [&](uint64_t a, uint64_t b) {
        char* p = out + (a - first_read) * fixed + (name_bytes_upto(a) - name_bytes_upto(first_read));
        uint8_t codes[BCS_MAX_READ], q[BCS_MAX_READ];
        const uint32_t R = cfg->read_len;
        for (uint64_t i = a; i < b; i++) {
            gen_read(*cfg, refs_host, i, codes, q);
            *p++ = '@';
            *p++ = 'r';
            p += snprintf(p, 24, "%llu", (unsigned long long)i);
            *p++ = '\n';
            for (uint32_t k = 0; k < R; k++) *p++ = "ACGTN"[codes[k]];
            *p++ = '\n';
            *p++ = '+';
            *p++ = '\n';
            memcpy(p, q, R);
            p += R;
            *p++ = '\n';
        }
    };
    if (threads == 1 || n_reads < 4096) {
        work(first_read, first_read + n_reads);
    } else {
        std::vector<std::thread> pool;
        const uint64_t per = (n_reads + threads - 1) / threads;
        for (unsigned t = 0; t < threads; t++) {
            const uint64_t a = first_read + (uint64_t)t * per;
            const uint64_t b = a + per < first_read + n_reads ? a + per : first_read + n_reads;
            if (a < b) pool.emplace_back(work, a, b);
        }
        for (auto& th : pool) th.join();
    }
    return need;
}

}  // extern "C"
