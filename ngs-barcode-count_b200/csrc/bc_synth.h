/* bc_synth.h — deterministic synthetic-read generator for bench.py and the tests (NOT part of the drop-in boundary).
 *
 * Read i of a workload is a pure function of (config, i): a counter-based splitmix64 stream and integer-only
 * arithmetic, so the same read comes out of the CUDA kernel (packed bc_batch layout, straight into HBM) and of the
 * host code (FASTQ text for the CPU oracle / the FASTQ ingest path) without ever storing the text of 10^8..10^9
 * reads.  SURVEY.md §8(d) lists the workloads; ngs-barcode-count_b200/synth.py builds their configurations.
 */
#ifndef BC_SYNTH_H
#define BC_SYNTH_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCS_MAX_SLOTS 16
#define BCS_MAX_READ 256

typedef struct {
    uint8_t kind;       /* 'S' 'B' 'R' */
    uint8_t skew;       /* abundance: 0 uniform, 1 ~u^2, 2 ~u^3 over the reference index / pool id */
    uint16_t offset;    /* position in the template */
    uint16_t len;       /* bases in the template */
    uint16_t ref_len;   /* bases per reference barcode (min(len, ref_len) are copied, the rest stays random) */
    uint32_t n_ref;     /* reference barcodes to draw from; 0 = none */
    uint32_t ref_off;   /* byte offset of the first one in the `refs` blob (codes 0..3, ref_len bytes each) */
    uint64_t pool;      /* n_ref == 0 only: > 0 draws an id in [0, pool) and derives the bases from it (lineage
                           barcodes); 0 leaves the bases random */
} bcs_slot;

typedef struct {
    uint64_t seed;
    uint32_t read_len, template_len, n_slots, max_start;
    uint8_t template_codes[BCS_MAX_READ]; /* 0..3 at constant positions; ignored inside slots */
    bcs_slot slots[BCS_MAX_SLOTS];
    uint32_t p_junk;      /* P(read carries no template) as a u32 threshold (p * 2^32) */
    uint32_t p_lowq;      /* P(read is a low-quality read: mean Phred 8..15) */
    uint32_t p_enriched;  /* molecule mode: P(read comes from one of n_enriched compounds) */
    uint16_t p_sub16;     /* per-base substitution probability * 65536 */
    uint16_t p_n16;       /* per-base N probability * 65536 */
    uint32_t n_enriched;
    uint64_t molecule_pool; /* > 0: (all S/B slots, UMI) are a function of one molecule id drawn in [0, pool):
                               PCR duplicates; 0: every slot is drawn independently */
    uint8_t q_mean, q_spread; /* good reads: mean Phred q_mean +- q_spread, per-base jitter +-6, clipped to 2..41 */
} bcs_config;

/* Packed batch straight into device memory (bc_batch layout of include/bc_b200.h); qual may be NULL.
 * `refs_dev` is the reference blob in device memory.  Runs on `cuda_stream`; returns a cudaError_t as int. */
int bcs_generate_device(const bcs_config *cfg, const uint8_t *refs_dev, uint64_t first_read, uint64_t n_reads,
                        uint32_t max_read_len, uint32_t *planes, uint16_t *read_len, uint8_t *qual, void *cuda_stream);

/* Register-only microbenchmarks of the INT pipes on the current device, in 10^12 lane-operations per second:
 * out[0] LOP3 (3-input logic), out[1] POPC, out[2] SHF (funnel shift). */
int bcs_measure_int_peaks(double *out);

/* The same reads as FASTQ text ("@r<i>\nSEQ\n+\nQUAL\n") into `out`; returns bytes written, 0 when `cap` is too
 * small.  bcs_fastq_bytes gives the exact size. */
size_t bcs_fastq_bytes(const bcs_config *cfg, uint64_t first_read, uint64_t n_reads);
size_t bcs_generate_fastq(const bcs_config *cfg, const uint8_t *refs_host, uint64_t first_read, uint64_t n_reads, char *out,
                          size_t cap, unsigned threads);

#ifdef __cplusplus
}
#endif
#endif
