/* include/bc_host.h — host side of the drop-in: the reference's run set-up (format file, conversion CSVs,
 * error caps), FASTQ ingest + packing into bc_batch, and the count/merged/enrichment CSV writers, as a C API
 * over the C++ implementation in ngs-barcode-count_b200/csrc/host/.  It replaces, for this path only:
 *   SequenceFormat::parse_format_file      info.rs:215-310
 *   BarcodeConversions::*                  info.rs:364-456
 *   MaxSeqErrors::new                      info.rs:490-543
 *   input::read_fastq                      input.rs:24-148   (records -> packed pinned batches instead of a deque)
 *   WriteFiles::write_counts_files         output.rs:74-485
 * The compute itself goes through include/bc_b200.h; nothing here decodes a read on the CPU.
 */
#ifndef BC_HOST_H
#define BC_HOST_H

#include "bc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bch_run bch_run;

/* arguments.rs:6-20, the part that shapes the path.  Negative max_err_* = flag not given (20 % default). */
typedef struct {
    const char *format_path;            /* --sequence-format (required) */
    const char *sample_barcodes_path;   /* --sample-barcodes or NULL */
    const char *counted_barcodes_path;  /* --counted-barcodes or NULL */
    int max_errors_counted_barcode;
    int max_errors_sample;
    int max_errors_constant;
    float min_quality;                  /* --min-quality */
    uint32_t max_read_len;              /* longest read to expect (0: 2 x template length, at least 160) */
} bch_args;

/* Parses the three input files and derives the caps; on failure returns NULL and writes the message. */
bch_run *bch_open(const bch_args *args, char *err, int errlen);
void bch_close(bch_run *run);
/* Tuning / test switches of the host side (results never change).  "lean_writer_min_rows": count tables of at least this
 * many rows are written by the streaming CSV writer (text straight from the packed keys on all host threads; default 4 M);
 * "wire_batches": 0 makes bch_count_fastq hand the GPU plain bc_batch arrays instead of the transfer form (default 1);
 * "fused_ingest": 0 makes it frame a whole block of a plain file before packing it instead of doing both in one pass (default 1). */
int bch_set_option(bch_run *run, const char *name, long long value);
/* The reference rewrites "Total sequences: N" in place while it reads (input.rs:54-57, 151-158): fn is called with the
 * number of records handed to the GPU so far, after every batch of bch_count_fastq[_multi]. */
typedef void (*bch_progress_fn)(uint64_t reads_so_far, void *user);
void bch_set_progress(bch_run *run, bch_progress_fn fn, void *user);
/* The configuration to hand to bc_create (owned by `run`). */
const bc_config *bch_config(const bch_run *run);
/* "-FORMAT-" and "-BARCODE INFO-" blocks as the reference prints them (info.rs:313-335, 618-659). */
const char *bch_describe(const bch_run *run);
uint32_t bch_barcode_num(const bch_run *run);
/* reference DNA / ID of a slot's i-th barcode (NULL when out of range) */
const char *bch_ref_dna(const bch_run *run, uint32_t slot, uint32_t i);
const char *bch_ref_name(const bch_run *run, uint32_t slot, uint32_t i);

/* Packs n reads (text) into caller-provided batch arrays sized with bc_plane_stride / bc_qual_stride.
 * quals may be NULL (then qual_out is not touched).  Reads longer than max_read_len give BC_EINVAL (the caller chose
 * the geometry).  A quality string shorter than its sequence is packed so that the filter behaves as the reference's
 * zip of scores and region codes does (parse.rs:338-343: a barcode run the string does not outlast is never tested);
 * a longer one is cut at the sequence length.  threads <= 1 packs on the calling thread. */
int bch_pack(uint32_t max_read_len, uint32_t n, const char *const *seqs, const char *const *quals, uint32_t *planes_out,
             uint16_t *read_len_out, uint8_t *qual_out, unsigned threads);
/* Same, for reads given as one '\n'-separated text block each (convenient from ctypes). */
int bch_pack_lines(uint32_t max_read_len, uint32_t n, const char *seq_lines, const char *qual_lines, uint32_t *planes_out,
                   uint16_t *read_len_out, uint8_t *qual_out, unsigned threads);

/* The packers and the line-end scanner exist in scalar, AVX2 and AVX-512 (BW + VBMI) form and use the best the CPU has.
 * Test hook: cap them at level 0 (scalar), 1 (AVX2) or 2 (AVX-512); returns the level now in use.  Process-wide. */
int bch_set_simd_level(int level);

/* The transfer form of a host batch (bc_wire_batch, bc_submit_wire): bch_wire_from_batch converts a host bc_batch of
 * geometry max_read_len into arrays laid out in `buf` (at least bch_wire_bound bytes; pinned memory for full-rate copies)
 * and fills *out with pointers into it.  qual_bits = 0 picks the narrowest quality form that holds every character of the
 * batch (2 or 4 bits with a dictionary, 6 bits, or plain bytes); 2 / 4 / 6 / 8 ask for that form (BC_EINVAL when the
 * batch does not fit it).  bch_count_fastq packs reads straight into this form (6-bit quality, N calls as a list). */
size_t bch_wire_bound(uint32_t n_reads, uint32_t max_read_len, int with_qual);
int bch_wire_from_batch(const bc_batch *batch, uint32_t max_read_len, uint32_t qual_bits, void *buf, size_t buf_bytes,
                        bc_wire_batch *out);
/* Wall time of the phases of the last bch_count_fastq[_multi] on the ingest thread: seconds[6] = split (record framing),
 * pack, submit calls, waiting for copies / the GPU, total, mapping the file; counts[3] = batches, batches whose quality went as plain bytes,
 * batches whose N calls went as a dense plane. */
int bch_ingest_stats(const bch_run *run, double *seconds, uint64_t *counts);

/* Host-only test hook of the record framing (input.rs:115-148): streams a .fastq / .fastq.gz file (plain, gzip with any
 * number of members, or bgzip — whose members are inflated on `threads` host threads) through the same block reader
 * bch_count_fastq uses and reports the number of records, of bases, and the CRC-32 of every record's sequence and
 * quality bytes in file order.  No GPU work. */
int bch_scan_fastq(const char *fastq_path, unsigned threads, uint64_t *n_records, uint64_t *n_bases, uint32_t *crc, char *err,
                   int errlen);

/* Host-only test hook of the plain-file path of bch_count_fastq: the file is mapped and cut into blocks of block_bytes
 * (0: 64 MB) that `threads` host threads split into records slice by slice (a slice is at least min_slice_bytes, 0:
 * 64 KB as in bch_count_fastq); reports like bch_scan_fastq.  No GPU work. */
int bch_split_fastq(const char *fastq_path, unsigned threads, size_t block_bytes, size_t min_slice_bytes, uint64_t *n_records,
                    uint64_t *n_bases, uint32_t *crc, char *err, int errlen);

/* Host-only test hook of the one-pass walker behind bch_count_fastq on plain files (a host thread frames a cache-sized chunk
 * of the mapping and packs its records at once into rows it reserves in the batch): batches of batch_rows rows, chunks of
 * chunk_bytes (0: 256 KB).  Checks that every row of every batch is filled exactly once; digest = sum over the records of
 * the CRC-32 of their sequence and quality bytes (rows of a batch are in no particular order).  No GPU work. */
int bch_walk_fastq(const char *fastq_path, unsigned threads, size_t chunk_bytes, uint32_t batch_rows, uint64_t *n_records,
                   uint64_t *n_bases, uint64_t *digest, uint64_t *n_batches, char *err, int errlen);

/* input::read_fastq replacement: streams a .fastq / .fastq.gz file through pinned double-buffered batches of
 * `batch_reads` reads into ctx (bc_submit).  threads = host threads (a pool kept by the run) that split and pack.
 * Reads of any length up to BC_MAX_READ_LEN are decoded: a batch that holds a read longer than the run's
 * max_read_len is packed with a wider stride (the reference has no length limit, input.rs:115-148); longer reads
 * still are counted, as unsupported.  Returns BC_OK and the number of records. */
int bch_count_fastq(bch_run *run, bc_ctx *ctx, const char *fastq_path, unsigned threads, uint32_t batch_reads,
                    uint64_t *total_reads, char *err, int errlen);

/* The same over several GPUs in one process — one context per GPU, all created from bch_config(run): batches go to
 * the contexts in turn (pinned staging buffers on each GPU's own NUMA node); after the last batch the contexts merge as
 * SURVEY.md section 8(e) prescribes: hashed keys through one exchange of the records over NVLink peer memory
 * (bc_exchange_*: every context then owns a disjoint set of keys), a dense count table by adding the tables into
 * ctxs[0].  bch_counters_multi / bch_write_counts_multi then read the merged result. */
int bch_count_fastq_multi(bch_run *run, bc_ctx *const *ctxs, int n_ctx, const char *fastq_path, unsigned threads,
                          uint32_t batch_reads, uint64_t *total_reads, char *err, int errlen);
int bch_counters_multi(bc_ctx *const *ctxs, int n_ctx, uint64_t out[BC_N_COUNTERS]);

/* WriteFiles::write_counts_files: per-sample count CSVs, merged CSV, Single/Double enrichment CSVs, same names
 * and layout as the reference; rows are written sorted.  names_out receives one "file name<TAB>barcode rows" line
 * per file written (what the reference keeps for its stats file, output.rs:143-165).  Returns the number of files. */
int bch_write_counts(bch_run *run, bc_ctx *ctx, const char *output_dir, const char *prefix, int merge_output, int enrich,
                     char *names_out, int names_len, char *err, int errlen);
/* After bch_count_fastq_multi: the rows of every owner form one table (output.rs:74-181); enrichment marginals of the
 * owners are summed (on the device when they are dense counter arrays). */
int bch_write_counts_multi(bch_run *run, bc_ctx *const *ctxs, int n_ctx, const char *output_dir, const char *prefix,
                           int merge_output, int enrich, char *names_out, int names_len, char *err, int errlen);

#ifdef __cplusplus
}
#endif
#endif /* BC_HOST_H */
