/* include/bc_b200.h — C ABI of the B200-native decode-and-count path of NGS-Barcode-Count.
 *
 * The reference (Rust crate barcode-count 0.11.1) has no FFI: its hot path is
 * `SequenceParser::new(..).parse()` (src/parse.rs:28-76) feeding `Results::add_count`
 * (src/info.rs:735-808) and the six `SequenceErrors` counters (src/info.rs:16-139), with
 * `ResultsEnrichment` (src/info.rs:840-904) computed at output time.  This header is the boundary a
 * host in any language (the reference's Rust via `extern "C"`, this repo's C++ CLI, Python/ctypes)
 * binds to replace exactly that path.  Plain pointers and sizes only; one bc_ctx per GPU; a ctx is
 * driven from one host thread at a time.  Every entry point returns BC_OK (0) or a negative BC_E*
 * code; the message is available from bc_last_error().  There is no CPU fallback: if no CUDA device
 * is usable bc_create() fails with BC_ECUDA.
 *
 * Citations are file:line into the reference repository.
 */
#ifndef BC_B200_H
#define BC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BC_ABI_VERSION 1
#define BC_MAX_SLOTS 16        /* sample + counted + random barcodes in one scheme */
#define BC_MAX_TEMPLATE 256    /* template (format_string) length limit, bases */
#define BC_MAX_READ_LEN 1024   /* read length limit, bases */
#define BC_MAX_REF_LEN 32      /* reference barcode length limit (bases compared per candidate) */
#define BC_MAX_KEY_BITS 126    /* packed (sample, counted.., UMI) key must fit 126 bits */

enum {
    BC_OK = 0,
    BC_EINVAL = -1,       /* bad argument / malformed config */
    BC_ECUDA = -2,        /* CUDA runtime error (message has the CUDA error string) */
    BC_ENOMEM = -3,
    BC_EUNSUPPORTED = -4, /* valid for the reference but outside this build's limits */
    BC_ESTATE = -5        /* call sequence error */
};

/* Per-read outcome.  Exactly one per read, as the reference bumps exactly one SequenceErrors counter per
 * read (info.rs:60-127).  BC_ST_MATCHED/BC_ST_DUPLICATE: parse.rs:65-69; BC_ST_CONSTANT: parse.rs:144-147;
 * BC_ST_LOW_QUALITY: parse.rs:105-113; BC_ST_SAMPLE: parse.rs:132-135; BC_ST_COUNTED: parse.rs:137-140.
 * BC_ST_UNSUPPORTED has no reference counterpart: the read holds a character outside {A,C,G,T,N} (the packer
 * flags it); such reads are counted separately and never silently mis-decoded. */
enum {
    BC_ST_MATCHED = 0,
    BC_ST_DUPLICATE = 1,
    BC_ST_CONSTANT = 2,
    BC_ST_LOW_QUALITY = 3,
    BC_ST_SAMPLE = 4,
    BC_ST_COUNTED = 5,
    BC_ST_UNSUPPORTED = 6
};

/* Order of the counters everywhere in this ABI = display order of SequenceErrors (info.rs:146-151),
 * plus the unsupported-read counter. */
enum {
    BC_CNT_MATCHED = 0,
    BC_CNT_CONSTANT = 1,
    BC_CNT_SAMPLE = 2,
    BC_CNT_COUNTED = 3,
    BC_CNT_DUPLICATES = 4,
    BC_CNT_LOW_QUALITY = 5,
    BC_CNT_UNSUPPORTED = 6,
    BC_N_COUNTERS = 7
};

/* One barcode of the scheme.  kind: 'S' sample `[n]`, 'B' counted `{n}`, 'R' random `(n)` (info.rs:240-249).
 * n_ref == 0 means raw-key mode: the captured bases (N included) are the key (parse.rs:453-454, 484-505,
 * 512-513).  Reference barcodes may be longer or shorter than `len`: comparisons run over the shorter of
 * the two, exact membership needs equal length (parse.rs:568, Q10 in SURVEY.md). */
typedef struct {
    uint8_t kind;
    uint16_t offset;  /* first template position of the barcode */
    uint16_t len;
    uint16_t max_err; /* MaxSeqErrors (info.rs:499-523) */
    uint32_t n_ref;
    const char *const *ref_seqs; /* n_ref NUL-terminated strings over {A,C,G,T,N} */
} bc_slot;

/* Run configuration: what SequenceFormat / BarcodeConversions / MaxSeqErrors hand to SequenceParser::new
 * (parse.rs:28-52, main.rs:93-113). */
typedef struct {
    uint32_t abi_version;        /* BC_ABI_VERSION */
    const char *template_chars;  /* SequenceFormat::format_string (info.rs:177): 'N' at barcode and format-N positions */
    uint32_t template_len;
    const char *region_codes;    /* SequenceFormat::regions_string (info.rs:178): S/B/R/C, shorter than the template when
                                    the format holds N runs (info.rs:287-295) */
    uint32_t region_len;
    uint32_t n_slots;
    bc_slot slots[BC_MAX_SLOTS]; /* in template order */
    uint16_t max_const_err;      /* MaxSeqErrors::max_constant_errors (info.rs:527-531) */
    float min_quality;           /* --min-quality; 0 disables the filter (parse.rs:98) */
    uint32_t max_read_len;       /* longest read any batch will carry */
} bc_config;

enum { BC_LOC_HOST = 0, BC_LOC_DEVICE = 1 };

/* Fixed-stride packed batch of reads.
 *   planes  : n_reads records of `plane_stride` u32 words.  A record holds three bit planes of W =
 *             bc_plane_words(max_read_len) words each — lo, hi, nmask — base i at bit (i & 31) of word (i >> 5):
 *             A=(0,0) C=(1,0) G=(0,1) T=(1,1) as (lo,hi); nmask=1 where the read has 'N' (then lo=hi=0).
 *             Bits at and beyond the read length are 0.  plane_stride = bc_plane_stride(max_read_len) (3W, made odd).
 *   read_len: bases per read; bit 15 (BC_READ_UNSUPPORTED) set when the read held a character outside ACGTN or a
 *             quality character below '!' (the reference's `q - 33` underflows there, parse.rs:326).
 *   qual    : n_reads records of `qual_stride` bytes of raw FASTQ quality characters (Phred+33), or NULL when
 *             min_quality == 0.  qual_stride = bc_qual_stride(max_read_len).
 * `location` says whether the three pointers are host (pinned or pageable) or device memory. */
#define BC_READ_UNSUPPORTED 0x8000u
typedef struct {
    uint32_t n_reads;
    uint32_t plane_stride;
    uint32_t qual_stride;
    int32_t location;
    const uint32_t *planes;
    const uint16_t *read_len;
    const uint8_t *qual;
} bc_batch;

uint32_t bc_plane_words(uint32_t max_read_len);
uint32_t bc_plane_stride(uint32_t max_read_len);
uint32_t bc_qual_stride(uint32_t max_read_len);

typedef struct bc_ctx bc_ctx;

/* Replaces the construction of the worker pool (main.rs:93-113, parse.rs:28-52) and of Results/SequenceErrors
 * (info.rs:678-732, 40-49).  expected_reads sizes the device tables (they grow on demand; 0 = small default). */
int bc_create(const bc_config *cfg, int device, uint64_t expected_reads, bc_ctx **out);
void bc_destroy(bc_ctx *ctx);
/* Message of the last failure on `ctx`; ctx == NULL gives the last bc_create failure of this thread. */
const char *bc_last_error(const bc_ctx *ctx);

/* Run all work on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream). */
int bc_set_stream(bc_ctx *ctx, void *cuda_stream);

/* Replaces one pass of SequenceParser::parse (parse.rs:53-76) over a batch: locate (parse.rs:89-96, 151-163,
 * 287-313), quality filter (parse.rs:98-119, 331-375), barcode correction (parse.rs:439-524, 553-593), count
 * (info.rs:735-808) and the outcome counters.  Asynchronous; host batches are copied through internal pinned
 * staging.  Batch memory may be rewritten after bc_wait_copies (host batches) / bc_sync (device batches). */
int bc_submit(bc_ctx *ctx, const bc_batch *batch);
int bc_sync(bc_ctx *ctx);
/* Blocks until the host->device copies of every submitted host batch are done (their memory may then be
 * rewritten) without waiting for the kernels. */
int bc_wait_copies(bc_ctx *ctx);

/* SequenceErrors (info.rs:16-23) so far, BC_CNT_* order.  Synchronises. */
int bc_get_counters(bc_ctx *ctx, uint64_t out[BC_N_COUNTERS]);

/* Test hooks (no counting): per-read result of the locate step / of the whole decode.  Output arrays are host
 * memory with n_reads (× n_slots) elements; any may be NULL. */
typedef struct {
    uint8_t *status;     /* BC_ST_MATCHED when located, else BC_ST_CONSTANT / BC_ST_UNSUPPORTED */
    int16_t *offset;     /* scheme start in the read, -1 when not located */
    uint8_t *repaired;   /* 1 when located by the constant-region repair (parse.rs:287-313) */
} bc_locate_out;
typedef struct {
    uint8_t *status;     /* BC_ST_* before de-duplication (never BC_ST_DUPLICATE) */
    int16_t *offset;
    uint8_t *repaired;
    int32_t *slot_index; /* [n_reads][n_slots] index into ref_seqs, -1 for raw slots or when not reached */
    uint64_t *key_lo;    /* packed key of matched reads (decode with bc_key_decode) */
    uint64_t *key_hi;
} bc_decode_out;
int bc_locate_only(bc_ctx *ctx, const bc_batch *batch, bc_locate_out *out);
int bc_decode_only(bc_ctx *ctx, const bc_batch *batch, bc_decode_out *out);

/* A table of (key, count) rows in host memory, owned by the library until bc_table_free.  `mask` (enrichment
 * tables only) has bit k set when counted barcode k is part of the row.  key_hi is NULL when every key fits 64
 * bits.  The rows of bc_finish live in pinned memory owned by the ctx (flags has BC_TABLE_BORROWED): they stay
 * valid until the next bc_finish / bc_destroy on that ctx, and bc_table_free only clears the struct. */
#define BC_TABLE_BORROWED 1u
typedef struct {
    uint64_t n_rows;
    uint64_t *key_lo;
    uint64_t *key_hi;
    uint64_t *count;
    uint32_t *mask;
    uint32_t flags;
} bc_table;
void bc_table_free(bc_table *t);

/* Replaces reading Results at output time (output.rs:226-272): one row per (sample, counted barcodes) key with
 * its count — the number of reads, or of distinct random barcodes when the scheme has one (info.rs:780-791,
 * output.rs:265-270).  Keys exclude the random barcode.  Synchronises; may be called repeatedly. */
int bc_finish(bc_ctx *ctx, bc_table *rows);

/* Replaces ResultsEnrichment::add_single / add_double (info.rs:840-904) over the final table: marginal sums
 * over one and over two counted barcodes, per sample.  `doubles` may be NULL. */
int bc_enrich(bc_ctx *ctx, bc_table *singles, bc_table *doubles);

/* Decode a key: per slot (scheme order) either the reference index or, for raw slots, the DNA string.
 * with_umi = 0 for table rows (bc_finish / bc_enrich: the random barcode is not part of the key), 1 for the
 * per-read keys of bc_decode_only (random barcode included).  mask = 0 means every counted barcode is present;
 * for enrichment rows pass the row's mask.  idx_out has n_slots entries (-1 for raw / absent slots); str_out is
 * n_slots strings of `str_stride` bytes (empty for indexed / absent slots).  Host only, no GPU work. */
int bc_key_decode(const bc_ctx *ctx, uint64_t key_lo, uint64_t key_hi, uint32_t mask, int with_umi, int32_t *idx_out,
                  char *str_out, uint32_t str_stride);

/* ---- multi-GPU (one process per GPU; the caller owns the communicator, e.g. torch.distributed/NCCL) ----------
 * Without a random barcode each rank counts its own reads and the tables are merged at the end:
 * bc_export_rows gives device-resident (key_lo,key_hi,count) arrays to all-gather, bc_import_rows adds them.
 * With a random barcode, de-duplication must be global: bc_decode_route decodes a batch and writes the
 * (key,UMI) records of matched reads into n_ranks device buckets by owner = hash(key) % n_ranks instead of
 * inserting them; after the all-to-all the owner calls bc_insert_records. */
typedef struct {
    uint64_t lo, hi;
} bc_record;
int bc_decode_route(bc_ctx *ctx, const bc_batch *batch, uint32_t n_ranks, bc_record *dev_buckets,
                    uint64_t bucket_capacity, uint32_t *dev_bucket_counts);
int bc_insert_records(bc_ctx *ctx, const bc_record *dev_records, uint64_t n);

/* Fused routing (one process per GPU on one NVLink/NVSwitch box, at most 8 ranks): the decode kernel stores every
 * matched (key, UMI) record straight into its owner's receive region over NVLink peer memory — no bucket + copy
 * step.  Each rank opens its receive buffer (2 parities x n_ranks sources x capacity records) and gets a
 * BC_IPC_HANDLE_BYTES handle; the caller exchanges the handles (any transport) and passes all of them, rank order,
 * to bc_route_connect.  Per batch: bc_route_submit(parity = batch index & 1) decodes and routes, leaving in
 * dev_counts[r] the number of records sent to rank r; the caller all-gathers the counts (that collective is also the
 * barrier that makes the peer stores visible) and hands the owner the counts of what it received:
 * dev_counts_from[s * count_stride] records from source rank s.  expected_records bounds the table growth check. */
#define BC_IPC_HANDLE_BYTES 64
int bc_route_open(bc_ctx *ctx, uint32_t n_ranks, uint32_t rank, uint64_t capacity, void *ipc_handle_out);
int bc_route_connect(bc_ctx *ctx, const void *ipc_handles);
int bc_route_submit(bc_ctx *ctx, const bc_batch *batch, uint32_t parity, uint32_t *dev_counts);
int bc_route_insert(bc_ctx *ctx, uint32_t parity, const uint32_t *dev_counts_from, uint32_t count_stride,
                    uint64_t expected_records);
int bc_export_rows(bc_ctx *ctx, uint64_t **dev_key_lo, uint64_t **dev_key_hi, uint64_t **dev_count, uint64_t *n_rows);
int bc_import_rows(bc_ctx *ctx, const uint64_t *dev_key_lo, const uint64_t *dev_key_hi, const uint64_t *dev_count,
                   uint64_t n_rows);
/* When the count table is a dense array (small key space, no random barcode) ranks can merge with ONE in-place
 * all-reduce (sum) over it instead of exchanging rows: *dev_counts / *n give the device array (u64 per key), or
 * NULL / 0 when the table is not dense.  Work queued on the ctx stream so far is ordered before the caller's use. */
int bc_dense_counts(bc_ctx *ctx, uint64_t **dev_counts, uint64_t *n);
int bc_add_counters(bc_ctx *ctx, const uint64_t add[BC_N_COUNTERS]);
int bc_reset(bc_ctx *ctx); /* clear tables and counters, keep configuration */

/* ---- measurement ------------------------------------------------------------------------------------------ */
enum { BC_K_DECODE = 0, BC_K_SCAN = 1, BC_K_INSERT = 2, BC_K_FINISH = 3, BC_K_OTHER = 4, BC_N_KERNELS = 5 };
typedef struct {
    uint64_t launches[BC_N_KERNELS];
    double ms[BC_N_KERNELS]; /* CUDA-event time on the ctx stream, only while profiling is on */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t table_capacity, table_entries;
    uint32_t key_bits, wide_keys, dense_table;
    uint32_t deferred_count; /* 1: matched reads are appended to a record buffer and counted at bc_finish /
                                bc_get_counters (partitioned, in shared memory); 0: tables updated read by read */
    uint32_t flushed_global; /* 1: the last flush fell back to the global-memory tables */
    uint32_t flush_stages;   /* the last shared-memory flush: 1 = no hot key, partitioned by key and counted in one pass;
                                2 = partitioned by (key, random barcode), then by key; 3 = one pass, with the few hot
                                keys' partitions set aside and sent through the two stages */
} bc_profile;
int bc_set_profiling(bc_ctx *ctx, int on);
int bc_get_profile(bc_ctx *ctx, bc_profile *out); /* synchronises */
int bc_reset_profile(bc_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* BC_B200_H */
