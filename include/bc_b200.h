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

#define BC_ABI_VERSION 2
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
    uint32_t flags;              /* BC_CFG_* (0 for production use) */
} bc_config;

/* bc_config.flags.  BC_CFG_INLINE_COUNT: update global hash tables read by read inside the decode kernel instead of
 * appending records and counting them at the flush (measurement aid: the design the deferred flush replaced). */
#define BC_CFG_INLINE_COUNT 1u
/* The decode kernel exists twice: generic (any scheme, run constants read from a kernel parameter) and specialised for
 * the run by NVRTC (same device code, constants folded in; about two seconds of compilation, cached per process and
 * configuration).  bc_create specialises when expected_reads >= 2^22; these flags force it (bc_create fails when the
 * specialisation cannot be built) or forbid it.  Results are identical either way. */
#define BC_CFG_SPECIALIZE 2u
#define BC_CFG_NO_SPECIALIZE 4u

enum { BC_LOC_HOST = 0, BC_LOC_DEVICE = 1 };

/* Fixed-stride packed batch of reads.
 *   planes  : n_reads records of `plane_stride` u32 words.  A record holds three bit planes of W words each — lo, hi,
 *             nmask — base i at bit (i & 31) of word (i >> 5); W = plane_stride / 3 is the batch's own geometry
 *             (reads of up to 32 W bases): normally bc_plane_words(cfg.max_read_len), wider for a batch that holds a
 *             longer read (up to BC_MAX_READ_LEN) —
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

/* Transfer form of a HOST batch ("wire batch"): the same reads as a bc_batch in fewer bytes, for the PCIe crossing that
 * bounds the end-to-end rate (DEL, 150-nt reads with quality: 218 -> 159 bytes per read).  bc_submit_wire copies the arrays
 * as they are and expands them on the device into the bc_batch layout before the decode kernel runs; results are those
 * of bc_submit on the equivalent bc_batch.  Geometry: W = bc_plane_words(max_read_len).
 *   lohi    : n_reads records of 2W words — the lo plane, then the hi plane (both 0 where the read has 'N').
 *   read_len: as in bc_batch (bit 15 = BC_READ_UNSUPPORTED).
 *   N calls : either `nmask`, n_reads records of W words (the N plane), or — nmask == NULL — a list of n_calls
 *             (n_read[i] = read index in the batch, n_pos[i] = base position) pairs: 6 bytes per call against 4W bytes
 *             per read, so the list is the smaller form below about 2W/3 N calls per read — every real run.
 *   qual    : the first bc_wire_qual_codes(max_read_len) quality characters of every read as `qual_bits`-bit codes,
 *             code i at bits [i * qual_bits, (i + 1) * qual_bits) of the record's little-endian bit stream, records of
 *             qual_stride = bc_wire_qual_stride(max_read_len, qual_bits) bytes.  qual_bits 8: the characters themselves;
 *             6: character - 33 (characters '!' .. '_'; code 63 stands for the byte 0xFF that the packer writes where a
 *             quality line ends before its sequence, see bc_batch / parse.rs:338-343); 4 and 2: index into qual_dict
 *             (instruments that bin their qualities emit 4 to 8 distinct characters).  Ignored when min_quality == 0. */
typedef struct {
    uint32_t n_reads;
    uint32_t max_read_len;
    uint32_t qual_bits;
    uint32_t qual_stride;
    uint32_t n_calls;
    uint8_t qual_dict[16];
    const uint32_t *lohi;
    const uint16_t *read_len;
    const uint32_t *nmask;
    const uint32_t *n_read;
    const uint16_t *n_pos;
    const uint8_t *qual;
} bc_wire_batch;
uint32_t bc_wire_qual_codes(uint32_t max_read_len);                 /* max_read_len rounded up to a multiple of 4 */
uint32_t bc_wire_qual_stride(uint32_t max_read_len, uint32_t bits); /* bytes per read, a multiple of 4 */

typedef struct bc_ctx bc_ctx;

/* Replaces the construction of the worker pool (main.rs:93-113, parse.rs:28-52) and of Results/SequenceErrors
 * (info.rs:678-732, 40-49).  expected_reads sizes the device tables (they grow on demand; 0 = small default). */
int bc_create(const bc_config *cfg, int device, uint64_t expected_reads, bc_ctx **out);
void bc_destroy(bc_ctx *ctx);
int bc_device_of(const bc_ctx *ctx); /* the CUDA device the context was created on */
/* "specialised", or why the context runs the generic decode kernel (small job, no libnvrtc, compile log). */
const char *bc_specialization_note(const bc_ctx *ctx);
/* Host-only check that the specialised decode kernel of `cfg` compiles for sm_100a (NVRTC, no GPU needed): BC_OK and
 * "cubin bytes: N" in log, or an error code and the reason. */
int bc_jit_check(const bc_config *cfg, char *log, int loglen);
/* Message of the last failure on `ctx`; ctx == NULL gives the last bc_create failure of this thread. */
const char *bc_last_error(const bc_ctx *ctx);

/* Run all work on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream). */
int bc_set_stream(bc_ctx *ctx, void *cuda_stream);

/* Replaces one pass of SequenceParser::parse (parse.rs:53-76) over a batch: locate (parse.rs:89-96, 151-163,
 * 287-313), quality filter (parse.rs:98-119, 331-375), barcode correction (parse.rs:439-524, 553-593), count
 * (info.rs:735-808) and the outcome counters.  Asynchronous; host batches are copied through internal pinned
 * staging.  Batch memory may be rewritten after bc_wait_copies (host batches) / bc_sync (device batches). */
int bc_submit(bc_ctx *ctx, const bc_batch *batch);
/* bc_submit for a host batch in its transfer form (see bc_wire_batch).  Asynchronous; the arrays may be rewritten after
 * bc_wait_copies. */
int bc_submit_wire(bc_ctx *ctx, const bc_wire_batch *batch);
int bc_sync(bc_ctx *ctx);
/* Blocks until the host->device copies of every submitted host batch are done (their memory may then be
 * rewritten) without waiting for the kernels. */
int bc_wait_copies(bc_ctx *ctx);
/* For a host that alternates between two batch buffers: blocks until the copies of every submitted host batch EXCEPT the
 * most recent one are done, i.e. until the buffer handed over two submits ago may be rewritten, while the copy just queued
 * keeps running.  (Batches of one form only — all bc_submit or all bc_submit_wire — between two calls.) */
int bc_wait_older_copies(bc_ctx *ctx);

/* Test / measurement switches of the counting step (results never change): "flush_global" = 1 sends the flush through
 * the global-memory hash tables (the fallback of oversized partitions), "flush_two_stage" = 1 always partitions by
 * (key, random barcode) first (the path of hot keys). */
int bc_set_option(bc_ctx *ctx, const char *name, int value);

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
 * over one and over two counted barcodes, per sample.  `doubles` may be NULL.  With index-coded barcodes all
 * marginals are dense counter arrays filled in one pass over the rows. */
int bc_enrich(bc_ctx *ctx, bc_table *singles, bc_table *doubles);
/* The dense marginal counters of this context's rows (computed on first use): *dev_counters / *n give the device
 * array (u64 each), or NULL / 0 when the scheme has raw barcodes (no dense index space).  Ranks of a multi-GPU job
 * whose rows are partitioned by key sum these arrays (all-reduce, or bc_peer_add) before one of them calls bc_enrich. */
int bc_marginals(bc_ctx *ctx, uint64_t **dev_counters, uint64_t *n);

/* Decode a key: per slot (scheme order) either the reference index or, for raw slots, the DNA string.
 * with_umi = 0 for table rows (bc_finish / bc_enrich: the random barcode is not part of the key), 1 for the
 * per-read keys of bc_decode_only (random barcode included).  mask = 0 means every counted barcode is present;
 * for enrichment rows pass the row's mask.  idx_out has n_slots entries (-1 for raw / absent slots); str_out is
 * n_slots strings of `str_stride` bytes (empty for indexed / absent slots).  Host only, no GPU work. */
int bc_key_decode(const bc_ctx *ctx, uint64_t key_lo, uint64_t key_hi, uint32_t mask, int with_umi, int32_t *idx_out,
                  char *str_out, uint32_t str_stride);

/* ---- multi-GPU: reads shard over the GPUs of one box; every rank (one bc_ctx per GPU, in one process or one process
 * per GPU) decodes its own shard with bc_submit.  What follows the last batch depends on the counting state:
 *
 *  dense count table (small index-coded key space without a random barcode, e.g. a CRISPR screen): ranks sum their
 *    tables once — bc_dense_counts + an all-reduce, or bc_peer_add inside one process.
 *
 *  hashed keys (any scheme with a random barcode, raw or large key spaces): every (key[, UMI]) record crosses NVLink once.
 *    record -> owner = hash(key without the random barcode) % n_ranks; a scatter kernel writes each record
 *    straight into its owner's receive buffer over NVLink peer memory, and every owner then de-duplicates and counts
 *    the keys it owns, so de-duplication is globally exact.  Afterwards a rank's rows are
 *    its owned keys (disjoint across ranks), its matched / duplicates counters are those of the records it owns and
 *    its other counters those of the reads it decoded: every statistic sums over ranks to the single-GPU value.
 *
 *    set-up   bc_exchange_open(capacity)  this rank's receive buffers, in records; the same capacity on every rank
 *             one process per GPU: bc_exchange_handle -> caller exchanges the BC_IPC_HANDLE_BYTES handles ->
 *                                  bc_exchange_connect(all handles, rank order)
 *             one process:         bc_exchange_connect_local(all contexts, rank order)
 *    per job  bc_submit ...       when the exchange is connected before a job's first batch the exchange is STREAMED: every
 *                                 batch's records leave for their owners right after its decode, on a side stream, under
 *                                 the decode of the next batch (a tile reserves its run in the owner's buffer with one
 *                                 atomic on the owner's receive cursor over NVLink); otherwise they leave in bulk below
 *             bc_exchange_count   -> sent[r] = records of this rank owned by rank r (synchronises)
 *             caller: all-gather the n_ranks x n_ranks matrix; first[r] = sum of sent[r] of the ranks before this one;
 *                     received = sum over ranks of what they send to this one; if any rank's total exceeds the
 *                     capacity, every rank disconnects, re-opens larger, reconnects and calls bc_exchange_count again
 *                     (the job's exchange then starts over, in bulk from the record buffers)
 *             bc_exchange_scatter(first)  bulk: asynchronous on the ctx stream; streamed: nothing left to move
 *             caller: a barrier ordered after every rank's bc_exchange_count / _scatter (a collective, or stream syncs)
 *             bc_exchange_finish(received)
 *    Until bc_exchange_finish, bc_get_counters / bc_finish on a rank of a multi-GPU job fail with BC_ESTATE.  A rank holds
 *    two receive buffers that alternate job by job, so a fast rank may start streaming the next job while a slow one still
 *    counts this one; the ranks must run the same sequence of jobs.  bc_set_option("exchange_bulk", 1) forbids streaming. */
#define BC_IPC_HANDLE_BYTES 64
int bc_exchange_open(bc_ctx *ctx, uint32_t n_ranks, uint32_t rank, uint64_t capacity);
int bc_exchange_handle(bc_ctx *ctx, void *ipc_handle_out);
/* Receive capacity (records) of an exchange that is open and connected to every peer, else 0. */
uint64_t bc_exchange_capacity(const bc_ctx *ctx);
/* Unmaps the other ranks' buffers.  Before re-opening larger, every rank disconnects and the caller synchronises the ranks:
 * a buffer must not be freed while another process still maps it. */
int bc_exchange_disconnect(bc_ctx *ctx);
int bc_exchange_connect(bc_ctx *ctx, const void *ipc_handles);
int bc_exchange_connect_local(bc_ctx *ctx, bc_ctx *const *ranks);
int bc_exchange_count(bc_ctx *ctx, uint64_t *sent);
int bc_exchange_scatter(bc_ctx *ctx, const uint64_t *first);
int bc_exchange_finish(bc_ctx *ctx, uint64_t n_received);
/* One process, several contexts: dst += src for the dense count table (BC_ADD_DENSE_COUNTS) or the dense marginals
 * (BC_ADD_MARGINALS); src may be on another GPU.  Both contexts must be idle.  Synchronises dst. */
enum { BC_ADD_DENSE_COUNTS = 0, BC_ADD_MARGINALS = 1 };
int bc_peer_add(bc_ctx *dst, bc_ctx *src, int what);
/* Without a random barcode and without a dense table, rows can also be merged explicitly: bc_export_rows gives
 * device-resident (key_lo, key_hi, count) arrays, bc_import_rows adds such rows to this rank's. */
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
enum { BC_K_DECODE = 0, BC_K_SCAN = 1, BC_K_INSERT = 2, BC_K_FINISH = 3, BC_K_OTHER = 4, BC_K_ENRICH = 5, BC_K_EXCHANGE = 6,
       BC_N_KERNELS = 7 };
typedef struct {
    uint64_t launches[BC_N_KERNELS];
    double ms[BC_N_KERNELS]; /* CUDA-event time on the ctx stream, only while profiling is on */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t table_capacity, table_entries;
    uint32_t key_bits, wide_keys, dense_table;
    uint32_t deferred_count; /* 1: matched reads are appended to a record buffer and counted at bc_finish /
                                bc_get_counters (partitioned, in shared memory); 0: tables updated read by read */
    uint32_t flushed_global; /* 1: the last flush fell back to the global-memory tables */
    uint64_t specialized_launches, generic_launches; /* decode launches through the NVRTC-specialised / the generic kernel */
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
