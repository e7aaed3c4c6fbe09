// bc_api.cu — the C ABI of include/bc_b200.h: context, device tables, staging, finish/enrichment.
// No CPU fallback lives here: every compute entry point launches the kernels of bc_kernels.cu or fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bc_b200.h"
#include "bc_kernels.h"
#include "bc_jit.h"

using namespace bc;

namespace {

thread_local std::string g_create_error;

struct Staging {
    uint32_t* planes = nullptr;
    uint16_t* read_len = nullptr;
    uint8_t* qual = nullptr;
    size_t cap_planes = 0, cap_len = 0, cap_qual = 0;  // bytes
    cudaEvent_t free_ev = nullptr;    // recorded on the main stream after the last kernel that read the buffer
    cudaEvent_t copied_ev = nullptr;  // recorded on the copy stream after the H2D copies
};

struct KeyField {
    int slot;        // index into cfg.slots
    uint32_t shift;  // first bit in the full key (UMI field at bit 0)
    uint32_t bits;
    bool raw;
};

struct ProfEvent {
    cudaEvent_t a, b;
    int kind;
};

struct ItemBuf {  // structure-of-arrays device buffer of (lo, hi, w) items, grow-only
    unsigned long long *lo = nullptr, *hi = nullptr, *w = nullptr;
    unsigned long long cap = 0;
};

}  // namespace

struct bc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = true;
    DevCfg cfg{};
    std::vector<KeyField> fields;      // per slot, scheme order
    std::vector<int> counted_slots;    // slot indices of counted barcodes in order
    int sample_slot = -1, umi_slot = -1;
    uint32_t max_read_len = 0, W = 0, plane_stride = 0, qual_stride = 0;
    bool quality_on = false;
    // reference accelerators
    uint4* d_refs = nullptr;
    uint32_t* d_tables = nullptr;
    unsigned long long* d_hash_keys = nullptr;
    uint32_t* d_hash_idx = nullptr;
    unsigned long long* d_half = nullptr;
    DevDeep* d_deep = nullptr;
    uint32_t* d_csr = nullptr;
    uint4* d_bref = nullptr;
    DevAux aux{};
    // reads of the batch in flight whose barcode step needs a search (filled by k_decode, drained by k_resolve)
    uint2* d_def_items = nullptr;
    uint32_t* d_def_count = nullptr;  // two counters, used by alternate batches
    uint32_t def_parity = 0;
    uint64_t def_cap = 0;
    // multi-GPU exchange at the flush (bc_exchange_*): this rank's receive buffer [lo: cap][hi: cap], the peers' buffers
    // (CUDA IPC mappings, or plain pointers of contexts in this process), and what the last exchange delivered
    unsigned long long* d_xrecv = nullptr;
    unsigned long long xcap = 0;
    uint32_t x_ranks = 0, x_rank = 0;
    unsigned long long* x_peer[kMaxRanks] = {nullptr};
    bool x_peer_ipc[kMaxRanks] = {false};
    uint32_t* d_xcursor = nullptr;         // kMaxRanks cursors / owner counts
    uint32_t h_xcur[kMaxRanks] = {0};
    unsigned long long x_sent[kMaxRanks] = {0};
    unsigned long long x_local_valid = 0;  // records of this rank that were not holes at the last bc_exchange_count
    int x_state = 0;                       // 0: records not exchanged; 1: counted; 2: scattered; 3: finished (rows valid)
    unsigned long long x_received = 0;
    // streamed exchange: the receive allocation holds TWO buffers that alternate job by job (a fast rank may stream the next
    // job's records while this one still counts the last job's) and, behind them, one 64-bit receive cursor per buffer that
    // the senders advance over NVLink.  A batch's records leave on x_stream right after its decode.
    cudaStream_t x_stream = nullptr;
    cudaEvent_t x_ev = nullptr;
    uint32_t x_epoch = 0;                  // jobs finished since the buffers were opened: parity picks the buffer
    int x_mode = 0;                        // of the job in progress: 0 undecided (nothing submitted), 1 streamed, 2 bulk
    bool x_overflowed = false, x_dirty = false, opt_exchange_bulk = false;
    unsigned long long* d_xsent = nullptr; // kMaxRanks counters: records sent to each owner by the streamed scatters
    unsigned int* d_xoverflow = nullptr;
    // options (bc_set_option): measurement / test switches of the flush
    bool opt_flush_global = false, opt_flush_two_stage = false;
    uint32_t cfg_flags = 0;
    // the decode kernel specialised for this run by NVRTC (bc_jit.cu); kernel == nullptr: the generic kernel is used
    JitDecode jit{};
    std::string jit_note;
    uint64_t jit_launches = 0, generic_launches = 0;
    // K4: dense enrichment marginals (valid for the current rows while marg_valid)
    MargPlan marg{};
    bool marg_dense = false, marg_valid = false;
    unsigned long long* d_marg = nullptr;
    // deferred counting (bc_partition.cu): matched reads append their packed key to `rec`; bc_finish / bc_get_counters
    // de-duplicate and count the whole buffer partition by partition in shared memory
    bool deferred = false;
    ItemBuf rec;                          // the record buffer: slot (cursor + read index) per read, kEmpty = hole
    unsigned long long rec_n = 0;            // slots used: every appended batch takes exactly n_reads of them
    bool stripes_dirty = false;              // k_decode's striped counters hold counts not yet folded into d_counters
    ItemBuf left;                         // records of the hot partitions set aside by the one-stage flush
    ItemBuf part, w1, w2, tmp;            // scratch: partitioned records, (key, weight) items before / after partitioning,
                                          // and the output of the first radix level when two are needed
    uint32_t* d_l1 = nullptr;             // histogram / starts / cursors of the first radix level
    ItemBuf imp;                          // rows imported from other ranks (bc_import_rows), added before stage B
    unsigned long long imp_n = 0;
    uint32_t *d_hist = nullptr, *d_starts = nullptr, *d_cursor = nullptr;
    uint32_t *d_hot = nullptr, *d_hot_n = nullptr;  // partitions the keyed reduce handed to the two-pass kernel
    uint32_t* d_big = nullptr;            // k_gather_big's list of oversized partitions: {start, size, place} x part_cap
    unsigned long long part_cap = 0;
    FlushStats* d_flush = nullptr;
    unsigned long long last_valid = 0, last_unique = 0;  // of the last flush
    unsigned long long dup_applied = 0;   // duplicates already moved from "matched" to "duplicates" by earlier flushes
    bool flushed_global = false;          // the last flush went through the global-memory tables
    uint32_t flush_stages = 0;            // partition / reduce stages of the last shared-memory flush (1 or 2)
    // counting state: map (key -> count) and, with a random barcode, the (key, UMI) set
    Tables tables{};
    unsigned long long expected_reads = 0;
    unsigned long long entries_upper = 0;  // host-side upper bound of entries added to either table
    unsigned long long imported_rows = 0;  // rows merged in from other ranks (they add keys without bumping "matched")
    unsigned long long* d_counters = nullptr;  // BC_N_COUNTERS + 2 (then: map entries, set entries)
    unsigned long long* d_stripes = nullptr;   // k_decode's striped copies of them, folded in after every launch
    // staging for host batches
    Staging staging[2];
    // host batches in their transfer form (bc_submit_wire): two arenas that alternate like `staging`, and the one set of
    // expanded arrays the decode kernel reads (written and read on the main stream, so one set is enough)
    struct WireStage {
        unsigned char* buf = nullptr;
        size_t cap = 0;
        cudaEvent_t free_ev = nullptr, copied_ev = nullptr;
    } wire[2];
    int wire_cur = 0;
    uint32_t* x_planes = nullptr;
    uint16_t* x_len = nullptr;
    uint8_t* x_qual = nullptr;
    size_t x_cap_planes = 0, x_cap_len = 0, x_cap_qual = 0;
    int cur = 0;
    bool copies_pending = false;
    // scratch for the test hooks
    // final rows: device buffers (grow-only, reused by every finish / export) and their pinned host mirror
    unsigned long long *d_row_lo = nullptr, *d_row_hi = nullptr, *d_row_cnt = nullptr, *d_row_n = nullptr;
    unsigned long long row_cap = 0, n_rows = 0;
    bool rows_valid = false;
    uint64_t *h_row_lo = nullptr, *h_row_hi = nullptr, *h_row_cnt = nullptr;
    unsigned long long h_row_cap = 0;
    // profiling
    bool profiling = false;
    std::vector<ProfEvent> prof_events;
    bc_profile prof{};
    std::string err;
};

namespace {

int fail(bc_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

#define CK(ctx, call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return fail(ctx, BC_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                            \
    } while (0)

unsigned long long pow2_at_least(unsigned long long v) {
    unsigned long long p = 1024;
    while (p < v) p <<= 1;
    return p;
}

uint32_t bits_for(uint32_t n) {  // bits to hold indices 0..n-1, at least 1
    uint32_t b = 1;
    while ((1ull << b) < n) b++;
    return b;
}

struct ProfScope {  // CUDA events around one launch, only while profiling is on
    bc_ctx* ctx;
    int kind;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t on;
    ProfScope(bc_ctx* c, int k, cudaStream_t stream = nullptr) : ctx(c), kind(k), on(stream ? stream : c->stream) {
        ctx->prof.launches[kind]++;
        if (ctx->profiling) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, on);
        }
    }
    ~ProfScope() {
        if (a) {
            cudaEventRecord(b, on);
            ctx->prof_events.push_back({a, b, kind});
        }
    }
};

void drain_profile(bc_ctx* ctx) {
    for (ProfEvent& p : ctx->prof_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) ctx->prof.ms[p.kind] += ms;
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    ctx->prof_events.clear();
}

void free_table(DevTable& t) {
    if (t.data) cudaFree(t.data);
    t.data = nullptr;
}

size_t table_bytes(const DevTable& t) { return (size_t)t.cap * table_stride(t.kind, t.wide) * sizeof(unsigned long long); }

int clear_table(bc_ctx* ctx, DevTable& t) {
    if (!t.data) return BC_OK;
    if (t.kind == 1) {  // {key = empty, count = 0} slots: one pass of 16-byte stores
        ProfScope p(ctx, BC_K_OTHER);
        CK(ctx, launch_clear_map(t, ctx->stream));
    } else {
        CK(ctx, cudaMemsetAsync(t.data, t.kind == 0 ? 0 : 0xFF, table_bytes(t), ctx->stream));
    }
    return BC_OK;
}

// kind: 0 dense (cap = number of counters), 1 hash map, 2 hash set
int alloc_table(bc_ctx* ctx, DevTable& t, int kind, int wide, unsigned long long cap, unsigned long long* n_entries) {
    t = DevTable{};
    t.kind = kind;
    t.wide = wide;
    t.cap = cap;
    t.n_entries = n_entries;
    CK(ctx, cudaMalloc(&t.data, table_bytes(t)));
    return clear_table(ctx, t);
}

// slots for `entries` keys at load factor <= 0.6
unsigned long long slots_for(unsigned long long entries) {
    const unsigned long long c = entries + entries * 2 / 3 + 1024;
    return c;
}

void drop_rows(bc_ctx* ctx) {
    ctx->n_rows = 0;
    ctx->rows_valid = false;
    ctx->marg_valid = false;
}

int reserve_rows(bc_ctx* ctx, unsigned long long n, bool wide) {
    if (n <= ctx->row_cap && (!wide || ctx->d_row_hi)) return BC_OK;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_row_lo) cudaFree(ctx->d_row_lo);
    if (ctx->d_row_hi) cudaFree(ctx->d_row_hi);
    if (ctx->d_row_cnt) cudaFree(ctx->d_row_cnt);
    ctx->d_row_lo = ctx->d_row_hi = ctx->d_row_cnt = nullptr;
    ctx->row_cap = 0;
    const unsigned long long cap = std::max<unsigned long long>(n + n / 8, 1024);
    CK(ctx, cudaMalloc(&ctx->d_row_lo, cap * sizeof(unsigned long long)));
    CK(ctx, cudaMalloc(&ctx->d_row_cnt, cap * sizeof(unsigned long long)));
    if (wide) CK(ctx, cudaMalloc(&ctx->d_row_hi, cap * sizeof(unsigned long long)));
    ctx->row_cap = cap;
    return BC_OK;
}

void free_items(ItemBuf& b) {
    if (b.lo) cudaFree(b.lo);
    if (b.hi) cudaFree(b.hi);
    if (b.w) cudaFree(b.w);
    b = ItemBuf{};
}

// room for n items; with keep, the first `used` items survive a reallocation
int reserve_items(bc_ctx* ctx, ItemBuf& b, unsigned long long n, bool wide, bool weighted, unsigned long long used = 0) {
    if (n <= b.cap && (!wide || b.hi) && (!weighted || b.w)) return BC_OK;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned long long cap = std::max<unsigned long long>(std::max(n, b.cap) + (used ? n / 2 : n / 16), 1024);
    ItemBuf nb;
    nb.cap = cap;
    CK(ctx, cudaMalloc(&nb.lo, cap * sizeof(unsigned long long)));
    if (wide) CK(ctx, cudaMalloc(&nb.hi, cap * sizeof(unsigned long long)));
    if (weighted) CK(ctx, cudaMalloc(&nb.w, cap * sizeof(unsigned long long)));
    if (used) {
        CK(ctx, cudaMemcpy(nb.lo, b.lo, used * sizeof(unsigned long long), cudaMemcpyDeviceToDevice));
        if (wide && b.hi) CK(ctx, cudaMemcpy(nb.hi, b.hi, used * sizeof(unsigned long long), cudaMemcpyDeviceToDevice));
        if (weighted && b.w) CK(ctx, cudaMemcpy(nb.w, b.w, used * sizeof(unsigned long long), cudaMemcpyDeviceToDevice));
    }
    free_items(b);
    b = nb;
    return BC_OK;
}

int reserve_parts(bc_ctx* ctx, unsigned long long n_parts) {
    if (n_parts + 1 <= ctx->part_cap) return BC_OK;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_hist) cudaFree(ctx->d_hist);
    if (ctx->d_starts) cudaFree(ctx->d_starts);
    if (ctx->d_cursor) cudaFree(ctx->d_cursor);
    if (ctx->d_hot) cudaFree(ctx->d_hot);
    if (ctx->d_big) cudaFree(ctx->d_big);
    ctx->d_hist = ctx->d_starts = ctx->d_cursor = ctx->d_hot = ctx->d_big = nullptr;
    ctx->part_cap = 0;
    const unsigned long long cap = n_parts + n_parts / 8 + 1024;
    CK(ctx, cudaMalloc(&ctx->d_hot, cap * sizeof(uint32_t)));
    CK(ctx, cudaMalloc(&ctx->d_big, cap * 4 * sizeof(uint32_t)));
    if (!ctx->d_hot_n) CK(ctx, cudaMalloc(&ctx->d_hot_n, sizeof(uint32_t)));
    CK(ctx, cudaMalloc(&ctx->d_hist, cap * sizeof(uint32_t)));
    CK(ctx, cudaMalloc(&ctx->d_starts, cap * sizeof(uint32_t)));
    CK(ctx, cudaMalloc(&ctx->d_cursor, cap * sizeof(uint32_t)));
    ctx->part_cap = cap;
    return BC_OK;
}

// Room in the record buffer for `extra` more slots (the buffer grows, keeping what it holds)
int reserve_records(bc_ctx* ctx, unsigned long long extra) {
    const bool wide = ctx->cfg.wide != 0;
    if (ctx->rec_n + extra <= ctx->rec.cap && ctx->rec.lo) return BC_OK;
    return reserve_items(ctx, ctx->rec, ctx->rec_n + extra, wide, false, ctx->rec_n);
}

// first append of a job: size the buffer for the whole job (bc_create's expected_reads) plus `slack`
int prime_records(bc_ctx* ctx, unsigned long long slack) {
    if (ctx->rec.lo) return BC_OK;
    return reserve_records(ctx, ctx->expected_reads + ctx->expected_reads / 8 + slack);
}

RecOut rec_out(const bc_ctx* ctx) { return RecOut{ctx->rec.lo, ctx->cfg.wide ? ctx->rec.hi : nullptr, ctx->rec_n}; }

// k_decode adds its outcome counters to striped copies; they are folded into d_counters when somebody reads them
int fold_counters(bc_ctx* ctx) {
    if (!ctx->stripes_dirty) return BC_OK;
    ProfScope p(ctx, BC_K_OTHER);
    CK(ctx, launch_fold_counters(ctx->d_stripes, ctx->d_counters, ctx->stream));
    ctx->stripes_dirty = false;
    return BC_OK;
}

// Smallest integer sum S with fl32(fl32(S) / fl32(len)) >= min_quality: the reference's f32 mean test
// (parse.rs:352-355) as an exact integer threshold (Q12).  255*len+1 when no sum passes.
uint32_t quality_threshold(uint32_t len, float min_quality) {
    const uint32_t top = 255u * len;
    for (uint32_t s = 0; s <= top; s++) {
        volatile float avg = static_cast<float>(s) / static_cast<float>(len);
        if (!(avg < min_quality)) return s;
    }
    return top + 1;
}

// A batch carries its own geometry: W = plane_stride / 3 words per plane, i.e. reads of up to 32 W bases.  The context's
// max_read_len is only the default; a host that meets a longer read packs that batch wider (up to BC_MAX_READ_LEN).
int validate_batch(bc_ctx* ctx, const bc_batch* b) {
    if (!b) return fail(ctx, BC_EINVAL, "batch is NULL");
    if (b->n_reads == 0) return BC_OK;
    if (!b->planes || !b->read_len) return fail(ctx, BC_EINVAL, "batch.planes / batch.read_len is NULL");
    const uint32_t W = b->plane_stride / 3;
    if (W == 0 || b->plane_stride != ((3 * W) | 1u) || 32 * W > BC_MAX_READ_LEN)
        return fail(ctx, BC_EINVAL, "batch.plane_stride %u is not bc_plane_stride(n) of a read length up to %d", b->plane_stride, BC_MAX_READ_LEN);
    if (32 * W < ctx->cfg.L)
        return fail(ctx, BC_EINVAL, "batch.plane_stride %u: reads of at most %u bases cannot hold the %u-base scheme", b->plane_stride, 32 * W, ctx->cfg.L);
    if (ctx->quality_on) {
        if (!b->qual) return fail(ctx, BC_EINVAL, "min_quality > 0 but batch.qual is NULL");
        if ((b->qual_stride & 3u) || b->qual_stride < 32 * (W - 1) + 1)
            return fail(ctx, BC_EINVAL, "batch.qual_stride %u does not go with plane_stride %u (use bc_qual_stride(n))", b->qual_stride, b->plane_stride);
    }
    if (b->location != BC_LOC_HOST && b->location != BC_LOC_DEVICE) return fail(ctx, BC_EINVAL, "batch.location");
    return BC_OK;
}

// Make the batch visible to the kernels: device batches are used in place, host batches go through the
// double-buffered staging on the copy stream (H2D of batch i+1 overlaps the kernels of batch i).
int stage_batch(bc_ctx* ctx, const bc_batch* b, BatchView* view) {
    view->n_reads = b->n_reads;
    view->plane_stride = b->plane_stride;
    view->qual_stride = b->qual_stride;
    view->W = b->plane_stride / 3;
    view->rep_chunks = (32 * view->W - ctx->cfg.L + 31) / 32;
    const bool want_qual = ctx->quality_on;
    if (b->location == BC_LOC_DEVICE) {
        view->planes = b->planes;
        view->read_len = b->read_len;
        view->qual = want_qual ? b->qual : nullptr;
        return BC_OK;
    }
    Staging& s = ctx->staging[ctx->cur];
    ctx->cur ^= 1;
    if (!s.free_ev) {
        CK(ctx, cudaEventCreateWithFlags(&s.free_ev, cudaEventDisableTiming));
        CK(ctx, cudaEventCreateWithFlags(&s.copied_ev, cudaEventDisableTiming));
    }
    const size_t pb = (size_t)b->n_reads * b->plane_stride * sizeof(uint32_t);
    const size_t lb = (size_t)b->n_reads * sizeof(uint16_t);
    const size_t qb = want_qual ? (size_t)b->n_reads * b->qual_stride : 0;
    if (s.cap_planes < pb || s.cap_len < lb || s.cap_qual < qb) {
        CK(ctx, cudaEventSynchronize(s.free_ev));
        if (s.planes) cudaFree(s.planes);
        if (s.read_len) cudaFree(s.read_len);
        if (s.qual) cudaFree(s.qual);
        s.planes = nullptr; s.read_len = nullptr; s.qual = nullptr;
        s.cap_planes = std::max(s.cap_planes, pb + pb / 8);
        s.cap_len = std::max(s.cap_len, lb + lb / 8);
        s.cap_qual = std::max(s.cap_qual, qb + qb / 8);
        CK(ctx, cudaMalloc(&s.planes, s.cap_planes));
        CK(ctx, cudaMalloc(&s.read_len, s.cap_len));
        if (want_qual) CK(ctx, cudaMalloc(&s.qual, s.cap_qual));
    }
    CK(ctx, cudaStreamWaitEvent(ctx->copy_stream, s.free_ev, 0));
    CK(ctx, cudaMemcpyAsync(s.planes, b->planes, pb, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(ctx, cudaMemcpyAsync(s.read_len, b->read_len, lb, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (want_qual) CK(ctx, cudaMemcpyAsync(s.qual, b->qual, qb, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(ctx, cudaEventRecord(s.copied_ev, ctx->copy_stream));
    CK(ctx, cudaStreamWaitEvent(ctx->stream, s.copied_ev, 0));
    ctx->prof.h2d_bytes += pb + lb + qb;
    ctx->copies_pending = true;
    view->planes = s.planes;
    view->read_len = s.read_len;
    view->qual = s.qual;
    // the caller records s.free_ev after its kernels
    return 1;  // staged
}

int release_staging(bc_ctx* ctx, int staged) {
    if (staged == 1) {
        Staging& s = ctx->staging[ctx->cur ^ 1];
        CK(ctx, cudaEventRecord(s.free_ev, ctx->stream));
    }
    return BC_OK;
}

int grow_table(bc_ctx* ctx, DevTable& t, unsigned long long need_entries) {
    if (t.kind == 0 || !t.data) return BC_OK;
    if (slots_for(need_entries) <= t.cap) return BC_OK;
    unsigned long long cap = t.cap;
    while (cap < slots_for(need_entries)) cap *= 2;
    DevTable bigger;
    int rc = alloc_table(ctx, bigger, t.kind, t.wide, cap, t.n_entries);
    if (rc != BC_OK) return rc;
    {
        ProfScope p(ctx, BC_K_OTHER);
        CK(ctx, launch_rehash(t, bigger, ctx->stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    free_table(t);
    t = bigger;
    return BC_OK;
}

// keep the load factor of the hash tables <= 0.6 (entries are bounded by the reads submitted so far)
int ensure_capacity(bc_ctx* ctx, unsigned long long incoming) {
    Tables& T = ctx->tables;
    if (ctx->deferred || (T.map.kind == 0 && !T.has_set)) return BC_OK;
    ctx->entries_upper += incoming;
    const unsigned long long smallest = std::min(T.map.kind ? T.map.cap : ~0ull, T.has_set ? T.set.cap : ~0ull);
    if (slots_for(ctx->entries_upper) <= smallest) return BC_OK;
    {
        const int rc_ = fold_counters(ctx);
        if (rc_ != BC_OK) return rc_;
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long n[2] = {0, 0};
    CK(ctx, cudaMemcpy(n, ctx->d_counters + BC_N_COUNTERS, sizeof n, cudaMemcpyDeviceToHost));
    ctx->entries_upper = std::max(n[0], n[1]) + incoming;
    int rc = grow_table(ctx, T.map, n[0] + incoming);
    if (rc == BC_OK && T.has_set) rc = grow_table(ctx, T.set, n[1] + incoming);
    return rc;
}

}  // namespace

extern "C" {

uint32_t bc_plane_words(uint32_t max_read_len) { return (max_read_len + 31) / 32; }
uint32_t bc_plane_stride(uint32_t max_read_len) { return (3 * bc_plane_words(max_read_len)) | 1u; }
uint32_t bc_qual_stride(uint32_t max_read_len) { return (((max_read_len + 3) / 4) | 1u) * 4; }

const char* bc_last_error(const bc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int bc_device_of(const bc_ctx* ctx) { return ctx ? ctx->device : -1; }

void bc_destroy(bc_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    drain_profile(ctx);
    if (ctx->d_row_lo) cudaFree(ctx->d_row_lo);
    if (ctx->d_row_hi) cudaFree(ctx->d_row_hi);
    if (ctx->d_row_cnt) cudaFree(ctx->d_row_cnt);
    if (ctx->h_row_lo) cudaFreeHost(ctx->h_row_lo);
    if (ctx->h_row_hi) cudaFreeHost(ctx->h_row_hi);
    if (ctx->h_row_cnt) cudaFreeHost(ctx->h_row_cnt);
    if (ctx->d_row_n) cudaFree(ctx->d_row_n);
    free_table(ctx->tables.map);
    free_table(ctx->tables.set);
    free_items(ctx->rec);
    free_items(ctx->part);
    free_items(ctx->w1);
    free_items(ctx->w2);
    free_items(ctx->tmp);
    free_items(ctx->left);
    if (ctx->d_l1) cudaFree(ctx->d_l1);
    free_items(ctx->imp);
    if (ctx->d_hist) cudaFree(ctx->d_hist);
    if (ctx->d_starts) cudaFree(ctx->d_starts);
    if (ctx->d_cursor) cudaFree(ctx->d_cursor);
    if (ctx->d_hot) cudaFree(ctx->d_hot);
    if (ctx->d_hot_n) cudaFree(ctx->d_hot_n);
    if (ctx->d_big) cudaFree(ctx->d_big);
    if (ctx->d_flush) cudaFree(ctx->d_flush);
    for (Staging& s : ctx->staging) {
        if (s.planes) cudaFree(s.planes);
        if (s.read_len) cudaFree(s.read_len);
        if (s.qual) cudaFree(s.qual);
        if (s.free_ev) cudaEventDestroy(s.free_ev);
        if (s.copied_ev) cudaEventDestroy(s.copied_ev);
    }
    for (bc_ctx::WireStage& s : ctx->wire) {
        if (s.buf) cudaFree(s.buf);
        if (s.free_ev) cudaEventDestroy(s.free_ev);
        if (s.copied_ev) cudaEventDestroy(s.copied_ev);
    }
    if (ctx->x_planes) cudaFree(ctx->x_planes);
    if (ctx->x_len) cudaFree(ctx->x_len);
    if (ctx->x_qual) cudaFree(ctx->x_qual);
    if (ctx->d_refs) cudaFree(ctx->d_refs);
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    if (ctx->d_hash_keys) cudaFree(ctx->d_hash_keys);
    if (ctx->d_hash_idx) cudaFree(ctx->d_hash_idx);
    if (ctx->d_half) cudaFree(ctx->d_half);
    if (ctx->d_deep) cudaFree(ctx->d_deep);
    if (ctx->d_csr) cudaFree(ctx->d_csr);
    if (ctx->d_bref) cudaFree(ctx->d_bref);
    if (ctx->d_def_items) cudaFree(ctx->d_def_items);
    if (ctx->d_def_count) cudaFree(ctx->d_def_count);
    for (uint32_t r = 0; r < ctx->x_ranks; r++)
        if (ctx->x_peer_ipc[r] && ctx->x_peer[r]) cudaIpcCloseMemHandle(ctx->x_peer[r]);
    if (ctx->d_xrecv) cudaFree(ctx->d_xrecv);
    if (ctx->d_xcursor) cudaFree(ctx->d_xcursor);
    if (ctx->d_xsent) cudaFree(ctx->d_xsent);
    if (ctx->d_xoverflow) cudaFree(ctx->d_xoverflow);
    if (ctx->x_stream) {
        cudaStreamSynchronize(ctx->x_stream);
        cudaStreamDestroy(ctx->x_stream);
    }
    if (ctx->x_ev) cudaEventDestroy(ctx->x_ev);
    if (ctx->d_marg) cudaFree(ctx->d_marg);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_stripes) cudaFree(ctx->d_stripes);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

// Host copies of the reference-set accelerators, built with the run constants and uploaded by bc_create
struct HostRefs {
    std::vector<uint4> refs;
    std::vector<unsigned long long> hkeys;
    std::vector<uint32_t> hidx;
    std::vector<unsigned long long> half;
    std::vector<DevDeep> deep;
    std::vector<uint32_t> csr;
    std::vector<uint4> bref;
    size_t table_u16 = 0;
};

// bc_config -> the run constants (ctx->cfg: what the kernels read; also what bc_jit.cu specialises the decode kernel on),
// the key layout and the host side of the look-up structures.  No CUDA call in here.
static int build_run_constants(const bc_config* cfg, bc_ctx* ctx, HostRefs& H) {
#define FAILC(code, ...) return fail(nullptr, code, __VA_ARGS__)
    const uint32_t L = cfg->template_len;
    DevCfg& d = ctx->cfg;
    d.L = L;
    d.TW = (L + 31) / 32;
    d.n_slots = cfg->n_slots;
    d.max_const_err = std::min<uint32_t>(cfg->max_const_err, L);  // a cap above the scheme length admits nothing more
    ctx->cfg_flags = cfg->flags;
    ctx->max_read_len = cfg->max_read_len;
    ctx->W = bc_plane_words(cfg->max_read_len);
    ctx->plane_stride = bc_plane_stride(cfg->max_read_len);
    ctx->qual_stride = bc_qual_stride(cfg->max_read_len);
    ctx->quality_on = cfg->min_quality > 0.0f;

    // ---- template planes (info.rs:283-299): barcode positions are free, 'N' outside barcodes is format-N
    std::vector<int> owner(L, -1);
    for (uint32_t s = 0; s < cfg->n_slots; s++) {
        const bc_slot& S = cfg->slots[s];
        if (S.kind != 'S' && S.kind != 'B' && S.kind != 'R') FAILC(BC_EINVAL, "slot %u: kind must be 'S', 'B' or 'R'", s);
        if (S.len == 0 || (uint32_t)S.offset + S.len > L) FAILC(BC_EINVAL, "slot %u lies outside the template", s);
        if (S.len > BC_MAX_REF_LEN) FAILC(BC_EUNSUPPORTED, "slot %u: barcodes longer than %d bases", s, BC_MAX_REF_LEN);
        for (uint32_t p = S.offset; p < (uint32_t)S.offset + S.len; p++) {
            if (owner[p] >= 0) FAILC(BC_EINVAL, "slots %d and %u overlap", owner[p], s);
            owner[p] = (int)s;
        }
        if (S.kind == 'S') {
            if (ctx->sample_slot >= 0) FAILC(BC_EINVAL, "two sample barcodes (the reference rejects this too, info.rs:308)");
            ctx->sample_slot = (int)s;
        } else if (S.kind == 'R') {
            if (ctx->umi_slot >= 0) FAILC(BC_EINVAL, "two random barcodes (the reference rejects this too, info.rs:308)");
            if (S.n_ref) FAILC(BC_EINVAL, "the random barcode cannot have a reference set");
            ctx->umi_slot = (int)s;
        } else {
            ctx->counted_slots.push_back((int)s);
        }
    }
    for (uint32_t p = 0; p < L; p++) {
        const char ch = cfg->template_chars[p];
        const uint32_t w = p >> 5, bit = 1u << (p & 31);
        if (owner[p] >= 0) continue;
        switch (ch) {
            case 'A': d.t_cm[w] |= bit; break;
            case 'C': d.t_cm[w] |= bit; d.t_lo[w] |= bit; break;
            case 'G': d.t_cm[w] |= bit; d.t_hi[w] |= bit; break;
            case 'T': d.t_cm[w] |= bit; d.t_lo[w] |= bit; d.t_hi[w] |= bit; break;
            case 'N': d.t_fn[w] |= bit; d.has_fn = 1; break;
            default:
                FAILC(BC_EUNSUPPORTED, "template position %u holds '%c': only upper-case A/C/G/T/N constants are supported "
                      "(lower case never matches in the reference's repair step, Q11)", p, ch);
        }
    }

    {  // locate prefilter: the template word with the most constant bases
        int bestw = 0, bestc = -1;
        for (uint32_t w = 0; w < d.TW; w++) {
            const int c = __builtin_popcount(d.t_cm[w]);
            if (c > bestc) {
                bestc = c;
                bestw = (int)w;
            }
        }
        d.pivot = (uint32_t)bestw;
        // exact-match prefilter (phase A): four constant positions of that word per base, spread over the word; a base with
        // fewer than four repeats its last one
        d.xpivot = (uint32_t)bestw;
        for (uint32_t b = 0; b < 4; b++) {
            std::vector<uint32_t> pos;
            for (uint32_t q = 0; q < 32; q++)
                if (((d.t_cm[bestw] >> q) & 1u) && ((((d.t_lo[bestw] >> q) & 1u) | (((d.t_hi[bestw] >> q) & 1u) << 1)) == b)) pos.push_back(q);
            if (pos.empty()) continue;
            d.xs_has |= 1u << b;
            for (uint32_t i = 0; i < 4; i++) d.xs_sh[b][i] = pos[std::min<size_t>(pos.size() - 1, i * pos.size() / 4)];
        }
        d.bs_ok = d.max_const_err <= 15 ? 1u : 0u;
        d.bs_k = 15u - std::min<uint32_t>(d.max_const_err, 15u);
        for (uint32_t q = 0; q < 32; q++) {
            if (!((d.t_cm[bestw] >> q) & 1u)) continue;
            const uint32_t b = ((d.t_lo[bestw] >> q) & 1u) | (((d.t_hi[bestw] >> q) & 1u) << 1);
            const uint32_t i = d.pv_n[b]++;
            d.pv_sh4[b][i >> 2] |= q << (8 * (i & 3));
        }
        // Static-block variant over two template words: pick the pair (w, w + 1) with the most positions left after
        // rounding every (word, base) group down to blocks of four (two blocks at most).  Used when it keeps at least
        // as many positions as the single pivot word has (never a weaker filter than the one it replaces).
        // Measured: with a single 32-offset chunk per read (CRISPR: 16 windows) the static blocks win (0.513 -> 0.480 ms per
        // batch: no per-base loop set-up, no remainder code); with two or more chunks (DEL: 65 windows) the per-base
        // loops over one word win (0.531 vs 0.543 ms: one plane word less to load and combine).
        const bool want_two = cfg->max_read_len - L + 1 <= 32;
        if (d.bs_ok && d.TW >= 2 && want_two) {
            auto kept = [&](uint32_t w, uint32_t b) {
                uint32_t n = 0;
                for (uint32_t q = 0; q < 32; q++)
                    if (((d.t_cm[w] >> q) & 1u) && ((((d.t_lo[w] >> q) & 1u) | (((d.t_hi[w] >> q) & 1u) << 1)) == b)) n++;
                return std::min<uint32_t>(n / 4, 2) * 4;
            };
            int best2 = -1;
            uint32_t best_n = 0;
            for (uint32_t w = 0; w + 1 < d.TW; w++) {
                uint32_t n = 0;
                for (uint32_t b = 0; b < 4; b++) n += kept(w, b) + kept(w + 1, b);
                if (n > best_n) {
                    best_n = n;
                    best2 = (int)w;
                }
            }
            if (best2 >= 0 && best_n >= (uint32_t)bestc && best_n >= 16) {
                d.bs_two = 1;
                d.pivot = (uint32_t)best2;
                // no more positions than the single pivot word would have used (the kernel is bound by the alu pipe:
                // every position is a funnel shift and 2.25 LOP3 per 32 offsets): drop whole blocks beyond that
                uint32_t blocks[2][4], total = 0;
                for (uint32_t ww = 0; ww < 2; ww++)
                    for (uint32_t b = 0; b < 4; b++) total += (blocks[ww][b] = kept(best2 + ww, b) / 4);
                const uint32_t want = ((uint32_t)bestc + 3) / 4;
                for (int ww = 1; ww >= 0 && total > want; ww--)
                    for (int b = 3; b >= 0 && total > want; b--)
                        while (blocks[ww][b] > 0 && total > want) {
                            blocks[ww][b]--;
                            total--;
                        }
                for (uint32_t ww = 0; ww < 2; ww++)
                    for (uint32_t b = 0; b < 4; b++) {
                        const uint32_t w = best2 + ww, keep = blocks[ww][b] * 4;
                        d.bs2_n[ww][b] = keep / 4;
                        uint32_t i = 0;
                        for (uint32_t q = 0; q < 32 && i < keep; q++)
                            if (((d.t_cm[w] >> q) & 1u) && ((((d.t_lo[w] >> q) & 1u) | (((d.t_hi[w] >> q) & 1u) << 1)) == b))
                                d.bs2_sh[ww][b][i++] = q;
                    }
            }
        }
    }

    // ---- quality runs (parse.rs:340-374): maximal runs of one region code; a non-constant run is tested only when
    // another code follows it inside the walked range (Q8); format-N shortens the code string (Q9)
    if (ctx->quality_on) {
        uint32_t i = 0;
        while (i < cfg->region_len) {
            uint32_t j = i;
            while (j < cfg->region_len && cfg->region_codes[j] == cfg->region_codes[i]) j++;
            if (cfg->region_codes[i] != 'C' && j < cfg->region_len) {
                if (d.n_qruns == kMaxQRuns) FAILC(BC_EUNSUPPORTED, "more than %d quality-tested runs", kMaxQRuns);
                // the kernel sums the raw Phred+33 bytes, so the per-byte offset goes into the threshold
                {
                    const uint32_t len = j - i, rem = len & 3u;
                    DevQRun q{};
                    q.off = (uint16_t)i;
                    q.len = (uint16_t)len;
                    q.n_words = (uint16_t)((len + 3) / 4);
                    q.tail_mask = rem ? (1u << (8 * rem)) - 1u : 0xFFFFFFFFu;
                    q.thresh = quality_threshold(len, cfg->min_quality) + 33u * len;
                    d.qruns[d.n_qruns++] = q;
                }
            }
            i = j;
        }
    }

    // ---- processing order (parse.rs:451, 483, 512) and key layout: [UMI | counted 1..k | sample]
    {
        uint32_t n = 0;
        if (ctx->sample_slot >= 0) d.order[n++] = (uint8_t)ctx->sample_slot;
        for (int s : ctx->counted_slots) d.order[n++] = (uint8_t)s;
        if (ctx->umi_slot >= 0) d.order[n++] = (uint8_t)ctx->umi_slot;
    }
    ctx->fields.resize(cfg->n_slots);
    uint32_t shift = 0;
    auto add_field = [&](int s) {
        const bc_slot& S = cfg->slots[s];
        const bool raw = S.n_ref == 0;
        const uint32_t bits = raw ? 3u * S.len : bits_for(S.n_ref);
        ctx->fields[s] = KeyField{s, shift, bits, raw};
        d.slots[s].key_shift = (uint16_t)shift;
        d.slots[s].key_bits = (uint16_t)bits;
        shift += bits;
    };
    if (ctx->umi_slot >= 0) {
        add_field(ctx->umi_slot);
        d.has_umi = 1;
        d.umi_bits = shift;
    }
    for (int s : ctx->counted_slots) add_field(s);
    if (ctx->sample_slot >= 0) add_field(ctx->sample_slot);
    if (shift > BC_MAX_KEY_BITS)
        FAILC(BC_EUNSUPPORTED, "packed key needs %u bits (> %d): too many / too long raw barcodes", shift, BC_MAX_KEY_BITS);
    d.key_bits = shift;
    d.wide = shift > 63;

    // ---- reference sets -> {lo,hi,nm,len} words, direct tables, exact-match hashes
    std::vector<uint4>& refs = H.refs;
    std::vector<unsigned long long>& hkeys = H.hkeys;
    std::vector<uint32_t>& hidx = H.hidx;
    std::vector<unsigned long long>& half = H.half;
    std::vector<DevDeep>& deep = H.deep;
    std::vector<uint32_t>& csr = H.csr;
    std::vector<uint4>& bref = H.bref;
    size_t& table_u16 = H.table_u16;
    for (uint32_t s = 0; s < cfg->n_slots; s++) {
        const bc_slot& S = cfg->slots[s];
        DevSlot& D = d.slots[s];
        D.offset = S.offset;
        D.len = S.len;
        D.max_err = S.max_err;
        D.kind = S.kind;
        D.n_ref = S.n_ref;
        D.mode = MODE_RAW;
        if (S.n_ref == 0) continue;
        if (!S.ref_seqs) FAILC(BC_EINVAL, "slot %u: ref_seqs is NULL", s);
        D.ref_off = (uint32_t)refs.size();
        bool any_exactable = false, indexable = true, nfree = true, one_len = true;
        size_t first_len = 0;
        for (uint32_t i = 0; i < S.n_ref; i++) {
            const char* r = S.ref_seqs[i];
            const size_t rl = r ? strlen(r) : 0;
            if (rl > BC_MAX_REF_LEN) FAILC(BC_EUNSUPPORTED, "slot %u: reference barcode %u longer than %d bases", s, i, BC_MAX_REF_LEN);
            uint4 v = make_uint4(0, 0, 0, (uint32_t)rl);
            for (size_t p = 0; p < rl; p++) {
                const uint32_t bit = 1u << p;
                switch (r[p]) {
                    case 'A': break;
                    case 'C': v.x |= bit; break;
                    case 'G': v.y |= bit; break;
                    case 'T': v.x |= bit; v.y |= bit; break;
                    case 'N': v.z |= bit; break;
                    default: FAILC(BC_EUNSUPPORTED, "slot %u: reference barcode '%s' holds a character outside ACGTN", s, r);
                }
            }
            if (rl == S.len && v.z == 0) any_exactable = true;
            else indexable = false;
            if (v.z) nfree = false;
            if (i == 0) first_len = rl;
            else if (rl != first_len) one_len = false;
            refs.push_back(v);
        }
        if (S.len <= 10 && S.n_ref < 0xFFFFu) {
            D.mode = MODE_TABLE;
            D.n_inline = nfree && one_len;
            D.aux_off = (uint32_t)table_u16;
            table_u16 += (size_t)1 << (2 * S.len);
        } else if (any_exactable) {
            D.mode = MODE_HASH;
            const unsigned long long cap = pow2_at_least(2ull * S.n_ref);
            D.aux_off = (uint32_t)hkeys.size();
            D.aux_mask = (uint32_t)(cap - 1);
            hkeys.resize(hkeys.size() + cap, 0ull);
            hidx.resize(hidx.size() + cap, kFail);
            for (uint32_t i = 0; i < S.n_ref; i++) {
                const uint4 v = refs[D.ref_off + i];
                if (v.w != S.len || v.z != 0) continue;
                const unsigned long long k = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
                unsigned long long h = mix64(k) & D.aux_mask;
                while (hidx[D.aux_off + h] != kFail && hkeys[D.aux_off + h] != k) h = (h + 1) & D.aux_mask;
                hkeys[D.aux_off + h] = k;
                hidx[D.aux_off + h] = i;  // identical strings cannot repeat in a set; the last one wins like a map insert
            }
            if (indexable) {
                // half index: every reference within distance 1 of a query shares one of its halves
                D.has_half = 1;
                D.half_len0 = (uint16_t)(S.len / 2);
                D.half_off = (uint32_t)half.size();
                D.half_mask = (uint32_t)(cap - 1);
                half.resize(half.size() + 2 * cap, kEmpty);
                for (uint32_t i = 0; i < S.n_ref; i++) {
                    const uint4 v = refs[D.ref_off + i];
                    for (int hh = 0; hh < 2; hh++) {
                        const uint32_t pos0 = hh ? D.half_len0 : 0u, hl = hh ? (uint32_t)S.len - D.half_len0 : (uint32_t)D.half_len0;
                        const uint32_t hm = hl >= 32 ? 0xFFFFFFFFu : ((1u << hl) - 1u);
                        const uint32_t key = ((v.x >> pos0) & hm) | (((v.y >> pos0) & hm) << 16);
                        uint32_t x = key;  // mix32 of bc_kernels.cu
                        x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
                        size_t pos = x & D.half_mask;
                        unsigned long long* tab = half.data() + D.half_off + (hh ? cap : 0);
                        while (tab[pos] != kEmpty) pos = (pos + 1) & D.half_mask;
                        tab[pos] = (unsigned long long)key | ((unsigned long long)i << 32);
                    }
                }
                // block index: one level per distance cap k = 2 (or max_err if smaller) .. max_err, k+1 blocks each,
                // bucketed by (up to) the first kMaxBlockKey bases of each block
                if ((uint32_t)S.max_err + 1u <= (uint32_t)kMaxBlocks && S.len / ((uint32_t)S.max_err + 1u) >= 3 && S.n_ref >= 256) {
                    D.deep_off = (uint32_t)deep.size();
                    for (uint32_t k = std::min<uint32_t>(2, S.max_err); k <= S.max_err; k++) {
                        const uint32_t P = k + 1u;
                        DevDeep dd{};
                        dd.n_blocks = P;
                        dd.cap = k;
                        for (uint32_t p = 0; p < P; p++) {
                            const uint32_t b0 = p * S.len / P, b1 = (p + 1) * S.len / P;
                            const uint32_t kl = std::min<uint32_t>(b1 - b0, kMaxBlockKey);
                            dd.key_pos[p] = (uint8_t)b0;
                            dd.key_len[p] = (uint8_t)kl;
                            const uint32_t nb = 1u << (2 * kl), km = (1u << kl) - 1u;
                            dd.start_off[p] = (uint32_t)csr.size();
                            csr.resize(csr.size() + nb + 1, 0u);
                            dd.ids_off[p] = (uint32_t)bref.size();
                            bref.resize(bref.size() + S.n_ref);
                            uint32_t* start = csr.data() + dd.start_off[p];
                            uint4* ids = bref.data() + dd.ids_off[p];
                            auto bucket = [&](uint32_t i) {
                                const uint4 v = refs[D.ref_off + i];
                                return ((v.x >> b0) & km) | (((v.y >> b0) & km) << kl);
                            };
                            for (uint32_t i = 0; i < S.n_ref; i++) start[bucket(i) + 1]++;
                            for (uint32_t bb = 0; bb < nb; bb++) start[bb + 1] += start[bb];
                            std::vector<uint32_t> fill(start, start + nb);
                            for (uint32_t i = 0; i < S.n_ref; i++) {
                                const uint4 v = refs[D.ref_off + i];
                                ids[fill[bucket(i)]++] = make_uint4(v.x, v.y, i, 0u);
                            }
                        }
                        deep.push_back(dd);
                        D.n_levels++;
                    }
                }
            }
        } else {
            D.mode = MODE_SCAN;
        }
    }
    return BC_OK;
#undef FAILC
}

static void plan_marginals(bc_ctx* ctx);

int bc_create(const bc_config* cfg, int device, uint64_t expected_reads, bc_ctx** out) {
    if (!cfg || !out) return fail(nullptr, BC_EINVAL, "cfg / out is NULL");
    *out = nullptr;
    if (cfg->abi_version != BC_ABI_VERSION) return fail(nullptr, BC_EINVAL, "abi_version %u != %u", cfg->abi_version, BC_ABI_VERSION);
    const uint32_t L = cfg->template_len;
    if (L == 0 || L > BC_MAX_TEMPLATE) return fail(nullptr, BC_EUNSUPPORTED, "template length %u outside 1..%d", L, BC_MAX_TEMPLATE);
    if (cfg->max_read_len < L || cfg->max_read_len > BC_MAX_READ_LEN)
        return fail(nullptr, BC_EUNSUPPORTED, "max_read_len %u outside %u..%d", cfg->max_read_len, L, BC_MAX_READ_LEN);
    if (cfg->n_slots > BC_MAX_SLOTS) return fail(nullptr, BC_EUNSUPPORTED, "more than %d barcodes", BC_MAX_SLOTS);
    if (cfg->region_len > L) return fail(nullptr, BC_EINVAL, "region_codes longer than the template");
    if (!cfg->template_chars || (cfg->region_len && !cfg->region_codes)) return fail(nullptr, BC_EINVAL, "template / region pointers");

    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev == 0)
        return fail(nullptr, BC_ECUDA, "no usable CUDA device (%s); this library has no CPU path", cudaGetErrorString(ce));
    if (device < 0 || device >= n_dev) return fail(nullptr, BC_EINVAL, "device %d of %d", device, n_dev);

    bc_ctx* ctx = new bc_ctx();
    ctx->device = device;
#define CKC(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) {                                                                               \
            int rc_ = fail(nullptr, BC_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            bc_destroy(ctx);                                                                                   \
            return rc_;                                                                                        \
        }                                                                                                      \
    } while (0)
#define FAILC(code, ...)                              \
    do {                                              \
        int rc_ = fail(nullptr, code, __VA_ARGS__);   \
        bc_destroy(ctx);                              \
        return rc_;                                   \
    } while (0)

    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));

    HostRefs H;
    {
        const int rc_ = build_run_constants(cfg, ctx, H);
        if (rc_ != BC_OK) {
            bc_destroy(ctx);
            return rc_;
        }
    }
    DevCfg& d = ctx->cfg;
    std::vector<uint4>& refs = H.refs;
    std::vector<unsigned long long>& hkeys = H.hkeys;
    std::vector<uint32_t>& hidx = H.hidx;
    std::vector<unsigned long long>& half = H.half;
    std::vector<DevDeep>& deep = H.deep;
    std::vector<uint32_t>& csr = H.csr;
    std::vector<uint4>& bref = H.bref;
    const size_t table_u16 = H.table_u16;
    if (!refs.empty()) {
        CKC(cudaMalloc(&ctx->d_refs, refs.size() * sizeof(uint4)));
        CKC(cudaMemcpyAsync(ctx->d_refs, refs.data(), refs.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!hkeys.empty()) {
        CKC(cudaMalloc(&ctx->d_hash_keys, hkeys.size() * sizeof(unsigned long long)));
        CKC(cudaMalloc(&ctx->d_hash_idx, hidx.size() * sizeof(uint32_t)));
        CKC(cudaMemcpyAsync(ctx->d_hash_keys, hkeys.data(), hkeys.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
        CKC(cudaMemcpyAsync(ctx->d_hash_idx, hidx.data(), hidx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!half.empty()) {
        CKC(cudaMalloc(&ctx->d_half, half.size() * sizeof(unsigned long long)));
        CKC(cudaMemcpyAsync(ctx->d_half, half.data(), half.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!deep.empty()) {
        CKC(cudaMalloc(&ctx->d_deep, deep.size() * sizeof(DevDeep)));
        CKC(cudaMemcpyAsync(ctx->d_deep, deep.data(), deep.size() * sizeof(DevDeep), cudaMemcpyHostToDevice, ctx->stream));
        CKC(cudaMalloc(&ctx->d_csr, csr.size() * sizeof(uint32_t)));
        CKC(cudaMemcpyAsync(ctx->d_csr, csr.data(), csr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CKC(cudaMalloc(&ctx->d_bref, bref.size() * sizeof(uint4)));
        CKC(cudaMemcpyAsync(ctx->d_bref, bref.data(), bref.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (table_u16) CKC(cudaMalloc(&ctx->d_tables, table_u16 * sizeof(uint32_t)));
    CKC(cudaMalloc(&ctx->d_def_count, 2 * sizeof(uint32_t)));
    CKC(cudaMemsetAsync(ctx->d_def_count, 0, 2 * sizeof(uint32_t), ctx->stream));
    ctx->aux = DevAux{ctx->d_refs, ctx->d_tables, ctx->d_hash_keys, ctx->d_hash_idx, ctx->d_half, ctx->d_deep, ctx->d_csr, ctx->d_bref};
    for (uint32_t s = 0; s < cfg->n_slots; s++) {
        if (d.slots[s].mode != MODE_TABLE) continue;
        ctx->prof.launches[BC_K_OTHER]++;
        CKC(launch_build_table(d.slots[s], ctx->aux, ctx->d_tables + d.slots[s].aux_off, ctx->stream));
    }
    CKC(cudaStreamSynchronize(ctx->stream));  // host vectors above go out of scope

    // ---- counters and the tables
    CKC(cudaMalloc(&ctx->d_counters, (BC_N_COUNTERS + 2) * sizeof(unsigned long long)));
    CKC(cudaMemsetAsync(ctx->d_counters, 0, (BC_N_COUNTERS + 2) * sizeof(unsigned long long), ctx->stream));
    CKC(cudaMalloc(&ctx->d_stripes, (size_t)kCounterStripes * kCounterStride * sizeof(unsigned long long)));
    CKC(cudaMemsetAsync(ctx->d_stripes, 0, (size_t)kCounterStripes * kCounterStride * sizeof(unsigned long long), ctx->stream));
    CKC(cudaMalloc(&ctx->d_row_n, sizeof(unsigned long long)));
    const unsigned long long hint = expected_reads ? expected_reads : (1ull << 20);
    int rc;
    Tables& T = ctx->tables;
    T.umi_bits = d.umi_bits;
    T.has_set = d.has_umi ? 1 : 0;
    const uint32_t map_bits = d.key_bits - d.umi_bits;
    // Dense counters (index = key: one fire-and-forget RED per update, no key storage, no probe) whenever the key space
    // is small — or not much larger than the job itself and affordable in HBM (DEL: 3 x 10 bits = 2^30 counters = 8.6 GB)
    bool dense = map_bits <= 27;
    // Deferred counting whenever read-by-read updates would be random DRAM traffic: a (key, UMI) set, or a hashed map.
    // A small dense count array without a random barcode lives in L2 and keeps the inline RED (CRISPR screens).
    // BC_CFG_INLINE_COUNT keeps the read-by-read tables (measurement aid; the same tables are the flush's fallback).
    ctx->deferred = (T.has_set || !dense) && !(cfg->flags & BC_CFG_INLINE_COUNT);
    ctx->expected_reads = hint;
    CKC(cudaMalloc(&ctx->d_flush, sizeof(FlushStats)));
    CKC(cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
    if (ctx->deferred) {
        // tables exist only as descriptors (kind / width) until a flush needs the global path
        T.map = DevTable{};
        T.map.kind = 1;
        T.map.wide = map_bits > 63;
        T.map.n_entries = ctx->d_counters + BC_N_COUNTERS;
        T.set = DevTable{};
        T.set.kind = 2;
        T.set.wide = d.wide;
        T.set.n_entries = ctx->d_counters + BC_N_COUNTERS + 1;
        rc = BC_OK;
    } else {
        if (!dense && map_bits <= 31 && (1ull << map_bits) <= 4 * hint) {
            size_t free_b = 0, total_b = 0;
            dense = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (sizeof(unsigned long long) << map_bits) <= free_b / 4;
        }
        if (dense) rc = alloc_table(ctx, T.map, 0, 0, 1ull << map_bits, ctx->d_counters + BC_N_COUNTERS);
        else rc = alloc_table(ctx, T.map, 1, map_bits > 63, slots_for(hint), ctx->d_counters + BC_N_COUNTERS);
        if (rc == BC_OK && T.has_set) rc = alloc_table(ctx, T.set, 2, d.wide, slots_for(hint), ctx->d_counters + BC_N_COUNTERS + 1);
    }
    if (rc != BC_OK) {
        g_create_error = ctx->err;
        bc_destroy(ctx);
        return rc;
    }
    plan_marginals(ctx);
    // Specialise the decode kernel for this run when the job is large enough to repay a second or two of compilation
    // (or when asked to): same device code, run constants folded in.
    if (!(cfg->flags & BC_CFG_NO_SPECIALIZE) && ((cfg->flags & BC_CFG_SPECIALIZE) || expected_reads >= (1ull << 22))) {
        ctx->jit = jit_decode(d, ctx->W, ctx->plane_stride, ctx->qual_stride, &ctx->jit_note);
        if (!ctx->jit.kernel && (cfg->flags & BC_CFG_SPECIALIZE)) FAILC(BC_ECUDA, "BC_CFG_SPECIALIZE: %s", ctx->jit_note.c_str());
    } else {
        ctx->jit_note = "not requested (small job)";
    }
    CKC(cudaStreamSynchronize(ctx->stream));
    *out = ctx;
    return BC_OK;
#undef CKC
#undef FAILC
}

const char* bc_specialization_note(const bc_ctx* ctx) { return ctx ? (ctx->jit.kernel ? "specialised" : ctx->jit_note.c_str()) : ""; }

int bc_jit_check(const bc_config* cfg, char* log, int loglen) {
    if (!cfg || !log || loglen <= 0) return BC_EINVAL;
    log[0] = 0;
    if (cfg->abi_version != BC_ABI_VERSION || cfg->template_len == 0 || cfg->template_len > BC_MAX_TEMPLATE || cfg->n_slots > BC_MAX_SLOTS ||
        cfg->max_read_len < cfg->template_len || cfg->max_read_len > BC_MAX_READ_LEN || !cfg->template_chars)
        return BC_EINVAL;
    bc_ctx* ctx = new bc_ctx();
    HostRefs H;
    int rc = build_run_constants(cfg, ctx, H);
    std::string msg;
    size_t bytes = 0;
    if (rc == BC_OK) {
        bytes = jit_compile_check(ctx->cfg, bc_plane_words(cfg->max_read_len), bc_plane_stride(cfg->max_read_len), bc_qual_stride(cfg->max_read_len), &msg);
        if (bytes == 0) rc = BC_ECUDA;
    } else {
        msg = g_create_error;
    }
    snprintf(log, (size_t)loglen, "%s", rc == BC_OK ? ("cubin bytes: " + std::to_string(bytes)).c_str() : msg.c_str());
    delete ctx;
    return rc;
}

int bc_set_stream(bc_ctx* ctx, void* cuda_stream) {
    if (!ctx) return BC_EINVAL;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    return BC_OK;
}

static SplitLevel owner_level(const bc_ctx* ctx);

// receive allocation of a rank (local, or a peer's mapping): [buffer 0][buffer 1][cursor 0, cursor 1], a buffer being
// xcap keys (lo) followed, for wide keys, by xcap high words
static unsigned long long* xbuf_lo(const bc_ctx* ctx, unsigned long long* base, uint32_t parity) {
    return base + (unsigned long long)parity * ctx->xcap * (ctx->cfg.wide ? 2 : 1);
}
static unsigned long long* xbuf_cursor(const bc_ctx* ctx, unsigned long long* base, uint32_t parity) {
    return base + 2 * ctx->xcap * (ctx->cfg.wide ? 2 : 1) + parity;
}


// Streamed exchange: records [first, first + n) of the record buffer — the batch the main stream has just decoded — go to their
// owners' receive buffers on x_stream, under the decode of the next batch.  The mode of a job is fixed by its first batch:
// streamed when every peer is connected by then, else the bulk exchange after the last batch (bc_exchange_count / _scatter).
static int stream_records(bc_ctx* ctx, unsigned long long first, unsigned long long n) {
    if (ctx->x_dirty)
        return fail(ctx, BC_ESTATE, "a multi-GPU job was reset after some of its records had been streamed to their owners: re-open the "
                                    "exchange (bc_exchange_open + connect) on every rank before the next job");
    if (ctx->x_mode == 0) {
        bool connected = !ctx->opt_exchange_bulk;
        for (uint32_t r = 0; r < ctx->x_ranks; r++) connected = connected && ctx->x_peer[r] != nullptr;
        ctx->x_mode = connected ? 1 : 2;
        if (connected) {
            CK(ctx, cudaMemsetAsync(ctx->d_xsent, 0, kMaxRanks * sizeof(unsigned long long), ctx->stream));
            CK(ctx, cudaMemsetAsync(ctx->d_xoverflow, 0, sizeof(unsigned int), ctx->stream));
            ctx->x_overflowed = false;
        }
    }
    if (ctx->x_mode != 1 || n == 0) return BC_OK;
    const bool wide = ctx->cfg.wide != 0;
    const uint32_t parity = ctx->x_epoch & 1u;
    PeerOut peers{};
    for (uint32_t r = 0; r < ctx->x_ranks; r++) {
        peers.lo[r] = xbuf_lo(ctx, ctx->x_peer[r], parity);
        peers.hi[r] = wide ? peers.lo[r] + ctx->xcap : nullptr;
        peers.cursor[r] = xbuf_cursor(ctx, ctx->x_peer[r], parity);
    }
    peers.cap = ctx->xcap;
    peers.sent = ctx->d_xsent;
    peers.overflow = ctx->d_xoverflow;
    CK(ctx, cudaEventRecord(ctx->x_ev, ctx->stream));
    CK(ctx, cudaStreamWaitEvent(ctx->x_stream, ctx->x_ev, 0));
    ProfScope p(ctx, BC_K_EXCHANGE, ctx->x_stream);
    CK(ctx, launch_owner_scatter(wide, ItemView{ctx->rec.lo + first, wide ? ctx->rec.hi + first : nullptr, nullptr}, peers, n, owner_level(ctx),
                                 ctx->d_xcursor, ctx->x_stream));
    return BC_OK;
}

static int run_decode(bc_ctx* ctx, const bc_batch* batch, int flags, const DecodeOut& out, unsigned long long* counters) {
    int rc = validate_batch(ctx, batch);
    if (rc != BC_OK) return rc;
    if (batch->n_reads == 0) return BC_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    if (flags & F_INSERT) {
        rc = ensure_capacity(ctx, batch->n_reads);
        if (rc != BC_OK) return rc;
    }
    if (flags & F_APPEND) {
        rc = prime_records(ctx, batch->n_reads);
        if (rc == BC_OK) rc = reserve_records(ctx, batch->n_reads);
        if (rc != BC_OK) return rc;
    }
    if (flags & (F_INSERT | F_APPEND)) {
        ctx->rows_valid = false;
        ctx->marg_valid = false;
        ctx->x_state = 0;
    }
    BatchView view{};
    const int staged = stage_batch(ctx, batch, &view);
    if (staged < 0) return staged;
    if (ctx->def_cap < batch->n_reads) {  // worst case every read of the batch is deferred
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->d_def_items) cudaFree(ctx->d_def_items);
        ctx->d_def_items = nullptr;
        CK(ctx, cudaMalloc(&ctx->d_def_items, (size_t)batch->n_reads * sizeof(uint2)));
        ctx->def_cap = batch->n_reads;
    }
    // the two counters of the deferred list alternate: this batch's k_resolve zeroes the one the next batch will fill
    const Deferred deferred{ctx->d_def_items, ctx->d_def_count + ctx->def_parity, ctx->d_def_count + (ctx->def_parity ^ 1u)};
    ctx->def_parity ^= 1u;
    {
        ProfScope p(ctx, BC_K_DECODE);
        const bool fits = ctx->jit.kernel && view.W == ctx->jit.W && view.plane_stride == ctx->jit.plane_stride &&
                          (!ctx->quality_on || view.qual_stride == ctx->jit.qual_stride) && decode_smem_bytes(view) <= 48 * 1024;
        if (fits) {
            ctx->jit_launches++;
            CK(ctx, launch_decode_jit(ctx->jit.kernel, view, ctx->aux, ctx->tables, counters ? ctx->d_stripes : nullptr, out, rec_out(ctx),
                                      deferred, flags, ctx->stream));
        } else {
            ctx->generic_launches++;
            CK(ctx, launch_decode(ctx->cfg, view, ctx->aux, ctx->tables, counters ? ctx->d_stripes : nullptr, out, rec_out(ctx), deferred,
                                  flags, ctx->stream));
        }
    }
    if (counters) ctx->stripes_dirty = true;
    {   // also after a locate-only launch (nothing deferred): it is k_resolve that zeroes the next batch's list counter
        ProfScope p(ctx, BC_K_SCAN);
        CK(ctx, launch_resolve(ctx->cfg, view, ctx->aux, ctx->tables, counters, out, rec_out(ctx), deferred, flags, ctx->stream));
    }
    if ((flags & F_APPEND) && ctx->x_ranks > 1) {
        rc = stream_records(ctx, ctx->rec_n, batch->n_reads);
        if (rc != BC_OK) return rc;
    }
    if (flags & F_APPEND) ctx->rec_n += batch->n_reads;
    return release_staging(ctx, staged);
}

int bc_submit(bc_ctx* ctx, const bc_batch* batch) {
    if (!ctx) return BC_EINVAL;
    return run_decode(ctx, batch, ctx->deferred ? F_APPEND : F_INSERT, DecodeOut{}, ctx->d_counters);
}

uint32_t bc_wire_qual_codes(uint32_t max_read_len) { return (max_read_len + 3u) & ~3u; }
uint32_t bc_wire_qual_stride(uint32_t max_read_len, uint32_t bits) {
    return (uint32_t)(((unsigned long long)bc_wire_qual_codes(max_read_len) * bits + 31) / 32 * 4);
}

// A host batch in its transfer form: copied to the device as it is (fewer bytes across PCIe), expanded there into the
// bc_batch layout, then decoded like any device batch.
int bc_submit_wire(bc_ctx* ctx, const bc_wire_batch* wb) {
    if (!ctx) return BC_EINVAL;
    if (!wb) return fail(ctx, BC_EINVAL, "wire batch is NULL");
    if (wb->n_reads == 0) return BC_OK;
    const uint32_t mrl = wb->max_read_len, W = bc_plane_words(mrl);
    if (mrl == 0 || mrl > BC_MAX_READ_LEN) return fail(ctx, BC_EINVAL, "wire batch: max_read_len %u", mrl);
    if (!wb->lohi || !wb->read_len) return fail(ctx, BC_EINVAL, "wire batch: lohi / read_len is NULL");
    if (!wb->nmask && wb->n_calls && (!wb->n_read || !wb->n_pos)) return fail(ctx, BC_EINVAL, "wire batch: n_calls > 0 but no list");
    const bool want_qual = ctx->quality_on;
    const uint32_t bits = wb->qual_bits;
    if (want_qual) {
        if (!wb->qual) return fail(ctx, BC_EINVAL, "min_quality > 0 but the wire batch has no quality codes");
        if (bits != 8 && bits != 6 && bits != 4 && bits != 2) return fail(ctx, BC_EINVAL, "wire batch: qual_bits %u (8, 6, 4 or 2)", bits);
        if (wb->qual_stride != bc_wire_qual_stride(mrl, bits))
            return fail(ctx, BC_EINVAL, "wire batch: qual_stride %u is not bc_wire_qual_stride(%u, %u)", wb->qual_stride, mrl, bits);
    }
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t n = wb->n_reads;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const bool dense_n = wb->nmask != nullptr;
    const size_t b_lohi = n * 2 * W * 4, b_len = n * 2, b_nm = dense_n ? n * W * 4 : 0, b_nr = dense_n ? 0 : (size_t)wb->n_calls * 4,
                 b_np = dense_n ? 0 : (size_t)wb->n_calls * 2, b_q = want_qual ? n * wb->qual_stride : 0;
    const size_t o_lohi = 0, o_len = o_lohi + up(b_lohi), o_nm = o_len + up(b_len), o_nr = o_nm + up(b_nm), o_np = o_nr + up(b_nr),
                 o_q = o_np + up(b_np), need = o_q + up(b_q);
    bc_ctx::WireStage& s = ctx->wire[ctx->wire_cur];
    ctx->wire_cur ^= 1;
    if (!s.free_ev) {
        CK(ctx, cudaEventCreateWithFlags(&s.free_ev, cudaEventDisableTiming));
        CK(ctx, cudaEventCreateWithFlags(&s.copied_ev, cudaEventDisableTiming));
    }
    if (s.cap < need) {
        CK(ctx, cudaEventSynchronize(s.free_ev));
        if (s.buf) cudaFree(s.buf);
        s.buf = nullptr;
        s.cap = need + need / 8;
        CK(ctx, cudaMalloc(&s.buf, s.cap));
    }
    // the expanded arrays: the batch's own geometry (like a device bc_batch)
    const uint32_t ps = bc_plane_stride(mrl), qs = bc_qual_stride(mrl);
    const size_t xb_p = n * ps * 4, xb_l = n * 2, xb_q = want_qual ? n * qs : 0;
    if (ctx->x_cap_planes < xb_p || ctx->x_cap_len < xb_l || ctx->x_cap_qual < xb_q) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));  // an earlier batch's kernels may still read them
        if (ctx->x_planes) cudaFree(ctx->x_planes);
        if (ctx->x_len) cudaFree(ctx->x_len);
        if (ctx->x_qual) cudaFree(ctx->x_qual);
        ctx->x_planes = nullptr; ctx->x_len = nullptr; ctx->x_qual = nullptr;
        ctx->x_cap_planes = std::max(ctx->x_cap_planes, xb_p + xb_p / 8);
        ctx->x_cap_len = std::max(ctx->x_cap_len, xb_l + xb_l / 8);
        ctx->x_cap_qual = std::max(ctx->x_cap_qual, xb_q + xb_q / 8);
        CK(ctx, cudaMalloc(&ctx->x_planes, ctx->x_cap_planes));
        CK(ctx, cudaMalloc(&ctx->x_len, ctx->x_cap_len));
        if (want_qual) CK(ctx, cudaMalloc(&ctx->x_qual, ctx->x_cap_qual));
    }
    CK(ctx, cudaStreamWaitEvent(ctx->copy_stream, s.free_ev, 0));
    auto h2d = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
        return bytes ? cudaMemcpyAsync(s.buf + off, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream) : cudaSuccess;
    };
    CK(ctx, h2d(o_lohi, wb->lohi, b_lohi));
    CK(ctx, h2d(o_len, wb->read_len, b_len));
    CK(ctx, h2d(o_nm, wb->nmask, b_nm));
    CK(ctx, h2d(o_nr, wb->n_read, b_nr));
    CK(ctx, h2d(o_np, wb->n_pos, b_np));
    CK(ctx, h2d(o_q, wb->qual, b_q));
    CK(ctx, cudaEventRecord(s.copied_ev, ctx->copy_stream));
    CK(ctx, cudaStreamWaitEvent(ctx->stream, s.copied_ev, 0));
    ctx->prof.h2d_bytes += b_lohi + b_len + b_nm + b_nr + b_np + b_q;
    ctx->copies_pending = true;
    WireView v{};
    v.lohi = reinterpret_cast<const uint32_t*>(s.buf + o_lohi);
    v.read_len = reinterpret_cast<const uint16_t*>(s.buf + o_len);
    v.nmask = dense_n ? reinterpret_cast<const uint32_t*>(s.buf + o_nm) : nullptr;
    v.n_read = reinterpret_cast<const uint32_t*>(s.buf + o_nr);
    v.n_pos = reinterpret_cast<const uint16_t*>(s.buf + o_np);
    v.qual = s.buf + o_q;
    v.n_reads = wb->n_reads;
    v.W = W;
    v.n_calls = dense_n ? 0u : wb->n_calls;
    v.qual_bits = bits;
    v.qual_stride = wb->qual_stride;
    v.n_codes = bc_wire_qual_codes(mrl);
    memcpy(v.dict.c, wb->qual_dict, sizeof v.dict.c);
    {
        ProfScope p(ctx, BC_K_OTHER);
        CK(ctx, launch_wire_expand(v, ctx->x_planes, ctx->x_len, want_qual ? ctx->x_qual : nullptr, ps, qs, ctx->stream));
    }
    CK(ctx, cudaEventRecord(s.free_ev, ctx->stream));  // the arena is free once it is expanded
    bc_batch b{};
    b.n_reads = wb->n_reads;
    b.plane_stride = ps;
    b.qual_stride = qs;
    b.location = BC_LOC_DEVICE;
    b.planes = ctx->x_planes;
    b.read_len = ctx->x_len;
    b.qual = want_qual ? ctx->x_qual : nullptr;
    return run_decode(ctx, &b, ctx->deferred ? F_APPEND : F_INSERT, DecodeOut{}, ctx->d_counters);
}

int bc_set_option(bc_ctx* ctx, const char* name, int value) {
    if (!ctx || !name) return BC_EINVAL;
    if (!strcmp(name, "flush_global")) ctx->opt_flush_global = value != 0;
    else if (!strcmp(name, "flush_two_stage")) ctx->opt_flush_two_stage = value != 0;
    else if (!strcmp(name, "exchange_bulk")) ctx->opt_exchange_bulk = value != 0;
    else return fail(ctx, BC_EINVAL, "bc_set_option: unknown option '%s'", name);
    ctx->rows_valid = false;
    return BC_OK;
}

int bc_sync(bc_ctx* ctx) {
    if (!ctx) return BC_EINVAL;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->copy_stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->x_stream) CK(ctx, cudaStreamSynchronize(ctx->x_stream));
    ctx->copies_pending = false;
    return BC_OK;
}

int bc_wait_copies(bc_ctx* ctx) {
    if (!ctx) return BC_EINVAL;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->copy_stream));
    ctx->copies_pending = false;
    return BC_OK;
}

// Double buffering on the caller's side: before it rewrites the batch it handed over two submits ago it only has to wait
// for THAT copy, not for the one just queued.  The staging slots alternate, so the slot the next submit will use is the one
// whose copy must be over.
int bc_wait_older_copies(bc_ctx* ctx) {
    if (!ctx) return BC_EINVAL;
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->staging[ctx->cur].copied_ev) CK(ctx, cudaEventSynchronize(ctx->staging[ctx->cur].copied_ev));
    if (ctx->wire[ctx->wire_cur].copied_ev) CK(ctx, cudaEventSynchronize(ctx->wire[ctx->wire_cur].copied_ev));
    return BC_OK;
}

// ---------------------------------------------------------------------------------------------- deferred counting
// The whole record buffer -> final rows (ctx row buffers) and the matched / duplicates split.  A pure function of the
// buffer (plus the imported rows), so it may run any number of times as more batches arrive.
struct FlushSrc {  // what a flush reads: the record buffer, or (multi-GPU) what the exchange delivered
    ItemView v;
    unsigned long long n, n_valid;
};
static int flush_global(bc_ctx* ctx, const FlushSrc& in);

static int apply_duplicates(bc_ctx* ctx, unsigned long long dup_now) {
    // k_decode counted every appended record as "matched"; move the repeats to "duplicates" (parse.rs:65-69)
    if (dup_now == ctx->dup_applied) return BC_OK;
    unsigned long long h[BC_N_COUNTERS];
    CK(ctx, cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    const long long delta = (long long)dup_now - (long long)ctx->dup_applied;
    h[BC_CNT_MATCHED] -= delta;
    h[BC_CNT_DUPLICATES] += delta;
    CK(ctx, cudaMemcpy(ctx->d_counters, h, sizeof h, cudaMemcpyHostToDevice));
    ctx->dup_applied = dup_now;
    return BC_OK;
}

// Hash-partitions n items (about n_valid of them not holes) into *n_parts pieces of ~reduce_fill items each
// (ctx->d_starts = their offsets in `out`).  One radix level for up to 2048 partitions, otherwise two (through `tmp`).
// drop_bits > 0 partitions by the key without its random barcode; max_part > 0 asks for every partition to hold at
// most that many items.  Returns 1 when the input is too large for two levels (the caller then takes the global-table
// path); with max_part: 2 when so many items sit in partitions larger than max_part that partitioning by key is
// pointless (nothing was moved to `out`), 3 when a few partitions are larger (moved like the others; *big_items says
// how many items they hold).
static int partition_items(bc_ctx* ctx, const ItemView& in, bool wide, unsigned long long n, unsigned long long n_valid, bool weighted,
                           const ItemView& out, bool count_valid, uint32_t drop_bits, uint32_t max_part, unsigned long long* n_parts,
                           unsigned long long* big_items = nullptr, unsigned long long salt = 0) {
    unsigned long long P = (n_valid + reduce_fill(wide) - 1) / reduce_fill(wide);
    if (P < 1) P = 1;
    const uint32_t mb = split_max_bits();
    uint32_t l2 = 0;
    unsigned long long F1 = P;
    if (P > (1ull << mb)) {
        uint32_t lg = 0;
        while ((1ull << lg) < P) lg++;
        l2 = (lg + 1) / 2;
        F1 = (P + (1ull << l2) - 1) >> l2;
        P = F1 << l2;
        if (l2 > mb || F1 > (1ull << mb)) return 1;
    }
    *n_parts = P;
    int rc = reserve_parts(ctx, P);
    if (rc != BC_OK) return rc;
    ProfScope p(ctx, BC_K_FINISH);
    int verdict = BC_OK;
    auto too_big = [&](bool* yes) -> int {  // partitions of the histogram just scanned that exceed max_part
        FlushStats st{};
        CK(ctx, cudaMemcpyAsync(&st, ctx->d_flush, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        *yes = st.big_items * 5 > n_valid;  // more than a fifth of the items under hot keys: not worth it
        if (st.max_bin > max_part) verdict = 3;
        if (big_items) *big_items = st.big_items;
        return BC_OK;
    };
    CK(ctx, cudaMemsetAsync(ctx->d_hist, 0, (P + 1) * sizeof(uint32_t), ctx->stream));
    if (l2 == 0) {
        const SplitLevel lv{P, 0u, 0xFFFFFFFFu, (uint32_t)P, drop_bits, salt};
        CK(ctx, launch_split(false, wide, in, out, nullptr, 1, n, lv, ctx->d_hist, ctx->d_flush, count_valid, ctx->stream));
        CK(ctx, launch_seg_scan(ctx->d_hist, 1, (uint32_t)P, nullptr, ctx->d_starts, ctx->d_cursor, max_part ? ctx->d_flush : nullptr, max_part, ctx->stream));
        if (max_part) {
            bool big = false;
            rc = too_big(&big);
            if (rc != BC_OK) return rc;
            if (big) return 2;
        }
        CK(ctx, launch_split(true, wide, in, out, nullptr, 1, n, lv, ctx->d_cursor, ctx->d_flush, false, ctx->stream));
        return verdict;
    }
    rc = reserve_items(ctx, ctx->tmp, n, wide, weighted);
    if (rc != BC_OK) return rc;
    if (!ctx->d_l1) CK(ctx, cudaMalloc(&ctx->d_l1, 3 * ((1u << mb) + 1) * sizeof(uint32_t)));
    uint32_t *hist1 = ctx->d_l1, *starts1 = hist1 + (1u << mb) + 1, *cursor1 = starts1 + (1u << mb) + 1;
    const ItemView tmp{ctx->tmp.lo, wide ? ctx->tmp.hi : nullptr, weighted ? ctx->tmp.w : nullptr};
    const SplitLevel lv1{P, l2, 0xFFFFFFFFu, (uint32_t)F1, drop_bits, salt}, lv2{P, 0u, (1u << l2) - 1u, 1u << l2, drop_bits, salt};
    CK(ctx, cudaMemsetAsync(hist1, 0, (F1 + 1) * sizeof(uint32_t), ctx->stream));
    CK(ctx, launch_split(false, wide, in, tmp, nullptr, 1, n, lv1, hist1, ctx->d_flush, count_valid, ctx->stream));
    CK(ctx, launch_seg_scan(hist1, 1, (uint32_t)F1, nullptr, starts1, cursor1, max_part ? ctx->d_flush : nullptr, 0xFFFFFFFFu, ctx->stream));
    if (max_part) {  // a first-level segment 2^l2 times the limit cannot split into small enough partitions
        FlushStats st{};
        CK(ctx, cudaMemcpyAsync(&st, ctx->d_flush, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        if (st.max_bin > ((unsigned long long)max_part << l2)) return 2;  // a segment whose AVERAGE partition is too large
        CK(ctx, cudaMemsetAsync(&ctx->d_flush->max_bin, 0, sizeof(unsigned long long), ctx->stream));
    }
    CK(ctx, launch_split(true, wide, in, tmp, nullptr, 1, n, lv1, cursor1, ctx->d_flush, false, ctx->stream));
    CK(ctx, launch_split(false, wide, tmp, out, starts1, (uint32_t)F1, n, lv2, ctx->d_hist, ctx->d_flush, false, ctx->stream));
    CK(ctx, launch_seg_scan(ctx->d_hist, (uint32_t)F1, 1u << l2, starts1, ctx->d_starts, ctx->d_cursor, max_part ? ctx->d_flush : nullptr, max_part, ctx->stream));
    if (max_part) {
        bool big = false;
        rc = too_big(&big);
        if (rc != BC_OK) return rc;
        if (big) return 2;
    }
    CK(ctx, launch_split(true, wide, tmp, out, starts1, (uint32_t)F1, n, lv2, ctx->d_cursor, ctx->d_flush, false, ctx->stream));
    return verdict;
}

static int go_global(bc_ctx* ctx, const FlushSrc& in, int where, unsigned long long code) {
    (void)where;
    (void)code;
    return flush_global(ctx, in);
}

// Items of `in` -> final rows (ctx row buffers); ctx->last_valid / last_unique = records that were not holes / distinct
// (key, random barcode) pairs among them.
static int flush_core(bc_ctx* ctx, const FlushSrc& in) {
    drop_rows(ctx);
    ctx->flushed_global = false;
    ctx->flush_stages = 0;
    ctx->last_valid = ctx->last_unique = 0;
    const unsigned long long n_rec = in.n, n_valid = in.n_valid;
    const bool has_umi = ctx->cfg.has_umi != 0;
    const bool wide_in = ctx->cfg.wide != 0, wide_out = ctx->tables.map.wide != 0;
    FlushStats st{};
    if (n_rec + ctx->imp_n == 0) {
        ctx->rows_valid = true;
        return BC_OK;
    }
    if (ctx->opt_flush_global || n_rec + ctx->imp_n >= 0xFFFFFFF0ULL) return go_global(ctx, in, 1, 0);

    int rc;
    CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
    ItemView src = in.v;
    unsigned long long n_src = n_rec, src_valid = n_valid, rows_done = 0, valid_total = 0, unique_total = 0;
    bool count_valid = has_umi;
    uint32_t stages = 0;
    unsigned long long salt = 0;

    // ---- one stage when (almost) no key is hot: partition by the key WITHOUT its random barcode, so a partition holds
    // every pair of its keys and the reduce kernel's (key, pairs) output is final.  A key with more pairs than a CTA's
    // key store makes its partition too large, which shows in the histogram before anything is moved: a few such
    // partitions are set aside (k_gather_big) and go through the two stages below on their own; if they hold more
    // than a fifth of the records, the one-stage attempt is abandoned.
    if (has_umi && n_rec && ctx->imp_n == 0 && !ctx->opt_flush_two_stage) {
        unsigned long long n_parts = 0, big = 0;
        rc = reserve_items(ctx, ctx->part, n_rec, wide_in, false);
        if (rc == BC_OK) rc = reserve_rows(ctx, n_valid, wide_out);
        if (rc != BC_OK) return rc;
        const ItemView part_v{ctx->part.lo, wide_in ? ctx->part.hi : nullptr, nullptr};
        const uint32_t cap = reduce_capacity(wide_in);
        rc = partition_items(ctx, src, wide_in, n_rec, n_valid, false, part_v, true, ctx->cfg.umi_bits, cap, &n_parts, &big);
        if (rc == 1) return go_global(ctx, in, 2, 0);
        if (rc == BC_OK || rc == 3) {
            {
                // one pass for partitions whose keys have few random barcodes each; the ones it gives up on (listed on the
                // device) go through the two-pass kernel right after — their output is final as well, the partition being by key
                const ItemView rows_out{ctx->d_row_lo, wide_out ? ctx->d_row_hi : nullptr, ctx->d_row_cnt};
                ProfScope p(ctx, BC_K_FINISH);
                CK(ctx, cudaMemsetAsync(ctx->d_hot_n, 0, sizeof(uint32_t), ctx->stream));
                CK(ctx, launch_reduce(RED_DEDUPE_KEYED, wide_in, part_v, ctx->d_starts, n_rec, n_parts, 0, ctx->cfg.umi_bits, rows_out,
                                      ctx->row_cap, ctx->d_flush, rc == 3 ? cap : 0u, HotList{ctx->d_hot, ctx->d_hot_n, false}, ctx->stream));
                CK(ctx, launch_reduce(RED_DEDUPE, wide_in, part_v, ctx->d_starts, n_rec, n_parts, 0, ctx->cfg.umi_bits, rows_out,
                                      ctx->row_cap, ctx->d_flush, rc == 3 ? cap : 0u, HotList{ctx->d_hot, ctx->d_hot_n, true}, ctx->stream));
            }
            CK(ctx, cudaMemcpyAsync(&st, ctx->d_flush, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
            CK(ctx, cudaStreamSynchronize(ctx->stream));
            CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
            if (st.overflow) return go_global(ctx, in, 3, st.overflow);
            rows_done = st.n_out;
            valid_total = st.valid;
            unique_total = st.unique;
            stages = 1;
            if (rc == BC_OK) {
                ctx->n_rows = rows_done;
                ctx->rows_valid = true;
                ctx->flush_stages = 1;
                ctx->last_valid = valid_total;
                ctx->last_unique = unique_total;
                return BC_OK;
            }
            // the hot partitions, gathered: the input of the two stages
            int rc2 = reserve_items(ctx, ctx->left, big, wide_in, false);
            if (rc2 != BC_OK) return rc2;
            const ItemView left_v{ctx->left.lo, wide_in ? ctx->left.hi : nullptr, nullptr};
            {
                ProfScope p(ctx, BC_K_FINISH);
                CK(ctx, launch_gather_big(part_v, ctx->d_starts, n_parts, cap, left_v, ctx->d_flush, ctx->d_big, ctx->d_hot_n, ctx->stream));
            }
            CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
            src = left_v;
            n_src = src_valid = big;
            count_valid = false;
            salt = 0x6A09E667F3BCC909ULL;  // these records share the hash bits that chose their partitions: use another hash
        } else if (rc == 2) {
            CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));  // hot keys everywhere: two stages
        } else {
            return rc;
        }
    }

    // ---- stage A: records -> (key, pairs) items in w1
    rc = reserve_items(ctx, ctx->w1, src_valid + ctx->imp_n, wide_out, true);
    if (rc != BC_OK) return rc;
    const ItemView w1_v{ctx->w1.lo, wide_out ? ctx->w1.hi : nullptr, ctx->w1.w};
    if (n_src) {
        if (has_umi) {
            unsigned long long n_parts = 0;
            rc = reserve_items(ctx, ctx->part, n_src, wide_in, false);
            if (rc != BC_OK) return rc;
            const ItemView part_v{ctx->part.lo, wide_in ? ctx->part.hi : nullptr, nullptr};
            rc = partition_items(ctx, src, wide_in, n_src, src_valid, false, part_v, count_valid, 0, 0, &n_parts, nullptr, salt);
            if (rc == 1) return go_global(ctx, in, 4, 0);
            if (rc != BC_OK) return rc;
            ProfScope p(ctx, BC_K_FINISH);
            CK(ctx, launch_reduce(RED_DEDUPE, wide_in, part_v, ctx->d_starts, n_src, n_parts, 0, ctx->cfg.umi_bits, w1_v, ctx->w1.cap,
                                  ctx->d_flush, 0, HotList{}, ctx->stream));
        } else {
            const uint32_t chunk = reduce_chunk(wide_in);
            ProfScope p(ctx, BC_K_FINISH);
            CK(ctx, launch_reduce(RED_COUNT, wide_in, src, nullptr, n_src, (n_src + chunk - 1) / chunk, chunk, 0, w1_v, ctx->w1.cap,
                                  ctx->d_flush, 0, HotList{}, ctx->stream));
        }
    }
    CK(ctx, cudaMemcpyAsync(&st, ctx->d_flush, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
    if (st.overflow) return go_global(ctx, in, 5, st.overflow);
    unsigned long long n1 = st.n_out;
    if (count_valid) valid_total = st.valid;
    unique_total += st.unique;
    if (ctx->imp_n) {  // rows of other ranks join as (key, count) items
        rc = reserve_items(ctx, ctx->w1, n1 + ctx->imp_n, wide_out, true, n1);
        if (rc != BC_OK) return rc;
        CK(ctx, cudaMemcpyAsync(ctx->w1.lo + n1, ctx->imp.lo, ctx->imp_n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (wide_out) CK(ctx, cudaMemcpyAsync(ctx->w1.hi + n1, ctx->imp.hi, ctx->imp_n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(ctx, cudaMemcpyAsync(ctx->w1.w + n1, ctx->imp.w, ctx->imp_n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        n1 += ctx->imp_n;
    }
    // ---- stage B: items partitioned by key, summed per key -> rows (after the rows of the one-stage part, if any)
    ctx->n_rows = rows_done;
    if (n1) {
        unsigned long long n_parts = 0;
        rc = reserve_items(ctx, ctx->w2, n1, wide_out, true);
        if (rc != BC_OK) return rc;
        if (rows_done == 0) rc = reserve_rows(ctx, n1, wide_out);
        else if (rows_done + n1 > ctx->row_cap) return go_global(ctx, in, 6, 0);  // cannot happen: sized for every valid record
        if (rc != BC_OK) return rc;
        const ItemView w1_now{ctx->w1.lo, wide_out ? ctx->w1.hi : nullptr, ctx->w1.w};
        rc = partition_items(ctx, w1_now, wide_out, n1, n1, true, ItemView{ctx->w2.lo, wide_out ? ctx->w2.hi : nullptr, ctx->w2.w}, false, 0,
                             0, &n_parts, nullptr, salt);
        if (rc == 1) return go_global(ctx, in, 7, 0);
        if (rc != BC_OK) return rc;
        CK(ctx, cudaMemcpyAsync(&ctx->d_flush->n_out, &rows_done, sizeof rows_done, cudaMemcpyHostToDevice, ctx->stream));
        {
            ProfScope p(ctx, BC_K_FINISH);
            CK(ctx, launch_reduce(RED_COUNT, wide_out, ItemView{ctx->w2.lo, wide_out ? ctx->w2.hi : nullptr, ctx->w2.w}, ctx->d_starts, n1,
                                  n_parts, 0, 0, ItemView{ctx->d_row_lo, wide_out ? ctx->d_row_hi : nullptr, ctx->d_row_cnt}, ctx->row_cap,
                                  ctx->d_flush, 0, HotList{}, ctx->stream));
        }
        CK(ctx, cudaMemcpyAsync(&st, ctx->d_flush, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
        if (st.overflow) return go_global(ctx, in, 8, st.overflow);
        ctx->n_rows = st.n_out;
    }
    ctx->rows_valid = true;
    ctx->flush_stages = stages ? 3 : 2;  // 3: one stage plus two stages for the hot keys
    ctx->last_valid = has_umi ? valid_total : n_valid;
    ctx->last_unique = has_umi ? unique_total : n_valid;
    return BC_OK;
}

// Fallback: the items through the global-memory set / map of bc_device.cuh (random DRAM accesses; counts add in 64 bits).
static int flush_global(bc_ctx* ctx, const FlushSrc& in) {
    Tables& T = ctx->tables;
    const bool wide_in = ctx->cfg.wide != 0, wide_out = T.map.wide != 0;
    drop_rows(ctx);
    ctx->flushed_global = true;
    free_table(T.map);
    free_table(T.set);
    CK(ctx, cudaMemsetAsync(ctx->d_counters + BC_N_COUNTERS, 0, 2 * sizeof(unsigned long long), ctx->stream));
    CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
    int rc = alloc_table(ctx, T.map, 1, wide_out, slots_for(in.n + ctx->imp_n), ctx->d_counters + BC_N_COUNTERS);
    if (rc == BC_OK && T.has_set) rc = alloc_table(ctx, T.set, 2, wide_in, slots_for(in.n), ctx->d_counters + BC_N_COUNTERS + 1);
    if (rc != BC_OK) return rc;
    {
        ProfScope p(ctx, BC_K_FINISH);
        CK(ctx, launch_insert_items(T, in.v, wide_in, in.n, ctx->d_flush, ctx->stream));
        if (ctx->imp_n) CK(ctx, launch_insert(T, ctx->imp.lo, wide_out ? ctx->imp.hi : nullptr, ctx->imp.w, ctx->imp_n, ctx->stream));
    }
    FlushStats st{};
    unsigned long long keys = 0;
    CK(ctx, cudaMemcpyAsync(&st, ctx->d_flush, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaMemcpyAsync(&keys, ctx->d_counters + BC_N_COUNTERS, sizeof keys, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    rc = reserve_rows(ctx, keys, wide_out);
    if (rc != BC_OK) return rc;
    CK(ctx, cudaMemsetAsync(ctx->d_row_n, 0, sizeof(unsigned long long), ctx->stream));
    {
        ProfScope p(ctx, BC_K_FINISH);
        CK(ctx, launch_compact(T.map, ctx->d_row_lo, wide_out ? ctx->d_row_hi : nullptr, ctx->d_row_cnt, ctx->d_row_n, ctx->stream));
    }
    CK(ctx, cudaMemcpyAsync(&ctx->n_rows, ctx->d_row_n, sizeof ctx->n_rows, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
    free_table(T.map);
    free_table(T.set);
    ctx->rows_valid = true;
    ctx->last_valid = st.valid;
    ctx->last_unique = T.has_set ? st.unique : st.valid;
    return BC_OK;
}

// The whole record buffer -> final rows and the matched / duplicates split.  A pure function of the buffer (plus the
// imported rows), so it may run any number of times as more batches arrive.
static int flush_records(bc_ctx* ctx) {
    if (ctx->rows_valid) return BC_OK;
    if (ctx->x_ranks > 1)
        return fail(ctx, BC_ESTATE, "this context is one rank of a multi-GPU job: run bc_exchange_count / _scatter / _finish after the "
                                    "last bc_submit before asking for counters or rows");
    int rc = fold_counters(ctx);
    if (rc == BC_OK) rc = bc_sync(ctx);
    if (rc != BC_OK) return rc;
    const unsigned long long n_rec = ctx->rec_n;
    // records that are not holes = reads counted "matched" so far (the repeats already moved to "duplicates" included)
    unsigned long long h[BC_N_COUNTERS];
    CK(ctx, cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    const unsigned long long n_valid = std::min(n_rec, h[BC_CNT_MATCHED] + h[BC_CNT_DUPLICATES]);
    const FlushSrc in{ItemView{ctx->rec.lo, ctx->cfg.wide ? ctx->rec.hi : nullptr, nullptr}, n_rec, n_valid};
    rc = flush_core(ctx, in);
    if (rc != BC_OK) return rc;
    return apply_duplicates(ctx, ctx->cfg.has_umi ? ctx->last_valid - ctx->last_unique : 0);
}

int bc_get_counters(bc_ctx* ctx, uint64_t out[BC_N_COUNTERS]) {
    if (!ctx || !out) return BC_EINVAL;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = fold_counters(ctx);
    if (rc == BC_OK) rc = bc_sync(ctx);
    if (rc != BC_OK) return rc;
    if (ctx->deferred) {  // the matched / duplicates split is known once the records are de-duplicated
        CK(ctx, cudaSetDevice(ctx->device));
        rc = flush_records(ctx);
        if (rc != BC_OK) return rc;
    }
    unsigned long long h[BC_N_COUNTERS];
    CK(ctx, cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    ctx->prof.d2h_bytes += sizeof h;
    for (int i = 0; i < BC_N_COUNTERS; i++) out[i] = h[i];
    return BC_OK;
}

static int decode_hook(bc_ctx* ctx, const bc_batch* batch, int flags, uint8_t* status, int16_t* offset, uint8_t* repaired,
                       int32_t* slot_index, uint64_t* key_lo, uint64_t* key_hi) {
    if (!ctx || !batch) return BC_EINVAL;
    const size_t n = batch->n_reads;
    if (n == 0) return BC_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    DecodeOut d{};
    const size_t ns = n * ctx->cfg.n_slots;
    auto cleanup = [&]() {
        if (d.status) cudaFree(d.status);
        if (d.offset) cudaFree(d.offset);
        if (d.repaired) cudaFree(d.repaired);
        if (d.slot_index) cudaFree(d.slot_index);
        if (d.key_lo) cudaFree(d.key_lo);
        if (d.key_hi) cudaFree(d.key_hi);
    };
#define CKH(call)                                                                                                  \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) {                                                                                   \
            cleanup();                                                                                             \
            return fail(ctx, BC_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);       \
        }                                                                                                          \
    } while (0)
    if (status) CKH(cudaMalloc(&d.status, n));
    if (offset) CKH(cudaMalloc(&d.offset, n * sizeof(int16_t)));
    if (repaired) CKH(cudaMalloc(&d.repaired, n));
    if (slot_index && ns) {
        CKH(cudaMalloc(&d.slot_index, ns * sizeof(int32_t)));
        CKH(cudaMemsetAsync(d.slot_index, 0xFF, ns * sizeof(int32_t), ctx->stream));
    }
    if (key_lo) CKH(cudaMalloc(&d.key_lo, n * sizeof(unsigned long long)));
    if (key_hi) CKH(cudaMalloc(&d.key_hi, n * sizeof(unsigned long long)));
    int rc = run_decode(ctx, batch, flags | F_EMIT, d, nullptr);
    if (rc != BC_OK) {
        cleanup();
        return rc;
    }
    CKH(cudaStreamSynchronize(ctx->stream));
    if (status) CKH(cudaMemcpy(status, d.status, n, cudaMemcpyDeviceToHost));
    if (offset) CKH(cudaMemcpy(offset, d.offset, n * sizeof(int16_t), cudaMemcpyDeviceToHost));
    if (repaired) CKH(cudaMemcpy(repaired, d.repaired, n, cudaMemcpyDeviceToHost));
    if (slot_index && ns) CKH(cudaMemcpy(slot_index, d.slot_index, ns * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (key_lo) CKH(cudaMemcpy(key_lo, d.key_lo, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (key_hi) CKH(cudaMemcpy(key_hi, d.key_hi, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
#undef CKH
    cleanup();
    return BC_OK;
}

int bc_locate_only(bc_ctx* ctx, const bc_batch* batch, bc_locate_out* out) {
    if (!out) return BC_EINVAL;
    return decode_hook(ctx, batch, F_LOCATE_ONLY, out->status, out->offset, out->repaired, nullptr, nullptr, nullptr);
}

int bc_decode_only(bc_ctx* ctx, const bc_batch* batch, bc_decode_out* out) {
    if (!out) return BC_EINVAL;
    return decode_hook(ctx, batch, 0, out->status, out->offset, out->repaired, out->slot_index, out->key_lo, out->key_hi);
}

// ---------------------------------------------------------------------------------------------- finish / enrich

void bc_table_free(bc_table* t) {
    if (!t) return;
    if (!(t->flags & BC_TABLE_BORROWED)) {
        free(t->key_lo);
        free(t->key_hi);
        free(t->count);
        free(t->mask);
    }
    memset(t, 0, sizeof *t);
}

// appends device rows to a malloc-owned host table (small tables: enrichment marginals)
static int rows_to_host(bc_ctx* ctx, const unsigned long long* lo, const unsigned long long* hi, const unsigned long long* cnt,
                        unsigned long long n, uint32_t mask, bc_table* t) {
    const uint64_t old = t->n_rows;
    const uint64_t tot = old + n;
    if (tot == 0) return BC_OK;
    t->key_lo = (uint64_t*)realloc(t->key_lo, tot * sizeof(uint64_t));
    t->key_hi = (uint64_t*)realloc(t->key_hi, tot * sizeof(uint64_t));
    t->count = (uint64_t*)realloc(t->count, tot * sizeof(uint64_t));
    t->mask = (uint32_t*)realloc(t->mask, tot * sizeof(uint32_t));
    if (!t->key_lo || !t->key_hi || !t->count || !t->mask) return fail(ctx, BC_ENOMEM, "host rows");
    if (n) {
        CK(ctx, cudaMemcpy(t->key_lo + old, lo, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        if (hi) CK(ctx, cudaMemcpy(t->key_hi + old, hi, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        else memset(t->key_hi + old, 0, n * sizeof(uint64_t));
        CK(ctx, cudaMemcpy(t->count + old, cnt, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        ctx->prof.d2h_bytes += (hi ? 3 : 2) * n * sizeof(uint64_t);
    }
    for (uint64_t i = old; i < tot; i++) t->mask[i] = mask;
    t->n_rows = tot;
    return BC_OK;
}

// the map's occupied entries -> ctx row buffers (device)
static int build_rows(bc_ctx* ctx) {
    if (ctx->rows_valid) return BC_OK;
    if (ctx->deferred) return flush_records(ctx);
    int rc = fold_counters(ctx);
    if (rc == BC_OK) rc = bc_sync(ctx);
    if (rc != BC_OK) return rc;
    drop_rows(ctx);
    const DevTable& M = ctx->tables.map;
    unsigned long long upper;
    if (M.kind == 0) {  // at most one row per counter and per matched read
        unsigned long long h[BC_N_COUNTERS];
        CK(ctx, cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
        upper = std::min<unsigned long long>(h[BC_CNT_MATCHED] + ctx->imported_rows, M.cap);
    } else {
        CK(ctx, cudaMemcpy(&upper, ctx->d_counters + BC_N_COUNTERS, sizeof upper, cudaMemcpyDeviceToHost));
    }
    rc = reserve_rows(ctx, upper, M.wide != 0);
    if (rc != BC_OK) return rc;
    CK(ctx, cudaMemsetAsync(ctx->d_row_n, 0, sizeof(unsigned long long), ctx->stream));
    {
        ProfScope p(ctx, BC_K_FINISH);
        CK(ctx, launch_compact(M, ctx->d_row_lo, M.wide ? ctx->d_row_hi : nullptr, ctx->d_row_cnt, ctx->d_row_n, ctx->stream));
    }
    CK(ctx, cudaMemcpyAsync(&ctx->n_rows, ctx->d_row_n, sizeof ctx->n_rows, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->rows_valid = true;
    return BC_OK;
}

int bc_finish(bc_ctx* ctx, bc_table* rows) {
    if (!ctx || !rows) return BC_EINVAL;
    CK(ctx, cudaSetDevice(ctx->device));
    memset(rows, 0, sizeof *rows);
    int rc = build_rows(ctx);
    if (rc != BC_OK) return rc;
    const unsigned long long n = ctx->n_rows;
    const bool wide = ctx->tables.map.wide != 0;
    if (n > ctx->h_row_cap || (wide && !ctx->h_row_hi)) {  // pinned mirror, grow-only
        if (ctx->h_row_lo) cudaFreeHost(ctx->h_row_lo);
        if (ctx->h_row_hi) cudaFreeHost(ctx->h_row_hi);
        if (ctx->h_row_cnt) cudaFreeHost(ctx->h_row_cnt);
        ctx->h_row_lo = ctx->h_row_hi = ctx->h_row_cnt = nullptr;
        ctx->h_row_cap = 0;
        const unsigned long long cap = std::max<unsigned long long>(n + n / 8, 1024);
        CK(ctx, cudaHostAlloc((void**)&ctx->h_row_lo, cap * sizeof(uint64_t), cudaHostAllocDefault));
        CK(ctx, cudaHostAlloc((void**)&ctx->h_row_cnt, cap * sizeof(uint64_t), cudaHostAllocDefault));
        if (wide) CK(ctx, cudaHostAlloc((void**)&ctx->h_row_hi, cap * sizeof(uint64_t), cudaHostAllocDefault));
        ctx->h_row_cap = cap;
    }
    if (n) {
        CK(ctx, cudaMemcpyAsync(ctx->h_row_lo, ctx->d_row_lo, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaMemcpyAsync(ctx->h_row_cnt, ctx->d_row_cnt, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (wide) CK(ctx, cudaMemcpyAsync(ctx->h_row_hi, ctx->d_row_hi, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->prof.d2h_bytes += (wide ? 3 : 2) * n * sizeof(uint64_t);
    }
    rows->n_rows = n;
    rows->key_lo = ctx->h_row_lo;
    rows->key_hi = wide ? ctx->h_row_hi : nullptr;
    rows->count = ctx->h_row_cnt;
    rows->flags = BC_TABLE_BORROWED;
    return BC_OK;
}

// key mask keeping the sample field and the counted barcodes listed in `keep` (bit k = k-th counted barcode)
static void marginal_mask(const bc_ctx* ctx, uint32_t keep, Key* mask, uint32_t* kept_bits) {
    Key m{0, 0};
    uint32_t bits = 0;
    auto add = [&](const KeyField& f) {
        const uint32_t sh = f.shift - ctx->cfg.umi_bits;
        for (uint32_t b = sh; b < sh + f.bits; b++) {
            if (b < 64) m.lo |= 1ull << b;
            else m.hi |= 1ull << (b - 64);
        }
        bits += f.bits;
    };
    for (size_t k = 0; k < ctx->counted_slots.size(); k++)
        if (keep & (1u << k)) add(ctx->fields[ctx->counted_slots[k]]);
    if (ctx->sample_slot >= 0) add(ctx->fields[ctx->sample_slot]);
    *mask = m;
    *kept_bits = bits;
}

static int one_marginal(bc_ctx* ctx, uint32_t keep, bc_table* out) {
    Key mask;
    uint32_t kept_bits;
    marginal_mask(ctx, keep, &mask, &kept_bits);
    unsigned long long entries = ctx->n_rows ? ctx->n_rows : 1;
    if (kept_bits < 40 && (1ull << kept_bits) < entries) entries = 1ull << kept_bits;
    unsigned long long* d_n = nullptr;
    CK(ctx, cudaMalloc(&d_n, sizeof(unsigned long long)));
    CK(ctx, cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), ctx->stream));
    DevTable t;
    const int wide = ctx->tables.map.wide;
    int rc = alloc_table(ctx, t, 1, wide, slots_for(entries), d_n);
    unsigned long long *lo = nullptr, *hi = nullptr, *cnt = nullptr, n = 0;
    if (rc == BC_OK) {
        {
            ProfScope p(ctx, BC_K_ENRICH);
            cudaError_t e = launch_marginal(ctx->d_row_lo, wide ? ctx->d_row_hi : nullptr, ctx->d_row_cnt, ctx->n_rows, mask, t, ctx->stream);
            if (e != cudaSuccess) rc = fail(ctx, BC_ECUDA, "launch_marginal: %s", cudaGetErrorString(e));
        }
        unsigned long long keys = 0;
        if (rc == BC_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, BC_ECUDA, "marginal sync");
        if (rc == BC_OK && cudaMemcpy(&keys, d_n, sizeof keys, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(ctx, BC_ECUDA, "marginal count");
        const unsigned long long cap = keys ? keys : 1;
        if (rc == BC_OK && (cudaMalloc(&lo, cap * 8) != cudaSuccess || cudaMalloc(&cnt, cap * 8) != cudaSuccess ||
                            (wide && cudaMalloc(&hi, cap * 8) != cudaSuccess)))
            rc = fail(ctx, BC_ENOMEM, "marginal rows");
        if (rc == BC_OK) {
            cudaMemsetAsync(ctx->d_row_n, 0, sizeof(unsigned long long), ctx->stream);
            {
                ProfScope p(ctx, BC_K_ENRICH);
                cudaError_t e = launch_compact(t, lo, hi, cnt, ctx->d_row_n, ctx->stream);
                if (e != cudaSuccess) rc = fail(ctx, BC_ECUDA, "launch_compact: %s", cudaGetErrorString(e));
            }
            if (rc == BC_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, BC_ECUDA, "compact sync");
            if (rc == BC_OK && cudaMemcpy(&n, ctx->d_row_n, sizeof n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(ctx, BC_ECUDA, "compact count");
        }
        if (rc == BC_OK) rc = rows_to_host(ctx, lo, hi, cnt, n, keep, out);
    }
    if (lo) cudaFree(lo);
    if (hi) cudaFree(hi);
    if (cnt) cudaFree(cnt);
    free_table(t);
    cudaFree(d_n);
    return rc;
}

// Dense plan for the marginals: possible when every counted barcode (and the sample barcode, if any) is index-coded and
// the counters of all singles and doubles together stay within a budget.
static void plan_marginals(bc_ctx* ctx) {
    MargPlan& P = ctx->marg;
    P = MargPlan{};
    ctx->marg_dense = false;
    const uint32_t k = (uint32_t)ctx->counted_slots.size();
    if (k == 0 || k > (uint32_t)kMaxSlots) return;
    const uint32_t drop = ctx->cfg.umi_bits;
    for (uint32_t a = 0; a < k; a++) {
        const KeyField& f = ctx->fields[ctx->counted_slots[a]];
        if (f.raw || f.bits > 24) return;
        P.f_shift[a] = f.shift - drop;
        P.f_bits[a] = f.bits;
    }
    if (ctx->sample_slot >= 0) {
        const KeyField& f = ctx->fields[ctx->sample_slot];
        if (f.raw || f.bits > 24) return;
        P.s_shift = f.shift - drop;
        P.s_bits = f.bits;
    }
    const unsigned long long budget = 1ULL << 26;  // counters (512 MB)
    unsigned long long off = 0;
    P.k = k;
    for (uint32_t a = 0; a < k; a++) {
        P.s_off[a] = off;
        off += 1ULL << (P.s_bits + P.f_bits[a]);
        if (off > budget) return;
    }
    P.n_single = off;
    for (uint32_t a = 0; a + 1 < k; a++)
        for (uint32_t b = a + 1; b < k; b++) {
            const uint32_t bits = P.s_bits + P.f_bits[a] + P.f_bits[b];
            if (bits > 40) return;
            P.pa[P.n_pairs] = (uint8_t)a;
            P.pb[P.n_pairs] = (uint8_t)b;
            P.d_off[P.n_pairs++] = off;
            off += 1ULL << bits;
            if (off > budget) return;
        }
    P.n_total = off;
    ctx->marg_dense = true;
}

static int compute_marginals(bc_ctx* ctx) {
    if (ctx->marg_valid) return BC_OK;
    if (!ctx->d_marg) CK(ctx, cudaMalloc(&ctx->d_marg, ctx->marg.n_total * sizeof(unsigned long long)));
    CK(ctx, cudaMemsetAsync(ctx->d_marg, 0, ctx->marg.n_total * sizeof(unsigned long long), ctx->stream));
    {
        ProfScope p(ctx, BC_K_ENRICH);
        CK(ctx, launch_marginals_dense(ctx->d_row_lo, ctx->tables.map.wide ? ctx->d_row_hi : nullptr, ctx->d_row_cnt, ctx->n_rows, ctx->marg,
                                       ctx->d_marg, ctx->stream));
    }
    ctx->marg_valid = true;
    return BC_OK;
}

// marginals [m_first, m_first + m_count) of the dense arrays -> host rows
static int marginals_to_host(bc_ctx* ctx, uint32_t m_first, uint32_t m_count, bc_table* out) {
    if (m_count == 0) return BC_OK;
    unsigned long long n = 0;
    CK(ctx, cudaMemsetAsync(ctx->d_row_n, 0, sizeof(unsigned long long), ctx->stream));
    {
        ProfScope p(ctx, BC_K_ENRICH);
        CK(ctx, launch_marginals_rows(ctx->marg, m_first, m_count, ctx->d_marg, nullptr, nullptr, nullptr, nullptr, ctx->d_row_n, ctx->stream));
    }
    CK(ctx, cudaMemcpyAsync(&n, ctx->d_row_n, sizeof n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (n == 0) return BC_OK;
    unsigned long long *lo = nullptr, *hi = nullptr, *cnt = nullptr;
    uint32_t* mask = nullptr;
    int rc = BC_OK;
    if (cudaMalloc(&lo, n * 8) != cudaSuccess || cudaMalloc(&hi, n * 8) != cudaSuccess || cudaMalloc(&cnt, n * 8) != cudaSuccess ||
        cudaMalloc(&mask, n * 4) != cudaSuccess)
        rc = fail(ctx, BC_ENOMEM, "marginal rows");
    if (rc == BC_OK) {
        cudaMemsetAsync(ctx->d_row_n, 0, sizeof(unsigned long long), ctx->stream);
        ProfScope p(ctx, BC_K_ENRICH);
        cudaError_t e = launch_marginals_rows(ctx->marg, m_first, m_count, ctx->d_marg, lo, hi, cnt, mask, ctx->d_row_n, ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, BC_ECUDA, "launch_marginals_rows: %s", cudaGetErrorString(e));
    }
    if (rc == BC_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, BC_ECUDA, "marginal rows sync");
    if (rc == BC_OK) {
        const uint64_t old = out->n_rows, tot = old + n;
        out->key_lo = (uint64_t*)realloc(out->key_lo, tot * sizeof(uint64_t));
        out->key_hi = (uint64_t*)realloc(out->key_hi, tot * sizeof(uint64_t));
        out->count = (uint64_t*)realloc(out->count, tot * sizeof(uint64_t));
        out->mask = (uint32_t*)realloc(out->mask, tot * sizeof(uint32_t));
        if (!out->key_lo || !out->key_hi || !out->count || !out->mask) rc = fail(ctx, BC_ENOMEM, "host rows");
        if (rc == BC_OK &&
            (cudaMemcpy(out->key_lo + old, lo, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
             cudaMemcpy(out->key_hi + old, hi, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
             cudaMemcpy(out->count + old, cnt, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
             cudaMemcpy(out->mask + old, mask, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess))
            rc = fail(ctx, BC_ECUDA, "marginal rows D2H");
        if (rc == BC_OK) {
            out->n_rows = tot;
            ctx->prof.d2h_bytes += n * 28;
        }
    }
    if (lo) cudaFree(lo);
    if (hi) cudaFree(hi);
    if (cnt) cudaFree(cnt);
    if (mask) cudaFree(mask);
    return rc;
}

int bc_marginals(bc_ctx* ctx, uint64_t** dev_counters, uint64_t* n) {
    if (!ctx || !dev_counters || !n) return BC_EINVAL;
    *dev_counters = nullptr;
    *n = 0;
    CK(ctx, cudaSetDevice(ctx->device));
    if (!ctx->marg_dense) return BC_OK;
    int rc = build_rows(ctx);
    if (rc == BC_OK) rc = compute_marginals(ctx);
    if (rc != BC_OK) return rc;
    *dev_counters = reinterpret_cast<uint64_t*>(ctx->d_marg);
    *n = ctx->marg.n_total;
    return BC_OK;
}

int bc_enrich(bc_ctx* ctx, bc_table* singles, bc_table* doubles) {
    if (!ctx || !singles) return BC_EINVAL;
    CK(ctx, cudaSetDevice(ctx->device));
    memset(singles, 0, sizeof *singles);
    if (doubles) memset(doubles, 0, sizeof *doubles);
    int rc = build_rows(ctx);
    if (rc != BC_OK) return rc;
    const uint32_t k = (uint32_t)ctx->counted_slots.size();
    if (ctx->marg_dense) {  // one pass over the rows fills every marginal; the arrays may have been merged across ranks since
        rc = compute_marginals(ctx);
        if (rc == BC_OK) rc = marginals_to_host(ctx, 0, k, singles);
        if (rc == BC_OK && doubles) rc = marginals_to_host(ctx, k, ctx->marg.n_pairs, doubles);
        return rc;
    }
    // raw (file-less) barcodes have no dense index space: one hash map per marginal
    for (uint32_t a = 0; a < k; a++) {  // info.rs:840-866
        rc = one_marginal(ctx, 1u << a, singles);
        if (rc != BC_OK) return rc;
    }
    if (doubles) {
        for (uint32_t a = 0; a + 1 < k; a++)  // info.rs:869-904
            for (uint32_t b = a + 1; b < k; b++) {
                rc = one_marginal(ctx, (1u << a) | (1u << b), doubles);
                if (rc != BC_OK) return rc;
            }
    }
    return BC_OK;
}

int bc_key_decode(const bc_ctx* ctx, uint64_t key_lo, uint64_t key_hi, uint32_t mask, int with_umi, int32_t* idx_out,
                  char* str_out, uint32_t str_stride) {
    if (!ctx || !idx_out) return BC_EINVAL;
    auto get_bits = [&](uint32_t shift, uint32_t bits) -> uint64_t {
        if (bits == 0) return 0;
        uint64_t v;
        if (shift >= 64) v = key_hi >> (shift - 64);
        else v = (key_lo >> shift) | (shift ? (key_hi << (64 - shift)) : 0);
        return bits >= 64 ? v : (v & ((1ull << bits) - 1));
    };
    const uint32_t drop = with_umi ? 0 : ctx->cfg.umi_bits;
    for (uint32_t s = 0; s < ctx->cfg.n_slots; s++) {
        idx_out[s] = -1;
        if (str_out) str_out[(size_t)s * str_stride] = 0;
        const KeyField& f = ctx->fields[s];
        const DevSlot& D = ctx->cfg.slots[s];
        if ((int)s == ctx->umi_slot && !with_umi) continue;
        if (D.kind == 'B' && mask) {
            uint32_t k = 0;
            for (; k < ctx->counted_slots.size(); k++)
                if (ctx->counted_slots[k] == (int)s) break;
            if (!(mask & (1u << k))) continue;
        }
        const uint32_t sh = f.shift - drop;
        if (!f.raw) {
            idx_out[s] = (int32_t)get_bits(sh, f.bits);
        } else if (str_out) {
            if (str_stride < (uint32_t)D.len + 1) return BC_EINVAL;
            const uint32_t lo = (uint32_t)get_bits(sh, D.len), hi = (uint32_t)get_bits(sh + D.len, D.len),
                           nm = (uint32_t)get_bits(sh + 2 * D.len, D.len);
            char* o = str_out + (size_t)s * str_stride;
            for (uint32_t p = 0; p < D.len; p++) {
                const uint32_t b = 1u << p;
                o[p] = (nm & b) ? 'N' : "ACGT"[((lo & b) ? 1 : 0) | ((hi & b) ? 2 : 0)];
            }
            o[D.len] = 0;
        }
    }
    return BC_OK;
}

// ---------------------------------------------------------------------------------------------- multi-GPU
// One exchange at the flush: record -> owner = hash(key without its random barcode) % n_ranks, written by the partitioning
// kernel straight into the owner's receive buffer (NVLink peer memory); every owner then runs the usual flush over what
// it received.  The owner hash uses its own salt so that it is independent of the hash that partitions an owner's records.
static const unsigned long long kOwnerSalt = 0xBB67AE8584CAA73BULL;

static SplitLevel owner_level(const bc_ctx* ctx) {
    return SplitLevel{ctx->x_ranks, 0u, 0xFFFFFFFFu, ctx->x_ranks, ctx->cfg.umi_bits, kOwnerSalt};
}

static void close_peers(bc_ctx* ctx) {
    for (uint32_t r = 0; r < (uint32_t)kMaxRanks; r++) {
        if (ctx->x_peer_ipc[r] && ctx->x_peer[r]) cudaIpcCloseMemHandle(ctx->x_peer[r]);
        ctx->x_peer[r] = nullptr;
        ctx->x_peer_ipc[r] = false;
    }
}

int bc_exchange_open(bc_ctx* ctx, uint32_t n_ranks, uint32_t rank, uint64_t capacity) {
    if (!ctx || n_ranks == 0 || n_ranks > (uint32_t)kMaxRanks || rank >= n_ranks || capacity == 0) return BC_EINVAL;
    if (!ctx->deferred)
        return fail(ctx, BC_ESTATE, "bc_exchange_open: this scheme counts into a dense table; merge ranks with bc_dense_counts / bc_peer_add");
    if (capacity >= 0xFFFFFFF0ULL) return fail(ctx, BC_EUNSUPPORTED, "bc_exchange_open: more than 2^32 records per owner");
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    close_peers(ctx);
    if (ctx->d_xrecv) cudaFree(ctx->d_xrecv);
    ctx->d_xrecv = nullptr;
    const size_t buf_words = (size_t)capacity * (ctx->cfg.wide ? 2 : 1);
    CK(ctx, cudaMalloc(&ctx->d_xrecv, (2 * buf_words + 32) * sizeof(unsigned long long)));
    CK(ctx, cudaMemset(ctx->d_xrecv + 2 * buf_words, 0, 32 * sizeof(unsigned long long)));  // the two receive cursors
    if (!ctx->d_xcursor) CK(ctx, cudaMalloc(&ctx->d_xcursor, kMaxRanks * sizeof(uint32_t)));
    if (!ctx->d_xsent) {
        CK(ctx, cudaMalloc(&ctx->d_xsent, kMaxRanks * sizeof(unsigned long long)));
        CK(ctx, cudaMalloc(&ctx->d_xoverflow, sizeof(unsigned int)));
        CK(ctx, cudaStreamCreateWithFlags(&ctx->x_stream, cudaStreamNonBlocking));
        CK(ctx, cudaEventCreateWithFlags(&ctx->x_ev, cudaEventDisableTiming));
    }
    ctx->xcap = capacity;
    ctx->x_ranks = n_ranks;
    ctx->x_rank = rank;
    ctx->x_peer[rank] = ctx->d_xrecv;
    ctx->x_state = 0;
    ctx->x_epoch = 0;
    ctx->x_mode = ctx->rec_n ? 2 : 0;  // records decoded before the buffers existed (a skewed job growing them) go in bulk
    ctx->x_overflowed = ctx->x_dirty = false;
    ctx->rows_valid = false;
    return BC_OK;
}

uint64_t bc_exchange_capacity(const bc_ctx* ctx) {
    if (!ctx || !ctx->d_xrecv || ctx->x_ranks == 0 || ctx->x_dirty) return 0;
    for (uint32_t r = 0; r < ctx->x_ranks; r++)
        if (!ctx->x_peer[r]) return 0;
    return ctx->xcap;
}

int bc_exchange_disconnect(bc_ctx* ctx) {
    if (!ctx) return BC_EINVAL;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    close_peers(ctx);
    if (ctx->d_xrecv) ctx->x_peer[ctx->x_rank] = ctx->d_xrecv;
    return BC_OK;
}

int bc_exchange_handle(bc_ctx* ctx, void* ipc_handle_out) {
    if (!ctx || !ipc_handle_out) return BC_EINVAL;
    if (!ctx->d_xrecv) return fail(ctx, BC_ESTATE, "bc_exchange_handle before bc_exchange_open");
    static_assert(sizeof(cudaIpcMemHandle_t) == BC_IPC_HANDLE_BYTES, "BC_IPC_HANDLE_BYTES");
    CK(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(ctx, cudaIpcGetMemHandle(&h, ctx->d_xrecv));
    memcpy(ipc_handle_out, &h, sizeof h);
    return BC_OK;
}

int bc_exchange_connect(bc_ctx* ctx, const void* ipc_handles) {
    if (!ctx || !ipc_handles) return BC_EINVAL;
    if (!ctx->d_xrecv) return fail(ctx, BC_ESTATE, "bc_exchange_connect before bc_exchange_open");
    CK(ctx, cudaSetDevice(ctx->device));
    for (uint32_t r = 0; r < ctx->x_ranks; r++) {
        if (r == ctx->x_rank) continue;
        if (ctx->x_peer_ipc[r] && ctx->x_peer[r]) cudaIpcCloseMemHandle(ctx->x_peer[r]);
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(ipc_handles) + (size_t)r * sizeof h, sizeof h);
        void* p = nullptr;
        CK(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->x_peer[r] = static_cast<unsigned long long*>(p);
        ctx->x_peer_ipc[r] = true;
    }
    return BC_OK;
}

static int enable_peer(bc_ctx* ctx, int other_device) {
    if (other_device == ctx->device) return BC_OK;
    int can = 0;
    CK(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, other_device));
    if (!can) return fail(ctx, BC_EUNSUPPORTED, "device %d cannot access device %d (no NVLink / PCIe peer path)", ctx->device, other_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(other_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
    else if (e != cudaSuccess) return fail(ctx, BC_ECUDA, "cudaDeviceEnablePeerAccess(%d): %s", other_device, cudaGetErrorString(e));
    return BC_OK;
}

int bc_exchange_connect_local(bc_ctx* ctx, bc_ctx* const* ranks) {
    if (!ctx || !ranks) return BC_EINVAL;
    if (!ctx->d_xrecv) return fail(ctx, BC_ESTATE, "bc_exchange_connect_local before bc_exchange_open");
    CK(ctx, cudaSetDevice(ctx->device));
    for (uint32_t r = 0; r < ctx->x_ranks; r++) {
        if (r == ctx->x_rank) continue;
        bc_ctx* o = ranks[r];
        if (!o || !o->d_xrecv || o->xcap != ctx->xcap || o->x_ranks != ctx->x_ranks || o->x_rank != r || o->cfg.wide != ctx->cfg.wide)
            return fail(ctx, BC_EINVAL, "bc_exchange_connect_local: rank %u is not open with the same geometry", r);
        int rc = enable_peer(ctx, o->device);
        if (rc != BC_OK) return rc;
        if (ctx->x_peer_ipc[r] && ctx->x_peer[r]) cudaIpcCloseMemHandle(ctx->x_peer[r]);
        ctx->x_peer[r] = o->d_xrecv;
        ctx->x_peer_ipc[r] = false;
    }
    return BC_OK;
}

int bc_exchange_count(bc_ctx* ctx, uint64_t* sent) {
    if (!ctx || !sent) return BC_EINVAL;
    if (ctx->x_ranks == 0) return fail(ctx, BC_ESTATE, "bc_exchange_count before bc_exchange_open");
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = fold_counters(ctx);
    if (rc == BC_OK) rc = bc_sync(ctx);
    if (rc != BC_OK) return rc;
    const unsigned long long n_rec = ctx->rec_n;
    if (n_rec >= 0xFFFFFFF0ULL) return fail(ctx, BC_EUNSUPPORTED, "bc_exchange_count: more than 2^32 records on one rank");
    if (ctx->x_mode == 1) {  // streamed: the records have left already, batch by batch; what went where was counted on the way
        CK(ctx, cudaStreamSynchronize(ctx->x_stream));
        unsigned long long hs[kMaxRanks];
        unsigned int over = 0;
        CK(ctx, cudaMemcpy(hs, ctx->d_xsent, sizeof hs, cudaMemcpyDeviceToHost));
        CK(ctx, cudaMemcpy(&over, ctx->d_xoverflow, sizeof over, cudaMemcpyDeviceToHost));
        ctx->x_overflowed = over != 0;  // an owner's buffer was too small: the caller sees it in the matrix and re-opens larger
        ctx->x_local_valid = 0;
        for (uint32_t r = 0; r < ctx->x_ranks; r++) {
            sent[r] = ctx->x_sent[r] = hs[r];
            ctx->x_local_valid += hs[r];
        }
        ctx->x_state = 1;
        return BC_OK;
    }
    uint32_t h[kMaxRanks] = {0};
    FlushStats st{};
    if (n_rec) {
        const bool wide = ctx->cfg.wide != 0;
        CK(ctx, cudaMemsetAsync(ctx->d_xcursor, 0, kMaxRanks * sizeof(uint32_t), ctx->stream));
        CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
        {
            ProfScope p(ctx, BC_K_EXCHANGE);
            CK(ctx, launch_split(false, wide, ItemView{ctx->rec.lo, wide ? ctx->rec.hi : nullptr, nullptr}, ItemView{}, nullptr, 1, n_rec,
                                 owner_level(ctx), ctx->d_xcursor, ctx->d_flush, true, ctx->stream));
        }
        CK(ctx, cudaMemcpyAsync(h, ctx->d_xcursor, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaMemcpyAsync(&st, ctx->d_flush, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
    }
    ctx->x_local_valid = st.valid;
    for (uint32_t r = 0; r < ctx->x_ranks; r++) sent[r] = ctx->x_sent[r] = h[r];
    ctx->x_state = 1;
    return BC_OK;
}

int bc_exchange_scatter(bc_ctx* ctx, const uint64_t* first) {
    if (!ctx || !first) return BC_EINVAL;
    if (ctx->x_state != 1) return fail(ctx, BC_ESTATE, "bc_exchange_scatter: call bc_exchange_count first (and nothing may be submitted in between)");
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->x_mode == 1) {
        if (ctx->x_overflowed)
            return fail(ctx, BC_EINVAL, "bc_exchange_scatter: an owner's receive buffer was too small for the streamed records: every rank "
                                        "re-opens larger (bc_exchange_open) and the exchange starts over");
        ctx->x_state = 2;  // nothing left to move
        return BC_OK;
    }
    uint32_t cur[kMaxRanks] = {0};
    PeerOut peers{};
    const uint32_t parity = ctx->x_epoch & 1u;
    for (uint32_t r = 0; r < ctx->x_ranks; r++) {
        if (!ctx->x_peer[r]) return fail(ctx, BC_ESTATE, "bc_exchange_scatter: rank %u is not connected", r);
        if (first[r] + ctx->x_sent[r] > ctx->xcap)
            return fail(ctx, BC_EINVAL, "bc_exchange_scatter: rank %u would receive past its capacity (%llu + %llu > %llu): reopen larger", r,
                        (unsigned long long)first[r], ctx->x_sent[r], ctx->xcap);
        cur[r] = (uint32_t)first[r];
        peers.lo[r] = xbuf_lo(ctx, ctx->x_peer[r], parity);
        peers.hi[r] = ctx->cfg.wide ? peers.lo[r] + ctx->xcap : nullptr;
    }
    const bool wide = ctx->cfg.wide != 0;
    memcpy(ctx->h_xcur, cur, sizeof cur);  // outlives this call: the copy below is asynchronous
    CK(ctx, cudaMemcpyAsync(ctx->d_xcursor, ctx->h_xcur, sizeof cur, cudaMemcpyHostToDevice, ctx->stream));
    {
        ProfScope p(ctx, BC_K_EXCHANGE);
        CK(ctx, launch_owner_scatter(wide, ItemView{ctx->rec.lo, wide ? ctx->rec.hi : nullptr, nullptr}, peers, ctx->rec_n, owner_level(ctx),
                                     ctx->d_xcursor, ctx->stream));
    }
    ctx->x_state = 2;
    return BC_OK;
}

int bc_exchange_finish(bc_ctx* ctx, uint64_t n_received) {
    if (!ctx) return BC_EINVAL;
    if (ctx->x_state != 2) return fail(ctx, BC_ESTATE, "bc_exchange_finish: call bc_exchange_scatter first");
    if (n_received > ctx->xcap) return fail(ctx, BC_EINVAL, "bc_exchange_finish: %llu records exceed the receive capacity", (unsigned long long)n_received);
    CK(ctx, cudaSetDevice(ctx->device));
    const bool wide = ctx->cfg.wide != 0;
    const uint32_t parity = ctx->x_epoch & 1u;
    {   // streamed senders advanced this buffer's cursor (also when this rank itself decoded nothing in this job): what arrived
        // must agree with the caller's plan; a bulk exchange leaves the cursor at 0
        unsigned long long got = 0;
        CK(ctx, cudaMemcpy(&got, xbuf_cursor(ctx, ctx->d_xrecv, parity), sizeof got, cudaMemcpyDeviceToHost));
        if (got != 0 && got != n_received)
            return fail(ctx, BC_ESTATE, "bc_exchange_finish: %llu records arrived, the caller's plan says %llu (was the barrier after the "
                                        "last rank's bc_exchange_count skipped?)", got, (unsigned long long)n_received);
        if (got) CK(ctx, cudaMemset(xbuf_cursor(ctx, ctx->d_xrecv, parity), 0, sizeof got));  // for the job after next
    }
    unsigned long long* recv = xbuf_lo(ctx, ctx->d_xrecv, parity);
    const FlushSrc in{ItemView{recv, wide ? recv + ctx->xcap : nullptr, nullptr}, n_received, n_received};
    ctx->x_epoch++;
    int rc = flush_core(ctx, in);
    if (rc != BC_OK) return rc;
    // this rank's share of the job's outcome: the records it owns (matched = distinct pairs, duplicates = their repeats)
    rc = fold_counters(ctx);
    if (rc == BC_OK) rc = bc_sync(ctx);
    if (rc != BC_OK) return rc;
    unsigned long long h[BC_N_COUNTERS];
    CK(ctx, cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    h[BC_CNT_MATCHED] = ctx->last_unique;
    h[BC_CNT_DUPLICATES] = ctx->last_valid - ctx->last_unique;
    CK(ctx, cudaMemcpy(ctx->d_counters, h, sizeof h, cudaMemcpyHostToDevice));
    ctx->dup_applied = 0;
    ctx->x_received = n_received;
    ctx->x_state = 3;
    return BC_OK;
}

// dst += src for the dense count table (what = 0) or the dense enrichment marginals (what = 1) of two contexts of this
// process; src may live on another GPU (peer access is enabled as needed).  Both must be idle (synchronised).
int bc_peer_add(bc_ctx* dst, bc_ctx* src, int what) {
    if (!dst || !src || dst == src) return BC_EINVAL;
    CK(dst, cudaSetDevice(dst->device));
    int rc = enable_peer(dst, src->device);
    if (rc != BC_OK) return rc;
    if (what == BC_ADD_DENSE_COUNTS) {
        if (dst->deferred || src->deferred || dst->tables.map.kind != 0 || src->tables.map.kind != 0 || dst->tables.map.cap != src->tables.map.cap)
            return fail(dst, BC_ESTATE, "bc_peer_add: both contexts need the same dense count table");
        CK(dst, launch_add_u64(dst->tables.map.data, src->tables.map.data, dst->tables.map.cap, dst->stream));
        dst->imported_rows = dst->tables.map.cap;
        dst->rows_valid = false;
        dst->marg_valid = false;
    } else if (what == BC_ADD_MARGINALS) {
        if (!dst->marg_dense || !src->marg_dense || !dst->marg_valid || !src->marg_valid || dst->marg.n_total != src->marg.n_total)
            return fail(dst, BC_ESTATE, "bc_peer_add: both contexts need computed dense marginals (bc_marginals)");
        CK(dst, launch_add_u64(dst->d_marg, src->d_marg, dst->marg.n_total, dst->stream));
    } else {
        return BC_EINVAL;
    }
    CK(dst, cudaStreamSynchronize(dst->stream));
    return BC_OK;
}

int bc_export_rows(bc_ctx* ctx, uint64_t** dev_key_lo, uint64_t** dev_key_hi, uint64_t** dev_count, uint64_t* n_rows) {
    if (!ctx || !dev_key_lo || !dev_key_hi || !dev_count || !n_rows) return BC_EINVAL;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = build_rows(ctx);
    if (rc != BC_OK) return rc;
    *dev_key_lo = reinterpret_cast<uint64_t*>(ctx->d_row_lo);
    *dev_key_hi = ctx->tables.map.wide ? reinterpret_cast<uint64_t*>(ctx->d_row_hi) : nullptr;
    *dev_count = reinterpret_cast<uint64_t*>(ctx->d_row_cnt);
    *n_rows = ctx->n_rows;
    return BC_OK;
}

// Adds (key, count) rows — keys WITHOUT the random barcode — into this rank's table.  Only meaningful for
// schemes without a random barcode (with one, de-duplication is routed per record, see bc_decode_route).
int bc_import_rows(bc_ctx* ctx, const uint64_t* dev_key_lo, const uint64_t* dev_key_hi, const uint64_t* dev_count, uint64_t n_rows) {
    if (!ctx) return BC_EINVAL;
    if (ctx->cfg.has_umi) return fail(ctx, BC_ESTATE, "bc_import_rows: scheme has a random barcode; exchange records instead (bc_exchange_*)");
    if (n_rows == 0) return BC_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->deferred) {  // kept aside; the next flush sums them with this rank's own (key, count) items
        const bool wide = ctx->tables.map.wide != 0;
        if (wide && !dev_key_hi) return fail(ctx, BC_EINVAL, "bc_import_rows: keys wider than 63 bits need dev_key_hi");
        int rc = reserve_items(ctx, ctx->imp, ctx->imp_n + n_rows, wide, true, ctx->imp_n);
        if (rc != BC_OK) return rc;
        CK(ctx, cudaMemcpyAsync(ctx->imp.lo + ctx->imp_n, dev_key_lo, n_rows * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (wide) CK(ctx, cudaMemcpyAsync(ctx->imp.hi + ctx->imp_n, dev_key_hi, n_rows * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(ctx, cudaMemcpyAsync(ctx->imp.w + ctx->imp_n, dev_count, n_rows * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->imp_n += n_rows;
        ctx->imported_rows += n_rows;
        ctx->rows_valid = false;
        return BC_OK;
    }
    int rc = ensure_capacity(ctx, n_rows);
    if (rc != BC_OK) return rc;
    ctx->rows_valid = false;
    ctx->imported_rows += n_rows;
    ProfScope p(ctx, BC_K_INSERT);
    CK(ctx, launch_insert(ctx->tables, reinterpret_cast<const unsigned long long*>(dev_key_lo),
                          reinterpret_cast<const unsigned long long*>(dev_key_hi),
                          reinterpret_cast<const unsigned long long*>(dev_count), n_rows, ctx->stream));
    return BC_OK;
}

int bc_dense_counts(bc_ctx* ctx, uint64_t** dev_counts, uint64_t* n) {
    if (!ctx || !dev_counts || !n) return BC_EINVAL;
    *dev_counts = nullptr;
    *n = 0;
    if (ctx->deferred || ctx->tables.map.kind != 0 || ctx->tables.has_set) return BC_OK;  // not a dense, UMI-free table: merge rows instead
    *dev_counts = reinterpret_cast<uint64_t*>(ctx->tables.map.data);
    *n = ctx->tables.map.cap;
    ctx->imported_rows = ctx->tables.map.cap;  // the caller is about to add other ranks' counts in place
    ctx->rows_valid = false;
    ctx->marg_valid = false;
    return BC_OK;
}

int bc_add_counters(bc_ctx* ctx, const uint64_t add[BC_N_COUNTERS]) {
    if (!ctx || !add) return BC_EINVAL;
    uint64_t cur[BC_N_COUNTERS];
    int rc = bc_get_counters(ctx, cur);
    if (rc != BC_OK) return rc;
    unsigned long long h[BC_N_COUNTERS];
    for (int i = 0; i < BC_N_COUNTERS; i++) h[i] = cur[i] + add[i];
    CK(ctx, cudaMemcpy(ctx->d_counters, h, sizeof h, cudaMemcpyHostToDevice));
    return BC_OK;
}

int bc_reset(bc_ctx* ctx) {
    if (!ctx) return BC_EINVAL;
    int rc = bc_sync(ctx);
    if (rc != BC_OK) return rc;
    drop_rows(ctx);
    CK(ctx, cudaMemsetAsync(ctx->d_counters, 0, (BC_N_COUNTERS + 2) * sizeof(unsigned long long), ctx->stream));
    rc = clear_table(ctx, ctx->tables.map);
    if (rc == BC_OK && ctx->tables.has_set) rc = clear_table(ctx, ctx->tables.set);
    CK(ctx, cudaMemsetAsync(ctx->d_stripes, 0, (size_t)kCounterStripes * kCounterStride * sizeof(unsigned long long), ctx->stream));
    CK(ctx, cudaMemsetAsync(ctx->d_flush, 0, sizeof(FlushStats), ctx->stream));
    ctx->stripes_dirty = false;
    ctx->rec_n = 0;
    ctx->imp_n = 0;
    ctx->dup_applied = 0;
    ctx->marg_valid = false;
    // a streamed job dropped half way has left records in its owners' buffers: the exchange must be re-opened
    if (ctx->x_mode == 1 && ctx->x_state != 3) ctx->x_dirty = true;
    ctx->x_mode = 0;
    ctx->x_state = 0;
    ctx->entries_upper = 0;
    ctx->imported_rows = 0;
    return rc;  // asynchronous: later work on the ctx stream is ordered after the clears
}

// ---------------------------------------------------------------------------------------------- measurement

int bc_set_profiling(bc_ctx* ctx, int on) {
    if (!ctx) return BC_EINVAL;
    ctx->profiling = on != 0;
    return BC_OK;
}

int bc_get_profile(bc_ctx* ctx, bc_profile* out) {
    if (!ctx || !out) return BC_EINVAL;
    int rc = fold_counters(ctx);
    if (rc == BC_OK) rc = bc_sync(ctx);
    if (rc != BC_OK) return rc;
    drain_profile(ctx);
    unsigned long long n = 0;
    CK(ctx, cudaMemcpy(&n, ctx->d_counters + BC_N_COUNTERS, sizeof n, cudaMemcpyDeviceToHost));
    ctx->prof.table_capacity = ctx->tables.map.cap;
    ctx->prof.table_entries = n;
    ctx->prof.key_bits = ctx->cfg.key_bits;
    ctx->prof.wide_keys = ctx->cfg.wide;
    ctx->prof.dense_table = ctx->tables.map.kind == 0;
    ctx->prof.deferred_count = ctx->deferred ? 1u : 0u;
    ctx->prof.flushed_global = ctx->flushed_global ? 1u : 0u;
    ctx->prof.flush_stages = ctx->flush_stages;
    ctx->prof.specialized_launches = ctx->jit_launches;
    ctx->prof.generic_launches = ctx->generic_launches;
    *out = ctx->prof;
    return BC_OK;
}

int bc_reset_profile(bc_ctx* ctx) {
    if (!ctx) return BC_EINVAL;
    int rc = bc_sync(ctx);
    if (rc != BC_OK) return rc;
    drain_profile(ctx);
    ctx->prof = bc_profile{};
    ctx->jit_launches = ctx->generic_launches = 0;
    return BC_OK;
}

}  // extern "C"
