// bc_kernels.h — host-callable launchers of the sm_100a kernels (defined in bc_kernels.cu).
#pragma once
#include "bc_device.cuh"

namespace bc {

constexpr int kMaxRanks = 8;  // GPUs of one box (NVLink / NVSwitch peers)

// multi-GPU exchange (launch_owner_scatter): the record buffers of the owner ranks, local or mapped over NVLink
struct PeerOut {
    unsigned long long* lo[kMaxRanks];
    unsigned long long* hi[kMaxRanks];  // nullptr for keys of at most 63 bits
    // streamed exchange (a batch's records leave right after its decode): a tile reserves its run in owner b's buffer with
    // one system-scope atomic on the OWNER's receive cursor (in the owner's memory, reached over NVLink like the buffer);
    // nullptr: the positions come from the local cursors handed to the launch (the bulk exchange after the last batch)
    unsigned long long* cursor[kMaxRanks];
    unsigned long long cap;        // records an owner's buffer holds: a run that would pass it is not written...
    unsigned long long* sent;      // ...but still counted here (local, one counter per owner)
    unsigned int* overflow;        // and this local flag is raised: the host then redoes the exchange in bulk, buffers re-opened larger
};

cudaError_t launch_fold_counters(unsigned long long* stripes, unsigned long long* counters, cudaStream_t stream);
cudaError_t launch_decode(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const Tables& tables,
                          unsigned long long* counters, const DecodeOut& out, const RecOut& rec, const Deferred& deferred, int flags,
                          cudaStream_t stream);
cudaError_t launch_decode_jit(const void* kernel, const BatchView& batch, const DevAux& aux, const Tables& tables, unsigned long long* counters,
                              const DecodeOut& out, const RecOut& rec, const Deferred& deferred, int flags, cudaStream_t stream);
cudaError_t launch_resolve(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const Tables& tables,
                           unsigned long long* counters, const DecodeOut& out, const RecOut& rec, const Deferred& deferred, int flags,
                           cudaStream_t stream);

// fills MODE_TABLE lookups with the exact correction result for every N-free barcode value
cudaError_t launch_build_table(const DevSlot& slot, const DevAux& aux, uint32_t* table, cudaStream_t stream);

// n (key, count) rows added to the map
cudaError_t launch_insert(const Tables& tables, const unsigned long long* key_lo, const unsigned long long* key_hi,
                          const unsigned long long* counts, unsigned long long n, cudaStream_t stream);

// occupied entries of `t` (dense: non-zero counts) appended to the row arrays; *n_rows is a device counter
cudaError_t launch_compact(const DevTable& t, unsigned long long* key_lo, unsigned long long* key_hi,
                           unsigned long long* count, unsigned long long* n_rows, cudaStream_t stream);

// dst[key & mask] += count over the rows
cudaError_t launch_marginal(const unsigned long long* key_lo, const unsigned long long* key_hi,
                            const unsigned long long* count, unsigned long long n_rows, Key mask, const DevTable& dst,
                            cudaStream_t stream);

cudaError_t launch_clear_map(const DevTable& t, cudaStream_t stream);

// ---- K4, enrichment marginals (info.rs:840-904) over index-coded keys: dense counter arrays, one pass over the rows ----
// single a   : counter [(sample << bits_a) | f_a]                       at s_off[a]
// double a<b : counter [(((sample << bits_b) | f_b) << bits_a) | f_a]   at d_off[pair]
constexpr int kMaxPairs = kMaxSlots * (kMaxSlots - 1) / 2;
struct MargPlan {
    uint32_t k;  // counted barcodes
    uint32_t f_shift[kMaxSlots], f_bits[kMaxSlots];  // their fields in a row key (random barcode already dropped)
    uint32_t s_shift, s_bits;                        // the sample field (0 bits: none)
    uint32_t n_pairs;                                // 0: singles only
    uint8_t pa[kMaxPairs], pb[kMaxPairs];
    unsigned long long s_off[kMaxSlots], d_off[kMaxPairs];
    unsigned long long n_single, n_total;  // counters in the singles / in all arrays
};
size_t marginal_smem_limit();
cudaError_t launch_marginals_dense(const unsigned long long* key_lo, const unsigned long long* key_hi, const unsigned long long* count,
                                   unsigned long long n_rows, const MargPlan& plan, unsigned long long* dense, cudaStream_t stream);
// non-zero counters of marginals [m_first, m_first + m_count) (singles first, then pairs) -> (key, count, mask) rows;
// key_lo == nullptr only counts them into *n_out
cudaError_t launch_marginals_rows(const MargPlan& plan, uint32_t m_first, uint32_t m_count, const unsigned long long* dense,
                                  unsigned long long* key_lo, unsigned long long* key_hi, unsigned long long* count, uint32_t* mask,
                                  unsigned long long* n_out, cudaStream_t stream);
// dst[i] += src[i] (src may be another GPU's memory)
cudaError_t launch_add_u64(unsigned long long* dst, const unsigned long long* src, unsigned long long n, cudaStream_t stream);

// A host batch in its transfer form, staged on the device (bc_wire_batch of include/bc_b200.h) -> bc_batch layout
struct WireDict {
    uint8_t c[16];
};
struct WireView {
    const uint32_t* lohi;     // n_reads x 2W words
    const uint16_t* read_len;
    const uint32_t* nmask;    // n_reads x W words, or nullptr: the N calls come as a list
    const uint32_t* n_read;
    const uint16_t* n_pos;
    const uint8_t* qual;      // packed codes, qual_stride bytes per read (a multiple of 4)
    uint32_t n_reads, W, n_calls, qual_bits, qual_stride, n_codes;
    WireDict dict;
};
cudaError_t launch_wire_expand(const WireView& w, uint32_t* planes, uint16_t* read_len, uint8_t* qual, uint32_t plane_stride,
                               uint32_t qual_stride, cudaStream_t stream);

// move every entry of `src` (hash kinds) into `dst`
cudaError_t launch_rehash(const DevTable& src, const DevTable& dst, cudaStream_t stream);

size_t decode_smem_bytes(const BatchView& batch);

// ---- deferred partitioned counting (bc_partition.cu) ---------------------------------------------------------------
// Items are structure-of-arrays (lo, hi, w): hi == nullptr for keys of at most 63 bits, w == nullptr for weight 1.
// An item whose key is kEmpty (narrow: lo, wide: hi) is a hole and is skipped everywhere.
struct ItemView {
    unsigned long long* lo;
    unsigned long long* hi;
    unsigned long long* w;
};
struct FlushStats {  // device-side results of one flush (u64 each)
    unsigned long long n_out;     // items written by the reduce kernel in flight (weighted keys, or final rows)
    unsigned long long overflow;  // != 0: a partition held more distinct keys than the shared-memory table -> global path
    unsigned long long valid;     // records that were not holes
    unsigned long long unique;    // distinct (key, random barcode) pairs
    unsigned long long max_bin;   // largest partition of the last histogram that asked for it
    unsigned long long big_items; // items in partitions larger than the limit given to that scan
};
enum ReduceMode { RED_DEDUPE = 0, RED_COUNT = 1, RED_DEDUPE_KEYED = 2 };
uint32_t reduce_fill(bool wide);   // target items per hashed partition
uint32_t reduce_chunk(bool wide);  // items per fixed chunk (pre-aggregation of schemes without a random barcode)

// per segment s of bins_per_seg bins: starts = seg_base[s] (0 when nullptr) + exclusive prefix of the segment's histogram,
// cursor = copy of starts; starts[n_seg * bins_per_seg] = grand total
cudaError_t launch_seg_scan(const uint32_t* hist, uint32_t n_seg, uint32_t bins_per_seg, const uint32_t* seg_base, uint32_t* starts,
                            uint32_t* cursor, FlushStats* stats /* nullable: records the largest bin and the items in bins > limit */,
                            uint32_t limit, cudaStream_t stream);
// partitions larger than `limit` copied to `out` (stats->n_out counts the items moved)
cudaError_t launch_gather_big(const ItemView& in, const uint32_t* starts, unsigned long long n_parts, uint32_t limit, const ItemView& out,
                              FlushStats* stats, uint32_t* scratch /* 4 x n_parts u32, 16-byte aligned */, uint32_t* scratch_n, cudaStream_t stream);
uint32_t reduce_capacity(bool wide);  // items a partition may hold for the staged (single pass) reduce
// One radix level of the hash partitioning: partition p of an item = mulhi(hash(key), P); the level's bin is
// (p >> shift) & mask, F bins.  seg_starts == nullptr: one segment [0, n_total); else n_seg segments, each split on its
// own (bins of segment s at bins + s * F).  scatter = false adds to the histogram `bins`; scatter = true moves the
// items to out, `bins` being the cursors (initialised with the exclusive prefix of the histogram).
struct SplitLevel {
    unsigned long long P;
    uint32_t shift, mask, F;
    uint32_t drop_bits;  // hash the key shifted right by this many bits (the random barcode) instead of the whole item
    unsigned long long salt;  // xor-ed into the key before hashing: an independent partitioning of the same keys
};
cudaError_t launch_split(bool scatter, bool wide, const ItemView& in, const ItemView& out, const uint32_t* seg_starts, uint32_t n_seg,
                         unsigned long long n_total, const SplitLevel& lv, uint32_t* bins, FlushStats* stats, bool count_valid,
                         cudaStream_t stream);
uint32_t split_max_bits();
// One CTA per partition [starts[p], starts[p+1]) — or, with starts == nullptr, per fixed chunk of `chunk` items.
//   RED_DEDUPE: distinct records of the partition, then their keys (record >> umi_bits) combined: out = (key, pairs)
//   RED_COUNT : out = (key, sum of weights)
// hot: RED_DEDUPE_KEYED appends the partitions it gives up on (a key with very many random barcodes) to hot.list / *hot.n
// (device memory); with hot.consume the launch instead processes exactly the listed partitions.
struct HotList {
    uint32_t* list = nullptr;
    uint32_t* n = nullptr;
    bool consume = false;
};
cudaError_t launch_reduce(int mode, bool wide, const ItemView& in, const uint32_t* starts, unsigned long long n_items,
                          unsigned long long n_ranges, uint32_t chunk, uint32_t umi_bits, const ItemView& out,
                          unsigned long long out_cap, FlushStats* stats, uint32_t skip_over /* > 0: leave larger partitions alone */,
                          const HotList& hot, cudaStream_t stream);
// multi-GPU exchange: valid items of `in` -> peers.lo/hi[owner] at cursors[owner]++ (see bc_partition.cu)
cudaError_t launch_owner_scatter(bool wide, const ItemView& in, const PeerOut& peers, unsigned long long n_total, const SplitLevel& lv,
                                 uint32_t* cursors, cudaStream_t stream);
// global-table path over the record buffer (oversized partitions, forced by BC_FLUSH_GLOBAL): counts like k_insert
cudaError_t launch_insert_items(const Tables& tables, const ItemView& in, bool wide, unsigned long long n, FlushStats* stats,
                                cudaStream_t stream);

}  // namespace bc
