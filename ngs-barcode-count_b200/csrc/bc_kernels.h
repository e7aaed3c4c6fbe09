// bc_kernels.h — host-callable launchers of the sm_100a kernels (defined in bc_kernels.cu).
#pragma once
#include "bc_device.cuh"

namespace bc {

struct DecodeOut {  // all optional (test hooks / decode_only)
    uint8_t* status;
    int16_t* offset;
    uint8_t* repaired;
    int32_t* slot_index;
    unsigned long long* key_lo;
    unsigned long long* key_hi;
};

constexpr int kMaxRanks = 8;  // one box

// multi-GPU: matched (key, UMI) records are written to their owner rank's bucket instead of being counted.  dst[r] is a
// local bucket (bc_decode_route) or rank r's receive region mapped over NVLink (bc_route_submit): plain peer stores
// from inside the decode kernel, cursors stay local — the transfer overlaps the decode, no separate copy step.
struct RouteOut {
    Key* dst[kMaxRanks];
    unsigned long long capacity;  // records per destination
    uint32_t* counts;             // local cursors: counts[r * count_stride]
    uint32_t count_stride;        // cursors of different ranks on different cache lines (same-sector atomics serialise)
    uint32_t n_ranks;
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
    uint32_t* count;
};

enum DecodeFlags { F_INSERT = 1, F_EMIT = 2, F_ROUTE = 4, F_LOCATE_ONLY = 8 };

// counters: BC_N_COUNTERS u64 on the device; n_new: entries newly claimed in `table`
cudaError_t launch_decode(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const Tables& tables,
                          unsigned long long* counters, const DecodeOut& out, const RouteOut& route, const Deferred& deferred,
                          int flags, cudaStream_t stream);
cudaError_t launch_resolve(const DevCfg& cfg, const BatchView& batch, const DevAux& aux, const Tables& tables,
                           unsigned long long* counters, const DecodeOut& out, const RouteOut& route, const Deferred& deferred,
                           int flags, cudaStream_t stream);

// fills MODE_TABLE lookups with the exact correction result for every N-free barcode value
cudaError_t launch_build_table(const DevSlot& slot, const DevAux& aux, uint32_t* table, cudaStream_t stream);

// records != nullptr: n reads (key incl. random barcode) counted like local ones, bumping matched/duplicates;
// otherwise n (key, count) rows added to the map
cudaError_t launch_insert(const Tables& tables, const unsigned long long* key_lo, const unsigned long long* key_hi,
                          const Key* records, const unsigned long long* counts, unsigned long long n,
                          unsigned long long* counters, cudaStream_t stream);

// occupied entries of `t` (dense: non-zero counts) appended to the row arrays; *n_rows is a device counter
cudaError_t launch_compact(const DevTable& t, unsigned long long* key_lo, unsigned long long* key_hi,
                           unsigned long long* count, unsigned long long* n_rows, cudaStream_t stream);

// dst[key & mask] += count over the rows
cudaError_t launch_marginal(const unsigned long long* key_lo, const unsigned long long* key_hi,
                            const unsigned long long* count, unsigned long long n_rows, Key mask, const DevTable& dst,
                            cudaStream_t stream);

cudaError_t launch_clear_map(const DevTable& t, cudaStream_t stream);

// copies the local buckets (local.counts[r] records each) to remote.dst[r] with coalesced 16-byte lanes
cudaError_t launch_push(const RouteOut& local, const RouteOut& remote, uint32_t* compact_counts, cudaStream_t stream);

// routed records received from every rank: segment s holds counts[s * count_stride] records at records + s * capacity
cudaError_t launch_insert_segments(const Tables& tables, const Key* records, unsigned long long capacity, const uint32_t* counts,
                                   uint32_t count_stride, uint32_t n_segments, unsigned long long* counters, cudaStream_t stream);

// move every entry of `src` (hash kinds) into `dst`
cudaError_t launch_rehash(const DevTable& src, const DevTable& dst, cudaStream_t stream);

size_t decode_smem_bytes(const BatchView& batch);

}  // namespace bc
