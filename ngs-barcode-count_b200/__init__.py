"""ctypes binding of the B200 decode-and-count library (include/bc_b200.h, include/bc_host.h).

The product is the C++/CUDA library in csrc/ (the reference is compiled Rust, so the host side is compiled code
too); this module is the thin binding the tests, bench.py and __graft_entry__ drive it through.  There is no CPU
path here: `lib()` raises when the shared library is missing and `Counter` raises when no CUDA device is usable.

The directory name holds a hyphen, so import it through the root-level shim: `import ngs_barcode_count_b200`.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB_PATH = os.path.join(PKG, "lib", "libbc_b200.so")
CLI_PATH = os.path.join(PKG, "bin", "barcode-count")

BC_ABI_VERSION = 2
BC_MAX_SLOTS = 16
BC_N_COUNTERS = 7
BC_N_KERNELS = 7
BC_LOC_HOST, BC_LOC_DEVICE = 0, 1
BC_READ_UNSUPPORTED = 0x8000
STATUS_NAMES = ["matched", "duplicate", "constant_region", "low_quality", "sample_barcode", "barcode", "unsupported"]
COUNTER_NAMES = ["matched", "constant_region", "sample_barcode", "barcode", "duplicates", "low_quality", "unsupported"]
KERNEL_NAMES = ["decode", "scan", "insert", "finish", "other", "enrich", "exchange"]
BC_CFG_INLINE_COUNT, BC_CFG_SPECIALIZE, BC_CFG_NO_SPECIALIZE = 1, 2, 4
BC_ADD_DENSE_COUNTS, BC_ADD_MARGINALS = 0, 1


class BcError(RuntimeError):
    pass


class bc_slot(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("offset", C.c_uint16), ("len", C.c_uint16), ("max_err", C.c_uint16),
                ("n_ref", C.c_uint32), ("ref_seqs", C.POINTER(C.c_char_p))]


class bc_config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("template_chars", C.c_char_p), ("template_len", C.c_uint32),
                ("region_codes", C.c_char_p), ("region_len", C.c_uint32), ("n_slots", C.c_uint32),
                ("slots", bc_slot * BC_MAX_SLOTS), ("max_const_err", C.c_uint16), ("min_quality", C.c_float),
                ("max_read_len", C.c_uint32), ("flags", C.c_uint32)]


class bc_batch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("plane_stride", C.c_uint32), ("qual_stride", C.c_uint32),
                ("location", C.c_int32), ("planes", C.c_void_p), ("read_len", C.c_void_p), ("qual", C.c_void_p)]


class bc_wire_batch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("max_read_len", C.c_uint32), ("qual_bits", C.c_uint32), ("qual_stride", C.c_uint32),
                ("n_calls", C.c_uint32), ("qual_dict", C.c_uint8 * 16), ("lohi", C.c_void_p), ("read_len", C.c_void_p),
                ("nmask", C.c_void_p), ("n_read", C.c_void_p), ("n_pos", C.c_void_p), ("qual", C.c_void_p)]


class bc_locate_out(C.Structure):
    _fields_ = [("status", C.c_void_p), ("offset", C.c_void_p), ("repaired", C.c_void_p)]


class bc_decode_out(C.Structure):
    _fields_ = [("status", C.c_void_p), ("offset", C.c_void_p), ("repaired", C.c_void_p), ("slot_index", C.c_void_p),
                ("key_lo", C.c_void_p), ("key_hi", C.c_void_p)]


class bc_table(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("key_lo", C.POINTER(C.c_uint64)), ("key_hi", C.POINTER(C.c_uint64)),
                ("count", C.POINTER(C.c_uint64)), ("mask", C.POINTER(C.c_uint32)), ("flags", C.c_uint32)]


class bc_profile(C.Structure):
    _fields_ = [("launches", C.c_uint64 * BC_N_KERNELS), ("ms", C.c_double * BC_N_KERNELS), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("table_capacity", C.c_uint64), ("table_entries", C.c_uint64),
                ("key_bits", C.c_uint32), ("wide_keys", C.c_uint32), ("dense_table", C.c_uint32),
                ("deferred_count", C.c_uint32), ("flushed_global", C.c_uint32), ("specialized_launches", C.c_uint64),
                ("generic_launches", C.c_uint64), ("flush_stages", C.c_uint32)]


def scan_fastq(path, threads=0):
    """(records, bases, crc32 of all sequence + quality bytes) of a FASTQ file read by the library's block reader (host only)."""
    n, b, c = C.c_uint64(), C.c_uint64(), C.c_uint32()
    err = C.create_string_buffer(512)
    rc = lib().bch_scan_fastq(os.fsencode(path), threads, C.byref(n), C.byref(b), C.byref(c), err, 512)
    if rc != 0:
        raise BcError("bch_scan_fastq: " + err.value.decode())
    return int(n.value), int(b.value), int(c.value)


def split_fastq(path, threads=0, block_bytes=0, min_slice=0):
    """Like scan_fastq, through the mapped-file splitter of bch_count_fastq's plain-file path (host only)."""
    n, b, c = C.c_uint64(), C.c_uint64(), C.c_uint32()
    err = C.create_string_buffer(512)
    rc = lib().bch_split_fastq(os.fsencode(path), threads, block_bytes, min_slice, C.byref(n), C.byref(b), C.byref(c), err, 512)
    if rc != 0:
        raise BcError("bch_split_fastq: " + err.value.decode())
    return int(n.value), int(b.value), int(c.value)


def walk_fastq(path, threads=0, chunk_bytes=0, batch_rows=1 << 16):
    """(records, bases, sum of per-record crc32, batches) through the one-pass walker of bch_count_fastq's plain-file path (host only)."""
    n, b, d, k = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    err = C.create_string_buffer(512)
    rc = lib().bch_walk_fastq(os.fsencode(path), threads, chunk_bytes, batch_rows, C.byref(n), C.byref(b), C.byref(d), C.byref(k), err, 512)
    if rc != 0:
        raise BcError("bch_walk_fastq: " + err.value.decode())
    return int(n.value), int(b.value), int(d.value), int(k.value)


def count_fastq_multi(run, counters, path, threads=0, batch_reads=1 << 20):
    """bch_count_fastq_multi over several Counters (one per GPU, or several on one GPU in tests) -> reads"""
    arr = (C.c_void_p * len(counters))(*[c.h for c in counters])
    total = C.c_uint64(0)
    err = C.create_string_buffer(2048)
    rc = lib().bch_count_fastq_multi(run.h, arr, len(counters), _b(path), threads, batch_reads, C.byref(total), err, 2048)
    if rc != 0:
        raise BcError("bch_count_fastq_multi: " + err.value.decode())
    return total.value


def counters_multi(counters):
    arr = (C.c_void_p * len(counters))(*[c.h for c in counters])
    out = (C.c_uint64 * BC_N_COUNTERS)()
    if lib().bch_counters_multi(arr, len(counters), out) != 0:
        raise BcError("bch_counters_multi: " + lib().bc_last_error(counters[0].h).decode())
    return dict(zip(COUNTER_NAMES, [int(x) for x in out]))


def write_counts_multi(run, counters, outdir, prefix, merge=False, enrich=False):
    arr = (C.c_void_p * len(counters))(*[c.h for c in counters])
    names = C.create_string_buffer(1 << 20)
    err = C.create_string_buffer(2048)
    n = lib().bch_write_counts_multi(run.h, arr, len(counters), _b(outdir), _b(prefix), int(merge), int(enrich), names, 1 << 20, err, 2048)
    if n < 0:
        raise BcError("bch_write_counts_multi: " + err.value.decode())
    return [x.split("\t")[0] for x in names.value.decode().split("\n") if x]


class bch_args(C.Structure):
    _fields_ = [("format_path", C.c_char_p), ("sample_barcodes_path", C.c_char_p), ("counted_barcodes_path", C.c_char_p),
                ("max_errors_counted_barcode", C.c_int), ("max_errors_sample", C.c_int), ("max_errors_constant", C.c_int),
                ("min_quality", C.c_float), ("max_read_len", C.c_uint32)]


# every entry point include/bc_b200.h and include/bc_host.h declare (tests/test_abi.py checks the list against the headers)
_PROTOS = {
    "bc_plane_words": (C.c_uint32, [C.c_uint32]),
    "bc_plane_stride": (C.c_uint32, [C.c_uint32]),
    "bc_qual_stride": (C.c_uint32, [C.c_uint32]),
    "bc_create": (C.c_int, [C.POINTER(bc_config), C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]),
    "bc_destroy": (None, [C.c_void_p]),
    "bc_device_of": (C.c_int, [C.c_void_p]),
    "bc_specialization_note": (C.c_char_p, [C.c_void_p]),
    "bc_jit_check": (C.c_int, [C.POINTER(bc_config), C.c_char_p, C.c_int]),
    "bc_last_error": (C.c_char_p, [C.c_void_p]),
    "bc_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bc_submit": (C.c_int, [C.c_void_p, C.POINTER(bc_batch)]),
    "bc_wire_qual_codes": (C.c_uint32, [C.c_uint32]),
    "bc_wire_qual_stride": (C.c_uint32, [C.c_uint32, C.c_uint32]),
    "bc_submit_wire": (C.c_int, [C.c_void_p, C.POINTER(bc_wire_batch)]),
    "bc_sync": (C.c_int, [C.c_void_p]),
    "bc_wait_copies": (C.c_int, [C.c_void_p]),
    "bc_wait_older_copies": (C.c_int, [C.c_void_p]),
    "bc_get_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "bc_locate_only": (C.c_int, [C.c_void_p, C.POINTER(bc_batch), C.POINTER(bc_locate_out)]),
    "bc_decode_only": (C.c_int, [C.c_void_p, C.POINTER(bc_batch), C.POINTER(bc_decode_out)]),
    "bc_table_free": (None, [C.POINTER(bc_table)]),
    "bc_finish": (C.c_int, [C.c_void_p, C.POINTER(bc_table)]),
    "bc_enrich": (C.c_int, [C.c_void_p, C.POINTER(bc_table), C.POINTER(bc_table)]),
    "bc_key_decode": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_int32), C.c_char_p,
                                C.c_uint32]),
    "bc_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "bc_marginals": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "bc_exchange_open": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64]),
    "bc_exchange_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bc_exchange_capacity": (C.c_uint64, [C.c_void_p]),
    "bc_exchange_disconnect": (C.c_int, [C.c_void_p]),
    "bc_exchange_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bc_exchange_connect_local": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "bc_exchange_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "bc_exchange_scatter": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "bc_exchange_finish": (C.c_int, [C.c_void_p, C.c_uint64]),
    "bc_peer_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "bc_export_rows": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_uint64)]),
    "bc_import_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "bc_dense_counts": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "bc_add_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "bc_reset": (C.c_int, [C.c_void_p]),
    "bc_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "bc_get_profile": (C.c_int, [C.c_void_p, C.POINTER(bc_profile)]),
    "bc_reset_profile": (C.c_int, [C.c_void_p]),
    "bch_open": (C.c_void_p, [C.POINTER(bch_args), C.c_char_p, C.c_int]),
    "bch_close": (None, [C.c_void_p]),
    "bch_set_progress": (None, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bch_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_longlong]),
    "bch_config": (C.POINTER(bc_config), [C.c_void_p]),
    "bch_describe": (C.c_char_p, [C.c_void_p]),
    "bch_barcode_num": (C.c_uint32, [C.c_void_p]),
    "bch_ref_dna": (C.c_char_p, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "bch_ref_name": (C.c_char_p, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "bch_pack": (C.c_int, [C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_uint]),
    "bch_pack_lines": (C.c_int, [C.c_uint32, C.c_uint32, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]),
    "bch_set_simd_level": (C.c_int, [C.c_int]),
    "bch_wire_bound": (C.c_size_t, [C.c_uint32, C.c_uint32, C.c_int]),
    "bch_wire_from_batch": (C.c_int, [C.POINTER(bc_batch), C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(bc_wire_batch)]),
    "bch_ingest_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "bch_scan_fastq": (C.c_int, [C.c_char_p, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_char_p,
                                 C.c_int]),
    "bch_split_fastq": (C.c_int, [C.c_char_p, C.c_uint, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                  C.c_char_p, C.c_int]),
    "bch_walk_fastq": (C.c_int, [C.c_char_p, C.c_uint, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_char_p, C.c_int]),
    "bch_count_fastq_multi": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.c_uint, C.c_uint32,
                                        C.POINTER(C.c_uint64), C.c_char_p, C.c_int]),
    "bch_counters_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_uint64)]),
    "bch_write_counts_multi": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                         C.c_char_p, C.c_int, C.c_char_p, C.c_int]),
    "bch_count_fastq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint, C.c_uint32, C.POINTER(C.c_uint64), C.c_char_p,
                                  C.c_int]),
    "bch_write_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                   C.c_char_p, C.c_int]),
}

_lib = None


def lib():
    """The shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BcError(f"{LIB_PATH} is missing: run `python ngs-barcode-count_b200/build.py` (there is no CPU path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _b(s):
    return None if s is None else os.fsencode(s)


def _ptr(a):
    return None if a is None else a.ctypes.data


class Run:
    """Run set-up: the scheme file, the two conversion CSVs and the error caps (bch_open).
    Mirrors SequenceFormat::parse_format_file + BarcodeConversions + MaxSeqErrors::new (info.rs:215-543)."""

    def __init__(self, fmt, samples=None, counted=None, min_quality=0.0, max_barcode=None, max_sample=None,
                 max_constant=None, max_read_len=0):
        neg = lambda v: -1 if v is None else int(v)
        self._keep = [_b(fmt), _b(samples), _b(counted)]
        a = bch_args(self._keep[0], self._keep[1], self._keep[2], neg(max_barcode), neg(max_sample), neg(max_constant),
                     float(min_quality), int(max_read_len))
        err = C.create_string_buffer(2048)
        self.h = lib().bch_open(C.byref(a), err, 2048)
        if not self.h:
            raise BcError(err.value.decode())
        self.cfg = lib().bch_config(self.h).contents
        self.max_read_len = self.cfg.max_read_len
        self.quality_on = self.cfg.min_quality > 0
        self.plane_stride = lib().bc_plane_stride(self.max_read_len)
        self.qual_stride = lib().bc_qual_stride(self.max_read_len)

    def close(self):
        if getattr(self, "h", None):
            lib().bch_close(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def n_slots(self):
        return self.cfg.n_slots

    def slot(self, i):
        return self.cfg.slots[i]

    def describe(self):
        return lib().bch_describe(self.h).decode()

    def set_option(self, name, value):
        if lib().bch_set_option(self.h, name.encode(), int(value)) != 0:
            raise BcError(f"bch_set_option: unknown option {name!r}")

    def jit_check(self):
        """Compiles the run's specialised decode kernel with NVRTC (no GPU needed) -> (ok, log)"""
        log = C.create_string_buffer(8192)
        rc = lib().bc_jit_check(C.byref(self.cfg), log, 8192)
        return rc == 0, log.value.decode()

    def ref_dna(self, slot, i):
        v = lib().bch_ref_dna(self.h, slot, i)
        return None if v is None else v.decode()

    def ref_name(self, slot, i):
        v = lib().bch_ref_name(self.h, slot, i)
        return None if v is None else v.decode()

    def pack(self, seqs, quals=None, threads=1):
        """Text reads -> a host Batch (bit planes, lengths and, when the quality filter is on, Phred bytes)."""
        n = len(seqs)
        planes = np.zeros((max(n, 1), self.plane_stride), dtype=np.uint32)
        read_len = np.zeros(max(n, 1), dtype=np.uint16)
        want_q = self.quality_on
        qual = np.zeros((max(n, 1), self.qual_stride), dtype=np.uint8) if want_q else None
        if n:
            if want_q and quals is None:
                raise BcError("min_quality > 0 needs the quality strings")
            rc = lib().bch_pack_lines(self.max_read_len, n, "\n".join(seqs).encode(),
                                      "\n".join(quals).encode() if want_q else None, _ptr(planes), _ptr(read_len),
                                      _ptr(qual), threads)
            if rc != 0:
                raise BcError("bch_pack_lines: a read is longer than max_read_len")
        return Batch(n, self.plane_stride, self.qual_stride, planes, read_len, qual)


class Batch:
    """A packed batch in host (numpy) or device (torch) memory; see bc_batch in include/bc_b200.h."""

    def __init__(self, n, plane_stride, qual_stride, planes, read_len, qual, device=False):
        self.n, self.plane_stride, self.qual_stride = n, plane_stride, qual_stride
        self.planes, self.read_len, self.qual, self.device = planes, read_len, qual, device

    def slice(self, a, b):
        return Batch(b - a, self.plane_stride, self.qual_stride, self.planes[a:b], self.read_len[a:b],
                     None if self.qual is None else self.qual[a:b], self.device)

    def c_struct(self):
        if self.device:
            p, r, q = self.planes.data_ptr(), self.read_len.data_ptr(), None if self.qual is None else self.qual.data_ptr()
        else:
            p, r, q = _ptr(self.planes), _ptr(self.read_len), _ptr(self.qual)
        return bc_batch(self.n, self.plane_stride, self.qual_stride, BC_LOC_DEVICE if self.device else BC_LOC_HOST, p, r, q)

    def to_device(self, device="cuda:0"):
        import torch
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a).view(
            np.int32 if a.dtype == np.uint32 else np.int16 if a.dtype == np.uint16 else np.uint8)).to(device)
        return Batch(self.n, self.plane_stride, self.qual_stride, t(self.planes), t(self.read_len), t(self.qual), True)


class WireBatch:
    """A host Batch in its transfer form (bc_wire_batch): lo / hi planes, N calls as a list (or a dense plane), quality
    characters as 8-, 6-, 4- or 2-bit codes, all inside one buffer — `buf` if given (a uint8 numpy array, e.g. the view
    of a pinned torch tensor, of at least WireBatch.bound(...) bytes), else a fresh numpy array."""

    @staticmethod
    def bound(n, max_read_len, with_qual):
        return int(lib().bch_wire_bound(n, max_read_len, int(with_qual)))

    def __init__(self, batch, max_read_len, qual_bits=0, buf=None):
        if batch.device:
            raise BcError("WireBatch: the transfer form is made from a host batch")
        need = self.bound(batch.n, max_read_len, batch.qual is not None)
        self.buf = np.zeros(max(need, 1), np.uint8) if buf is None else buf
        self.n, self.max_read_len = batch.n, max_read_len
        self.c = bc_wire_batch()
        b = batch.c_struct()
        rc = lib().bch_wire_from_batch(C.byref(b), max_read_len, qual_bits, _ptr(self.buf), self.buf.nbytes, C.byref(self.c))
        if rc != 0:
            raise BcError(f"bch_wire_from_batch failed ({rc}): the batch does not fit the requested form")
        w = lib().bc_plane_words(max_read_len)
        dense = bool(self.c.nmask)
        self.nbytes = (batch.n * (2 * w * 4 + 2) + (batch.n * w * 4 if dense else self.c.n_calls * 6) + batch.n * self.c.qual_stride)

    def c_struct(self):
        return self.c


class Counter:
    """One GPU's decode-and-count context (bc_ctx): stands where the reference has its SequenceParser worker pool,
    the shared Results and the SequenceErrors counters (parse.rs:28-76, info.rs:16-139, 661-808)."""

    def __init__(self, run, device=0, expected_reads=0, flags=0):
        self.run = run
        h = C.c_void_p()
        cfg = bc_config.from_buffer_copy(run.cfg)  # the run owns the strings the copy points to
        cfg.flags = flags
        rc = lib().bc_create(C.byref(cfg), device, expected_reads, C.byref(h))
        if rc != 0:
            raise BcError(f"bc_create failed ({rc}): " + lib().bc_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().bc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc, what):
        if rc < 0:
            raise BcError(f"{what} failed ({rc}): " + lib().bc_last_error(self.h).decode())
        return rc

    def set_stream(self, cuda_stream):
        self._ck(lib().bc_set_stream(self.h, C.c_void_p(cuda_stream)), "bc_set_stream")

    def set_option(self, name, value):
        self._ck(lib().bc_set_option(self.h, name.encode(), int(value)), "bc_set_option")

    def submit(self, batch):
        b = batch.c_struct()
        if isinstance(batch, WireBatch):
            self._ck(lib().bc_submit_wire(self.h, C.byref(b)), "bc_submit_wire")
        else:
            self._ck(lib().bc_submit(self.h, C.byref(b)), "bc_submit")

    def sync(self):
        self._ck(lib().bc_sync(self.h), "bc_sync")

    def wait_copies(self):
        self._ck(lib().bc_wait_copies(self.h), "bc_wait_copies")

    def reset(self):
        self._ck(lib().bc_reset(self.h), "bc_reset")

    def counters(self):
        out = (C.c_uint64 * BC_N_COUNTERS)()
        self._ck(lib().bc_get_counters(self.h, out), "bc_get_counters")
        return dict(zip(COUNTER_NAMES, [int(x) for x in out]))

    def add_counters(self, d):
        arr = (C.c_uint64 * BC_N_COUNTERS)(*[int(d.get(k, 0)) for k in COUNTER_NAMES])
        self._ck(lib().bc_add_counters(self.h, arr), "bc_add_counters")

    def locate_only(self, batch):
        n = batch.n
        st, off, rep = np.zeros(n, np.uint8), np.zeros(n, np.int16), np.zeros(n, np.uint8)
        out = bc_locate_out(_ptr(st), _ptr(off), _ptr(rep))
        b = batch.c_struct()
        self._ck(lib().bc_locate_only(self.h, C.byref(b), C.byref(out)), "bc_locate_only")
        return st, off, rep

    def decode_only(self, batch):
        n, ns = batch.n, self.run.n_slots
        st, off, rep = np.zeros(n, np.uint8), np.zeros(n, np.int16), np.zeros(n, np.uint8)
        idx = np.zeros((n, ns), np.int32)
        lo, hi = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        out = bc_decode_out(_ptr(st), _ptr(off), _ptr(rep), _ptr(idx), _ptr(lo), _ptr(hi))
        b = batch.c_struct()
        self._ck(lib().bc_decode_only(self.h, C.byref(b), C.byref(out)), "bc_decode_only")
        return dict(status=st, offset=off, repaired=rep, slot_index=idx, key_lo=lo, key_hi=hi)

    def key_decode(self, lo, hi, mask=0, with_umi=False):
        """-> per slot (scheme order): reference DNA for indexed slots, captured DNA for raw slots, None when absent."""
        ns = self.run.n_slots
        idx = (C.c_int32 * ns)()
        stride = 40
        buf = C.create_string_buffer(ns * stride)
        self._ck(lib().bc_key_decode(self.h, int(lo), int(hi), int(mask), int(with_umi), idx, buf, stride), "bc_key_decode")
        out = []
        for s in range(ns):
            if idx[s] >= 0:
                out.append(self.run.ref_dna(s, idx[s]))
            else:
                txt = buf.raw[s * stride:(s + 1) * stride].split(b"\0", 1)[0].decode()
                out.append(txt if txt else None)
        return out

    def _table(self, t):
        n = int(t.n_rows)
        get = lambda p, dt: np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy() if n and p else np.zeros(0, dt)
        rows = dict(key_lo=get(t.key_lo, np.uint64), key_hi=get(t.key_hi, np.uint64) if t.key_hi else np.zeros(n, np.uint64),
                    count=get(t.count, np.uint64),
                    mask=get(t.mask, np.uint32) if t.mask else None)
        lib().bc_table_free(C.byref(t))
        return rows

    def finish(self):
        t = bc_table()
        self._ck(lib().bc_finish(self.h, C.byref(t)), "bc_finish")
        return self._table(t)

    def finish_view(self):
        """bc_finish without copying: (n_rows, key_lo, key_hi or None, count) as numpy views of the ctx-owned pinned
        rows, valid until the next finish on this Counter."""
        t = bc_table()
        self._ck(lib().bc_finish(self.h, C.byref(t)), "bc_finish")
        n = int(t.n_rows)
        view = lambda p: np.ctypeslib.as_array(p, shape=(n,)) if n and p else np.zeros(0, np.uint64)
        return n, view(t.key_lo), (view(t.key_hi) if t.key_hi else None), view(t.count)

    def enrich(self, doubles=True):
        s, d = bc_table(), bc_table()
        self._ck(lib().bc_enrich(self.h, C.byref(s), C.byref(d) if doubles else None), "bc_enrich")
        return self._table(s), (self._table(d) if doubles else None)

    def count_fastq(self, path, threads=0, batch_reads=1 << 20):
        total = C.c_uint64(0)
        err = C.create_string_buffer(2048)
        rc = lib().bch_count_fastq(self.run.h, self.h, _b(path), threads, batch_reads, C.byref(total), err, 2048)
        if rc != 0:
            raise BcError("bch_count_fastq: " + err.value.decode())
        return total.value

    def ingest_stats(self):
        """phases of the last count_fastq on the ingest thread (seconds) and its batch counts"""
        sec, cnt = (C.c_double * 6)(), (C.c_uint64 * 3)()
        lib().bch_ingest_stats(self.run.h, sec, cnt)
        return dict(split_s=sec[0], pack_s=sec[1], submit_s=sec[2], wait_s=sec[3], total_s=sec[4], map_s=sec[5], batches=int(cnt[0]),
                    batches_qual8=int(cnt[1]), batches_dense_n=int(cnt[2]))

    def write_counts(self, outdir, prefix, merge=False, enrich=False):
        names = C.create_string_buffer(1 << 20)
        err = C.create_string_buffer(2048)
        n = lib().bch_write_counts(self.run.h, self.h, _b(outdir), _b(prefix), int(merge), int(enrich), names, 1 << 20, err, 2048)
        if n < 0:
            raise BcError("bch_write_counts: " + err.value.decode())
        return [x.split("\t")[0] for x in names.value.decode().split("\n") if x]

    # ---- multi-GPU building blocks ----
    def exchange_open(self, n_ranks, rank, capacity):
        self._ck(lib().bc_exchange_open(self.h, n_ranks, rank, int(capacity)), "bc_exchange_open")

    def exchange_disconnect(self):
        self._ck(lib().bc_exchange_disconnect(self.h), "bc_exchange_disconnect")

    def exchange_handle(self):
        """-> the CUDA IPC handle (bytes) of this rank's receive buffer"""
        h = C.create_string_buffer(64)
        self._ck(lib().bc_exchange_handle(self.h, h), "bc_exchange_handle")
        return h.raw

    def exchange_connect(self, handles):
        self._ck(lib().bc_exchange_connect(self.h, b"".join(handles)), "bc_exchange_connect")

    def exchange_connect_local(self, counters):
        arr = (C.c_void_p * len(counters))(*[c.h for c in counters])
        self._ck(lib().bc_exchange_connect_local(self.h, arr), "bc_exchange_connect_local")

    def exchange_count(self, n_ranks):
        out = (C.c_uint64 * n_ranks)()
        self._ck(lib().bc_exchange_count(self.h, out), "bc_exchange_count")
        return [int(x) for x in out]

    def exchange_scatter(self, first):
        arr = (C.c_uint64 * len(first))(*[int(x) for x in first])
        self._ck(lib().bc_exchange_scatter(self.h, arr), "bc_exchange_scatter")

    def exchange_finish(self, n_received):
        self._ck(lib().bc_exchange_finish(self.h, int(n_received)), "bc_exchange_finish")

    def peer_add(self, src, what):
        self._ck(lib().bc_peer_add(self.h, src.h, int(what)), "bc_peer_add")

    def marginals(self):
        """-> (device pointer, n) of the dense enrichment counters, or (None, 0) for schemes with raw barcodes"""
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(lib().bc_marginals(self.h, C.byref(p), C.byref(n)), "bc_marginals")
        return p.value, int(n.value)

    def export_rows(self):
        lo, hi, cnt, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._ck(lib().bc_export_rows(self.h, C.byref(lo), C.byref(hi), C.byref(cnt), C.byref(n)), "bc_export_rows")
        return lo.value, hi.value, cnt.value, int(n.value)

    def dense_counts(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(lib().bc_dense_counts(self.h, C.byref(p), C.byref(n)), "bc_dense_counts")
        return p.value, int(n.value)

    def import_rows(self, lo, hi, cnt, n):
        self._ck(lib().bc_import_rows(self.h, C.c_void_p(lo.data_ptr()), C.c_void_p(hi.data_ptr()) if hi is not None else None,
                                      C.c_void_p(cnt.data_ptr()), n), "bc_import_rows")

    # ---- measurement ----
    def set_profiling(self, on):
        self._ck(lib().bc_set_profiling(self.h, int(on)), "bc_set_profiling")

    def reset_profile(self):
        self._ck(lib().bc_reset_profile(self.h), "bc_reset_profile")

    def profile(self):
        p = bc_profile()
        self._ck(lib().bc_get_profile(self.h, C.byref(p)), "bc_get_profile")
        return dict(launches=dict(zip(KERNEL_NAMES, [int(x) for x in p.launches])),
                    ms=dict(zip(KERNEL_NAMES, [float(x) for x in p.ms])), h2d_bytes=int(p.h2d_bytes),
                    d2h_bytes=int(p.d2h_bytes), table_capacity=int(p.table_capacity), table_entries=int(p.table_entries),
                    key_bits=int(p.key_bits), wide_keys=int(p.wide_keys), dense_table=int(p.dense_table),
                    deferred_count=int(p.deferred_count), flushed_global=int(p.flushed_global), flush_stages=int(p.flush_stages),
                    specialized_launches=int(p.specialized_launches), generic_launches=int(p.generic_launches),
                    specialization=lib().bc_specialization_note(self.h).decode())
