"""Shared test plumbing: the oracle behind ctypes, golden-case loading, canonical CSV comparison."""
import ctypes as C
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

STATUS_NAMES = ["matched", "duplicate", "constant_region", "low_quality", "sample_barcode", "barcode"]
COUNTER_NAMES = ["matched", "constant_region", "sample_barcode", "barcode", "duplicates", "low_quality"]


class OrcOutcome(C.Structure):
    _fields_ = [("status", C.c_int), ("offset", C.c_long), ("repaired", C.c_int), ("has_random", C.c_int),
                ("sample", C.c_char * 512), ("barcodes", C.c_char * 4096), ("random", C.c_char * 512)]


_oracle_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        lib = C.CDLL(os.path.join(ROOT, "oracle", "build", "liboracle.so"))
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                   C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_process_read.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(OrcOutcome)]
        lib.orc_decode_read.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(OrcOutcome)]
        lib.orc_process_block.restype = C.c_long
        lib.orc_process_block.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_long]
        lib.orc_counters.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        lib.orc_write_files.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        lib.orc_run_fastq.restype = C.c_double
        lib.orc_run_fastq.argtypes = [C.c_void_p, C.c_char_p, C.c_uint, C.POINTER(C.c_ulonglong), C.c_char_p, C.c_int]
        lib.orc_fix_error.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_char_p, C.c_int]
        lib.orc_max_errors.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ushort), C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.orc_format_info.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _oracle_lib = lib
    return _oracle_lib


def _b(s):
    return None if s is None else os.fsencode(s)


class Oracle:
    """oracle::Pipeline through the C entry points of oracle/oracle_capi.cpp."""

    def __init__(self, fmt, samples=None, counted=None, min_quality=0.0, merge=False, enrich=False, outdir="./",
                 prefix="oracle", max_barcode=None, max_sample=None, max_constant=None):
        self.lib = oracle_lib()
        err = C.create_string_buffer(1024)
        neg = lambda v: -1 if v is None else int(v)
        self.h = self.lib.orc_create(_b(fmt), _b(samples), _b(counted), neg(max_barcode), neg(max_sample),
                                     neg(max_constant), float(min_quality), int(merge), int(enrich), _b(outdir),
                                     _b(prefix), err, 1024)
        if not self.h:
            raise RuntimeError("oracle: " + err.value.decode())

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _outcome(self, fn, seq, qual):
        o = OrcOutcome()
        rc = fn(self.h, seq.encode(), qual.encode(), C.byref(o))
        assert rc == 0
        return dict(status=STATUS_NAMES[o.status], offset=o.offset, repaired=bool(o.repaired),
                    sample=o.sample.decode(), barcodes=o.barcodes.decode(),
                    random=o.random.decode() if o.has_random else None)

    def process(self, seq, qual):
        return self._outcome(self.lib.orc_process_read, seq, qual)

    def decode(self, seq, qual):
        return self._outcome(self.lib.orc_decode_read, seq, qual)

    def process_block(self, seqs, quals):
        n = len(seqs)
        st = (C.c_int * n)()
        got = self.lib.orc_process_block(self.h, "\n".join(seqs).encode(), "\n".join(quals).encode(), st, n)
        assert got == n, (got, n)
        return list(st)

    def counters(self):
        out = (C.c_ulonglong * 6)()
        self.lib.orc_counters(self.h, out)
        return dict(zip(COUNTER_NAMES, out))

    def write_files(self):
        names = C.create_string_buffer(1 << 20)
        err = C.create_string_buffer(1024)
        n = self.lib.orc_write_files(self.h, names, 1 << 20, err, 1024)
        if n < 0:
            raise RuntimeError("oracle: " + err.value.decode())
        return [x for x in names.value.decode().split("\n") if x]

    def run_fastq(self, path, threads):
        total = C.c_ulonglong(0)
        err = C.create_string_buffer(1024)
        secs = self.lib.orc_run_fastq(self.h, _b(path), threads, C.byref(total), err, 1024)
        if secs < 0:
            raise RuntimeError("oracle: " + err.value.decode())
        return secs, total.value

    def format_info(self):
        fs, rs = C.create_string_buffer(8192), C.create_string_buffer(8192)
        cl, bn, mc, ms = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        mb = (C.c_int * 64)()
        self.lib.orc_format_info(self.h, fs, rs, 8192, C.byref(cl), C.byref(bn), C.byref(mc), C.byref(ms), mb)
        return dict(format_string=fs.value.decode(), regions_string=rs.value.decode(), constant_len=cl.value,
                    barcode_num=bn.value, max_constant=mc.value, max_sample=ms.value,
                    max_barcode=list(mb[:bn.value]))


def golden_cases():
    return sorted(d for d in os.listdir(GOLDEN) if os.path.exists(os.path.join(GOLDEN, d, "expected.json")))


def load_golden(name):
    d = os.path.join(GOLDEN, name)
    with open(os.path.join(d, "expected.json")) as f:
        exp = json.load(f)
    paths = dict(fmt=os.path.join(d, "scheme.txt"),
                 samples=os.path.join(d, "samples.csv") if os.path.exists(os.path.join(d, "samples.csv")) else None,
                 counted=os.path.join(d, "barcodes.csv") if os.path.exists(os.path.join(d, "barcodes.csv")) else None,
                 fastq=os.path.join(d, "reads.fastq"))
    return exp, paths


def read_fastq(path):
    with open(path) as f:
        lines = f.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return [(lines[i + 1], lines[i + 3]) for i in range(0, len(lines) - 3, 4)]


def canonical_csv(text):
    """Header verbatim + byte-sorted data rows (the reference's row order is ahash order)."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return [lines[0]] + sorted(lines[1:]) if lines else []


def canonical_merged(lines):
    """Merged files without a sample file have arbitrary column order in the reference: compare sample
    columns as a sorted multiset of (name, value) per row."""
    header = lines[0].split(",")
    n_bar = sum(1 for h in header if h.startswith("Barcode"))
    out = []
    for row in lines[1:]:
        f = row.split(",")
        out.append((tuple(f[:n_bar]), tuple(sorted(zip(header[n_bar:], f[n_bar:])))))
    return (tuple(header[:n_bar]), tuple(sorted(header[n_bar:]))), sorted(out)


def read_csv_dir(outdir, prefix):
    files = {}
    for fn in sorted(os.listdir(outdir)):
        if fn.startswith(prefix + "_") and fn.endswith(".csv"):
            with open(os.path.join(outdir, fn)) as f:
                files[fn] = canonical_csv(f.read())
    return files


def assert_same_csv_set(got, want):
    assert sorted(got) == sorted(want), (sorted(set(got) ^ set(want)))
    for fn in want:
        if ".all." in fn:
            assert canonical_merged(got[fn]) == canonical_merged(want[fn]), fn
        else:
            assert got[fn] == want[fn], fn


def bgzf_compress(data, chunk=20000, level=6):
    """bgzip-style gzip: independent members of at most 64 KB with their compressed size in a 'BC' extra subfield, ended
    by the empty end-of-file member (SAM specification, section 4.1)."""
    import struct
    import zlib
    out = bytearray()
    i = 0
    while True:
        block = data[i:i + chunk]
        i += chunk
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        cd = c.compress(block) + c.flush()
        out += b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(cd) + 25)
        out += cd + struct.pack("<II", zlib.crc32(block) & 0xFFFFFFFF, len(block))
        if not block:
            break
    return bytes(out)
