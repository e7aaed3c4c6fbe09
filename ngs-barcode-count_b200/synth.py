"""Synthetic workloads of BASELINE.json: scheme / conversion files on disk (what both the GPU
library and the CPU oracle consume) plus the generator configuration for csrc/bc_synth.cu, which produces read i of
a workload identically on the device (packed batches in HBM) and on the host (FASTQ text).

Bench / test tooling, not part of the drop-in boundary.
"""
import ctypes as C
import os
import re

import numpy as np

from . import PKG, BcError

SYNTH_LIB_PATH = os.path.join(PKG, "lib", "libbc_synth.so")
BCS_MAX_SLOTS, BCS_MAX_READ = 16, 256
BASE_SEED = 20261018


class bcs_slot(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("skew", C.c_uint8), ("offset", C.c_uint16), ("len", C.c_uint16), ("ref_len", C.c_uint16),
                ("n_ref", C.c_uint32), ("ref_off", C.c_uint32), ("pool", C.c_uint64)]


class bcs_config(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("read_len", C.c_uint32), ("template_len", C.c_uint32), ("n_slots", C.c_uint32),
                ("max_start", C.c_uint32), ("template_codes", C.c_uint8 * BCS_MAX_READ), ("slots", bcs_slot * BCS_MAX_SLOTS),
                ("p_junk", C.c_uint32), ("p_lowq", C.c_uint32), ("p_enriched", C.c_uint32), ("p_sub16", C.c_uint16),
                ("p_n16", C.c_uint16), ("n_enriched", C.c_uint32), ("molecule_pool", C.c_uint64), ("q_mean", C.c_uint8),
                ("q_spread", C.c_uint8)]


_lib = None


def synth_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise BcError(f"{SYNTH_LIB_PATH} is missing: run `python ngs-barcode-count_b200/build.py`")
        l = C.CDLL(SYNTH_LIB_PATH)
        l.bcs_generate_device.restype = C.c_int
        l.bcs_generate_device.argtypes = [C.POINTER(bcs_config), C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
        l.bcs_fastq_bytes.restype = C.c_size_t
        l.bcs_fastq_bytes.argtypes = [C.POINTER(bcs_config), C.c_uint64, C.c_uint64]
        l.bcs_generate_fastq.restype = C.c_size_t
        l.bcs_generate_fastq.argtypes = [C.POINTER(bcs_config), C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_uint]
        l.bcs_measure_int_peaks.restype = C.c_int
        l.bcs_measure_int_peaks.argtypes = [C.POINTER(C.c_double)]
        _lib = l
    return _lib


def measure_int_peaks():
    """Measured INT-pipe peaks of the current GPU, in 10^12 lane-ops/s: dict(lop3=, popc=, shf=)."""
    out = (C.c_double * 3)()
    rc = synth_lib().bcs_measure_int_peaks(out)
    if rc != 0:
        raise BcError(f"bcs_measure_int_peaks: cudaError {rc}")
    return dict(lop3=out[0], popc=out[1], shf=out[2])


# ---- barcode sets with pairwise Hamming distance >= 3: shortened Hamming codes over GF(4) -------------------------
_GF4_MUL = np.array([[0, 0, 0, 0], [0, 1, 2, 3], [0, 2, 3, 1], [0, 3, 1, 2]], dtype=np.uint8)


def _projective_points():
    """The 21 points of PG(2,4): pairwise linearly independent columns of a [21,18,3] Hamming parity-check matrix."""
    pts = []
    for a in range(4):
        for b in range(4):
            for c in range(4):
                v = (a, b, c)
                lead = next((x for x in v if x), 0)
                if lead == 1:
                    pts.append(v)
    assert len(pts) == 21
    return pts


def hamming_code_words(n, count, rng):
    """`count` distinct words of a shortened [n, n-3, 3] GF(4) Hamming code (n <= 21), as an (count, n) uint8 array of
    codes 0..3, translated by a random coset vector and with a random column order: any two words differ in >= 3
    positions, so one substitution is always uniquely correctable."""
    k = n - 3
    assert 1 <= k <= 18 and count <= 4 ** k
    pts = [p for p in _projective_points() if p not in ((1, 0, 0), (0, 1, 0), (0, 0, 1))]
    cols = np.array([pts[i] for i in rng.permutation(len(pts))[:k]], dtype=np.uint8)  # (k, 3)
    if 4 ** k <= 4 * count or k <= 10:
        ids = rng.permutation(4 ** k)[:count].astype(np.uint64)
    else:
        ids = np.unique(rng.integers(0, 4 ** k, size=int(count * 1.2) + 16, dtype=np.uint64))
        rng.shuffle(ids)
        ids = ids[:count]
        assert len(ids) == count
    msg = np.zeros((count, k), dtype=np.uint8)
    for j in range(k):
        msg[:, j] = (ids >> np.uint64(2 * j)) & np.uint64(3)
    par = np.zeros((count, 3), dtype=np.uint8)
    for j in range(k):
        for r in range(3):
            par[:, r] ^= _GF4_MUL[msg[:, j], cols[j, r]]
    words = np.concatenate([msg, par], axis=1)
    words = words[:, rng.permutation(n)]
    words ^= rng.integers(0, 4, size=n, dtype=np.uint8)[None, :]
    return words


def codes_to_dna(words):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    return [bytes(lut[w]).decode() for w in words]


def dna_to_codes(s):
    return np.array(["ACGT".index(ch) for ch in s], dtype=np.uint8)


# ---- workloads ------------------------------------------------------------------------------------------------------
WORKLOADS = {
    # BASELINE.json configs[0]: the reference's example files as shipped (7-nt barcodes in 6/10-nt slots, Q10)
    "example": dict(scheme=None, read_len=100, reads=1_000_000, min_quality=0.0, merge=True, enrich=True),
    # configs[1]: CRISPR screen, [8] x 96 samples + {20} x 80k guides, 75-nt reads
    "crispr": dict(scheme="[8]GTTTTAGAGCTAGAAATAGC{20}AAGTTAAAATAA", read_len=75, reads=100_000_000, min_quality=0.0,
                   merge=True, enrich=False, n_sample=96, n_counted=[80_000]),
    # configs[2] (and [4] per GPU): DEL 3-cycle, 3 x {8} x 1024 + (10) UMI, --min-quality 20 --enrich, 150-nt reads
    "del3": dict(scheme="ACGTTGCAGTCCAGTA{8}GATTACAG{8}CCTGAAGT{8}TGCATGCATGCA(10)AGGCTTAC", read_len=150,
                 reads=400_000_000, min_quality=20.0, merge=False, enrich=True, n_counted=[1024, 1024, 1024],
                 molecule_ratio=1.3125),
    # configs[3]: lineage tracing, raw {30} keys + (12) UMI, 100-nt reads
    "lineage": dict(scheme="TGACCTGAAGTCCATGCAAT{30}ACGGTACCTA(12)GGATCCTA", read_len=100, reads=100_000_000,
                    min_quality=0.0, merge=False, enrich=False, lineage_pool=30_000_000),
}

EXAMPLE_SCHEME = ("[10]AGCTACGAATCG{6}TGGA{6}TGGA{6}ACTAGAT(8)TAGA")
EXAMPLE_SAMPLES = [("AGCATAC", "Sample_name_1"), ("AACTTAC", "Sample_name_2")]
EXAMPLE_BARCODES = [("CAGAGAC", "Barcode_name_1", 1), ("TGATTGC", "Barcode_name_2", 1), ("ATGAAAT", "Barcode_name_3", 2),
                    ("GCGCCAT", "Barcode_name_4", 2), ("GATAGCT", "Barcode_name_5", 3), ("TTAGCTA", "Barcode_name_6", 3)]


class Workload:
    """Files + generator configuration of one named workload."""

    def __init__(self, name, workdir, reads=None, seed=BASE_SEED):
        if name not in WORKLOADS:
            raise BcError(f"unknown workload {name!r} (have {sorted(WORKLOADS)})")
        w = dict(WORKLOADS[name])
        self.name, self.workdir = name, workdir
        self.reads = int(reads if reads else w["reads"])
        self.read_len = w["read_len"]
        self.min_quality, self.merge, self.enrich = w["min_quality"], w["merge"], w["enrich"]
        os.makedirs(workdir, exist_ok=True)
        rng = np.random.default_rng(seed)
        self.fmt = os.path.join(workdir, "scheme.txt")
        self.samples = self.counted = None
        sample_set, counted_sets = None, None
        if name == "example":
            # the golden directory holds byte copies of the reference's three example files
            gold = os.path.join(os.path.dirname(PKG), "tests", "golden", "example")
            scheme = EXAMPLE_SCHEME
            for fn in ("scheme.txt", "samples.csv", "barcodes.csv"):
                with open(os.path.join(gold, fn), "rb") as f, open(os.path.join(workdir, fn), "wb") as g:
                    g.write(f.read())
            self.samples, self.counted = os.path.join(workdir, "samples.csv"), os.path.join(workdir, "barcodes.csv")
            sample_set = [d for d, _ in EXAMPLE_SAMPLES]
            counted_sets = [[d for d, _, k in EXAMPLE_BARCODES if k == j] for j in (1, 2, 3)]
        else:
            scheme = w["scheme"]
            with open(self.fmt, "w") as f:
                f.write(scheme + "\n")
        tokens = re.findall(r"\{\d+\}|\[\d+\]|\(\d+\)|[ACGT]+", scheme)
        slots, template, pos = [], [], 0
        for t in tokens:
            if t[0] in "{[(":
                n = int(t[1:-1])
                slots.append(dict(kind={"{": "B", "[": "S", "(": "R"}[t[0]], offset=pos, len=n))
                template += [255] * n
                pos += n
            else:
                template += list(dna_to_codes(t))
                pos += len(t)
        self.template_len = pos
        if name != "example":
            if w.get("n_sample"):
                s_len = next(s["len"] for s in slots if s["kind"] == "S")
                sample_set = codes_to_dna(hamming_code_words(s_len, w["n_sample"], rng))
                self.samples = os.path.join(workdir, "samples.csv")
                with open(self.samples, "w") as f:
                    f.write("Barcode,Sample_ID\n" + "".join(f"{d},sample_{i + 1:02d}\n" for i, d in enumerate(sample_set)))
            if w.get("n_counted"):
                counted_sets = []
                lines = ["Barcode,Barcode_ID,Barcode_Number"]
                bslots = [s for s in slots if s["kind"] == "B"]
                for k, (s, n) in enumerate(zip(bslots, w["n_counted"])):
                    st = codes_to_dna(hamming_code_words(s["len"], n, rng))
                    counted_sets.append(st)
                    lines += [f"{d},bb{k + 1}_{i + 1:05d},{k + 1}" for i, d in enumerate(st)]
                self.counted = os.path.join(workdir, "barcodes.csv")
                with open(self.counted, "w") as f:
                    f.write("\n".join(lines) + "\n")
        # generator configuration
        cfg = bcs_config()
        cfg.seed = seed
        cfg.read_len, cfg.template_len, cfg.n_slots = self.read_len, self.template_len, len(slots)
        cfg.max_start = min(8, self.read_len - self.template_len - 1)
        for i in range(BCS_MAX_READ):
            cfg.template_codes[i] = template[i] if i < len(template) else 255
        blob = []
        blob_len = 0
        k_counted = 0
        for i, s in enumerate(slots):
            d = cfg.slots[i]
            d.kind, d.offset, d.len = ord(s["kind"]), s["offset"], s["len"]
            refs = None
            if s["kind"] == "S" and sample_set:
                refs, d.skew = sample_set, 0
            elif s["kind"] == "B":
                if counted_sets:
                    refs, d.skew = counted_sets[k_counted], (2 if name == "crispr" else 0)
                elif w.get("lineage_pool"):
                    d.pool, d.skew = w["lineage_pool"], 2
                k_counted += 1
            if refs:
                arr = np.stack([dna_to_codes(r) for r in refs])
                d.n_ref, d.ref_len, d.ref_off = arr.shape[0], arr.shape[1], blob_len
                blob.append(arr.reshape(-1))
                blob_len += arr.size
        self.refs_blob = np.concatenate(blob) if blob else np.zeros(1, np.uint8)
        p = lambda x: min(0xFFFFFFFF, int(x * 2 ** 32))
        cfg.p_junk, cfg.p_lowq = p(0.02), p(0.03 if self.min_quality > 0 else 0.0)
        cfg.p_sub16, cfg.p_n16 = int(0.005 * 65536), int(0.001 * 65536)
        if w.get("molecule_ratio"):
            cfg.molecule_pool = max(1, int(self.reads * w["molecule_ratio"]))
            cfg.p_enriched, cfg.n_enriched = p(0.01), 1000
        cfg.q_mean, cfg.q_spread = 34, 8
        self.cfg = cfg
        self._refs_dev = None

    # the same files through the product's own set-up
    def run(self, bc, max_read_len=None):
        return bc.Run(self.fmt, self.samples, self.counted, min_quality=self.min_quality,
                      max_read_len=max_read_len or self.read_len)

    def generate_device(self, run, first, n, device="cuda:0", stream=None):
        """Reads [first, first+n) as a device-resident Batch (torch tensors)."""
        import torch
        from . import Batch
        if self._refs_dev is None or str(self._refs_dev.device) != str(torch.device(device)):
            self._refs_dev = torch.from_numpy(self.refs_blob).to(device)
        planes = torch.empty((n, run.plane_stride), dtype=torch.int32, device=device)
        read_len = torch.empty(n, dtype=torch.int16, device=device)
        qual = torch.empty((n, run.qual_stride), dtype=torch.uint8, device=device) if run.quality_on else None
        st = stream if stream is not None else torch.cuda.current_stream(device).cuda_stream
        rc = synth_lib().bcs_generate_device(C.byref(self.cfg), self._refs_dev.data_ptr(), first, n, run.max_read_len,
                                             planes.data_ptr(), read_len.data_ptr(), qual.data_ptr() if qual is not None else None,
                                             C.c_void_p(st))
        if rc != 0:
            raise BcError(f"bcs_generate_device: cudaError {rc}")
        return Batch(n, run.plane_stride, run.qual_stride, planes, read_len, qual, device=True)

    def fastq_bytes(self, first, n):
        return synth_lib().bcs_fastq_bytes(C.byref(self.cfg), first, n)

    def generate_fastq(self, first, n, threads=8):
        """Reads [first, first+n) as FASTQ text (a numpy uint8 array)."""
        need = self.fastq_bytes(first, n)
        out = np.empty(need, dtype=np.uint8)
        got = synth_lib().bcs_generate_fastq(C.byref(self.cfg), self.refs_blob.ctypes.data, first, n, out.ctypes.data, need, threads)
        if got != need:
            raise BcError("bcs_generate_fastq failed")
        return out

    def write_fastq(self, path, first, n, threads=8, chunk=2_000_000):
        with open(path, "wb") as f:
            for a in range(first, first + n, chunk):
                f.write(self.generate_fastq(a, min(chunk, first + n - a), threads).tobytes())
        return path
