"""The transfer form of a host batch (bc_wire_batch): what bch_wire_from_batch writes, expanded again by a numpy
restatement of the device kernels (k_wire_planes / k_wire_ncalls / k_wire_qual), gives back the bc_batch arrays bit for
bit — every quality form (8 / 6 / 4 / 2 bits), N calls as a list and as a dense plane, short quality lines (the 0xFF
mark), reads flagged unsupported.  CPU only; the GPU side is tests/test_gpu_parity.py::test_wire_*."""
import ctypes as C
import random

import numpy as np
import pytest

import ngs_barcode_count_b200 as bc
from helpers import load_golden


def as_array(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()


def expand(wb, run_qual_stride, plane_stride):
    """numpy restatement of launch_wire_expand (csrc/bc_kernels.cu)"""
    c = wb.c
    n, mrl = c.n_reads, c.max_read_len
    w = (mrl + 31) // 32
    lohi = as_array(c.lohi, n * 2 * w, np.uint32).reshape(n, 2 * w)
    planes = np.zeros((n, plane_stride), np.uint32)
    if c.nmask:
        nm = as_array(c.nmask, n * w, np.uint32).reshape(n, w)
    else:
        nm = np.zeros((n, w), np.uint32)
        reads, pos = as_array(c.n_read, c.n_calls, np.uint32), as_array(c.n_pos, c.n_calls, np.uint16)
        for r, p in zip(reads.tolist(), pos.tolist()):
            nm[r, p >> 5] |= np.uint32(1 << (p & 31))
    planes[:, :w] = lohi[:, :w] & ~nm
    planes[:, w:2 * w] = lohi[:, w:] & ~nm
    planes[:, 2 * w:3 * w] = nm
    read_len = as_array(c.read_len, n, np.uint16)
    qual = None
    if c.qual:
        bits, n_codes = c.qual_bits, (mrl + 3) // 4 * 4
        raw = as_array(c.qual, n * c.qual_stride, np.uint8).reshape(n, c.qual_stride)
        qual = np.full((n, run_qual_stride), ord("!"), np.uint8)
        if bits == 8:
            qual[:, :n_codes] = raw[:, :n_codes]
        else:
            stream = np.unpackbits(raw, axis=1, bitorder="little")[:, :n_codes * bits].reshape(n, n_codes, bits)
            codes = (stream * (1 << np.arange(bits, dtype=np.uint32))).sum(axis=2).astype(np.uint32)
            if bits == 6:
                qual[:, :n_codes] = np.where(codes == 63, 255, codes + 33).astype(np.uint8)
            else:
                qual[:, :n_codes] = np.array(list(c.qual_dict), np.uint8)[codes]
    return planes, read_len, qual


def random_reads(rng, n, max_len, alphabet_q, p_n=0.01, ragged=True):
    seqs, quals = [], []
    for _ in range(n):
        ln = rng.randint(1, max_len) if ragged and rng.random() < 0.3 else max_len
        s = "".join("N" if rng.random() < p_n else rng.choice("ACGT") for _ in range(ln))
        if rng.random() < 0.02:
            s = s[:ln // 2] + "X" + s[ln // 2 + 1:]  # unsupported character
        ql = ln if rng.random() < 0.9 else rng.randint(0, ln)  # short quality lines
        seqs.append(s)
        quals.append("".join(rng.choice(alphabet_q) for _ in range(ql)))
    return seqs, quals


ALPHABETS = {
    2: "#-7F",                                           # four bins (e.g. NovaSeq) -> 2 bits... unless a short line adds 0xFF
    4: "#,-27<AFJ",                                      # binned, up to 16 symbols
    6: "".join(chr(c) for c in range(33, 75)),           # unbinned Phred+33 up to 'J'
    8: "".join(chr(c) for c in range(33, 127)),          # up to '~' (long-read instruments)
}


@pytest.mark.parametrize("want_bits", [2, 4, 6, 8])
@pytest.mark.parametrize("max_len", [20, 75, 150, 201])
def test_wire_round_trip(want_bits, max_len, tmp_path):
    _, p = load_golden("example_q20")
    run = bc.Run(p["fmt"], p["samples"], p["counted"], min_quality=20.0, max_read_len=max(max_len, 100))
    rng = random.Random(want_bits * 1000 + max_len)
    ragged = want_bits != 2  # a short quality line adds the 0xFF mark: a fifth symbol
    seqs, quals = random_reads(rng, 300, max_len, ALPHABETS[want_bits], ragged=ragged)
    if not ragged:
        quals = [q.ljust(len(s), "#")[:len(s)] for s, q in zip(seqs, quals)]
    batch = run.pack(seqs, quals)
    mrl = run.max_read_len
    wb = bc.WireBatch(batch, mrl)
    assert wb.c.qual_bits <= want_bits
    planes, read_len, qual = expand(wb, run.qual_stride, run.plane_stride)
    assert np.array_equal(planes, batch.planes)
    assert np.array_equal(read_len, batch.read_len)
    # quality characters of a read's own positions (what lies beyond its length is never looked at and travels as code 0)
    inside = np.arange(run.qual_stride)[None, :] < (batch.read_len & 0x7FFF)[:, None]
    assert np.array_equal(qual[inside], batch.qual[inside])
    # every wider form holds the same batch
    for bits in (4, 6, 8):
        if bits > wb.c.qual_bits and (bits != 6 or want_bits <= 6):
            wider = bc.WireBatch(batch, mrl, qual_bits=bits)
            assert wider.c.qual_bits == bits
            assert np.array_equal(expand(wider, run.qual_stride, run.plane_stride)[2][inside], batch.qual[inside])
    # and a narrower one is refused
    if wb.c.qual_bits > 2:
        with pytest.raises(bc.BcError):
            bc.WireBatch(batch, mrl, qual_bits=2)


def test_wire_n_calls_list_and_dense():
    _, p = load_golden("example")
    run = bc.Run(p["fmt"], p["samples"], p["counted"])
    rng = random.Random(5)
    for p_n, dense in ((0.002, False), (0.5, True)):
        seqs, _ = random_reads(rng, 500, 100, "I", p_n=p_n)
        batch = run.pack(seqs)
        wb = bc.WireBatch(batch, run.max_read_len)
        assert bool(wb.c.nmask) == dense and wb.c.qual_bits == 0 and not wb.c.qual
        planes, read_len, _ = expand(wb, run.qual_stride, run.plane_stride)
        assert np.array_equal(planes, batch.planes) and np.array_equal(read_len, batch.read_len)
        if not dense:
            assert wb.c.n_calls == sum(s.count("N") + s.count("X") for s in seqs)


def test_wire_strides():
    lib = bc.lib()
    for mrl in (1, 4, 75, 150, 151, 1024):
        codes = lib.bc_wire_qual_codes(mrl)
        assert codes >= mrl and codes % 4 == 0 and codes - mrl < 4
        for bits in (2, 4, 6, 8):
            qs = lib.bc_wire_qual_stride(mrl, bits)
            assert qs % 4 == 0 and qs * 8 >= codes * bits and (qs - 4) * 8 < codes * bits
    # DEL geometry: 150-nt reads cross PCIe in 40 + 2 + 116 bytes (+ the N list) instead of 60 + 2 + 156
    assert lib.bc_wire_qual_stride(150, 6) == 116


@pytest.mark.parametrize("level", [0, 1, 2])
def test_packers_agree_across_simd_levels(level):
    """The scalar, AVX2 and AVX-512 forms of the base packer, the 6-bit quality packer and the line-end scanner give the
    same bytes (the level is capped by what the CPU has)."""
    lib = bc.lib()
    _, p = load_golden("example_q20")
    rng = random.Random(11)
    try:
        outs = []
        for lv in (0, level):
            lib.bch_set_simd_level(lv)
            res = []
            for max_len in (1, 31, 32, 33, 63, 64, 65, 100, 127, 128, 150, 151, 250):
                run = bc.Run(p["fmt"], p["samples"], p["counted"], min_quality=20.0, max_read_len=max(max_len, 100))
                r2 = random.Random(max_len)
                seqs, quals = random_reads(r2, 200, max_len, ALPHABETS[6])
                batch = run.pack(seqs, quals)
                wb = bc.WireBatch(batch, run.max_read_len, qual_bits=6)
                res.append((batch.planes.copy(), batch.read_len.copy(), batch.qual.copy(), bytes(wb.buf[:wb.bound(batch.n, run.max_read_len, True)])))
            outs.append(res)
        for a, b in zip(*outs):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    finally:
        lib.bch_set_simd_level(2)
