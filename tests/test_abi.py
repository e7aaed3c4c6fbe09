"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the headers declare, the host
side (scheme / CSV set-up, packing) agrees with the golden fixtures, and without a GPU the library fails loudly
instead of computing anything on the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ngs_barcode_count_b200 as bc
from helpers import GOLDEN, ROOT, golden_cases, load_golden, read_fastq


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bch?_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = bc.lib()
    names = declared_functions("bc_b200.h") + declared_functions("bc_host.h")
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/ but not exported by libbc_b200.so"
        assert name in bc._PROTOS, f"{name} has no ctypes prototype"
    assert sorted(bc._PROTOS) == sorted(names)


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the ABI structs have the size the C compiler gives the header's structs (a field added on
    one side only would shift everything behind it)."""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "bc_b200.h"\n#include "bc_host.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(bc_slot), sizeof(bc_config), sizeof(bc_batch), '
                   'sizeof(bc_table), sizeof(bc_profile), sizeof(bc_decode_out), sizeof(bch_args), sizeof(bc_wire_batch)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(t) for t in (bc.bc_slot, bc.bc_config, bc.bc_batch, bc.bc_table, bc.bc_profile, bc.bc_decode_out, bc.bch_args,
                               bc.bc_wire_batch)]
    assert got == want


def test_cli_binary_exists_and_prints_version():
    import subprocess
    out = subprocess.run([bc.CLI_PATH, "--version"], capture_output=True, text=True)
    assert out.returncode == 0 and "NGS-Barcode-Count" in out.stdout


def test_strides():
    lib = bc.lib()
    for r in (1, 31, 32, 33, 75, 100, 150, 151, 1024):
        w = lib.bc_plane_words(r)
        assert w == (r + 31) // 32
        assert lib.bc_plane_stride(r) == (3 * w) | 1  # odd: conflict-free per-thread rows in shared memory
        qs = lib.bc_qual_stride(r)
        assert qs >= r and qs % 4 == 0 and (qs // 4) % 2 == 1


@pytest.mark.parametrize("case", golden_cases())
def test_host_setup_matches_golden(case):
    exp, p = load_golden(case)
    fl = exp["flags"]
    run = bc.Run(p["fmt"], p["samples"], p["counted"], min_quality=fl["min_quality"], max_barcode=fl["max_barcode"],
                 max_sample=fl["max_sample"], max_constant=fl["max_constant"])
    cfg = run.cfg
    assert cfg.template_chars[:cfg.template_len].decode() == exp["format_string"]
    assert cfg.region_codes[:cfg.region_len].decode() == exp["regions_string"]
    assert cfg.max_const_err == exp["caps"]["constant"]
    counted = [run.slot(i).max_err for i in range(run.n_slots) if run.slot(i).kind == ord("B")]
    assert counted == exp["caps"]["barcode"]
    sample = [run.slot(i).max_err for i in range(run.n_slots) if run.slot(i).kind == ord("S")]
    if sample and exp["caps"]["sample"] is not None:
        assert sample[0] == exp["caps"]["sample"]
    # slots sit where the template has its N runs
    for i in range(run.n_slots):
        s = run.slot(i)
        assert exp["format_string"][s.offset:s.offset + s.len] == "N" * s.len
        assert set(exp["regions_string"][:0]) <= set("SBRC")


def test_setup_errors():
    ex = os.path.join(GOLDEN, "example")
    with pytest.raises(bc.BcError):
        bc.Run(os.path.join(ex, "missing.txt"))
    with pytest.raises(bc.BcError):  # sample file without [n] in the scheme: the reference loses every count (Q15)
        bc.Run(os.path.join(GOLDEN, "lineage_raw", "scheme.txt"), samples=os.path.join(ex, "samples.csv"))


def py_pack(seq, W):
    lo, hi, nm = [0] * W, [0] * W, [0] * W
    for i, ch in enumerate(seq):
        w, b = i >> 5, 1 << (i & 31)
        if ch == "C":
            lo[w] |= b
        elif ch == "G":
            hi[w] |= b
        elif ch == "T":
            lo[w] |= b
            hi[w] |= b
        elif ch != "A":
            nm[w] |= b
    return lo + hi + nm


def test_pack_bit_layout():
    ex = os.path.join(GOLDEN, "example_q20")
    run = bc.Run(os.path.join(ex, "scheme.txt"), os.path.join(ex, "samples.csv"), os.path.join(ex, "barcodes.csv"),
                 min_quality=20.0)
    reads = read_fastq(os.path.join(ex, "reads.fastq"))[:64]
    seqs = [r[0] for r in reads] + ["", "A", "ACGTN" * 20, "ACGTX"]
    quals = [r[1] for r in reads] + ["", "I", "I" * 100, "IIIII"]
    b = run.pack(seqs, quals)
    W = bc.lib().bc_plane_words(run.max_read_len)
    for i, s in enumerate(seqs):
        assert list(b.planes[i, :3 * W]) == py_pack(s, W), i
        want_len = len(s) | (bc.BC_READ_UNSUPPORTED if "X" in s else 0)
        assert b.read_len[i] == want_len
        assert bytes(b.qual[i, :len(s)]) == quals[i].encode()
        assert set(bytes(b.qual[i, len(s):])) <= {ord("!")}
    with pytest.raises(bc.BcError):
        run.pack(["A" * (run.max_read_len + 1)], ["I" * (run.max_read_len + 1)])
    # a quality line shorter than its sequence is not an error (the reference zips scores with region codes and simply
    # stops with the line, parse.rs:338-343): from the line's last score on the row holds 255, which no threshold rejects
    b = run.pack(["ACGTACGTAC", "ACGT", "ACGT"], ["IIII", "", "IIIIIIII"])
    assert bytes(b.qual[0, :3]) == b"III" and set(bytes(b.qual[0, 3:])) == {255}
    assert set(bytes(b.qual[1])) == {255}
    assert bytes(b.qual[2, :4]) == b"IIII" and set(bytes(b.qual[2, 4:])) == {ord("!")}  # a longer line is cut
    assert list(b.read_len[:3]) == [10, 4, 4]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ex = os.path.join(GOLDEN, "example")
    run = bc.Run(os.path.join(ex, "scheme.txt"), os.path.join(ex, "samples.csv"), os.path.join(ex, "barcodes.csv"))
    with pytest.raises(bc.BcError, match="no usable CUDA device|CUDA"):
        bc.Counter(run)
    import subprocess
    out = subprocess.run([bc.CLI_PATH, "-f", os.path.join(ex, "reads.fastq"), "-q", os.path.join(ex, "scheme.txt")],
                         capture_output=True, text=True)
    assert out.returncode != 0 and "CUDA" in out.stderr


@pytest.mark.parametrize("case", ["del3_umi", "crispr", "lineage_raw", "format_n", "refs_mixed_n", "long_reads"])
def test_specialized_decode_kernel_compiles_for_sm_100a(case):
    """bc_jit.cu hands NVRTC the library's own device code (csrc/bc_decode.cuh, embedded at build time) with the run
    constants as a constant object.  No GPU is needed to compile it: the cubin must come out for every kind of scheme."""
    from helpers import load_golden
    exp, p = load_golden(case)
    run = bc.Run(p["fmt"], p["samples"], p["counted"], min_quality=exp["flags"]["min_quality"])
    ok, log = run.jit_check()
    assert ok and log.startswith("cubin bytes: ") and int(log.split(": ")[1]) > 10_000, log
