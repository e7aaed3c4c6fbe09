"""CPU tests that pin the oracle (oracle/) — the reference's own known-answer tests for this path, the
hand-derived vectors of SURVEY.md §8(c), and the committed golden fixtures produced by the independent
Python mirror (tests/golden/make_golden.py)."""
import ctypes as C
import os

import pytest

from helpers import GOLDEN, Oracle, assert_same_csv_set, golden_cases, load_golden, oracle_lib, read_csv_dir, read_fastq

EX = os.path.join(GOLDEN, "example")


def fix_error(seq, cands, m):
    lib = oracle_lib()
    arr = (C.c_char_p * len(cands))(*[c.encode() for c in cands])
    out = C.create_string_buffer(256)
    return out.value.decode() if lib.orc_fix_error(seq.encode(), arr, len(cands), m, out, 256) else None


def max_errors(sample_err, sample_size, barcode_err, sizes, const_err, const_size):
    lib = oracle_lib()
    neg = lambda v: -1 if v is None else v
    arr = (C.c_ushort * len(sizes))(*sizes)
    c, s = C.c_int(), C.c_int()
    b = (C.c_int * len(sizes))()
    lib.orc_max_errors(neg(sample_err), neg(sample_size), neg(barcode_err), arr, len(sizes), neg(const_err), const_size,
                       C.byref(c), C.byref(s), b)
    return c.value, s.value, list(b)


# ---- the reference's rustdoc known-answer tests -----------------------------------------------------------------
def test_kat_fix_error_doctest():  # parse.rs:540-551
    assert fix_error("AGTAG", ["AGCAG", "ACAAG", "AGCAA"], 5 // 5) == "AGCAG"
    assert fix_error("AGTAG", ["AGCAG", "AGAAG", "AGCAA"], 5 // 5) is None


def test_kat_max_constant_errors():  # info.rs:547-564
    assert max_errors(None, 10, None, [8, 8, 8], None, 30)[0] == 6
    assert max_errors(None, 10, None, [8, 8, 8], 3, 30)[0] == 3


def test_kat_max_sample_errors():  # info.rs:571-588
    assert max_errors(None, 10, None, [8, 8, 8], None, 30)[1] == 2
    assert max_errors(3, 10, None, [8, 8, 8], None, 30)[1] == 3


def test_kat_max_barcode_errors():  # info.rs:595-612
    assert max_errors(None, 10, None, [8, 8, 8], None, 30)[2] == [1, 1, 1]
    assert max_errors(None, 10, 2, [8, 8, 8], None, 30)[2] == [2, 2, 2]


def test_fix_error_properties():
    # order independence, N wildcards on both sides, truncation to the shorter string (Q10)
    assert fix_error("AGTAG", ["AGCAA", "ACAAG", "AGCAG"], 1) == "AGCAG"
    assert fix_error("ANTAG", ["AGTAG", "ACTAC"], 1) == "AGTAG"
    assert fix_error("AGTAG", ["AGNAG", "CCCCC"], 0) == "AGNAG"
    assert fix_error("CAGAGA", ["CAGAGAC", "TGATTGC"], 1) == "CAGAGAC"
    assert fix_error("AAAA", ["AAAT", "AATA"], 1) is None  # tie at the minimum
    assert fix_error("AAAA", ["TTTT"], 1) is None
    assert fix_error("AAAA", [], 1) is None


# ---- hand-derived vectors on the reference's example files ------------------------------
SAMPLE, C1, B1, C2, B2, C3, B3, C4, UMI, C5 = ("AGCATACGGG", "AGCTACGAATCG", "CAGAGA", "TGGA", "ATGAAA", "TGGA",
                                               "GATAGC", "ACTAGAT", "ACGTACGT", "TAGA")
G1 = SAMPLE + C1 + B1 + C2 + B2 + C3 + B3 + C4 + UMI + C5
ROW = "CAGAGAC,ATGAAAT,GATAGCT"


def example_oracle(**kw):
    return Oracle(os.path.join(EX, "scheme.txt"), os.path.join(EX, "samples.csv"), os.path.join(EX, "barcodes.csv"), **kw)


def test_example_format_and_caps():
    o = example_oracle()
    info = o.format_info()
    assert info["format_string"] == "N" * 10 + C1 + "N" * 6 + C2 + "N" * 6 + C3 + "N" * 6 + C4 + "N" * 8 + C5
    assert info["regions_string"] == "S" * 10 + "C" * 12 + "B" * 6 + "C" * 4 + "B" * 6 + "C" * 4 + "B" * 6 + "C" * 7 + "R" * 8 + "C" * 4
    assert (info["constant_len"], info["max_constant"], info["max_sample"], info["max_barcode"]) == (31, 6, 2, [1, 1, 1])


def test_g1_exact_and_g2_umi_duplicate():
    o = example_oracle()
    q = "I" * (len(G1) + 1)
    r = o.process(G1 + "A", q)
    assert (r["status"], r["offset"], r["repaired"], r["sample"], r["barcodes"], r["random"]) == (
        "matched", 0, False, "AGCATAC", ROW, UMI)
    assert o.process(G1 + "A", q)["status"] == "duplicate"
    assert o.counters() == dict(matched=1, duplicates=1, constant_region=0, low_quality=0, sample_barcode=0, barcode=0)


def test_g3_g4_barcode_fix_and_reject():
    o = example_oracle()
    q = "I" * 68
    assert o.process(G1.replace(B1, "CAGTGA", 1) + "A", q)["barcodes"] == ROW
    assert o.process(G1.replace(B1, "CTGTGA", 1) + "A", q)["status"] == "barcode"


def test_g5_sample_tie():
    o = example_oracle()
    assert o.process("AACATAC" + G1[7:] + "A", "I" * 68)["status"] == "sample_barcode"


def test_g6_repair_and_quality_window_at_zero():
    bad_c1 = "AGCTACGTATCG"
    read = "TTTTT" + G1.replace(C1, bad_c1, 1) + "A"
    o = example_oracle()
    r = o.process(read, "I" * len(read))
    assert (r["status"], r["offset"], r["repaired"], r["barcodes"]) == ("matched", 5, True, ROW)
    # Q6: after a repair the quality window is q[0..67), not q[5..72): low scores at 0..9 (the 'S' run when read
    # from offset 0) must drop the read, low scores at 5..14 only partly overlap and do not.
    o = example_oracle(min_quality=20.0)
    q = "#" * 10 + "I" * (len(read) - 10)
    assert o.process(read, q)["status"] == "low_quality"
    o = example_oracle(min_quality=20.0)
    q = "I" * 5 + "#" * 4 + "I" * (len(read) - 9)  # mean of q[0..10) = (6*40+4*2)/10 = 24.8 >= 20
    assert o.process(read, q)["status"] == "matched"


def test_g7_last_offset_not_repaired_but_exact_matches():
    bad = G1.replace(C1, "AGCTACGTATCG", 1)
    o = example_oracle()
    assert o.process("TTTTT" + bad, "I" * (len(bad) + 5))["status"] == "constant_region"  # Q3
    assert o.process("TTTTT" + G1, "I" * (len(G1) + 5))["status"] == "matched"
    assert o.process(bad, "I" * len(bad))["status"] == "constant_region"  # R == L: no window at all


def test_g8_read_n_in_constant():
    read = G1.replace(C1, "AGCTANGAATCG", 1) + "A"
    o = example_oracle()
    r = o.process(read, "I" * len(read))
    assert (r["status"], r["offset"], r["repaired"], r["barcodes"]) == ("matched", 0, True, ROW)


def test_q8_last_run_not_quality_tested(tmp_path):
    fmt = tmp_path / "f.txt"
    fmt.write_text("ACGTAC{4}GGTT(4)")  # template ends in the random barcode: its quality is never looked at
    o = Oracle(str(fmt), min_quality=30.0)
    read = "ACGTAC" + "TTTT" + "GGTT" + "CCCC" + "A"
    assert o.process(read, "I" * 14 + "#" * 5)["status"] == "matched"
    assert o.process(read, "I" * 6 + "#" * 4 + "I" * 9)["status"] == "low_quality"


def test_q16_every_listed_sample_gets_a_file(tmp_path):
    o = example_oracle(outdir=str(tmp_path), prefix="t", merge=True, enrich=True)
    o.process(G1 + "A", "I" * 68)
    names = o.write_files()
    assert "t_Sample_name_2_counts.csv" in names
    files = read_csv_dir(str(tmp_path), "t")
    assert files["t_Sample_name_2_counts.csv"] == ["Barcode_1,Barcode_2,Barcode_3,Count"]
    assert files["t_Sample_name_1_counts.csv"] == ["Barcode_1,Barcode_2,Barcode_3,Count",
                                                   "Barcode_name_1,Barcode_name_3,Barcode_name_5,1"]
    assert files["t_counts.all.csv"] == ["Barcode_1,Barcode_2,Barcode_3,Sample_name_1,Sample_name_2",
                                         "Barcode_name_1,Barcode_name_3,Barcode_name_5,1,0"]
    assert files["t_Sample_name_1_counts.Single.csv"][1:] == [",,Barcode_name_5,1", ",Barcode_name_3,,1", "Barcode_name_1,,,1"]
    assert files["t_Sample_name_1_counts.Double.csv"][1:] == [",Barcode_name_3,Barcode_name_5,1", "Barcode_name_1,,Barcode_name_5,1",
                                                              "Barcode_name_1,Barcode_name_3,,1"]


# ---- golden fixtures from the independent Python mirror -----------------------------------------------------------
@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_golden(case, tmp_path):
    exp, paths = load_golden(case)
    fl = exp["flags"]
    o = Oracle(paths["fmt"], paths["samples"], paths["counted"], min_quality=fl["min_quality"], merge=fl["merge"],
               enrich=fl["enrich"], outdir=str(tmp_path), prefix="golden", max_barcode=fl["max_barcode"],
               max_sample=fl["max_sample"], max_constant=fl["max_constant"])
    info = o.format_info()
    assert info["format_string"] == exp["format_string"]
    assert info["regions_string"] == exp["regions_string"]
    assert (info["max_constant"], info["max_sample"], info["max_barcode"]) == (
        exp["caps"]["constant"], exp["caps"]["sample"], exp["caps"]["barcode"])
    reads = read_fastq(paths["fastq"])
    assert len(reads) == len(exp["outcomes"])
    for i, ((seq, qual), want) in enumerate(zip(reads, exp["outcomes"])):
        got = o.process(seq, qual)
        assert got["status"] == want["status"], (i, got, want)
        assert got["offset"] == want["offset"] and got["repaired"] == want["repaired"], (i, got, want)
        if want["status"] in ("matched", "duplicate"):
            assert (got["sample"], got["barcodes"], got["random"]) == (want["sample"], want["barcodes"], want["random"]), i
    assert o.counters() == exp["counters"]
    o.write_files()
    assert_same_csv_set(read_csv_dir(str(tmp_path), "golden"), exp["files"])


@pytest.mark.parametrize("case", ["example", "del3_umi"])
def test_threaded_run_equals_serial(case, tmp_path):
    exp, paths = load_golden(case)
    fl = exp["flags"]
    o = Oracle(paths["fmt"], paths["samples"], paths["counted"], min_quality=fl["min_quality"], merge=fl["merge"],
               enrich=fl["enrich"], outdir=str(tmp_path), prefix="golden")
    secs, total = o.run_fastq(paths["fastq"], 4)
    assert total == len(exp["outcomes"])
    assert o.counters() == exp["counters"]
    o.write_files()
    assert_same_csv_set(read_csv_dir(str(tmp_path), "golden"), exp["files"])
