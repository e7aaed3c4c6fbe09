"""Differential fuzz of the two independent restatements of the reference's hot path — the C++ oracle (oracle/) and the
Python mirror (tests/mirror.py, which drives `re` with the regex string the reference builds) — over hundreds of random
schemes: format-N runs, sample / counted / random barcodes in any order, conversion files with references shorter and
longer than their slot and with N in them, cap overrides, the quality filter, ragged reads (shorter than the scheme
too) and quality lines shorter than their sequence.  Per read: status, located offset, repaired flag, sample, counted
barcodes and random barcode must agree; per scheme: the six counters and the whole canonical CSV set.

The Rust reference cannot run in this image, so this does not pin the oracle to the reference; it shows that two
readings of parse.rs / info.rs / output.rs written separately do not diverge anywhere the generator reaches
(parse.rs:89-163, 270-375, 439-593; info.rs:215-310, 364-456, 490-543, 662-808, 840-904; output.rs:74-485)."""
import os
import random

import pytest

import mirror
from helpers import Oracle, assert_same_csv_set, read_csv_dir

N_SCHEMES = int(os.environ.get("BC_FUZZ_SCHEMES", "520"))
READS_PER_SCHEME = 36


def random_scheme(rng):
    """-> (format text, sample csv text or None, counted csv text or None, flags)"""
    n_counted = rng.choice([1, 1, 2, 2, 3, 4])
    with_sample = rng.random() < 0.5
    with_random = rng.random() < 0.6
    slots = ["B"] * n_counted + (["S"] if with_sample else []) + (["R"] if with_random else [])
    rng.shuffle(slots)
    if with_sample and rng.random() < 0.6:  # usual layout: sample first
        slots.remove("S")
        slots.insert(0, "S")
    parts, lens = [], {}
    counted_lens = []
    for i, s in enumerate(slots):
        if rng.random() < 0.85 or i == 0:
            parts.append(mirror.rand_seq(rng, rng.randint(2, 18)))
            if rng.random() < 0.2:  # a format-N run inside / after a constant
                parts.append("N" * rng.randint(1, 3))
                parts.append(mirror.rand_seq(rng, rng.randint(2, 6)))
        ln = rng.randint(3, 14) if s != "B" else rng.randint(3, 22)
        if s == "B":
            counted_lens.append(ln)
            parts.append("{%d}" % ln)
        elif s == "S":
            lens["S"] = ln
            parts.append("[%d]" % ln)
        else:
            parts.append("(%d)" % ln)
    if rng.random() < 0.85:
        parts.append(mirror.rand_seq(rng, rng.randint(2, 12)))
    seps = ["", "", "\n", " "]
    text = "# fuzz\n" + "".join(p + rng.choice(seps) for p in parts) + "\n"

    def ref_set(n, slot_len):
        out = set()
        while len(out) < n:
            ln = slot_len
            r = rng.random()
            if r < 0.12:
                ln = max(1, slot_len - rng.randint(1, 2))  # shorter than the slot (Q10)
            elif r < 0.24:
                ln = slot_len + rng.randint(1, 2)
            d = mirror.rand_seq(rng, ln)
            if rng.random() < 0.08:
                k = rng.randrange(ln)
                d = d[:k] + "N" + d[k + 1:]
            out.add(d)
        return sorted(out)

    sample_text = counted_text = None
    if with_sample and rng.random() < 0.7:
        sample_text = "Barcode,Sample_ID\n" + "".join("%s,smp%d\n" % (d, i) for i, d in enumerate(ref_set(rng.randint(1, 5), lens["S"])))
    if rng.random() < 0.7:
        counted_text = "Barcode,Barcode_ID,Barcode_Number\n"
        for k, ln in enumerate(counted_lens):
            for i, d in enumerate(ref_set(rng.randint(1, 7), ln)):
                counted_text += "%s,b%d_%d,%d\n" % (d, k + 1, i, k + 1)
    flags = dict(min_quality=rng.choice([0.0, 0.0, 12.0, 20.0, 27.5]), merge=rng.random() < 0.5, enrich=rng.random() < 0.5,
                 max_constant=rng.choice([None, None, None, 0, 1, 3]), max_sample=rng.choice([None, None, 0, 1, 2]),
                 max_barcode=rng.choice([None, None, 0, 1, 2]))
    return text, sample_text, counted_text, flags


def ragged(rng, reads, L):
    """make_reads gives full-length pairs; cut some reads (below the scheme length too) and some quality lines."""
    out = []
    for seq, qual in reads:
        r = rng.random()
        if r < 0.10:
            seq = seq[:rng.randint(max(1, L - 6), len(seq))]
            qual = qual[:len(seq)]
        elif r < 0.14:
            seq = seq[:rng.randint(1, L)]
            qual = qual[:len(seq)]
        r = rng.random()
        if r < 0.10:
            qual = qual[:rng.randint(0, len(qual))]  # a quality line shorter than its sequence (the zip just stops)
        elif r < 0.13:
            qual = qual + "I" * rng.randint(1, 5)
        out.append((seq, qual))
    return out


@pytest.mark.parametrize("chunk", range(8))
def test_oracle_equals_mirror_on_random_schemes(chunk, tmp_path):
    per_chunk = (N_SCHEMES + 7) // 8
    checked = 0
    for k in range(per_chunk):
        seed = "fuzz-%d-%d" % (chunk, k)
        rng = random.Random(seed)
        text, sample_text, counted_text, fl = random_scheme(rng)
        fmt = mirror.SequenceFormat(text)
        if fmt.barcode_num == 0:
            continue
        d = tmp_path / ("s%d" % k)
        d.mkdir()
        (d / "scheme.txt").write_text(text)
        s_path = c_path = None
        if sample_text:
            s_path = str(d / "samples.csv")
            (d / "samples.csv").write_text(sample_text)
        if counted_text:
            c_path = str(d / "barcodes.csv")
            (d / "barcodes.csv").write_text(counted_text)
        samples_hash = mirror.sample_conversion(sample_text) if sample_text else {}
        counted_hash = mirror.barcode_conversion(counted_text, fmt.barcode_num) if counted_text else []
        caps = mirror.max_seq_errors(fl["max_sample"], fmt.sample_length_option, fl["max_barcode"], fmt.barcode_lengths,
                                     fl["max_constant"], fmt.constant_region_length)
        dec = mirror.Decoder(fmt, samples_hash, counted_hash, caps, fl["min_quality"])
        read_len = fmt.length + rng.randint(0, 40)
        reads = ragged(rng, mirror.make_reads(rng, fmt, samples_hash, counted_hash, READS_PER_SCHEME, read_len,
                                              sub_rate=0.03, n_rate=0.012), fmt.length)
        out_dir = d / "out"
        out_dir.mkdir()
        orc = Oracle(str(d / "scheme.txt"), s_path, c_path, min_quality=fl["min_quality"], merge=fl["merge"], enrich=fl["enrich"],
                     outdir=str(out_dir), prefix="p", max_barcode=fl["max_barcode"], max_sample=fl["max_sample"],
                     max_constant=fl["max_constant"])
        info = orc.format_info()
        assert info["format_string"] == fmt.format_string and info["regions_string"] == fmt.regions_string, seed
        assert (info["max_constant"], info["max_sample"], info["max_barcode"]) == (caps[0], caps[1], caps[2]), seed
        for i, (seq, qual) in enumerate(reads):
            want = dec.process(seq, qual)
            got = orc.process(seq, qual)
            if want["status"] in ("matched", "duplicate"):
                assert got == want, (seed, i, seq, qual, got, want)
            else:
                assert (got["status"], got["offset"], got["repaired"]) == (want["status"], want["offset"], want["repaired"]), \
                    (seed, i, seq, qual, got, want)
        assert orc.counters() == dec.counters, seed
        orc.write_files()
        want_files = {fn: lines for fn, lines in dec.write("p", fl["merge"], fl["enrich"]).items()}
        assert_same_csv_set(read_csv_dir(str(out_dir), "p"), want_files)
        orc.close()
        checked += 1
    assert checked >= per_chunk * 0.9
