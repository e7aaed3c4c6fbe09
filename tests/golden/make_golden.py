#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

The Rust reference cannot be built in this image (no cargo/rustc), so the fixtures come from a second,
independent restatement of the reference's hot path: the line-by-line Python mirror in tests/mirror.py (it
drives Python's `re` with the *same regex string* the reference builds, info.rs:263-298).  The C++ oracle
(oracle/) and the CUDA path are both checked against the vectors this script writes.  Run it from the repo
root in the build container:

    python tests/golden/make_golden.py            # needs /root/reference for the example files

It copies the reference's three example *data* files (scheme / barcode / sample CSV — config C1 of
BASELINE.json) into tests/golden/example/ because /root/reference does not exist on the GPU box.
Citations are file:line into the reference repository.
"""
import json
import os
import random
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


sys.path.insert(0, os.path.dirname(HERE))
from mirror import (Decoder, SequenceFormat, barcode_conversion, make_reads, max_seq_errors, rand_seq,  # noqa: E402,F401
                    sample_conversion)


CASES = {
    # name: (format text or None for the example scheme, sample csv, counted csv, read_len, n_reads, flags)
    "example": dict(example=True, read_len=100, n=400, min_quality=0.0, merge=True, enrich=True),
    "example_q20": dict(example=True, read_len=90, n=400, min_quality=20.0, merge=True, enrich=True),
    "example_tight": dict(example=True, read_len=68, n=300, min_quality=0.0, merge=False, enrich=False,
                          max_constant=2, max_sample=1, max_barcode=0),
    "del3_umi": dict(fmt="ACGTTGCAGTCCAGTA\n{8}\nGATTACAG\n{8}\nCCTGAAGT\n{8}\nTGCATGCATGCA\n(10)\nAGGCTTAC\n",
                     gen_counted=(3, 12, 8), read_len=110, n=500, min_quality=20.0, merge=False, enrich=True),
    "crispr": dict(fmt="[8]GTTTTAGAGCTAGAAATAGC{20}AAGTTAAAATAA\n", gen_samples=(6, 8), gen_counted=(1, 40, 20),
                   read_len=75, n=500, min_quality=0.0, merge=True, enrich=False),
    "lineage_raw": dict(fmt="# raw keys\nTGACCTGAAGTCCATGCAAT{30}ACGGTACCTA(12)GGATCCTA", read_len=100, n=400,
                        min_quality=0.0, merge=False, enrich=False),
    "sample_raw_two": dict(fmt="[2]ACGTACGTAC{5}{4}TTGACA", read_len=46, n=400, min_quality=25.0, merge=True,
                           enrich=True),
    "format_n": dict(fmt="ACGTACNNGTAC{6}TTGNCA(5)GG", read_len=48, n=400, min_quality=22.0, merge=False,
                     enrich=False, gen_counted=(1, 5, 6)),
    # sample barcode + three counted barcodes + random barcode + quality filter, merged and enriched output
    "del3_sample_umi": dict(fmt="[6]ACGTTGCAGTCCAGTA{8}GATTACAG{8}CCTGAAGT{8}TGCATGCATGCA(10)AGGCTTAC\n",
                            gen_samples=(4, 6), gen_counted=(3, 10, 8), read_len=120, n=500, min_quality=18.0,
                            merge=True, enrich=True),
    # reference barcodes of mixed lengths (shorter and longer than their slot, Q10) and with N in them: the whole-set
    # search with fix_error's "N never counts" rule on both sides (parse.rs:568-577)
    "refs_mixed_n": dict(fmt="[6]ACGTACGGTCAGT{9}TTGACCAT(6)GGCA\n",
                         sample_text="Barcode,Sample_ID\nACGTAC,s1\nTTGCAN,s2\nGGATCCA,s3\n",
                         counted_text="Barcode,Barcode_ID,Barcode_Number\nACGTTGCAA,b1,1\nTTGACNGTA,b2,1\nGGCATCAG,b3,1\n"
                                      "CATGGTACCA,b4,1\nAGTCAGTCA,b5,1\nNNGTCATGA,b6,1\nTCAGGTCAGTT,b7,1\n",
                         read_len=60, n=500, min_quality=0.0, merge=True, enrich=False),
    # reads far longer than the scheme: more than 64 window offsets (three 32-offset chunks in the GPU locate step)
    "long_reads": dict(fmt="GTCAGTTACGCATGCA{7}TTGACCAGTGCA{7}CAGGTTCAATGC(8)ACGT\n", gen_counted=(2, 9, 7), read_len=160, n=500,
                       min_quality=0.0, merge=False, enrich=True),
}


def gen_set(rng, n, length, min_dist=3):
    out = []
    tries = 0
    while len(out) < n and tries < 200000:
        tries += 1
        c = rand_seq(rng, length)
        if all(sum(a != b for a, b in zip(c, o)) >= min_dist for o in out):
            out.append(c)
    assert len(out) == n
    return out


def main():
    os.makedirs(os.path.join(HERE, "example"), exist_ok=True)
    if os.path.isdir(REF):
        for src, dst in (("scheme.example.txt", "scheme.txt"), ("barcode.example.csv", "barcodes.csv"),
                         ("sample_barcode.example.csv", "samples.csv")):
            with open(os.path.join(REF, src)) as f, open(os.path.join(HERE, "example", dst), "w") as g:
                g.write(f.read())
    for name, c in CASES.items():
        rng = random.Random("golden-" + name)
        case_dir = os.path.join(HERE, name)
        os.makedirs(case_dir, exist_ok=True)
        if c.get("example"):
            fmt_text = open(os.path.join(HERE, "example", "scheme.txt")).read()
            sample_text = open(os.path.join(HERE, "example", "samples.csv")).read()
            counted_text = open(os.path.join(HERE, "example", "barcodes.csv")).read()
        else:
            fmt_text = c["fmt"]
            sample_text, counted_text = c.get("sample_text"), c.get("counted_text")
            if "gen_samples" in c:
                n, ln = c["gen_samples"]
                sample_text = "Barcode,Sample_ID\n" + "".join(
                    "%s,sample_%02d\n" % (s, i + 1) for i, s in enumerate(gen_set(rng, n, ln)))
            if "gen_counted" in c:
                slots, n, ln = c["gen_counted"]
                counted_text = "Barcode,Barcode_ID,Barcode_Number\n"
                for k in range(slots):
                    for i, s in enumerate(gen_set(rng, n, ln)):
                        counted_text += "%s,bb%d_%03d,%d\n" % (s, k + 1, i + 1, k + 1)
        fmt = SequenceFormat(fmt_text)
        samples_hash = sample_conversion(sample_text) if sample_text else {}
        counted_hash = barcode_conversion(counted_text, fmt.barcode_num) if counted_text else []
        caps = max_seq_errors(c.get("max_sample"), fmt.sample_length_option, c.get("max_barcode"), fmt.barcode_lengths,
                              c.get("max_constant"), fmt.constant_region_length)
        dec = Decoder(fmt, samples_hash, counted_hash, caps, c["min_quality"])
        reads = make_reads(rng, fmt, samples_hash, counted_hash, c["n"], c["read_len"])
        outcomes = [dec.process(s, q) for s, q in reads]
        files = dec.write("golden", c["merge"], c["enrich"])
        with open(os.path.join(case_dir, "scheme.txt"), "w") as f:
            f.write(fmt_text)
        if sample_text:
            with open(os.path.join(case_dir, "samples.csv"), "w") as f:
                f.write(sample_text)
        if counted_text:
            with open(os.path.join(case_dir, "barcodes.csv"), "w") as f:
                f.write(counted_text)
        with open(os.path.join(case_dir, "reads.fastq"), "w") as f:
            for i, (s, q) in enumerate(reads):
                f.write("@r%d\n%s\n+\n%s\n" % (i, s, q))
        expected = dict(
            case=name,
            flags=dict(min_quality=c["min_quality"], merge=c["merge"], enrich=c["enrich"],
                       max_constant=c.get("max_constant"), max_sample=c.get("max_sample"),
                       max_barcode=c.get("max_barcode")),
            format_string=fmt.format_string, regions_string=fmt.regions_string, regex_string=fmt.regex_string,
            caps=dict(constant=caps[0], sample=caps[1], barcode=caps[2]),
            counters=dec.counters, outcomes=outcomes, files=files)
        with open(os.path.join(case_dir, "expected.json"), "w") as f:
            json.dump(expected, f, indent=0, sort_keys=True)
        print(name, dec.counters, "files:", len(files))


if __name__ == "__main__":
    sys.exit(main())
