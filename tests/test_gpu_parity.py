"""GPU parity tests: the CUDA decode-and-count path, called through the C ABI (include/bc_b200.h, bc_host.h), against
the committed golden fixtures (tests/golden/, produced by the Python mirror of the reference) and against the CPU
oracle (oracle/) on seeded random inputs.  Bit-exact: per-read status / offset / repaired flag / decoded barcodes,
the outcome counters and the canonical CSV set must all be equal."""
import os
import random
import subprocess

import numpy as np
import pytest

import ngs_barcode_count_b200 as bc
from helpers import (GOLDEN, Oracle, assert_same_csv_set, bgzf_compress, golden_cases, load_golden, read_csv_dir, read_fastq)

pytestmark = pytest.mark.gpu

ORACLE_TO_GPU_STATUS = {"matched": "matched", "duplicate": "matched", "constant_region": "constant_region",
                        "low_quality": "low_quality", "sample_barcode": "sample_barcode", "barcode": "barcode"}


def make_run(paths, fl, **kw):
    return bc.Run(paths["fmt"], paths["samples"], paths["counted"], min_quality=fl["min_quality"],
                  max_barcode=fl["max_barcode"], max_sample=fl["max_sample"], max_constant=fl["max_constant"], **kw)


def decoded_strings(run, ctr, lo, hi):
    """(sample, 'b1,b2,..', random) of one matched read, as the oracle reports them."""
    per_slot = ctr.key_decode(lo, hi, with_umi=True)
    sample, counted, rnd = "barcode", [], None
    for s in range(run.n_slots):
        kind = chr(run.slot(s).kind)
        if kind == "S":
            sample = per_slot[s]
        elif kind == "B":
            counted.append(per_slot[s])
        else:
            rnd = per_slot[s]
    return sample, ",".join(counted), rnd


def check_reads_against(run, ctr, reads, outcomes):
    batch = run.pack([r[0] for r in reads], [r[1] for r in reads])
    got = ctr.decode_only(batch)
    st, off, rep = ctr.locate_only(batch)
    for i, want in enumerate(outcomes):
        g_status = bc.STATUS_NAMES[got["status"][i]]
        assert g_status == ORACLE_TO_GPU_STATUS[want["status"]], (i, g_status, want, reads[i][0])
        located = want["offset"] >= 0
        assert (st[i] == 0) == located, (i, want)
        assert off[i] == want["offset"] and bool(rep[i]) == want["repaired"], (i, off[i], rep[i], want)
        assert got["offset"][i] == want["offset"] and bool(got["repaired"][i]) == want["repaired"], (i, want)
        if want["status"] in ("matched", "duplicate"):
            assert decoded_strings(run, ctr, got["key_lo"][i], got["key_hi"][i]) == (
                want["sample"], want["barcodes"], want["random"]), (i, want)
    return batch


@pytest.mark.parametrize("case", golden_cases())
def test_golden_per_read_and_csv(case, tmp_path):
    exp, paths = load_golden(case)
    fl = exp["flags"]
    run = make_run(paths, fl)
    ctr = bc.Counter(run)
    reads = read_fastq(paths["fastq"])
    batch = check_reads_against(run, ctr, reads, exp["outcomes"])
    # hooks do not count
    assert sum(ctr.counters().values()) == 0
    # count in three uneven batches (the last one a single read), then compare counters and files
    n = batch.n
    for a, b in ((0, n // 3), (n // 3, n - 1), (n - 1, n)):
        ctr.submit(batch.slice(a, b))
    c = ctr.counters()
    assert c.pop("unsupported") == 0
    assert c == exp["counters"]
    ctr.write_counts(str(tmp_path), "golden", merge=fl["merge"], enrich=fl["enrich"])
    assert_same_csv_set(read_csv_dir(str(tmp_path), "golden"), exp["files"])
    # reset clears everything; a second pass gives the same answer
    ctr.reset()
    assert sum(ctr.counters().values()) == 0
    ctr.submit(batch)
    c = ctr.counters()
    c.pop("unsupported")
    assert c == exp["counters"]


@pytest.mark.parametrize("case", golden_cases())
def test_golden_specialized_kernel(case, tmp_path):
    """The decode kernel compiled for the run by NVRTC (bc_jit.cu: the same device code with the run constants folded in,
    what bc_create does on its own for jobs of 2^22 reads and more) against the golden per-read outcomes, counters and CSV
    set — and the launches really went through it."""
    exp, paths = load_golden(case)
    fl = exp["flags"]
    run = make_run(paths, fl)
    ctr = bc.Counter(run, flags=bc.BC_CFG_SPECIALIZE)
    assert ctr.profile()["specialization"] == "specialised"
    reads = read_fastq(paths["fastq"])
    batch = check_reads_against(run, ctr, reads, exp["outcomes"])
    n = batch.n
    for a, b in ((0, n // 2), (n // 2, n - 3), (n - 3, n)):
        ctr.submit(batch.slice(a, b))
    c = ctr.counters()
    assert c.pop("unsupported") == 0
    assert c == exp["counters"]
    prof = ctr.profile()
    assert prof["specialized_launches"] >= 5 and prof["generic_launches"] == 0, prof
    ctr.write_counts(str(tmp_path), "golden", merge=fl["merge"], enrich=fl["enrich"])
    assert_same_csv_set(read_csv_dir(str(tmp_path), "golden"), exp["files"])
    # a batch of another geometry (longer reads) falls back to the generic kernel of the same context
    wide = bc.Run(paths["fmt"], paths["samples"], paths["counted"], min_quality=fl["min_quality"], max_barcode=fl["max_barcode"],
                  max_sample=fl["max_sample"], max_constant=fl["max_constant"], max_read_len=run.max_read_len + 64)
    ctr.reset()
    ctr.reset_profile()
    ctr.submit(wide.pack([r[0] for r in reads], [r[1] for r in reads]))
    c = ctr.counters()
    c.pop("unsupported")
    assert c == exp["counters"] and ctr.profile()["generic_launches"] == 1


@pytest.mark.parametrize("case", golden_cases())
def test_golden_lean_csv_writer(case, tmp_path):
    """Large count tables are written by the streaming writer (text straight from the packed keys on all host threads);
    forced here for every golden case: same CSV set (it steps aside by itself when a merged file is due)."""
    exp, paths = load_golden(case)
    fl = exp["flags"]
    run = make_run(paths, fl)
    run.set_option("lean_writer_min_rows", 0)
    ctr = bc.Counter(run)
    reads = read_fastq(paths["fastq"])
    ctr.submit(run.pack([r[0] for r in reads], [r[1] for r in reads]))
    ctr.write_counts(str(tmp_path), "golden", merge=fl["merge"], enrich=fl["enrich"])
    assert_same_csv_set(read_csv_dir(str(tmp_path), "golden"), exp["files"])
    # without --merge-output every case goes through it, whatever its sample barcodes are
    out2 = tmp_path / "nomerge"
    out2.mkdir()
    ctr.write_counts(str(out2), "golden", merge=False, enrich=fl["enrich"])
    want = {fn: lines for fn, lines in exp["files"].items() if "_counts.all" not in fn}
    assert_same_csv_set(read_csv_dir(str(out2), "golden"), want)


@pytest.mark.parametrize("case", golden_cases())
def test_golden_fastq_ingest_small_batches(case, tmp_path):
    """bch_count_fastq (the read_fastq replacement) with batches far smaller than the file: pinned double
    buffering, staging reuse and table growth are all exercised."""
    exp, paths = load_golden(case)
    fl = exp["flags"]
    run = make_run(paths, fl)
    ctr = bc.Counter(run)
    total = ctr.count_fastq(paths["fastq"], threads=2, batch_reads=37)
    assert total == len(exp["outcomes"])
    c = ctr.counters()
    c.pop("unsupported")
    assert c == exp["counters"]
    ctr.write_counts(str(tmp_path), "golden", merge=fl["merge"], enrich=fl["enrich"])
    assert_same_csv_set(read_csv_dir(str(tmp_path), "golden"), exp["files"])


@pytest.mark.parametrize("case", golden_cases())
def test_golden_wire_batches(case, tmp_path):
    """Host batches in their transfer form (bc_submit_wire: lo / hi planes, N calls as a list, quality as 2- to 8-bit codes,
    expanded on the device) give the golden counters and CSV set, in every quality form that holds the batch."""
    exp, paths = load_golden(case)
    fl = exp["flags"]
    run = make_run(paths, fl)
    ctr = bc.Counter(run)
    reads = read_fastq(paths["fastq"])
    batch = run.pack([r[0] for r in reads], [r[1] for r in reads])
    n = batch.n
    auto = bc.WireBatch(batch, run.max_read_len)
    forms = [0] + ([b for b in (4, 6, 8) if b > auto.c.qual_bits] if run.quality_on else [])
    for bits in forms:
        ctr.reset()
        for a, b in ((0, n // 3), (n // 3, n - 1), (n - 1, n)):
            ctr.submit(bc.WireBatch(batch.slice(a, b), run.max_read_len, qual_bits=bits))
        c = ctr.counters()
        assert c.pop("unsupported") == 0
        assert c == exp["counters"], bits
    assert ctr.profile()["h2d_bytes"] > 0
    ctr.write_counts(str(tmp_path), "golden", merge=fl["merge"], enrich=fl["enrich"])
    assert_same_csv_set(read_csv_dir(str(tmp_path), "golden"), exp["files"])
    # a wider geometry than the run's (a batch that met a longer read) and the specialised kernel's context
    ctr2 = bc.Counter(run, flags=bc.BC_CFG_SPECIALIZE)
    wide = bc.Run(paths["fmt"], paths["samples"], paths["counted"], min_quality=fl["min_quality"], max_barcode=fl["max_barcode"],
                  max_sample=fl["max_sample"], max_constant=fl["max_constant"], max_read_len=run.max_read_len + 70)
    ctr2.submit(bc.WireBatch(batch.slice(0, n // 2), run.max_read_len))
    ctr2.submit(bc.WireBatch(wide.pack([r[0] for r in reads[n // 2:]], [r[1] for r in reads[n // 2:]]), wide.max_read_len))
    c = ctr2.counters()
    c.pop("unsupported")
    assert c == exp["counters"]
    prof = ctr2.profile()
    assert prof["specialized_launches"] == 1 and prof["generic_launches"] == 1, prof


@pytest.mark.parametrize("case", ["del3_umi", "example_q20", "crispr"])
def test_golden_fastq_ingest_plain_batches(case, tmp_path):
    """bch_count_fastq with the transfer form switched off (plain bc_batch arrays across PCIe): same result."""
    exp, paths = load_golden(case)
    fl = exp["flags"]
    run = make_run(paths, fl)
    run.set_option("wire_batches", 0)
    ctr = bc.Counter(run)
    assert ctr.count_fastq(paths["fastq"], threads=2, batch_reads=41) == len(exp["outcomes"])
    c = ctr.counters()
    c.pop("unsupported")
    assert c == exp["counters"]
    plain_bytes = ctr.profile()["h2d_bytes"]
    run.set_option("wire_batches", 1)
    ctr.reset()
    ctr.reset_profile()
    assert ctr.count_fastq(paths["fastq"], threads=2, batch_reads=41) == len(exp["outcomes"])
    c = ctr.counters()
    c.pop("unsupported")
    assert c == exp["counters"]
    assert ctr.profile()["h2d_bytes"] < plain_bytes


def test_ingest_wire_fallbacks(tmp_path):
    """A batch with a quality character beyond '_' crosses as plain bytes, a batch full of N calls sends its N plane whole:
    the ingest notices both by itself and the result equals the plain-batch path's."""
    exp, paths = load_golden("example_q20")
    fl = exp["flags"]
    rng = random.Random(9)
    reads = read_fastq(paths["fastq"])
    out = []
    for i, (s, q) in enumerate(reads):
        if i % 50 == 7:
            q = q[:5] + "~" + q[6:]  # Phred 93: long-read instruments
        if 100 <= i < 160:
            s = "".join("N" if rng.random() < 0.4 else ch for ch in s)
        out.append((s, q))
    fq = tmp_path / "fallbacks.fastq"
    with open(fq, "w") as f:
        for i, (s, q) in enumerate(out):
            f.write(f"@r{i}\n{s}\n+\n{q}\n")
    run = make_run(paths, fl)
    ctr = bc.Counter(run)
    got = {}
    for wire in (0, 1):
        run.set_option("wire_batches", wire)
        ctr.reset()
        assert ctr.count_fastq(str(fq), threads=2, batch_reads=20) == len(out)
        got[wire] = ctr.counters()
        st = ctr.ingest_stats()
        if wire:
            assert st["batches"] == (len(out) + 19) // 20 and 0 < st["batches_qual8"] < st["batches"] and 0 < st["batches_dense_n"] <= 4, st
    assert got[0] == got[1] and got[1]["matched"] > 0
    # and against the oracle
    orc = Oracle(paths["fmt"], paths["samples"], paths["counted"], min_quality=fl["min_quality"], outdir=str(tmp_path), prefix="o")
    for s, q in out:
        orc.process(s, q)
    want = orc.counters()
    g = dict(got[1])
    g.pop("unsupported")
    assert g == want


@pytest.mark.parametrize("case", ["example", "crispr", "del3_umi", "lineage_raw"])
@pytest.mark.parametrize("gz", [False, True, "bgzf"])
def test_cli_drop_in(case, gz, tmp_path):
    """The `barcode-count` binary with the reference's flags writes the reference's CSV set (plain FASTQ, two concatenated
    gzip members, and bgzip members inflated in parallel)."""
    exp, paths = load_golden(case)
    fl = exp["flags"]
    fastq = paths["fastq"]
    if gz == "bgzf":
        fastq = str(tmp_path / "reads.fastq.gz")
        with open(fastq, "wb") as f:
            f.write(bgzf_compress(open(paths["fastq"], "rb").read(), chunk=3000))
    elif gz:
        import gzip
        fastq = str(tmp_path / "reads.fastq.gz")
        data = open(paths["fastq"], "rb").read()
        half = data.find(b"\n@", len(data) // 2) + 1
        with open(fastq, "wb") as f:  # two concatenated gzip members (MultiGzDecoder, input.rs:63)
            f.write(gzip.compress(data[:half]))
            f.write(gzip.compress(data[half:]))
    out = tmp_path / "out"
    out.mkdir()
    cmd = [bc.CLI_PATH, "--fastq", fastq, "--sequence-format", paths["fmt"], "--output-dir", str(out), "--prefix", "golden",
           "--threads", "2", "--min-quality", str(fl["min_quality"])]
    if paths["samples"]:
        cmd += ["--sample-barcodes", paths["samples"]]
    if paths["counted"]:
        cmd += ["--counted-barcodes", paths["counted"]]
    if fl["merge"]:
        cmd.append("--merge-output")
    if fl["enrich"]:
        cmd.append("--enrich")
    for flag, key in (("--max-errors-counted-barcode", "max_barcode"), ("--max-errors-sample", "max_sample"),
                      ("--max-errors-constant", "max_constant")):
        if fl[key] is not None:
            cmd += [flag, str(fl[key])]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert_same_csv_set(read_csv_dir(str(out), "golden"), exp["files"])
    c = exp["counters"]
    assert f"Correctly matched sequences: {c['matched']:,}" in r.stdout
    assert f"Constant region mismatches:  {c['constant_region']:,}" in r.stdout
    assert f"Duplicates:                  {c['duplicates']:,}" in r.stdout
    # the run record the reference appends (output.rs:488-576): same sections, one line per file written
    stats = (out / "golden_barcode_stats.txt").read_text()
    for section in ("-TIME INFORMATION-", "-INPUT FILES-", "-FORMAT-", "-BARCODE INFO-", "-RESULTS-", "-OUTPUT FILES-"):
        assert section in stats
    assert f"Correctly matched sequences: {c['matched']:,}" in stats
    assert stats.count("File & barcodes counted: ") == len(exp["files"])
    for fn, lines in exp["files"].items():
        assert f"File & barcodes counted: {fn}\t{len(lines) - 1:,}" in stats


@pytest.mark.parametrize("case", ["del3_umi", "crispr", "example", "sample_raw_two"])
def test_cli_several_contexts_and_clap_syntax(case, tmp_path):
    """`--devices` shards the reads over several contexts inside one process (here twice the same GPU, so that a
    one-GPU box covers it): batches alternate between them and the result is merged by one record exchange (hashed keys)
    or by adding the dense tables (crispr) — the CSV set must still be the reference's.  The command line is clap's:
    `--name=value`, bundled short flags, attached short values (arguments.rs:27-124)."""
    exp, paths = load_golden(case)
    fl = exp["flags"]
    out = tmp_path / "out"
    out.mkdir()
    cmd = [bc.CLI_PATH, f"--fastq={paths['fastq']}", "-q", paths["fmt"], f"-o={out}", "-pgolden", "-t3", "--devices=0,0,0",
           "--batch-reads=64", f"--min-quality={fl['min_quality']}"]
    if paths["samples"]:
        cmd += ["-s", paths["samples"]]
    if paths["counted"]:
        cmd += [f"--counted-barcodes={paths['counted']}"]
    short = ("m" if fl["merge"] else "") + ("e" if fl["enrich"] else "")
    if short:
        cmd.append("-" + short)
    for flag, key in (("--max-errors-counted-barcode", "max_barcode"), ("--max-errors-sample", "max_sample"),
                      ("--max-errors-constant", "max_constant")):
        if fl[key] is not None:
            cmd += [f"{flag}={fl[key]}"]
    r = subprocess.run(cmd, capture_output=True)  # bytes: text mode would translate the carriage returns
    assert r.returncode == 0, r.stderr
    stdout = r.stdout.decode()
    assert_same_csv_set(read_csv_dir(str(out), "golden"), exp["files"])
    c = exp["counters"]
    n_reads = len(exp["outcomes"])
    assert f"Total sequences:             {n_reads:,}\r\n" in stdout  # input.rs:151-158: rewritten in place, then a newline
    assert f"Correctly matched sequences: {c['matched']:,}" in stdout
    assert f"Duplicates:                  {c['duplicates']:,}" in stdout
    assert f"Low quality barcodes:        {c['low_quality']:,}" in stdout
    for fn, lines in exp["files"].items():  # output.rs:143-165, 355-359: file name, then its number of barcode rows
        end = "\n" if "_counts.all" in fn else "\r\n"
        assert f"{fn}\nBarcodes counted: {len(lines) - 1:,}{end}" in stdout, fn


def test_cli_gzip_total_has_the_reference_phantom_read(tmp_path):
    """Q21: on the gzip path the reference's reader counts one more 'read' at the end of the stream (input.rs:69-73, 129-133)
    and warns about short gzip files (output.rs:566-571); the counters and CSVs are those of the real reads."""
    import gzip
    exp, paths = load_golden("example")
    fastq = str(tmp_path / "reads.fastq.gz")
    with open(fastq, "wb") as f:
        f.write(gzip.compress(open(paths["fastq"], "rb").read()))
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([bc.CLI_PATH, "-f", fastq, "-q", paths["fmt"], "-s", paths["samples"], "-c", paths["counted"], "-o", str(out),
                        "-p", "golden", "-me"], capture_output=True)
    assert r.returncode == 0, r.stderr
    stdout = r.stdout.decode()
    assert stdout.startswith("-FORMAT-")
    assert "If this program stops reading before the expected number of sequencing reads" in stdout
    assert f"Total sequences:             {len(exp['outcomes']) + 1:,}\r\n" in stdout
    assert "WARNING: The program may have stopped early with the gzipped file." in stdout
    assert_same_csv_set(read_csv_dir(str(out), "golden"), exp["files"])


def test_reads_of_any_length_and_ragged_quality_lines(tmp_path):
    """The reference has no read-length limit and zips a quality line of any length with the scheme (input.rs:115-148,
    parse.rs:287-313, 338-343).  A FASTQ whose later reads are far longer than the first ones (the CLI sizes its batches
    from the first reads), up to the build's 1024-base limit, with quality lines shorter and longer than their sequences:
    same counters and CSV set as the oracle; nothing aborts; reads beyond 1024 bases are counted as unsupported."""
    import mirror
    exp, paths = load_golden("del3_umi")
    rng = random.Random(77)
    base = read_fastq(paths["fastq"])
    reads = list(base[:150])
    for k, (s, q) in enumerate(base[150:400]):
        extra = rng.choice([0, 30, 200, 500, 1024 - len(s)])
        pre = mirror.rand_seq(rng, rng.randint(0, extra))
        s2 = pre + s + mirror.rand_seq(rng, extra - len(pre))
        q2 = "".join(chr(33 + rng.randint(15, 40)) for _ in range(len(s2)))
        r = rng.random()
        if r < 0.25:
            q2 = q2[:rng.randint(0, len(q2))]
        elif r < 0.35:
            q2 += "IIII"
        reads.append((s2, q2))
    too_long = [("ACGT" * 300, "I" * 1200), (base[0][0] + "A" * 1100, "I" * (len(base[0][0]) + 1100))]
    fq = tmp_path / "ragged.fastq"
    with open(fq, "w") as f:
        for i, (s, q) in enumerate(reads + too_long):
            f.write(f"@r{i}\n{s}\n+\n{q}\n")
    o_dir, g_dir = tmp_path / "o", tmp_path / "g"
    o_dir.mkdir()
    g_dir.mkdir()
    fl = exp["flags"]
    orc = Oracle(paths["fmt"], paths["samples"], paths["counted"], min_quality=fl["min_quality"], merge=fl["merge"], enrich=fl["enrich"],
                 outdir=str(o_dir), prefix="p")
    for s, q in reads:
        orc.process(s, q)
    orc.write_files()
    want = orc.counters()
    assert want["low_quality"] > 0 and want["matched"] > 100
    # the library's ingest with a run sized for the short reads only
    run = bc.Run(paths["fmt"], paths["samples"], paths["counted"], min_quality=fl["min_quality"], max_read_len=len(base[0][0]))
    ctr = bc.Counter(run)
    for wire in (0, 1):  # plain bc_batch arrays, then (the default) the transfer form
        run.set_option("wire_batches", wire)
        ctr.reset()
        assert ctr.count_fastq(str(fq), threads=3, batch_reads=64) == len(reads) + 2
        got = ctr.counters()
        assert got.pop("unsupported") == 2 and got == want, wire
    ctr.write_counts(str(g_dir), "p", merge=fl["merge"], enrich=fl["enrich"])
    assert_same_csv_set(read_csv_dir(str(g_dir), "p"), read_csv_dir(str(o_dir), "p"))
    # and the command line, which probes the first reads for its default geometry
    c_dir = tmp_path / "c"
    c_dir.mkdir()
    r = subprocess.run([bc.CLI_PATH, "-f", str(fq), "-q", paths["fmt"], "-c", paths["counted"], "-o", str(c_dir), "-p", "p",
                        f"--min-quality={fl['min_quality']}", "--batch-reads=100"] + (["-e"] if fl["enrich"] else []),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert_same_csv_set(read_csv_dir(str(c_dir), "p"), read_csv_dir(str(o_dir), "p"))
    assert f"Correctly matched sequences: {want['matched']:,}" in r.stdout and "2 reads hold characters outside" in r.stderr


# ---- hand-derived vectors on the reference's example files ------------------------------
SAMPLE, C1, B1, C2, B2, C3, B3, C4, UMI, C5 = ("AGCATACGGG", "AGCTACGAATCG", "CAGAGA", "TGGA", "ATGAAA", "TGGA",
                                               "GATAGC", "ACTAGAT", "ACGTACGT", "TAGA")
G1 = SAMPLE + C1 + B1 + C2 + B2 + C3 + B3 + C4 + UMI + C5
ROW = "CAGAGAC,ATGAAAT,GATAGCT"
EX = os.path.join(GOLDEN, "example")


def example(min_quality=0.0):
    run = bc.Run(os.path.join(EX, "scheme.txt"), os.path.join(EX, "samples.csv"), os.path.join(EX, "barcodes.csv"),
                 min_quality=min_quality)
    return run, bc.Counter(run)


def one(run, ctr, seq, qual=None):
    qual = qual or "I" * len(seq)
    got = ctr.decode_only(run.pack([seq], [qual]))
    status = bc.STATUS_NAMES[got["status"][0]]
    strings = decoded_strings(run, ctr, got["key_lo"][0], got["key_hi"][0]) if status == "matched" else None
    return status, int(got["offset"][0]), bool(got["repaired"][0]), strings


def test_hand_vectors():
    run, ctr = example()
    assert one(run, ctr, G1 + "A") == ("matched", 0, False, ("AGCATAC", ROW, UMI))  # G1
    assert one(run, ctr, G1.replace(B1, "CAGTGA", 1) + "A")[3][1] == ROW  # G3: d=1 corrected
    assert one(run, ctr, G1.replace(B1, "CTGTGA", 1) + "A")[0] == "barcode"  # G4: d=2 rejected
    assert one(run, ctr, "AACATAC" + G1[7:] + "A")[0] == "sample_barcode"  # G5: tie
    bad = G1.replace(C1, "AGCTACGTATCG", 1)
    assert one(run, ctr, "TTTTT" + bad + "A") == ("matched", 5, True, ("AGCATAC", ROW, UMI))  # G6
    assert one(run, ctr, "TTTTT" + bad)[0] == "constant_region"  # G7 (Q3): last offset not repaired
    assert one(run, ctr, "TTTTT" + G1)[:3] == ("matched", 5, False)  # ... but matched exactly there
    assert one(run, ctr, bad)[0] == "constant_region"  # R == L
    assert one(run, ctr, G1.replace(C1, "AGCTANGAATCG", 1) + "A") == ("matched", 0, True, ("AGCATAC", ROW, UMI))  # G8
    assert one(run, ctr, G1[:40])[0] == "constant_region"  # shorter than the scheme (Q4 fenced off)
    assert one(run, ctr, G1.replace("G", "g", 1) + "A")[0] == "unsupported"
    # G2: UMI duplicate through the counting path
    b = run.pack([G1 + "A", G1 + "A", G1 + "C"])
    ctr.submit(b)
    c = ctr.counters()
    assert (c["matched"], c["duplicates"]) == (1, 2)
    rows = ctr.finish()
    assert list(rows["count"]) == [1]


def test_quality_window_after_repair_q6():
    read = "TTTTT" + G1.replace(C1, "AGCTACGTATCG", 1) + "A"
    run, ctr = example(min_quality=20.0)
    assert one(run, ctr, read, "#" * 10 + "I" * (len(read) - 10))[0] == "low_quality"
    assert one(run, ctr, read, "I" * 5 + "#" * 4 + "I" * (len(read) - 9))[0] == "matched"


def test_empty_batch_and_empty_table(tmp_path):
    run, ctr = example()
    ctr.submit(run.pack([]))
    assert sum(ctr.counters().values()) == 0
    assert ctr.finish()["count"].size == 0
    names = ctr.write_counts(str(tmp_path), "t", merge=True, enrich=True)
    files = read_csv_dir(str(tmp_path), "t")
    assert files["t_Sample_name_1_counts.csv"] == ["Barcode_1,Barcode_2,Barcode_3,Count"]  # Q16
    assert "t_counts.all.csv" in names


# ---- seeded random schemes and reads against the oracle ---------------------------------------------------------
def rand_dna(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, p_sub, p_n):
    out = []
    for ch in s:
        u = rng.random()
        if u < p_n:
            out.append("N")
        elif u < p_n + p_sub:
            out.append(rng.choice("ACGT"))
        else:
            out.append(ch)
    return "".join(out)


def random_case(rng, tmp_path, idx):
    """A random scheme with its conversion files and reads; exercises ties, N, raw slots, mixed lengths."""
    pieces, kinds = [], []
    n_counted = rng.randint(1, 3)
    layout = ["S"] * rng.randint(0, 1) + ["B"] * n_counted + ["R"] * rng.randint(0, 1)
    rng.shuffle(layout)
    scheme = ""
    slot_lens = []
    for k, kind in enumerate(layout):
        if k > 0 or rng.random() < 0.8:
            if k == 0 or rng.random() < 0.85:  # sometimes two barcodes touch
                c = rand_dna(rng, rng.randint(3, 14))
                scheme += c
        ln = rng.choice([4, 5, 6, 8, 10, 12, 16, 20] if kind != "R" else [4, 6, 8, 10, 12])
        slot_lens.append(ln)
        scheme += {"S": "[%d]", "B": "{%d}", "R": "(%d)"}[kind] % ln
    if rng.random() < 0.8:
        scheme += rand_dna(rng, rng.randint(3, 12))
    if not any(ch in "ACGT" for ch in scheme):
        scheme = "ACGTTGCA" + scheme
    fmt = tmp_path / f"scheme{idx}.txt"
    fmt.write_text(scheme + "\n")
    # reference sets: close neighbours on purpose so that ties and corrections happen
    def make_set(ln, n):
        base = [rand_dna(rng, ln) for _ in range(max(1, n // 3))]
        out = set(base)
        while len(out) < n:
            out.add(mutate(rng, rng.choice(base), 0.25, 0.0))
        return sorted(out)
    samples_path = counted_path = None
    sample_set, counted_sets = None, None
    if "S" in layout and rng.random() < 0.7:
        ln = slot_lens[layout.index("S")]
        sample_set = make_set(ln, rng.randint(2, 12))
        samples_path = tmp_path / f"samples{idx}.csv"
        samples_path.write_text("Barcode,Sample_ID\n" + "".join(f"{d},S{j}\n" for j, d in enumerate(sample_set)))
    if rng.random() < 0.75:
        counted_sets = []
        lines = ["Barcode,Barcode_ID,Barcode_Number"]
        for k, s in enumerate(i for i, kd in enumerate(layout) if kd == "B"):
            st = make_set(slot_lens[s], rng.randint(2, 40))
            counted_sets.append(st)
            lines += [f"{d},B{k}_{j},{k + 1}" for j, d in enumerate(st)]
        counted_path = tmp_path / f"counted{idx}.csv"
        counted_path.write_text("\n".join(lines) + "\n")
    # template instance builder
    import re
    tokens = re.findall(r"\{\d+\}|\[\d+\]|\(\d+\)|[ACGT]+", scheme)
    L = sum(int(t[1:-1]) if t[0] in "{[(" else len(t) for t in tokens)
    reads = []
    R = L + rng.randint(1, 40)
    for _ in range(600):
        body, k = "", 0
        for t in tokens:
            if t[0] == "[":
                body += rng.choice(sample_set) if sample_set and rng.random() < 0.9 else rand_dna(rng, int(t[1:-1]))
            elif t[0] == "{":
                st = counted_sets[k] if counted_sets else None
                body += rng.choice(st) if st and rng.random() < 0.9 else rand_dna(rng, int(t[1:-1]))
                k += 1
            elif t[0] == "(":
                body += rand_dna(rng, int(t[1:-1]), "ACGGT")
            else:
                body += t
        u = rng.random()
        if u < 0.08:
            seq = rand_dna(rng, R)
        else:
            start = rng.randint(0, R - L)
            seq = rand_dna(rng, start) + mutate(rng, body, rng.choice([0.0, 0.01, 0.04]), rng.choice([0.0, 0.01])) + rand_dna(rng, R - L - start)
            if rng.random() < 0.05:
                seq = seq[:rng.randint(max(1, L - 5), R)]  # ragged lengths, some shorter than the scheme
        qual = "".join(chr(33 + max(2, min(40, int(rng.gauss(30, 8))))) for _ in seq)
        reads.append((seq, qual))
    return dict(fmt=str(fmt), samples=str(samples_path) if samples_path else None,
                counted=str(counted_path) if counted_path else None), reads, len(layout) - layout.count("R") - layout.count("S")


# BC_TEST_SEEDS=N widens the campaign (one-off soak runs); the committed default keeps the suite short
@pytest.mark.parametrize("seed", range(int(os.environ.get("BC_TEST_SEEDS", "24"))))
def test_random_schemes_against_oracle(seed, tmp_path):
    rng = random.Random(1000 + seed)
    paths, reads, n_counted = random_case(rng, tmp_path, seed)
    min_q = rng.choice([0.0, 0.0, 20.0, 27.5])
    caps = dict(max_barcode=rng.choice([None, None, 0, 2]), max_sample=rng.choice([None, None, 1]),
                max_constant=rng.choice([None, None, 1, 4]))
    merge, enrich = rng.random() < 0.5, rng.random() < 0.6
    o_dir, g_dir = tmp_path / "o", tmp_path / "g"
    o_dir.mkdir()
    g_dir.mkdir()
    orc = Oracle(paths["fmt"], paths["samples"], paths["counted"], min_quality=min_q, merge=merge, enrich=enrich,
                 outdir=str(o_dir), prefix="p", **caps)
    # the fenced-off reference behaviour (Q4: reads shorter than the scheme underflow in the reference) is
    # defined by the oracle as constant_region_error, like the GPU path
    outcomes = [orc.process(s, q) for s, q in reads]
    run = bc.Run(paths["fmt"], paths["samples"], paths["counted"], min_quality=min_q, **caps)
    try:
        ctr = bc.Counter(run)
    except bc.BcError as e:  # documented limit: raw barcodes of more than 42 bases in all do not fit a key
        if "packed key needs" in str(e):
            pytest.skip(str(e))
        raise
    batch = check_reads_against(run, ctr, reads, outcomes)
    if seed % 2:  # every other scheme crosses PCIe in the transfer form, in two batches
        half = batch.n // 2
        ctr.submit(bc.WireBatch(batch.slice(0, half), run.max_read_len))
        ctr.submit(bc.WireBatch(batch.slice(half, batch.n), run.max_read_len))
    else:
        ctr.submit(batch)
    c = ctr.counters()
    assert c.pop("unsupported") == 0
    assert c == orc.counters()
    orc.write_files()
    ctr.write_counts(str(g_dir), "p", merge=merge, enrich=enrich and n_counted >= 2)
    assert_same_csv_set(read_csv_dir(str(g_dir), "p"), read_csv_dir(str(o_dir), "p"))


def test_device_resident_batch_and_caller_stream():
    import torch
    exp, paths = load_golden("del3_umi")
    run = make_run(paths, exp["flags"])
    ctr = bc.Counter(run)
    stream = torch.cuda.Stream()
    ctr.set_stream(stream.cuda_stream)
    reads = read_fastq(paths["fastq"])
    dev = run.pack([r[0] for r in reads], [r[1] for r in reads]).to_device()
    with torch.cuda.stream(stream):
        ctr.submit(dev)
    c = ctr.counters()
    c.pop("unsupported")
    assert c == exp["counters"]
    assert ctr.profile()["h2d_bytes"] == 0


def test_table_growth_and_many_keys():
    """Raw 30-mer keys + UMI: 126-bit records, tables start tiny (expected_reads=0 -> growth by rehash)."""
    exp, paths = load_golden("lineage_raw")
    run = make_run(paths, exp["flags"])
    rng = random.Random(7)
    reads = read_fastq(paths["fastq"])
    tmpl = [r for r, o in zip(reads, exp["outcomes"]) if o["status"] == "matched" and not o["repaired"]][0][0]
    off = [o for o in exp["outcomes"] if o["status"] == "matched" and not o["repaired"]][0]["offset"]
    seqs, want = [], {}
    n = 6000
    for i in range(n):
        key, umi = rand_dna(rng, 30), rand_dna(rng, 12)
        s = tmpl[:off + 20] + key + tmpl[off + 50:off + 60] + umi + tmpl[off + 72:]
        reps = rng.randint(1, 3)
        for _ in range(reps):
            seqs.append(s)
        want[key] = want.get(key, 0) + 1
    ctr = bc.Counter(run, expected_reads=0)
    batch = run.pack(seqs)
    for a in range(0, batch.n, 1000):
        ctr.submit(batch.slice(a, min(batch.n, a + 1000)))
    c = ctr.counters()
    assert c["matched"] == n and c["duplicates"] == len(seqs) - n
    rows = ctr.finish()
    got = {}
    for lo, hi, cnt in zip(rows["key_lo"], rows["key_hi"], rows["count"]):
        got[ctr.key_decode(lo, hi)[0]] = int(cnt)
    assert got == want


def _counter(run, mode, **kw):
    """deferred: records appended, partitioned shared-memory flush (default); two_stage: the flush's hot-key path for every
    key; global: same records flushed through the global-memory tables (the overflow fallback); inline: tables updated
    read by read inside k_decode (bc_config.flags = BC_CFG_INLINE_COUNT)."""
    ctr = bc.Counter(run, flags=bc.BC_CFG_INLINE_COUNT if mode == "inline" else 0, **kw)
    if mode == "global":
        ctr.set_option("flush_global", 1)
    elif mode == "two_stage":
        ctr.set_option("flush_two_stage", 1)
    return ctr


@pytest.mark.parametrize("mode", ["deferred", "two_stage", "global", "inline"])
def test_counting_modes_skewed_keys(mode):
    """What the deferred counting has to survive: one (key, UMI) pair repeated far beyond a partition's table size, one hot
    key with tens of thousands of distinct UMIs, many singleton keys — in several submits, with counters read (= a flush)
    in between.  Expected counts are computed in Python from the construction."""
    exp, paths = load_golden("lineage_raw")
    run = make_run(paths, exp["flags"])
    rng = random.Random(11)
    reads = read_fastq(paths["fastq"])
    tmpl = [r for r, o in zip(reads, exp["outcomes"]) if o["status"] == "matched" and not o["repaired"]][0][0]
    off = [o for o in exp["outcomes"] if o["status"] == "matched" and not o["repaired"]][0]["offset"]

    def read(key, umi):
        return tmpl[:off + 20] + key + tmpl[off + 50:off + 60] + umi + tmpl[off + 72:]

    seqs, pairs = [], set()
    hot, rep_key, rep_umi = rand_dna(rng, 30), rand_dna(rng, 30), rand_dna(rng, 12)
    seqs += [read(rep_key, rep_umi)] * 20000
    pairs.add((rep_key, rep_umi))
    for _ in range(30000):
        u = rand_dna(rng, 12)
        seqs.append(read(hot, u))
        pairs.add((hot, u))
    for _ in range(9000):
        k, u = rand_dna(rng, 30), rand_dna(rng, 12)
        seqs.append(read(k, u))
        pairs.add((k, u))
    seqs.append("ACGT" * 25)  # no scheme in it
    rng.shuffle(seqs)
    want = {}
    for k, _ in pairs:
        want[k] = want.get(k, 0) + 1
    ctr = _counter(run, mode, expected_reads=0)
    assert ctr.profile()["deferred_count"] == (0 if mode == "inline" else 1)
    batch = run.pack(seqs)
    cuts = [0, 7000, 7001, 40000, batch.n]
    for a, b in zip(cuts, cuts[1:]):
        ctr.submit(batch.slice(a, b))
        c = ctr.counters()  # flushes in the deferred modes; later submits must still add up
        assert sum(c.values()) == b
    assert c["matched"] == len(pairs) and c["duplicates"] == len(seqs) - 1 - len(pairs) and c["constant_region"] == 1
    prof = ctr.profile()
    assert prof["flushed_global"] == (1 if mode == "global" else 0)
    # the hot key (30000 pairs) does not fit one partition: the default flush has to notice and take two stages
    assert prof["flush_stages"] == (2 if mode in ("deferred", "two_stage") else 0)
    rows = ctr.finish()
    got = {}
    for lo, hi, cnt in zip(rows["key_lo"], rows["key_hi"], rows["count"]):
        k = ctr.key_decode(lo, hi)[0]
        assert k not in got
        got[k] = int(cnt)
    assert got == want
    ctr.reset()
    ctr.submit(batch.slice(0, 5000))
    assert sum(ctr.counters().values()) == 5000


@pytest.mark.parametrize("mode", ["two_stage", "global", "inline"])
@pytest.mark.parametrize("case", ["del3_umi", "example", "sample_raw_two", "example_q20"])
def test_golden_csv_other_counting_modes(case, mode, tmp_path):
    """The golden CSV sets through the non-default counting paths (the default is covered by test_golden_per_read_and_csv)."""
    exp, paths = load_golden(case)
    fl = exp["flags"]
    run = make_run(paths, fl)
    ctr = _counter(run, mode)
    reads = read_fastq(paths["fastq"])
    ctr.submit(run.pack([r[0] for r in reads], [r[1] for r in reads]))
    c = ctr.counters()
    c.pop("unsupported")
    assert c == exp["counters"]
    ctr.write_counts(str(tmp_path), "golden", merge=fl["merge"], enrich=fl["enrich"])
    assert_same_csv_set(read_csv_dir(str(tmp_path), "golden"), exp["files"])


@pytest.mark.parametrize("blen,max_err,n_ref", [(20, None, 700), (20, 1, 300), (20, 7, 400), (16, 2, 5000), (12, None, 900),
                                                (24, 5, 260), (11, 0, 300), (20, None, 100)])
def test_long_barcode_search_paths(blen, max_err, n_ref, tmp_path):
    """Barcodes longer than the direct-table limit: exact hash, half index (distance 1), block index (distance 2..cap,
    N wildcards expanded) and the whole-set fallback (3+ N, small sets) must all reproduce fix_error exactly —
    including ties at the minimum, which the engineered near-duplicate references provoke."""
    rng = random.Random(blen * 1000 + n_ref + (max_err or 0))
    refs = set()
    while len(refs) < n_ref * 2 // 3:
        refs.add(rand_dna(rng, blen))
    base = sorted(refs)
    while len(refs) < n_ref:  # close neighbours: distance 1..3 from an existing reference
        r = list(rng.choice(base))
        for pos in rng.sample(range(blen), rng.randint(1, 3)):
            r[pos] = rng.choice("ACGT")
        refs.add("".join(r))
    # a crowded half: many references sharing their first half
    head = rand_dna(rng, blen // 2)
    for _ in range(40):
        refs.add(head + rand_dna(rng, blen - blen // 2))
    refs = sorted(refs)
    fmt = tmp_path / "scheme.txt"
    fmt.write_text("ACGTACGTACGG{%d}TTGACAGTCA\n" % blen)
    counted = tmp_path / "counted.csv"
    counted.write_text("Barcode,ID,N\n" + "".join(f"{d},id{i},1\n" for i, d in enumerate(refs)))
    reads = []
    for i in range(3000):
        b = list(rng.choice(refs))
        for pos in rng.sample(range(blen), rng.choice([0, 1, 1, 1, 2, 2, 3, 4, 5, 6])):
            b[pos] = rng.choice("ACGT")
        for pos in rng.sample(range(blen), rng.choice([0, 0, 0, 1, 1, 2, 3, 4, 5])):
            b[pos] = "N"
        seq = rand_dna(rng, rng.randint(0, 6)) + "ACGTACGTACGG" + "".join(b) + "TTGACAGTCA" + rand_dna(rng, rng.randint(1, 6))
        reads.append((seq, "I" * len(seq)))
    orc = Oracle(str(fmt), None, str(counted), max_barcode=max_err)
    outcomes = [orc.process(s, q) for s, q in reads]
    assert sum(o["status"] == "matched" for o in outcomes) > 100
    run = bc.Run(str(fmt), None, str(counted), max_barcode=max_err)
    ctr = bc.Counter(run)
    batch = check_reads_against(run, ctr, reads, outcomes)
    ctr.submit(batch)
    c = ctr.counters()
    assert c.pop("unsupported") == 0 and c == orc.counters()
    assert ctr.profile()["launches"]["scan"] >= 1


def test_import_rows_merge_without_random_barcode():
    """The final-merge path of schemes without a random barcode whose key space is too large for a dense table (raw
    keys): two contexts count half of the reads each, the rows of one are imported into the other (bc_export_rows ->
    bc_import_rows, what ranks do at the end of a multi-GPU job) and the merged rows must equal one context over all reads."""
    from ngs_barcode_count_b200.multi import dev_tensor
    exp, paths = load_golden("sample_raw_two")
    run = make_run(paths, exp["flags"])
    reads = read_fastq(paths["fastq"]) * 3  # repeats: counts above one
    batch = run.pack([r[0] for r in reads], [r[1] for r in reads])

    def rows_of(ctr):
        rows = ctr.finish()
        return sorted(zip([int(x) for x in rows["key_hi"]], [int(x) for x in rows["key_lo"]], [int(x) for x in rows["count"]]))

    whole = bc.Counter(run)
    assert whole.profile()["deferred_count"] == 1 and whole.profile()["dense_table"] == 0
    whole.submit(batch)
    want = rows_of(whole)
    half = batch.n // 2
    a, b = bc.Counter(run), bc.Counter(run)
    a.submit(batch.slice(0, half))
    b.submit(batch.slice(half, batch.n))
    assert rows_of(a) != want  # flushed once already: the import below has to re-flush with the extra rows
    lo, hi, cnt, n = b.export_rows()
    t_lo, t_cnt = dev_tensor(lo, n, "cuda:0"), dev_tensor(cnt, n, "cuda:0")
    t_hi = dev_tensor(hi, n, "cuda:0") if hi else None
    a.import_rows(t_lo, t_hi, t_cnt, n)
    assert rows_of(a) == want
    ca, cb, cw = a.counters(), b.counters(), whole.counters()
    assert ca["matched"] + cb["matched"] == cw["matched"] and sum(r[2] for r in want) == cw["matched"]


@pytest.mark.parametrize("bulk", [False, True])
@pytest.mark.parametrize("n_ranks", [2, 3])
@pytest.mark.parametrize("case", ["del3_umi", "lineage_raw", "example", "sample_raw_two"])
def test_exchange_between_contexts_on_one_gpu(case, n_ranks, bulk):
    """The multi-GPU exchange (bc_exchange_*: records scattered into their owner's receive buffer by the partitioning
    kernel, every owner de-duplicating what it received) with the ranks being contexts on ONE GPU: the union of the
    owners' rows, and the sum of their counters, must be exactly what a single context gives on all the reads.  Covers
    the N > 1 device code on a one-GPU box (tests/test_multi_gpu.py runs it over NVLink when the box has more)."""
    from ngs_barcode_count_b200.multi import exchange_plan
    exp, paths = load_golden(case)
    run = make_run(paths, exp["flags"])
    reads = read_fastq(paths["fastq"])
    batch = run.pack([r[0] for r in reads], [r[1] for r in reads])

    def rows_of(ctr):
        rows = ctr.finish()
        return sorted(zip([int(x) for x in rows["key_hi"]], [int(x) for x in rows["key_lo"]], [int(x) for x in rows["count"]]))

    direct = bc.Counter(run)
    direct.submit(batch)
    want_rows, want_c = rows_of(direct), direct.counters()
    if not direct.profile()["deferred_count"]:
        pytest.skip("dense count table: ranks merge with an all-reduce, not an exchange")
    ranks = [bc.Counter(run) for _ in range(n_ranks)]
    for r, c in enumerate(ranks):
        c.exchange_open(n_ranks, r, batch.n + 16)
    for c in ranks:
        c.exchange_connect_local(ranks)
        if bulk:  # the records leave after the last batch instead of batch by batch
            c.set_option("exchange_bulk", 1)
    cuts = [batch.n * r // n_ranks for r in range(n_ranks + 1)]
    for rep in range(5):  # several jobs: a reset in between must leave the exchange usable (the receive buffers alternate)
        if rep == 3:  # back to roomy buffers; in jobs 3 and 4 rank 0 decodes nothing at all (its shard is empty) but still owns keys
            for c in ranks:
                c.exchange_disconnect()
            for r, c in enumerate(ranks):
                c.exchange_open(n_ranks, r, batch.n + 16)
            for c in ranks:
                c.exchange_connect_local(ranks)
            cuts = [0] + [batch.n * r // (n_ranks - 1) for r in range(n_ranks)]
        if rep == 2:  # buffers far too small: the streamed records do not fit, the job's exchange starts over in bulk
            for c in ranks:
                c.exchange_disconnect()
            for r, c in enumerate(ranks):
                c.exchange_open(n_ranks, r, max(1, batch.n // (4 * n_ranks)))
            for c in ranks:
                c.exchange_connect_local(ranks)
        for r, c in enumerate(ranks):
            c.reset()
            if cuts[r] == cuts[r + 1]:
                continue
            mid = (cuts[r] + cuts[r + 1]) // 2
            c.submit(batch.slice(cuts[r], mid))  # two batches per rank: the streamed exchange moves them one by one
            c.submit(batch.slice(mid, cuts[r + 1]))
        with pytest.raises(bc.BcError):
            ranks[0].counters()  # a rank of a multi-GPU job has no counters before the exchange
        matrix = [c.exchange_count(n_ranks) for c in ranks]
        assert sum(map(sum, matrix)) == want_c["matched"] + want_c["duplicates"]
        if rep == 2:
            need = exchange_plan(matrix, 0)[2]
            assert need > batch.n // (4 * n_ranks)
            if not bulk:  # the rank(s) whose run passed the end of an owner's buffer refuse to go on
                refused = 0
                for r, c in enumerate(ranks):
                    try:
                        c.exchange_scatter(exchange_plan(matrix, r)[0])
                    except bc.BcError:
                        refused += 1
                assert refused >= 1
            for c in ranks:
                c.exchange_disconnect()
            for r, c in enumerate(ranks):
                c.exchange_open(n_ranks, r, need)
            for c in ranks:
                c.exchange_connect_local(ranks)
            assert [c.exchange_count(n_ranks) for c in ranks] == matrix
        for r, c in enumerate(ranks):
            c.exchange_scatter(exchange_plan(matrix, r)[0])
        for c in ranks:
            c.sync()  # the barrier of the in-process form
        for r, c in enumerate(ranks):
            c.exchange_finish(exchange_plan(matrix, r)[1])
        got_rows, got_c = [], {k: 0 for k in want_c}
        for c in ranks:
            got_rows += rows_of(c)
            for k, v in c.counters().items():
                got_c[k] += v
        assert got_c == want_c
        assert sorted(got_rows) == want_rows
        keys = [(h, l) for h, l, _ in got_rows]
        assert len(set(keys)) == len(keys)  # owners hold disjoint key sets
