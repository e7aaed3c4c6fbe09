"""The synthetic workloads (bench tooling): barcode-set construction and host generator on the CPU, and on the GPU the
device generator against the host one plus oracle parity of the whole job on every BASELINE.json workload."""
import numpy as np
import pytest

import ngs_barcode_count_b200 as bc
from ngs_barcode_count_b200 import synth
from helpers import Oracle, assert_same_csv_set, read_csv_dir


def test_hamming_sets_have_distance_3():
    rng = np.random.default_rng(5)
    w = synth.hamming_code_words(8, 1024, rng)
    assert len({bytes(x) for x in w}) == 1024
    d = (w[:, None, :] != w[None, :, :]).sum(-1)
    np.fill_diagonal(d, 99)
    assert d.min() == 3
    g = synth.hamming_code_words(20, 80_000, rng)
    assert len({bytes(x) for x in g}) == 80_000
    sub = g[rng.permutation(80_000)[:1500]]
    d = (sub[:, None, :] != sub[None, :, :]).sum(-1)
    np.fill_diagonal(d, 99)
    assert d.min() >= 3


@pytest.mark.parametrize("name", ["example", "crispr", "del3", "lineage"])
def test_host_generator_is_deterministic_and_well_formed(name, tmp_path):
    wl = synth.Workload(name, str(tmp_path / name), reads=10_000)
    a = wl.generate_fastq(100, 300, threads=1).tobytes()
    b = wl.generate_fastq(0, 1000, threads=3).tobytes()
    assert len(a) == wl.fastq_bytes(100, 300) and len(b) == wl.fastq_bytes(0, 1000)
    lines = b.decode().split("\n")
    assert lines[-1] == "" and len(lines) == 4001
    assert a.decode().split("\n")[:4] == lines[400:404]  # read 100 is the same whatever the range / thread count
    for i in range(0, 4000, 4):
        assert lines[i] == f"@r{i // 4}" and lines[i + 2] == "+"
        assert len(lines[i + 1]) == wl.read_len == len(lines[i + 3])
        assert set(lines[i + 1]) <= set("ACGTN")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["example", "crispr", "del3", "lineage"])
def test_device_generator_equals_host_and_job_equals_oracle(name, tmp_path):
    import torch
    n = 20_000
    wl = synth.Workload(name, str(tmp_path / name), reads=n)
    run = wl.run(bc)
    text = wl.generate_fastq(0, n, threads=4).tobytes().decode().split("\n")
    seqs, quals = text[1::4], text[3::4]
    host = run.pack(seqs, quals)
    dev = wl.generate_device(run, 0, n)
    torch.cuda.synchronize()
    assert np.array_equal(dev.planes.cpu().numpy().view(np.uint32), host.planes)
    assert np.array_equal(dev.read_len.cpu().numpy().view(np.uint16), host.read_len)
    if run.quality_on:
        assert np.array_equal(dev.qual.cpu().numpy(), host.qual)
    # whole job on the device batch vs the oracle on the text
    ctr = bc.Counter(run, expected_reads=n)
    ctr.submit(dev)
    got = ctr.counters()
    o_dir, g_dir = tmp_path / "o", tmp_path / "g"
    o_dir.mkdir()
    g_dir.mkdir()
    orc = Oracle(wl.fmt, wl.samples, wl.counted, min_quality=wl.min_quality, merge=wl.merge, enrich=wl.enrich,
                 outdir=str(o_dir), prefix="p")
    orc.process_block(seqs, quals)
    assert got.pop("unsupported") == 0 and got == orc.counters()
    assert got["matched"] > n // 2
    orc.write_files()
    ctr.write_counts(str(g_dir), "p", merge=wl.merge, enrich=wl.enrich)
    assert_same_csv_set(read_csv_dir(str(g_dir), "p"), read_csv_dir(str(o_dir), "p"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["del3", "crispr", "lineage"])
def test_job_is_independent_of_batching_and_order(name, tmp_path):
    """Size-independent property at a size the oracle would need minutes for: the final table and the counters do not
    depend on how the reads are cut into batches nor on the order of the batches (every statistic of the path is an
    order-independent sum or set union); a re-run after bc_reset reproduces them; duplicates +
    matched == reads that passed the filters."""
    import torch
    n = 3_000_000
    wl = synth.Workload(name, str(tmp_path / name), reads=n)
    run = wl.run(bc)
    whole = wl.generate_device(run, 0, n)
    torch.cuda.synchronize()

    def job(cuts, reverse=False):
        ctr = bc.Counter(run, expected_reads=0)  # tiny tables: growth by rehash is exercised too
        edges = list(zip([0] + cuts, cuts + [n]))
        if reverse:
            edges = edges[::-1]
        for a, b in edges:
            ctr.submit(whole.slice(a, b))
        c = ctr.counters()
        k, lo, hi, cnt = ctr.finish_view()
        hi = hi if hi is not None else np.zeros(k, np.uint64)
        order = np.lexsort((lo, hi))
        rows = np.stack([hi[order], lo[order], cnt[order]], axis=1).copy()
        return c, rows

    c0, r0 = job([])
    assert sum(c0.values()) == n
    assert int(r0[:, 2].sum()) == c0["matched"]
    for cuts, rev in (([1_000_000, 2_000_001], False), ([123, 128, 70_000, 1_500_000, 2_999_999], True)):
        c1, r1 = job(cuts, rev)
        assert c1 == c0
        assert r1.shape == r0.shape and bool((r1 == r0).all())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["del3", "lineage"])
def test_exchange_on_one_gpu_equals_single_context(name, tmp_path):
    """Four contexts on one GPU play the ranks of a multi-GPU job on a BASELINE workload (hundreds of partitions per
    owner, hot keys for lineage): decode a quarter of the reads each, exchange once (bc_exchange_*), and the union of the
    owners' rows / the sum of their counters must equal the single-context job."""
    import torch
    from ngs_barcode_count_b200.multi import exchange_plan
    n, batch, world = 2_000_000, 250_000, 4
    wl = synth.Workload(name, str(tmp_path / name), reads=n)
    run = wl.run(bc)
    batches = [wl.generate_device(run, a, min(batch, n - a)) for a in range(0, n, batch)]
    torch.cuda.synchronize()

    def rows(ctr):
        k, lo, hi, cnt = ctr.finish_view()
        hi = hi if hi is not None else np.zeros(k, np.uint64)
        return np.stack([hi, lo, cnt], axis=1).copy()

    def canon(r):
        return r[np.lexsort((r[:, 1], r[:, 0]))]

    single = bc.Counter(run, expected_reads=n)
    for b in batches:
        single.submit(b)
    c0, r0 = single.counters(), canon(rows(single))
    ranks = [bc.Counter(run, expected_reads=n // world) for _ in range(world)]
    for r, c in enumerate(ranks):
        c.exchange_open(world, r, 64)  # far too small: the plan below has to ask for more
    for i, b in enumerate(batches):
        ranks[i * world // len(batches)].submit(b)
    matrix = [c.exchange_count(world) for c in ranks]
    need = max(exchange_plan(matrix, r)[2] for r in range(world))
    assert need > 64
    with pytest.raises(bc.BcError):
        ranks[0].exchange_connect_local(ranks)
        ranks[0].exchange_scatter(exchange_plan(matrix, 0)[0])  # past the capacity: refused, nothing written
    for r, c in enumerate(ranks):
        c.exchange_open(world, r, need)  # re-open larger: the records are still there
    for c in ranks:
        c.exchange_connect_local(ranks)
    matrix2 = [c.exchange_count(world) for c in ranks]
    assert matrix2 == matrix
    for r, c in enumerate(ranks):
        c.exchange_scatter(exchange_plan(matrix, r)[0])
    for c in ranks:
        c.sync()
    for r, c in enumerate(ranks):
        c.exchange_finish(exchange_plan(matrix, r)[1])
    c1 = {k: sum(c.counters()[k] for c in ranks) for k in c0}
    r1 = canon(np.concatenate([rows(c) for c in ranks]))
    assert c1 == c0 and sum(c0.values()) == n and (name != "del3" or c0["duplicates"] > 0)
    assert r0.shape == r1.shape and bool((r0 == r1).all())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["del3", "crispr", "lineage", "example"])
def test_specialized_kernel_equals_generic_on_workloads(name, tmp_path):
    """Millions of reads of every BASELINE workload through the NVRTC-specialised decode kernel and through the generic
    one: identical counters and identical (key, count) rows."""
    import torch
    n, batch = 3_000_000, 1_000_000
    wl = synth.Workload(name, str(tmp_path / name), reads=n)
    run = wl.run(bc)
    batches = [wl.generate_device(run, a, min(batch, n - a)) for a in range(0, n, batch)]
    torch.cuda.synchronize()

    def result(flags):
        ctr = bc.Counter(run, expected_reads=n, flags=flags)
        for b in batches:
            ctr.submit(b)
        c = ctr.counters()
        k, lo, hi, cnt = ctr.finish_view()
        hi = hi if hi is not None else np.zeros(k, np.uint64)
        order = np.lexsort((lo, hi))
        rows = np.stack([hi[order], lo[order], cnt[order]], axis=1).copy()
        prof = ctr.profile()
        ctr.close()
        return c, rows, prof

    c0, r0, p0 = result(bc.BC_CFG_NO_SPECIALIZE)
    c1, r1, p1 = result(bc.BC_CFG_SPECIALIZE)
    assert p0["specialized_launches"] == 0 and p1["generic_launches"] == 0 and p1["specialized_launches"] == len(batches)
    assert c0 == c1 and sum(c0.values()) == n
    assert r0.shape == r1.shape and bool((r0 == r1).all())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["del3", "lineage", "example"])
def test_counting_modes_agree_on_workloads(name, tmp_path):
    """Deferred partitioned counting (default), its global-table fallback and the read-by-read inline tables must give the
    same counters and the same (key, count) rows on the BASELINE workloads (hundreds of partitions per stage)."""
    import torch
    from ngs_barcode_count_b200.multi import Job
    n, batch = 1_500_000, 400_000
    wl = synth.Workload(name, str(tmp_path / name), reads=n)
    run = wl.run(bc)
    has_umi = any(run.slot(i).kind == ord("R") for i in range(run.n_slots))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        batches = [wl.generate_device(run, a, min(batch, n - a), stream=stream.cuda_stream) for a in range(0, n, batch)]
    stream.synchronize()

    def result(mode):
        ctr = bc.Counter(run, expected_reads=n, flags=bc.BC_CFG_INLINE_COUNT if mode == "inline" else 0)
        if mode == "two_stage":
            ctr.set_option("flush_two_stage", 1)
        if mode == "global":
            ctr.set_option("flush_global", 1)
        ctr.set_stream(stream.cuda_stream)
        job = Job(bc, ctr, run, 1, 0, "cuda:0", stream, has_umi, n)
        for _ in range(2):
            job.step(batches, to_host=True)
        k, lo, hi, cnt = ctr.finish_view()
        hi = hi if hi is not None else np.zeros(k, np.uint64)
        order = np.lexsort((lo, hi))
        rows = np.stack([hi[order], lo[order], cnt[order]], axis=1).copy()
        prof = ctr.profile()
        c = ctr.counters()
        ctr.close()
        return c, rows, prof

    c0, r0, p0 = result("deferred")
    assert sum(c0.values()) == n and int(r0[:, 2].sum()) == c0["matched"]
    if p0["dense_table"] and not has_umi:
        pytest.skip("dense inline counting: no deferred path for this scheme")
    assert p0["deferred_count"] == 1 and p0["flushed_global"] == 0
    # del3: one stage, its 1000 enriched compounds set aside (3) or not even that (1); the 16 keys of the example files
    # are all hot (two stages); the Zipf lineage barcodes are in between at this size
    if has_umi:
        assert p0["flush_stages"] in {"del3": (1, 3), "example": (2,), "lineage": (2, 3)}[name], p0
    for mode in ("two_stage", "global", "inline"):
        c1, r1, p1 = result(mode)
        assert c1 == c0, mode
        assert r1.shape == r0.shape and bool((r1 == r0).all()), mode
