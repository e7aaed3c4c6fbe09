"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): N ranks over NCCL give exactly the single-GPU
result — counters, number of rows, every (key, count) row and the enrichment marginals — both for the one-exchange UMI de-duplication (DEL, lineage)
and for the final table merge (CRISPR)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["del3", "crispr", "lineage"])
def test_n_ranks_equal_one_gpu(name):
    n = n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + os.getpid() % 300), os.path.join(ROOT, "tests", "multi_gpu_check.py"), name, "300000", "70000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_CHECK" in r.stdout and " OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["del3_umi", "crispr", "lineage_raw", "example"])
def test_cli_devices_on_real_gpus(case, tmp_path):
    """`barcode-count --devices all`: one process, one context per GPU of the box (bch_count_fastq_multi: batches in turn,
    records streamed to their owner GPU over NVLink peer memory or dense tables added, rows of all owners gathered into one
    CSV set) writes the reference's CSV set for the golden cases — and, on a larger synthetic file, the same files as one GPU."""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ngs_barcode_count_b200 as bc
    from helpers import assert_same_csv_set, load_golden, read_csv_dir
    exp, paths = load_golden(case)
    fl = exp["flags"]

    def run(fastq, out, devices, batch):
        out.mkdir()
        cmd = [bc.CLI_PATH, "-f", str(fastq), "-q", paths["fmt"], "-o", str(out), "-p", "golden", "--devices", devices, f"--batch-reads={batch}",
               f"--min-quality={fl['min_quality']}"]
        if paths["samples"]:
            cmd += ["-s", paths["samples"]]
        if paths["counted"]:
            cmd += ["-c", paths["counted"]]
        cmd += (["-m"] if fl["merge"] else []) + (["-e"] if fl["enrich"] else [])
        for flag, key in (("--max-errors-counted-barcode", "max_barcode"), ("--max-errors-sample", "max_sample"),
                          ("--max-errors-constant", "max_constant")):
            if fl[key] is not None:
                cmd += [f"{flag}={fl[key]}"]
        r = subprocess.run(cmd, capture_output=True)
        assert r.returncode == 0, r.stderr[-2000:]
        return read_csv_dir(str(out), "golden"), r.stdout.decode()

    got, stdout = run(paths["fastq"], tmp_path / "all", "all", 37)
    assert_same_csv_set(got, exp["files"])
    assert f"Correctly matched sequences: {exp['counters']['matched']:,}" in stdout
    # the same reads repeated 100 times with their order shuffled: several batches per GPU, real duplicates for the UMI schemes
    import random
    lines = open(paths["fastq"]).read().split("\n")
    recs = ["\n".join(lines[i:i + 4]) + "\n" for i in range(0, len(lines) - 3, 4)]
    rng = random.Random(1)
    big = tmp_path / "big.fastq"
    with open(big, "w") as f:
        for rep in range(100):
            rng.shuffle(recs)
            f.write("".join(recs))
    one, _ = run(big, tmp_path / "one", "0", 5000)
    many, _ = run(big, tmp_path / "many", "all", 5000)
    assert_same_csv_set(many, one)
