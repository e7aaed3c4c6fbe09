#!/usr/bin/env python3
"""bench.py — reads/s decoded+counted by the B200 decode-and-count path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config del3|crispr|lineage|example] [--reads PER_GPU]
    python bench.py --impl reference ...      # the reference algorithm on the host cores (CPU oracle, oracle/)

A step is one pass of the whole job over the workload's reads: reset the tables, decode+count every batch, extract
the (key, count) rows.  `value` has the packed reads resident in HBM when the timed region starts (CUDA events);
`e2e` runs the same job through the C ABI from pinned HOST batches (H2D of every batch and D2H of the result rows
inside the timed region, wall clock between device synchronisations).  One process per GPU; reads are sharded
across ranks (weak scaling); with a random barcode the matched (key, UMI) records are routed to an owner rank by
key hash (NCCL all-to-all) so that de-duplication is globally exact, otherwise tables merge once at the end.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "reads/sec decoded+counted"
UNIT = "reads/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="del3", choices=["del3", "crispr", "lineage", "example"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: the workload's size)")
    ap.add_argument("--batch-reads", type=int, default=1 << 23)
    ap.add_argument("--e2e-reads", type=int, default=1 << 25, help="reads per GPU of the host-buffer (e2e) leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the cpu_baseline sample")
    ap.add_argument("--fastq-reads", type=int, default=8_000_000, help="reads of the FASTQ-file leg (e2e_fastq)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workdir", default=os.path.join(tempfile.gettempdir(), "bc_b200_bench"))
    return ap.parse_args()


def algorithmic_bytes_per_read(read_len, quality_on):
    """2-bit bases + N mask (+ Phred bytes when the quality filter is on) + the length word."""
    return (read_len + 3) // 4 + (read_len + 7) // 8 + (read_len if quality_on else 0) + 2


# --------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        note = None
        if not any(len(l.split(",")) >= 7 for l in self.lines):  # region shorter than nvidia-smi's start-up: one query right after it
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout
                self.lines = [l.strip() for l in out.splitlines() if l.strip()]
                note = "timed region shorter than the sampler start-up: one sample taken right after it"
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        d = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
             "samples": len(sm)}
        if note:
            d["note"] = note
        return d


# --------------------------------------------------------------------------------------------------- reference arm
def oracle_rate(wl, n_reads, threads, workdir, first=0):
    """The CPU oracle (reference algorithm, reference threading shape) over reads [first, first+n) as a FASTQ file."""
    from helpers import Oracle
    path = os.path.join(workdir, f"cpu_sample_{wl.name}.fastq")
    wl.write_fastq(path, first, n_reads, threads=threads)
    orc = Oracle(wl.fmt, wl.samples, wl.counted, min_quality=wl.min_quality, merge=wl.merge, enrich=wl.enrich,
                 outdir=workdir, prefix="cpu")
    secs, total = orc.run_fastq(path, threads)
    counters = orc.counters()
    orc.close()
    return total / secs, secs, total, counters, path


def calibrated_sample(wl, threads, workdir, target_s):
    rate, secs, _, _, _ = oracle_rate(wl, 100_000, threads, workdir)
    n = int(min(max(rate * target_s, 50_000), 20_000_000))
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    import ngs_barcode_count_b200 as bc  # noqa: F401  (workload files only; no GPU work in this arm)
    from ngs_barcode_count_b200 import synth
    from helpers import Oracle
    threads = os.cpu_count() or 1
    per_gpu = args.reads or synth.WORKLOADS[args.config]["reads"]
    os.makedirs(args.workdir, exist_ok=True)
    wl = synth.Workload(args.config, os.path.join(args.workdir, f"ref_{args.config}"), reads=per_gpu * args.gpus)
    # a bounded sample of the workload per step, sized so that (steps + warmup) steps end within a few minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    n = calibrated_sample(wl, threads, args.workdir, min(budget, args.cpu_seconds))
    path = os.path.join(args.workdir, f"cpu_sample_{wl.name}.fastq")
    wl.write_fastq(path, 0, n, threads=threads)
    times = []
    for it in range(args.warmup + args.steps):
        orc = Oracle(wl.fmt, wl.samples, wl.counted, min_quality=wl.min_quality, merge=wl.merge, enrich=wl.enrich,
                     outdir=args.workdir, prefix="cpu")
        secs, total = orc.run_fastq(path, threads)
        orc.close()
        if it >= args.warmup:
            times.append(secs)
    t = sum(times) / len(times)
    value = n / t
    sample = f"first {n} reads of the {wl.name} workload as a FASTQ file, {threads} threads (1 reader + {threads - 1} workers)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(wl, per_gpu, args, extra={"sample_reads": n}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU oracle (oracle/): C++ restatement of the reference algorithm in the reference's threading shape; the Rust "
                "reference itself cannot be built in this image (no cargo/rustc)",
    }))


def workload_config(wl, per_gpu, args, extra=None):
    c = {"workload": f"{wl.name}: {wl.read_len}-nt reads, scheme of {wl.template_len} nt, {per_gpu} reads per GPU",
         "reads_per_gpu": per_gpu, "read_len": wl.read_len, "min_quality": wl.min_quality, "enrich": wl.enrich,
         "batch_reads": args.batch_reads, "l2": "inputs larger than L2 (every batch is read once per step)"}
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------------------- B200 arm
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import ngs_barcode_count_b200 as bc
    from ngs_barcode_count_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the decode-and-count path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device(dev))
    per_gpu = args.reads or synth.WORKLOADS[args.config]["reads"]
    os.makedirs(args.workdir, exist_ok=True)
    wl = synth.Workload(args.config, os.path.join(args.workdir, f"{args.config}_r{rank}"), reads=per_gpu * world)
    run = wl.run(bc)
    has_umi = any(run.slot(i).kind == ord("R") for i in range(run.n_slots))
    ctr = bc.Counter(run, device=local, expected_reads=per_gpu)
    stream = torch.cuda.Stream(device=dev)
    ctr.set_stream(stream.cuda_stream)

    # ---- inputs: this rank's shard, generated straight into HBM
    first = rank * per_gpu
    batches = []
    with torch.cuda.stream(stream):
        for a in range(0, per_gpu, args.batch_reads):
            n = min(args.batch_reads, per_gpu - a)
            batches.append(wl.generate_device(run, first + a, n, device=dev, stream=stream.cuda_stream))
    stream.synchronize()

    from ngs_barcode_count_b200.multi import Job
    job = Job(bc, ctr, run, world, rank, dev, stream, has_umi, args.batch_reads)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident timing (CUDA events on the ctx stream, max over ranks)
    sampler = ClockSampler(local)
    sampler.start()  # sampled from the warm-up on (same load), so that short timed regions still get samples
    for _ in range(args.warmup):
        job.step(batches)
    barrier()
    ctr.reset_profile()
    ctr.set_profiling(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for _ in range(args.steps):
        n_rows = job.step(batches)
    with torch.cuda.stream(stream):
        ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    ctr.set_profiling(False)
    prof = ctr.profile()
    counters = job.global_counters()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    total_reads = per_gpu * world
    value = total_reads / (ms_step * 1e-3)
    assert sum(counters.values()) == total_reads * 1, (counters, total_reads)  # every read has exactly one outcome

    # roofline of the dominant kernel (decode): algorithmic bytes / average launch time
    bpr = algorithmic_bytes_per_read(wl.read_len, run.quality_on)
    dec_launches = prof["launches"]["decode"]
    dec_ms = prof["ms"]["decode"]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # the counting step adds "one table slot RMW (random access => count 32 B sector per probe)".  A table
    # that does not fit the 126 MB L2 costs DRAM sectors: read + write-back of the (key, UMI) set sector for every read
    # that reaches the count, and of the key->count sector for every pair seen for the first time.
    reached = (counters["matched"] + counters["duplicates"]) / total_reads
    fresh = counters["matched"] / total_reads
    slot_b = 8 if prof["dense_table"] else (32 if prof["wide_keys"] else 16)
    map_in_dram = prof["table_capacity"] * slot_b > 126e6
    table_bpr = (64.0 * reached if has_umi else 0.0) + (64.0 * (fresh if has_umi else reached) if map_in_dram else 0.0)
    if job.routed:  # multi-GPU with a random barcode: the decode kernel only buckets records, the owner's insert kernel counts
        table_bpr = 16.0 * reached
    elif prof["deferred_count"]:
        # deferred counting: k_decode touches no table; every read owns one record slot (8 or 16 bytes, holes included)
        # that the kernel writes once.  De-duplication / counting run once per job in the flush kernels (kernel_ms.finish).
        table_bpr = 16.0 if prof["wide_keys"] else 8.0
    achieved_input = (bpr * per_gpu * args.steps / 1e9) / (dec_ms * 1e-3) if dec_ms > 0 else 0.0
    achieved = ((bpr + table_bpr) * per_gpu * args.steps / 1e9) / (dec_ms * 1e-3) if dec_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.config)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "k_decode", "bytes_per_read": bpr + table_bpr, "bytes_per_read_input": bpr,
                "bytes_per_read_table_sectors": table_bpr, "achieved_input_only": achieved_input,
                "frac_input_only": achieved_input / peak, "reads_per_launch": per_gpu * args.steps / max(1, dec_launches),
                "avg_launch_ms": dec_ms / max(1, dec_launches), "kernel_share_of_step": dec_ms / (ms_total if ms_total else 1),
                "peak_source": peak_src, "kernel_ms": prof["ms"], "kernel_launches": prof["launches"],
                "counting": ("deferred: records appended by k_decode, partitioned shared-memory de-duplication + counting at the "
                             "flush (kernel_ms.finish)" if prof["deferred_count"] else
                             "inline: tables updated read by read inside k_decode"),
                "flushed_global": prof["flushed_global"], "flush_stages": prof["flush_stages"]}
    gpu_launches = sum(prof["launches"].values())
    if rank == 0:
        try:  # INT-pipe peaks of this GPU (register-only microbenchmarks) and the pivot-test rate the kernel reaches
            peaks = synth.measure_int_peaks()
            windows = wl.read_len - wl.template_len + 1
            roofline["int"] = {"peak_lop3_tops": peaks["lop3"], "peak_popc_tops": peaks["popc"], "peak_shf_tops": peaks["shf"],
                               "windows_per_read": windows,
                               "achieved_window_tests_tera_per_s": windows * per_gpu * args.steps / (dec_ms * 1e-3) / 1e12 if dec_ms else 0.0,
                               "note": "peaks are lane-ops/s of the alu pipe (LOP3, SHF) and of POPC; the pivot prefilter is bit-sliced: per "
                                       "32 window offsets and constant position one funnel shift + 2.25 LOP3 (carry-save adds), no POPC"}
        except Exception as e:  # the microbenchmark is informational
            roofline["int"] = {"error": str(e)}

    # ---- end to end from pinned host batches through the C ABI
    e2e = None
    if not args.no_e2e:
        e2e_n = min(per_gpu, args.e2e_reads)
        host_batches, h2d = [], 0
        done = 0
        for b in batches:
            if done >= e2e_n:
                break
            n = min(b.n, e2e_n - done)
            hb = job.to_pinned(b.slice(0, n))
            host_batches.append(hb)
            h2d += hb.nbytes
            done += n
        # a context sized for this leg's read count (tables are cleared and scanned once per job)
        ctr2 = bc.Counter(run, device=local, expected_reads=e2e_n)
        ctr2.set_stream(stream.cuda_stream)
        job2 = Job(bc, ctr2, run, world, rank, dev, stream, has_umi, args.batch_reads)
        for _ in range(2):
            job2.step(host_batches, to_host=True)
        barrier()
        ctr2.reset_profile()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rows = job2.step(host_batches, to_host=True)
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        p2 = ctr2.profile()
        ctr2.close()
        e2e = {"value": e2e_n * world / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": p2["h2d_bytes"] // args.steps,
               "d2h_bytes_per_step": p2["d2h_bytes"] // args.steps, "reads_per_gpu": e2e_n,
               "what": "pinned host bc_batch buffers -> bc_submit/bc_decode_route (H2D inside) -> bc_finish rows on the host"}
        del host_batches

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded FASTQ sample of the same workload
    cpu = None
    fastq_leg = None
    if rank == 0 and world == 1 and not args.no_cpu:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        threads = os.cpu_count() or 1
        n = calibrated_sample(wl, threads, args.workdir, args.cpu_seconds)
        rate, secs, total, cpu_counters, path = oracle_rate(wl, n, threads, args.workdir)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n} reads of the workload as a FASTQ file, {secs:.1f} s, 1 reader + {threads - 1} workers"}
        # the same FASTQ file through the product's own ingest (parse + pack on host threads, H2D, kernels) and a
        # parity check of the counters at this size
        # parity first: the very file the oracle just read
        ctr3 = bc.Counter(run, device=local, expected_reads=max(total, args.fastq_reads))
        got_n = ctr3.count_fastq(path, threads=threads, batch_reads=1 << 20)
        got = ctr3.counters()
        got.pop("unsupported")
        assert got_n == total and got == cpu_counters, ("GPU/oracle counters differ on the CPU sample", got, cpu_counters)
        # throughput on a larger file of the same workload (the CPU sample is over in milliseconds here)
        big = os.path.join(args.workdir, f"fastq_leg_{wl.name}.fastq")
        try:
            wl.write_fastq(big, 0, args.fastq_reads, threads=threads)
        except OSError:  # no room for the larger file: time the CPU sample file instead
            big = path
        ctr3.reset()
        ctr3.count_fastq(big, threads=threads, batch_reads=1 << 20)  # warm-up pass: page cache
        ctr3.reset()
        t0 = time.perf_counter()
        big_n = ctr3.count_fastq(big, threads=threads, batch_reads=1 << 20)
        ctr3.counters()
        dt = time.perf_counter() - t0
        ctr3.close()
        if big != path:
            os.remove(big)
        fastq_leg = {"value": big_n / dt, "unit": UNIT, "reads": big_n, "host_threads": threads,
                     "what": "plain FASTQ file (page cache) -> bch_count_fastq (mmap, host split+pack, H2D, kernels) -> counters; "
                             "the GPU counters equal the oracle's on the CPU sample file"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(wl, per_gpu, args, extra={"parallelism": job.parallelism}),
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
            "e2e_fastq": fastq_leg, "counters": counters, "rows": int(n_rows),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
