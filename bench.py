#!/usr/bin/env python3
"""bench.py — reads/s decoded+counted by the B200 decode-and-count path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config del3|crispr|lineage|example] [--reads PER_GPU]
    python bench.py --impl reference ...      # the reference algorithm on the host cores (CPU oracle, oracle/)

A step is one pass of the whole job over the workload's reads: reset, decode every batch, de-duplicate + count (the
flush), extract the (key, count) rows.  `value` has the packed reads resident in HBM when the timed region starts (CUDA
events); `e2e` runs the same job through the C ABI from pinned HOST batches — in their transfer form (bc_submit_wire:
159 instead of 218 bytes per read across PCIe) unless --e2e-form plain — with the H2D of every batch and the D2H of the
result rows inside the timed region, wall clock between device synchronisations; `e2e.binned_quality` repeats it on the
same reads with binned qualities (2-bit codes); `e2e_fastq` starts from a FASTQ file (host framing + packing included,
phases of the ingest thread reported).  One process per GPU; reads are sharded across ranks (weak scaling).  With hashed
keys (any scheme with a random barcode) every batch's records leave for their owner rank hash(key) % N right after its
decode — a scatter kernel on a side stream writes them into the owner's receive buffer over NVLink peer memory — and
every owner de-duplicates what it owns, so de-duplication is globally exact; dense count tables (CRISPR) merge with one
all-reduce.  For N > 1 the line carries `parity_n_ranks`: the N-rank job and
a single-GPU job over the same reads (a reduced range) must have identical counters and the same row digest.
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
    ap.add_argument("--e2e-reads", type=int, default=0,
                    help="reads per GPU of the host-buffer (e2e) leg (default: 2^27 on one GPU, 2^25 per GPU on several)")
    ap.add_argument("--parity-reads", type=int, default=1 << 24, help="total reads of the N-rank == 1-GPU parity job (N > 1)")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE workloads (other_configs)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the cpu_baseline sample")
    ap.add_argument("--fastq-reads", type=int, default=8_000_000, help="reads of the FASTQ-file leg (e2e_fastq)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-binned", action="store_true", help="skip the binned-quality variant of the e2e leg")
    ap.add_argument("--e2e-form", choices=["wire", "plain"], default="wire",
                    help="host batches of the e2e leg: the transfer form (bc_submit_wire: 6-bit quality codes, N calls as a list) or plain "
                         "bc_batch arrays (bc_submit)")
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
    emit_json({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(wl, per_gpu, args, extra={"sample_reads": n}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU oracle (oracle/): C++ restatement of the reference algorithm in the reference's threading shape; the Rust "
                "reference itself cannot be built in this image (no cargo/rustc)",
    })


def workload_config(wl, per_gpu, args, extra=None):
    c = {"workload": f"{wl.name}: {wl.read_len}-nt reads, scheme of {wl.template_len} nt, {per_gpu} reads per GPU",
         "reads_per_gpu": per_gpu, "read_len": wl.read_len, "min_quality": wl.min_quality, "enrich": wl.enrich,
         "batch_reads": args.batch_reads, "l2": "inputs larger than L2 (every batch is read once per step)"}
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------------------- B200 arm
def numa_cpus_of_gpu(index):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when the box does not say."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id  # 0000:1B:00.0
    except Exception:
        try:
            bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()[-12:]
        except Exception:
            return None, None
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus.lower()}/numa_node").read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return node, cpus
    except Exception:
        return None, None


class Leg:
    """One workload on this rank's GPU: its files, context, device-resident batches and job."""

    def __init__(self, bc, synth, args, name, per_gpu, world, rank, local, stream, tag=""):
        import torch
        self.wl = synth.Workload(name, os.path.join(args.workdir, f"{name}{tag}_r{rank}"), reads=per_gpu * world)
        self.run = self.wl.run(bc)
        self.has_umi = any(self.run.slot(i).kind == ord("R") for i in range(self.run.n_slots))
        self.per_gpu, self.world, self.rank, self.dev, self.stream = per_gpu, world, rank, f"cuda:{local}", stream
        self.ctr = bc.Counter(self.run, device=local, expected_reads=per_gpu)
        self.ctr.set_stream(stream.cuda_stream)
        self.batches = []
        first = rank * per_gpu
        with torch.cuda.stream(stream):
            for a in range(0, per_gpu, args.batch_reads):
                n = min(args.batch_reads, per_gpu - a)
                self.batches.append(self.wl.generate_device(self.run, first + a, n, device=self.dev, stream=stream.cuda_stream))
        stream.synchronize()
        from ngs_barcode_count_b200.multi import Job
        self.job = Job(bc, self.ctr, self.run, world, rank, self.dev, stream, self.has_umi, per_gpu)

    def close(self):
        self.batches = []
        self.ctr.close()


def timed_steps(leg, steps, warmup, barrier):
    """-> (ms per step: CUDA events on the ctx stream, max over ranks; profile; rows; counters summed over ranks)"""
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        leg.job.step(leg.batches)
    barrier()
    leg.ctr.reset_profile()
    leg.ctr.set_profiling(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(leg.stream):
        ev0.record(leg.stream)
    for _ in range(steps):
        n_rows = leg.job.step(leg.batches)
    with torch.cuda.stream(leg.stream):
        ev1.record(leg.stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    leg.ctr.set_profiling(False)
    prof = leg.ctr.profile()
    t = torch.tensor([ms_total], dtype=torch.float64, device=leg.dev)
    if leg.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps, ms_total, prof, n_rows, leg.job.global_counters()


def decode_roofline(leg, prof, counters, steps, ms_total, peak, config):
    """HBM roofline of the dominant kernel (k_decode): SURVEY.md §8(d) bytes per read x reads per launch / average launch time."""
    wl, per_gpu, total_reads = leg.wl, leg.per_gpu, leg.per_gpu * leg.world
    bpr = algorithmic_bytes_per_read(wl.read_len, leg.run.quality_on)
    dec_launches, dec_ms = prof["launches"]["decode"], prof["ms"]["decode"]
    # what the kernel writes / touches beside its input.  Deferred counting: every read owns one record slot (8 or 16 bytes,
    # holes included) that the kernel writes once; the tables are not touched (the flush counts: kernel_ms.finish).
    # Inline dense table (CRISPR): one RED per matched read into a counter array; counted as a 32-byte sector read + written
    # back only when the array does not fit the 126 MB L2.
    reached = (counters["matched"] + counters["duplicates"]) / total_reads
    if prof["deferred_count"]:
        table_bpr = 16.0 if prof["wide_keys"] else 8.0
    else:
        slot_b = 8 if prof["dense_table"] else (32 if prof["wide_keys"] else 16)
        in_dram = prof["table_capacity"] * slot_b > 126e6
        table_bpr = (64.0 * reached if leg.has_umi else 0.0) + (64.0 * reached if in_dram else 0.0)
    achieved_input = (bpr * per_gpu * steps / 1e9) / (dec_ms * 1e-3) if dec_ms > 0 else 0.0
    achieved = ((bpr + table_bpr) * per_gpu * steps / 1e9) / (dec_ms * 1e-3) if dec_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(config)
    return {"bound": "hbm", "achieved": achieved, "peak": peak[0], "unit": "GB/s", "frac": achieved / peak[0], "traffic": traffic,
            "kernel": "k_decode_jit (the decode kernel specialised for the run by NVRTC)" if prof["specialized_launches"] else "k_decode",
            "specialization": prof["specialization"], "bytes_per_read": bpr + table_bpr, "bytes_per_read_input": bpr,
            "bytes_per_read_written_or_touched": table_bpr, "achieved_input_only": achieved_input,
            "frac_input_only": achieved_input / peak[0], "reads_per_launch": per_gpu * steps / max(1, dec_launches),
            "avg_launch_ms": dec_ms / max(1, dec_launches), "kernel_share_of_step": dec_ms / (ms_total if ms_total else 1),
            "peak_source": peak[1], "kernel_ms": prof["ms"], "kernel_launches": prof["launches"],
            "counting": ("deferred: records appended by k_decode, partitioned shared-memory de-duplication + counting at the "
                         "flush (kernel_ms.finish)" if prof["deferred_count"] else
                         "inline: one RED per matched read into a dense count array inside k_decode"),
            "flushed_global": prof["flushed_global"], "flush_stages": prof["flush_stages"]}


def oracle_parity(bc, leg, n, workdir, threads):
    """Counters and canonical CSV set of the GPU job vs the CPU oracle on the first n reads of the workload (text)."""
    from helpers import Oracle, assert_same_csv_set, read_csv_dir
    wl = leg.wl
    text = wl.generate_fastq(0, n, threads=threads).tobytes().decode().split("\n")
    seqs, quals = text[1::4], text[3::4]
    o_dir, g_dir = os.path.join(workdir, f"par_o_{wl.name}"), os.path.join(workdir, f"par_g_{wl.name}")
    for d in (o_dir, g_dir):
        os.makedirs(d, exist_ok=True)
        for f in os.listdir(d):
            os.remove(os.path.join(d, f))
    orc = Oracle(wl.fmt, wl.samples, wl.counted, min_quality=wl.min_quality, merge=wl.merge, enrich=wl.enrich, outdir=o_dir, prefix="p")
    orc.process_block(seqs, quals)
    ctr = bc.Counter(leg.run, device=int(leg.dev.split(":")[1]), expected_reads=n)
    ctr.submit(leg.run.pack(seqs, quals, threads=threads))
    got = ctr.counters()
    ok = got.pop("unsupported") == 0 and got == orc.counters()
    if ok:
        orc.write_files()
        ctr.write_counts(g_dir, "p", merge=wl.merge, enrich=wl.enrich)
        try:
            assert_same_csv_set(read_csv_dir(g_dir, "p"), read_csv_dir(o_dir, "p"))
        except AssertionError:
            ok = False
    orc.close()
    ctr.close()
    return ok


_JSON_FD = None


def quiet_stdout():
    """Libraries print to stdout (NCCL its version, torchrun its banner): keep stdout for the ONE JSON line."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit_json(obj):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(obj) + "\n").encode())


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    import ngs_barcode_count_b200 as bc
    from ngs_barcode_count_b200 import synth
    from ngs_barcode_count_b200.multi import Job

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
    stream = torch.cuda.Stream(device=dev)
    threads = os.cpu_count() or 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = (json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)")
    else:
        peak = (6650.0, "fallback (B200_PROFILING.md)")

    # ---- the headline workload: this rank's shard, generated straight into HBM
    leg = Leg(bc, synth, args, args.config, per_gpu, world, rank, local, stream)
    wl, run, ctr, job = leg.wl, leg.run, leg.ctr, leg.job

    # ---- kernel-resident timing (CUDA events on the ctx stream, max over ranks)
    sampler = ClockSampler(local)
    sampler.start()  # sampled from the warm-up on (same load), so that short timed regions still get samples
    ms_step, ms_total, prof, n_rows, counters = timed_steps(leg, args.steps, args.warmup, barrier)
    clocks = sampler.stop()
    total_reads = per_gpu * world
    value = total_reads / (ms_step * 1e-3)
    assert sum(counters.values()) == total_reads, (counters, total_reads)  # every read has exactly one outcome
    roofline = decode_roofline(leg, prof, counters, args.steps, ms_total, peak, args.config)
    gpu_launches = sum(prof["launches"].values())
    if rank == 0:
        try:  # INT-pipe peaks of this GPU (register-only microbenchmarks) and the window-test rate the kernel reaches
            peaks = synth.measure_int_peaks()
            windows = wl.read_len - wl.template_len + 1
            dec_ms = prof["ms"]["decode"]
            roofline["int"] = {"peak_lop3_tops": peaks["lop3"], "peak_popc_tops": peaks["popc"], "peak_shf_tops": peaks["shf"],
                               "windows_per_read": windows,
                               "achieved_window_tests_tera_per_s": windows * per_gpu * args.steps / (dec_ms * 1e-3) / 1e12 if dec_ms else 0.0,
                               "note": "peaks are lane-ops/s of the alu pipe (LOP3, SHF) and of POPC; the exact-match prefilter is bit-sliced: "
                                       "per 32 window offsets 16 funnel shifts + 8 LOP3; the repair scan (reads without an exact window) adds "
                                       "one funnel shift + 2.25 LOP3 per pivot position and 32 offsets, no POPC"}
        except Exception as e:  # the microbenchmark is informational
            roofline["int"] = {"error": str(e)}

    # ---- K4 and the row read-back, timed on their own (the reference's Compute vs Total split, main.rs:127-164)
    extras = {}
    if job.owns_rows():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = ctr.finish_view()[0]
        torch.cuda.synchronize()
        extras["rows_d2h_ms"] = (time.perf_counter() - t0) * 1e3
        extras["rows_this_rank"] = int(k)
    if wl.enrich:
        ctr.reset_profile()
        ctr.set_profiling(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_marg = job.merged_marginals()  # dense marginals of this rank's rows (+ all-reduce over the owners)
        torch.cuda.synchronize()
        extras["enrich_wall_ms"] = (time.perf_counter() - t0) * 1e3
        ctr.set_profiling(False)
        extras["enrich_ms"] = ctr.profile()["ms"]["enrich"]
        extras["enrich_dense_counters"] = n_marg
        if rank == 0:
            t0 = time.perf_counter()
            s_tab, d_tab = ctr.enrich(doubles=True)
            extras["enrich_tables_to_host_ms"] = (time.perf_counter() - t0) * 1e3
            extras["enrich_rows"] = {"single": int(len(s_tab["count"])), "double": int(len(d_tab["count"]))}
            # every single-barcode marginal sums to the number of counted molecules
            k_counted = sum(1 for i in range(run.n_slots) if run.slot(i).kind == ord("B"))
            assert int(s_tab["count"].sum()) == k_counted * counters["matched"], (int(s_tab["count"].sum()), counters)
            del s_tab, d_tab

    # ---- N ranks == 1 GPU (driver-verifiable): a reduced range of the same workload through both
    parity_n = None
    if world > 1:
        small_per = max(args.batch_reads // 8, args.parity_reads // world)
        pleg = Leg(bc, synth, args, args.config, small_per, world, rank, local, stream, tag="_par")
        pleg.job.step(pleg.batches)
        got_c, got_d = pleg.job.global_counters(), pleg.job.checksum()
        parity_n = {"reads": small_per * world}
        if rank == 0:
            single = bc.Counter(pleg.run, device=local, expected_reads=small_per * world)
            single.set_stream(stream.cuda_stream)
            sjob = Job(bc, single, pleg.run, 1, 0, dev, stream, pleg.has_umi, small_per * world)
            with torch.cuda.stream(stream):
                whole = [pleg.wl.generate_device(pleg.run, a, min(args.batch_reads, small_per * world - a), device=dev, stream=stream.cuda_stream)
                         for a in range(0, small_per * world, args.batch_reads)]
            stream.synchronize()
            sjob.step(whole)
            want_c, want_d = sjob.global_counters(), sjob.checksum()
            ok = got_c == want_c and got_d == want_d
            parity_n.update({"result": "ok" if ok else "MISMATCH", "rows": got_d[0], "digest": f"{got_d[2]:016x}",
                             "single_gpu_rows": want_d[0], "single_gpu_digest": f"{want_d[2]:016x}",
                             "what": "counters and (rows, sum of counts, sum of mix64(key)*count) of the N-rank job vs one GPU over the same reads"})
            del whole
            single.close()
            flag = torch.tensor([1 if ok else 0], device=dev)
        else:
            flag = torch.tensor([0], device=dev)
        dist.broadcast(flag, src=0)
        pleg.close()
        if not int(flag.item()):
            if rank == 0:
                emit_json({"error": "N-rank result differs from the single-GPU result", "parity_n_ranks": parity_n})
            dist.destroy_process_group()
            sys.exit(1)

    # ---- end to end from pinned host batches through the C ABI
    e2e = None
    if not args.no_e2e:
        e2e_n = min(per_gpu, args.e2e_reads or ((1 << 27) if world == 1 else (1 << 25)))
        node, cpus = numa_cpus_of_gpu(local)
        old_aff = None
        if cpus:  # pinned staging buffers on the GPU's own NUMA node (first touch by this process)
            try:
                old_aff = os.sched_getaffinity(0)
                os.sched_setaffinity(0, cpus & old_aff or old_aff)
            except OSError:
                old_aff = None
        host_batches, done = [], 0
        for b in leg.batches:
            if done >= e2e_n:
                break
            n = min(b.n, e2e_n - done)
            host_batches.append(job.to_pinned(b.slice(0, n), wire=args.e2e_form == "wire"))
            done += n
        # a context sized for this leg's read count
        ctr2 = bc.Counter(run, device=local, expected_reads=e2e_n)
        ctr2.set_stream(stream.cuda_stream)
        job2 = Job(bc, ctr2, run, world, rank, dev, stream, leg.has_umi, e2e_n)
        for _ in range(2):
            job2.step(host_batches, to_host=True)
        barrier()
        ctr2.reset_profile()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            job2.step(host_batches, to_host=True)
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        p2 = ctr2.profile()
        ctr2.close()
        e2e = {"value": e2e_n * world / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": p2["h2d_bytes"] // args.steps,
               "d2h_bytes_per_step": p2["d2h_bytes"] // args.steps, "reads_per_gpu": e2e_n, "numa_node_of_gpu": node,
               "h2d_bytes_per_read": p2["h2d_bytes"] / args.steps / e2e_n, "host_batch_form": args.e2e_form,
               "what": ("pinned host batches in their transfer form (bc_wire_batch: lo/hi planes, N calls as a list, 6-bit quality codes) -> "
                        "bc_submit_wire (H2D + expansion on the device inside)" if args.e2e_form == "wire" else
                        "pinned host bc_batch buffers -> bc_submit (H2D inside)") +
                       " [-> one record exchange over NVLink] -> bc_finish rows on the host"}
        del host_batches
        # ---- the same leg on reads whose qualities are binned the way current Illumina instruments emit them (RTA3: Q2 / Q12 /
        # Q23 / Q37): four distinct characters travel as 2-bit codes.  Another input (the quality filter sees other scores), so
        # it is reported beside `e2e`, not as it; N = 1 only.
        if world == 1 and wl.min_quality > 0 and args.e2e_form == "wire" and not args.no_binned:
            bn = min(e2e_n, 1 << 26)
            u8 = lambda v: torch.tensor(v, dtype=torch.uint8, device=dev)

            def bin4(qual):  # Phred+33 characters -> '#' (Q2), '-' (Q12), '8' (Q23), 'F' (Q37)
                return torch.where(qual < 40, u8(35), torch.where(qual < 51, u8(45), torch.where(qual < 63, u8(56), u8(70))))
            binned, done = [], 0
            for b in leg.batches:
                if done >= bn:
                    break
                n = min(b.n, bn - done)
                sl = b.slice(0, n)
                qb = bin4(sl.qual)
                binned.append(job.to_pinned(bc.Batch(n, sl.plane_stride, sl.qual_stride, sl.planes, sl.read_len, qb, device=True), wire=True))
                del qb
                done += n
            ctr4 = bc.Counter(run, device=local, expected_reads=bn)
            ctr4.set_stream(stream.cuda_stream)
            job4 = Job(bc, ctr4, run, 1, 0, dev, stream, leg.has_umi, bn)
            job4.step(binned, to_host=True)
            ctr4.reset_profile()
            t0 = time.perf_counter()
            for _ in range(3):
                job4.step(binned, to_host=True)
            torch.cuda.synchronize()
            dtb = (time.perf_counter() - t0) / 3
            p4 = ctr4.profile()
            cb = ctr4.counters()
            assert sum(cb.values()) == bn, cb
            ctr4.close()
            e2e["binned_quality"] = {"value": bn / dtb, "unit": UNIT, "reads": bn, "h2d_bytes_per_read": p4["h2d_bytes"] / 3 / bn,
                                     "quality_code_bits": binned[0].qual_bits, "counters": cb,
                                     "what": "the e2e leg on the same reads with their qualities binned to Q2 / Q12 / Q23 / Q37 (four characters -> "
                                             "2-bit codes in the transfer form)"}
            del binned
        if old_aff:
            os.sched_setaffinity(0, old_aff)

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded FASTQ sample of the same workload
    cpu = None
    fastq_leg = None
    others = None
    if rank == 0 and world == 1 and not args.no_cpu:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        n = calibrated_sample(wl, threads, args.workdir, args.cpu_seconds)
        rate, secs, total, cpu_counters, path = oracle_rate(wl, n, threads, args.workdir)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n} reads of the workload as a FASTQ file, {secs:.1f} s, 1 reader + {threads - 1} workers",
               "note": "C++ restatement of the reference algorithm (std::map under one mutex where the reference has ahash maps): "
                       "read every GPU/CPU ratio as an upper bound"}
        # the same FASTQ file through the product's own ingest (parse + pack on host threads, H2D, kernels) and a
        # parity check of the counters at this size: first the very file the oracle just read
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
        best = None
        for _ in range(3):
            ctr3.reset()
            t0 = time.perf_counter()
            big_n = ctr3.count_fastq(big, threads=threads, batch_reads=1 << 20)
            ctr3.counters()
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, phases = dt, ctr3.ingest_stats()
        ctr3.close()
        if big != path:
            os.remove(big)
        fastq_leg = {"value": big_n / best, "unit": UNIT, "reads": big_n, "host_threads": threads, "passes": "best of 3 after one warm-up",
                     "ingest_thread_phases": phases,
                     "what": "plain FASTQ file (page cache) -> bch_count_fastq (mmap, host split+pack, H2D, kernels) -> counters; "
                             "the GPU counters equal the oracle's on the CPU sample file"}

    # ---- the other BASELINE workloads, kernel-resident, with an oracle parity flag each (N=1 only)
    if world == 1 and not args.no_others:
        leg.batches = []  # frees the headline workload's reads
        torch.cuda.empty_cache()
        others = {}
        for name in ("crispr", "lineage", "example", "del3"):
            if name == args.config:
                continue
            n_o = min(synth.WORKLOADS[name]["reads"], 100_000_000)
            oleg = Leg(bc, synth, args, name, n_o, 1, 0, local, stream, tag="_oc")
            o_ms, o_total, o_prof, o_rows, o_cnt = timed_steps(oleg, 3, 2, barrier)
            o_roof = decode_roofline(oleg, o_prof, o_cnt, 3, o_total, peak, name)
            ok = sum(o_cnt.values()) == n_o
            par = oracle_parity(bc, oleg, 60_000, args.workdir, threads) if not args.no_cpu else None
            others[name] = {"reads": n_o, "value": n_o / (o_ms * 1e-3), "unit": UNIT, "ms_per_step": o_ms, "rows": int(o_rows),
                            "roofline_frac": o_roof["frac"], "roofline_frac_input_only": o_roof["frac_input_only"],
                            "decode_avg_launch_ms": o_roof["avg_launch_ms"], "kernel_ms": o_prof["ms"],
                            "flush_stages": o_prof["flush_stages"], "every_read_has_one_outcome": ok,
                            "parity_vs_oracle_60k_reads": ("ok" if par else "MISMATCH") if par is not None else None}
            assert ok and par is not False, (name, others[name])
            oleg.close()
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(wl, per_gpu, args, extra={"parallelism": job.parallelism}),
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
            "e2e_fastq": fastq_leg, "counters": counters, "rows": int(n_rows), "parity_n_ranks": parity_n,
            "other_configs": others,
        }
        line.update(extras)
        emit_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
