"""Run under torchrun with N >= 2 GPUs (tests/test_multi_gpu.py launches it): every rank decodes its shard of one
global read range through ngs-barcode-count_b200/multi.py (one exchange of the records over NVLink, or one all-reduce of
the dense table); the merged result must be identical to a single-GPU job over the whole range: counters, every
(key, count) row, the order-independent row digest bench.py prints, and the enrichment marginals."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ngs_barcode_count_b200 as bc  # noqa: E402
from ngs_barcode_count_b200 import synth  # noqa: E402
from ngs_barcode_count_b200.multi import Job  # noqa: E402


def rows_of(ctr):
    n, lo, hi, cnt = ctr.finish_view()
    hi = hi if hi is not None else np.zeros(n, np.uint64)
    order = np.lexsort((lo, hi))
    return np.stack([hi[order], lo[order], cnt[order]], axis=1).copy()


def marg_of(ctr):
    s, d = ctr.enrich(doubles=True)
    out = []
    for t in (s, d):
        out += sorted(zip(t["mask"].tolist(), t["key_hi"].tolist(), t["key_lo"].tolist(), t["count"].tolist()))
    return out


def main():
    name, per_gpu, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    wl = synth.Workload(name, f"/tmp/bc_mgpu_{name}_r{rank}", reads=per_gpu * world)
    run = wl.run(bc)
    has_umi = any(run.slot(i).kind == ord("R") for i in range(run.n_slots))
    stream = torch.cuda.Stream(device=dev)
    ctr = bc.Counter(run, device=local, expected_reads=per_gpu * 2)
    ctr.set_stream(stream.cuda_stream)
    # a receive buffer sized for a tenth of the shard: the first job has to grow it (collectively)
    job = Job(bc, ctr, run, world, rank, dev, stream, has_umi, max(1, per_gpu // 10))
    with torch.cuda.stream(stream):
        batches = [wl.generate_device(run, rank * per_gpu + a, min(batch, per_gpu - a), device=dev, stream=stream.cuda_stream)
                   for a in range(0, per_gpu, batch)]
    for _ in range(2):  # twice: reset between jobs must work
        n_rows = job.step(batches, to_host=True)
    counters = job.global_counters()
    digest = job.checksum()
    n_marg = job.merged_marginals() if wl.enrich else 0
    # gather every rank's rows on rank 0 (exchange: disjoint key partitions; dense merge: rank 0 holds everything)
    mine = rows_of(ctr) if job.owns_rows() else np.zeros((0, 3), np.uint64)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        multi_rows = np.concatenate(gathered)
        multi_rows = multi_rows[np.lexsort((multi_rows[:, 1], multi_rows[:, 0]))]
        single = bc.Counter(run, device=local, expected_reads=per_gpu * world)
        whole = [wl.generate_device(run, a, min(batch, per_gpu * world - a), device=dev) for a in range(0, per_gpu * world, batch)]
        torch.cuda.synchronize()  # device batches must stay alive and complete until bc_sync (bc_submit is asynchronous)
        for b in whole:
            single.submit(b)
        want_counters = single.counters()
        want_rows = rows_of(single)
        sjob = Job(bc, single, run, 1, 0, dev, torch.cuda.current_stream(), has_umi, per_gpu * world)
        ok = counters == want_counters and multi_rows.shape == want_rows.shape and bool((multi_rows == want_rows).all()) \
            and n_rows == want_rows.shape[0] and digest == sjob.checksum()
        if ok and wl.enrich and n_marg:
            ok = marg_of(ctr) == marg_of(single)
        print(f"MULTI_GPU_CHECK {name} world={world} reads={per_gpu * world} rows={n_rows} counters={counters} digest={digest} "
              f"{'OK' if ok else 'MISMATCH ' + str(want_counters) + ' rows ' + str(want_rows.shape)}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
