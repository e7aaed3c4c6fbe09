"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in ngs-barcode-count_b200/multi.py: the variable-length
record exchange used for hash-routed UMI de-duplication and the row gather used for the final table merge."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ngs_barcode_count_b200  # noqa: F401
    from ngs_barcode_count_b200 import multi
    try:
        g = torch.Generator().manual_seed(100 + rank)
        cap = 64
        # records: lo = global unique id, hi = owner rank; bucket r of this rank holds records owned by r
        counts = torch.tensor([int(torch.randint(0, cap, (1,), generator=g)) for _ in range(world)], dtype=torch.int32)
        if rank == 0:
            counts[1] = 0  # an empty bucket
        send = torch.full((world, cap, 2), -1, dtype=torch.int64)
        for r in range(world):
            n = int(counts[r])
            send[r, :n, 0] = torch.arange(n) + 1000 * rank + 100000 * r
            send[r, :n, 1] = r
        recv = torch.full((world * cap, 2), -7, dtype=torch.int64)
        rcounts = torch.zeros(world, dtype=torch.int32)
        n = multi.exchange_records(send, counts, rcounts, recv, world)
        got = recv[:n]
        assert bool((got[:, 1] == rank).all()), "a record landed on a rank that does not own it"
        # every rank reports what it sent / received; rank 0 checks the union is preserved
        allc = [torch.zeros(world, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(allc, counts)
        want = sorted(int(x) for src in range(world) for x in (torch.arange(int(allc[src][rank])) + 1000 * src + 100000 * rank))
        assert sorted(int(x) for x in got[:, 0]) == want
        # overflow is an error, not silent loss
        bad = counts.clone()
        bad[0] = cap + 1
        try:
            multi.exchange_records(send, bad, rcounts, recv, world)
            raised = False
        except RuntimeError:
            raised = True
        assert raised
        dist.barrier()
        # row gather with different lengths per rank (rank 1 has none)
        nrows = 5 if rank == 0 else 0
        cols = [torch.arange(nrows, dtype=torch.int64) + 10 * c + 100 * rank for c in range(3)]
        parts, sizes = multi.gather_rows_to_root(cols, nrows, rank, world, "cpu")
        assert sizes == [5, 0]
        if rank == 0:
            assert [p[0].tolist() for p in parts] == [c.tolist() for c in cols]
        nrows = 3 + rank
        cols = [torch.arange(nrows, dtype=torch.int64) + 10 * c + 100 * rank for c in range(3)]
        parts, sizes = multi.gather_rows_to_root(cols, nrows, rank, world, "cpu")
        if rank == 0:
            assert sizes == [3, 4]
            assert parts[2][1].tolist() == [120, 121, 122, 123]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_exchange_and_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
