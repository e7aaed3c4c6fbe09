"""world_size-2 gloo test (CPU) of the live multi-GPU host logic in ngs-barcode-count_b200/multi.py: Job.step ->
Job._exchange drives the library's bc_exchange_* calls and plans where every rank's run lands in every owner's receive
buffer.  The Counter is replaced by a recording fake (no GPU here); the ranks then compare notes: every owner's buffer
must be tiled exactly - no gap, no overlap - by the runs the ranks were told to write, in rank order, and a job whose
owners would overflow the receive capacity must make every rank re-open larger before anything is scattered."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class FakeCounter:
    """Records the exchange calls Job makes; `sent` is what this rank pretends to hold for every owner."""

    def __init__(self, sent):
        self.sent, self.calls, self.cap, self.first, self.received = sent, [], None, None, None

    def profile(self):
        return {"deferred_count": 1}

    def reset(self):
        self.calls.append("reset")

    def submit(self, batch):
        self.calls.append("submit")

    def exchange_open(self, world, rank, capacity):
        self.calls.append(("open", capacity))
        self.cap = capacity

    def exchange_disconnect(self):
        self.calls.append("disconnect")

    def exchange_handle(self):
        return b"h" * 64

    def exchange_connect(self, handles):
        assert len(handles) == 2 and all(len(h) == 64 for h in handles)
        self.calls.append("connect")

    def exchange_count(self, world):
        self.calls.append("count")
        return list(self.sent)

    def exchange_scatter(self, first):
        assert all(f + s <= self.cap for f, s in zip(first, self.sent)), "scatter past the receive capacity"
        self.calls.append("scatter")
        self.first = list(first)

    def exchange_finish(self, received):
        self.calls.append("finish")
        self.received = received

    def finish_view(self):
        return (3, None, None, None)

    def export_rows(self):
        return (0, 0, 0, 3)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ngs_barcode_count_b200  # noqa: F401
    from ngs_barcode_count_b200 import multi
    try:
        # the pure plan first
        first, received, need = multi.exchange_plan([[5, 7], [11, 13]], 1)
        assert (first, received, need) == ([5, 7], 20, 20)
        assert multi.exchange_plan([[5, 7], [11, 13]], 0) == ([0, 0], 16, 20)
        for sent_by_rank, expected in (([[100, 40], [60, 300]], 1000), ([[5000, 10], [7000, 20]], 1000)):
            ctr = FakeCounter(sent_by_rank[rank])
            job = multi.Job(None, ctr, None, world, rank, "cpu", None, True, expected, deferred=True)
            assert job.exchange and ctr.calls[0] == ("open", int(expected * 1.25) + 4096) and ctr.calls[1] == "connect"
            n_rows = job.step(["b0", "b1", "b2"])
            assert n_rows == 3 * world  # the all-reduced row count
            totals = [sum(sent_by_rank[s][o] for s in range(world)) for o in range(world)]
            grew = max(totals) > int(expected * 1.25) + 4096
            want = ["reset", "submit", "submit", "submit", "count"] + \
                (["disconnect", ("open", int(max(totals) * 1.1) + 4096), "connect", "count"] if grew else []) + ["scatter", "finish"]
            assert ctr.calls[2:] == want, ctr.calls
            assert ctr.received == totals[rank]
            # every owner's buffer is tiled by the ranks' runs, in rank order
            plans = [None] * world
            dist.all_gather_object(plans, ctr.first)
            for o in range(world):
                at = 0
                for s in range(world):
                    assert plans[s][o] == at, (plans, o, s)
                    at += sent_by_rank[s][o]
                assert at == totals[o]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_exchange_host_logic_gloo_world2():
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


def test_rows_checksum_is_order_independent_and_additive():
    from ngs_barcode_count_b200.multi import rows_checksum
    g = torch.Generator().manual_seed(3)
    lo = torch.randint(-2**62, 2**62, (1000,), generator=g, dtype=torch.int64)
    hi = torch.randint(0, 2**40, (1000,), generator=g, dtype=torch.int64)
    cnt = torch.randint(1, 1000, (1000,), generator=g, dtype=torch.int64)
    whole = rows_checksum(lo, hi, cnt)
    perm = torch.randperm(1000, generator=g)
    assert rows_checksum(lo[perm], hi[perm], cnt[perm]) == whole
    a, b = rows_checksum(lo[:300], hi[:300], cnt[:300]), rows_checksum(lo[300:], hi[300:], cnt[300:])
    assert [(x + y) & ((1 << 64) - 1) for x, y in zip(a, b)] == whole
    changed = cnt.clone()
    changed[17] += 1
    assert rows_checksum(lo, hi, changed) != whole
    assert rows_checksum(lo[:0], None, cnt[:0]) == [0, 0, 0]
    # narrow keys (no high word) hash like a zero high word
    assert rows_checksum(lo, None, cnt) == rows_checksum(lo, torch.zeros_like(hi), cnt)
