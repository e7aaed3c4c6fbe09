"""One whole decode-and-count job over a rank's shard of the reads, for 1..N GPUs (one process per GPU).

Sharding follows SURVEY.md §8(e): reads are independent, so every rank decodes its own contiguous range with the
single-GPU kernels.  What follows the last batch depends on the counting state of the scheme:
  * dense count table (small index-coded key space, no random barcode — CRISPR): ONE in-place all-reduce of the table;
  * hashed keys (any scheme with a random barcode, raw or large key spaces): ONE exchange of the (key[, UMI]) records —
    record -> owner = hash(key without UMI) % N, written by the library's partitioning kernel straight into the owner's
    receive buffer over NVLink peer memory (bc_exchange_*); every owner then de-duplicates and counts the keys it owns,
    so de-duplication is globally exact.  The only collectives are an all-gather of the N x N count matrix and the
    tiny all-reduce that orders the owners' flush after every rank's scatter.
torch.distributed is plumbing only (communicator + a few integers); all compute is in the CUDA library.  The same steps
inside one process (one context per GPU, no communicator) are `bch_count_fastq_multi` in csrc/host/bc_host.cpp.
"""
import numpy as np
import torch
import torch.distributed as dist


class _DevArray:
    """A device pointer owned by the library, viewed as a torch tensor through the CUDA array interface."""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def dev_tensor(ptr, n, device):
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=device)
    return torch.as_tensor(_DevArray(ptr, n), device=device)


def exchange_plan(matrix, rank):
    """matrix[s][o] = records rank s holds that rank o owns.  Owner o's receive buffer takes the ranks' runs in rank
    order, so this rank's run starts at first[o] = sum of matrix[s][o] over s < rank.
    -> (first, received by this rank, largest total any owner receives)"""
    world = len(matrix)
    first = [sum(int(matrix[s][o]) for s in range(rank)) for o in range(world)]
    totals = [sum(int(matrix[s][o]) for s in range(world)) for o in range(world)]
    return first, totals[rank], max(totals)


_MASK64 = (1 << 64) - 1


def _lsr(x, s):
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix64(x):
    """bc::mix64 on int64 tensors (two's complement wrap-around = arithmetic mod 2^64)."""
    c1 = torch.tensor(0xff51afd7ed558ccd - (1 << 64), dtype=torch.int64, device=x.device)
    c2 = torch.tensor(0xc4ceb9fe1a85ec53 - (1 << 64), dtype=torch.int64, device=x.device)
    x = x ^ _lsr(x, 33)
    x = x * c1
    x = x ^ _lsr(x, 33)
    x = x * c2
    return x ^ _lsr(x, 33)


def rows_checksum(lo, hi, cnt):
    """Order-independent digest of a set of (key, count) rows: (rows, sum of counts, sum of mix64(key) * count), the
    sums mod 2^64.  Digests of disjoint row sets add up to the digest of their union."""
    n = lo.numel()
    if n == 0:
        return [0, 0, 0]
    h = _mix64(lo ^ _mix64(hi + 0x1234567)) if hi is not None else _mix64(lo ^ _mix64(torch.full_like(lo, 0x1234567)))
    return [n, int(cnt.sum().item()) & _MASK64, int((h * cnt).sum().item()) & _MASK64]


class HostBatch:
    """A bc_batch in pinned host memory (torch pinned tensors viewed as numpy)."""

    def __init__(self, bc, dev_batch):
        pin = lambda t: None if t is None else torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
        self._t = [pin(dev_batch.planes), pin(dev_batch.read_len), pin(dev_batch.qual)]
        np_ = [None if t is None else t.numpy() for t in self._t]
        self.batch = bc.Batch(dev_batch.n, dev_batch.plane_stride, dev_batch.qual_stride, np_[0], np_[1], np_[2], device=False)
        self.n = dev_batch.n
        self.nbytes = sum(t.numel() * t.element_size() for t in self._t if t is not None)


class HostWireBatch:
    """The same batch in its transfer form (bc_wire_batch) inside one pinned buffer: what a host that packs for the PCIe
    crossing hands to bc_submit_wire."""

    def __init__(self, bc, dev_batch, max_read_len):
        host = bc.Batch(dev_batch.n, dev_batch.plane_stride, dev_batch.qual_stride, dev_batch.planes.cpu().numpy(),
                        dev_batch.read_len.cpu().numpy(), None if dev_batch.qual is None else dev_batch.qual.cpu().numpy(), device=False)
        need = bc.WireBatch.bound(dev_batch.n, max_read_len, dev_batch.qual is not None)
        self._t = torch.empty(max(need, 1), dtype=torch.uint8, pin_memory=True)
        self.batch = bc.WireBatch(host, max_read_len, buf=self._t.numpy())
        self.n = dev_batch.n
        self.nbytes = self.batch.nbytes
        self.qual_bits = self.batch.c.qual_bits
        self.n_calls_listed = self.batch.c.n_calls if not self.batch.c.nmask else None


class Job:
    """ctr: this rank's Counter (or any object with the same methods: the CPU test drives the host logic with a fake).
    expected_reads: reads per rank, sizes the exchange's receive buffer (it grows when a job needs more)."""

    def __init__(self, bc, ctr, run, world, rank, device, stream, has_umi, expected_reads, deferred=None):
        self.bc, self.ctr, self.run = bc, ctr, run
        self.world, self.rank, self.device, self.stream, self.has_umi = world, rank, device, stream, has_umi
        self.deferred = bool(ctr.profile()["deferred_count"]) if deferred is None else deferred
        self.exchange = world > 1 and self.deferred
        if world == 1:
            self.parallelism = "1 GPU"
        elif self.exchange:
            self.parallelism = (f"reads sharded over {world} GPUs; every batch's (key,UMI) records leave for owner = hash(key) % {world} "
                                f"right after its decode: a scatter kernel on a side stream writes them into the owner GPU's memory over "
                                f"NVLink (runs reserved with one atomic on the owner's receive cursor); each owner de-duplicates and "
                                f"counts its keys; bulk exchange after the last batch when a receive buffer turns out too small")
        else:
            self.parallelism = (f"reads sharded over {world} GPUs; one in-place all-reduce of the dense count table at the end")
        if self.exchange:
            self.cap = 0
            self._open(int(expected_reads * 1.25) + 4096)

    def _open(self, capacity):
        """(re)allocate the receive buffers and connect every rank to every other (collective)."""
        if self.cap:  # growing: nobody frees a buffer that another rank still maps
            self.ctr.exchange_disconnect()
            dist.barrier()
        self.ctr.exchange_open(self.world, self.rank, capacity)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.ctr.exchange_handle())
        self.ctr.exchange_connect(handles)
        self.cap = capacity
        dist.barrier()

    def to_pinned(self, dev_batch, wire=False):
        return HostWireBatch(self.bc, dev_batch, self.run.max_read_len) if wire else HostBatch(self.bc, dev_batch)

    def _b(self, b):
        return b.batch if isinstance(b, (HostBatch, HostWireBatch)) else b

    def step(self, batches, to_host=False):
        """reset -> decode every batch -> merge across ranks -> rows.  Returns the number of (key, count) rows of the
        whole job (to_host: this rank's rows are also copied to the host, as a drop-in caller would need them)."""
        ctr = self.ctr
        ctr.reset()
        with torch.cuda.stream(self.stream):
            for b in batches:
                ctr.submit(self._b(b))
            if self.exchange:
                self._exchange()
            elif self.world > 1:
                return self._merge_dense(to_host)
            n = ctr.finish_view()[0] if to_host else ctr.export_rows()[3]
            if self.world > 1:
                t = torch.tensor([n], dtype=torch.int64, device=self.device)
                dist.all_reduce(t)
                n = int(t.item())
            return n

    def _exchange(self):
        """The one exchange step of a job with hashed keys.  Every rank computes the same plan from the same matrix."""
        sent = self.ctr.exchange_count(self.world)
        mine = torch.tensor(sent, dtype=torch.int64, device=self.device)
        matrix = torch.empty(self.world * self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(matrix, mine)
        matrix = matrix.view(self.world, self.world).cpu().tolist()
        first, received, need = exchange_plan(matrix, self.rank)
        if need > self.cap:  # a skewed job (hot keys): every rank sees the same matrix and grows together
            self._open(int(need * 1.1) + 4096)
            assert self.ctr.exchange_count(self.world) == sent  # a re-opened exchange starts over; the records have not changed
        self.ctr.exchange_scatter(first)
        # orders the owners' flush after every rank's scatter: a rank leaves the all-reduce only once all have entered it,
        # and each enters it on the stream its scatter kernel runs on
        dist.all_reduce(torch.zeros(1, dtype=torch.int32, device=self.device))
        self.ctr.exchange_finish(received)
        self.last_matrix = matrix

    def _merge_dense(self, to_host):
        ptr, n_dense = self.ctr.dense_counts()
        if not n_dense:
            raise RuntimeError("multi-GPU jobs need deferred counting (exchange) or a dense count table (all-reduce); "
                               "BC_CFG_INLINE_COUNT hash tables are a single-GPU measurement aid")
        dist.all_reduce(dev_tensor(ptr, n_dense, self.device))
        if self.rank == 0:  # the merged table is on every rank; rank 0 extracts the rows
            n_rows = self.ctr.finish_view()[0] if to_host else self.ctr.export_rows()[3]
        else:
            n_rows = 0
        t = torch.tensor([n_rows], dtype=torch.int64, device=self.device)
        dist.broadcast(t, src=0)
        return int(t.item())

    def owns_rows(self):
        """whether this rank's rows are part of the job's result (exchange: every owner; dense merge: rank 0 only)"""
        return self.world == 1 or self.exchange or self.rank == 0

    def checksum(self):
        """Digest of the job's final rows, summed over the ranks (see rows_checksum)."""
        d = [0, 0, 0]
        if self.owns_rows():
            lo, hi, cnt, n = self.ctr.export_rows()
            with torch.cuda.stream(self.stream):
                d = rows_checksum(dev_tensor(lo, n, self.device), dev_tensor(hi, n, self.device) if hi else None,
                                  dev_tensor(cnt, n, self.device))
        if self.world > 1:
            t = torch.tensor([x - (1 << 64) if x >= (1 << 63) else x for x in d], dtype=torch.int64, device=self.device)
            dist.all_reduce(t)
            d = [int(x) & _MASK64 for x in t.tolist()]
        return d

    def merged_marginals(self):
        """Enrichment marginals of the whole job (dense counters summed over the owners); rank 0 then calls enrich()."""
        ptr, n = self.ctr.marginals()
        if self.world > 1 and self.exchange and n:
            with torch.cuda.stream(self.stream):
                dist.all_reduce(dev_tensor(ptr, n, self.device))
            torch.cuda.current_stream().synchronize()
            self.stream.synchronize()
        return n

    def global_counters(self):
        c = self.ctr.counters()
        if self.world > 1:
            names = list(c)
            t = torch.tensor([c[k] for k in names], dtype=torch.int64, device=self.device)
            dist.all_reduce(t)
            c = dict(zip(names, [int(x) for x in t.tolist()]))
        return c
