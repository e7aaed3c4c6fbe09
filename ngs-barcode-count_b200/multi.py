"""One whole decode-and-count job over a rank's shard of the reads, for 1..N GPUs (one process per GPU).

Sharding follows SURVEY.md §8(e): reads are independent, so every rank decodes its own contiguous range.
  * scheme without a random barcode: ranks count locally; the (key, count) rows are merged once at the end
    (all-gather of the rows, summed into rank 0's table with bc_import_rows);
  * scheme with a random barcode (UMI): de-duplication has to be global, so bc_decode_route buckets the matched
    (key, UMI) records by owner = hash(key) % N on the device, the buckets are exchanged with an NCCL all-to-all
    and every owner inserts what it received (bc_insert_records); the owner decides matched vs duplicate.
torch.distributed is plumbing only (communicator + buffers); all compute is in the CUDA library.
"""
import os

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


def exchange_records(send, counts, rcounts, recv, world):
    """All-to-all of variable-length record buckets.  send: [world, cap, 2] int64 (bucket r holds counts[r] records for
    rank r); recv: [>= total, 2] int64.  Returns the number of records received (packed at the front of recv).
    Device-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
    dist.all_to_all_single(rcounts, counts)
    sc = counts.cpu().tolist()
    rc = rcounts.cpu().tolist()
    if max(sc) > send.shape[1]:
        raise RuntimeError("route bucket overflow: a bucket received more records than its capacity")
    if sum(rc) > recv.shape[0]:
        raise RuntimeError("route receive buffer too small")
    rank = dist.get_rank()
    ops, off = [], 0
    for r in range(world):
        out = recv[off:off + rc[r]]
        off += rc[r]
        if r == rank:
            out.copy_(send[r, :sc[r]])
            continue
        if rc[r]:
            ops.append(dist.P2POp(dist.irecv, out, r))
        if sc[r]:
            ops.append(dist.P2POp(dist.isend, send[r, :sc[r]], r))
    if ops:  # one grouped NCCL call (ncclGroupStart/End) == an all-to-all-v; plain isend/irecv pairs under gloo
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return off


def gather_rows_to_root(cols, n, rank, world, device):
    """Gathers every rank's row columns (same length n per rank, differing across ranks) to rank 0.
    Returns (parts, sizes): parts[c][r] is column c of rank r (rank 0 only), sizes[r] the row count of rank r."""
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, torch.tensor([n], dtype=torch.int64, device=device))
    sizes = sizes.cpu().tolist()
    parts = []
    for src in cols:
        if rank == 0:
            p = [torch.empty(s, dtype=torch.int64, device=device) for s in sizes]
            _gather_var(src, p, sizes, rank, world)
            parts.append(p)
        else:
            _gather_var(src, None, sizes, rank, world)
    return parts, sizes


def _gather_var(src, parts, sizes, rank, world):
    # variable-length gather as point-to-point sends (dist.gather needs equal sizes)
    if rank == 0:
        parts[0].copy_(src)
        reqs = [dist.irecv(parts[r], src=r) for r in range(1, world) if sizes[r]]
        for q in reqs:
            q.wait()
    elif sizes[rank]:
        dist.send(src.contiguous(), dst=0)


class HostBatch:
    """A bc_batch in pinned host memory (torch pinned tensors viewed as numpy)."""

    def __init__(self, bc, dev_batch):
        pin = lambda t: None if t is None else torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
        self._t = [pin(dev_batch.planes), pin(dev_batch.read_len), pin(dev_batch.qual)]
        np_ = [None if t is None else t.numpy() for t in self._t]
        self.batch = bc.Batch(dev_batch.n, dev_batch.plane_stride, dev_batch.qual_stride, np_[0], np_[1], np_[2], device=False)
        self.n = dev_batch.n
        self.nbytes = sum(t.numel() * t.element_size() for t in self._t if t is not None)


class Job:
    def __init__(self, bc, ctr, run, world, rank, device, stream, has_umi, batch_reads):
        self.bc, self.ctr, self.run = bc, ctr, run
        self.world, self.rank, self.device, self.stream, self.has_umi = world, rank, device, stream, has_umi
        if world == 1:
            self.parallelism = "1 GPU"
        elif has_umi:
            self.parallelism = (f"reads sharded over {world} GPUs; (key,UMI) records stored by the decode kernel into the owner "
                                f"GPU hash(key)%{world} over NVLink peer memory; one tiny NCCL all-gather of counts per batch")
        else:
            self.parallelism = (f"reads sharded over {world} GPUs; tables merged once at the end (in-place all-reduce of the "
                                f"dense count table, or gather of rows for hashed tables)")
        # measurement aid: BC_SPLIT_COUNT=1 uses the routed path on one GPU too (decode and table updates in separate,
        # overlapping kernels instead of one fused kernel)
        self.routed = has_umi and (world > 1 or bool(os.environ.get("BC_SPLIT_COUNT")))
        if self.routed:
            # fused routing: every rank maps every other rank's receive buffer (CUDA IPC over NVLink); the decode
            # kernel stores records straight into the owner's memory.  A (source, owner) region can hold a whole
            # batch, so no key skew can overflow it.
            self.cap = batch_reads
            handles = [None] * world
            mine = ctr.route_open(world, rank, self.cap)
            if world > 1:
                dist.all_gather_object(handles, mine)
            else:
                handles = [mine]
            ctr.route_connect(handles)
            self.counts = torch.zeros((2, world), dtype=torch.int32, device=device)          # what I sent, per parity
            self.all_counts = torch.zeros((2, world * world), dtype=torch.int32, device=device)  # [source][owner]
            self.parity = 0
            if world > 1:
                dist.barrier()

    def to_pinned(self, dev_batch):
        return HostBatch(self.bc, dev_batch)

    def _b(self, b):
        return b.batch if isinstance(b, HostBatch) else b

    def step(self, batches, to_host=False):
        """reset -> decode+count every batch -> rows.  Returns the number of (key, count) rows of the whole job
        (to_host: the rows themselves are also copied to the host, as a drop-in caller would need them)."""
        ctr = self.ctr
        ctr.reset()
        with torch.cuda.stream(self.stream):
            if self.routed:
                for b in batches:
                    self._routed(self._b(b))
            else:
                for b in batches:
                    ctr.submit(self._b(b))
            if self.world > 1 and not self.has_umi:
                return self._merge_rows(to_host)
            if to_host:
                n = ctr.finish_view()[0]
            else:
                n = ctr.export_rows()[3]
            if self.world > 1:
                t = torch.tensor([n], dtype=torch.int64, device=self.device)
                dist.all_reduce(t)
                n = int(t.item())
            return n

    def _routed(self, batch):
        """decode + route one batch.  No host synchronisation: the counts stay on the device; the tiny all-gather is
        the only collective and doubles as the barrier that makes every rank's peer stores visible to the owner."""
        p = self.parity
        self.parity ^= 1
        self.ctr.route_submit(batch, p, self.counts[p])
        if self.world > 1:
            dist.all_gather_into_tensor(self.all_counts[p], self.counts[p])
        else:
            self.all_counts[p].copy_(self.counts[p])
        # records sent to me by source s: all_counts[p][s * world + rank]
        self.ctr.route_insert(p, self.all_counts[p].data_ptr() + 4 * self.rank, self.world, int(batch.n * 1.5))

    def _merge_rows(self, to_host):
        ptr, n_dense = self.ctr.dense_counts()
        if n_dense:  # dense table: one in-place all-reduce, then rank 0 extracts the rows
            dist.all_reduce(dev_tensor(ptr, n_dense, self.device))
            if self.rank == 0:
                n_rows = self.ctr.finish_view()[0] if to_host else self.ctr.export_rows()[3]
            else:
                n_rows = 0
            t = torch.tensor([n_rows], dtype=torch.int64, device=self.device)
            dist.broadcast(t, src=0)
            return int(t.item())
        lo, hi, cnt, n = self.ctr.export_rows()
        wide = hi is not None and hi != 0
        cols = [dev_tensor(p, n, self.device) for p in ((lo, hi, cnt) if wide else (lo, cnt))]
        parts, sizes = gather_rows_to_root(cols, n, self.rank, self.world, self.device)
        if self.rank == 0:
            for r in range(1, self.world):
                if sizes[r]:
                    self.ctr.import_rows(parts[0][r], parts[1][r] if wide else None, parts[-1][r], sizes[r])
            n_rows = self.ctr.finish_view()[0] if to_host else self.ctr.export_rows()[3]
        else:
            n_rows = 0
        t = torch.tensor([n_rows], dtype=torch.int64, device=self.device)
        dist.broadcast(t, src=0)
        return int(t.item())

    def global_counters(self):
        c = self.ctr.counters()
        if self.world > 1:
            names = list(c)
            t = torch.tensor([c[k] for k in names], dtype=torch.int64, device=self.device)
            dist.all_reduce(t)
            c = dict(zip(names, [int(x) for x in t.tolist()]))
        return c
