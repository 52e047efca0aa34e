"""Multi-GPU `triangle` (one process per GPU, torch.distributed for the plumbing).

The reference is a single process (rayon threads); the pair loop of src/triangle.rs:71-105 shards naturally:
  1. every rank sketches a contiguous block of genomes (seeding is independent per genome);
  2. ONE exchange step: the ranks' sketch blocks are all-gathered (NCCL over NVLink/NVSwitch) so every GPU holds all
     N sketches in rank-major = global genome order (10k genomes ~ 12 GB; 4.4 GB of seeds + views);
  3. every rank screens the triangle (cheap) and chains every world-th passing pair, no further communication
     (sk_screen_triangle_rows offers a row-cyclic partition of the screen itself for sets where screening dominates);
  4. results stay on their rank (the reference's sparse output order is nondeterministic anyway).
"""
import ctypes as C

import numpy as np

from . import host as H


def shard_range(n_items, world, rank):
    """Contiguous block partition used for the seeding stage."""
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def rows_of_rank(n_rows, world, rank):
    """Row-cyclic partition of the triangle's rows (row i has N-1-i columns: cyclic interleave balances to < 1/N)."""
    return range(rank, n_rows, world)


def pairs_of_rank(sorted_pairs, world, rank):
    """Cyclic split of the sorted passing-pair list: every rank gets the same number of pairs (+-1)."""
    return np.ascontiguousarray(sorted_pairs[rank::world])


def gather_variable(dist, local, world, device):
    """All-gather 1-D tensors of different lengths: returns the list of per-rank tensors (views into one buffer).
    Works on any backend (gloo on CPU for tests, nccl on GPUs)."""
    import torch
    n = torch.tensor([local.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    padded = torch.zeros(mx, dtype=local.dtype, device=device)
    padded[:local.numel()] = local
    out = torch.empty(world * mx, dtype=local.dtype, device=device)
    if hasattr(dist, "all_gather_into_tensor") and device != "cpu" and str(device) != "cpu":
        dist.all_gather_into_tensor(out, padded)
    else:
        parts = [out[r * mx:(r + 1) * mx] for r in range(world)]
        dist.all_gather(parts, padded)
    return [out[r * mx:r * mx + sizes[r]] for r in range(world)]


class DistTriangle:
    def __init__(self, ctx, world, rank, sp, mp):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ctx, self.world, self.rank, self.sp, self.mp = ctx, world, rank, sp, mp
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.last_results = []

    def exchange(self, local_set):
        """All-gather the ranks' sketch sets -> one set holding every genome, on this GPU."""
        import os
        import time
        torch, L, ctx = self.torch, self.ctx.L, self.ctx
        trace = os.environ.get("SK_TRACE") and self.rank == 0
        t0 = time.perf_counter()
        nbytes, nwords = C.c_uint64(), C.c_uint64()
        ctx.check(L.sk_sketch_set_blob_size(local_set.h, C.byref(nbytes), C.byref(nwords)))
        meta = np.zeros(nwords.value, np.uint64)
        # sizes first (tiny), so the blob can be packed straight into its slot of the gathered buffer (no padded copy)
        sz = torch.tensor([nbytes.value, nwords.value], dtype=torch.int64, device=self.device)
        allsz = torch.empty(2 * self.world, dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(allsz, sz)
        allsz = allsz.cpu().numpy().reshape(self.world, 2)
        mx = int(allsz[:, 0].max())
        mx = (mx + 255) & ~255
        out = torch.empty(self.world * mx, dtype=torch.uint8, device=self.device)
        mine = out[self.rank * mx:(self.rank + 1) * mx]
        ctx.check(L.sk_sketch_set_pack(local_set.h, mine.data_ptr(), meta.ctypes.data))
        t1 = time.perf_counter()
        self.dist.all_gather_into_tensor(out, mine)              # in place: rank r's slice is already at offset r * mx
        mw = int(allsz[:, 1].max())
        mloc = torch.zeros(mw, dtype=torch.int64, device=self.device)
        mloc[:nwords.value] = torch.from_numpy(meta.view(np.int64)).to(self.device)
        mall = torch.empty(self.world * mw, dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(mall, mloc)
        mall = mall.cpu().numpy().view(np.uint64).reshape(self.world, mw)
        t2 = time.perf_counter()
        metas = [np.ascontiguousarray(mall[r, :int(allsz[r, 1])]) for r in range(self.world)]
        bp = (C.c_void_p * self.world)(*[out.data_ptr() + r * mx for r in range(self.world)])
        mp_ = (C.c_void_p * self.world)(*[m.ctypes.data for m in metas])
        res = C.c_void_p()
        ctx.check(L.sk_sketch_set_unpack(ctx.h, self.world, bp, mp_, C.byref(res)))
        if trace:
            print("[multi_gpu rank0] exchange: sizes+pack %.1f ms, all-gather (%.2f GB) %.1f ms, unpack+tables %.1f ms" %
                  ((t1 - t0) * 1e3, self.world * mx / 1e9, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3), flush=True)
        return H.SketchSet(ctx, res)

    def step(self, host_bases, dev_ptr, off, goc, nloc, g0, n_total):
        """One whole triangle over all ranks; returns this rank's number of kept pairs."""
        import os
        import time
        ctx, L = self.ctx, self.ctx.L
        trace = os.environ.get("SK_TRACE") and self.rank == 0
        t0 = time.perf_counter()
        if host_bases is not None:
            local = H.sketch_contigs(ctx, host_bases, off, goc, nloc, self.sp)
        else:
            local = H.sketch_contigs(ctx, None, off, goc, nloc, self.sp, device_ptr=dev_ptr)
        t1 = time.perf_counter()
        allset = self.exchange(local)
        local.free()
        assert len(allset) == n_total
        t2 = time.perf_counter()
        # every rank screens the whole triangle (a few ms: one sort of all markers) and takes every world-th passing pair of
        # the sorted list: an even split of the CHAINING work, which a row partition does not give when related genomes are
        # adjacent (rows i mod 8 of a 20-genome cluster carry 33 vs 16 pairs)
        pairs = pairs_of_rank(H.screen_triangle(ctx, allset, self.mp), self.world, self.rank)
        t3 = time.perf_counter()
        res = H.chain_pairs(ctx, allset, allset, pairs, self.mp, as_array=True)
        allset.free()
        if trace:
            print("[multi_gpu rank0] sketch %.1f ms  exchange %.1f ms  screen %.1f ms  chain %.1f ms (%d pairs)" %
                  ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (time.perf_counter() - t3) * 1e3, len(pairs)), flush=True)
        self.last_results = res[res["ani"] > 0.1]               # src/triangle.rs:99 (numpy structured array)
        return len(self.last_results)
