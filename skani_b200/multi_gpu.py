"""Multi-GPU `triangle` (one process per GPU, torch.distributed for the plumbing; the one-process form of the same
scheme is sk_triangle_multi, skani_b200/csrc/multi.cu).

The reference is a single process (rayon threads); the pair loop of src/triangle.rs:71-105 shards naturally:
  1. every rank owns a contiguous block of genomes and runs the pipelined single-GPU triangle on it (sk_triangle_local:
     upload || seed || screen || chain of the pairs INSIDE the block) and keeps the block's sketch set;
  2. the MARKERS of every genome (8 B x L/1000 per genome: 0.4 GB for 10k genomes) are all-gathered and every rank
     screens the whole triangle (a few ms) -> the same sorted list of passing pairs on every rank; pairs inside one
     block are dropped (done in step 1);
  3. the remaining CROSS-BLOCK pairs are cut into `world` contiguous slices (equal chaining work by construction); a rank
     needs the full sketches of just the genomes its slice touches; they arrive in ONE variable all-to-all (NCCL over
     NVLink/NVSwitch) of per-destination sub-blobs that carry the k-mer hash tables along (sk_sketch_set_pack_subset,
     SK_PACK_TABLES: nothing is rebuilt on arrival).  With related genomes adjacent in the input there are hardly any
     cross-block pairs; with a random input order (1 - 1/world) of all pairs are cross-block and the exchange approaches
     the volume of a full all-gather, never more;
  4. each rank rebuilds a working set from what it received (rank-major = ascending global genome order), chains its
     slice, maps the ids back to global ones.  Results stay on their rank (the reference's sparse output order is
     nondeterministic anyway).
`exchange()` (full all-gather of the sketches) is kept for callers that want every sketch everywhere (search/dist DBs).
"""
import ctypes as C
import os
import time

import numpy as np

from . import host as H

PACK_MARKERS_ONLY = 1        # include/skani_b200.h SK_PACK_MARKERS_ONLY
PACK_TABLES = 2              # include/skani_b200.h SK_PACK_TABLES


def shard_range(n_items, world, rank):
    """Contiguous block partition used for the seeding stage (and for the slices of the sorted pair list)."""
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def rows_of_rank(n_rows, world, rank):
    """Row-cyclic partition of the triangle's rows (row i has N-1-i columns: cyclic interleave balances to < 1/N)."""
    return range(rank, n_rows, world)


def pairs_of_rank(sorted_pairs, world, rank):
    """Cyclic split of the sorted passing-pair list: every rank gets the same number of pairs (+-1)."""
    return np.ascontiguousarray(sorted_pairs[rank::world])


def partition_pairs(sorted_pairs, world):
    """Split the sorted cross-block pair list over the ranks so that a rank has to FETCH little: the pairs of one connected
    component of the pair graph (a cluster of related genomes) stay together, so the rank needs that cluster's genomes once --
    with a genome order unrelated to relatedness a contiguous slice of the sorted list needs ~5x more sketches (8 GPUs, 10 k
    genomes: 6 800 instead of ~1 300 per rank).  Components larger than half a rank's fair share are cut into runs of
    consecutive pairs; items go to the least loaded rank, largest first (deterministic: every rank computes the same split).
    Returns `world` sorted uint64 arrays."""
    p = np.ascontiguousarray(sorted_pairs, np.uint64)
    n = len(p)
    if world <= 1 or n == 0:
        return [p] + [p[:0] for _ in range(max(world, 1) - 1)]
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    i = (p >> np.uint64(32)).astype(np.int64); j = (p & np.uint64(0xFFFFFFFF)).astype(np.int64)
    ids, inv = np.unique(np.concatenate([i, j]), return_inverse=True)
    a, b = inv[:n], inv[n:]
    ncomp, lab = connected_components(csr_matrix((np.ones(n, np.int8), (a, b)), shape=(len(ids), len(ids))), directed=False)
    comp = lab[a]
    first = np.full(ncomp, n, np.int64)
    np.minimum.at(first, comp, np.arange(n))                   # component -> index of its first pair in the sorted list
    order = np.lexsort((np.arange(n), first[comp]))            # pairs grouped by component (in order of first appearance), sorted inside
    counts = np.bincount(comp, minlength=ncomp)[np.argsort(first, kind="stable")]
    cap = max(1, -(-n // (2 * world)))
    items = []                                                 # (-size, start in `order`)
    pos = 0
    for cnt in counts.tolist():
        for s0 in range(0, cnt, cap):
            items.append((-min(cap, cnt - s0), pos + s0))
        pos += cnt
    items.sort()
    import heapq
    heap = [(0, r) for r in range(world)]
    chunks = [[] for _ in range(world)]
    for neg, start in items:
        load, r = heapq.heappop(heap)
        chunks[r].append(order[start:start - neg])
        heapq.heappush(heap, (load - neg, r))
    return [np.sort(p[np.concatenate(c)]) if c else p[:0] for c in chunks]


def pair_slice_of_rank(sorted_pairs, world, rank):
    """This rank's part of the cross-block pair list (see partition_pairs)."""
    return partition_pairs(sorted_pairs, world)[rank]


def genomes_of_pairs(pairs):
    """Ascending distinct genome ids that appear in a list of (i << 32 | j) pairs."""
    p = np.asarray(pairs, np.uint64)
    if len(p) == 0:
        return np.zeros(0, np.uint32)
    return np.unique(np.concatenate([(p >> np.uint64(32)).astype(np.uint32), (p & np.uint64(0xFFFFFFFF)).astype(np.uint32)]))


def fetch_plan(sorted_pairs, world, rank, bounds, parts=None):
    """Who needs what.  bounds[r] .. bounds[r+1] = the genome block sketched by rank r.
    Returns (need, send, recv_counts): need = ascending global ids this rank chains; send[d] = LOCAL indices (into this
    rank's block) of the genomes rank d needs from here; recv_counts[r] = how many genomes arrive from rank r.
    parts = partition_pairs(sorted_pairs, world) if the caller has it already."""
    bounds = np.asarray(bounds, np.int64)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    if parts is None:
        parts = partition_pairs(sorted_pairs, world)
    send, need = [], None
    for d in range(world):
        nd = genomes_of_pairs(parts[d])
        if d == rank:
            need = nd
        a, b = np.searchsorted(nd, [lo, hi])
        send.append((nd[a:b] - lo).astype(np.uint32))
    cut = np.searchsorted(need, bounds)
    recv_counts = np.diff(cut).astype(np.int64)
    return need, send, recv_counts


def allgather_sorted_u64(dist, torch, part, world, device):
    """All ranks contribute a uint64 array of any length; everyone gets the sorted concatenation (two collectives: lengths,
    then the arrays padded to the longest).  `device` = where the collective's tensors live (cuda for NCCL, cpu for gloo)."""
    part = np.ascontiguousarray(part, np.uint64)
    cnt = torch.tensor([len(part)], dtype=torch.int64, device=device)
    allcnt = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allcnt, cnt)
    allcnt = allcnt.cpu().numpy()
    mx = int(allcnt.max()) if world else 0
    if mx == 0:
        return np.zeros(0, np.uint64)
    buf = np.zeros(mx, np.int64)
    buf[:len(part)] = part.view(np.int64)
    mine = torch.from_numpy(buf).to(device)
    allbuf = torch.empty(world * mx, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allbuf, mine)
    allbuf = allbuf.cpu().numpy().view(np.uint64).reshape(world, mx)
    out = np.concatenate([allbuf[r, :int(allcnt[r])] for r in range(world)])
    out.sort()
    return out


def cross_block_pairs(sorted_pairs, bounds):
    """The pairs whose two genomes live in different blocks (bounds[r] .. bounds[r+1] = block of rank r)."""
    p = np.asarray(sorted_pairs, np.uint64)
    if len(p) == 0:
        return p
    b = np.asarray(bounds, np.int64)[1:]
    bi = np.searchsorted(b, (p >> np.uint64(32)).astype(np.int64), side="right")
    bj = np.searchsorted(b, (p & np.uint64(0xFFFFFFFF)).astype(np.int64), side="right")
    return np.ascontiguousarray(p[bi != bj])


def remap_pairs(pairs, need):
    """(i << 32 | j) in global ids -> the same pairs in indices of the ascending id list `need`."""
    p = np.asarray(pairs, np.uint64)
    i = np.searchsorted(need, (p >> np.uint64(32)).astype(np.uint32)).astype(np.uint64)
    j = np.searchsorted(need, (p & np.uint64(0xFFFFFFFF)).astype(np.uint32)).astype(np.uint64)
    return np.ascontiguousarray((i << np.uint64(32)) | j)


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


def alltoall_variable(dist, send_parts, world, device, src=None):
    """Variable all-to-all of 1-D tensors (send_parts[d] goes to rank d); returns the list of tensors received, by source
    rank.  One tiny all-to-all for the sizes, one for the payload.  Any backend (gloo on CPU for tests, nccl on GPUs).
    src: the buffer the parts are consecutive views of, if they are (saves the concatenating copy)."""
    import torch
    dtype = send_parts[0].dtype
    n_in = torch.tensor([p.numel() for p in send_parts], dtype=torch.int64, device=device)
    n_out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(n_out, n_in)
    in_split = [int(p.numel()) for p in send_parts]
    out_split = [int(x) for x in n_out.cpu().tolist()]
    if src is None:
        src = torch.cat(send_parts) if sum(in_split) else torch.empty(0, dtype=dtype, device=device)
    dst = torch.empty(sum(out_split), dtype=dtype, device=device)
    dist.all_to_all_single(dst, src, output_split_sizes=out_split, input_split_sizes=in_split)
    outs, o = [], 0
    for r in range(world):
        outs.append(dst[o:o + out_split[r]])
        o += out_split[r]
    return outs


class DistTriangle:
    def __init__(self, ctx, world, rank, sp, mp, name_ranks=None):
        """name_ranks: optional per-GLOBAL-genome order of the file names (see SketchSet.set_name_ranks); default = index."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ctx, self.world, self.rank, self.sp, self.mp = ctx, world, rank, sp, mp
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.name_ranks = None if name_ranks is None else np.ascontiguousarray(name_ranks, np.uint64)
        self.last_results = []
        self.last_need = None

    # ---- full exchange: every rank ends up with every genome (markers only, or whole sketches) -------------------
    def exchange(self, local_set, flags=0):
        """All-gather the ranks' sketch sets -> one set holding every genome (rank-major order), on this GPU."""
        torch, L, ctx = self.torch, self.ctx.L, self.ctx
        trace = os.environ.get("SK_TRACE") and self.rank == 0
        t0 = time.perf_counter()
        nbytes, nwords = local_set.subset_blob_size(None, flags)
        # sizes first (tiny), so the blob can be packed straight into its slot of the gathered buffer (no padded copy)
        sz = torch.tensor([nbytes, nwords], dtype=torch.int64, device=self.device)
        allsz = torch.empty(2 * self.world, dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(allsz, sz)
        allsz = allsz.cpu().numpy().reshape(self.world, 2)
        mx = int(allsz[:, 0].max())
        mx = (mx + 255) & ~255
        out = torch.empty(self.world * mx, dtype=torch.uint8, device=self.device)
        mine = out[self.rank * mx:(self.rank + 1) * mx]
        meta = local_set.pack_subset(None, flags, mine.data_ptr(), nwords)
        t1 = time.perf_counter()
        self.dist.all_gather_into_tensor(out, mine)              # in place: rank r's slice is already at offset r * mx
        mw = int(allsz[:, 1].max())
        mloc = torch.zeros(mw, dtype=torch.int64, device=self.device)
        mloc[:nwords] = torch.from_numpy(meta.view(np.int64)).to(self.device)
        mall = torch.empty(self.world * mw, dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(mall, mloc)
        mall = mall.cpu().numpy().view(np.uint64).reshape(self.world, mw)
        t2 = time.perf_counter()
        metas = [np.ascontiguousarray(mall[r, :int(allsz[r, 1])]) for r in range(self.world)]
        res = self._unpack([out.data_ptr() + r * mx for r in range(self.world)], metas)
        if trace:
            print("[multi_gpu rank0] exchange(flags=%d): sizes+pack %.1f ms, all-gather (%.2f GB) %.1f ms, unpack+tables %.1f ms" %
                  (flags, (t1 - t0) * 1e3, self.world * mx / 1e9, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3), flush=True)
        return res

    def _unpack(self, blob_ptrs, metas):
        ctx, L = self.ctx, self.ctx.L
        n = len(blob_ptrs)
        bp = (C.c_void_p * n)(*blob_ptrs)
        mp_ = (C.c_void_p * n)(*[m.ctypes.data for m in metas])
        res = C.c_void_p()
        ctx.check(L.sk_sketch_set_unpack(ctx.h, n, bp, mp_, C.byref(res)))
        return H.SketchSet(ctx, res)

    # ---- partial exchange: every rank receives the sketches of the genomes its pair slice touches ----------------
    def fetch(self, local_set, send, recv_counts):
        """send[d] = local genome indices for rank d (ascending).  Returns the working set: the received genomes in
        source-rank order (= ascending global order).  Sub-blobs carry the k-mer tables (no rebuild on arrival)."""
        torch = self.torch
        sizes = [local_set.subset_blob_size(send[d], PACK_TABLES) for d in range(self.world)]
        blobs = torch.empty(sum(s[0] for s in sizes), dtype=torch.uint8, device=self.device)
        parts, metas, o = [], [], 0
        for d in range(self.world):
            part = blobs[o:o + sizes[d][0]]
            metas.append(local_set.pack_subset(send[d], PACK_TABLES, part.data_ptr(), sizes[d][1]))
            parts.append(part)
            o += sizes[d][0]
        got = alltoall_variable(self.dist, parts, self.world, self.device, src=blobs)
        mparts = [torch.from_numpy(m.view(np.int64)).to(self.device) for m in metas]
        gmeta = [g.cpu().numpy().view(np.uint64) for g in alltoall_variable(self.dist, mparts, self.world, self.device)]
        for r in range(self.world):
            assert int(gmeta[r][0]) == int(recv_counts[r]), "fetch plan mismatch between ranks"
        # all_to_all_single's output is one contiguous tensor; sub-blob sizes are multiples of 256 bytes, so alignment holds
        work = self._unpack([g.data_ptr() for g in got], [np.ascontiguousarray(m) for m in gmeta])
        return work, sum(int(g.numel()) for r, g in enumerate(got) if r != self.rank)

    def step(self, host_bases, dev_ptr, off, goc, nloc, g0, n_total, packed=None):
        """One whole triangle over all ranks; returns this rank's number of kept pairs (results in self.last_results,
        ref_id / query_id = GLOBAL genome indices)."""
        ctx = self.ctx
        torch = self.torch
        trace = os.environ.get("SK_TRACE") and self.rank == 0
        t0 = time.perf_counter()
        ranks_local = None if self.name_ranks is None else self.name_ranks[g0:g0 + nloc]
        if packed is not None:          # (units, nmask, contig_len): the block's genomes already 2-bit packed on the host
            res_local, local, _st = H.triangle_2bit(ctx, packed[0], packed[1], packed[2], goc, nloc, self.sp, self.mp, name_ranks=ranks_local, keep_set=True)
        elif host_bases is not None:    # pipelined: upload || seed || screen || chain inside the block
            res_local, local, _st = H.triangle_local(ctx, host_bases, off, goc, nloc, self.sp, self.mp, name_ranks=ranks_local)
        else:                           # sequences already resident on the device: sketch -> screen -> chain on one stream
            local = H.sketch_contigs(ctx, None, off, goc, nloc, self.sp, device_ptr=dev_ptr)
            if ranks_local is not None:
                local.set_name_ranks(ranks_local)
            lp = H.screen_triangle(ctx, local, self.mp)
            res_local = H.chain_pairs(ctx, local, local, lp, self.mp, as_array=True)
            res_local = res_local[res_local["ani"] > 0.1]               # src/triangle.rs:99
        res_local["ref_id"] += np.uint32(g0)
        res_local["query_id"] += np.uint32(g0)
        t1 = time.perf_counter()
        # genome blocks of all ranks
        blk = torch.tensor([g0, nloc], dtype=torch.int64, device=self.device)
        allblk = torch.empty(2 * self.world, dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(allblk, blk)
        allblk = allblk.cpu().numpy().reshape(self.world, 2)
        bounds = np.concatenate([allblk[:, 0], [allblk[-1, 0] + allblk[-1, 1]]])
        assert bounds[0] == 0 and bounds[-1] == n_total and np.all(np.diff(bounds) == allblk[:, 1]), "ranks must hold consecutive blocks"
        # markers everywhere -> SHARDED screen: this rank screens the rows of its own block against every genome before them
        # (pairs inside the block are already done) and the partial cross-block pair lists are all-gathered and merged, so every
        # rank holds the same sorted list
        mk = self.exchange(local, PACK_MARKERS_ONLY)
        assert len(mk) == n_total
        part = H.screen_triangle_block(ctx, mk, g0, g0 + nloc, self.mp)
        mk.free()
        part = part[(part >> np.uint64(32)) < np.uint64(g0)]
        pairs = allgather_sorted_u64(self.dist, torch, part, self.world, self.device)
        t2 = time.perf_counter()
        parts = partition_pairs(pairs, self.world)
        need, send, recv_counts = fetch_plan(pairs, self.world, self.rank, bounds, parts)
        mine = parts[self.rank]
        work, remote_bytes = self.fetch(local, send, recv_counts)
        local.free()
        assert len(work) == len(need)
        t3 = time.perf_counter()
        if len(mine):
            work.set_name_ranks(need.astype(np.uint64) if self.name_ranks is None else self.name_ranks[need])
            res = H.chain_pairs(ctx, work, work, remap_pairs(mine, need), self.mp, as_array=True)
            res["ref_id"] = need[res["ref_id"]]
            res["query_id"] = need[res["query_id"]]
            res = np.concatenate([res_local, res[res["ani"] > 0.1]])
        else:
            res = res_local
        work.free()
        if trace:
            print("[multi_gpu rank0] block triangle %.1f ms (%d pairs kept)  markers+screen %.1f ms  fetch %.1f ms (%d genomes, %.1f MB remote)  "
                  "chain %.1f ms (%d cross-block pairs)" % ((t1 - t0) * 1e3, len(res_local), (t2 - t1) * 1e3, (t3 - t2) * 1e3, len(need),
                                                            remote_bytes / 1e6, (time.perf_counter() - t3) * 1e3, len(mine)), flush=True)
        self.last_results = res
        self.last_need = need
        return len(self.last_results)
