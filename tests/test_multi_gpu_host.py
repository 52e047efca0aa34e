"""Host-side logic of the multi-GPU path on CPU: partitions, and the variable-length all-gather over gloo with
world_size 2 (the same code path NCCL takes on the GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_partitions_cover_everything():
    from skani_b200.multi_gpu import shard_range, rows_of_rank, pairs_of_rank
    import numpy as np
    pl = np.arange(95000, dtype=np.uint64)
    parts = [pairs_of_rank(pl, 8, r) for r in range(8)]
    assert sorted(np.concatenate(parts).tolist()) == pl.tolist() and max(map(len, parts)) - min(map(len, parts)) <= 1
    for n in (0, 1, 7, 10000):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
            rows = sorted(i for r in range(world) for i in rows_of_rank(max(n - 1, 0), world, r))
            assert rows == list(range(max(n - 1, 0)))
    # row-cyclic balance of the pair counts for the 10k triangle: every rank within 0.2% of the mean
    n, world = 10000, 8
    loads = [sum(n - 1 - i for i in rows_of_rank(n - 1, world, r)) for r in range(world)]
    assert sum(loads) == n * (n - 1) // 2
    assert (max(loads) - min(loads)) / (sum(loads) / world) < 2e-3


def test_fetch_plan_is_consistent_across_ranks():
    """Every rank computes the plan from the same sorted pair list: what rank r plans to send to d must be exactly what
    d expects from r, the union over r must be d's need list, and the remapped pairs must point back at the same ids."""
    from skani_b200.multi_gpu import fetch_plan, pair_slice_of_rank, remap_pairs, genomes_of_pairs, shard_range
    rng = np.random.default_rng(7)
    for n, world, clustered in ((200, 2, True), (1000, 4, True), (333, 8, False), (50, 3, True), (10, 4, False)):
        if clustered:      # clusters of 20 consecutive genomes (the bench workload), all pairs inside a cluster
            pl = [(i << 32) | j for i in range(n) for j in range(i + 1, min(n, (i // 20 + 1) * 20))]
        else:              # relatedness unrelated to the input order
            ii = rng.integers(0, n - 1, 400); jj = rng.integers(1, n, 400)
            pl = sorted({(int(min(a, b)) << 32) | int(max(a, b)) for a, b in zip(ii, jj) if a != b})
        pairs = np.array(pl, np.uint64)
        bounds = [shard_range(n, world, r)[0] for r in range(world)] + [n]
        plans = [fetch_plan(pairs, world, r, bounds) for r in range(world)]
        covered = []
        for d in range(world):
            need, _send, recv_counts = plans[d]
            mine = pair_slice_of_rank(pairs, world, d)
            covered.append(mine)
            assert np.array_equal(need, genomes_of_pairs(mine))
            got = np.concatenate([plans[r][1][d].astype(np.int64) + bounds[r] for r in range(world)]) if world else []
            assert np.array_equal(got, need.astype(np.int64))                       # rank-major == ascending global order
            assert [len(plans[r][1][d]) for r in range(world)] == recv_counts.tolist()
            rp = remap_pairs(mine, need)
            back = (need[(rp >> np.uint64(32)).astype(np.int64)].astype(np.uint64) << np.uint64(32)) | need[(rp & np.uint64(0xFFFFFFFF)).astype(np.int64)].astype(np.uint64)
            assert np.array_equal(back, mine)
        assert np.array_equal(np.sort(np.concatenate(covered)), pairs)              # a partition of the pair list
        sizes = [len(c) for c in covered]
        assert all(np.all(np.diff(c.astype(np.int64)) > 0) for c in covered if len(c) > 1)
        # load balance: whole components stay together, items are at most half a fair share: nobody above 1.5x the mean (+ one item)
        assert max(sizes) <= 1.5 * len(pairs) / world + max(1, -(-len(pairs) // (2 * world)))
    # no pairs at all
    need, send, rc = fetch_plan(np.zeros(0, np.uint64), 2, 1, [0, 5, 10])
    assert len(need) == 0 and all(len(x) == 0 for x in send) and rc.tolist() == [0, 0]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from skani_b200.multi_gpu import gather_variable
    local = torch.arange(5 + 3 * rank, dtype=torch.int64) + 100 * rank
    parts = gather_variable(dist, local, world, "cpu")
    ok = all(torch.equal(parts[r], torch.arange(5 + 3 * r, dtype=torch.int64) + 100 * r) for r in range(world))
    b = torch.full((7 - 7 * rank,), rank + 1, dtype=torch.uint8)   # rank 1 contributes an empty blob
    pb = gather_variable(dist, b, world, "cpu")
    ok = ok and pb[0].numel() == 7 and pb[1].numel() == 0 and int(pb[0].sum()) == 7
    # variable all-to-all (the sketch fetch): rank r sends 3r+d bytes of value 10r+d to rank d, including nothing at all
    from skani_b200.multi_gpu import alltoall_variable
    parts = [torch.full((3 * rank + d,), 10 * rank + d, dtype=torch.uint8) for d in range(world)]
    got = alltoall_variable(dist, parts, world, "cpu")
    ok = ok and all(got[r].tolist() == [10 * r + rank] * (3 * r + rank) for r in range(world))
    buf = torch.cat(parts)
    views, o = [], 0
    for p_ in parts:
        views.append(buf[o:o + p_.numel()]); o += p_.numel()
    got2 = alltoall_variable(dist, views, world, "cpu", src=buf)
    ok = ok and all(torch.equal(a, b) for a, b in zip(got, got2))
    # partial pair lists of the sharded screen -> the same sorted list everywhere (incl. an empty contribution, values >= 2^63)
    from skani_b200.multi_gpu import allgather_sorted_u64
    mine_u = np.array([], np.uint64) if rank == 1 else np.array([(5 << 32) | 9, (1 << 63) | 7, 3], np.uint64)
    allp = allgather_sorted_u64(dist, torch, mine_u, world, "cpu")
    ok = ok and allp.dtype == np.uint64 and allp.tolist() == [3, (5 << 32) | 9, (1 << 63) | 7]
    allp = allgather_sorted_u64(dist, torch, np.array([10 * rank + 2, 10 * rank + 1], np.uint64), world, "cpu")
    ok = ok and allp.tolist() == [1, 2, 11, 12]
    ok = ok and len(allgather_sorted_u64(dist, torch, np.zeros(0, np.uint64), world, "cpu")) == 0
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_variable_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_cross_block_pairs_and_bench_helpers():
    """cross_block_pairs keeps exactly the pairs whose genomes live in different blocks (the pairs inside a block are chained
    by the block's own pipelined triangle), for the block layouts sk_triangle_multi and bench.py use; bench.py's result
    checksum is order independent and sensitive to every compared field; the permuted generator is a permutation."""
    from skani_b200.multi_gpu import cross_block_pairs, shard_range
    rng = np.random.default_rng(3)
    for n, world in ((40, 2), (1000, 8), (17, 3)):
        ii = rng.integers(0, n - 1, 500); jj = rng.integers(1, n, 500)
        pl = np.array(sorted({(int(min(a, b)) << 32) | int(max(a, b)) for a, b in zip(ii, jj) if a != b}), np.uint64)
        bounds = [shard_range(n, world, r)[0] for r in range(world)] + [n]
        blk = lambda g: max(r for r in range(world) if bounds[r] <= g)
        want = [int(x) for x in pl if blk(int(x) >> 32) != blk(int(x) & 0xFFFFFFFF)]
        assert cross_block_pairs(pl, bounds).tolist() == want
    assert len(cross_block_pairs(np.zeros(0, np.uint64), [0, 5, 10])) == 0
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from skani_b200.host import RESULT_DTYPE
    res = np.zeros(50, RESULT_DTYPE)
    res["ref_id"] = rng.integers(0, 1000, 50); res["query_id"] = rng.integers(0, 1000, 50)
    res["ani"] = rng.random(50).astype(np.float32); res["af_ref"] = rng.random(50).astype(np.float32); res["af_query"] = rng.random(50).astype(np.float32)
    c0 = bench.result_checksum(res)
    assert c0 == bench.result_checksum(res[rng.permutation(50)]) and 0 < c0 < 2 ** 64
    parts = [res[:20], res[20:]]
    assert (bench.result_checksum(parts[0]) + bench.result_checksum(parts[1])) % 2 ** 64 == c0      # what the all-reduce adds up
    for f in ("ref_id", "query_id", "ani", "af_ref", "af_query"):
        r2 = res.copy()
        r2[f][7] = r2[f][7] + (1 if f.endswith("_id") else np.float32(1e-6))
        assert bench.result_checksum(r2) != c0, f
    assert bench.result_checksum(res[:0]) == 0
    assert bench.expected_pairs(10000, 20) == 95000 and bench.expected_pairs(45, 20) == 2 * 190 + 10
    from bench_support import synth
    ids = synth.shuffled_ids(1000, 12345)
    assert sorted(ids.tolist()) == list(range(1000)) and ids.tolist() != list(range(1000))
    assert np.array_equal(ids, synth.shuffled_ids(1000, 12345))


def test_partition_pairs_keeps_clusters_together():
    """Shuffled genome order (relatedness unrelated to the index): the component-wise partition makes a rank fetch about one
    cluster's genomes per 190 pairs, where contiguous slices of the sorted list touch several times more genomes; a single giant
    component degrades to slices (no rank above its fair share + one item)."""
    from skani_b200.multi_gpu import partition_pairs, genomes_of_pairs, shard_range
    rng = np.random.default_rng(11)
    n, G, world = 2000, 20, 8
    perm = rng.permutation(n)                                            # genome g sits at index perm[g]
    pl = sorted({(int(min(perm[a], perm[b])) << 32) | int(max(perm[a], perm[b]))
                 for c in range(n // G) for a in range(c * G, (c + 1) * G) for b in range(a + 1, (c + 1) * G)})
    pairs = np.array(pl, np.uint64)
    parts = partition_pairs(pairs, world)
    assert np.array_equal(np.sort(np.concatenate(parts)), pairs)
    need_comp = [len(genomes_of_pairs(p)) for p in parts]
    need_slice = [len(genomes_of_pairs(pairs[slice(*shard_range(len(pairs), world, r))])) for r in range(world)]
    assert max(need_comp) <= 1.2 * n / world and sum(need_slice) > 3 * sum(need_comp), (need_comp, need_slice)
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 190                                # whole clusters: at most one cluster of difference
    # one giant component (a dense set): cut into runs, balanced
    dense = np.array([(i << 32) | j for i in range(60) for j in range(i + 1, 60)], np.uint64)
    dp = partition_pairs(dense, 4)
    assert np.array_equal(np.sort(np.concatenate(dp)), dense)
    assert max(len(x) for x in dp) <= len(dense) / 4 + len(dense) / 8 + 1
    assert [len(x) for x in partition_pairs(dense[:0], 3)] == [0, 0, 0] and len(partition_pairs(dense, 1)[0]) == len(dense)
