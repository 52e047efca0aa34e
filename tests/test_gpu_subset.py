"""GPU: sub-blobs of a sketch set (sk_sketch_set_pack_subset -> sk_sketch_set_unpack), the pieces of the multi-GPU
"fetch what you chain" exchange, on one GPU: a subset must carry exactly its genomes' sketches, a markers-only blob must
screen like the full set, and chaining pairs on a working set of fetched genomes must give the results of the full set."""
import ctypes as C

import numpy as np
import pytest

from bench_support import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import skani_b200 as sk
    c = sk.Context(0)
    yield c
    c.close()


def unpack(ctx, torch, sets_and_genomes, flags=0):
    """pack each (set, genomes) into its own device blob, then unpack all of them into one set"""
    import skani_b200 as sk
    blobs, metas = [], []
    for s, genomes in sets_and_genomes:
        nb, nw = s.subset_blob_size(genomes, flags)
        assert nb % 256 == 0
        t = torch.empty(nb, dtype=torch.uint8, device="cuda")
        metas.append(s.pack_subset(genomes, flags, t.data_ptr(), nw))
        blobs.append(t)
    n = len(blobs)
    bp = (C.c_void_p * n)(*[b.data_ptr() for b in blobs])
    mp_ = (C.c_void_p * n)(*[m.ctypes.data for m in metas])
    res = C.c_void_p()
    ctx.check(ctx.L.sk_sketch_set_unpack(ctx.h, n, bp, mp_, C.byref(res)))
    torch.cuda.synchronize()
    return sk.SketchSet(ctx, res)


def test_subset_roundtrip_markers_only_and_working_set(ctx):
    import torch
    import skani_b200 as sk
    from skani_b200.multi_gpu import genomes_of_pairs, remap_pairs, PACK_MARKERS_ONLY
    n, L, G = 16, 200_000, 4
    bases, off, goc = synth.generate(0, n, L, G=G)          # members m % 4 == 2 have 50 contigs
    sp, mp = sk.sketch_params(), sk.map_params()
    full = sk.sketch_contigs(ctx, bases, off, goc, n, sp)
    # (1) arbitrary order, runs and singletons, duplicates
    pick = np.array([5, 6, 7, 2, 15, 0, 1, 6], np.uint32)
    sub = unpack(ctx, torch, [(full, pick)])
    assert len(sub) == len(pick)
    for i, g in enumerate(pick):
        a, b = sub.export(i), full.export(int(g))
        for key in ("kmer", "pos", "cc", "markers", "contig_lengths"):
            assert np.array_equal(a[key], b[key]), (key, i, g)
        assert sub.info(i) == full.info(int(g))
    sub.free()
    # (2) all genomes (NULL list) in two parts == the set itself; empty subset is legal
    halves = unpack(ctx, torch, [(full, np.arange(0, 7, dtype=np.uint32)), (full, np.zeros(0, np.uint32)), (full, np.arange(7, n, dtype=np.uint32))])
    whole = unpack(ctx, torch, [(full, None)])
    pairs_full = sk.screen_triangle(ctx, full, mp)
    assert len(pairs_full) >= 17
    for s in (halves, whole):
        assert len(s) == n
        assert np.array_equal(sk.screen_triangle(ctx, s, mp), pairs_full)
    res_full = sk.chain_pairs(ctx, full, full, pairs_full, mp, as_array=True)
    res_halves = sk.chain_pairs(ctx, halves, halves, pairs_full, mp, as_array=True)
    assert res_full.tobytes() == res_halves.tobytes()
    halves.free(); whole.free()
    # (3) markers only: same screen, no seeds, chains to "no anchors"
    mk = unpack(ctx, torch, [(full, None)], PACK_MARKERS_ONLY)
    assert len(mk) == n and mk.info(3)["n_records"] == 0 and mk.info(3)["n_markers"] == full.info(3)["n_markers"]
    assert np.array_equal(sk.screen_triangle(ctx, mk, mp), pairs_full)
    r = sk.chain_pairs(ctx, mk, mk, pairs_full[:3], mp, as_array=True)
    assert np.all(np.isnan(r["ani"]))
    mk.free()
    # (4) a rank's view: chain a slice of the pair list on a working set holding only the genomes the slice touches
    mine = pairs_full[5:17]
    need = genomes_of_pairs(mine)
    assert 0 < len(need) < n
    work = unpack(ctx, torch, [(full, need[:3]), (full, need[3:])])
    work.set_name_ranks(need.astype(np.uint64))
    rw = sk.chain_pairs(ctx, work, work, remap_pairs(mine, need), mp, as_array=True)
    rw["ref_id"] = need[rw["ref_id"]]
    rw["query_id"] = need[rw["query_id"]]
    assert rw.tobytes() == res_full[5:17].tobytes()
    work.free(); full.free()


@pytest.mark.parametrize("flags", [0, 2])
def test_scattered_subset_uses_the_batched_copy(ctx, flags):
    """More than 8 separate runs of genomes (the cross-block fetch of a genome order unrelated to relatedness) go through one
    batched device memcpy instead of thousands of cudaMemcpyAsync calls: same sketches (and, with SK_PACK_TABLES, chaining on
    the unpacked set without rebuilding tables gives the full set's results)."""
    import torch
    import skani_b200 as sk
    n, L, G = 40, 120_000, 4
    bases, off, goc = synth.generate(0, n, L, G=G)
    mp = sk.map_params()
    full = sk.sketch_contigs(ctx, bases, off, goc, n)
    pick = np.array([0, 2, 3, 5, 8, 9, 11, 14, 16, 17, 20, 23, 25, 26, 27, 30, 33, 35, 38, 39], np.uint32)   # 14 runs
    sub = unpack(ctx, torch, [(full, pick[:11]), (full, pick[11:])], flags)
    assert len(sub) == len(pick)
    for i, g in enumerate(pick):
        a, b = sub.export(i), full.export(int(g))
        for key in ("kmer", "pos", "cc", "markers", "contig_lengths"):
            assert np.array_equal(a[key], b[key]), (key, i, g)
    pairs_full = sk.screen_triangle(ctx, full, mp)
    keep = pairs_full[np.isin(pairs_full >> np.uint64(32), pick) & np.isin(pairs_full & np.uint64(0xFFFFFFFF), pick)]
    assert len(keep) >= 5
    from skani_b200.multi_gpu import remap_pairs
    sub.set_name_ranks(pick.astype(np.uint64))
    rs = sk.chain_pairs(ctx, sub, sub, remap_pairs(keep, pick), mp, as_array=True)
    rf = sk.chain_pairs(ctx, full, full, keep, mp, as_array=True)
    rs["ref_id"] = pick[rs["ref_id"]]; rs["query_id"] = pick[rs["query_id"]]
    assert rs.tobytes() == rf.tobytes()
    sub.free(); full.free()
