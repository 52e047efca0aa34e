"""Multi-GPU triangle on ONE device: sk_triangle_multi with several contexts bound to GPU 0 runs the complete multi-GPU
path -- genome blocks, per-block pipelined triangle, marker exchange, replicated screen, cross-block pair slices, fetch
of sub-blobs with their k-mer tables, working-set chaining, id mapping -- with device copies standing in for the NVLink
peer copies.  The result SET must equal the single-context triangle byte for byte and the oracle within 1e-4."""
import numpy as np
import pytest

import oracle_py as O
from bench_support import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def as_dict(res):
    return {(int(r["ref_id"]), int(r["query_id"])): r for r in res}


def oracle_triangle(bases, off, goc, n):
    osk = [O.sketch_from_contigs("g%06d" % g, [bases[int(off[i]):int(off[i + 1])] for i in np.nonzero(goc == g)[0]]) for g in range(n)]
    ores, _ = O.triangle(osk, O.cmd())
    return {(r.ref_id, r.query_id): r for r in ores}


@pytest.mark.parametrize("world,order", [(2, "contiguous"), (3, "contiguous"), (2, "shuffled"), (4, "shuffled")])
def test_multi_context_triangle_matches_single_and_oracle(world, order):
    import skani_b200 as sk
    n, L, G = 18, 300_000, 6          # blocks of 9 / 6 cut clusters of 6
    ids = np.arange(n, dtype=np.uint64) if order == "contiguous" else synth.shuffled_ids(n, 11)
    bases, off, goc = synth.generate_ids(ids, L, G=G)
    ctxs = [sk.Context(0) for _ in range(world)]
    try:
        single, _ = sk.triangle(ctxs[0], bases, off, goc, n, as_array=True)
        multi, st = sk.triangle_multi(ctxs, bases, off, goc, n)
        a = np.sort(single, order=["ref_id", "query_id"]); b = np.sort(multi, order=["ref_id", "query_id"])
        assert len(a) == len(b) == n // G * (G * (G - 1) // 2)
        assert a.tobytes() == b.tobytes()
        assert st.n_pairs_screened == len(b)
        exp = oracle_triangle(bases, off, goc, n)
        got = as_dict(b)
        assert set(got) == set(exp)
        for k, e in exp.items():
            for f in ("ani", "af_ref", "af_query"):
                assert abs(float(got[k][f]) - getattr(e, f)) <= TOL, (k, f)
    finally:
        for c in ctxs:
            c.close()


def test_multi_context_name_ranks_and_empty_blocks():
    """-i style name ranks (records of one file share a rank) reach the cross-block chains; more contexts than genomes and
    genomes without sequence are handled."""
    import skani_b200 as sk
    n, L, G = 8, 300_000, 4
    bases, off, goc = synth.generate(0, n, L, G=G)
    ranks = np.array([0, 0, 0, 0, 1, 1, 1, 1], np.uint64)
    ctxs = [sk.Context(0) for _ in range(3)]
    try:
        gs = sk.sketch_contigs(ctxs[0], bases, off, goc, n)
        gs.set_name_ranks(ranks)
        pairs = sk.screen_triangle(ctxs[0], gs)
        ref = sk.chain_pairs(ctxs[0], gs, gs, pairs, as_array=True)
        ref = np.sort(ref[ref["ani"] > 0.1], order=["ref_id", "query_id"])
        gs.free()
        multi, _ = sk.triangle_multi(ctxs, bases, off, goc, n, name_ranks=ranks)
        assert np.sort(multi, order=["ref_id", "query_id"]).tobytes() == ref.tobytes()
        # two trailing genomes without any contig + more contexts than genomes with sequence
        m2, _ = sk.triangle_multi(ctxs, bases[:2 * L], off[:np.searchsorted(goc, 2) + 1], goc[goc < 2], 4)
        s2, _ = sk.triangle(ctxs[0], bases[:2 * L], off[:np.searchsorted(goc, 2) + 1], goc[goc < 2], 4, as_array=True)
        assert np.sort(m2, order=["ref_id", "query_id"]).tobytes() == np.sort(s2, order=["ref_id", "query_id"]).tobytes()
    finally:
        for c in ctxs:
            c.close()


def test_subset_blobs_with_tables_roundtrip():
    """SK_PACK_TABLES sub-blobs rebuild a working set that chains exactly like the source set (tables copied, not rebuilt)."""
    import skani_b200 as sk
    import torch
    n, L, G = 8, 300_000, 4
    bases, off, goc = synth.generate(0, n, L, G=G)
    ctx = sk.Context(0)
    try:
        gs = sk.sketch_contigs(ctx, bases, off, goc, n)
        pick = np.array([1, 2, 3, 5, 6], np.uint32)
        nb, nw = gs.subset_blob_size(pick, 2)
        nb0, _ = gs.subset_blob_size(pick, 0)
        assert nb > nb0
        buf = torch.empty(nb, dtype=torch.uint8, device="cuda")
        meta = gs.pack_subset(pick, 2, buf.data_ptr(), nw)
        from skani_b200.multi_gpu import DistTriangle
        dt = DistTriangle.__new__(DistTriangle)
        dt.ctx = ctx
        work = dt._unpack([buf.data_ptr()], [meta])
        assert len(work) == len(pick)
        for (a, b) in [(0, 1), (1, 2), (3, 4), (0, 3)]:
            x = sk.chain_pairs(ctx, work, work, [(a << 32) | b], as_array=True)[0]
            y = sk.chain_pairs(ctx, gs, gs, [(int(pick[a]) << 32) | int(pick[b])], as_array=True)[0]
            for f in ("ani", "af_ref", "af_query", "std", "ci_lower", "ci_upper"):
                assert (np.isnan(x[f]) and np.isnan(y[f])) or x[f] == y[f], f
    finally:
        ctx.close()
