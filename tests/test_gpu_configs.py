"""Scaled-down versions of BASELINE.json configs 3 and 5 on the GPU vs the oracle (parity at the configurations'
parameter settings: search-mode screening, c=30 / m=200 / rescue off on thousands of short contigs)."""
import os

import numpy as np
import pytest

import oracle_py as O
from bench_support import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import skani_b200 as sk
    c = sk.Context(0)
    yield c
    c.close()


def close(g, o, tol=1e-4):
    return all(abs(float(g[f]) - getattr(o, f)) <= tol for f in ("ani", "af_query", "af_ref"))


def test_config5_small_genomes_triangle(ctx):
    """`skani triangle -i --small-genomes` (= -c 30 -m 200 --faster-small, src/parse.rs:847-853): 2000 contigs of
    2-120 kb in clusters of 10."""
    import skani_b200 as sk
    rng = np.random.default_rng(5)
    contigs = []
    for cl in range(200):
        L = int(np.exp(rng.uniform(np.log(2000), np.log(120000))))
        anc = rng.choice(np.frombuffer(b"ACGT", np.uint8), L)
        for m in range(10):
            s = anc.copy()
            mut = rng.random(L) < rng.uniform(0.001, 0.04)
            s[mut] = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(mut.sum()))
            contigs.append(s)
    kw = dict(c=30, k=15, marker_c=200)
    gs = sk.sketch_sequences(ctx, [contigs], sk.sketch_params(**kw), individual_contig=True)
    assert len(gs) == 2000
    osk = [O.sketch_from_contigs("f", [c], **kw) for c in contigs]
    mp = sk.map_params(learned_ani=False, rescue_small=False)
    pairs = sk.screen_triangle(ctx, gs, mp)
    ro, cols = O.screen_triangle(osk, rescue_small=False)
    exp_pairs = sorted((i << 32) | int(j) for i in range(2000) for j in cols[int(ro[i]):int(ro[i + 1])])
    assert pairs.tolist() == exp_pairs and len(exp_pairs) > 5000
    res = sk.chain_pairs(ctx, gs, gs, pairs, mp, as_array=True)
    ores, _ = O.triangle(osk, O.cmd(learned_ani=False, rescue_small=False), threads=16)
    exp = {(r.ref_id, r.query_id): r for r in ores}
    got = {(int(r["ref_id"]), int(r["query_id"])): r for r in res if r["ani"] > 0.1}
    assert set(got) == set(exp)
    assert all(close(got[k], exp[k]) for k in exp)
    # the probe kernel stages small k-mer tables in shared memory with a TMA bulk copy (cp.async.bulk + mbarrier); the
    # plain global-memory probe (SK_PROBE_TMA=0, the A/B baseline) must give byte-identical results
    os.environ["SK_PROBE_TMA"] = "0"
    try:
        res0 = sk.chain_pairs(ctx, gs, gs, pairs, mp, as_array=True)
    finally:
        del os.environ["SK_PROBE_TMA"]
    assert res.tobytes() == res0.tobytes()


def test_config3_search_queries_vs_db(ctx):
    """`skani search`: query genomes vs a pre-sketched DB (here: 40 queries x 400 refs of 300 kb); both screening modes of
    src/search.rs:123-140 and the ani > 0.5 keep rule."""
    import skani_b200 as sk
    L, G = 300_000, 20
    rb, roff, rgoc = synth.generate(0, 400, L, G=G)
    refs = sk.sketch_contigs(ctx, rb, roff, rgoc, 400)
    oref = [O.sketch_from_contigs("r%06d" % g, [rb[int(roff[i]):int(roff[i + 1])] for i in np.nonzero(rgoc == g)[0]]) for g in range(400)]
    # queries: fresh members of random clusters = the ancestor (member 0) re-mutated
    rng = np.random.default_rng(3)
    qs_host, oq = [], []
    for q in range(40):
        cl = int(rng.integers(0, 400 // G))
        a = rb[cl * G * L:(cl * G + 1) * L].copy()
        mut = rng.random(L) < rng.uniform(0.002, 0.03)
        a[mut] = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(mut.sum()))
        qs_host.append([a]); oq.append(O.sketch_from_contigs("q%03d" % q, [a]))
    qs = sk.sketch_sequences(ctx, qs_host)
    # switch_qr's tie-break compares FILE NAMES (src/chain.rs:19-21): the oracle sketches are named q%03d / r%06d
    qs.set_name_ranks(np.arange(40)); refs.set_name_ranks(40 + np.arange(400))
    mp = sk.map_params(min_af=-1.0, rescue_small=False)   # search defaults (src/parse.rs:962-990)
    for mode, use_index in ((1, False), (3, True)):
        pairs = sk.host.screen_query_ref(ctx, refs, qs, mp, mode=mode)
        res = sk.chain_pairs(ctx, refs, qs, pairs, mp, as_array=True)
        got = {(int(r["ref_id"]), int(r["query_id"])): r for r in res if r["ani"] > 0.5}
        ores = O.search(oref, oq, O.cmd(min_af=-1.0, rescue_small=False), use_index=use_index, threads=16)
        exp = {(r.ref_id, r.query_id): r for r in ores}
        assert set(got) == set(exp) and len(exp) >= 40 * (G - 1)
        assert all(close(got[k], exp[k]) for k in exp)
