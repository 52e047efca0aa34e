"""GPU: sk_sketch_set_import_batch (host sketches -> device set; the path a skani database takes into HBM) round-trips
against sk_sketch_set_export, chains identically to the set the sketches came from, and reproduces the README pair
(golden G13) from the ORACLE's sketches of the two E. coli fixtures."""
import os

import numpy as np
import pytest

import oracle_py as O
from bench_support import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    import skani_b200 as sk
    c = sk.Context(0)
    yield c
    c.close()


def test_import_batch_roundtrip_and_chain(ctx):
    import skani_b200 as sk
    n, L, G = 8, 150_000, 4
    bases, off, goc = synth.generate(0, n, L, G=G)          # members m % 4 == 2 have 50 contigs
    sp, mp = sk.sketch_params(), sk.map_params()
    full = sk.sketch_contigs(ctx, bases, off, goc, n, sp)
    exports = [full.export(g) for g in range(n)]
    rng = np.random.default_rng(5)
    shuffled = []
    for e in exports:                                       # records arrive in arbitrary (hash map) order
        perm = rng.permutation(len(e["kmer"]))
        shuffled.append(dict(kmer=e["kmer"][perm], pos=e["pos"][perm], cc=e["cc"][perm], markers=rng.permutation(e["markers"]),
                             contig_lengths=e["contig_lengths"]))
    imp = sk.import_sketches(ctx, shuffled, sp)
    assert len(imp) == n
    for g in range(n):
        a = imp.export(g)
        for key in ("kmer", "pos", "cc", "markers", "contig_lengths"):
            assert np.array_equal(a[key], exports[g][key]), (key, g)
        assert imp.info(g) == full.info(g)
    pairs = sk.screen_triangle(ctx, full, mp)
    assert len(pairs) >= 8 and np.array_equal(sk.screen_triangle(ctx, imp, mp), pairs)
    r_full = sk.chain_pairs(ctx, full, full, pairs, mp, as_array=True)
    r_imp = sk.chain_pairs(ctx, imp, imp, pairs, mp, as_array=True)
    assert r_full.tobytes() == r_imp.tobytes()
    r_mixed = sk.chain_pairs(ctx, imp, full, pairs, mp, as_array=True)     # imported refs, sketched queries (search's shape)
    assert np.array_equal(r_mixed["ani"], r_full["ani"], equal_nan=True)
    # markers only (markers.bin): same screen, no seeds
    mk = sk.import_sketches(ctx, [dict(markers=e["markers"], total_len=L) for e in exports], sp, seeds=False)
    assert mk.info(2)["n_records"] == 0 and mk.info(2)["total_len"] == L
    assert np.array_equal(sk.screen_triangle(ctx, mk, mp), pairs)
    q = sk.screen_query_ref(ctx, mk, full, mp, mode=1)
    assert np.array_equal(q, sk.screen_query_ref(ctx, full, full, mp, mode=1))
    # a single sketch through the one-genome entry point, an empty sketch in a batch
    one = sk.import_sketches(ctx, [shuffled[3]], sp)
    assert np.array_equal(one.export(0)["kmer"], exports[3]["kmer"])
    z = dict(kmer=np.zeros(0, np.uint32), pos=np.zeros(0, np.uint32), cc=np.zeros(0, np.uint32), markers=np.zeros(0, np.uint64),
             contig_lengths=np.zeros(0, np.uint32))
    with_empty = sk.import_sketches(ctx, [shuffled[0], z, shuffled[1]], sp)
    assert with_empty.info(1)["n_records"] == 0 and np.array_equal(with_empty.export(2)["pos"], exports[1]["pos"])
    r = sk.chain_pairs(ctx, with_empty, with_empty, np.array([(0 << 32) | 2, (0 << 32) | 1], np.uint64), mp, as_array=True)
    assert abs(r["ani"][0] - r_full["ani"][list(pairs).index((0 << 32) | 1)]) == 0 and np.isnan(r["ani"][1])
    for s in (one, with_empty, mk, imp, full):
        s.free()


def test_import_oracle_sketches_reproduces_g13(ctx):
    """oracle sketches of EC590 / K12 -> device -> chain: README ANI 99.39, AF 91.89 / 92.46"""
    import skani_b200 as sk
    osk, _ = O.sketch_files([os.path.join(GOLD, "e.coli-EC590.fasta.gz"), os.path.join(GOLD, "e.coli-K12.fasta.gz")])
    ex = []
    for s in osk:
        e = s.export()
        e["total_len"] = s.total_len
        ex.append(e)
    imp = sk.import_sketches(ctx, ex)
    r = sk.chain_pairs(ctx, imp, imp, np.array([(1 << 32) | 0], np.uint64), sk.map_params(), as_array=True)   # ref K12, query EC590
    f2 = lambda x: "%.2f" % float(np.float32(x) * np.float32(100.0))
    assert (f2(r["ani"][0]), f2(r["af_ref"][0]), f2(r["af_query"][0])) == ("99.39", "91.89", "92.46")
    imp.free()
