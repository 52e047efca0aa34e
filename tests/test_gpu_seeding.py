"""GPU parity: seeding kernels (pack -> hashpass -> expand -> views) vs the CPU oracle, bit-exact, through the C ABI."""
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


def parity_set(rng, n_contigs=40):
    """contigs with N runs at quarter-lane boundaries, lowercase, IUPAC codes, lengths 499/500/501 and every
    residue of (len - 20) mod 4 (SURVEY.md section 8d 'parity set')."""
    out = []
    for t in range(n_contigs):
        n = [500, 501, 502, 503, 777, 1024, 4099, 20000, 65537, 131075][t % 10] + int(rng.integers(0, 4))
        s = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
        fl = t % 6
        if fl == 1:
            s[rng.integers(0, n, 5)] = ord("N")
        if fl == 2:
            q = (n - 20) // 4
            for l in range(4):
                for d in range(-2, 24):
                    p = l * q + d + int(rng.integers(0, 3))
                    if 0 <= p < n and rng.random() < 0.33:
                        s[p] = ord("N")
        if fl == 3:
            m = rng.random(n) < 0.15
            s[m] = s[m] + 32
        if fl == 4:
            s[rng.integers(0, n, 30)] = rng.choice(np.frombuffer(b"RYKMSWnuUBDHV-*", np.uint8), 30)
        if fl == 5:
            a = int(rng.integers(0, n - 1)); b = min(n, a + 1 + int(rng.integers(0, 300)))
            s[a:b] = ord("N")
        out.append(s.astype(np.uint8))
    return out


def compare(gset, g, osk):
    e = gset.export(g)
    o = osk.export()
    assert gset.info(g)["n_kmers"] == osk.n_kmers
    for key in ("kmer", "pos", "cc", "markers", "contig_lengths"):
        assert np.array_equal(e[key], o[key]), key
    assert gset.info(g)["total_len"] == osk.total_len


@pytest.mark.parametrize("c,k,mc", [(125, 15, 1000), (30, 15, 200), (200, 15, 1000), (10, 13, 40), (125, 16, 1000)])
def test_seed_parity_edge_cases(ctx, c, k, mc):
    import skani_b200 as sk
    rng = np.random.default_rng(1234 + c)
    contigs = parity_set(rng)
    genomes = [contigs[0:7], contigs[7:8], contigs[8:25], contigs[25:40]]
    sp = sk.sketch_params(c, k, mc)
    gs = sk.sketch_sequences(ctx, genomes, sp)
    assert len(gs) == 4
    for g, ctgs in enumerate(genomes):
        osk = O.sketch_from_contigs("g%d" % g, ctgs, c=c, k=k, marker_c=mc)
        compare(gs, g, osk)
    # individual-contig mode (-i): one sketch per kept record
    gi = sk.sketch_sequences(ctx, [contigs[:9]], sp, individual_contig=True)
    assert len(gi) == 9
    for g in range(9):
        compare(gi, g, O.sketch_from_contigs("x", [contigs[g]], c=c, k=k, marker_c=mc))


def test_seed_parity_synthetic_genomes(ctx):
    import skani_b200 as sk
    L = 600_000
    bases, off, goc = synth.generate(0, 8, L, G=4)
    gs = sk.sketch_contigs(ctx, bases, off, goc, 8)
    for g in range(8):
        idx = np.nonzero(goc == g)[0]
        ctgs = [bases[int(off[i]):int(off[i + 1])] for i in idx]
        compare(gs, g, O.sketch_from_contigs("g", ctgs))
    # append() concatenates sets without changing per-genome content
    a = sk.sketch_contigs(ctx, bases[:4 * L], off[:np.searchsorted(goc, 4) + 1], goc[goc < 4], 4)
    n4 = int(np.searchsorted(goc, 4))
    b = sk.sketch_contigs(ctx, bases[4 * L:], off[n4:] - off[n4], goc[n4:] - 4, 4)
    a.append(b)
    assert len(a) == 8
    for g in range(8):
        ea, eg = a.export(g), gs.export(g)
        for key in ea:
            assert np.array_equal(ea[key], eg[key])


def test_empty_and_tiny_inputs(ctx):
    import skani_b200 as sk
    # a genome whose only contig is < 42 bases seeds nothing; zero contigs -> empty sketches
    gs = sk.sketch_contigs(ctx, np.frombuffer(b"ACGT" * 10, np.uint8), [0, 40], [0], 2)
    assert len(gs) == 2 and gs.info(0)["n_records"] == 0 and gs.info(1)["n_contigs"] == 0
    with pytest.raises(Exception):
        sk.sketch_contigs(ctx, np.zeros(600, np.uint8), [0, 600], [0], 1, sk.sketch_params(2000, 15, 1000))  # c > marker_c


@pytest.mark.parametrize("variant", ["1", "2"])
def test_hashpass_arithmetic_variants_bit_exact(ctx, monkeypatch, variant):
    """The A/B variants of hashpass_kernel (xor-shift right shifts issued as mul.hi on the FMA pipe, SK_HASHPASS_VARIANT)
    select exactly the same windows as the default arithmetic."""
    import skani_b200 as sk
    monkeypatch.setenv("SK_HASHPASS_VARIANT", variant)
    rng = np.random.default_rng(99)
    contigs = parity_set(rng, 30)
    for c, k, mc in ((125, 15, 1000), (10, 13, 40)):
        gs = sk.sketch_sequences(ctx, [contigs[:12], contigs[12:30]], sk.sketch_params(c, k, mc))
        compare(gs, 0, O.sketch_from_contigs("a", contigs[:12], c=c, k=k, marker_c=mc))
        compare(gs, 1, O.sketch_from_contigs("b", contigs[12:30], c=c, k=k, marker_c=mc))


def test_scalar_seeder_semantics_bit_exact(ctx):
    """SK_SEED_SCALAR: the device reproduces seeding::fmh_seeds (src/seeding.rs:225-323, the path of hosts without AVX2):
    single lane, no dropped tail windows, 'N' AND 'n' suppress the next k windows -- against the oracle's scalar seeder on the
    edge-case set (which contains lowercase n, N runs, IUPAC codes, lengths 500..131k)."""
    import skani_b200 as sk
    rng = np.random.default_rng(4321)
    contigs = parity_set(rng, 40)
    contigs[3][100:110] = ord("n"); contigs[5][25:40] = ord("n"); contigs[7][-3:] = ord("n")
    ctx.set_seeding_semantics(scalar=True)
    try:
        for c, k, mc in ((125, 15, 1000), (30, 15, 200), (10, 13, 40), (125, 16, 1000)):
            genomes = [contigs[0:9], contigs[9:10], contigs[10:40]]
            gs = sk.sketch_sequences(ctx, genomes, sk.sketch_params(c, k, mc))
            for g, ctgs in enumerate(genomes):
                compare(gs, g, O.sketch_from_contigs("g%d" % g, ctgs, c=c, k=k, marker_c=mc, avx2sem=False))
        # and the two semantics really differ on this set
        a = sk.sketch_sequences(ctx, [contigs[0:9]], sk.sketch_params()).export(0)
        ctx.set_seeding_semantics(scalar=False)
        b = sk.sketch_sequences(ctx, [contigs[0:9]], sk.sketch_params()).export(0)
        assert len(a["kmer"]) != len(b["kmer"]) or not np.array_equal(a["pos"], b["pos"])
    finally:
        ctx.set_seeding_semantics(scalar=False)
