"""GPU parity of the packed-input paths: sk_sketch_batch with every share of host-side 2-bit packing (SK_HOST_PACK = 0, a
fraction, 1, adaptive), many small sub-batches (the pack / upload / seed pipeline with its double-buffered slots), and
sk_sketch_batch_2bit with and without an N mask -- all must give the oracle's sketches bit for bit on the edge-case set
(N runs at quarter-lane boundaries, lowercase, IUPAC, lengths 500..131k)."""
import numpy as np
import pytest

import oracle_py as O
from bench_support import synth
from test_gpu_seeding import compare, parity_set

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import skani_b200 as sk
    c = sk.Context(0)
    yield c
    c.close()


def layout(genomes):
    arrs, goc = [], []
    for g, ctgs in enumerate(genomes):
        for c in ctgs:
            arrs.append(c); goc.append(g)
    off = np.zeros(len(arrs) + 1, np.uint64)
    off[1:] = np.cumsum([len(a) for a in arrs])
    return np.concatenate(arrs), off, np.asarray(goc, np.uint32)


@pytest.mark.parametrize("share", ["0", "0.35", "1", None])
@pytest.mark.parametrize("subbatch", [None, "70000"])
def test_hybrid_host_pack_parity(ctx, monkeypatch, share, subbatch):
    import skani_b200 as sk
    rng = np.random.default_rng(77)
    contigs = parity_set(rng, 60)
    genomes = [contigs[0:7], contigs[7:8], contigs[8:25], contigs[25:40], contigs[40:41], contigs[41:60]]
    bases, off, goc = layout(genomes)
    if share is None:
        monkeypatch.delenv("SK_HOST_PACK", raising=False)
    else:
        monkeypatch.setenv("SK_HOST_PACK", share)
    if subbatch:
        monkeypatch.setenv("SK_SUBBATCH_BYTES", subbatch)
    for pinned in (False, True):
        if pinned:
            import torch
            t = torch.empty(len(bases), dtype=torch.uint8, pin_memory=True)
            t.numpy()[:] = bases
            src = t.numpy()
        else:
            src = bases
        gs = sk.sketch_contigs(ctx, src, off, goc, len(genomes))
        if share in ("0", "1"):
            assert abs(ctx.last_pack_share - float(share)) < 1e-9
        for g, ctgs in enumerate(genomes):
            compare(gs, g, O.sketch_from_contigs("g%d" % g, ctgs))
        gs.free()


@pytest.mark.parametrize("c,k,mc", [(125, 15, 1000), (30, 15, 200), (10, 13, 40)])
def test_2bit_entry_parity(ctx, monkeypatch, c, k, mc):
    import skani_b200 as sk
    rng = np.random.default_rng(5 + c)
    contigs = parity_set(rng, 30)
    genomes = [contigs[0:5], contigs[5:6], contigs[6:30]]
    bases, off, goc = layout(genomes)
    sp = sk.sketch_params(c, k, mc)
    units, nmask, lens = sk.pack_contigs(ctx.L, bases, off)
    for sub in (None, "50000"):
        if sub:
            monkeypatch.setenv("SK_SUBBATCH_BYTES", sub)
        gs = sk.sketch_contigs_2bit(ctx, units, nmask, lens, goc, len(genomes), sp)
        for g, ctgs in enumerate(genomes):
            compare(gs, g, O.sketch_from_contigs("g%d" % g, ctgs, c=c, k=k, marker_c=mc))
        gs.free()
    monkeypatch.delenv("SK_SUBBATCH_BYTES", raising=False)
    # without a mask the 'N's read as 'A' (code 0): equals the oracle on the N-free rewrite of the same contigs
    clean = [np.where(x == ord("N"), ord("A"), x).astype(np.uint8) for x in contigs]
    cg = [clean[0:5], clean[5:6], clean[6:30]]
    gs = sk.sketch_contigs_2bit(ctx, units, None, lens, goc, len(genomes), sp)
    for g, ctgs in enumerate(cg):
        compare(gs, g, O.sketch_from_contigs("g%d" % g, ctgs, c=c, k=k, marker_c=mc))


def test_2bit_synthetic_and_triangle_hybrid(ctx, monkeypatch):
    """whole triangle through sk_triangle with host packing forced on / off: identical result bytes"""
    import skani_b200 as sk
    n, L, G = 12, 300_000, 4
    bases, off, goc = synth.generate(0, n, L, G=G)
    outs = []
    for share in ("0", "1", "0.5"):
        monkeypatch.setenv("SK_HOST_PACK", share)
        r, _ = sk.triangle(ctx, bases, off, goc, n, as_array=True)
        outs.append(np.sort(r, order=["ref_id", "query_id"]).tobytes())
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) > 0
    units, nmask, lens = sk.pack_contigs(ctx.L, bases, off)
    gs = sk.sketch_contigs_2bit(ctx, units, None, lens, goc, n)
    ga = sk.sketch_contigs(ctx, bases, off, goc, n)
    for g in range(n):
        ea, eb = gs.export(g), ga.export(g)
        for key in ea:
            assert np.array_equal(ea[key], eb[key])


def test_host_packer_implementation_and_bytes():
    """The GPU boxes' CPUs have AVX-512 VBMI: the table-driven packer must be the one in use there (its self-check passed), and
    sk_pack_contig must agree with the byte table (oracle ascii semantics, every byte value at every position of a unit)
    whatever implementation runs.  (Marked gpu only because that is where the AVX-512 paths exist; no kernel is launched.)"""
    import skani_b200 as sk
    from skani_b200 import _lib
    L = _lib.load()
    impl = L.sk_pack_impl().decode()
    flags = open("/proc/cpuinfo").read()
    if "avx512vbmi" in flags and "avx512bw" in flags:
        assert impl == "avx512vbmi", impl
    rng = np.random.default_rng(5)
    parts = [np.arange(256, dtype=np.uint8).repeat(3), rng.integers(0, 256, 100_003).astype(np.uint8),
             np.frombuffer(bytes(rng.choice(list(b"ACGTNacgtnUuRYKM-*"), 70_001).tolist()), np.uint8)]
    for shift in (0, 1, 31, 33, 63, 64, 65):
        seq = np.concatenate([np.frombuffer(b"A" * shift, np.uint8)] + parts)
        units, nmask, clen = sk.pack_contigs(L, seq, np.array([0, len(seq)], np.uint64))
        code = np.zeros(256, np.uint8)
        for ch, v in ((b"C", 1), (b"G", 2), (b"T", 3), (b"U", 3), (b"c", 1), (b"g", 2), (b"t", 3), (b"u", 3)):
            code[ch[0]] = v
        code[:4] = np.arange(4)
        c = code[seq].astype(np.uint64)
        pad = (-len(seq)) % 32
        c = np.concatenate([c, np.zeros(pad, np.uint64)]).reshape(-1, 32)
        want = (c << (2 * np.arange(32, dtype=np.uint64))).sum(axis=1, dtype=np.uint64)      # disjoint bit fields: sum == or
        isn = np.concatenate([(seq == 78), np.zeros(pad, bool)]).reshape(-1, 32).astype(np.uint64)
        want_n = (isn << np.arange(32, dtype=np.uint64)).sum(axis=1, dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(units, want), shift
        assert np.array_equal(nmask, want_n), shift


@pytest.mark.parametrize("pipelined", [False, True])
def test_triangle_2bit_equals_ascii_triangle(ctx, monkeypatch, pipelined):
    """sk_triangle_2bit (genomes already packed on the host, with an N mask) gives the result bytes of sk_triangle on the ASCII
    form, through the one-shot path and through the upload || seed || chain pipeline with several sub-batches and waves."""
    import skani_b200 as sk
    n, L, G = 14, 250_000, 4
    bases, off, goc = synth.generate(0, n, L, G=G)
    bases = bases.copy()
    bases[1000:1040] = ord("N"); bases[3 * L + 77: 3 * L + 79] = ord("n")          # an N run and lower-case n (not an 'N' for the mask)
    r0, st0 = sk.triangle(ctx, bases, off, goc, n, as_array=True)
    units, nmask, lens = sk.pack_contigs(ctx.L, bases, off)
    if pipelined:
        monkeypatch.setenv("SK_FORCE_PIPELINE", "1")
        monkeypatch.setenv("SK_SUBBATCH_BYTES", "600000")
    r1, st1 = sk.triangle_2bit(ctx, units, nmask, lens, goc, n)
    r2, kept_set, _ = sk.triangle_2bit(ctx, units, nmask, lens, goc, n, keep_set=True)
    assert len(kept_set) == n
    kept_set.free()
    k0 = np.sort(r0, order=["ref_id", "query_id"]).tobytes()
    assert len(r0) > 0 and k0 == np.sort(r1, order=["ref_id", "query_id"]).tobytes() == np.sort(r2, order=["ref_id", "query_id"]).tobytes()
    assert st0.n_pairs_screened == st1.n_pairs_screened
