"""Multi-GPU triangle (markers all-gather + screen everywhere, contiguous pair slices, variable all-to-all fetch of the
sketches a rank chains; n=18, G=6 puts a cluster across the block boundary) == single-process oracle."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import oracle_py as O
from bench_support import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_triangle_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n, L, G = 18, 300_000, 6
    with tempfile.TemporaryDirectory() as d:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py"), d, str(n), str(L), str(G)]
        subprocess.check_call(cmd, cwd=ROOT, timeout=600)
        bases, off, goc = synth.generate(0, n, L, G=G)
        osk = [O.sketch_from_contigs("g%06d" % g, [bases[int(off[i]):int(off[i + 1])] for i in np.nonzero(goc == g)[0]]) for g in range(n)]
        ores, _ = O.triangle(osk, O.cmd())
        exp = {(r.ref_id, r.query_id): r for r in ores}
        for use_host in (0, 1):
            rows = np.concatenate([np.load(os.path.join(d, "rank%d_%d.npy" % (r, use_host))) for r in range(2)])
            got = {(int(x[0]), int(x[1])): x for x in rows}
            assert set(got) == set(exp) and len(rows) == len(exp)
            # both ranks chain: their own block's pairs plus a share of the cross-block component (items of <= 3 of its 9 pairs)
            sizes = [len(np.load(os.path.join(d, "rank%d_%d.npy" % (r, use_host)))) for r in range(2)]
            assert min(sizes) > 0 and abs(sizes[0] - sizes[1]) <= 3
            for k, e in exp.items():
                assert abs(got[k][2] - e.ani) <= 1e-4 and abs(got[k][3] - e.af_ref) <= 1e-4 and abs(got[k][4] - e.af_query) <= 1e-4


def test_one_process_multi_device_triangle_matches_single():
    """sk_triangle_multi with one context per PHYSICAL GPU (peer copies over NVLink): same result set as one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import skani_b200 as sk
    n, L, G = 18, 300_000, 6
    for ids in (np.arange(n, dtype=np.uint64), synth.shuffled_ids(n, 11)):
        bases, off, goc = synth.generate_ids(ids, L, G=G)
        ctxs = [sk.Context(d) for d in range(2)]
        try:
            single, _ = sk.triangle(ctxs[0], bases, off, goc, n, as_array=True)
            multi, _ = sk.triangle_multi(ctxs, bases, off, goc, n)
            a = np.sort(single, order=["ref_id", "query_id"]); b = np.sort(multi, order=["ref_id", "query_id"])
            assert len(a) == len(b) == n // G * (G * (G - 1) // 2) and a.tobytes() == b.tobytes()
        finally:
            for c in ctxs:
                c.close()
