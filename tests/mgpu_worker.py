"""torchrun worker for tests/test_gpu_multi.py: multi-GPU triangle on a small synthetic set; every rank writes its
kept results to <out>/rank<k>.npy as rows (ref, query, ani, af_ref, af_query)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skani_b200 as sk  # noqa: E402
from skani_b200.multi_gpu import DistTriangle, shard_range  # noqa: E402
from bench_support import synth  # noqa: E402

out_dir, n, L, G = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
g0, g1 = shard_range(n, world, rank)
bases, off, goc = synth.generate(g0, g1, L, G=G)
ctx = sk.Context(local)
tri = DistTriangle(ctx, world, rank, sk.sketch_params(), sk.map_params())
for use_host in (True, False):
    dev = torch.from_numpy(bases).cuda()
    kept = tri.step(bases if use_host else None, dev.data_ptr(), off, goc, g1 - g0, g0, n)
    lr = tri.last_results
    rows = np.stack([lr["ref_id"], lr["query_id"], lr["ani"], lr["af_ref"], lr["af_query"]], axis=1).astype(np.float64).reshape(-1, 5)
    np.save(os.path.join(out_dir, "rank%d_%d.npy" % (rank, int(use_host))), rows)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
