#!/usr/bin/env python3
"""bench.py -- genome-pairs/sec of the `skani triangle` hot path on synthetic 5 Mbp bacterial genomes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--genomes G] ...
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

A "step" is one full pass of the hot path (FracMinHash seeding -> marker screen -> chaining/ANI) over the whole
synthetic genome set.  `value` = genome pairs / second with the ASCII genomes already resident in HBM; `e2e` is the same
metric through the C ABI with HOST (pinned) buffers: H2D of every base and D2H of every result inside the timed region.
`--impl reference` times the reference's CPU algorithm (the C++ restatement in oracle/: the Rust crate cannot be built
in this image) on all host cores over a bounded sample of the same workload, scaled as stated in `cpu_baseline.sample`.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "genome-pairs/sec, skani triangle 10k x 5 Mbp synthetic"
UNIT = "genome-pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genomes", type=int, default=10000)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--cluster", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=400, help="genomes in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle = C++ restatement of the reference algorithm; "port")
# ------------------------------------------------------------------------------------------------------------
def host_threads():
    """Threads for the CPU arm: all host CPUs this process may actually use.  The GPU boxes expose 128 logical CPUs but the
    container's cgroup grants a CPU-time quota (cpu.max, 16 CPUs on the round-1 boxes); running 128 threads under that quota
    is SLOWER than 32 (measured: seeding speed-up 15.9x at 16 threads, 18.7x at 32, 9.2x at 128), so use 2 threads per
    granted CPU, capped by the visible CPUs."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            quota = float(q) / float(p)
    except Exception:
        pass
    if quota:
        n = max(1, min(n, int(round(2 * quota))))
    return n, quota



def cpu_triangle_sample(n_sample, L, G, threads):
    """Times seeding + screen + chain of the oracle on n_sample synthetic genomes. Returns seconds and details."""
    import oracle_py as O
    from bench_support import synth
    bases, off, goc = synth.generate(0, n_sample, L, G=G)
    t0 = time.perf_counter()
    # seeding: OpenMP threads over genomes inside the oracle, the reference's own parallel structure (src/file_io.rs:149)
    sk = O.sketch_many(bases, off, goc, n_sample, threads=threads)
    t1 = time.perf_counter()
    res, info = O.triangle(sk, O.cmd(), threads=threads)
    t2 = time.perf_counter()
    return dict(t_seed=t1 - t0, t_pairs=t2 - t1, t_total=t2 - t0, n_kept=len(res), n_chained=info["n_chained"],
                t_screen=info["t_screen"], t_chain=info["t_chain"])


def cpu_single_thread(L, G):
    """Per-thread rates of the oracle (1 thread): ms per genome seeded, ms per chained pair."""
    import oracle_py as O
    from bench_support import synth
    n = min(G, 8)
    bases, off, goc = synth.generate(0, n, L, G=G, threads=1)
    t0 = time.perf_counter()
    sk = [O.sketch_from_contigs("g%06d" % g, [bases[int(off[i]):int(off[i + 1])] for i in np.nonzero(goc == g)[0]]) for g in range(n)]
    t1 = time.perf_counter()
    res, info = O.triangle(sk, O.cmd(), threads=1)
    t2 = time.perf_counter()
    return {"seed_ms_per_genome": (t1 - t0) * 1e3 / n, "chain_ms_per_pair": (t2 - t1) * 1e3 / max(info["n_chained"], 1)}


def cpu_baseline(args, threads):
    S = min(args.cpu_sample, args.genomes)
    S = max(args.cluster, (S // args.cluster) * args.cluster)
    d = cpu_triangle_sample(S, args.genome_len, args.cluster, threads)
    N = args.genomes
    t_full = d["t_total"] * (N / S)   # every stage is linear in N for the clustered set (fixed cluster size)
    value = (N * (N - 1) / 2) / t_full
    one = cpu_single_thread(args.genome_len, args.cluster)
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "single_thread": one,
            "cgroup_cpu_quota": host_threads()[1],
            "sample": "oracle (C++ restatement of skani 0.3.0; Rust reference not buildable here) triangle on %d of the %d genomes "
                      "(%d clusters): %.2f s (seeding %.2f s; serial marker index + screen %.2f s; chain %.2f s for %d pairs) scaled "
                      "x%.1f to the full set (all stages linear in N at fixed cluster size)" % (S, N, S // args.cluster, d["t_total"],
                                                                                              d["t_seed"], d["t_screen"], d["t_chain"],
                                                                                              d["n_chained"], N / S)}, d


def workload_name(N, L, G):
    """config.workload, identical for both arms (same synthetic set, same sketch parameters)"""
    return ("triangle %d x %d bp synthetic clustered (G=%d, subst 0.1-5%%, inversions, 50-contig members), c=125 k=15 m=1000; "
            "inputs %.1f GB >> L2 (no flush needed)" % (N, L, G, N * L / 1e9))


# ------------------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads, _quota = host_threads()
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        cb, d = cpu_baseline(args, threads)
        if i >= args.warmup:
            vals.append((cb["value"], d["t_total"]))
        last = cb
        if i == 0 and d["t_total"] > 30:   # keep the whole run within a few minutes
            args.warmup, args.steps = 0, min(args.steps, 2)
            vals.append((cb["value"], d["t_total"]))
            break
    v = float(np.mean([x[0] for x in vals]))
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
            "ms_per_step": float(np.mean([x[1] for x in vals])) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args.genomes, args.genome_len, args.cluster), "genomes": args.genomes,
                       "genome_len": args.genome_len, "pairs": args.genomes * (args.genomes - 1) // 2},
            "cpu_baseline": last, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------
# b200 arm
# ------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import skani_b200 as sk
    from skani_b200 import _lib
    from bench_support import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)

    N, L, G = args.genomes, args.genome_len, args.cluster
    # genome shard of this rank (contiguous block; global genome order = rank-major)
    g0, g1 = (N * rank) // world, (N * (rank + 1)) // world
    nloc = g1 - g0
    pinned = torch.empty(nloc * L, dtype=torch.uint8, pin_memory=True)
    host = pinned.numpy()
    t_gen0 = time.perf_counter()
    synth.generate(g0, g1, L, G=G, out=host)
    off, goc = synth.layout(g0, g1, L, G)
    t_gen = time.perf_counter() - t_gen0
    ctx = sk.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    sp, mp = sk.sketch_params(), sk.map_params()
    dev_bases = pinned.to("cuda", non_blocking=False)
    total_pairs = N * (N - 1) // 2
    result_bytes = [0]
    tri = None
    if world > 1:
        from skani_b200.multi_gpu import DistTriangle
        tri = DistTriangle(ctx, world, rank, sp, mp)

    def step(e2e):
        if world == 1:
            if e2e:
                res, st = sk.triangle(ctx, host, off, goc, nloc, sp, mp, as_array=True)
                result_bytes[0] = len(res) * C.sizeof(_lib.AniResult)
                return len(res), st
            gs = sk.sketch_contigs(ctx, None, off, goc, nloc, sp, device_ptr=dev_bases.data_ptr())
            pairs = sk.screen_triangle(ctx, gs, mp)
            res = sk.chain_pairs(ctx, gs, gs, pairs, mp, as_array=True)
            kept = int((res["ani"] > 0.1).sum())
            gs.free()
            return kept, None
        kept = tri.step(host if e2e else None, dev_bases.data_ptr() if dev_bases is not None else 0, off, goc, nloc, g0, N)
        if e2e:
            result_bytes[0] = kept * C.sizeof(_lib.AniResult)
        return kept, None

    def timed(e2e, n_steps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        t0 = time.perf_counter()
        ev0.record(stream)
        kept = 0
        for _ in range(n_steps):
            kept, _st = step(e2e)
        ev1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        ms = max(dev_ms, wall * 1e3)  # the step blocks on its own D2H copies; wall >= device span. report the larger
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n_steps, kept, ctx.launches - l0

    # value leg first (needs the device-resident copy of the bases), then the roofline probe, then drop the 50 GB device
    # copy before the end-to-end leg so that both legs run with comfortable HBM headroom
    for _ in range(max(args.warmup, 0)):
        step(False)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_val, _, launches = timed(False, args.steps)
    roof = None
    if rank == 0:
        roof = measure_hashpass(ctx, stream, dev_bases, off, goc, nloc, sp, L)
    dev_bases = None
    torch.cuda.empty_cache()
    for _ in range(max(args.warmup, 0)):
        step(True)
    ms_e2e, kept, _ = timed(True, args.steps)
    clk = clocks.stop() if rank == 0 else None

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu, _d = cpu_baseline(args, host_threads()[0])
    if rank == 0:
        line = {"metric": METRIC, "value": total_pairs / (ms_val * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_val, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload_name(N, L, G), "genomes": N, "genome_len": L, "pairs": total_pairs, "kept_pairs_rank0": kept,
                           "kept_pairs_expected_all_ranks": N // G * (G * (G - 1) // 2),
                           "host_gen_s": round(t_gen, 2)},
                "e2e": {"value": total_pairs / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(N * L), "d2h_bytes_per_step": int(result_bytes[0])},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def measure_hashpass(ctx, stream, dev_bases, off, goc, nloc, sp, L):
    """Roofline of the dominant kernel (hashpass_kernel), timed live with CUDA events on the launch stream
    (sk_ctx_set_timing brackets every launch).  Algorithmic bytes: SURVEY.md section 8(d) seeding figure
    L/4 + L/8 + 12 L/c + 8 L/m per genome (2.395 MB at L = 5 Mbp), times the genomes one launch processes."""
    import skani_b200 as sk
    n = min(nloc, 200)
    n_contigs = int(np.searchsorted(goc, n))
    o, g = off[:n_contigs + 1], goc[:n_contigs]
    sk.sketch_contigs(ctx, None, o, g, n, sp, device_ptr=dev_bases.data_ptr()).free()   # warm
    ctx.get_timing(reset=True)
    ctx.set_timing(True)
    reps = 3
    for _ in range(reps):
        sk.sketch_contigs(ctx, None, o, g, n, sp, device_ptr=dev_bases.data_ptr()).free()
    t = ctx.get_timing(reset=True)
    ctx.set_timing(False)
    ms, launches = t["hashpass_kernel"]
    bytes_per_genome = L / 4.0 + L / 8.0 + 12.0 * L / sp.c + 8.0 * L / sp.marker_c
    genomes_per_launch = n * reps / launches
    sec_per_launch = ms * 1e-3 / launches
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = genomes_per_launch * bytes_per_genome / sec_per_launch / 1e9
    stage_ms = sum(v[0] for v in t.values())
    # DRAM traffic per genome of this kernel from the committed ncu --set full capture (profiles/r01_ncu_full_top5.md:
    # dram__bytes_read 250.05 MB + dram__bytes_write 98.18 MB for a 100-genome launch), scaled to this launch size
    traffic = (250.05184e6 + 98.18496e6) / 100.0 * genomes_per_launch
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_ncu_full_top5.md, per genome x genomes_per_launch",
            "kernel": "hashpass_kernel", "launch_ms": sec_per_launch * 1e3, "genomes_per_launch": genomes_per_launch,
            "algorithmic_bytes_per_genome": bytes_per_genome,
            "gbases_per_s": genomes_per_launch * L / sec_per_launch / 1e9,
            "kernel_ms": {k: round(v[0] / reps, 3) for k, v in t.items()}, "timed_kernels_ms_per_rep": round(stage_ms / reps, 3),
            "note": "hashpass is integer-ALU bound (64-bit hash per base at 0.375 B/base), not HBM bound: see DESIGN.md",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}


if __name__ == "__main__":
    sys.exit(main())
