"""BASELINE config 5: Chamfer MMD-CD / COV-CD / 1-NNA-CD over a generated and a reference set
sharded across ranks (one process per GPU; NCCL all-gather of the sets, local CD row blocks).

    torchrun --nproc-per-node N tools/eval_sets.py --total 8192        (or plain python for N=1)
Synthetic clouds per SURVEY 8(d): randn(2048,3) * diag(s), s ~ U(0.2,1)^3; reference seed 11, generated seed 13."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402


def synth(seed, start, count, N, device):
    out = torch.empty(count, N, 3)
    for i in range(count):
        g = torch.Generator().manual_seed(seed * 1_000_003 + start + i)
        out[i] = torch.randn(N, 3, generator=g) * (0.2 + 0.8 * torch.rand(3, generator=g))
    return out.to(device)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=512)
    ap.add_argument("--points", type=int, default=2048)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    start, count = pcd_b200.shard_range(args.total, rank, world)
    G, R = synth(13, start, count, args.points, dev), synth(11, start, count, args.points, dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    res = pcd_b200.evaluate_sets(G, R)          # all-gathers + the three CD matrices + reductions, all on the current stream
    e1.record()
    torch.cuda.synchronize()
    dt_dev = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev)
    if world > 1:
        dist.all_reduce(dt_dev, op=dist.ReduceOp.MAX)      # device time, max over ranks
        dist.barrier()
    wall = time.perf_counter() - t0
    dt = float(dt_dev[0])
    if rank == 0:
        n = args.total
        tile = min(512, max(128, -(-n // 32 // 64) * 64))          # evaluate_sets' default block size
        nb = (n + tile - 1) // tile
        # G x R in full + the upper-triangle blocks of the symmetric G x G and R x R (diagonal blocks: their own upper triangle)
        bs = [min(tile, n - i * tile) for i in range(nb)]
        pairs = float(n) * n + 2.0 * sum(bs[i] * (bs[i] + 1) // 2 if i == j else bs[i] * bs[j] for i in range(nb) for j in range(i, nb))
        res.update({"n_gpus": world, "clouds_per_set": args.total, "points": args.points, "seconds": dt,
                    "timing": "CUDA events around evaluate_sets, max over ranks", "wall_seconds": wall,
                    "cloud_pairs_evaluated": pairs, "pairs_if_all_three_matrices_were_full": 3.0 * n * n, "cloud_pairs_per_s": pairs / dt, "evals_per_s": pairs * args.points ** 2 / dt,
                    "all_gather_bytes_per_set": args.total * args.points * 12})
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
