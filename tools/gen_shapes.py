"""BASELINE config 3: batch-sharded DDPM-1000 generation of `--total` shapes across the ranks of one box
(one process per GPU, no in-loop communication; Philox noise keyed by the global sample index).

    torchrun --nproc-per-node N tools/gen_shapes.py --total 8192 [--steps 1000] [--kind ddpm]
Prints one JSON line on rank 0: whole-job shapes/s (max over ranks), plus a checksum of sample 0 that must
not depend on N."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--kind", default="ddpm", choices=["ddpm", "ddim"])
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--max-batch", type=int, default=512)
    ap.add_argument("--precision", default="f16mix")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = pcd_b200.PointCloudDiffusion(args.points, precision=args.precision)
    model.load_state_dict(pcd_b200.synthetic_state_dict(model, alpha=1.0 / 3300.0), strict=True)
    model = model.eval().to(dev)
    # warm-up: plan + graph for the batch sizes that will be used
    pcd_b200.sample_sharded(model, min(args.total, world * 2), args.points, 3, args.kind, seed=5)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out, start = pcd_b200.sample_sharded(model, args.total, args.points, args.steps, args.kind, seed=5, max_batch=args.max_batch)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    finite = torch.tensor([float(torch.isfinite(out).all())], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"workload": f"{args.kind.upper()}-{args.steps} generation of {args.total} shapes x {args.points} pts, {args.precision}",
                          "n_gpus": world, "seconds": float(dt), "shapes_per_s": args.total / float(dt),
                          "shapes_this_rank": int(out.shape[0]), "all_finite": bool(finite.item()),
                          "sample0_checksum": float(out[0].double().abs().sum()), "sample0_first_point": out[0, 0].tolist()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
