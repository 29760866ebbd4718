"""Voxel-VAE decoder (VAE3DLarge.decode): per-layer parity against the oracle and per-launch timing.
    python tools/bench_vae3d.py [--batch 128] [--precisions bf16x3,f16mix,bf16,f16] [--parity-batch 3]
Prints one JSON line per precision: per-tap relative L2 at the parity batch, decode ms and per-layer TFLOP/s at --batch."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from oracle import pointdiff_oracle as O  # noqa: E402  (checker only)


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--parity-batch", type=int, default=3)
    ap.add_argument("--precisions", default="bf16x3,f16mix,bf16,f16")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    sd = O.make_synthetic_vae3d_decoder_checkpoint()
    g = torch.Generator().manual_seed(3)
    zp = torch.randn(args.parity_batch, 256, generator=g)
    taps = {}
    t0 = time.time()
    want = O.vae3d_decode(sd, zp, taps=taps)
    cpu_s = time.time() - t0
    zb = torch.randn(args.batch, 256, generator=g).cuda()
    for prec in args.precisions.split(","):
        eng = pcd_b200.Vae3dEngine(sd, torch.device("cuda", 0), prec)
        errs = {str(i): rel_l2(eng.tap(zp.cuda(), i), taps[i]) for i in (0, 2, 3, 5, 6, 8, 9, 11)}
        got = eng.decode(zp.cuda()).cpu()
        vox_err = float((got - want).abs().max())
        flips = int(((got > 0.4) != (want > 0.4)).sum())
        for _ in range(3):
            eng.decode(zb)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
        ev[0].record()
        for i in range(args.reps):
            vox = eng.decode(zb)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps))[args.reps // 2]
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); clouds = pcd_b200.voxel_tensor_to_point_clouds(vox, 0.4); t1.record(); torch.cuda.synchronize()
        prof = eng.profile(zb)
        agg = {}
        for name, m, fl in prof:
            key = name.split("[")[0]
            a = agg.setdefault(key, [0.0, 0.0]); a[0] += m; a[1] += fl
        total_fl = sum(fl for _, _, fl in prof)
        print(json.dumps({
            "precision": prec, "batch": args.batch, "decode_ms": ms, "shapes_per_s": args.batch / ms * 1e3,
            "algorithmic_tflops": total_fl / ms * 1e-9, "gflop_per_sample": total_fl / args.batch * 1e-9,
            "voxel_to_points_ms": t0.elapsed_time(t1), "mean_points": sum(len(c) for c in clouds) / len(clouds),
            "parity": {"batch": args.parity_batch, "tap_rel_l2": errs, "vox_max_abs": vox_err, "occupancy_flips": flips},
            "cpu_oracle_s_per_sample": cpu_s / args.parity_batch,
            "layers": {k: {"ms": round(v[0], 4), "tflops": round(v[1] / v[0] * 1e-9, 1) if v[0] > 0 else 0} for k, v in agg.items()},
        }), flush=True)
        eng.close()
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
