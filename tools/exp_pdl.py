"""A/B of programmatic dependent launch (PCD_PDL=0 / 1, read once per process) on the point model's reverse loop:
batch 4 DDPM-1000 (BASELINE configs[0], launch-gap bound) and batch 512 DDIM-50 (configs[1]).  One subprocess per setting."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    import pcd_b200
    from oracle import pointdiff_oracle as O
    precision = os.environ.get("EXP_PRECISION", "f16mix")
    N = 2048
    for B, kind, S, alpha in ((4, "ddpm", 1000, 1 / 3300.0), (64, "ddpm", 200, 1 / 3300.0), (512, "ddim", 50, 1 / 33.0)):
        sd = O.make_synthetic_checkpoint(seed=24, alpha=alpha)
        m = pcd_b200.PointCloudDiffusion(N, precision=precision)
        m.load_state_dict(sd, strict=True)
        m = m.eval().cuda()
        xT = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(5)).cuda()
        fn = (lambda s: m.sample2(B, N, num_steps=s, x_T=xT, seed=5)) if kind == "ddpm" else (lambda s: m.sample(B, N, num_steps=s, x_T=xT))
        fn(4)
        fn(S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(S)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"pdl": os.environ.get("PCD_PDL", "1"), "precision": precision, "batch": B, "loop": f"{kind}-{S}", "ms": ms,
                          "us_per_reverse_step": ms / S * 1e3, "shapes_per_s_at_full_loop": B / (ms / S * (1000 if kind == "ddpm" else 50)) * 1e3,
                          "checksum": float(out.double().abs().sum()), "finite": bool(torch.isfinite(out).all())}), flush=True)
        m.model.engine().close()
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for prec in ("f16mix", "bf16"):
            for pdl in ("0", "1"):
                env = dict(os.environ, PCD_PDL=pdl, EXP_PRECISION=prec)
                r = subprocess.run([sys.executable, __file__, "one"], env=env, timeout=900)
                if r.returncode != 0:
                    print(json.dumps({"pdl": pdl, "precision": prec, "failed": r.returncode}), flush=True)
