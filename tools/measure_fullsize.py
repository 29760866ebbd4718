"""Prints the full-size (N = 2048, alpha = 1/33) parity numbers of every precision mode against tests/golden/fullsize_golden.pt
(the UNMODIFIED reference's outputs): per-forward eps error, DDIM-50 and DDPM-20 final-sample rel-L2 and Chamfer distance.
The bounds asserted in tests/test_gpu_fullsize.py were set from this table (profiles/fullsize_parity_r2.jsonl)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from oracle import pointdiff_oracle as O  # noqa: E402


def rel(a, b):
    return float((a.double().cpu() - b.double()).norm() / b.double().norm())


def main():
    fg = torch.load(os.path.join(ROOT, "tests", "golden", "fullsize_golden.pt"), weights_only=True)
    sd = O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 33.0)
    N = 2048
    g = torch.Generator().manual_seed(56)
    filler = torch.randn(4, N, 3, generator=g)
    xT = torch.cat([filler[:1], fg["a33.xT"][:1], filler[1:3], fg["a33.xT"][1:], filler[3:]])       # golden clouds at rows 1 and 4
    rows = [1, 4]
    S = int(fg["a33.ddpm20.S"])
    gn = torch.Generator().manual_seed(int(fg["a33.ddpm20.noise_seed"]))
    noise2 = torch.stack([torch.randn(2, N, 3, generator=gn) for _ in range(S - 1)])
    noise = torch.zeros(S - 1, 6, N, 3)
    noise[:, rows] = noise2
    for precision in sys.argv[1:] or ["fp32", "bf16x3", "f16mix", "f16", "bf16"]:
        m = pcd_b200.PointCloudDiffusion(N, precision=precision)
        m.load_state_dict(sd, strict=True)
        m = m.eval().cuda()
        eps = m.model(xT.cuda(), torch.ones(6).cuda())[rows]
        d50 = m.sample(6, N, num_steps=50, x_T=xT)[rows]
        p20 = m.sample2(6, N, num_steps=S, x_T=xT, noise=noise)[rows]
        rec = {"precision": precision, "eps_rel_l2": rel(eps, fg["a33.fwd.eps"]),
               "ddim50_rel_l2": rel(d50, fg["a33.ddim50.out"]),
               "ddim50_cd": [float(v) for v in pcd_b200.chamfer_distance_per_pair(d50, fg["a33.ddim50.out"].cuda())],
               "ddpm20_rel_l2": rel(p20, fg["a33.ddpm20.out"]),
               "ddpm20_cd": [float(v) for v in pcd_b200.chamfer_distance_per_pair(p20, fg["a33.ddpm20.out"].cuda())],
               "ref_absmax": float(fg["a33.ddim50.out"].abs().max()),
               "cd_between_the_two_reference_clouds": float(pcd_b200.chamfer_distance(fg["a33.ddim50.out"][:1].cuda(), fg["a33.ddim50.out"][1:].cuda()))}
        print(json.dumps(rec), flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
