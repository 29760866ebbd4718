"""f16mix layer-plan experiments (one process per setting: the plan is read from the environment when the handle is created).

For every setting: tap / eps error against the fp32 oracle on the alpha = 1/33 checkpoint at N = 2048, the per-layer profile
and the DDIM loop rate at batch 512.  Usage:
    python tools/exp_mix.py                      # runs the default list of settings, one subprocess each
    python tools/exp_mix.py one                  # this process, settings from the environment
    python tools/exp_mix.py '[{"PCD_TILE_ORDER": "0"}, {"PCD_TILE_ORDER": "1", "EXP_PRECISION": "bf16"}]'     # a list of environments
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SETTINGS = [
    {"PCD_MIX_C8": "none"},
    {},                                                              # default c8 set
    {"PCD_MIX_C8": "e3c1,e3c2,e3c3,e4c1,e4c2,e4c3,d4c2,d4c3,d3c1,d3c2,d3c3,d2c1,d2c2,d2c3"},
]


def one():
    import torch
    import pcd_b200
    from oracle import pointdiff_oracle as O

    def rel(a, b):
        return float((a.double().cpu() - b.double()).norm() / b.double().norm())

    precision = os.environ.get("EXP_PRECISION", "f16mix")
    os.environ["PCD_TAPS"] = "1"
    sd = O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 33.0)
    g = torch.Generator().manual_seed(77)
    B, N = 2, 2048
    x, t = torch.randn(B, N, 3, generator=g), torch.tensor([0.3, 0.9])
    taps = {}
    ref = O.denoiser_forward(sd, x, t, taps=taps)
    m = pcd_b200.PointCloudDiffusion(N, precision=precision)
    m.load_state_dict(sd, strict=True)
    m = m.eval().cuda()
    eps = m.model(x.cuda(), t.cuda())
    torch.cuda.synchronize()
    eng = m.model.engine()
    rec = {"env": {k: v for k, v in os.environ.items() if k.startswith("PCD_") and k != "PCD_TAPS"}, "precision": precision}
    for name, C in (("x1", 128), ("x2", 256), ("x3", 512), ("x4", 1024), ("d4", 512), ("d1", 64)):
        rec[name] = rel(eng.tap(name, (B, N, C)), taps[name].transpose(1, 2))
        rec[name + "_absmean"] = float(taps[name].abs().mean())
    rec["eps"] = rel(eps, ref)
    del m, eng
    torch.cuda.empty_cache()
    os.environ.pop("PCD_TAPS")
    m = pcd_b200.PointCloudDiffusion(N, precision=precision)
    m.load_state_dict(sd, strict=True)
    m = m.eval().cuda()
    eng = m.model.engine()
    Bb = int(os.environ.get("EXP_BATCH", "512"))
    xb, tb = torch.randn(Bb, N, 3, device="cuda"), torch.full((Bb,), 0.5, device="cuda")
    eng.profile(xb, tb)
    rows = eng.profile(xb, tb)
    rec["profile_ms"] = {r[0]: round(r[1], 4) for r in rows}
    rec["profile_step_ms"] = sum(r[1] for r in rows)
    S = 8
    m.sample(Bb, N, num_steps=2, x_T=xb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.sample(Bb, N, num_steps=S, x_T=xb); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / S
    rec["loop_ms_per_step"] = ms
    rec["ddim50_shapes_per_s"] = Bb / (ms * 50) * 1e3
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        settings = json.loads(sys.argv[1]) if len(sys.argv) > 1 else SETTINGS      # e.g. '[{"PCD_TILE_ORDER": "0"}, {"PCD_TILE_ORDER": "1"}]'
        for st in settings:
            env = dict(os.environ)
            env.update(st)
            r = subprocess.run([sys.executable, __file__, "one"], env=env, timeout=900)
            if r.returncode != 0:
                print(json.dumps({"env": st, "failed": r.returncode}), flush=True)
