"""f16mix experiments: eps error against the fp32 oracle (golden inputs + a 2 x 2048 cloud) and reverse-step time at batch 512
for a list of PCD_MIX_NP2 / PCD_MIX_EXTRA settings (read when the denoiser handle is created)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200
from oracle import pointdiff_oracle as O

def rel(a, b): return float((a.double().cpu() - b.double()).norm() / b.double().norm())

sd = O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 33.0)
g = torch.Generator().manual_seed(77)
xs = [(torch.randn(2, 2048, 3, generator=g), torch.tensor([0.3, 0.9])), (torch.randn(3, 1000, 3, generator=g), torch.tensor([0.05, 0.5, 1.0]))]
refs = [O.denoiser_forward(sd, x, t) for x, t in xs]
configs = [("", "")] + [(c, "") for c in sys.argv[1:]]
for np2, extra in configs:
    os.environ.pop("PCD_MIX_NP2", None)
    if np2: os.environ["PCD_MIX_NP2"] = np2
    m = pcd_b200.PointCloudDiffusion(2048, precision="f16mix")
    m.load_state_dict(sd, strict=True)
    m = m.eval().cuda()
    errs = [rel(m.model(x.cuda(), t.cuda()), r) for (x, t), r in zip(xs, refs)]
    B, S = 512, 6
    xT = torch.randn(B, 2048, 3, generator=g).cuda()
    m.sample(B, 2048, num_steps=2, x_T=xT)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.sample(B, 2048, num_steps=S, x_T=xT); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / S
    print(json.dumps({"np2": np2, "eps_rel_l2": errs, "ms_per_reverse_step_b512": ms, "ddim50_shapes_per_s": B / (ms * 50) * 1e3}), flush=True)
    del m
    torch.cuda.empty_cache()
