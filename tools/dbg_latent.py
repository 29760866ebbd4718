"""Temporary: persistent latent kernel vs the legacy CUDA-graph path."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200
from oracle import pointdiff_oracle as O

def rel(a, b): return float((a - b).norm() / b.norm())

NP = 256
sd = O.make_synthetic_latent_checkpoint(num_points=NP)
m = pcd_b200.LatentDiffusion(pcd_b200.SimplePointNetVAE(NP), is_voxel_based=False)
m.load_state_dict(sd, strict=False)
m = m.eval().cuda()
Bs = [int(a) for a in sys.argv[1:]] or [4, 128]
for B in Bs:
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, 256, generator=g).cuda()
    t = torch.rand(B, generator=g).cuda()
    def both(fn):
        os.environ.pop("PCD_LATENT_LEGACY", None)
        a = fn(); torch.cuda.synchronize()
        os.environ["PCD_LATENT_LEGACY"] = "1"
        b = fn(); torch.cuda.synchronize()
        os.environ.pop("PCD_LATENT_LEGACY", None)
        return a, b
    a, b = both(lambda: m.engine().forward(z, t))
    print(f"B={B} forward rel {rel(a, b):.3e}", flush=True)
    for S in (1, 2, 5):
        a, b = both(lambda: m.sample(B, num_steps=S, z_T=z.cpu(), return_latent=True))
        print(f"B={B} ddim S={S} rel {rel(a, b):.3e}  rows bad: {[(i, round(rel(a[i], b[i]), 6)) for i in range(min(B, 8))]}", flush=True)
    a, b = both(lambda: m.engine().decode(z))
    print(f"B={B} decode rel {rel(a, b):.3e}", flush=True)
