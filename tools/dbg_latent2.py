"""Temporary: persistent latent kernel vs the legacy CUDA-graph path (DDPM + large batch)."""
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
def both(fn):
    os.environ.pop("PCD_LATENT_LEGACY", None)
    a = fn(); torch.cuda.synchronize()
    os.environ["PCD_LATENT_LEGACY"] = "1"
    b = fn(); torch.cuda.synchronize()
    os.environ.pop("PCD_LATENT_LEGACY", None)
    return a, b
mode = sys.argv[1]
if mode == "ddpm":
    B = 4
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, 256, generator=g)
    for S in (2, 3, 8):
        noise = torch.randn(S - 1, B, 256, generator=g)
        a, b = both(lambda: m.sample2(B, num_steps=S, z_T=z, noise=noise, return_latent=True))
        print(f"ddpm injected S={S} rel {rel(a, b):.3e} rows {[round(rel(a[i], b[i]), 7) for i in range(B)]}", flush=True)
        a, b = both(lambda: m.sample2(B, num_steps=S, z_T=z, seed=5, return_latent=True))
        print(f"ddpm philox S={S} rel {rel(a, b):.3e} rows {[round(rel(a[i], b[i]), 7) for i in range(B)]}", flush=True)
else:
    B = int(mode)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, 256, generator=g).cuda()
    t = torch.rand(B, generator=g).cuda()
    a, b = both(lambda: m.engine().forward(z, t))
    print(f"B={B} forward rel {rel(a, b):.3e}", flush=True)
    for S in (1, 2, 5, 50):
        a, b = both(lambda: m.sample(B, num_steps=S, z_T=z.cpu(), return_latent=True))
        print(f"B={B} ddim S={S} rel {rel(a, b):.3e}", flush=True)
    a, b = both(lambda: m.engine().decode(z))
    print(f"B={B} decode rel {rel(a, b):.3e}", flush=True)
