"""Long latent loops on the persistent kernel: DDIM-1000 and DDPM-1000 (Philox) at batch 4 stay finite and agree with the legacy
fp32 CUDA-core path within the amplification of rounding noise over 1000 steps."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200
from oracle import pointdiff_oracle as O
sd = O.make_synthetic_latent_checkpoint(num_points=256)
m = pcd_b200.LatentDiffusion(pcd_b200.SimplePointNetVAE(256), is_voxel_based=False)
m.load_state_dict(sd, strict=False)
m = m.eval().cuda()
z = torch.randn(4, 256, generator=torch.Generator().manual_seed(1))
def rel(a, b): return float((a - b).norm() / b.norm())
for name, fn in (("ddim1000", lambda: m.sample(4, num_steps=1000, z_T=z, return_latent=True)),
                 ("ddpm1000", lambda: m.sample2(4, num_steps=1000, z_T=z, seed=3, return_latent=True))):
    os.environ.pop("PCD_LATENT_LEGACY", None)
    a = fn(); torch.cuda.synchronize()
    os.environ["PCD_LATENT_LEGACY"] = "1"
    b = fn(); torch.cuda.synchronize()
    os.environ.pop("PCD_LATENT_LEGACY", None)
    print(name, "finite", bool(torch.isfinite(a).all()), "rel vs legacy", rel(a, b))
