"""Does tcgen05 kind::tf32 ignore the 13 low mantissa bits of its fp32 containers?  Compare the persistent latent kernel with the
hi plane masked (default) against the raw values left in place (PCD_LT_DBG=32): identical bits <=> the hardware truncates."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200
from oracle import pointdiff_oracle as O
sd = O.make_synthetic_latent_checkpoint(num_points=256)
m = pcd_b200.LatentDiffusion(pcd_b200.SimplePointNetVAE(256), is_voxel_based=False)
m.load_state_dict(sd, strict=False)
m = m.eval().cuda()
g = torch.Generator().manual_seed(3)
z = torch.randn(128, 256, generator=g).cuda(); t = torch.rand(128, generator=g).cuda()
a = m.engine().forward(z, t); torch.cuda.synchronize()
os.environ["PCD_LT_DBG"] = "32"
b = m.engine().forward(z, t); torch.cuda.synchronize()
ref = O.latent_denoiser_forward(sd, z.cpu()[:16], t.cpu()[:16])
print("identical bits:", bool(torch.equal(a, b)), " max abs diff", float((a - b).abs().max()),
      " rel vs oracle masked", float((a[:16].cpu() - ref).norm() / ref.norm()), " raw", float((b[:16].cpu() - ref).norm() / ref.norm()))
