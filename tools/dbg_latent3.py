import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200
from oracle import pointdiff_oracle as O
NP = 256
sd = O.make_synthetic_latent_checkpoint(num_points=NP)
m = pcd_b200.LatentDiffusion(pcd_b200.SimplePointNetVAE(NP), is_voxel_based=False)
m.load_state_dict(sd, strict=False)
m = m.eval().cuda()
B = int(sys.argv[1])
g = torch.Generator().manual_seed(3)
z = torch.randn(B, 256, generator=g).cuda()
t = torch.rand(B, generator=g).cuda()
try:
    a = m.engine().forward(z, t); torch.cuda.synchronize()
    print(f"B={B} maxops={os.environ.get('PCD_LT_MAXOPS')} OK", flush=True)
except Exception as e:
    print(f"B={B} maxops={os.environ.get('PCD_LT_MAXOPS')} FAIL {str(e)[:60]}", flush=True)
