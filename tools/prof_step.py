"""Target program for ncu: a few eager denoiser steps of the point model at the bench shape (batch 512 x 2048 points).
    python tools/prof_step.py [precision=f16mix] [steps=3] [batch=512]
One step = 32 launches; the gemm_tc_kernel launches of a step are, in order: enc1.conv2, enc1.conv3, enc2.conv1-3, enc3.conv1-3,
enc4.conv1-3, global_feat.0, global_feat.3+maxpool (index 12), dec4.conv1 (13), dec4.conv2 (14), ... output.0+sampler (25)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from oracle import pointdiff_oracle as O  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "f16mix"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = int(sys.argv[3]) if len(sys.argv) > 3 else 512
sd = O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 33.0)
m = pcd_b200.PointCloudDiffusion(2048, precision=precision)
m.load_state_dict(sd, strict=True)
m = m.eval().cuda()
x = torch.randn(B, 2048, 3, device="cuda")
t = torch.full((B,), 0.5, device="cuda")
for _ in range(steps):
    eps = m.model(x, t)
torch.cuda.synchronize()
print("ok", float(eps.abs().mean()))
