"""Full-size golden vectors (N = 2048, the sizes BASELINE's configs are quoted on) from the UNMODIFIED reference imported
in place (oracle/ref_shim.py).  Run in the build container (takes a few minutes of CPU):

    python tests/golden/make_golden_fullsize.py

* point model, alpha = 1/33 checkpoint (the one SURVEY 8(d) prescribes for DDIM): `PointCloudDiffusion.sample` DDIM-50 on TWO
  2048-point clouds (x_T stored), the first step's predicted noise (a per-forward check at full size), and
  `PointCloudDiffusion.sample2` DDPM-20 on the same x_T with replayed noise (only the generator seed is stored: the test redraws
  the same CPU-generator stream).  The GPU tests run these clouds INSIDE a larger batch -- results are batch independent.
* latent path at num_points = 2048 (BASELINE config 4's decoder: the 6144 x 6144 output layer): `SimplePointNetVAE.decode` and the
  DDIM-8 latent loop + decode on two latents.
Weights are regenerated from seeds (checksums stored)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pointdiff_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fullsize_golden.pt")
NOISE_SEED = 61


def main():
    rd, rn, _ = ref_shim.load_reference()
    out = {}
    N, B = 2048, 2
    g = torch.Generator().manual_seed(55)
    sd = O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 33.0)
    m = rd.PointCloudDiffusion(num_points=N)
    m.load_state_dict(sd, strict=True)
    m.eval()
    xT = torch.randn(B, N, 3, generator=g)
    out["a33.xT"] = xT
    with torch.no_grad():
        out["a33.fwd.t"] = torch.ones(B)
        out["a33.fwd.eps"] = m.model(xT, torch.ones(B))
        with ref_shim.replay_randn([xT]):
            out["a33.ddim50.out"] = m.sample(B, N, num_steps=50)
        S = 20
        gn = torch.Generator().manual_seed(NOISE_SEED)
        noises = [torch.randn(B, N, 3, generator=gn) for _ in range(S - 1)]
        with ref_shim.replay_randn([xT] + noises):
            out["a33.ddpm20.out"] = m.sample2(B, N, num_steps=S)
        out["a33.ddpm20.noise_seed"], out["a33.ddpm20.S"] = NOISE_SEED, S
    del m
    # latent path at 2048 points
    sdl = O.make_synthetic_latent_checkpoint(num_points=N)
    lm = rd.LatentDiffusion(rn.SimplePointNetVAE(num_points=N), is_voxel_based=False)
    res = lm.load_state_dict(sdl, strict=False)
    assert not res.unexpected_keys
    lm.eval()
    out["latent.sd_checksum"] = sum(float(v.double().abs().sum()) for v in sdl.values())
    with torch.no_grad():
        z = torch.randn(2, 256, generator=g)
        out["latent.decode.z"], out["latent.decode.out"] = z, lm.vae.decode(z)
        zT = torch.randn(2, 256, generator=g)
        S = 8
        z_t, z_0 = zT, zT
        for step in range(S):      # the reference's sample() crashes for a point VAE (diffusion.py:650-653): same pieces, same order
            tt = torch.ones(2) - step * (1.0 / S)
            n, s = lm.diffusion_schedule(tt)
            eps = lm.model(z_t, tt)
            z_0 = lm.remove_noise(z_t, eps, n, s)
            n2, s2 = lm.diffusion_schedule(tt - 1.0 / S)
            z_t = s2.view(-1, 1) * z_0 + n2.view(-1, 1) * eps
        out["latent.ddim8.zT"], out["latent.ddim8.out"] = zT, lm.vae.decode(z_0)
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
