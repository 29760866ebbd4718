"""GPU: parity at the sizes BASELINE's configs are quoted on (N = 2048) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden_fullsize.py), on the alpha = 1/33 checkpoint SURVEY 8(d) prescribes for DDIM -- the one where
the denoiser's output actually steers the trajectory (on the alpha = 1/3300 checkpoint eps is ~0 and a loop test says little
about 16-bit drift).

The two golden clouds are sampled INSIDE a batch of 6 (rows 1 and 4): a cloud's result does not depend on its batch
(tests/test_gpu_samplers.py::test_full_size_ddim50_properties asserts that bit for bit at batch 512), so this pins config 2's
per-cloud arithmetic at full size.

Stated tolerances (north_star: per-step eps within 1e-3 relative L2 for the 16-bit path / 1e-5 for fp32 mode; final DDIM samples
within a stated Chamfer tolerance).  Measured, the final DDIM-50 sample's rel-L2 is about the per-step eps error (the returned
x0 is dominated by the last steps' predictions) and DDPM-20 with replayed noise sits ~4x above it.  Chamfer values are in the reference's units (x1e3, cube-normalised); two
DIFFERENT reference samples of this checkpoint are 102 apart (printed by tools/measure_fullsize.py).  Measured values:
profiles/fullsize_parity_r2.jsonl."""
import os

import pytest
import torch

import pcd_b200
from oracle import pointdiff_oracle as O
from conftest import rel_l2

pytestmark = pytest.mark.gpu
N = 2048
ROWS = [1, 4]

#            per-step eps rel-L2   DDIM-50 rel-L2   DDIM-50 Chamfer   DDPM-20 rel-L2      measured (profiles/fullsize_parity_r2.jsonl)
BOUNDS = {
    "fp32":   (1e-5,                 2e-5,            0.05,             1e-4),            # 3.5e-6   2.3e-6   0.004   1.2e-5
    "bf16x3": (1e-3,                 3e-4,            0.5,              1e-3),            # 4.6e-5   4.2e-5   0.10    1.6e-4
    "f16mix": (1e-3,                 2e-3,            2.5,              1e-2),            # 6.2e-4   4.8e-4   0.79    2.6e-3
    "f16":    (6e-3,                 1.5e-2,          15.0,             5e-2),            # 3.0e-3   3.6e-3   5.1     1.2e-2
    "bf16":   (3e-2,                 1.2e-1,          80.0,             3.5e-1),          # 2.3e-2   3.4e-2   42      1.0e-1
}
# Reading the Chamfer column: two DIFFERENT reference samples of this checkpoint are 102 apart, so single-pass bf16 (42) lands
# nearer to the reference sample than to an unrelated one but visibly off it; f16mix -- the default precision and the bench
# headline -- is within 1 (1 %), the split modes within 0.1.


@pytest.fixture(scope="module")
def fg():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "fullsize_golden.pt"), weights_only=True)


@pytest.fixture(scope="module")
def batch(fg):
    g = torch.Generator().manual_seed(56)
    filler = torch.randn(4, N, 3, generator=g)
    xT = torch.cat([filler[:1], fg["a33.xT"][:1], filler[1:3], fg["a33.xT"][1:], filler[3:]])
    S = int(fg["a33.ddpm20.S"])
    gn = torch.Generator().manual_seed(int(fg["a33.ddpm20.noise_seed"]))
    noise = torch.zeros(S - 1, 6, N, 3)
    noise[:, ROWS] = torch.stack([torch.randn(2, N, 3, generator=gn) for _ in range(S - 1)])
    return xT, noise, S


@pytest.mark.parametrize("precision", list(BOUNDS))
def test_fullsize_sd33_vs_reference_golden(fg, batch, sd33, precision):
    b_eps, b_ddim, b_cd, b_ddpm = BOUNDS[precision]
    xT, noise, S = batch
    m = pcd_b200.PointCloudDiffusion(N, precision=precision)
    m.load_state_dict(sd33, strict=True)
    m = m.eval().cuda()
    eps = m.model(xT.cuda(), torch.ones(6).cuda())[ROWS]
    assert rel_l2(eps, fg["a33.fwd.eps"]) < b_eps
    out = m.sample(6, N, num_steps=50, x_T=xT)[ROWS]                     # reference diffusion.py:261-289, DDIM-50
    assert bool(torch.isfinite(out).all())
    assert rel_l2(out, fg["a33.ddim50.out"]) < b_ddim
    cd = pcd_b200.chamfer_distance_per_pair(out, fg["a33.ddim50.out"].cuda())
    assert float(cd.max()) < b_cd, cd
    out2 = m.sample2(6, N, num_steps=S, x_T=xT, noise=noise)[ROWS]       # reference diffusion.py:225-259 with replayed noise
    assert rel_l2(out2, fg["a33.ddpm20.out"]) < b_ddpm


def test_fullsize_latent_decode_and_loop_2048(fg):
    """BASELINE config 4's decoder at its real size: SimplePointNetVAE(2048).decode (6144 x 6144 output layer: different tiles and
    split-K than the 256-point goldens) and the DDIM-8 latent loop + decode, against the reference's outputs."""
    sdl = O.make_synthetic_latent_checkpoint(num_points=N)
    assert abs(sum(float(v.double().abs().sum()) for v in sdl.values()) - float(fg["latent.sd_checksum"])) < 1e-6 * float(fg["latent.sd_checksum"])
    m = pcd_b200.LatentDiffusion(pcd_b200.SimplePointNetVAE(N), is_voxel_based=False)
    assert not m.load_state_dict(sdl, strict=False).unexpected_keys
    m = m.eval().cuda()
    out = m.engine().decode(fg["latent.decode.z"].cuda())
    assert out.shape == (2, N, 3) and rel_l2(out, fg["latent.decode.out"]) < 2e-5
    pts = m.sample(2, num_steps=8, z_T=fg["latent.ddim8.zT"])
    assert pts.shape == (2, N, 3) and rel_l2(pts, fg["latent.ddim8.out"]) < 5e-4
    # the standalone VAE module decodes through its own decoder-only handle (reference networks.py:1219-1231)
    vae = pcd_b200.SimplePointNetVAE(N)
    vae.load_state_dict({k[len("vae."):]: v for k, v in sdl.items() if k.startswith("vae.")}, strict=False)
    assert rel_l2(vae.eval().cuda().decode(fg["latent.decode.z"].cuda()), fg["latent.decode.out"]) < 2e-5


#              DDPM-1000 rel-L2 bound      measured (profiles/ddpm1000_parity_r2.jsonl)
DDPM1000 = {
    "fp32":   2e-5,                      # 2.8e-6
    "bf16x3": 3e-4,                      # 4.9e-5
    "f16mix": 5e-3,                      # 9.6e-4
    "f16":    2.5e-2,                    # 4.9e-3
    "bf16":   2e-1,                      # 4.3e-2
}


@pytest.mark.parametrize("precision", list(DDPM1000))
def test_ddpm1000_full_length_vs_reference_golden(sd3300, precision):
    """The metric's other loop at its real length: `sample2` with 1000 reverse steps (reference diffusion.py:225-259), x_T and all 999
    noise draws replayed, against the UNMODIFIED reference's output (tests/golden/make_golden_ddpm1000.py; 64 points keep the
    reference's run to a minute).  alpha = 1/3300: SURVEY 8(d)'s DDPM checkpoint -- with alpha = 1/33 the reference itself overflows
    to NaN within 1000 steps.  The state grows to |x| ~ 900 here, still inside fp16's range for every layer."""
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ddpm1000_golden.pt"), weights_only=True)
    B, n, S = int(g["B"]), int(g["N"]), int(g["S"])
    assert abs(sum(float(v.double().abs().sum()) for v in sd3300.values()) - float(g["sd_checksum"])) < 1e-6 * float(g["sd_checksum"])
    xT = torch.randn(B, n, 3, generator=torch.Generator().manual_seed(int(g["xT_seed"])))
    gn = torch.Generator().manual_seed(int(g["noise_seed"]))
    noise = torch.stack([torch.randn(B, n, 3, generator=gn) for _ in range(S - 1)])
    m = pcd_b200.PointCloudDiffusion(n, precision=precision)
    m.load_state_dict(sd3300, strict=True)
    out = m.eval().cuda().sample2(B, n, num_steps=S, x_T=xT, noise=noise)
    assert bool(torch.isfinite(out).all())
    assert rel_l2(out, g["out"]) < DDPM1000[precision]
