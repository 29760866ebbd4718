"""CPU: oracle restatement of the latent path against golden vectors from the reference."""
import pytest
import torch

from oracle import pointdiff_oracle as O
from oracle import ref_shim


@pytest.fixture(scope="module")
def lg():
    import os
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "latent_golden.pt"), weights_only=True)


@pytest.fixture(scope="module")
def lsd(lg):
    sd = O.make_synthetic_latent_checkpoint(num_points=int(lg["num_points"]))
    assert abs(sum(float(v.double().abs().sum()) for v in sd.values()) - lg["sd_checksum"]) < 1e-6 * lg["sd_checksum"]
    return sd


def test_latent_forward_decode_and_loops_match_reference_golden(lg, lsd):
    NP = int(lg["num_points"])
    assert torch.equal(O.latent_denoiser_forward(lsd, lg["fwd.z"], lg["fwd.t"]), lg["fwd.eps"])
    assert torch.equal(O.vae_decode(lsd, lg["decode.z"], NP), lg["decode.out"])
    S = int(lg["ddpm.S"])
    assert torch.equal(O.latent_ddpm_sample(lsd, lg["ddpm.zT"], list(lg["ddpm.noise"]), S, NP), lg["ddpm.out"])
    assert torch.equal(O.latent_ddim_sample(lsd, lg["ddpm.zT"], S, NP, decode=False), lg["ddim.z0"])
    assert torch.equal(O.latent_ddim_sample(lsd, lg["ddpm.zT"], S, NP), lg["ddim.out"])


def test_groupnorm_semantics_on_2d_input():
    # GroupNorm(8, C) on [B, C] = per-row standardisation over each contiguous C/8 channels, eps 1e-5
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 64, generator=g)
    w, b = torch.randn(64, generator=g), torch.randn(64, generator=g)
    y = torch.nn.functional.group_norm(x, 8, w, b, 1e-5)
    xg = x.view(3, 8, 8)
    ref = ((xg - xg.mean(-1, keepdim=True)) / torch.sqrt(xg.var(-1, unbiased=False, keepdim=True) + 1e-5)).view(3, 64) * w + b
    assert torch.allclose(y, ref, atol=1e-6)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_latent_state_dict_interoperates_and_reference_sample_crashes(lsd, lg):
    import pcd_b200
    rd, rn, _ = ref_shim.load_reference()
    NP = int(lg["num_points"])
    ref = rd.LatentDiffusion(rn.SimplePointNetVAE(num_points=NP), is_voxel_based=False)
    mine = pcd_b200.LatentDiffusion(pcd_b200.SimplePointNetVAE(NP), is_voxel_based=False)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(mine.state_dict(), strict=True)
    assert dict(mine.hparams) == {k: ref.hparams[k] for k in mine.hparams}
    with pytest.raises(UnboundLocalError):      # the reference's DDIM path is broken for point VAEs
        ref.eval()
        with torch.no_grad():
            ref.sample(2, num_steps=2)


def test_folding_decoder_matches_reference_golden(lg):
    """FoldingDecoder.forward (networks.py:1484-1509): bit-identical to the reference's output."""
    NPf = int(lg["fold.num_points"])
    fsd = O.make_synthetic_folding_checkpoint(num_points=NPf)
    assert abs(sum(float(v.double().abs().sum()) for v in fsd.values()) - lg["fold.sd_checksum"]) < 1e-6 * lg["fold.sd_checksum"]
    assert torch.equal(O.folding_grid(), lg["fold.grid"])
    out = O.folding_decode(fsd, lg["fold.z"])
    assert out.shape == (3, NPf, 3) and torch.equal(out, lg["fold.out"])


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_folding_decoder_container_matches_reference_keys():
    import pcd_b200
    _, rn, _ = ref_shim.load_reference()
    ref = rn.FoldingDecoder(256, 300)
    mine = pcd_b200.FoldingDecoder(256, 300)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict(), strict=True)
    assert torch.equal(mine.grid, ref.grid)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_latent_linear_schedule_loops_match_reference(lsd, lg):
    """noise_schedule='linear' in the latent loops: the reference's linear_diffusion_schedule cumprods over the BATCH axis
    (diffusion.py:553-569), so every sample of a batch gets its own rates.  Oracle == reference, bit for bit."""
    rd, rn, _ = ref_shim.load_reference()
    NP = int(lg["num_points"])
    m = rd.LatentDiffusion(rn.SimplePointNetVAE(num_points=NP), is_voxel_based=False, noise_schedule="linear")
    m.load_state_dict(lsd, strict=False)
    m.eval()
    g = torch.Generator().manual_seed(17)
    B, S = 5, 6
    zT = torch.randn(B, 256, generator=g)
    noises = [torch.randn(B, 256, generator=g) for _ in range(S - 1)]
    with torch.no_grad():
        with ref_shim.replay_randn([zT] + noises):
            ref_ddpm = m.sample2(B, num_steps=S)
        z_t, z_0 = zT, zT
        for step in range(S):                      # the reference's sample() crashes for a point VAE: its own pieces instead
            tt = torch.ones(B) - step * (1.0 / S)
            n, s = m.diffusion_schedule(tt)
            eps = m.model(z_t, tt)
            z_0 = m.remove_noise(z_t, eps, n, s)
            n2, s2 = m.diffusion_schedule(tt - 1.0 / S)
            z_t = s2.view(-1, 1) * z_0 + n2.view(-1, 1) * eps
    assert torch.equal(O.latent_ddpm_sample(lsd, zT, noises, S, NP, schedule="linear"), ref_ddpm)
    assert torch.equal(O.latent_ddim_sample(lsd, zT, S, NP, decode=False, schedule="linear"), z_0)
    n, _ = m.diffusion_schedule(torch.full((B,), 0.5))
    assert len(set(float(v) for v in n)) == B          # one rate per sample: the batch-axis cumprod


def test_oracle_latent_ddpm1000_is_bit_identical_to_the_reference_golden():
    """1000 reverse steps of `LatentDiffusion.sample2` + decode with replayed noise: the oracle restatement reproduces the unmodified
    reference's output bit for bit (golden: tests/golden/make_golden_ddpm1000.py), so the loop arithmetic the GPU path is checked
    against is the reference's own -- schedule, posterior update, noise injection -- over the metric's full length."""
    import os
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ddpm1000_golden.pt"), weights_only=True)
    B, S, NP = int(g["B"]), int(g["S"]), int(g["latent.num_points"])
    sdl = O.make_synthetic_latent_checkpoint(num_points=NP)
    zT = torch.randn(B, 256, generator=torch.Generator().manual_seed(int(g["latent.zT_seed"])))
    gl = torch.Generator().manual_seed(int(g["latent.noise_seed"]))
    noises = [torch.randn(B, 256, generator=gl) for _ in range(S - 1)]
    assert torch.equal(O.latent_ddpm_sample(sdl, zT, noises, S, NP), g["latent.out"])
