"""GPU parity: VAE3DLarge.decode (implicit-GEMM 3-D convolutions on the tcgen05 kernel) and the voxel -> points
compaction against the oracle / the reference's golden vectors."""
import os

import pytest
import torch

import pcd_b200
from oracle import pointdiff_oracle as O
from conftest import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# per-layer relative L2 of the activations / max abs error of the voxel probabilities, by precision
TAP_TOL = {"bf16x3": 2e-4, "f16mix": 2e-4, "bf16": 4e-2, "f16": 8e-3}
# (measured on B200: taps 1e-5..7e-5 / voxels 2.4e-4 split; taps 3e-3..7e-3 / voxels 7e-2 bf16; 4e-4..9e-4 / 7e-3 f16)
VOX_TOL = {"bf16x3": 5e-4, "f16mix": 5e-4, "bf16": 1.5e-1, "f16": 3e-2}


@pytest.fixture(scope="module")
def vg():
    return torch.load(os.path.join(ROOT, "tests", "golden", "vae3d_golden.pt"), weights_only=True)


@pytest.fixture(scope="module")
def vsd():
    return O.make_synthetic_vae3d_decoder_checkpoint()


def _engine(vsd, precision):
    return pcd_b200.Vae3dEngine(vsd, torch.device("cuda", 0), precision)


def test_voxel_points_bit_exact_vs_reference_golden(vg):
    for key, thr in (("glue.", 0.5),):
        clouds = pcd_b200.voxel_tensor_to_point_clouds(vg[key + "vox"].cuda(), thr)
        assert [len(c) for c in clouds] == vg[key + "counts"].tolist() and clouds[1].shape == (0, 3)
        assert torch.equal(torch.cat(clouds).cpu(), vg[key + "points"])
    vox = O.vae3d_decode(O.make_synthetic_vae3d_decoder_checkpoint(), vg["z"])
    clouds = pcd_b200.voxel_tensor_to_point_clouds(vox.cuda(), vg["threshold"])
    assert [len(c) for c in clouds] == vg["counts"].tolist()
    assert torch.equal(torch.cat(clouds).cpu(), vg["points"])
    # nothing above the threshold at all
    empty = pcd_b200.voxel_tensor_to_point_clouds(torch.zeros(2, 1, 8, 8, 8, device="cuda"), 0.5)
    assert [tuple(c.shape) for c in empty] == [(0, 3), (0, 3)]


@pytest.mark.parametrize("precision", ["bf16x3", "f16mix", "bf16", "f16"])
def test_decoder_taps_vs_oracle(vsd, vg, precision):
    """Every layer of the nn.Sequential against the oracle's activation (localises a wrong tap / parity class / fold)."""
    eng = _engine(vsd, precision)
    z = vg["z"]
    taps = {}
    O.vae3d_decode(vsd, z, taps=taps)
    lin = torch.nn.functional.linear(z, vsd["vae.decoder_input.weight"], vsd["vae.decoder_input.bias"]).view(-1, 512, 4, 4, 4)
    errs = {-1: rel_l2(eng.tap(z.cuda(), -1), lin)}
    for idx in (0, 2, 3, 5, 6, 8, 9, 11):
        errs[idx] = rel_l2(eng.tap(z.cuda(), idx), taps[idx])
    print(precision, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TAP_TOL[precision], errs


@pytest.mark.parametrize("precision", ["bf16x3", "f16mix", "bf16", "f16"])
def test_decode_vs_reference_golden(vsd, vg, precision):
    eng = _engine(vsd, precision)
    want = O.vae3d_decode(vsd, vg["z"])                       # == the reference (tests/test_vae3d_oracle.py)
    assert torch.equal(want.half(), vg["vox"])
    got = eng.decode(vg["z"].cuda()).cpu()
    err = float((got - want).abs().max())
    print(precision, "max abs voxel-probability error", err)
    assert got.shape == want.shape and err < VOX_TOL[precision]
    if precision in ("bf16x3", "f16mix"):
        # occupancy: identical except for voxels whose reference probability sits within the tolerance of the threshold
        thr = vg["threshold"]
        flips = (got > thr) != (want > thr)
        assert bool(((want - thr).abs()[flips] < VOX_TOL[precision]).all())
        assert int(flips.sum()) <= 3


@pytest.mark.parametrize("B", [1, 2, 5])
def test_batch_sizes_and_shard_invariance(vsd, B):
    eng = _engine(vsd, "bf16x3")
    g = torch.Generator().manual_seed(B)
    z = torch.randn(B, 256, generator=g)
    got = eng.decode(z.cuda()).cpu()
    want = O.vae3d_decode(vsd, z)
    assert float((got - want).abs().max()) < VOX_TOL["bf16x3"]
    one = torch.cat([eng.decode(z[i:i + 1].cuda()).cpu() for i in range(B)])
    assert torch.equal(one, got)                              # a sample's grid does not depend on its batch neighbours


def test_latent_diffusion_default_voxel_configuration(vsd):
    """LatentDiffusion(vae=VAE3DLarge, is_voxel_based=True).sample -> list of ragged point clouds (diffusion.py:619-653)."""
    lsd = O.make_synthetic_latent_checkpoint(num_points=256)
    vae = pcd_b200.VAE3DLarge()
    vae.load_state_dict({k[4:]: t for k, t in vsd.items()}, strict=False)
    m = pcd_b200.LatentDiffusion(vae)
    m.load_state_dict({k: v for k, v in lsd.items() if k.startswith("model.")}, strict=False)
    m = m.eval().cuda()
    g = torch.Generator().manual_seed(2)
    zT = torch.randn(3, 256, generator=g)
    z0 = m.sample(3, num_steps=6, z_T=zT, return_latent=True)
    clouds = m.sample(3, num_steps=6, threshold=0.4, z_T=zT)
    assert isinstance(clouds, list) and len(clouds) == 3 and all(c.is_cuda and c.shape[1] == 3 for c in clouds)
    want_vox = O.vae3d_decode(vsd, z0.cpu())
    want = O.voxel_tensor_to_point_clouds(want_vox, 0.4)
    near = int(((want_vox - 0.4).abs() < VOX_TOL["bf16x3"]).sum())
    for c, w in zip(clouds, want):
        a = {tuple(p) for p in c.cpu().tolist()}
        b = {tuple(p) for p in w.tolist()}
        assert len(a ^ b) <= near
        assert float(c.min()) >= -1.0 and float(c.max()) <= 1.0
    # the reference's own VAE class layout is recognised structurally as well (a torch module with the same tree)
    assert pcd_b200.voxel.is_vae3d_large(vae)
