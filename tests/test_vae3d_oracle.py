"""CPU: the voxel-VAE decoder oracle against the reference's golden vectors and, where the
reference tree is mounted, against the reference's own code; host-side container logic."""
import os

import pytest
import torch

import pcd_b200
from oracle import pointdiff_oracle as O
from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vg():
    return torch.load(os.path.join(ROOT, "tests", "golden", "vae3d_golden.pt"), weights_only=True)


@pytest.fixture(scope="module")
def vsd(vg):
    sd = O.make_synthetic_vae3d_decoder_checkpoint()
    assert sum(float(v.double().abs().sum()) for v in sd.values()) == vg["sd_checksum"]
    return sd


def test_decode_matches_reference_golden(vg, vsd):
    vox = O.vae3d_decode(vsd, vg["z"])
    assert vox.shape == (3, 1, 32, 32, 32)
    assert float(vox.double().sum()) == float(vg["vox_sum"])
    assert torch.equal(vox.half(), vg["vox"])
    clouds = O.voxel_tensor_to_point_clouds(vox, threshold=vg["threshold"])
    assert [len(c) for c in clouds] == vg["counts"].tolist()
    assert torch.equal(torch.cat(clouds), vg["points"])


def test_voxel_glue_matches_reference_golden(vg):
    clouds = O.voxel_tensor_to_point_clouds(vg["glue.vox"])
    assert [len(c) for c in clouds] == vg["glue.counts"].tolist() and clouds[1].shape == (0, 3)
    assert torch.equal(torch.cat(clouds), vg["glue.points"])
    assert float(vg["glue.points"].min()) == -1.0 and float(vg["glue.points"].max()) == 1.0


def test_transposed_conv_parity_decomposition(vsd):
    """The identity the kernels rely on: ConvTranspose3d(k=4, s=2, p=1) = 8 output-parity classes, each a 2x2x2-tap
    correlation over the INPUT grid (per axis: parity 0 -> (shift 0, k 1), (shift -1, k 3); parity 1 -> (0, 2), (+1, 0))."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 6, 4, 4, 4, generator=g)
    w = torch.randn(6, 5, 4, 4, 4, generator=g)
    ref = torch.nn.functional.conv_transpose3d(x.double(), w.double(), stride=2, padding=1)
    DI, KI = ((0, -1), (0, 1)), ((1, 3), (2, 0))
    xp = torch.nn.functional.pad(x.double(), (1, 1, 1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for pd in range(2):
        for ph in range(2):
            for pw in range(2):
                acc = torch.zeros(1, 5, 4, 4, 4, dtype=torch.float64)
                for td in range(2):
                    for th in range(2):
                        for tw in range(2):
                            dd, dh, dw = DI[pd][td], DI[ph][th], DI[pw][tw]
                            sl = xp[:, :, 1 + dd:5 + dd, 1 + dh:5 + dh, 1 + dw:5 + dw]
                            wk = w[:, :, KI[pd][td], KI[ph][th], KI[pw][tw]].double()
                            acc += torch.einsum("bidhw,io->bodhw", sl, wk)
                out[:, :, pd::2, ph::2, pw::2] = acc
    assert torch.allclose(out, ref, atol=1e-12)


def test_container_matches_reference_keys_and_loads_strictly(vsd):
    v = pcd_b200.VAE3DLarge()
    spec = O.vae3d_decoder_state_dict_spec(prefix="vae")
    sd = v.state_dict()
    for k, shape, _ in spec:
        assert tuple(sd[k[4:]].shape) == tuple(shape), k
    res = v.load_state_dict({k[4:]: t for k, t in vsd.items()}, strict=False)
    assert not res.unexpected_keys and all(k.startswith(("encoder", "fc_")) for k in res.missing_keys)
    m = pcd_b200.LatentDiffusion(v)                       # the reference's default: is_voxel_based=True
    assert m.hparams.is_voxel_based and any(k.startswith("vae.decoder.11.bn2") for k in m.state_dict())
    with pytest.raises(pcd_b200.PcdError):                # no CPU fallback
        v.decode(torch.zeros(1, 256))
    with pytest.raises(pcd_b200.PcdError):
        pcd_b200.voxel_tensor_to_point_clouds(torch.zeros(1, 1, 4, 4, 4))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_oracle_and_container_vs_reference_module(vsd):
    import importlib
    _, rn, _ = ref_shim.load_reference()
    ru = importlib.import_module("utils")
    ref = rn.VAE3DLarge()
    ours = pcd_b200.VAE3DLarge()
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
    assert all(a.shape == b.shape for a, b in zip(ref.state_dict().values(), ours.state_dict().values()))
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict({k[4:]: t for k, t in vsd.items()}, strict=False)
    ref.eval()
    g = torch.Generator().manual_seed(9)
    z = torch.randn(2, 256, generator=g)
    with torch.no_grad():
        want = ref.decode(z)
    got = O.vae3d_decode(vsd, z)
    assert torch.equal(want, got)
    a, b = ru.voxel_tensor_to_point_clouds(want, 0.4), O.voxel_tensor_to_point_clouds(got, 0.4)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
