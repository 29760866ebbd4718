"""CPU, build container only: the oracle against the reference's own code imported in place.
Skipped where /root/reference is not mounted (e.g. the GPU box)."""
import pytest
import torch

from oracle import pointdiff_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref_model(sd33):
    rd, _, _ = ref_shim.load_reference()
    m = rd.PointCloudDiffusion(num_points=128)
    assert list(m.state_dict().keys()) == list(sd33.keys())
    m.load_state_dict(sd33, strict=True)
    return m.eval()


def test_forward_bit_exact(ref_model, sd33):
    g = torch.Generator().manual_seed(11)
    x, t = torch.randn(2, 128, 3, generator=g), torch.tensor([0.1, 0.77])
    with torch.no_grad():
        assert torch.equal(ref_model.model(x, t), O.denoiser_forward(sd33, x, t))


def test_forward_bit_exact_other_widths():
    """dim == time_dim other than the default 256 (the reference's constructor takes them, diffusion.py:15-28; odd widths zero-pad
    the embedding, networks.py:836-837)."""
    rd, _, _ = ref_shim.load_reference()
    g = torch.Generator().manual_seed(14)
    x, t = torch.randn(2, 96, 3, generator=g), torch.tensor([0.2, 0.9])
    for T in (64, 129, 512):
        sd = O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 33.0, dim=T, time_dim=T)
        m = rd.PointCloudDiffusion(num_points=96, dim=T, time_dim=T)
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            assert torch.equal(m.eval().model(x, t), O.denoiser_forward(sd, x, t, time_dim=T)), T


def test_samplers_bit_exact(ref_model, sd33):
    g = torch.Generator().manual_seed(12)
    xT = torch.randn(2, 128, 3, generator=g)
    noises = [torch.randn(2, 128, 3, generator=g) for _ in range(4)]
    with torch.no_grad(), ref_shim.replay_randn([xT]):
        assert torch.equal(ref_model.sample(2, 128, num_steps=5), O.ddim_sample(sd33, xT, 5))
    with torch.no_grad(), ref_shim.replay_randn([xT] + noises):
        assert torch.equal(ref_model.sample2(2, 128, num_steps=5), O.ddpm_sample(sd33, xT, noises, 5))


def test_reference_units_py_inputs():
    _, _, rm = ref_shim.load_reference()
    torch.manual_seed(0)
    x, y = torch.randn(1, 994, 3), torch.randn(1, 948, 3)
    assert float(rm.chamfer_distance(x, y)) == float(O.chamfer_distance(x, y))


def test_compute_metrics_recon_term_bit_exact():
    """metrics.py:181 (voxel BCE through utils.voxelize): oracle restatement AND the product's host-side torch expression."""
    import pcd_b200
    _, _, rm = ref_shim.load_reference()
    import utils as ref_utils
    g = torch.Generator().manual_seed(17)
    x, y = torch.rand(3, 500, 3, generator=g) * 2.4 - 1.2, torch.rand(3, 400, 3, generator=g) * 2 - 1
    want = torch.nn.functional.binary_cross_entropy(ref_utils.voxelize(x), ref_utils.voxelize(y))
    assert torch.equal(O.voxel_bce(x, y), want)
    assert torch.equal(pcd_b200.metrics.voxelize(x), ref_utils.voxelize(x))
    assert torch.equal(O.voxelize(x[0]), ref_utils.voxelize(x[0]))


def test_product_state_dict_interoperates_with_reference(ref_model):
    import pcd_b200
    mine = pcd_b200.PointCloudDiffusion(128)
    mine.load_state_dict(ref_model.state_dict(), strict=True)
    ref_model.load_state_dict(mine.state_dict(), strict=True)
    assert [tuple(v.shape) for v in mine.state_dict().values()] == [tuple(v.shape) for v in ref_model.state_dict().values()]
    # same schedule bits
    t = torch.linspace(0, 1, 7)
    for a, b in zip(mine.diffusion_schedule(t), ref_model.diffusion_schedule(t)):
        assert torch.equal(a, b)


def test_sinkhorn_emd_bit_exact():
    _, _, rm = ref_shim.load_reference()
    g = torch.Generator().manual_seed(13)
    x, y = torch.randn(2, 200, 3, generator=g), torch.randn(2, 150, 3, generator=g) * 0.5
    assert float(rm.earth_mover_distance_gpu(x, y)) == float(O.sinkhorn_emd(x, y))
    assert float(rm.earth_mover_distance_gpu(x, y, epsilon=0.3, max_iter=7)) == float(O.sinkhorn_emd(x, y, epsilon=0.3, max_iter=7))


def test_linear_schedule_samplers_bit_exact(sd33):
    """noise_schedule='linear' (diffusion.py:189-205): the oracle reproduces the batch-axis cumprod exactly."""
    rd, _, _ = ref_shim.load_reference()
    m = rd.PointCloudDiffusion(num_points=64, noise_schedule="linear")
    m.load_state_dict(sd33, strict=True)
    m.eval()
    g = torch.Generator().manual_seed(14)
    xT = torch.randn(3, 64, 3, generator=g)
    noises = [torch.randn(3, 64, 3, generator=g) for _ in range(3)]
    with torch.no_grad(), ref_shim.replay_randn([xT]):
        assert torch.equal(m.sample(3, 64, num_steps=4), O.ddim_sample(sd33, xT, 4, schedule="linear"))
    with torch.no_grad(), ref_shim.replay_randn([xT] + noises):
        assert torch.equal(m.sample2(3, 64, num_steps=4), O.ddpm_sample(sd33, xT, noises, 4, schedule="linear"))
    x0 = 0.3 * torch.randn(3, 64, 3, generator=g)
    with torch.no_grad():
        assert torch.equal(m.sample3(3, 64, x=x0, start_t=torch.full((3,), 0.2), num_steps=3),
                           O.ddim3_sample(sd33, x0, torch.full((3,), 0.2), 3, schedule="linear"))
