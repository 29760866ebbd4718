"""GPU: the reverse loops (DDIM `sample`, DDPM `sample2`, DDIM-from-t `sample3`) through the C ABI
against the reference golden outputs / the oracle, with shared x_T and injected noise.

Tolerances:
  fp32 mode: relative L2 of the final sample < 2e-4 (the loop amplifies per-step 1e-6 noise:
             total linear gain of the DDIM map is s(0)/s(1) = 47.5)
  bf16 mode: Chamfer distance (reference units, x1e3, cube-normalised) between our sample and the
             reference's sample < 10  (unrelated clouds are ~300 apart; SURVEY H2 measured 8.8 for
             a torch bf16 emulation of DDIM-50).
"""
import pytest
import torch

import pcd_b200
from oracle import pointdiff_oracle as O
from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _model(sd, precision, n=256):
    m = pcd_b200.PointCloudDiffusion(n, precision=precision)
    m.load_state_dict(sd, strict=True)
    return m.eval().cuda()


def _cd(a, b):
    return float(O.chamfer_pairs(a.cpu(), b.cpu())[0].max())


@pytest.mark.parametrize("tag", ["a33", "a3300"])
def test_fp32_samplers_vs_reference_golden(golden, sd33, sd3300, tag):
    sd = sd33 if tag == "a33" else sd3300
    m = _model(sd, "fp32")
    S, xT = int(golden[f"{tag}.ddim.S"]), golden[f"{tag}.ddim.xT"]
    out = m.sample(2, 256, num_steps=S, x_T=xT)
    assert out.is_cuda and out.shape == (2, 256, 3)
    assert rel_l2(out, golden[f"{tag}.ddim.out"]) < 2e-4
    out = m.sample2(2, 256, num_steps=S, x_T=xT, noise=golden[f"{tag}.ddpm.noise"])
    assert rel_l2(out, golden[f"{tag}.ddpm.out"]) < 2e-4
    out = m.sample3(2, 256, x=golden[f"{tag}.ddim3.x"], start_t=golden[f"{tag}.ddim3.start_t"], num_steps=5)
    assert rel_l2(out, golden[f"{tag}.ddim3.out"]) < 2e-4


def test_bf16_samplers_vs_reference_golden(golden, sd3300):
    m = _model(sd3300, "bf16")
    S, xT = int(golden["a3300.ddim.S"]), golden["a3300.ddim.xT"]
    out = m.sample(2, 256, num_steps=S, x_T=xT)
    assert torch.isfinite(out).all()
    assert _cd(out, golden["a3300.ddim.out"]) < 10.0
    assert rel_l2(out, golden["a3300.ddim.out"]) < 5e-2
    out = m.sample2(2, 256, num_steps=S, x_T=xT, noise=golden["a3300.ddpm.noise"])
    assert _cd(out, golden["a3300.ddpm.out"]) < 10.0
    out = m.sample3(2, 256, x=golden["a3300.ddim3.x"], start_t=golden["a3300.ddim3.start_t"], num_steps=5)
    assert _cd(out, golden["a3300.ddim3.out"]) < 10.0


def test_bf16x3_samplers_meet_the_fp32_style_bound(golden, sd3300):
    """Split-bf16 mode: final samples within 2e-3 relative L2 and Chamfer < 0.5 of the reference's."""
    m = _model(sd3300, "bf16x3")
    S, xT = int(golden["a3300.ddim.S"]), golden["a3300.ddim.xT"]
    out = m.sample(2, 256, num_steps=S, x_T=xT)
    assert rel_l2(out, golden["a3300.ddim.out"]) < 2e-3 and _cd(out, golden["a3300.ddim.out"]) < 0.5
    out = m.sample2(2, 256, num_steps=S, x_T=xT, noise=golden["a3300.ddpm.noise"])
    assert rel_l2(out, golden["a3300.ddpm.out"]) < 2e-3 and _cd(out, golden["a3300.ddpm.out"]) < 0.5


def test_fp32_ddim50_vs_oracle(sd33):
    g = torch.Generator().manual_seed(31)
    xT = torch.randn(1, 128, 3, generator=g)
    m = _model(sd33, "fp32", 128)
    out = m.sample(1, 128, num_steps=50, x_T=xT)
    ref = O.ddim_sample(sd33, xT, 50)
    assert rel_l2(out, ref) < 1e-3


def test_graph_and_eager_loops_agree(sd3300, monkeypatch):
    g = torch.Generator().manual_seed(32)
    xT = torch.randn(2, 128, 3, generator=g)
    m = _model(sd3300, "bf16", 128)
    a = m.sample(2, 128, num_steps=6, x_T=xT)
    monkeypatch.setenv("PCD_NO_GRAPH", "1")
    b = m.sample(2, 128, num_steps=6, x_T=xT)
    assert torch.equal(a, b)


def test_philox_noise_is_standard_normal_and_shard_invariant():
    z = pcd_b200._lib.philox_normal(seed=5, sample_offset=0, step=3, B=8, N=2048, device="cuda")
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    assert abs(float((z ** 4).mean()) - 3.0) < 0.15                      # kurtosis of N(0,1)
    part = pcd_b200._lib.philox_normal(seed=5, sample_offset=4, step=3, B=4, N=2048, device="cuda")
    assert torch.equal(part, z[4:])                                     # keyed by GLOBAL sample index
    other = pcd_b200._lib.philox_normal(seed=5, sample_offset=0, step=4, B=8, N=2048, device="cuda")
    assert not torch.equal(other, z)
    assert abs(float((z * other).mean())) < 0.02                        # steps are uncorrelated


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ddpm_philox_path_matches_oracle_and_shards(sd3300, precision):
    """DDPM with in-kernel Philox noise: (a) equals the oracle fed with the very same noise,
    (b) sampling shapes [0,4) in one call equals [0,2) + [2,4) in two calls (multi-GPU sharding)."""
    g = torch.Generator().manual_seed(33)
    B, N, S, seed = 4, 128, 6, 77
    xT = torch.randn(B, N, 3, generator=g)
    m = _model(sd3300, precision, N)
    full = m.sample2(B, N, num_steps=S, x_T=xT, seed=seed)
    parts = torch.cat([m.sample2(2, N, num_steps=S, x_T=xT[:2], seed=seed, sample_offset=0),
                       m.sample2(2, N, num_steps=S, x_T=xT[2:], seed=seed, sample_offset=2)])
    assert torch.equal(full, parts)
    noises = [pcd_b200._lib.philox_normal(seed, 0, k, B, N, "cuda").cpu() for k in range(S - 1)]
    ref = O.ddpm_sample(sd3300, xT, noises, S)
    if precision == "fp32":
        assert rel_l2(full, ref) < 2e-4
    else:
        assert _cd(full, ref) < 10.0


def test_host_buffer_entry_point(sd3300):
    g = torch.Generator().manual_seed(34)
    xT = torch.randn(2, 128, 3, generator=g)
    m = _model(sd3300, "bf16", 128)
    dev = m.sample(2, 128, num_steps=4, x_T=xT)
    host = m.sample_host(xT.pin_memory(), 4, "ddim")
    assert not host.is_cuda and torch.equal(host, dev.cpu())
    # DDIM from a given x / start_t (`sample3`, diffusion.py:291-337) through the same host-buffer entry
    x0, st = 0.3 * torch.randn(2, 128, 3, generator=g), torch.full((2,), 0.2)
    assert torch.equal(m.sample_host(x0, 3, "ddim3", start_t=st), m.sample3(2, 128, x=x0, start_t=st, num_steps=3).cpu())
    # noise_schedule='linear' (one schedule row per sample: pcd_sample_host_rows)
    ml = pcd_b200.PointCloudDiffusion(128, noise_schedule="linear", precision="fp32")
    ml.load_state_dict(sd3300, strict=True)
    ml = ml.eval().cuda()
    got = ml.sample_host(xT, 5, "ddim")
    assert torch.equal(got, ml.sample(2, 128, num_steps=5, x_T=xT).cpu())
    assert rel_l2(got, O.ddim_sample(sd3300, xT, 5, schedule="linear")) < 2e-4


@pytest.mark.parametrize("precision", ["fp32", "f16mix", "bf16"])
def test_reconstruction_script_path(sd3300, precision):
    """The reference's reconstruction test (test_point_ddpm.py:58-122): noise validation clouds to t = 0.01
    (`add_noise`), run `sample3` from there, score every pair with `compute_metrics`.  Same calls, same order;
    the randn_like draw of add_noise is injected so the oracle sees identical inputs."""
    g = torch.Generator().manual_seed(35)
    B, N, S = 3, 256, 6
    x0 = 0.5 * torch.randn(B, N, 3, generator=g) * torch.tensor([1.0, 0.6, 0.3])
    t = torch.ones(B) * 0.010
    noise = torch.randn(B, N, 3, generator=g)
    x_t, _, n, s = O.add_noise(x0, t, noise)
    m = _model(sd3300, precision, N)
    n_dev, s_dev = m.diffusion_schedule(t.cuda())
    assert torch.equal(n_dev.cpu(), n) or torch.allclose(n_dev.cpu(), n, rtol=1e-6)
    recon = m.sample3(num_samples=B, num_points=N, x=x_t.cuda(), start_t=t.cuda(), num_steps=S)
    ref = O.ddim3_sample(sd3300, x_t, t, S)
    if precision == "fp32":
        assert rel_l2(recon, ref) < 2e-4
    elif precision == "f16mix":
        assert rel_l2(recon, ref) < 1e-3
    cds, emds = [], []
    for orig, rec in zip(x0.cuda(), recon):
        cd, emd, recon_loss = pcd_b200.compute_metrics(orig, rec, use_approximate_gpu_emd=True)
        assert cd.dim() == 0 and emd.dim() == 0 and recon_loss.dim() == 0
        assert float(recon_loss) == float(O.voxel_bce(orig.cpu(), rec.cpu()))      # metrics.py:181, same occupancy grids
        cds.append(float(cd)); emds.append(float(emd))
    want_cd = O.chamfer_pairs(x0, ref)[0]
    want_emd = torch.stack([O.sinkhorn_emd(x0[i], ref[i], exact=True) for i in range(B)])
    tol = {"fp32": 1e-4, "f16mix": 2e-3, "bf16": 5e-2}[precision]
    assert torch.allclose(torch.tensor(cds), want_cd, rtol=tol)
    assert torch.allclose(torch.tensor(emds), want_emd, rtol=tol)


@pytest.mark.parametrize("precision,bound", [("f16mix", 2e-4), ("f16", 2e-3)])
def test_fp16_modes_samplers_vs_reference_golden(golden, sd3300, precision, bound):
    """fp16 operand modes: final samples of all three loops against the reference's golden outputs."""
    m = _model(sd3300, precision)
    S, xT = int(golden["a3300.ddim.S"]), golden["a3300.ddim.xT"]
    assert rel_l2(m.sample(2, 256, num_steps=S, x_T=xT), golden["a3300.ddim.out"]) < bound
    assert rel_l2(m.sample2(2, 256, num_steps=S, x_T=xT, noise=golden["a3300.ddpm.noise"]), golden["a3300.ddpm.out"]) < bound
    out = m.sample3(2, 256, x=golden["a3300.ddim3.x"], start_t=golden["a3300.ddim3.start_t"], num_steps=5)
    assert rel_l2(out, golden["a3300.ddim3.out"]) < bound


@pytest.mark.parametrize("precision,bound", [("fp32", 2e-4), ("f16mix", 2e-3)])
def test_linear_schedule_loops_vs_oracle(sd3300, precision, bound):
    """noise_schedule='linear': per-sample schedule rows (the reference's batch-axis cumprod, diffusion.py:202)
    through pcd_sample_rows, all three loops against the oracle (itself bit-identical to the reference)."""
    g = torch.Generator().manual_seed(36)
    B, N, S = 3, 128, 5
    m = pcd_b200.PointCloudDiffusion(N, noise_schedule="linear", precision=precision)
    m.load_state_dict(sd3300, strict=True)
    m = m.eval().cuda()
    xT = torch.randn(B, N, 3, generator=g)
    noises = [torch.randn(B, N, 3, generator=g) for _ in range(S - 1)]
    assert rel_l2(m.sample(B, N, num_steps=S, x_T=xT), O.ddim_sample(sd3300, xT, S, schedule="linear")) < bound
    out = m.sample2(B, N, num_steps=S, x_T=xT, noise=torch.stack(noises))
    assert rel_l2(out, O.ddpm_sample(sd3300, xT, noises, S, schedule="linear")) < bound
    x0 = 0.3 * torch.randn(B, N, 3, generator=g)
    st = torch.full((B,), 0.2)
    assert rel_l2(m.sample3(B, N, x=x0, start_t=st, num_steps=3), O.ddim3_sample(sd3300, x0, st, 3, schedule="linear")) < bound
    # the quirk is real: sample 0 alone is NOT sample 0 of the batch (its rates depend on its index in the batch)
    alone = m.sample(1, N, num_steps=S, x_T=xT[:1])
    assert rel_l2(alone, O.ddim_sample(sd3300, xT[:1], S, schedule="linear")) < bound


def test_full_size_ddim50_properties(sd3300):
    """BASELINE configs[1] at its full size (batch 512 x 2048 points, DDIM-50, bf16), where the CPU oracle would take hours:
    size-independent properties instead.  (1) every sample is finite; (2) sharding: the first 256 samples equal a batch-256 call on
    the same x_T bit for bit (no cross-sample arithmetic, fixed tile order per sample); (3) point-permutation equivariance of the
    whole 50-step loop, bit for bit (per-point layers are row independent, the max-pool is order independent, DDIM is noise free);
    (4) the Chamfer distance of a cloud to its own permutation is 0 and the all-pairs matrix of the 512 samples is symmetric."""
    B, N, S = 512, 2048, 50
    m = _model(sd3300, "bf16", N)
    g = torch.Generator().manual_seed(99)
    xT = torch.randn(B, N, 3, generator=g)
    full = m.sample(B, N, num_steps=S, x_T=xT)
    assert full.shape == (B, N, 3) and bool(torch.isfinite(full).all())
    half = m.sample(256, N, num_steps=S, x_T=xT[:256])
    assert torch.equal(half, full[:256])
    perm = torch.randperm(N, generator=g)
    sub = m.sample(8, N, num_steps=S, x_T=xT[:8][:, perm].contiguous())
    assert torch.equal(sub, full[:8][:, perm.cuda()])
    cd_self = pcd_b200.chamfer_distance_per_pair(full[:8], sub)
    assert float(cd_self.abs().max()) == 0.0
    cdm = pcd_b200.chamfer_matrix(full[:128].contiguous(), full[:128].contiguous())
    assert torch.equal(cdm, cdm.t()) and float(cdm.diagonal().abs().max()) == 0.0


def test_edge_shapes_vs_oracle(sd33):
    """Smallest and empty shapes: one cloud of ONE point through a one-step loop of each sampler (a single padded row tile, the
    max-pool over one valid point), and an empty batch (the reference's loops return the empty tensor)."""
    g = torch.Generator().manual_seed(40)
    m = _model(sd33, "fp32", 1)
    xT = torch.randn(1, 1, 3, generator=g)
    assert rel_l2(m.sample(1, 1, num_steps=1, x_T=xT), O.ddim_sample(sd33, xT, 1)) < 1e-4
    assert rel_l2(m.sample2(1, 1, num_steps=1, x_T=xT), O.ddpm_sample(sd33, xT, [], 1)) < 1e-4
    st = torch.full((1,), 0.3)
    assert rel_l2(m.sample3(1, 1, x=xT, start_t=st, num_steps=1), O.ddim3_sample(sd33, xT, st, 1)) < 1e-4
    n2 = [torch.randn(1, 1, 3, generator=g)]
    assert rel_l2(m.sample2(1, 1, num_steps=2, x_T=xT, noise=torch.stack(n2)), O.ddpm_sample(sd33, xT, n2, 2)) < 1e-4
    mm = _model(sd33, "f16mix", 64)
    empty = mm.sample(0, 64, num_steps=3, x_T=torch.empty(0, 64, 3))
    assert tuple(empty.shape) == (0, 64, 3) and empty.is_cuda
    assert tuple(mm.model(torch.empty(0, 64, 3).cuda(), torch.empty(0).cuda()).shape) == (0, 64, 3)
    assert torch.isnan(pcd_b200.chamfer_distance(torch.empty(0, 5, 3).cuda(), torch.empty(0, 7, 3).cuda()))
    assert torch.isnan(O.chamfer_distance(torch.empty(0, 5, 3), torch.empty(0, 7, 3)))


def test_batch_above_bench_size_matches_small_batches_f16mix(sd33):
    """Maximum sizes: 1536 clouds x 2048 points (3.1 M rows, three times the bench batch) in the default precision.  Size-independent
    property: the first and the last four clouds equal batch-4 calls on the same x_T bit for bit -- the batch-4 plan runs the fused
    chain kernels and the skinny per-sample bias GEMM, the big plan neither, so this also pins those schedule choices against each
    other at full point count."""
    B, N, S = 1536, 2048, 2
    m = _model(sd33, "f16mix", N)
    g = torch.Generator().manual_seed(123)
    xT = torch.randn(B, N, 3, generator=g)
    full = m.sample(B, N, num_steps=S, x_T=xT)
    assert full.shape == (B, N, 3) and bool(torch.isfinite(full).all())
    assert torch.equal(m.sample(4, N, num_steps=S, x_T=xT[:4]), full[:4])
    assert torch.equal(m.sample(4, N, num_steps=S, x_T=xT[-4:]), full[-4:])
    # and against the oracle on one cloud (two steps of the alpha = 1/33 checkpoint)
    assert rel_l2(full[-1:], O.ddim_sample(sd33, xT[-1:], S)) < 2e-3
