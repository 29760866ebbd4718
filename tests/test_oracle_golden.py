"""CPU: the oracle restatement against the golden vectors generated from the reference
(tests/golden/make_golden.py) -- this is what pins the oracle (the reference ships
no golden tensors of its own for this path; units.py only asserts ranges)."""
import torch

from oracle import pointdiff_oracle as O


def _checksum(sd):
    return sum(float(v.double().abs().sum()) for v in sd.values() if v.dtype.is_floating_point)


def test_synthetic_checkpoint_is_reproducible(golden, sd33, sd3300):
    assert abs(_checksum(sd33) - golden["a33.sd_checksum"]) < 1e-6 * golden["a33.sd_checksum"]
    assert abs(_checksum(sd3300) - golden["a3300.sd_checksum"]) < 1e-6 * golden["a3300.sd_checksum"]
    spec = O.state_dict_spec()
    assert len(spec) == 203 and list(sd33.keys()) == [k for k, _, _ in spec]
    n_params = sum(v.numel() for k, v in sd33.items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n_params == 21485827


def test_forward_matches_reference_golden(golden, sd33):
    for tag in ("fwd", "fwd2"):
        eps = O.denoiser_forward(sd33, golden[f"a33.{tag}.x"], golden[f"a33.{tag}.t"])
        assert torch.equal(eps, golden[f"a33.{tag}.eps"])   # same ops, same machine family: bit-exact


def test_samplers_match_reference_golden(golden, sd33, sd3300):
    for tag, sd in (("a33", sd33), ("a3300", sd3300)):
        S = int(golden[f"{tag}.ddim.S"])
        xT = golden[f"{tag}.ddim.xT"]
        assert torch.equal(O.ddim_sample(sd, xT, S), golden[f"{tag}.ddim.out"])
        noise = golden[f"{tag}.ddpm.noise"]
        assert torch.equal(O.ddpm_sample(sd, xT, list(noise), S), golden[f"{tag}.ddpm.out"])
        out3 = O.ddim3_sample(sd, golden[f"{tag}.ddim3.x"], golden[f"{tag}.ddim3.start_t"], 5)
        assert torch.equal(out3, golden[f"{tag}.ddim3.out"])


def test_schedule_known_answers(golden):
    n, s = O.offset_cosine_schedule(golden["sched.t"])
    assert torch.equal(n, golden["sched.noise"]) and torch.equal(s, golden["sched.signal"])
    # t=1 -> (n .9998, s .02); t=0 -> (n .31225, s .95); signal^2 + noise^2 = 1
    assert abs(float(s[-1]) - 0.02) < 1e-6 and abs(float(s[0]) - 0.95) < 1e-6
    assert torch.allclose(n * n + s * s, torch.ones_like(n), atol=1e-6)


def test_linear_schedule_batch_cumprod_quirk():
    # t=[.5,.5,.5] -> signal [0.990, 0.980, 0.970]
    n, s = O.linear_schedule(torch.tensor([0.5, 0.5, 0.5]))
    assert torch.allclose(s, torch.tensor([0.98995, 0.98000, 0.97015]), atol=1e-4)


def test_chamfer_known_answers(golden):
    cd = O.chamfer_distance(golden["cd.units.x"], golden["cd.units.y"])
    assert float(cd) == float(golden["cd.units.value"]) == 142.71389770507812   # derived KAT, SURVEY section 4
    assert 0.0 <= float(cd) <= 200.0                                              # the reference's own assertion (units.py:25-26)
    cdb = O.chamfer_distance(golden["cd.batch.x"], golden["cd.batch.y"])
    assert float(cdb) == float(golden["cd.batch.value"])
    assert torch.equal(O.normalize_to_cube(golden["cd.batch.x"]), golden["cd.norm.x"])
    per_pair, ixy, iyx = O.chamfer_pairs(golden["cd.batch.x"], golden["cd.batch.y"])
    # exact (direct-difference) per-pair values agree with the reference's mm-path values to fp32 noise
    assert torch.allclose(per_pair, golden["cd.batch.per_pair"], rtol=2e-5)
    assert torch.allclose(per_pair.mean(), golden["cd.batch.value"], rtol=2e-5)
    assert ixy.shape == (4, 512) and iyx.shape == (4, 300)


def test_chamfer_properties():
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(3, 100, 3, generator=g), torch.randn(3, 80, 3, generator=g)
    a, _, _ = O.chamfer_pairs(x, y)
    b, _, _ = O.chamfer_pairs(y, x)
    assert torch.allclose(a, b, rtol=1e-6)                              # symmetry
    assert float(O.chamfer_pairs(x, x)[0].abs().max()) == 0.0           # exact path: CD(x,x) = 0
    M = O.chamfer_matrix(x, y)
    assert torch.allclose(torch.diagonal(M), a, rtol=1e-6)
    res = O.set_metrics_from_matrices(O.chamfer_matrix(x, y), O.chamfer_matrix(x, x), O.chamfer_matrix(y, y))
    assert 0.0 <= res["cov_cd"] <= 1.0 and 0.0 <= res["1nna_cd"] <= 1.0 and res["mmd_cd"] > 0


def test_denoiser_point_permutation_equivariance(sd33):
    g = torch.Generator().manual_seed(2)
    x, t = torch.randn(1, 64, 3, generator=g), torch.tensor([0.5])
    perm = torch.randperm(64, generator=g)
    a = O.denoiser_forward(sd33, x, t)[:, perm]
    b = O.denoiser_forward(sd33, x[:, perm], t)
    assert torch.allclose(a, b, atol=1e-5)


def test_sinkhorn_emd_known_answers(golden):
    """metrics.py:94-158.  The oracle restates the reference's update order exactly: bit-identical values on the
    reference's own unit-test inputs (units.py:7-11,25; the reference asserts 0 <= EMD <= 200) and on a ragged
    batch whose cost is normalised by ONE maximum over the batch."""
    X, Y = golden["cd.units.x"], golden["cd.units.y"]
    emd = O.sinkhorn_emd(X, Y)
    assert float(emd) == float(golden["emd.units.value"]) == 5.995205879211426     # derived KAT, SURVEY section 4
    assert 0.0 <= float(emd) <= 200.0
    xb, yb = golden["cd.batch.x"], golden["cd.batch.y"]
    assert float(O.sinkhorn_emd(xb, yb)) == float(golden["emd.batch.value"])
    assert float(O.sinkhorn_emd(xb, yb, epsilon=0.5)) == float(golden["emd.batch.eps05.value"])
    # direct-difference distances (what the CUDA kernel computes) change the value only at fp32 noise level
    exact, iters = O.sinkhorn_emd(xb, yb, exact=True, per_pair=True)
    assert abs(float(exact.mean()) - float(golden["emd.batch.value"])) < 2e-5 * float(golden["emd.batch.value"])
    assert 1 <= iters <= 100
    # the batch maximum couples the pairs: a pair evaluated alone differs from the same pair inside the batch
    alone = torch.stack([O.sinkhorn_emd(xb[i], yb[i]) for i in range(4)])
    assert torch.equal(alone, golden["emd.batch.per_pair_alone"])
    assert not torch.allclose(alone.mean(), golden["emd.batch.value"], rtol=1e-3)


def test_oracle_ddpm1000_is_bit_identical_to_the_reference_golden(sd3300):
    """`PointCloudDiffusion.sample2` at the metric's full length (1000 reverse steps, 2 x 64 points, x_T and 999 noise draws replayed):
    the oracle reproduces the unmodified reference's output bit for bit (golden: tests/golden/make_golden_ddpm1000.py; ~30 s)."""
    import os
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ddpm1000_golden.pt"), weights_only=True)
    B, S, N = int(g["B"]), int(g["S"]), int(g["N"])
    xT = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(int(g["xT_seed"])))
    gn = torch.Generator().manual_seed(int(g["noise_seed"]))
    noises = [torch.randn(B, N, 3, generator=gn) for _ in range(S - 1)]
    assert torch.equal(O.ddpm_sample(sd3300, xT, noises, S), g["out"])
