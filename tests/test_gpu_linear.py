"""GPU: the tcgen05 per-point layer kernel (pcd_linear_bf16) in isolation against a torch
reference computed from the SAME bf16 operands with fp32 accumulation.  Tolerance: the only
differences are fp32 summation order and the final bf16 rounding, so |err| <= 1 bf16 ulp of the
result (2^-8 relative) + tiny absolute slack."""
import pytest
import torch

import pcd_b200

pytestmark = pytest.mark.gpu


def _ref(a0, w, bias, a1=None, relu=True):
    a = a0.float() if a1 is None else torch.cat([a0.float(), a1.float()], dim=1)
    y = a @ w.float().t() + bias
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("M,K0,K1,Cout,relu", [
    (128, 64, 0, 64, True),        # one tile, one k-block
    (256, 64, 0, 128, True),
    (384, 128, 0, 256, False),
    (1024, 256, 0, 512, True),     # 2 n-blocks
    (512, 512, 512, 512, True),    # decoder concat [prev | skip]
    (256, 128, 128, 128, True),
    (19 * 128, 1024, 0, 2048, True),   # many tiles per CTA -> pipeline / phase wrap-around
    (148 * 2 * 128 + 128, 64, 0, 64, True),   # more tiles than SMs, ragged last wave
    (256, 2048, 0, 4096, True),    # long K, 16 n-blocks
])
def test_linear_matches_torch(M, K0, K1, Cout, relu):
    g = torch.Generator(device="cuda").manual_seed(M + K0 + Cout)
    a0 = torch.randn(M, K0, device="cuda", generator=g).bfloat16()
    a1 = torch.randn(M, K1, device="cuda", generator=g).bfloat16() if K1 else None
    w = (torch.randn(Cout, K0 + K1, device="cuda", generator=g) / (K0 + K1) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g)
    out = pcd_b200._lib.linear_bf16(a0, w, bias, a1, relu)
    torch.cuda.synchronize()
    ref = _ref(a0, w, bias, a1, relu)
    err = (out.float() - ref).abs()
    tol = ref.abs() * 2 ** -7 + 2e-2
    bad = (err > tol)
    assert not bool(bad.any()), f"{int(bad.sum())} / {bad.numel()} elements off; max err {float(err.max())}"
    # and tight in aggregate
    assert float((out.float() - ref).norm() / ref.norm()) < 4e-3


def test_linear_rejects_bad_shapes():
    a = torch.zeros(100, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(pcd_b200.PcdError):
        pcd_b200._lib.linear_bf16(a, w, torch.zeros(64, device="cuda"))
