"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference
(/root/reference imported in place through oracle/ref_shim.py).  Run in the build container:

    python tests/golden/make_golden.py

The weights are not stored (86 MB): they are regenerated from seeds by
`oracle.pointdiff_oracle.make_synthetic_checkpoint`, and a checksum of that state_dict is
stored so a drift of the generator is detected.  Everything else (inputs and the reference's
outputs) is stored as small fp32 tensors.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pointdiff_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "pointdiff_golden.pt")


def sd_checksum(sd):
    tot = 0.0
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            tot += float(v.double().abs().sum())
    return tot


def main():
    rd, rn, rm = ref_shim.load_reference()
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(5)
    out = {}
    for tag, alpha in (("a33", 1.0 / 33.0), ("a3300", 1.0 / 3300.0)):
        sd = O.make_synthetic_checkpoint(seed=24, alpha=alpha)
        m = rd.PointCloudDiffusion(num_points=256)
        m.load_state_dict(sd, strict=True)
        m.eval()
        out[f"{tag}.sd_checksum"] = sd_checksum(sd)
        B, N = 2, 256
        x = torch.randn(B, N, 3, generator=g)
        t = torch.tensor([0.3, 0.9])
        with torch.no_grad():
            out[f"{tag}.fwd.x"], out[f"{tag}.fwd.t"] = x, t
            out[f"{tag}.fwd.eps"] = m.model(x, t)
            # ragged N (not a multiple of 128) and B=3
            x2 = torch.randn(3, 200, 3, generator=g)
            t2 = torch.tensor([1.0, 0.5, 0.01])
            out[f"{tag}.fwd2.x"], out[f"{tag}.fwd2.t"] = x2, t2
            out[f"{tag}.fwd2.eps"] = m.model(x2, t2)
            xT = torch.randn(B, N, 3, generator=g)
            S = 8
            with ref_shim.replay_randn([xT]):
                out[f"{tag}.ddim.out"] = m.sample(B, N, num_steps=S)
            out[f"{tag}.ddim.xT"], out[f"{tag}.ddim.S"] = xT, S
            noises = [torch.randn(B, N, 3, generator=g) for _ in range(S - 1)]
            with ref_shim.replay_randn([xT] + noises):
                out[f"{tag}.ddpm.out"] = m.sample2(B, N, num_steps=S)
            out[f"{tag}.ddpm.noise"] = torch.stack(noises)
            x0 = 0.4 * torch.randn(B, N, 3, generator=g)
            st = torch.full((B,), 0.01)
            out[f"{tag}.ddim3.x"], out[f"{tag}.ddim3.start_t"] = x0, st
            out[f"{tag}.ddim3.out"] = m.sample3(B, N, x=x0, start_t=st, num_steps=5)
    # schedule known answers (diffusion.py:208-223)
    m = rd.PointCloudDiffusion(num_points=16)
    tt = torch.linspace(0, 1, 11)
    n, s = m.diffusion_schedule(tt)
    out["sched.t"], out["sched.noise"], out["sched.signal"] = tt, n, s
    # Chamfer: the reference's own unit-test inputs (units.py:7-11) and a few more
    torch.manual_seed(0)
    X, Y = torch.randn(1, 994, 3), torch.randn(1, 948, 3)
    out["cd.units.x"], out["cd.units.y"] = X, Y
    out["cd.units.value"] = rm.chamfer_distance(X, Y)
    xb = torch.randn(4, 512, 3, generator=g) * torch.tensor([1.0, 0.5, 0.25])
    yb = torch.randn(4, 300, 3, generator=g) * torch.tensor([0.3, 1.0, 0.6]) + 0.2
    out["cd.batch.x"], out["cd.batch.y"] = xb, yb
    out["cd.batch.value"] = rm.chamfer_distance(xb, yb)
    out["cd.batch.per_pair"] = torch.stack([rm.chamfer_distance(xb[i], yb[i]) for i in range(4)])
    out["cd.norm.x"] = rm.normalize_to_cube(xb)
    # Sinkhorn EMD (metrics.py:94-158): the reference's own unit-test inputs (units.py:7-11,25), a ragged batch
    # (ONE cost maximum over the batch) and a larger-epsilon case that needs more iterations
    out["emd.units.value"] = rm.earth_mover_distance_gpu(X, Y)
    out["emd.batch.value"] = rm.earth_mover_distance_gpu(xb, yb)
    out["emd.batch.eps05.value"] = rm.earth_mover_distance_gpu(xb, yb, epsilon=0.5)
    out["emd.batch.per_pair_alone"] = torch.stack([rm.earth_mover_distance_gpu(xb[i], yb[i]) for i in range(4)])
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(out), "entries")


if __name__ == "__main__":
    main()
