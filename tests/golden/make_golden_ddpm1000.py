"""Golden vector for the FULL-LENGTH DDPM loop (BASELINE's metric names DDPM-1000) from the UNMODIFIED reference imported in place
(oracle/ref_shim.py): `PointCloudDiffusion.sample2(2, 64, num_steps=1000)` on the alpha = 1/3300 checkpoint (the one that keeps 1000 DDPM steps bounded; with alpha = 1/33 the REFERENCE itself overflows to NaN) with x_T and all 999 noise
draws replayed from CPU-generator seeds (only the seeds are stored; the test redraws the same streams).  64 points keep the
reference's 1000 forwards to about a minute of CPU; the loop arithmetic (schedule table, posterior update, noise injection, 1000
dependent steps) is what this pins.

    python tests/golden/make_golden_ddpm1000.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pointdiff_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ddpm1000_golden.pt")
XT_SEED, NOISE_SEED = 71, 72


def main():
    rd, rn, _ = ref_shim.load_reference()
    N, B, S = 64, 2, 1000
    sd = O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 3300.0)
    m = rd.PointCloudDiffusion(num_points=N)
    m.load_state_dict(sd, strict=True)
    m.eval()
    xT = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(XT_SEED))
    gn = torch.Generator().manual_seed(NOISE_SEED)
    noises = [torch.randn(B, N, 3, generator=gn) for _ in range(S - 1)]
    with torch.no_grad(), ref_shim.replay_randn([xT] + noises):
        out = m.sample2(B, N, num_steps=S)
    rec = {"xT_seed": XT_SEED, "noise_seed": NOISE_SEED, "S": S, "N": N, "B": B, "out": out,
           "sd_checksum": sum(float(v.double().abs().sum()) for v in sd.values())}
    del m
    # latent model (BASELINE config 4's loop at DDPM length): LatentDiffusion.sample2(2, num_steps=1000) with a 256-point
    # SimplePointNetVAE (is_voxel_based=False), z_T and the 999 noise draws replayed from the same kind of seeded streams
    NP = 256
    sdl = O.make_synthetic_latent_checkpoint(num_points=NP)
    lm = rd.LatentDiffusion(rn.SimplePointNetVAE(num_points=NP), is_voxel_based=False)
    assert not lm.load_state_dict(sdl, strict=False).unexpected_keys
    lm.eval()
    zT = torch.randn(B, 256, generator=torch.Generator().manual_seed(XT_SEED + 10))
    gl = torch.Generator().manual_seed(NOISE_SEED + 10)
    lnoises = [torch.randn(B, 256, generator=gl) for _ in range(S - 1)]
    with torch.no_grad(), ref_shim.replay_randn([zT] + lnoises):
        lout = lm.sample2(B, num_steps=S)
    rec.update({"latent.zT_seed": XT_SEED + 10, "latent.noise_seed": NOISE_SEED + 10, "latent.num_points": NP, "latent.out": lout,
                "latent.sd_checksum": sum(float(v.double().abs().sum()) for v in sdl.values())})
    torch.save(rec, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; |out| mean", float(out.abs().mean()), "finite", bool(torch.isfinite(out).all()),
          "; latent |out| mean", float(lout.abs().mean()), "finite", bool(torch.isfinite(lout).all()))


if __name__ == "__main__":
    main()
