"""Golden vectors for the latent path (BASELINE config 4) from the UNMODIFIED reference
(LatentDiffusion + SimpleLatentUNetPointNet + SimplePointNetVAE, imported in place).
    python tests/golden/make_golden_latent.py
Weights are regenerated from seeds by oracle.make_synthetic_latent_checkpoint (checksum stored)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pointdiff_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "latent_golden.pt")
FOLD_NP = 300   # points of the synthetic FoldingDecoder
NP = 256   # points of the synthetic SimplePointNetVAE (keeps the 3*NP x 3*NP output layer small)


def main():
    rd, rn, _ = ref_shim.load_reference()
    sd = O.make_synthetic_latent_checkpoint(num_points=NP)
    m = rd.LatentDiffusion(rn.SimplePointNetVAE(num_points=NP), is_voxel_based=False)
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.startswith(("vae.encoder", "vae.fc_")) for k in res.missing_keys)
    m.eval()
    g = torch.Generator().manual_seed(5)
    out = {"num_points": NP, "sd_checksum": sum(float(v.double().abs().sum()) for v in sd.values())}
    with torch.no_grad():
        z, t = torch.randn(5, 256, generator=g), torch.tensor([0.05, 0.3, 0.6, 0.9, 1.0])
        out["fwd.z"], out["fwd.t"], out["fwd.eps"] = z, t, m.model(z, t)
        out["decode.z"] = z
        out["decode.out"] = m.vae.decode(z)
        zT = torch.randn(4, 256, generator=g)
        S = 8
        noises = [torch.randn(4, 256, generator=g) for _ in range(S - 1)]
        with ref_shim.replay_randn([zT] + noises):
            out["ddpm.out"] = m.sample2(4, num_steps=S)          # decoded clouds [4, NP, 3]
        out["ddpm.zT"], out["ddpm.noise"], out["ddpm.S"] = zT, torch.stack(noises), S
        # the reference's sample()/sample3() crash for a point VAE (diffusion.py:650-653, 704-707): record the
        # latent loop result through the reference's own pieces instead (model + schedule + remove_noise)
        z_t, z_0 = zT, zT
        for step in range(S):
            tt = torch.ones(4) - step * (1.0 / S)
            n, s = m.diffusion_schedule(tt)
            eps = m.model(z_t, tt)
            z_0 = m.remove_noise(z_t, eps, n, s)
            n2, s2 = m.diffusion_schedule(tt - 1.0 / S)
            z_t = s2.view(-1, 1) * z_0 + n2.view(-1, 1) * eps
        out["ddim.z0"] = z_0
        out["ddim.out"] = m.vae.decode(z_0)
        # FoldingDecoder (PointNetVAE.decode, networks.py:1449-1509) with seeded weights, ragged num_points (not 2^k)
        fsd = O.make_synthetic_folding_checkpoint(num_points=FOLD_NP)
        dec = rn.FoldingDecoder(256, FOLD_NP)
        dec.load_state_dict({k[len("vae.decoder."):]: v for k, v in fsd.items()}, strict=True)
        zf = torch.randn(3, 256, generator=g)
        out["fold.num_points"], out["fold.z"], out["fold.out"] = FOLD_NP, zf, dec(zf)
        out["fold.sd_checksum"] = sum(float(v.double().abs().sum()) for v in fsd.values())
        out["fold.grid"] = dec.grid
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
