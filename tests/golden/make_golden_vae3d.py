"""Golden vectors for the voxel-VAE decoder from the UNMODIFIED reference
(networks.VAE3DLarge.decode + utils.voxel_tensor_to_point_clouds, imported in place).
    python tests/golden/make_golden_vae3d.py
Weights are regenerated from seeds by oracle.make_synthetic_vae3d_decoder_checkpoint (checksum stored)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pointdiff_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "vae3d_golden.pt")


def main():
    _, rn, _ = ref_shim.load_reference()
    ru = importlib.import_module("utils")
    sd = O.make_synthetic_vae3d_decoder_checkpoint()
    vae = rn.VAE3DLarge()
    res = vae.load_state_dict({k[len("vae."):]: v for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys and all(k.startswith(("encoder", "fc_")) for k in res.missing_keys)
    vae.eval()
    g = torch.Generator().manual_seed(3)
    z = torch.randn(3, 256, generator=g)           # odd batch: exercises the even-batch padding of the 4^3 tiles
    out = {"sd_checksum": sum(float(v.double().abs().sum()) for v in sd.values()), "z": z, "threshold": 0.4}
    with torch.no_grad():
        vox = vae.decode(z)
    out["vox"] = vox.half()                        # 197 KB; fp16 rounding (<= 4.9e-4 abs) is far below what the fixture checks:
    out["vox_sum"] = vox.double().sum()            # the exact fp32 grid is pinned through its sum and the point clouds
    clouds = ru.voxel_tensor_to_point_clouds(vox, threshold=0.4)
    out["counts"] = torch.tensor([len(c) for c in clouds])
    out["points"] = torch.cat(clouds)
    # the glue on a small non-cubic grid with an empty and a full sample
    v2 = torch.rand(4, 1, 5, 7, 9, generator=g)
    v2[1] = 0.0
    v2[2] = 1.0
    c2 = ru.voxel_tensor_to_point_clouds(v2)       # default threshold 0.5
    out["glue.vox"], out["glue.counts"], out["glue.points"] = v2, torch.tensor([len(c) for c in c2]), torch.cat(c2)
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; counts", out["counts"].tolist(), out["glue.counts"].tolist())


if __name__ == "__main__":
    main()
