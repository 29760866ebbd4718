"""CPU: host-side logic of the product package and the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re
import sys

import pytest
import torch

import pcd_b200
from oracle import pointdiff_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pcd_b200.h")).read()
    declared = set(re.findall(r"\b(pcd_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(pcd_b200._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pcd_b200.h but not exported"
    assert set(pcd_b200._lib.EXPORTED_SYMBOLS) == declared
    assert pcd_b200._lib.lib().pcd_abi_version() == 1


def test_state_dict_layout_matches_reference_spec():
    m = pcd_b200.PointCloudDiffusion(2048)
    sd = m.state_dict()
    spec = O.state_dict_spec()
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    assert m.hparams.num_points == 2048 and m.hparams.noise_schedule == "cosine"


def test_tables_match_oracle_schedule():
    m = pcd_b200.PointCloudDiffusion(64)
    S = 10
    tab = m.ddim_table(S)
    for k in range(S):
        t = torch.ones(1) - k * (1.0 / S)
        n, s = O.offset_cosine_schedule(t)
        n2, s2 = O.offset_cosine_schedule(t - 1.0 / S)
        assert float(tab[k, 0]) == float(n) and float(tab[k, 1]) == float(s) and float(tab[k, 5]) == float(t)
        if k < S - 1:
            assert float(tab[k, 2]) == float(s2) and float(tab[k, 3]) == float(n2)
    assert tab[-1, 2] == 1.0 and tab[-1, 3] == 0.0 and float(tab[:, 4].abs().max()) == 0.0
    tab = m.ddpm_table(S)
    for k in range(S):
        i = S - 1 - k
        n, s = O.offset_cosine_schedule(torch.ones(1) * i / S)
        assert float(tab[k, 0]) == float(n) and float(tab[k, 1]) == float(s)
        if i > 0:
            n_p, s_p = O.offset_cosine_schedule(torch.ones(1) * (i - 1) / S)
            assert float(tab[k, 2]) == float(s_p) and float(tab[k, 4]) == float(torch.sqrt(n_p / n) * n)
    assert tab[-1, 2] == 1.0 and tab[-1, 4] == 0.0
    tab = m.ddim3_table(0.01, 5)
    steps = torch.linspace(0.01, 0.0, 5)
    assert [float(v) for v in tab[:, 5]] == [float(v) for v in steps]
    assert float(tab[-1, 5]) == 0.0 and tab[-1, 2] == 1.0


def test_add_and_remove_noise_follow_reference_formulas():
    m = pcd_b200.PointCloudDiffusion(32)
    torch.manual_seed(3)
    x0 = torch.randn(2, 32, 3)
    t = torch.tensor([0.2, 0.8])
    torch.manual_seed(4)
    x_t, noise, n, s = m.add_noise(x0, t)
    torch.manual_seed(4)
    ref_noise = torch.randn_like(x0)
    x_ref, _, n_ref, s_ref = O.add_noise(x0, t, ref_noise)
    assert torch.equal(noise, ref_noise) and torch.equal(x_t, x_ref) and torch.equal(n, n_ref) and torch.equal(s, s_ref)
    assert torch.allclose(m.remove_noise(x_t, noise, n, s), x0, atol=1e-5)


def test_lightning_checkpoint_round_trip(tmp_path):
    sd = O.make_synthetic_checkpoint()
    path = tmp_path / "m.ckpt"
    torch.save({"state_dict": sd, "hyper_parameters": {"num_points": 777, "dim": 256, "time_dim": 256, "lr": 1e-4,
                                                        "noise_schedule": "cosine"}}, path)
    m = pcd_b200.PointCloudDiffusion.load_from_checkpoint(str(path))
    assert m.num_points == 777
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    with pytest.raises(RuntimeError):
        bad = dict(sd); bad.pop("model.refine4.bias")
        m.load_state_dict(bad, strict=True)


def test_no_cpu_fallback():
    m = pcd_b200.PointCloudDiffusion(64)
    with pytest.raises(pcd_b200.PcdError):
        m.sample(1, 64, num_steps=2)
    with pytest.raises(pcd_b200.PcdError):
        pcd_b200.chamfer_distance(torch.randn(10, 3), torch.randn(12, 3))
    with pytest.raises(pcd_b200.PcdError):       # host buffers or not, the loop needs the GPU
        pcd_b200.PointCloudDiffusion(64, noise_schedule="linear").sample_host(torch.zeros(2, 64, 3), 4)
    with pytest.raises(pcd_b200.PcdError):
        pcd_b200.SimplePointNetVAE(64).decode(torch.zeros(1, 256))
    with pytest.raises(ValueError):
        pcd_b200.UNetPointNetLarge(dim=512, time_dim=256)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "3d-shape-generation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src, f"{fn} mentions the oracle"


def test_set_metric_definitions_agree_with_oracle():
    g = torch.Generator().manual_seed(9)
    D_gr, D_gg, D_rr = torch.rand(6, 5, generator=g), torch.rand(6, 6, generator=g), torch.rand(5, 5, generator=g)
    D_gg, D_rr = (D_gg + D_gg.t()) / 2, (D_rr + D_rr.t()) / 2
    assert pcd_b200.set_metrics_from_matrices(D_gr, D_gg, D_rr) == O.set_metrics_from_matrices(D_gr, D_gg, D_rr)


def test_linear_schedule_tables_reproduce_the_batch_cumprod_quirk():
    """diffusion.py:202 cumprods over the BATCH axis: with the reference's [B] vector of equal times every sample gets its
    own rates, so the tables carry one row per step AND sample; sample3's scalar t keeps a shared row."""
    m = pcd_b200.PointCloudDiffusion(16, noise_schedule="linear")
    S, B = 4, 3
    tab = m.ddim_table(S, B)
    assert tuple(tab.shape) == (S, B, 8)
    for step in range(S):
        t = torch.ones(B) - step * (1.0 / S)
        n, s = O.linear_schedule(t)
        n2, s2 = O.linear_schedule(t - 1.0 / S)
        assert torch.equal(tab[step, :, 0], n) and torch.equal(tab[step, :, 1], s)
        if step < S - 1:
            assert torch.equal(tab[step, :, 2], s2) and torch.equal(tab[step, :, 3], n2)
    assert not torch.equal(tab[0, 0], tab[0, 1])                   # rows differ between samples
    assert torch.equal(tab[-1, :, 2], torch.ones(B)) and torch.equal(tab[-1, :, 3], torch.zeros(B))
    tab2 = m.ddpm_table(S, B)
    assert tuple(tab2.shape) == (S, B, 8)
    n, s = O.linear_schedule(torch.ones(B) * 3 / S)
    n_p, s_p = O.linear_schedule(torch.ones(B) * 2 / S)
    assert torch.equal(tab2[0, :, 2], s_p) and torch.equal(tab2[0, :, 4], torch.sqrt(n_p / n) * n)
    assert tuple(m.ddim3_table(0.5, 3).shape) == (3, 8)
    # cosine: always one shared row per step, whatever the batch
    assert tuple(pcd_b200.PointCloudDiffusion(16).ddim_table(S, B).shape) == (S, 8)


def test_schedule_table_cache_returns_the_same_values_and_keys_on_the_schedule():
    """The per-call schedule tables are cached (3 ms of tiny torch CPU ops per 50-step call): a hit must equal a fresh build bit for bit,
    and models with different schedule parameters, samplers, step counts or start times must not share an entry."""
    from importlib import import_module
    D = import_module("3d-shape-generation_b200.diffusion")
    D._TABLE_CACHE.clear()
    m = pcd_b200.PointCloudDiffusion(64)
    a = m.ddim_table(20)
    assert torch.equal(a, D._build_ddim_table(m.diffusion_schedule, 20, 1)) and m.ddim_table(20) is a
    assert torch.equal(m.ddpm_table(20), D._build_ddpm_table(m.diffusion_schedule, 20, 1))
    assert torch.equal(m.ddim3_table(0.3, 7), D._build_ddim3_table(m.diffusion_schedule, 0.3, 7))
    assert not torch.equal(m.ddim3_table(0.3, 7), m.ddim3_table(0.31, 7))
    assert m.ddim_table(21).shape[0] == 21 and not torch.equal(m.ddim_table(20)[:, :5], m.ddpm_table(20)[:, :5])
    m2 = pcd_b200.PointCloudDiffusion(64)
    m2.cosine_max_signal_rate = 0.9                      # another schedule: must not hit m's entries
    b = m2.ddim_table(20)
    assert not torch.equal(a, b) and torch.equal(b, D._build_ddim_table(m2.diffusion_schedule, 20, 1))
    lin = pcd_b200.PointCloudDiffusion(64, noise_schedule="linear")
    assert lin.ddim_table(5, batch=3).shape == (5, 3, 8) and not torch.equal(lin.ddim_table(5, batch=3)[:, 0], lin.ddim_table(5, batch=3)[:, 2])
    assert len(D._TABLE_CACHE) <= 64


def test_latent_checkpoint_loader_and_add_noise(tmp_path):
    """`LatentDiffusion.load_from_checkpoint(path, vae=vae, is_voxel_based=...)` (reference call sites train_point_ldm.py:106,222)
    on a Lightning-style dict written without Lightning: hyper-parameters exclude the vae (diffusion.py:375), the state_dict carries
    `model.*` and the frozen `vae.*`; keyword overrides win.  `add_noise` (diffusion.py:490-504) is the reference's expression."""
    NP = 64
    sd = O.make_synthetic_latent_checkpoint(num_points=NP)
    src = pcd_b200.LatentDiffusion(pcd_b200.SimplePointNetVAE(NP), is_voxel_based=False)
    src.load_state_dict(sd, strict=False)
    full = src.state_dict()
    assert any(k.startswith("vae.") for k in full) and any(k.startswith("model.") for k in full)
    path = tmp_path / "ldm.ckpt"
    torch.save({"state_dict": full, "hyper_parameters": {"latent_dim": 256, "dim": 512, "time_dim": 256, "lr": 3e-4,
                                                         "noise_schedule": "cosine", "is_voxel_based": True}}, path)
    m = pcd_b200.LatentDiffusion.load_from_checkpoint(str(path), vae=pcd_b200.SimplePointNetVAE(NP), is_voxel_based=False)
    assert m.hparams.is_voxel_based is False and m.hparams.lr == 3e-4 and m.noise_schedule == "cosine"
    for k, v in m.state_dict().items():
        assert torch.equal(v, full[k]), k
    assert all(not p.requires_grad for p in m.vae.parameters())
    bad = dict(full); bad.pop("model.refine4.bias")
    torch.save({"state_dict": bad, "hyper_parameters": {}}, path)
    with pytest.raises(RuntimeError):
        pcd_b200.LatentDiffusion.load_from_checkpoint(str(path), vae=pcd_b200.SimplePointNetVAE(NP))
    # add_noise: same draws, same expression as the reference
    z0, t = torch.randn(3, 256), torch.tensor([0.1, 0.5, 0.9])
    torch.manual_seed(9)
    z_t, noise, n, s = m.add_noise(z0, t)
    torch.manual_seed(9)
    want_noise = torch.randn_like(z0)
    wn, ws = O.offset_cosine_schedule(t)
    assert torch.equal(noise, want_noise) and torch.equal(n, wn) and torch.equal(s, ws)
    assert torch.equal(z_t, ws.view(-1, 1) * z0 + wn.view(-1, 1) * want_noise)
    # VAE checkpoints load the same way (train_point_ldm.py:43)
    torch.save({"state_dict": src.vae.state_dict(), "hyper_parameters": {"num_points": NP, "latent_dim": 256, "hidden_dim": 512}}, path)
    vae = pcd_b200.SimplePointNetVAE.load_from_checkpoint(str(path))
    assert vae.hparams.num_points == NP and torch.equal(vae.output_layer.weight, src.vae.output_layer.weight)
    v3 = pcd_b200.VAE3DLarge(latent_dim=256)
    torch.save({"state_dict": v3.state_dict(), "hyper_parameters": {"input_shape": (32, 32, 32), "latent_dim": 256, "lr": 2e-4}}, path)
    v3b = pcd_b200.VAE3DLarge.load_from_checkpoint(str(path))       # test_point_ldm.py:157
    assert v3b.hparams.lr == 2e-4 and torch.equal(v3b.decoder_input.weight, v3.decoder_input.weight)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm) prints ONE JSON line with the contract's keys, at a size the CPU
    finishes in seconds (the arm times the oracle port = the reference's CPU arithmetic; no GPU, no library call)."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--points", "64", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "shapes/sec" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert set(d["config"]) == {"workload", "loop_steps", "batch_per_gpu", "points", "parallelism", "l2"}
