"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference in place.

This module exists to (a) validate the restatement in `oracle/pointdiff_oracle.py`
against the reference's own code and (b) generate the golden vectors committed
under `tests/golden/` (see `tests/golden/make_golden.py`).  It only works where
`/root/reference` is mounted (the build container); the GPU box does not have
it, so nothing in `-m gpu` tests, `smoke()` or `bench.py` may import this file.

The reference imports `pytorch_lightning`, `matplotlib` and `plyfile` at module
top level (diffusion.py:1-12, networks.py:1-12, utils.py:1-6) and none of them
is installed here, so four stub modules are placed in `sys.modules` first.
Nothing is copied out of the reference: its files are imported where they lie.
"""
from __future__ import annotations

import inspect
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("PCD_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "diffusion.py"))


class _AttrDict(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


class _LightningModule(nn.Module):
    """Smallest stand-in that lets PointCloudDiffusion/LatentDiffusion construct and sample."""

    logger = None

    def save_hyperparameters(self, *args, ignore=None, **kwargs):
        # the reference reads self.hparams.num_points / latent_dim / is_voxel_based / lr
        # (diffusion.py:349,589,611,414): capture the caller's __init__ locals.
        frame = inspect.currentframe().f_back
        ignore = set(ignore or [])
        hp = {k: v for k, v in frame.f_locals.items()
              if k not in ("self", "__class__") and k not in ignore}
        object.__setattr__(self, "_hparams", _AttrDict(hp))

    @property
    def hparams(self):
        return self._hparams

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass


def _install_stubs() -> None:
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = _LightningModule
        pl.LightningDataModule = object
        pl.seed_everything = lambda s, *a, **k: torch.manual_seed(s)
        sys.modules["pytorch_lightning"] = pl
    for name in ("matplotlib", "matplotlib.pyplot", "deepdish"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "plyfile" not in sys.modules:
        ply = types.ModuleType("plyfile")
        ply.PlyData = object
        ply.PlyElement = object
        sys.modules["plyfile"] = ply


_cached = None


def load_reference():
    """Return (diffusion, networks, metrics) modules of the unmodified reference."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import diffusion as ref_diffusion  # noqa: E402
    import metrics as ref_metrics  # noqa: E402
    import networks as ref_networks  # noqa: E402
    _cached = (ref_diffusion, ref_networks, ref_metrics)
    return _cached


class replay_randn:
    """Context manager that makes the reference's torch.randn / randn_like calls return
    pre-recorded tensors in order (diffusion.py:239,254,275 draw from the global RNG)."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        self.i = 0

    def _next(self, shape):
        t = self.tensors[self.i]
        self.i += 1
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t.clone()

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *s, **k: self._next(s[0] if len(s) == 1 and not isinstance(s[0], int) else s)
        torch.randn_like = lambda x, **k: self._next(x.shape)
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like
        return False
