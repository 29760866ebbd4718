"""Host-side mirror of the reference's denoiser module tree (networks.py:16-49, 724-838).

The nn.Modules here are *parameter containers* with exactly the reference's names and shapes,
so `state_dict()` / `load_state_dict(strict=True)` interoperate with the reference in both
directions.  `forward` does not run torch layers: it calls the sm_100a library through the
C ABI (`_lib.Denoiser`), and raises if that is impossible (no CUDA device / library missing).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class PointNetLayer(nn.Module):
    """conv1d(k=1) -> BatchNorm1d -> ReLU, three times (reference networks.py:16-49)."""

    def __init__(self, in_dim: int, mid_dim: int, out_dim: int | None = None):
        super().__init__()
        out_dim = mid_dim if out_dim is None else out_dim
        self.conv1 = nn.Conv1d(in_dim, mid_dim, 1)
        self.bn1 = nn.BatchNorm1d(mid_dim)
        self.conv2 = nn.Conv1d(mid_dim, mid_dim, 1)
        self.bn2 = nn.BatchNorm1d(mid_dim)
        self.conv3 = nn.Conv1d(mid_dim, out_dim, 1)
        self.bn3 = nn.BatchNorm1d(out_dim)

    def forward(self, x):  # pragma: no cover - never used: the whole network runs in the CUDA library
        raise _lib.PcdError("PointNetLayer is a parameter container; call UNetPointNetLarge.forward")


class UNetPointNetLarge(nn.Module):
    """Per-point shared-MLP U-Net denoiser (reference networks.py:724-838)."""

    def __init__(self, dim: int = 512, time_dim: int = 256, precision: str = "f16mix"):
        super().__init__()
        if dim != time_dim:
            # the reference itself only works for dim == time_dim: time_mlp emits `dim`,
            # enc1 expects 3 + time_dim input channels (networks.py:738-744)
            raise ValueError("UNetPointNetLarge requires dim == time_dim (reference networks.py:738-744)")
        if not (4 <= time_dim <= 4096):
            raise ValueError("time_dim out of range (4..4096)")
        self.time_dim = time_dim
        self.precision = precision
        self.time_mlp = nn.Sequential(nn.Linear(time_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.enc1 = PointNetLayer(3 + time_dim, 64, 128)
        self.enc2 = PointNetLayer(128, 128, 256)
        self.enc3 = PointNetLayer(256, 256, 512)
        self.enc4 = PointNetLayer(512, 512, 1024)
        self.global_feat = nn.Sequential(nn.Conv1d(1024, 2048, 1), nn.BatchNorm1d(2048), nn.ReLU(),
                                         nn.Conv1d(2048, 4096, 1), nn.BatchNorm1d(4096), nn.ReLU())
        self.dec4 = PointNetLayer(4096 + 1024, 1024, 512)
        self.dec3 = PointNetLayer(512 + 512, 512, 256)
        self.dec2 = PointNetLayer(256 + 256, 256, 128)
        self.dec1 = PointNetLayer(128 + 128, 128, 64)
        self.output = nn.Sequential(nn.Conv1d(64, 64, 1), nn.BatchNorm1d(64), nn.ReLU(), nn.Conv1d(64, 3, 1))
        self.refine1 = nn.Conv1d(128, 128, 1)
        self.refine2 = nn.Conv1d(256, 256, 1)
        self.refine3 = nn.Conv1d(512, 512, 1)
        self.refine4 = nn.Conv1d(1024, 1024, 1)
        self._engine = None
        self._engine_key = None

    # -- engine management ------------------------------------------------------------------
    def _weights_version(self):
        return tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())

    def engine(self) -> _lib.Denoiser:
        p = next(self.parameters())
        if p.device.type != "cuda":
            raise _lib.PcdError("model is on %s: the B200 sampling path has no CPU fallback; call .to('cuda')" % p.device)
        if self.training:
            raise _lib.PcdError("the B200 path implements eval-mode BatchNorm only; call .eval() (sample*() do)")
        key = (p.device, self.precision, self._weights_version())
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            sd = {"model." + k: v for k, v in self.state_dict().items()}
            self._engine = _lib.Denoiser(sd, p.device, self.precision)
            self._engine_key = key
        return self._engine

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """x [B,N,3], t [B] -> predicted noise [B,N,3] (reference networks.py:779-818)."""
        return self.engine().forward(x, t)

    def precision_gap(self, x: torch.Tensor, t: torch.Tensor, against: str = "bf16x3") -> float:
        """Relative L2 distance between this module's output and the same weights run in precision `against` on one forward.
        Not in the reference: a check for a new checkpoint.  The fp16 modes (`f16mix`, `f16`) clamp activations at +-65504
        instead of overflowing, so a checkpoint whose activations leave that range gives finite but wrong results; the bf16
        planes of `bf16x3` have fp32's exponent range.  Expect ~6e-4 for `f16mix`; a gap of 1e-2 or more means saturation --
        use `precision='bf16x3'` for that checkpoint."""
        own = self.forward(x, t)
        other = UNetPointNetLarge(self.time_dim, self.time_dim, precision=against)
        other.load_state_dict(self.state_dict(), strict=True)
        other = other.to(x.device).eval()
        ref = other.forward(x, t)
        gap = float((own.double() - ref.double()).norm() / ref.double().norm())
        other._engine.close()
        return gap

    def get_timestep_embedding(self, timesteps: torch.Tensor, embedding_dim: int) -> torch.Tensor:
        """Sinusoidal embedding (reference networks.py:820-838); host-side utility, the fused path
        recomputes it on the device."""
        half = embedding_dim // 2
        emb = torch.log(torch.tensor(10000.0, device=timesteps.device)) / (half - 1)
        emb = torch.exp(torch.arange(half, device=timesteps.device) * -emb)
        emb = timesteps[:, None] * emb[None, :]
        emb = torch.cat((torch.sin(emb), torch.cos(emb)), dim=-1)
        if embedding_dim % 2 == 1:
            emb = torch.nn.functional.pad(emb, (0, 1))
        return emb
