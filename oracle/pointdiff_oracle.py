"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's point-cloud diffusion
sampling hot path (fp32, torch CPU ops).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this; the product path
(`3d-shape-generation_b200/`) never does and fails loudly without its CUDA library.

Parity status: the reference ships no golden tensors for this path;
this restatement is pinned instead by (1) differential tests against the reference's own
code imported in place (`tests/test_oracle_vs_reference.py`, runs where /root/reference is
mounted) and (2) golden vectors generated from the reference and committed under
`tests/golden/` (`tests/golden/make_golden.py`), which travel to the GPU box.

Every function cites the reference file:line it follows.  The functions operate on a plain
`state_dict` (keys exactly as the reference's `PointCloudDiffusion.state_dict()`:
`model.enc1.conv1.weight` ...), so no reference class is needed at run time.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5  # nn.BatchNorm1d default, networks.py:30

# ----------------------------------------------------------------------------------------
# architecture table (networks.py:743-777): block -> (Cin, Cmid, Cout)
# ----------------------------------------------------------------------------------------
POINTNET_BLOCKS = {
    "enc1": (3 + 256, 64, 128),
    "enc2": (128, 128, 256),
    "enc3": (256, 256, 512),
    "enc4": (512, 512, 1024),
    "dec4": (4096 + 1024, 1024, 512),
    "dec3": (512 + 512, 512, 256),
    "dec2": (256 + 256, 256, 128),
    "dec1": (128 + 128, 128, 64),
}
REFINE = {"refine1": 128, "refine2": 256, "refine3": 512, "refine4": 1024}


def state_dict_spec(dim: int = 256, time_dim: int = 256) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every entry of PointCloudDiffusion.state_dict(), in module
    registration order (networks.py:737-777, PointNetLayer networks.py:29-34).
    kind in {conv_w, lin_w, bias, bn_w, bn_b, bn_mean, bn_var, bn_count}."""
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def lin(name, cin, cout):
        spec.append((f"{name}.weight", (cout, cin), "lin_w"))
        spec.append((f"{name}.bias", (cout,), "bias"))

    def conv(name, cin, cout):
        spec.append((f"{name}.weight", (cout, cin, 1), "conv_w"))
        spec.append((f"{name}.bias", (cout,), "bias"))

    def bn(name, c):
        spec.append((f"{name}.weight", (c,), "bn_w"))
        spec.append((f"{name}.bias", (c,), "bn_b"))
        spec.append((f"{name}.running_mean", (c,), "bn_mean"))
        spec.append((f"{name}.running_var", (c,), "bn_var"))
        spec.append((f"{name}.num_batches_tracked", (), "bn_count"))

    def block(name, cin, cmid, cout):
        conv(f"{name}.conv1", cin, cmid); bn(f"{name}.bn1", cmid)
        conv(f"{name}.conv2", cmid, cmid); bn(f"{name}.bn2", cmid)
        conv(f"{name}.conv3", cmid, cout); bn(f"{name}.bn3", cout)

    lin("model.time_mlp.0", time_dim, dim)
    lin("model.time_mlp.2", dim, dim)
    for name in ("enc1", "enc2", "enc3", "enc4"):
        cin, cmid, cout = POINTNET_BLOCKS[name]
        if name == "enc1":
            cin = 3 + time_dim
        block(f"model.{name}", cin, cmid, cout)
    conv("model.global_feat.0", 1024, 2048); bn("model.global_feat.1", 2048)
    conv("model.global_feat.3", 2048, 4096); bn("model.global_feat.4", 4096)
    for name in ("dec4", "dec3", "dec2", "dec1"):
        block(f"model.{name}", *POINTNET_BLOCKS[name])
    conv("model.output.0", 64, 64); bn("model.output.1", 64)
    conv("model.output.3", 64, 3)
    for name in ("refine1", "refine2", "refine3", "refine4"):
        conv(f"model.{name}", REFINE[name], REFINE[name])
    return spec


def make_synthetic_checkpoint(seed: int = 24, alpha: float = 1.0 / 33.0, bn_seed: int = 7,
                              dim: int = 256, time_dim: int = 256) -> SD:
    """Deterministic synthetic weights with the reference's *initialisation law*
    (diffusion.py:40-54: Kaiming-normal fan_out/ReLU on Conv1d/Linear, zero bias) but with
    (a) `output.3.weight` scaled by `alpha` so the sampling loop stays finite (SURVEY H1:
    random init has |eps_hat| ~ 33|x| and the DDIM map has gain 47.5) and (b) randomised
    BatchNorm running stats / affine so a BN-folding bug cannot hide behind identity BN.
    Biases are also randomised (small) so a dropped bias is visible."""
    g = torch.Generator().manual_seed(seed)
    gb = torch.Generator().manual_seed(bn_seed)
    sd: SD = {}
    for key, shape, kind in state_dict_spec(dim, time_dim):
        if kind in ("conv_w", "lin_w"):
            fan_out = shape[0]  # kernel size 1
            sd[key] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
        elif kind == "bias":
            sd[key] = torch.randn(shape, generator=gb) * 0.05
        elif kind == "bn_w":
            sd[key] = 1.0 + 0.2 * torch.randn(shape, generator=gb)
        elif kind == "bn_b":
            sd[key] = 0.1 * torch.randn(shape, generator=gb)
        elif kind == "bn_mean":
            sd[key] = 0.1 * torch.randn(shape, generator=gb)
        elif kind == "bn_var":
            sd[key] = 0.5 + torch.rand(shape, generator=gb)
        elif kind == "bn_count":
            sd[key] = torch.tensor(1, dtype=torch.int64)
    sd["model.output.3.weight"] = sd["model.output.3.weight"] * alpha
    sd["model.output.3.bias"] = sd["model.output.3.bias"] * alpha
    return sd


# ----------------------------------------------------------------------------------------
# denoiser (networks.py:779-838)
# ----------------------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, embedding_dim: int) -> torch.Tensor:
    """networks.py:820-838.  t is a float in [0,1], not an integer step."""
    half = embedding_dim // 2
    emb = torch.log(torch.tensor(10000.0)) / (half - 1)
    emb = torch.exp(torch.arange(half) * -emb)
    emb = t[:, None] * emb[None, :]
    emb = torch.cat((torch.sin(emb), torch.cos(emb)), dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def _bn_eval(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(x, sd[f"{name}.running_mean"], sd[f"{name}.running_var"],
                        sd[f"{name}.weight"], sd[f"{name}.bias"], False, 0.0, BN_EPS)


def _conv(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.conv1d(x, sd[f"{name}.weight"], sd[f"{name}.bias"])


def _pointnet_layer(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """networks.py:46-48: three times conv1d(k=1) -> BatchNorm1d(eval) -> ReLU."""
    for i in (1, 2, 3):
        x = F.relu(_bn_eval(sd, f"{name}.bn{i}", _conv(sd, f"{name}.conv{i}", x)))
    return x


def time_mlp(sd: SD, t: torch.Tensor, time_dim: Optional[int] = None) -> torch.Tensor:
    """networks.py:737-741,791-792: Linear -> SiLU -> Linear on the sinusoidal embedding (time_dim defaults to the checkpoint's)."""
    if time_dim is None:
        time_dim = sd["model.time_mlp.0.weight"].shape[1]
    e = timestep_embedding(t, time_dim)
    e = F.linear(e, sd["model.time_mlp.0.weight"], sd["model.time_mlp.0.bias"])
    e = F.silu(e)
    return F.linear(e, sd["model.time_mlp.2.weight"], sd["model.time_mlp.2.bias"])


def denoiser_forward(sd: SD, x: torch.Tensor, t: torch.Tensor, time_dim: Optional[int] = None,
                     taps: Optional[dict] = None) -> torch.Tensor:
    """UNetPointNetLarge.forward, networks.py:779-818.  x [B,N,3], t [B] -> eps_hat [B,N,3].
    `taps`, if given, receives intermediate activations (channel-major [B,C,N])."""
    temb = time_mlp(sd, t, time_dim)                                  # :791-792
    h = x.transpose(2, 1)                                            # :795
    h = torch.cat([h, temb.unsqueeze(2).expand(-1, -1, h.shape[2])], dim=1)  # :796-797
    x1 = _pointnet_layer(sd, "model.enc1", h)                         # :800-803
    x2 = _pointnet_layer(sd, "model.enc2", x1)
    x3 = _pointnet_layer(sd, "model.enc3", x2)
    x4 = _pointnet_layer(sd, "model.enc4", x3)
    g = F.relu(_bn_eval(sd, "model.global_feat.1", _conv(sd, "model.global_feat.0", x4)))   # :806
    g = F.relu(_bn_eval(sd, "model.global_feat.4", _conv(sd, "model.global_feat.3", g)))
    g = torch.max(g, 2, keepdim=True)[0]                             # :807
    if taps is not None:
        taps.update(temb=temb, x1=x1, x2=x2, x3=x3, x4=x4, g=g[:, :, 0])
    g = g.repeat(1, 1, h.shape[2])                                   # :808
    d = _pointnet_layer(sd, "model.dec4", torch.cat([g, _conv(sd, "model.refine4", x4)], dim=1))  # :811
    if taps is not None:
        taps["d4"] = d
    d = _pointnet_layer(sd, "model.dec3", torch.cat([d, _conv(sd, "model.refine3", x3)], dim=1))
    d = _pointnet_layer(sd, "model.dec2", torch.cat([d, _conv(sd, "model.refine2", x2)], dim=1))
    d = _pointnet_layer(sd, "model.dec1", torch.cat([d, _conv(sd, "model.refine1", x1)], dim=1))
    if taps is not None:
        taps["d1"] = d
    d = F.relu(_bn_eval(sd, "model.output.1", _conv(sd, "model.output.0", d)))   # :816
    d = _conv(sd, "model.output.3", d)
    return d.transpose(2, 1)                                         # :818


# ----------------------------------------------------------------------------------------
# schedules (diffusion.py:189-223)
# ----------------------------------------------------------------------------------------
COSINE_MIN_SIGNAL = 0.02   # diffusion.py:34
COSINE_MAX_SIGNAL = 0.95   # diffusion.py:35
LINEAR_MIN_RATE = 0.0001   # diffusion.py:32
LINEAR_MAX_RATE = 0.02     # diffusion.py:33


def offset_cosine_schedule(t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """diffusion.py:208-223 -> (noise_rates, signal_rates)."""
    a0 = torch.acos(torch.tensor(COSINE_MAX_SIGNAL))
    a1 = torch.acos(torch.tensor(COSINE_MIN_SIGNAL))
    ang = a0 + t * (a1 - a0)
    return torch.sin(ang), torch.cos(ang)


def linear_schedule(t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """diffusion.py:189-205.  NB: cumprod runs over the *batch* axis."""
    betas = LINEAR_MIN_RATE + t.clone() * (LINEAR_MAX_RATE - LINEAR_MIN_RATE)
    alpha_bars = torch.cumprod(1 - betas, dim=0)
    return 1 - alpha_bars, alpha_bars


def schedule_fn(name: str):
    return offset_cosine_schedule if name == "cosine" else linear_schedule


# ----------------------------------------------------------------------------------------
# samplers (diffusion.py:225-337)
# ----------------------------------------------------------------------------------------
def remove_noise(x_t, eps, noise_rates, signal_rates):
    """diffusion.py:154-168."""
    return (x_t - noise_rates.view(-1, 1, 1) * eps) / signal_rates.view(-1, 1, 1)


def ddim_sample(sd: SD, x_T: torch.Tensor, num_steps: int, schedule: str = "cosine",
                record: Optional[list] = None) -> torch.Tensor:
    """`PointCloudDiffusion.sample`, diffusion.py:261-289.  Returns the LAST x_0 (not x_t)."""
    sched = schedule_fn(schedule)
    B = x_T.shape[0]
    x_t = x_T
    step_size = 1.0 / num_steps
    x_0 = x_T
    for step in range(num_steps):
        t = torch.ones(B) - step * step_size
        n, s = sched(t)
        eps = denoiser_forward(sd, x_t, t)
        if record is not None:
            record.append(eps)
        x_0 = remove_noise(x_t, eps, n, s)
        n2, s2 = sched(t - step_size)
        x_t = s2.view(-1, 1, 1) * x_0 + n2.view(-1, 1, 1) * eps
    return x_0


def ddpm_sample(sd: SD, x_T: torch.Tensor, noises: Sequence[torch.Tensor], num_steps: int,
                schedule: str = "cosine") -> torch.Tensor:
    """`PointCloudDiffusion.sample2`, diffusion.py:225-259.  noises[j] is the j-th
    `randn_like` draw, i.e. the one used at i = num_steps-1-j (i > 0)."""
    sched = schedule_fn(schedule)
    B = x_T.shape[0]
    x_t = x_T
    j = 0
    for i in reversed(range(num_steps)):
        t = torch.ones(B) * i / num_steps
        n, s = sched(t)
        eps = denoiser_forward(sd, x_t, t)
        x_0 = remove_noise(x_t, eps, n, s)
        if i > 0:
            n_p, s_p = sched(torch.ones(B) * (i - 1) / num_steps)
            coefficient = torch.sqrt(n_p / n)
            z = noises[j]; j += 1
            x_t = s_p.view(-1, 1, 1) * x_0 + coefficient.view(-1, 1, 1) * n.view(-1, 1, 1) * z
        else:
            x_t = x_0
    return x_t


def ddim3_sample(sd: SD, x: torch.Tensor, start_t: torch.Tensor, num_steps: int,
                 schedule: str = "cosine") -> torch.Tensor:
    """`PointCloudDiffusion.sample3`, diffusion.py:291-337 (scalar t shared by the batch;
    last iteration evaluates the model at t=0 and does not re-noise)."""
    sched = schedule_fn(schedule)
    B = x.shape[0]
    steps = torch.linspace(float(start_t[0]), 0.0, num_steps)
    x_0 = x
    for i in range(num_steps):
        t = steps[i]
        n, s = sched(t)
        eps = denoiser_forward(sd, x, t.expand(B))
        x_0 = remove_noise(x, eps, n, s)
        if i < num_steps - 1:
            n2, s2 = sched(steps[i + 1])
            x = s2.view(-1, 1, 1) * x_0 + n2.view(-1, 1, 1) * eps
    return x_0


def add_noise(x_0: torch.Tensor, t: torch.Tensor, noise: torch.Tensor, schedule: str = "cosine"):
    """diffusion.py:138-152 with the `randn_like` draw injected."""
    n, s = schedule_fn(schedule)(t)
    return s.view(-1, 1, 1) * x_0 + n.view(-1, 1, 1) * noise, noise, n, s


# ----------------------------------------------------------------------------------------
# Chamfer core (metrics.py:7-47)
# ----------------------------------------------------------------------------------------
def normalize_to_cube(points: torch.Tensor) -> torch.Tensor:
    """metrics.py:7-21: AABB centre, then one scalar max-|.| scale per cloud."""
    center = (points.max(dim=1, keepdim=True)[0] + points.min(dim=1, keepdim=True)[0]) / 2
    points = points - center
    scale = points.abs().max(dim=1, keepdim=True)[0].max(dim=2, keepdim=True)[0]
    return points / scale


def chamfer_distance(x: torch.Tensor, y: torch.Tensor, scaling_factor: float = 1e3,
                     exact: bool = False) -> torch.Tensor:
    """metrics.py:23-47.  Batched input -> ONE scalar averaged over the batch.
    exact=True uses direct differences instead of cdist's matmul path."""
    x = x.unsqueeze(0) if x.dim() == 2 else x
    y = y.unsqueeze(0) if y.dim() == 2 else y
    x = normalize_to_cube(x)
    y = normalize_to_cube(y)
    dist = torch.cdist(x, y, compute_mode="donot_use_mm_for_euclid_dist" if exact else
                       "use_mm_for_euclid_dist_if_necessary")
    cd = torch.mean(torch.min(dist, dim=2)[0]) + torch.mean(torch.min(dist, dim=1)[0])
    return cd * scaling_factor


def chamfer_pairs(x: torch.Tensor, y: torch.Tensor, scaling_factor: float = 1e3):
    """Per-pair decomposition of metrics.py:35-47 used by the set metrics: returns
    (cd[B], idx_xy[B,N], idx_yx[B,M]) with direct-difference distances (exact NN)."""
    xn, yn = normalize_to_cube(x), normalize_to_cube(y)
    dist = torch.cdist(xn, yn, compute_mode="donot_use_mm_for_euclid_dist")
    dxy, ixy = torch.min(dist, dim=2)
    dyx, iyx = torch.min(dist, dim=1)
    return (dxy.mean(dim=1) + dyx.mean(dim=1)) * scaling_factor, ixy, iyx


def chamfer_matrix(G: torch.Tensor, R: torch.Tensor, scaling_factor: float = 1e3) -> torch.Tensor:
    """CD[i,j] = metrics.chamfer_distance(G[i], R[j]) for all pairs (reference per-pair
    semantics: non-squared L2, per-cloud cube normalisation, sum of directional means)."""
    Gn, Rn = normalize_to_cube(G), normalize_to_cube(R)
    out = torch.empty(G.shape[0], R.shape[0])
    for i in range(G.shape[0]):
        d = torch.cdist(Gn[i:i + 1].expand(R.shape[0], -1, -1), Rn,
                        compute_mode="donot_use_mm_for_euclid_dist")
        out[i] = (d.min(dim=2)[0].mean(dim=1) + d.min(dim=1)[0].mean(dim=1)) * scaling_factor
    return out


def set_metrics_from_matrices(D_gr: torch.Tensor, D_gg: torch.Tensor, D_rr: torch.Tensor) -> dict:
    """MMD-CD / COV-CD / 1-NNA-CD (Achlioptas et al. 2018; Yang et al. 2019) on top of the
    reference's per-pair CD.  Not in the reference; definitions:
      MMD = mean_r min_g D[g,r];  COV = |{argmin_r D[g,r]}| / |R|;
      1-NNA = leave-one-out 1-NN accuracy over G u R with the block matrix [[GG,GR],[RG,RR]]."""
    nG, nR = D_gr.shape
    mmd = D_gr.min(dim=0)[0].mean()
    cov = torch.unique(D_gr.argmin(dim=1)).numel() / nR
    top = torch.cat([D_gg, D_gr], dim=1)
    bot = torch.cat([D_gr.t(), D_rr], dim=1)
    full = torch.cat([top, bot], dim=0).clone()
    full.fill_diagonal_(float("inf"))
    nn_idx = full.argmin(dim=1)
    label = torch.cat([torch.zeros(nG), torch.ones(nR)])
    acc = (label[nn_idx] == label).float().mean()
    return {"mmd_cd": float(mmd), "cov_cd": float(cov), "1nna_cd": float(acc)}


# ----------------------------------------------------------------------------------------
# Sinkhorn EMD (metrics.py:94-158) -- SURVEY 8(f) rank 3
# ----------------------------------------------------------------------------------------
def voxelize(points: torch.Tensor, voxel_resolution: int = 32) -> torch.Tensor:
    """utils.py:488-509: occupancy grid of a cloud in [-1, 1]^3 (truncating cast, clamped, per-sample scatter loop)."""
    points = points.unsqueeze(0) if points.dim() == 2 else points
    points = (points + 1) * (voxel_resolution - 1) / 2
    points = points.long().clamp(0, voxel_resolution - 1)
    voxels = torch.zeros(points.size(0), voxel_resolution, voxel_resolution, voxel_resolution)
    for i in range(points.size(0)):
        voxels[i, points[i, :, 0], points[i, :, 1], points[i, :, 2]] = 1
    return voxels


def voxel_bce(generated: torch.Tensor, reference: torch.Tensor) -> torch.Tensor:
    """metrics.py:181: the `recon_loss` of compute_metrics."""
    return torch.nn.functional.binary_cross_entropy(voxelize(generated), voxelize(reference))


def sinkhorn_emd(x: torch.Tensor, y: torch.Tensor, epsilon: float = 1e-2, thresh: float = 1e-5, max_iter: int = 100,
                 scaling_factor: float = 1, exact: bool = False, per_pair: bool = False):
    """`earth_mover_distance_gpu`, metrics.py:94-158, with the reference's exact update order:
    cost = cdist / cdist.max() (ONE maximum over the whole batch, :124); duals start at 0 (:130-131);
    alpha <- eps*(log(mu+1e-10) - LSE_j(-C/eps + beta_j)) then beta from the NEW alpha (:142-145); the loop
    stops when BOTH max-abs dual changes over the whole batch drop below `thresh` (:148-151);
    P = exp(-C/eps + alpha + beta^T) and emd = sum(P*C) per pair (:154-157).
    exact=True uses direct-difference distances (what the CUDA kernel computes) instead of cdist's matmul path."""
    x = x.unsqueeze(0) if x.dim() == 2 else x
    y = y.unsqueeze(0) if y.dim() == 2 else y
    x = normalize_to_cube(x)
    y = normalize_to_cube(y)
    batch_size, n, _ = x.shape
    _, m, _ = y.shape
    C = torch.cdist(x, y, p=2, compute_mode="donot_use_mm_for_euclid_dist" if exact else "use_mm_for_euclid_dist_if_necessary")
    C = C / C.max()
    lambda_val = 1 / epsilon
    alpha = torch.zeros(batch_size, n, 1)
    beta = torch.zeros(batch_size, m, 1)
    mu = torch.ones(batch_size, n, 1) / n
    nu = torch.ones(batch_size, m, 1) / m
    iters = 0
    for _ in range(max_iter):
        alpha_prev, beta_prev = alpha, beta
        alpha = epsilon * (torch.log(mu + 1e-10) - torch.logsumexp(-lambda_val * C + beta.transpose(1, 2), dim=2, keepdim=True))
        beta = epsilon * (torch.log(nu + 1e-10) - torch.logsumexp(-lambda_val * C.transpose(1, 2) + alpha.transpose(1, 2), dim=2, keepdim=True))
        iters += 1
        if torch.abs(alpha - alpha_prev).max() < thresh and torch.abs(beta - beta_prev).max() < thresh:
            break
    P = torch.exp(-lambda_val * C + alpha + beta.transpose(1, 2))
    emd = torch.sum(P * C, dim=(1, 2))
    if per_pair:
        return emd * scaling_factor, iters
    return emd.mean() * scaling_factor


# ----------------------------------------------------------------------------------------
# latent path (BASELINE config 4): SimpleLatentUNetPointNet (networks.py:962-1106),
# LatentDiffusion samplers (diffusion.py:575-707), SimplePointNetVAE.decode (networks.py:1144-1154, 1219-1231)
# ----------------------------------------------------------------------------------------
GN_EPS = 1e-5  # nn.GroupNorm default


def latent_state_dict_spec(latent_dim: int = 256, dim: int = 512, time_dim: int = 256, num_points: int = 2048,
                           hidden_dim: int = 512):
    """(key, shape, kind) for the `model.*` (latent denoiser) and the decoder part of `vae.*` of
    LatentDiffusion(SimplePointNetVAE(num_points)).state_dict().  The VAE encoder entries
    (vae.encoder.*, vae.fc_mu, vae.fc_logvar) are not on the sampling path and are omitted here."""
    spec = []

    def lin(name, cin, cout):
        spec.append((f"{name}.weight", (cout, cin), "lin_w"))
        spec.append((f"{name}.bias", (cout,), "bias"))

    def gn(name, c):
        spec.append((f"{name}.weight", (c,), "bn_w"))
        spec.append((f"{name}.bias", (c,), "bn_b"))

    lin("vae.decoder.0", latent_dim, hidden_dim // 2)
    lin("vae.decoder.2", hidden_dim // 2, hidden_dim)
    lin("vae.decoder.4", hidden_dim, num_points * 3)
    lin("vae.output_layer", num_points * 3, num_points * 3)
    lin("model.time_mlp.0", time_dim, time_dim)
    lin("model.time_mlp.2", time_dim, time_dim)
    d4, d2 = dim // 4, dim // 2
    for name, cin, cout in (("enc1", latent_dim + time_dim, d4), ("enc2", d4, d2), ("enc3", d2, dim), ("enc4", dim, dim * 2)):
        lin(f"model.{name}.0", cin, cout); gn(f"model.{name}.1", cout)
    lin("model.global_feat.0", dim * 2, dim * 4); gn("model.global_feat.1", dim * 4)
    lin("model.global_feat.3", dim * 4, dim * 8); gn("model.global_feat.4", dim * 8)
    for name, cin, cout in (("dec4", dim * 8 + dim * 2, dim * 2), ("dec3", dim * 2 + dim, dim), ("dec2", dim + d2, d2),
                            ("dec1", d2 + d4, d4)):
        lin(f"model.{name}.0", cin, cout); gn(f"model.{name}.1", cout)
    lin("model.output.0", d4, d4)
    lin("model.output.2", d4, latent_dim)
    for name, c in (("refine1", d4), ("refine2", d2), ("refine3", dim), ("refine4", dim * 2)):
        lin(f"model.{name}", c, c)
    return spec


def make_synthetic_latent_checkpoint(seed: int = 24, alpha: float = 1.0 / 16.0, aff_seed: int = 7, num_points: int = 2048) -> SD:
    """Synthetic weights for the latent path (Kaiming fan_out like diffusion.py:391-408), randomised
    GroupNorm affine and biases, `output.2` scaled by alpha so the latent loop stays finite."""
    g = torch.Generator().manual_seed(seed)
    gb = torch.Generator().manual_seed(aff_seed)
    sd: SD = {}
    for key, shape, kind in latent_state_dict_spec(num_points=num_points):
        if kind == "lin_w":
            sd[key] = torch.randn(shape, generator=g) * math.sqrt(2.0 / shape[0])
        elif kind == "bias":
            sd[key] = torch.randn(shape, generator=gb) * 0.05
        elif kind == "bn_w":
            sd[key] = 1.0 + 0.2 * torch.randn(shape, generator=gb)
        elif kind == "bn_b":
            sd[key] = 0.1 * torch.randn(shape, generator=gb)
    sd["model.output.2.weight"] = sd["model.output.2.weight"] * alpha
    sd["model.output.2.bias"] = sd["model.output.2.bias"] * alpha
    return sd


def _lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[f"{name}.weight"], sd[f"{name}.bias"])


def _lin_gn_relu(sd: SD, name: str, idx: int, x: torch.Tensor) -> torch.Tensor:
    """nn.Sequential(Linear, GroupNorm(8, C), ReLU) (networks.py:984-1031)."""
    y = _lin(sd, f"{name}.{idx}", x)
    y = F.group_norm(y, 8, sd[f"{name}.{idx + 1}.weight"], sd[f"{name}.{idx + 1}.bias"], GN_EPS)
    return F.relu(y)


def latent_time_mlp(sd: SD, t: torch.Tensor, time_dim: int = 256) -> torch.Tensor:
    e = timestep_embedding(t, time_dim)
    e = F.silu(_lin(sd, "model.time_mlp.0", e))
    return _lin(sd, "model.time_mlp.2", e)


def latent_denoiser_forward(sd: SD, z: torch.Tensor, t: torch.Tensor, time_dim: int = 256, taps: Optional[dict] = None):
    """SimpleLatentUNetPointNet.forward (networks.py:1051-1086), eval mode (Dropout = identity)."""
    temb = latent_time_mlp(sd, t, time_dim)
    h = torch.cat([z, temb], dim=1)
    z1 = _lin_gn_relu(sd, "model.enc1", 0, h)
    z2 = _lin_gn_relu(sd, "model.enc2", 0, z1)
    z3 = _lin_gn_relu(sd, "model.enc3", 0, z2)
    z4 = _lin_gn_relu(sd, "model.enc4", 0, z3)
    g = _lin_gn_relu(sd, "model.global_feat", 0, z4)
    g = _lin_gn_relu(sd, "model.global_feat", 3, g)
    d = _lin_gn_relu(sd, "model.dec4", 0, torch.cat([g, _lin(sd, "model.refine4", z4)], dim=1))
    d = _lin_gn_relu(sd, "model.dec3", 0, torch.cat([d, _lin(sd, "model.refine3", z3)], dim=1))
    d = _lin_gn_relu(sd, "model.dec2", 0, torch.cat([d, _lin(sd, "model.refine2", z2)], dim=1))
    d = _lin_gn_relu(sd, "model.dec1", 0, torch.cat([d, _lin(sd, "model.refine1", z1)], dim=1))
    if taps is not None:
        taps.update(temb=temb, z1=z1, z4=z4, g=g, d1=d)
    d = F.relu(_lin(sd, "model.output.0", d))
    return _lin(sd, "model.output.2", d)


def vae_decode(sd: SD, z: torch.Tensor, num_points: int = 2048) -> torch.Tensor:
    """SimplePointNetVAE.decode (networks.py:1219-1231; layers :1144-1154), eval mode."""
    h = F.relu(_lin(sd, "vae.decoder.0", z))
    h = F.relu(_lin(sd, "vae.decoder.2", h))
    h = F.relu(_lin(sd, "vae.decoder.4", h))
    return _lin(sd, "vae.output_layer", h).view(-1, num_points, 3)


def latent_ddim_sample(sd: SD, z_T: torch.Tensor, num_steps: int, num_points: int = 2048, decode: bool = True,
                       schedule: str = "cosine"):
    """LatentDiffusion.sample (diffusion.py:619-653) for a point-based VAE.  NB: the reference crashes
    here when is_voxel_based=False (`point_clouds` is only assigned in the voxel branch, :650-653);
    the defined behaviour, mirroring sample2's else branch (:611-614), is to return vae.decode(z_0)."""
    B = z_T.shape[0]
    z_t = z_T
    step_size = 1.0 / num_steps
    z_0 = z_T
    sched = schedule_fn(schedule)     # 'linear' cumprods over the batch axis (diffusion.py:553-569 = :189-205)
    for step in range(num_steps):
        t = torch.ones(B) - step * step_size
        n, s = sched(t)
        eps = latent_denoiser_forward(sd, z_t, t)
        z_0 = (z_t - n.view(-1, 1) * eps) / s.view(-1, 1)
        n2, s2 = sched(t - step_size)
        z_t = s2.view(-1, 1) * z_0 + n2.view(-1, 1) * eps
    return vae_decode(sd, z_0, num_points) if decode else z_0


def latent_ddpm_sample(sd: SD, z_T: torch.Tensor, noises, num_steps: int, num_points: int = 2048, decode: bool = True,
                       schedule: str = "cosine"):
    """LatentDiffusion.sample2 (diffusion.py:575-616)."""
    B = z_T.shape[0]
    z_t = z_T
    j = 0
    offset_cosine_schedule = schedule_fn(schedule)      # the loop below is schedule agnostic
    for i in reversed(range(num_steps)):
        t = torch.ones(B) * i / num_steps
        n, s = offset_cosine_schedule(t)
        eps = latent_denoiser_forward(sd, z_t, t)
        z_0 = (z_t - n.view(-1, 1) * eps) / s.view(-1, 1)
        if i > 0:
            n_p, s_p = offset_cosine_schedule(torch.ones(B) * (i - 1) / num_steps)
            coefficient = torch.sqrt(n_p / n)
            z_t = s_p.view(-1, 1) * z_0 + coefficient.view(-1, 1) * n.view(-1, 1) * noises[j]
            j += 1
        else:
            z_t = z_0
    return vae_decode(sd, z_t, num_points) if decode else z_t


# ----------------------------------------------------------------------------------------
# FoldingDecoder (PointNetVAE.decode, networks.py:386-412, 1449-1509, 1579-1589) -- SURVEY 8(f) rank 2
# ----------------------------------------------------------------------------------------
def folding_state_dict_spec(latent_dim: int = 256, num_points: int = 2048, prefix: str = "vae.decoder"):
    """(key, shape) of FoldingDecoder's parameters.  FoldingLayer(cin, cout) = Conv1d(cin, cout, 1), ReLU,
    Conv1d(cout, cout, 1) (networks.py:396-400) -- there is NO activation between consecutive FoldingLayers."""
    spec = []
    for fold, cin in (("fold1", latent_dim + 2), ("fold2", latent_dim + 3)):
        for i, (ci, co) in enumerate(((cin, 512), (512, 512), (512, 3))):
            spec.append((f"{prefix}.{fold}.{i}.layer.0.weight", (co, ci, 1)))
            spec.append((f"{prefix}.{fold}.{i}.layer.0.bias", (co,)))
            spec.append((f"{prefix}.{fold}.{i}.layer.2.weight", (co, co, 1)))
            spec.append((f"{prefix}.{fold}.{i}.layer.2.bias", (co,)))
    spec.append((f"{prefix}.upsample.weight", (num_points, 1024)))
    spec.append((f"{prefix}.upsample.bias", (num_points,)))
    return spec


def make_synthetic_folding_checkpoint(seed: int = 31, latent_dim: int = 256, num_points: int = 2048,
                                      prefix: str = "vae.decoder") -> SD:
    """Seeded FoldingDecoder weights: PyTorch-default-like uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) with non-zero biases."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in folding_state_dict_spec(latent_dim, num_points, prefix):
        fan_in = shape[1] if len(shape) > 1 else None
        if fan_in is None:
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        else:
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
    return sd


def folding_grid() -> torch.Tensor:
    """networks.py:1463-1467: 32 x 32 grid on [-1, 1]^2, 'ij' meshgrid, as [2, 1024]."""
    r = torch.linspace(-1, 1, 32)
    xc, yc = torch.meshgrid(r, r, indexing="ij")
    return torch.stack([xc, yc], dim=-1).view(-1, 2).transpose(0, 1)


def _folding_layer(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return _conv(sd, f"{name}.layer.2", F.relu(_conv(sd, f"{name}.layer.0", x)))


def folding_decode(sd: SD, z: torch.Tensor, prefix: str = "vae.decoder") -> torch.Tensor:
    """FoldingDecoder.forward, networks.py:1484-1509: z [B, latent] -> [B, num_points, 3]."""
    B = z.size(0)
    grid = folding_grid().unsqueeze(0).repeat(B, 1, 1)                     # :1495
    zz = z.unsqueeze(2).repeat(1, 1, grid.size(2))                         # :1496
    h = torch.cat([zz, grid], dim=1)                                       # :1499
    for i in range(3):
        h = _folding_layer(sd, f"{prefix}.fold1.{i}", h)                   # :1500
    h = torch.cat([zz, h], dim=1)                                          # :1503
    for i in range(3):
        h = _folding_layer(sd, f"{prefix}.fold2.{i}", h)                   # :1504  [B, 3, 1024]
    # :1507-1508: Linear(1024 -> num_points) ACROSS the point axis, then [B, num_points, 3]
    return F.linear(h, sd[f"{prefix}.upsample.weight"], sd[f"{prefix}.upsample.bias"]).transpose(1, 2)


# ----------------------------------------------------------------------------------------
# VAE3DLarge.decode (networks.py:471-505, 2247-2264, 2327-2339) + voxel -> points glue (utils.py:511-539)
# -- the decoder of the reference's DEFAULT latent-diffusion configuration (is_voxel_based=True)
# ----------------------------------------------------------------------------------------
# (index in nn.Sequential, kind, cin, cout)
VAE3D_DECODER = [(0, "convT", 512, 256), (2, "res", 256, 256), (3, "convT", 256, 128), (5, "res", 128, 128),
                 (6, "convT", 128, 64), (8, "res", 64, 64), (9, "conv", 64, 32), (11, "res", 32, 32), (12, "conv", 32, 1)]


def vae3d_decoder_state_dict_spec(latent_dim: int = 256, prefix: str = "vae"):
    """(key, shape, kind) of the decoder half of VAE3DLarge's state_dict."""
    spec = [(f"{prefix}.decoder_input.weight", (512 * 64, latent_dim), "w"), (f"{prefix}.decoder_input.bias", (512 * 64,), "b")]
    for idx, kind, ci, co in VAE3D_DECODER:
        name = f"{prefix}.decoder.{idx}"
        if kind == "convT":      # nn.ConvTranspose3d weight is [cin, cout, 4, 4, 4]
            spec += [(f"{name}.weight", (ci, co, 4, 4, 4), "w"), (f"{name}.bias", (co,), "b")]
        elif kind == "conv":
            spec += [(f"{name}.weight", (co, ci, 3, 3, 3), "w"), (f"{name}.bias", (co,), "b")]
        else:                    # ResidualBlock3D(c, c): conv1, bn1, conv2, bn2 (no downsample when cin == cout)
            for j in (1, 2):
                spec += [(f"{name}.conv{j}.weight", (co, ci, 3, 3, 3), "w"), (f"{name}.conv{j}.bias", (co,), "b"),
                         (f"{name}.bn{j}.weight", (co,), "bn_w"), (f"{name}.bn{j}.bias", (co,), "bn_b"),
                         (f"{name}.bn{j}.running_mean", (co,), "bn_m"), (f"{name}.bn{j}.running_var", (co,), "bn_v"),
                         (f"{name}.bn{j}.num_batches_tracked", (), "nbt")]
    return spec


def make_synthetic_vae3d_decoder_checkpoint(seed: int = 41, latent_dim: int = 256, prefix: str = "vae", out_gain: float = 1.0) -> SD:
    """Seeded decoder weights with variance-preserving scales and randomised BatchNorm statistics (a default-init BN is
    an identity and would hide folding bugs).  The last conv is scaled so voxel logits spread over the sigmoid."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind in vae3d_decoder_state_dict_spec(latent_dim, prefix):
        if kind == "w":
            if len(shape) == 2:
                fan = shape[1]
            elif key.endswith(("decoder.0.weight", "decoder.3.weight", "decoder.6.weight")):
                fan = shape[0] * 8          # transposed conv k4 s2: 8 taps reach an output voxel
            else:
                fan = shape[1] * 27
            sd[key] = torch.randn(shape, generator=g) * (2.0 / fan) ** 0.5
        elif kind == "b":
            sd[key] = torch.randn(shape, generator=g) * 0.05
        elif kind == "bn_w":
            sd[key] = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif kind == "bn_b":
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_m":
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_v":
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        else:
            sd[key] = torch.tensor(0, dtype=torch.int64)
    sd[f"{prefix}.decoder.12.weight"] = sd[f"{prefix}.decoder.12.weight"] * out_gain
    return sd


def _bn3d_eval(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(x, sd[f"{name}.running_mean"], sd[f"{name}.running_var"], sd[f"{name}.weight"], sd[f"{name}.bias"],
                        False, 0.0, BN_EPS)


def _res_block3d(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """ResidualBlock3D.forward, networks.py:488-505 (cin == cout: identity shortcut)."""
    out = F.relu(_bn3d_eval(sd, f"{name}.bn1", F.conv3d(x, sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"], padding=1)))
    out = _bn3d_eval(sd, f"{name}.bn2", F.conv3d(out, sd[f"{name}.conv2.weight"], sd[f"{name}.conv2.bias"], padding=1))
    return F.relu(out + x)


def vae3d_decode(sd: SD, z: torch.Tensor, prefix: str = "vae", taps: Optional[dict] = None) -> torch.Tensor:
    """VAE3DLarge.decode, networks.py:2327-2339: z [B, 256] -> voxel probabilities [B, 1, 32, 32, 32]."""
    h = F.linear(z, sd[f"{prefix}.decoder_input.weight"], sd[f"{prefix}.decoder_input.bias"]).view(-1, 512, 4, 4, 4)
    for idx, kind, ci, co in VAE3D_DECODER:
        name = f"{prefix}.decoder.{idx}"
        if kind == "convT":
            h = F.relu(F.conv_transpose3d(h, sd[f"{name}.weight"], sd[f"{name}.bias"], stride=2, padding=1))
        elif kind == "res":
            h = _res_block3d(sd, name, h)
        elif idx == 9:
            h = F.relu(F.conv3d(h, sd[f"{name}.weight"], sd[f"{name}.bias"], padding=1))
        else:
            h = torch.sigmoid(F.conv3d(h, sd[f"{name}.weight"], sd[f"{name}.bias"], padding=1))
        if taps is not None:
            taps[idx] = h
    return h


def voxel_tensor_to_point_clouds(voxel_grid: torch.Tensor, threshold: float = 0.5):
    """utils.py:511-539: per sample, coordinates (x, y, z) of voxels above `threshold`, normalised to [-1, 1]."""
    _, _, depth, height, width = voxel_grid.shape
    out = []
    for i in range(voxel_grid.shape[0]):
        zz, yy, xx = torch.where(voxel_grid[i, 0] > threshold)
        if len(zz) > 0:
            pts = torch.stack([xx, yy, zz], dim=1).float()
            pts = 2 * pts / torch.tensor([width - 1, height - 1, depth - 1]) - 1
        else:
            pts = torch.empty((0, 3))
        out.append(pts)
    return out
