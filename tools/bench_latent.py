"""BASELINE config 4: latent DDIM-50 + SimplePointNetVAE decode to 2048 points.
Reports shapes/s (device-timed, z_T resident), us per reverse step, the algorithmic weight bytes per
step (this path is weight-bandwidth / latency bound) and a CPU baseline (oracle port).
Under torchrun (one rank per GPU) every rank samples its own shard of the batch (no collective in the loop); the
reported time is the max over ranks and shapes/s is the whole job's."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128     # 1024 latents over 8 GPUs
    S, NP = 50, 2048
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from oracle import pointdiff_oracle as O   # synthetic weights + CPU baseline only
    sd = O.make_synthetic_latent_checkpoint(num_points=NP)
    m = pcd_b200.LatentDiffusion(pcd_b200.SimplePointNetVAE(NP), is_voxel_based=False)
    m.load_state_dict(sd, strict=False)
    m = m.eval().cuda()
    zT = torch.randn(B, 256, generator=torch.Generator().manual_seed(5 + rank)).cuda()
    for _ in range(3):
        out = m.sample(B, num_steps=S, z_T=zT, sample_offset=rank * B)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    reps = 5
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(reps):
        z0 = m.sample(B, num_steps=S, z_T=zT, sample_offset=rank * B, return_latent=True)
    e[1].record()
    for _ in range(reps):
        out = m.engine().decode(z0)
    e[2].record()
    torch.cuda.synchronize()
    loop_ms, dec_ms = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
    if world > 1:
        tmax = torch.tensor([loop_ms, dec_ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        loop_ms, dec_ms = float(tmax[0]), float(tmax[1])
    n_model = sum(v.numel() for k, v in sd.items() if k.startswith("model."))
    n_vae = sum(v.numel() for k, v in sd.items() if k.startswith("vae."))
    res = {"workload": f"latent DDIM-{S} + VAE decode -> {NP} pts, batch {B} per GPU, fp32-class (3xTF32)", "n_gpus": world,
           "shapes_per_s": world * B / (loop_ms + dec_ms) * 1e3, "timing": "CUDA events, max over ranks" if world > 1 else "CUDA events",
           "loop_ms": loop_ms, "us_per_reverse_step": loop_ms / S * 1e3, "decode_ms": dec_ms,
           "weight_bytes_per_step_fp32": 4 * n_model, "weight_stream_gbs": 4 * n_model / (loop_ms / S * 1e-3) / 1e9,
           "decoder_weight_bytes_fp32": 4 * n_vae, "algorithmic_flops_per_sample_step": 38_174_720,
           "tflops_loop": 38_174_720 * B * world * S / (loop_ms * 1e-3) / 1e12, "finite": bool(torch.isfinite(out).all())}
    if rank != 0:
        dist.destroy_process_group()
        return
    # CPU baseline (bounded): oracle latent loop, 5 steps, same batch; decode once
    zc = zT.cpu()
    O.latent_ddim_sample(sd, zc[:8], 1, NP, decode=False)
    t0 = time.perf_counter(); O.latent_ddim_sample(sd, zc, 5, NP, decode=False); t_loop = (time.perf_counter() - t0) / 5 * S
    t0 = time.perf_counter(); O.vae_decode(sd, zc, NP); t_dec = time.perf_counter() - t0
    res["cpu_baseline"] = {"shapes_per_s": B / (t_loop + t_dec), "kind": "port", "cores": torch.get_num_threads(),
                           "sample": f"5 of {S} reverse steps (extrapolated) + 1 decode, batch {B}"}
    print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
