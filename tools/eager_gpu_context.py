"""Context figure, not a bench arm: the reference ARCHITECTURE (the oracle's plain-torch restatement of UNetPointNetLarge.forward,
bit-identical to the reference on CPU) run in eager PyTorch on the SAME B200, so that the CPU-host baseline of bench.py can be read
next to what stock cuDNN / cuBLAS kernels give on this GPU.  One denoiser forward is timed (the sampler update is negligible) and
quoted as DDIM-50 shapes/s = batch / (50 x forward time).

    python tools/eager_gpu_context.py [batch]        # one JSON object per line; copy into profiles/

Settings: strict fp32 (TF32 off), torch's default (cuDNN convolutions may use TF32 -- what the reference gets on an Ampere+ GPU out
of the box), and bf16 autocast."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pointdiff_oracle as O  # noqa: E402  (dev tool: the oracle as the stand-in for the reference model)


def timed(fn, warm=2, reps=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    N = 2048
    sd = {k: v.cuda() for k, v in O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 33.0).items()}
    torch.set_default_device("cuda")       # the oracle builds its small constants (frequency table) on the default device
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(B, N, 3, device="cuda", generator=g)
    t = torch.full((B,), 0.5, device="cuda")
    with torch.no_grad():
        ref = None
        for name, tf32, amp in (("fp32_strict", False, False), ("torch_default_tf32_convs", True, False), ("bf16_autocast", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False

            def fwd():
                if amp:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return O.denoiser_forward(sd, x, t)
                return O.denoiser_forward(sd, x, t)

            out = fwd().float()
            if ref is None:
                ref = out
            ms = timed(fwd)
            print(json.dumps({"what": "eager torch forward of the reference architecture on this GPU (context, not a bench arm)",
                              "setting": name, "batch": B, "points": N, "forward_ms": ms,
                              "ddim50_shapes_per_s": B / (50 * ms) * 1e3,
                              "eps_rel_l2_vs_fp32_strict": float((out - ref).norm() / ref.norm()),
                              "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)


if __name__ == "__main__":
    main()
