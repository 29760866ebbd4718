"""GPU diagnostics: each section runs in its own process (a trapped kernel poisons the CUDA
context) and prints error tables; output is meant for gpurun_out/."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sec_linear():
    import torch
    import pcd_b200
    for (M, K0, K1, Cout) in [(128, 64, 0, 64), (128, 128, 0, 64), (256, 64, 0, 128), (128, 64, 0, 256),
                              (1024, 256, 0, 512), (512, 512, 512, 512), (256, 2048, 0, 4096)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        a0 = torch.randn(M, K0, device="cuda", generator=g).bfloat16()
        a1 = torch.randn(M, K1, device="cuda", generator=g).bfloat16() if K1 else None
        w = (torch.randn(Cout, K0 + K1, device="cuda", generator=g) / (K0 + K1) ** 0.5).bfloat16()
        bias = torch.randn(Cout, device="cuda", generator=g)
        out = pcd_b200._lib.linear_bf16(a0, w, bias, a1, True)
        torch.cuda.synchronize()
        a = a0.float() if a1 is None else torch.cat([a0.float(), a1.float()], 1)
        ref = torch.relu(a @ w.float().t() + bias)
        err = float((out.float() - ref).norm() / ref.norm())
        print(f"linear M={M} K={K0}+{K1} Cout={Cout}: rel-L2 {err:.3e} max-abs {float((out.float()-ref).abs().max()):.3e}", flush=True)
        if err > 1e-2:
            # show structure of the error: per 32-column chunk / per row-group
            e = (out.float() - ref).abs()
            print("   col-chunk mean err:", [round(float(e[:, c:c + 32].mean()), 3) for c in range(0, min(Cout, 256), 32)])
            print("   row-group mean err:", [round(float(e[r:r + 8].mean()), 3) for r in range(0, 64, 8)])


def sec_forward(precision):
    import torch
    import pcd_b200
    from oracle import pointdiff_oracle as O
    os.environ["PCD_TAPS"] = "1"
    sd = O.make_synthetic_checkpoint()
    m = pcd_b200.PointCloudDiffusion(256, precision=precision)
    m.load_state_dict(sd)
    m = m.eval().cuda()
    g = torch.Generator().manual_seed(21)
    B, N = 2, 384
    x, t = torch.randn(B, N, 3, generator=g), torch.tensor([0.25, 0.8])
    taps = {}
    ref = O.denoiser_forward(sd, x, t, taps=taps)
    eps = m.model(x.cuda(), t.cuda())
    torch.cuda.synchronize()
    eng = m.model.engine()

    def rl(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())
    print(f"[{precision}] temb {rl(eng.tap('temb', (B, 256)), taps['temb']):.3e}")
    for name, C in (("x1", 128), ("x2", 256), ("x3", 512), ("x4", 1024), ("d4", 512), ("d1", 64)):
        print(f"[{precision}] {name} {rl(eng.tap(name, (B, N, C)), taps[name].transpose(1, 2)):.3e}")
    print(f"[{precision}] g {rl(eng.tap('g', (B, 4096)), taps['g']):.3e}")
    print(f"[{precision}] eps {rl(eps.cpu(), ref):.3e}", flush=True)


def sec_samplers(precision):
    """Final-sample error of the three samplers against the reference's golden vectors."""
    import torch
    import pcd_b200
    from oracle import pointdiff_oracle as O
    golden = torch.load(os.path.join(ROOT, "tests", "golden", "pointdiff_golden.pt"), weights_only=True)
    sd = O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 3300.0)
    m = pcd_b200.PointCloudDiffusion(256, precision=precision)
    m.load_state_dict(sd)
    m = m.eval().cuda()

    def rl(a, b):
        return float((a.double().cpu() - b.double()).norm() / b.double().norm())

    def cd(a, b):
        return float(pcd_b200.metrics.chamfer_distance(a.cuda(), b.cuda()))
    S, xT = int(golden["a3300.ddim.S"]), golden["a3300.ddim.xT"]
    out = m.sample(2, 256, num_steps=S, x_T=xT)
    print(f"[{precision}] ddim-{S}: rel-L2 {rl(out, golden['a3300.ddim.out']):.3e} CD {cd(out, golden['a3300.ddim.out']):.4f}")
    out = m.sample2(2, 256, num_steps=S, x_T=xT, noise=golden["a3300.ddpm.noise"])
    print(f"[{precision}] ddpm-{S}: rel-L2 {rl(out, golden['a3300.ddpm.out']):.3e} CD {cd(out, golden['a3300.ddpm.out']):.4f}")
    out = m.sample3(2, 256, x=golden["a3300.ddim3.x"], start_t=golden["a3300.ddim3.start_t"], num_steps=5)
    print(f"[{precision}] ddim3-5: rel-L2 {rl(out, golden['a3300.ddim3.out']):.3e} CD {cd(out, golden['a3300.ddim3.out']):.4f}", flush=True)


def sec_e2e(precision="bf16", B=512):
    """Per-call wall time of the host-buffer entry (pcd_sample_host) next to the device-resident loop."""
    import time
    import torch
    import pcd_b200
    from oracle import pointdiff_oracle as O
    sd = O.make_synthetic_checkpoint(seed=24, alpha=1.0 / 3300.0)
    m = pcd_b200.PointCloudDiffusion(2048, precision=precision)
    m.load_state_dict(sd)
    m = m.eval().cuda()
    eng = m.model.engine()
    table = m.ddim_table(50)
    xh = torch.randn(B, 2048, 3).pin_memory()
    oh = torch.empty_like(xh).pin_memory()
    xd = xh.cuda()
    for it in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = xd.clone(); eng.sample_(table, x, seed=5)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        eng.sample_host(table, xh, oh, seed=5)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"iter {it}: device-resident loop {1e3 * (t1 - t0):8.1f} ms   host-buffer call {1e3 * (t2 - t1):8.1f} ms", flush=True)


def sec_profile(batches=(4, 64, 512), precision="bf16"):
    import torch
    import pcd_b200
    from oracle import pointdiff_oracle as O
    sd = O.make_synthetic_checkpoint()
    m = pcd_b200.PointCloudDiffusion(2048, precision=precision)
    m.load_state_dict(sd)
    m = m.eval().cuda()
    for B in batches:
        x, t = torch.randn(B, 2048, 3, device="cuda"), torch.full((B,), 0.5, device="cuda")
        eng = m.model.engine()
        eng.profile(x, t)
        rows = eng.profile(x, t)
        tot = sum(r[1] for r in rows)
        fl = sum(r[2] for r in rows)
        print(f"--- B={B}: step {tot:.3f} ms, {fl / tot / 1e9:.1f} TFLOP/s algorithmic")
        for name, ms, f in rows:
            print(f"   {name:28s} {ms:9.4f} ms  {f / max(ms, 1e-9) / 1e9:9.1f} TFLOP/s")
        sys.stdout.flush()


if __name__ == "__main__":
    if len(sys.argv) > 2:     # e.g. `fwd f16mix`, `samplers f16`, `profile512 f16mix`, `profile256 bf16x3`
        sec, prec = sys.argv[1], sys.argv[2]
        if sec == "fwd":
            sec_forward(prec)
        elif sec == "samplers":
            sec_samplers(prec)
        elif sec == "e2e":
            sec_e2e(prec)
        elif sec.startswith("profile"):
            sec_profile((int(sec[len("profile"):] or 512),), prec)
        else:
            raise SystemExit(f"unknown section {sec}")
    elif len(sys.argv) > 1:
        {"linear": sec_linear, "fwd32": lambda: sec_forward("fp32"), "fwd16": lambda: sec_forward("bf16"), "fwdx3": lambda: sec_forward("bf16x3"),
         "profile": sec_profile, "profile64": lambda: sec_profile((64,)), "profile512": lambda: sec_profile((512,))}[sys.argv[1]]()
    else:
        for s in ("linear", "fwd32", "fwd16", "profile"):
            print(f"===== {s} =====", flush=True)
            r = subprocess.run([sys.executable, __file__, s], timeout=600)
            print(f"===== {s} exit {r.returncode} =====", flush=True)
