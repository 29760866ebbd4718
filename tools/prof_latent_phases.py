"""Per-phase timing of the persistent latent kernel: run the reverse loop with the program truncated after k phases
(PCD_LT_MAXOPS, read when the plan's program is built) and difference the per-step times.  Results are garbage numerically
(the update phase is cut off); only the timing is meaningful."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = ["emb", "time_mlp.0", "time_mlp.2", "enc1.temb_bias", "enc1 G", "enc1 N", "enc2 G", "enc2 N", "enc3 G", "enc3 N", "enc4 G", "enc4 N",
         "gf0 G", "gf0 N", "gf3 G", "gf3 N", "dec4 G", "dec4 N", "dec3 G", "dec3 N", "dec2 G", "dec2 N", "dec1 G", "dec1 N", "out0", "out2+update"]

def child(k, B, S):
    import torch
    import pcd_b200
    from oracle import pointdiff_oracle as O
    sd = O.make_synthetic_latent_checkpoint(num_points=256)
    m = pcd_b200.LatentDiffusion(pcd_b200.SimplePointNetVAE(256), is_voxel_based=False)
    m.load_state_dict(sd, strict=False)
    m = m.eval().cuda()
    zT = torch.randn(B, 256, generator=torch.Generator().manual_seed(5)).cuda()
    for _ in range(2):
        m.sample(B, num_steps=S, z_T=zT, return_latent=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.sample(B, num_steps=S, z_T=zT, return_latent=True)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"k": k, "us_per_step": e0.elapsed_time(e1) / 5 / S * 1e3}))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    S = 50
    prev = None
    for k in [4] + list(range(5, 27)):
        env = dict(os.environ, PCD_LT_MAXOPS=str(k))
        out = subprocess.run([sys.executable, __file__, "child", str(k), str(B), str(S)], env=env, capture_output=True, text=True)
        try:
            r = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception:
            print("k", k, "failed", out.stderr[-300:]); continue
        d = None if prev is None else r["us_per_step"] - prev
        print(f"k={k:2d} last phase {NAMES[k-1]:16s} cumulative {r['us_per_step']:8.2f} us/step  delta {'' if d is None else f'{d:7.2f}'}", flush=True)
        prev = r["us_per_step"]
