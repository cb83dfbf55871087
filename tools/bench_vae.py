"""3-D VAE decode on one B200: the reference's deployment call — 5 latent frames of 40 x 60 (17 frames of
320 x 480), tiling + slicing on, bf16 — through orv_b200.AutoencoderKLCogVideoX, next to the oracle restatement run by
eager torch on the same GPU in bf16 (cuDNN conv3d: what diffusers executes for the reference).

    python tools/bench_vae.py [--steps 5] [--warmup 2] [--profile] [--skip-eager] [--latent 5 40 60]

Prints one JSON line: ms per decode, algorithmic TFLOP (convolutions only, 2 * pixels * taps * c_in * c_out), fraction of
the sustained bf16 peak, per-kernel-class device times (CUPTI, --profile) and the eager-torch time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def conv_flops(vae, T, h, w) -> float:
    return vae.decode_conv_flops(T, h, w)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--latent", type=int, nargs=3, default=[5, 40, 60])
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--skip-eager", action="store_true")
    args = ap.parse_args()
    from oracle import vae_oracle as V
    from orv_b200 import AutoencoderKLCogVideoX
    dev = "cuda:0"
    cfg = V.default_config()
    sd = V.synthetic_state_dict(cfg, seed=1)
    vae = AutoencoderKLCogVideoX().eval()
    vae.load_state_dict(sd)
    vae = vae.to(dev, torch.bfloat16)
    vae.enable_slicing()
    vae.enable_tiling()
    T, h, w = args.latent
    z = torch.randn(1, 16, T, h, w, device=dev).bfloat16()
    flops = conv_flops(vae, T, h, w)
    for _ in range(args.warmup):
        out = vae.decode(z).sample
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(args.steps):
        out = vae.decode(z).sample
        ev[i + 1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    med = ms[len(ms) // 2]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    line = {"what": "vae_decode", "latent": [T, h, w], "frames": 4 * (T - 1) + 1, "out_shape": list(out.shape),
            "ms_per_decode": round(med, 3), "ms_min": round(ms[0], 3), "wall_ms": round(wall, 3),
            "launches": vae.last_launches, "conv_tflop": round(flops / 1e12, 3),
            "tflops": round(flops / med / 1e9, 1),
            "frac_of_sustained_bf16_peak": round(flops / med / 1e9 / peaks["bf16_tflops_sustained"], 4)}
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            vae.decode(z)
            torch.cuda.synchronize()
        agg = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                name = e.name.split("(")[0].split("<")[0].replace("orvb::", "")
                a = agg.setdefault(name, [0, 0.0])
                a[0] += 1
                a[1] += e.device_time
        top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]
        line["kernels"] = {k: {"launches": n, "ms": round(us / 1e3, 3)} for k, (n, us) in top}
    if not args.skip_eager:
        sdb = {k: v.to(dev, torch.bfloat16) for k, v in sd.items()}
        with torch.no_grad():
            for _ in range(1):
                ref = V.decode(sdb, cfg, z, tiling=True)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            n = max(1, args.steps // 2)
            for _ in range(n):
                ref = V.decode(sdb, cfg, z, tiling=True)
            b.record()
            torch.cuda.synchronize()
        line["torch_eager_bf16_ms"] = round(a.elapsed_time(b) / n, 3)
        line["speedup_vs_torch_eager"] = round(line["torch_eager_bf16_ms"] / med, 2)
        d = (out.float() - ref.float()).abs()
        line["vs_eager_bf16"] = {"mean_abs": float(d.mean()), "max_abs": float(d.max()), "ref_mean_abs": float(ref.float().abs().mean())}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
