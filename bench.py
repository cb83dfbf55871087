#!/usr/bin/env python
"""bench.py — denoised video frames/s of ORV's DiT hot path on B200 (BASELINE.json metric, config 2).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W     (the reference algorithm on the host CPU cores)
    python bench.py --impl torch-eager --steps K --warmup W   (the reference forward in eager torch on cuda:0; the
                                                               default run calls this in a child process and reports
                                                               it as `torch_eager_gpu`)

A "step" is one clip: the full 50-iteration denoise loop (transformer forward + CFG/scheduler update per
iteration) over one synthetic 17-frame 320x480 clip (latents [1,5,16,40,60], text [1,226,4096], 16 actions).
`value`  = clips * 16 frames / time with every input resident in HBM (device-timed, max over ranks).
`e2e`    = the same metric through the public pipeline call with HOST inputs (pinned) -> H2D inside the timed
           region and the final latents read back D2H.
Random-init weights of the CogVideoX-2B ORV architecture (std 0.02), synthetic inputs; VAE decode and T5 are not
part of the metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_INFERENCE_STEPS = 50
FRAMES_PER_CLIP = 16  # BASELINE.json: "320x480x16"; the 16 actions generate 16 new frames after the reference frame
FWD_TFLOP = 10.968    # measured on the reference forward (SURVEY probe P1), config 2, per sequence


def config2() -> dict:
    return dict(num_attention_heads=30, attention_head_dim=64, in_channels=32, out_channels=16, num_layers=30,
                sample_width=60, sample_height=40, sample_frames=17, modulate_encoder_hidden_states=True,
                text_embed_dim=4096, max_text_seq_length=226, time_embed_dim=512, patch_size=2,
                loaded_pretrained_model_name_or_path="THUDM/CogVideoX-2b")


def peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d.get("bf16_tflops_sustained", 1422.1), bf16_burst=d.get("bf16_tflops", 1687.1),
                    hbm=d.get("hbm_gbs", 6464.9), source="measured")
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int = 0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self) -> dict:
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                continue
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_step_seconds(n_timed: int, n_warm: int, budget_s: float = 150.0):
    """Times transformer forwards of the config-2 workload with the CPU oracle (bf16, all host threads — the
    dtype the reference deploys in).  Returns (seconds per FULL forward, description, cores)."""
    from oracle import flat_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.default_config(**config2())
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.02, dtype=torch.bfloat16)
    inp = O.synthetic_inputs(cfg, 1, 5, 40, 60, seed=1)
    hs, text, act = inp["hidden_states"].bfloat16(), inp["text"].bfloat16(), inp["actions"].bfloat16()
    t = torch.tensor([999], dtype=torch.int64)
    layers_full = cfg["num_layers"]
    # probe with 2 layers to size the sample
    cfg_probe = dict(cfg, num_layers=2)
    with torch.no_grad():
        O.forward(sd, cfg_probe, hs, text, t, actions=act)
        t0 = time.perf_counter()
        O.forward(sd, cfg_probe, hs, text, t, actions=act)
        per2 = time.perf_counter() - t0
    per_layer = per2 / 2
    total_fwd = max(n_timed + n_warm, 1)
    layers = int(max(1, min(layers_full, budget_s / (per_layer * total_fwd))))
    cfg_s = dict(cfg, num_layers=layers)
    times = []
    with torch.no_grad():
        for i in range(n_warm + n_timed):
            t0 = time.perf_counter()
            O.forward(sd, cfg_s, hs, text, t, actions=act)
            dt = time.perf_counter() - t0
            if i >= n_warm:
                times.append(dt)
    mean = sum(times) / len(times)
    full = mean * layers_full / layers if layers < layers_full else mean
    sample = (f"{n_timed} forward(s) of the config-2 batch (1 clip, S=3226 tokens), oracle port in torch-CPU bf16, "
              f"{layers}/{layers_full} transformer layers per forward"
              + (" scaled linearly to 30" if layers < layers_full else "")
              + f"; frames/s = 16 / (50 forwards); {cores} threads")
    return full, sample, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, sample, cores = cpu_reference_step_seconds(args.steps, args.warmup)
    clip_s = sec * NUM_INFERENCE_STEPS
    value = FRAMES_PER_CLIP / clip_s
    line = {
        "impl": "reference", "metric": "denoised_video_frames_per_sec", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": clip_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "config2: CogVideoX-2B ORV DiT, 17x320x480 clip (latents 5x40x60), B=1, 50 steps, "
                               "actions on, no CFG", "note": "step time = 50 x measured CPU forward"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# second baseline leg: the reference's DEPLOYMENT path — eager PyTorch on the same GPU (SURVEY §8d "beat this")
# ----------------------------------------------------------------------------------------------------------------
def run_torch_eager(args, device: str = "cuda:0", cfg_over: dict | None = None):
    """Times the restated reference forward (oracle/flat_oracle.py) executed by eager torch in bf16 on `device`:
    cuBLAS GEMMs + SDPA(flash) + the ~60 elementwise kernels per block the reference launches.  A baseline leg like
    `cpu_baseline` — never the product — run in its own process by `torch_eager_gpu_leg`."""
    from oracle import flat_oracle as O
    dev = torch.device(device)
    cfg = O.default_config(**dict(config2(), **(cfg_over or {})))
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {}
    for name, shape in O.param_shapes(cfg).items():
        t = torch.randn(shape, generator=g, device=dev) * 0.02
        if name.endswith(".weight") and len(shape) == 1:
            t = 1.0 + t
        sd[name] = t.to(torch.bfloat16)
    lat_h, lat_w = cfg["sample_height"], cfg["sample_width"]
    inp = O.synthetic_inputs(cfg, 1, 5, lat_h, lat_w, seed=1)
    hs, text, act = (inp[k].to(dev, torch.bfloat16) for k in ("hidden_states", "text", "actions"))
    t = torch.tensor([999], dtype=torch.int64, device=dev)
    cuda = dev.type == "cuda"
    steps, warm = max(args.steps, 1), max(args.warmup, 1)
    with torch.no_grad():
        for _ in range(warm):
            O.forward(sd, cfg, hs, text, t, actions=act)
        if cuda:
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            O.forward(sd, cfg, hs, text, t, actions=act)
        if cuda:
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
        else:
            ms = (time.perf_counter() - t0) * 1e3 / steps
    line = {"impl": "torch-eager", "metric": "denoised_video_frames_per_sec",
            "value": FRAMES_PER_CLIP / (NUM_INFERENCE_STEPS * ms * 1e-3), "unit": "frames/s", "ms_per_forward": ms,
            "forwards_timed": steps, "device": str(dev),
            "sample": "config-2 transformer forward (1 clip, S=3226 tokens, 30 layers), restated reference forward in "
                      "eager torch bf16 (cuBLAS + SDPA); frames/s = 16 / (50 forwards), sampler arithmetic excluded"}
    print(json.dumps(line), flush=True)
    return line


def torch_eager_gpu_leg(timeout_s: float = 90.0) -> dict:
    """Runs `bench.py --impl torch-eager` in a child process (a failure there cannot take the bench line with it)."""
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "torch-eager", "--steps", "5",
                              "--warmup", "2"], capture_output=True, text=True, timeout=timeout_s)
        lines = [x for x in res.stdout.splitlines() if x.startswith("{")]
        if res.returncode == 0 and lines:
            out = json.loads(lines[-1])
            out.pop("impl", None)
            out.pop("metric", None)
            return out
        return {"unavailable": ((res.stderr or "") + (res.stdout or ""))[-300:].strip() or f"exit code {res.returncode}"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:300]}


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def init_weights_(model, seed: int, std: float = 0.02):
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1 and name.endswith("weight"):
                p.copy_(1.0 + torch.randn(p.shape, generator=g, device="cuda") * std)
            else:
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") * std)


def run_ours(args):
    from orv_b200 import (CogVideoXDPMScheduler, CogVideoXImageToVideoPipelineTraj,
                          CogVideoXTransformer3DModelTraj, _lib as L)
    from orv_b200 import dist as D
    from orv_b200.models.pipeline_control import default_vae_config
    rank, local_rank, world = D.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = config2()
    with torch.device(dev):
        model = CogVideoXTransformer3DModelTraj(**cfg)
    if rank == 0:
        init_weights_(model, seed=0)
    model = model.to(torch.bfloat16).eval()
    model.action_embed.mask = False  # parity/bench runs pin the reference's stray action dropout off
    D.broadcast_weights(model, src=0)  # the single collective of the path
    arena_bytes = model.weight_arena().numel() * 2

    sched = CogVideoXDPMScheduler(timestep_spacing="trailing")
    pipe = CogVideoXImageToVideoPipelineTraj(None, None, default_vae_config(), model, sched)

    # ---- synthetic inputs, host-pinned masters + device-resident copies ----
    g = torch.Generator().manual_seed(1 + rank)
    image_h = torch.randn(1, 32, 1, 40, 60, generator=g).bfloat16().pin_memory()           # first-frame VAE moments
    text_h = (torch.randn(1, 226, 4096, generator=g) * 0.2).bfloat16().pin_memory()
    act_h = ((torch.rand(1, 16, 7, generator=g) * 2 - 1) * torch.tensor([20.0] * 6 + [1.0])).bfloat16().pin_memory()
    image_d, text_d, act_d = image_h.to(dev), text_h.to(dev), act_h.to(dev)

    def run_clip(image, text, act, seed):
        gen = torch.Generator().manual_seed(seed)
        # prompt AND prompt_embeds, as the reference programs call it (evaluation_control_to_video.py:321-322):
        # check_inputs is invoked positionally in the reference, so prompt_embeds alone raises (SURVEY probe P7)
        out = pipe(image=image, prompt="", prompt_embeds=text, height=320, width=480, num_frames=17,
                   num_inference_steps=NUM_INFERENCE_STEPS, guidance_scale=1.0, generator=gen,
                   controls_or_guidances={"actions": act}, output_type="latent", return_dict=False)[0]
        return out

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        return D.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)

    # ---- device-resident throughput ----
    for i in range(args.warmup):
        run_clip(image_d, text_d, act_d, 42 + i)
    with ClockSampler(local_rank) as clocks:
        secs = timed(lambda i: run_clip(image_d, text_d, act_d, 100 + i), args.steps)
    launches_per_clip = pipe.last_step_launches
    value = world * args.steps * FRAMES_PER_CLIP / secs

    # ---- end-to-end: host inputs in, latents out ----
    def e2e_clip(i):
        lat = run_clip(image_h.to(dev, non_blocking=True), text_h.to(dev, non_blocking=True),
                       act_h.to(dev, non_blocking=True), 200 + i)
        return lat.cpu()

    e2e_clip(0)
    e2e_secs = timed(e2e_clip, args.steps)
    e2e_value = world * args.steps * FRAMES_PER_CLIP / e2e_secs
    h2d = image_h.numel() * 2 + text_h.numel() * 2 + act_h.numel() * 2 + NUM_INFERENCE_STEPS * 5 * 16 * 40 * 60 * 2 \
        + 5 * 16 * 40 * 60 * 2  # + host-generated DPM noise per step + initial noise (CPU generator contract)
    d2h = 5 * 16 * 40 * 60 * 2

    # ---- per-kernel-class timing inside the real step (CUDA events on the launch stream) ----
    model.set_profile(True)
    hs = torch.randn(1, 5, 32, 40, 60, device=dev).bfloat16()
    tt = torch.full((1,), 499, device=dev, dtype=torch.int64)
    nprof = 3
    with torch.no_grad():
        for _ in range(nprof):
            model(hs, text_d, {"actions": act_d}, tt, return_dict=False)
    ms, cnt = model.get_profile()
    model.set_profile(False)
    names = ["prologue_adaln", "embed", "ln_modulate", "gemm_qkv", "attention", "gemm_out", "gemm_ff1", "gemm_ff2",
             "head"]
    S, Dm, FF, H = 3226, 1920, 7680, 30
    flops = {"gemm_qkv": 2.0 * S * 3 * Dm * Dm, "attention": 4.0 * S * S * Dm, "gemm_out": 2.0 * S * Dm * Dm,
             "gemm_ff1": 2.0 * S * FF * Dm, "gemm_ff2": 2.0 * S * Dm * FF}
    pk = peaks()
    kernels = {}
    for i, n in enumerate(names):
        if cnt[i] == 0:
            continue
        per_launch_us = ms[i] / cnt[i] * 1e3
        ent = {"us_per_launch": round(per_launch_us, 2), "launches_per_forward": cnt[i] // nprof,
               "ms_per_forward": round(ms[i] / nprof, 4)}
        if n in flops:
            ent["tflops"] = round(flops[n] / (per_launch_us * 1e-6) / 1e12, 1)
            ent["frac_of_peak"] = round(ent["tflops"] / pk["bf16"], 4)
        kernels[n] = ent
    fwd_ms = sum(ms) / nprof
    dom = max(flops.keys(), key=lambda n: kernels[n]["ms_per_forward"])
    # dram bytes per launch of the dominant kernel, from the committed ncu capture of the same build (profiles/)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        key = {"attention": "attention_kernel", "gemm_qkv": "gemm2_bf16_kernel<3>", "gemm_ff1": "gemm2_bf16_kernel<1>",
               "gemm_ff2": "gemm2_bf16_kernel<2>", "gemm_out": "gemm2_bf16_kernel<2>"}[dom]
        for kname, ent in tj.items():
            if key in kname and ent:
                traffic = ent[0]["dram_bytes"]
                break
    roofline = {"bound": "tensor", "kernel": dom, "achieved": kernels[dom]["tflops"], "peak": pk["bf16"],
                "unit": "TFLOP/s", "frac": round(kernels[dom]["tflops"] / pk["bf16"], 4), "traffic": traffic,
                "algorithmic_flop_per_launch": flops[dom],
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['source']})",
                "forward_ms_sum_of_kernels": round(fwd_ms, 3),
                "forward_tensor_frac": round(FWD_TFLOP / (fwd_ms * 1e-3) / pk["bf16"], 4)}

    clk = clocks.summary()
    line = {
        "metric": "denoised_video_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "config2: CogVideoX-2B ORV DiT (D=1920,H=30,L=30), 17x320x480 clip (latents 5x40x60, "
                               "S=3226 tokens), B=1 clip per GPU per step, 50 DPM-trailing iterations, actions on, "
                               "no CFG (guidance 1.0)",
                   "clips_per_gpu_per_step": 1, "iterations_per_clip": NUM_INFERENCE_STEPS,
                   "l2": "weights (%.2f GB per forward) stream from HBM every iteration, far above the 126 MB L2; "
                         "no explicit flush" % (arena_bytes / 1e9),
                   "parallelism": f"dp{world} (independent clips, one NCCL weight broadcast at init)"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_secs / args.steps * 1e3},
        "gpu_launches": int(launches_per_clip * args.steps),
        "tensor_frac_of_peak": round(NUM_INFERENCE_STEPS * FWD_TFLOP * world * args.steps / secs / (pk["bf16"] * world), 4),
        "tensor_frac_of_burst_peak": round(NUM_INFERENCE_STEPS * FWD_TFLOP * world * args.steps / secs
                                           / (pk["bf16_burst"] * world), 4),
        "roofline": roofline, "kernels": kernels, "clocks": clk,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, sample, cores = cpu_reference_step_seconds(1, 0, budget_s=25.0)
        line["cpu_baseline"] = {"value": FRAMES_PER_CLIP / (sec * NUM_INFERENCE_STEPS), "unit": "frames/s",
                                "cores": cores, "kind": "port", "sample": sample}
        # what the reference actually deploys: the same forward in eager torch on this GPU (own process)
        line["torch_eager_gpu"] = torch_eager_gpu_leg()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-eager":
        run_torch_eager(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
