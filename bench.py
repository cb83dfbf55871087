#!/usr/bin/env python
"""bench.py — denoised video frames/s of ORV's DiT hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --config {2,3,4,5} --clips-per-gpu B      (default: config 2, B = 1 — the configuration the
                                                               metric is quoted on; 3/4/5 are BASELINE.json's other
                                                               single-GPU-sized workloads, see CONFIGS)
    python bench.py --impl reference --steps K --warmup W     (the reference algorithm on the host CPU cores)
    python bench.py --impl torch-eager --steps K --warmup W   (the reference forward in eager torch on cuda:0; the
                                                               default run calls this in a child process and reports
                                                               it as `torch_eager_gpu`)

A "step" is one pipeline call: the full 50-iteration denoise loop (transformer forward + CFG/scheduler update per
iteration) over B synthetic clips per GPU.
`value`  = clips * frames per clip / time with every input resident in HBM (device-timed, max over ranks).
`e2e`    = the same metric through the public pipeline call with HOST inputs (pinned) -> H2D inside the timed
           region and the final latents read back D2H.
Per-kernel-class times come from device-side activity records (CUPTI through torch.profiler) of one pipeline call
replaying its captured CUDA graphs: for every kernel the completion-to-completion interval on the stream
(end - previous end), so host enqueue gaps cannot enter and the classes of a forward sum to its device span.
Random-init weights of the named architecture (std 0.02), synthetic inputs; VAE decode and T5 are not part of the
metric.
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

_BASE2B = dict(num_attention_heads=30, attention_head_dim=64, in_channels=32, out_channels=16, num_layers=30,
               modulate_encoder_hidden_states=True, text_embed_dim=4096, max_text_seq_length=226, time_embed_dim=512,
               patch_size=2)

# BASELINE.json `configs` (index = position in that list).  tflop = algorithmic forward TFLOP per clip and iteration
# (measured on the reference forward with FlopCounterMode; a CFG pair counts both sequences, a
# multiview clip all views).  frames = video frames one clip yields (16 new frames per view).
CONFIGS = {
    2: dict(name="config2: CogVideoX-2B ORV DiT (D=1920,H=30,L=30), 17x320x480 clip (latents 5x40x60, S=3226 tokens), "
                 "50 DPM-trailing iterations, actions on, no CFG (guidance 1.0)",
            model=dict(_BASE2B, sample_width=60, sample_height=40, sample_frames=17,
                       loaded_pretrained_model_name_or_path="THUDM/CogVideoX-2b"),
            px=(320, 480), views=1, controls=False, cfg_pair=False, tflop=10.968, frames=16, clips=1),
    3: dict(name="config3: CogVideoX-2B + occupancy condfull (depth + semantic-label latents), 17x320x480 clip, "
                 "50 DPM-trailing iterations, actions on, no CFG",
            model=dict(_BASE2B, sample_width=60, sample_height=40, sample_frames=17, visual_guidance=True,
                       num_control_blocks=2, loaded_pretrained_model_name_or_path="THUDM/CogVideoX-2b"),
            px=(320, 480), views=1, controls=True, cfg_pair=False, tflop=11.023, frames=16, clips=1),
    4: dict(name="config4: CogVideoX1.5-5B dims (D=3072,H=48,L=42, p_t=2, RoPE, ofs), 17x320x480 clip (6 latent frames "
                 "after p_t padding, S=2026), 50 DPM-trailing iterations, CFG pair per clip (guidance 6, caller-duplicated "
                 "actions)",
            model=dict(_BASE2B, num_attention_heads=48, num_layers=42, sample_width=60, sample_height=40,
                       sample_frames=21, patch_size_t=2, use_rotary_positional_embeddings=True, ofs_embed_dim=512,
                       patch_bias=False, loaded_pretrained_model_name_or_path="THUDM/CogVideoX1.5-5b-I2V"),
            px=(320, 480), views=1, controls=False, cfg_pair=True, tflop=2 * 21.404, frames=16, clips=1),
    5: dict(name="config5: CogVideoX-2B multiview (3 views, 30 temporal + 30 view blocks) + condfull, 17x256x384 per "
                 "view, 50 DPM-trailing iterations, actions on, no CFG",
            model=dict(_BASE2B, sample_width=48, sample_height=32, sample_frames=17, visual_guidance=True,
                       num_control_blocks=2, multiview=True, max_n_view=3,
                       loaded_pretrained_model_name_or_path="THUDM/CogVideoX-2b"),
            px=(256, 384), views=3, controls=True, cfg_pair=False, tflop=33.633, frames=48, clips=2),
}


def config2() -> dict:
    return dict(CONFIGS[2]["model"])


def workload_name(cid: int, clips: int) -> str:
    return CONFIGS[cid]["name"] + f", B={clips} clip(s) per GPU per step"


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
def _oracle_case(cid: int, clips: int, device="cpu", dtype=torch.bfloat16, seed_weights_on_device=False):
    """The config's transformer-forward inputs for the oracle (one iteration of the denoise loop)."""
    from oracle import flat_oracle as O
    c = CONFIGS[cid]
    cfg = O.default_config(**c["model"])
    V = c["views"]
    h, w = c["px"][0] // 8, c["px"][1] // 8
    pt = cfg["patch_size_t"] or 1
    F = -(-5 // pt) * pt
    B = clips * (2 if c["cfg_pair"] else 1)
    inp = O.synthetic_inputs(cfg, B, F * V, h, w, seed=1, with_controls=c["controls"], n_actions=16 if pt == 1 else 20)
    kw = dict(actions=inp["actions"].to(device, dtype))
    if c["controls"]:
        kw["depths"], kw["labels"] = inp["depths"].to(device, dtype), inp["labels"].to(device, dtype)
    if cfg["use_rotary_positional_embeddings"]:
        rp = O.pipeline_rope(cfg, c["px"][0], c["px"][1], F)
        kw["rope"] = (rp[0].to(device), rp[1].to(device))
    if cfg["ofs_embed_dim"] is not None:
        kw["ofs"] = torch.tensor([2.0], device=device)
    if V > 1:
        kw["num_views"] = V
    t = torch.full((B,), 999, dtype=torch.int64, device=device)
    return cfg, inp["hidden_states"].to(device, dtype), inp["text"].to(device, dtype), t, kw


def cpu_reference_forward_seconds(cid: int, clips: int, n_timed: int, n_warm: int, budget_s: float):
    """Times transformer forwards of the config's workload with the CPU oracle (bf16 — the dtype the reference
    deploys in —, all host threads).  The sample is bounded by the layer count (timed on `layers` of the model's
    layers and scaled linearly when the budget does not allow all of them).  Returns (seconds per FULL forward,
    description, cores)."""
    from oracle import flat_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, hs, text, t, kw = _oracle_case(cid, clips)
    sd = O.synthetic_state_dict(cfg, seed=0, std=0.02, dtype=torch.bfloat16)
    layers_full = cfg["num_layers"]
    cfg_probe = dict(cfg, num_layers=1)
    with torch.no_grad():
        O.forward(sd, cfg_probe, hs, text, t, **kw)
        t0 = time.perf_counter()
        O.forward(sd, cfg_probe, hs, text, t, **kw)
        per_layer = time.perf_counter() - t0
    total_fwd = max(n_timed + n_warm, 1)
    layers = int(max(1, min(layers_full, budget_s / (per_layer * total_fwd))))
    cfg_s = dict(cfg, num_layers=layers)
    times = []
    with torch.no_grad():
        for i in range(n_warm + n_timed):
            t0 = time.perf_counter()
            O.forward(sd, cfg_s, hs, text, t, **kw)
            dt = time.perf_counter() - t0
            if i >= n_warm:
                times.append(dt)
    mean = sum(times) / len(times)
    full = mean * layers_full / layers if layers < layers_full else mean
    sample = (f"{n_timed} timed forward(s) after {n_warm} warm-up of the config-{cid} batch ({clips} clip(s)), oracle port in "
              f"torch-CPU bf16, {layers}/{layers_full} transformer layers per forward"
              + (f" scaled linearly to {layers_full}" if layers < layers_full else "")
              + f"; frames/s = frames per clip / ({NUM_INFERENCE_STEPS} forwards); {cores} threads")
    return full, sample, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cid, clips = args.config, args.clips_per_gpu or CONFIGS[args.config]["clips"]
    sec, sample, cores = cpu_reference_forward_seconds(cid, clips, max(args.steps, 1), max(args.warmup, 1), budget_s=150.0)
    step_s = sec * NUM_INFERENCE_STEPS
    value = clips * CONFIGS[cid]["frames"] / step_s
    line = {
        "impl": "reference", "metric": "denoised_video_frames_per_sec", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(cid, clips), "clips_per_gpu_per_step": clips,
                   "iterations_per_clip": NUM_INFERENCE_STEPS,
                   "note": "each step is a bounded sample: one measured CPU forward x 50 iterations (a full CPU clip is "
                           "minutes); the reference's own Python needs diffusers (absent), so this is the oracle port"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# second baseline leg: the reference's DEPLOYMENT path — eager PyTorch on the same GPU (SURVEY §8d "beat this")
# ----------------------------------------------------------------------------------------------------------------
def run_torch_eager(args, device: str = "cuda:0"):
    """Times the restated reference forward (oracle/flat_oracle.py) executed by eager torch in bf16 on `device`:
    cuBLAS GEMMs + SDPA(flash) + the ~60 elementwise kernels per block the reference launches.  A baseline leg like
    `cpu_baseline` — never the product — run in its own process by `torch_eager_gpu_leg`."""
    from oracle import flat_oracle as O
    dev = torch.device(device)
    cid, clips = args.config, args.clips_per_gpu or CONFIGS[args.config]["clips"]
    cfg, hs, text, t, kw = _oracle_case(cid, clips, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {}
    for name, shape in O.param_shapes(cfg).items():
        w = torch.randn(shape, generator=g, device=dev) * 0.02
        if name.endswith(".weight") and len(shape) == 1:
            w = 1.0 + w
        sd[name] = w.to(torch.bfloat16)
    cuda = dev.type == "cuda"
    steps, warm = max(args.steps, 1), max(args.warmup, 1)
    clk = None
    with torch.no_grad(), torch.device(dev):
        for _ in range(warm):
            O.forward(sd, cfg, hs, text, t, **kw)
        if cuda:
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with ClockSampler(dev.index or 0) as cs:
                e0.record()
                for _ in range(steps):
                    O.forward(sd, cfg, hs, text, t, **kw)
                e1.record()
                torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            clk = cs.summary()
        else:
            t0 = time.perf_counter()
            for _ in range(steps):
                O.forward(sd, cfg, hs, text, t, **kw)
            ms = (time.perf_counter() - t0) * 1e3 / steps
    frames = clips * CONFIGS[cid]["frames"]
    line = {"impl": "torch-eager", "metric": "denoised_video_frames_per_sec",
            "value": frames / (NUM_INFERENCE_STEPS * ms * 1e-3), "unit": "frames/s", "ms_per_forward": ms,
            "forwards_timed": steps, "warmup_forwards": warm, "device": str(dev), "clocks": clk,
            "sample": f"config-{cid} transformer forward ({clips} clip(s)), restated reference forward in eager torch bf16 "
                      f"(cuBLAS + SDPA); frames/s = frames per clip / ({NUM_INFERENCE_STEPS} forwards), sampler "
                      "arithmetic excluded"}
    print(json.dumps(line), flush=True)
    return line


def torch_eager_gpu_leg(cid: int, clips: int, timeout_s: float = 150.0) -> dict:
    """Runs `bench.py --impl torch-eager` in a child process (a failure there cannot take the bench line with it)."""
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "torch-eager", "--steps", "20",
                              "--warmup", "3", "--config", str(cid), "--clips-per-gpu", str(clips)],
                             capture_output=True, text=True, timeout=timeout_s)
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


CLASS_NAMES = ["prologue_adaln", "embed", "ln_modulate", "gemm_qkv", "attention", "gemm_out", "gemm_ff1", "gemm_ff2",
               "head"]
# substrings of the kernel names liborv_b200.so launches (everything else under the profiler is torch's own)
OUR_KERNELS = ("gemm2_bf16_kernel", "gemm_bf16_kernel", "gemm2_chain_kernel", "attention_kernel", "ln_ab_kernel",
               "ln_modulate_kernel", "skinny_linear_kernel", "patchify_kernel", "unpatchify_kernel", "ab_combine_kernel",
               "build_emb_kernel", "actions_to_f32_kernel", "timestep_sinusoid_kernel", "add_hidden_kernel",
               "mv_gather_kernel", "fill_tables_kernel", "sampler_step_kernel")


def init_vae_weights_(vae, seed: int):
    """Variance-preserving random decoder weights (std 1/sqrt(fan_in); GroupNorm scales and the multiplicative conv_y
    branch around 1) so that activations keep unit scale through the 37 norm sites, as a trained decoder's do."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    for name, p in vae.named_parameters():
        shape = tuple(p.shape)
        if name.endswith("norm_layer.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.05 * torch.randn(shape, generator=g) + (1.0 if "conv_y" in name else 0.0)
        elif ".conv_y." in name:
            t = torch.randn(shape, generator=g) * (0.3 / shape[1] ** 0.5)
        else:
            t = torch.randn(shape, generator=g) / (p[0].numel() ** 0.5)
        p.data.copy_(t.to(p.dtype))


def ncu_traffic(cls: str):
    """DRAM bytes per launch of a kernel class from the COMMITTED ncu capture (profiles/traffic.json, written by
    tools/summarize_profiles.py) — context for `roofline.traffic`, which stays null because it cannot be measured in a
    timed run."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    key = {"attention": "attention_kernel", "gemm_qkv": "gemm2_bf16_kernel<3>", "gemm_ff1": "gemm2_bf16_kernel<1>",
           "gemm_ff2": "gemm2_bf16_kernel<2>", "gemm_out": "gemm2_bf16_kernel<2>"}.get(cls)
    if not key or not os.path.exists(tp):
        return None
    for kname, ent in json.load(open(tp)).items():
        if key in kname and ent:
            return {"dram_bytes_per_launch": ent[0]["dram_bytes"], "us_under_ncu": ent[0]["us"], "source": ent[0]["source"],
                    "note": "config-2 capture, serialised and cold-cache"}
    return None


def class_flops(model_cfg: dict, B: int, S: int, views: int, St: int, tok: int, frames_lat: int) -> dict:
    """Algorithmic FLOP per LAUNCH of the tensor-core classes (2 M N K; attention 4 S^2 D, no causal discount).
    With multiview the classes average over the temporal block's and the view block's launches."""
    D = model_cfg["num_attention_heads"] * 64
    R = B * S
    f = {"gemm_qkv": 2.0 * R * 3 * D * D, "attention": 4.0 * B * S * S * D, "gemm_out": 2.0 * R * D * D,
         "gemm_ff1": 2.0 * R * 4 * D * D, "gemm_ff2": 2.0 * R * 4 * D * D}
    if views > 1:
        clips = B // views
        Smv, q = views * (St + tok), views * tok
        att_mv = 4.0 * clips * frames_lat * q * Smv * D
        f["attention"] = (f["attention"] + att_mv) / 2
        Mv = clips * frames_lat * views * tok
        f["gemm_out"] = (f["gemm_out"] + 2 * 2.0 * Mv * D * D) / 3  # temporal out + view out + view proj_out
    return f


def attribute_device_time(evs, classes):
    """evs: (start_us, end_us, kernel name) of this library's kernels on one stream; classes: ORVB_PC_* of the
    launches of one forward, in order.  Forwards are delimited by the sampler-step kernel that follows each of them;
    only segments with exactly len(classes) kernels are used (the first iteration of a clip also builds the
    step-invariant cache and the modulation schedule, so it is longer).  Every kernel is charged its
    completion-to-completion interval, end - previous end (the first one end - start).  Returns (per-class [us, launches], per-class raw duration, sampler us, forward
    spans, forwards used) or None."""
    evs = sorted(evs)
    segs, cur = [], []
    for s, e, name in evs:
        if "sampler_step_kernel" in name:
            segs.append((cur, (s, e)))
            cur = []
        else:
            cur.append((s, e, name))
    good = [(seg, ss) for seg, ss in segs if classes and len(seg) == len(classes)]
    if not good:
        return None
    acc = {n: [0.0, 0] for n in CLASS_NAMES}
    raw = {n: 0.0 for n in CLASS_NAMES}
    samp, spans = 0.0, []
    for seg, (ss, se) in good:
        prev_end = seg[0][0]
        for (s, e, _), c in zip(seg, classes):
            excl = max(e - prev_end, 0.0)
            prev_end = max(prev_end, e)
            acc[CLASS_NAMES[c]][0] += excl
            acc[CLASS_NAMES[c]][1] += 1
            raw[CLASS_NAMES[c]] += e - s
        spans.append(prev_end - seg[0][0])
        samp += max(se - prev_end, 0.0)
    return acc, raw, samp, spans, len(good)


def profile_step(run_once, model, flops: dict, pk: dict):
    """Device-side per-class timing of one real pipeline call (its forwards are CUDA-graph replays): CUPTI activity
    records through torch.profiler.  Every kernel is charged its completion-to-completion interval on the stream (end -
    previous kernel's end): device-side launch gaps are included, host enqueue time between forwards and the overlap of
    programmatic dependent launches are not, and the classes of a forward sum to exactly its span.
    `us_kernel_duration` is the plain start-to-end duration (what ncu reports, up to its serialisation)."""
    from torch.autograd import DeviceType
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run_once()
        torch.cuda.synchronize()
    classes = list(getattr(model, "last_launch_classes", []) or [])
    evs = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events()
           if e.device_type == DeviceType.CUDA and any(k in e.name for k in OUR_KERNELS)]
    res = attribute_device_time(evs, classes)
    if res is None:
        return None, {"unavailable": f"no profiled forward matched the launch list ({len(classes)} launches, "
                                     f"{len(evs)} kernel records)"}
    acc, raw, samp, spans, nf = res
    kernels = {}
    for n, (us, cnt) in acc.items():
        if cnt == 0:
            continue
        per = us / cnt
        ent = {"us_per_launch": round(per, 2), "us_kernel_duration": round(raw[n] / cnt, 2),
               "launches_per_forward": cnt // nf, "ms_per_forward": round(us / nf / 1e3, 4)}
        if n in flops:
            ent["tflops"] = round(flops[n] / (per * 1e-6) / 1e12, 1)
            ent["frac_of_peak"] = round(ent["tflops"] / pk["bf16"], 4)
        kernels[n] = ent
    kernels["sampler_step"] = {"us_per_launch": round(samp / nf, 2), "launches_per_forward": 1,
                               "ms_per_forward": round(samp / nf / 1e3, 4)}
    info = {"forwards_profiled": nf, "forward_ms_sum_of_kernels": round(sum(v[0] for v in acc.values()) / nf / 1e3, 4),
            "forward_ms_span": round(sum(spans) / nf / 1e3, 4),
            "method": "CUPTI activity records (torch.profiler) of one pipeline call replaying its CUDA graphs; per kernel "
                      "end - previous end (us_per_launch) and end - start (us_kernel_duration)"}
    return kernels, info


def build_runner(cid: int, clips: int = 0, rank: int = 0, dev=None):
    """Model + pipeline + synthetic inputs of a config on `dev`; returns (run(seed, from_host=False) -> latents, model,
    info).  Shared by run_ours and the tools/ profilers."""
    from orv_b200 import (CogVideoXDPMScheduler, CogVideoXImageToVideoPipelineTraj,
                          CogVideoXTransformer3DModelTraj)
    from orv_b200 import dist as D
    from orv_b200.models.pipeline_control import default_vae_config
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    c = CONFIGS[cid]
    B = clips or c["clips"]
    V = c["views"]
    with torch.device(dev):
        model = CogVideoXTransformer3DModelTraj(**c["model"])
    if rank == 0:
        init_weights_(model, seed=0)
    model = model.to(torch.bfloat16).eval()
    model.action_embed.mask = False  # parity/bench runs pin the reference's stray action dropout off
    D.broadcast_weights(model, src=0)  # the single collective of the path
    sched = CogVideoXDPMScheduler(timestep_spacing="trailing")
    # 3-D VAE decoder with the reference script's settings (inference_control_to_video.py:98-99); random
    # variance-preserving weights.  The headline metric excludes decoding; the `decode` leg reports it.
    from orv_b200 import AutoencoderKLCogVideoX
    with torch.device(dev):
        vae = AutoencoderKLCogVideoX()
    init_vae_weights_(vae, seed=2)
    vae = vae.to(torch.bfloat16).eval()
    vae.enable_slicing()
    vae.enable_tiling()
    pipe = CogVideoXImageToVideoPipelineTraj(None, None, vae, model, sched)
    # The reference's 1.5-5B pipeline path cannot run with one view (`first_frame` is sliced with
    # size(1) of a 6-D tensor = n_views, cogvideox_control.py:1212-1214, so the image latents get one frame more than the
    # noise latents and :1413 fails).  The drop-in reproduces that by default; the bench asks for the intended padding.
    pipe.fix_patch_t_padding = c["model"].get("patch_size_t") is not None

    # ---- synthetic inputs, host-pinned masters + device-resident copies ----
    g = torch.Generator().manual_seed(1 + rank)
    h, w = c["px"][0] // 8, c["px"][1] // 8
    n_cfg = 2 if c["cfg_pair"] else 1
    host = {"image": torch.randn(B, 32, V, h, w, generator=g).bfloat16(),           # first-frame VAE moments per view
            "text": (torch.randn(B, 226, 4096, generator=g) * 0.2).bfloat16(),
            # with CFG the reference does not duplicate controls: the caller passes them per sequence
            "actions": ((torch.rand(B * n_cfg, 16, 7, generator=g) * 2 - 1) * torch.tensor([20.0] * 6 + [1.0])).bfloat16()}
    if c["cfg_pair"]:
        host["neg_text"] = torch.zeros_like(host["text"])
    if c["controls"]:  # depth / semantic-label VAE moments [B, 32, V*F, h, w]
        for key in ("depths", "labels"):
            m = torch.randn(B, 32, V * 5, h, w, generator=g)
            m[:, 16:] = m[:, 16:] * 0.5 - 3.0
            host[key] = m.bfloat16()
    host = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}

    def run_clip(inp, seed, output_type="latent"):
        gen = torch.Generator().manual_seed(seed)
        cg = {"actions": inp["actions"]}
        if c["controls"]:
            cg["depths"], cg["labels"] = inp["depths"], inp["labels"]
        kw = dict(prompt=[""] * B, negative_prompt_embeds=None)
        if c["cfg_pair"]:  # check_inputs quirk (:1261-1270): the embeds pair only validates with prompt=None
            kw = dict(prompt=None, negative_prompt_embeds=inp["neg_text"])
        # prompt AND prompt_embeds, as the reference programs call it (evaluation_control_to_video.py:321-322)
        return pipe(image=inp["image"], prompt_embeds=inp["text"], height=c["px"][0], width=c["px"][1], num_frames=17,
                    num_inference_steps=NUM_INFERENCE_STEPS, guidance_scale=6.0 if c["cfg_pair"] else 1.0, generator=gen,
                    controls_or_guidances=cg, output_type=output_type, return_dict=False, num_views=V, **kw)[0]

    def run(seed, from_host=False, output_type="latent"):
        if from_host:
            return run_clip({k: v.to(dev, non_blocking=True) for k, v in host.items()}, seed, output_type)
        return run_clip(resident, seed, output_type)

    info = dict(B=B, V=V, n_cfg=n_cfg, h=h, w=w, vae=vae, host_bytes=sum(v.numel() * v.element_size() for v in host.values()),
                pipe=pipe, arena_bytes=model.weight_arena().numel() * 2)
    return run, model, info


def run_ours(args):
    from orv_b200 import dist as D
    rank, local_rank, world = D.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cid = args.config
    c = CONFIGS[cid]
    run, model, info = build_runner(cid, args.clips_per_gpu, rank, dev)
    B, V, n_cfg, h, w, pipe, arena_bytes = (info[k] for k in ("B", "V", "n_cfg", "h", "w", "pipe", "arena_bytes"))

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

    frames_per_step = B * c["frames"]
    tflop_per_step = NUM_INFERENCE_STEPS * c["tflop"] * B
    # ---- device-resident throughput ----
    out = run(41)  # (always at least one untimed call: graph capture, allocations; also the shape of the result)
    for i in range(args.warmup):
        out = run(42 + i)
    with ClockSampler(local_rank) as clocks:
        secs = timed(lambda i: run(100 + i), args.steps)
    launches_per_step = pipe.last_step_launches
    value = world * args.steps * frames_per_step / secs

    # ---- end-to-end: host inputs in, latents out ----
    # The result of every clip is read back into pinned host memory inside the timed region.  The read is asynchronous
    # (as a serving loop would issue it): the host waits for clip i's latents only when their slot is needed again, i.e.
    # while the GPU already works on clip i + 1, so the next call's host prologue (CPU-generator draws, staging) overlaps
    # the tail of the current one.  `timed` synchronises at the end, so all K reads have landed when the clock stops.
    res_pin = [torch.empty(out.shape, dtype=out.dtype).pin_memory() for _ in range(2)]
    res_evt = [torch.cuda.Event() for _ in range(2)]

    def e2e_clip(i):
        slot = i & 1
        res_evt[slot].synchronize()
        res_pin[slot].copy_(run(200 + i, from_host=True), non_blocking=True)
        res_evt[slot].record()
        return res_pin[slot]

    if os.environ.get("ORVB_BENCH_BLOCKING_READBACK", "0") == "1":  # A/B: one blocking .cpu() per clip
        def e2e_clip(i):  # noqa: F811
            return run(200 + i, from_host=True).cpu()

    e2e_clip(0)
    e2e_secs = timed(e2e_clip, args.steps)
    e2e_value = world * args.steps * frames_per_step / e2e_secs
    lat_bytes = out.numel() * 2
    # + host-generated DPM noise per iteration + initial noise (CPU-generator contract of the reference scripts)
    h2d = info["host_bytes"] + (NUM_INFERENCE_STEPS + 1) * lat_bytes
    d2h = lat_bytes

    # ---- the same call with the 3-D VAE decode behind it (output_type="pt": frames in [0, 1], read back) ----
    decode = None
    try:
        vae = info["vae"]

        def full_clip(i):
            return run(400 + i, from_host=True, output_type="pt").cpu()

        vid = full_clip(0)
        full_secs = timed(full_clip, args.steps)
        lat = out.permute(0, 2, 1, 3, 4) / pipe.vae_scaling_factor_image
        vae.decode(lat)
        dec_secs = timed(lambda i: vae.decode(lat), args.steps)
        decode = {"e2e_with_decode": {"value": world * args.steps * frames_per_step / full_secs, "unit": "frames/s",
                                      "ms_per_step": full_secs / args.steps * 1e3,
                                      "d2h_bytes_per_step": int(vid.numel() * vid.element_size())},
                  "decode_ms_per_step": dec_secs / args.steps * 1e3, "decode_launches_per_step": int(vae.last_launches),
                  "decode_conv_tflop_per_step": round(lat.shape[0] * vae.decode_conv_flops(*lat.shape[2:]) / 1e12, 3),
                  "decode_tensor_frac_of_peak": round(lat.shape[0] * vae.decode_conv_flops(*lat.shape[2:]) / 1e12
                                                      / (dec_secs / args.steps) / peaks()["bf16"], 4),
                  "frames_shape": list(vid.shape), "tiling": True, "slicing": True,
                  "note": "AutoencoderKLCogVideoX.decode on orvb_conv_cl / orvb_spatial_norm_cl (SURVEY 8 f2); not part of "
                          "the headline metric, which is the denoise loop (SURVEY 8d)"}
    except Exception as e:  # noqa: BLE001 — extra leg: never take the bench line down
        decode = {"unavailable": repr(e)[:300]}

    # ---- per-kernel-class device times of one real step ----
    pk = peaks()
    mc = c["model"]
    pt = mc.get("patch_size_t") or 1
    Fl = -(-5 // pt)  # token frames
    tok = (h // 2) * (w // 2)
    S = 226 + Fl * tok
    flops = class_flops(mc, B * n_cfg * V, S, V, 226, tok, Fl)
    try:
        kernels, pinfo = profile_step(lambda: run(300), model, flops, pk)
    except Exception as e:  # noqa: BLE001 — measurement helper: never take the bench line down
        kernels, pinfo = None, {"unavailable": repr(e)[:300]}
    roofline = None
    if kernels:
        dom = max(flops.keys(), key=lambda n: kernels.get(n, {}).get("ms_per_forward", 0.0))
        roofline = {"bound": "tensor", "kernel": dom, "achieved": kernels[dom]["tflops"], "peak": pk["bf16"],
                    "unit": "TFLOP/s", "frac": round(kernels[dom]["tflops"] / pk["bf16"], 4),
                    # DRAM bytes need an ncu replay and cannot be measured inside a timed run: null here; the committed
                    # capture of this kernel is profiles/r02*_attn_raw.csv (dram__bytes_read.sum + dram__bytes_write.sum)
                    "traffic": None, "traffic_ncu": ncu_traffic(dom),
                    "algorithmic_flop_per_launch": flops[dom],
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['source']})",
                    "forward_ms_sum_of_kernels": pinfo["forward_ms_sum_of_kernels"],
                    "forward_tensor_frac": round(c["tflop"] * B / (pinfo["forward_ms_sum_of_kernels"] * 1e-3) / pk["bf16"], 4),
                    "sum_of_kernels_le_step": bool(pinfo["forward_ms_sum_of_kernels"]
                                                   <= secs / args.steps * 1e3 / NUM_INFERENCE_STEPS)}

    clk = clocks.summary()
    line = {
        "metric": "denoised_video_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(cid, B), "clips_per_gpu_per_step": B,
                   "iterations_per_clip": NUM_INFERENCE_STEPS, "frames_per_clip": c["frames"],
                   "l2": "weights (%.2f GB per forward) stream from HBM every iteration, far above the 126 MB L2; "
                         "no explicit flush" % (arena_bytes / 1e9),
                   "parallelism": f"dp{world} (independent clips, one NCCL weight broadcast at init)"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_secs / args.steps * 1e3},
        "gpu_launches": int(launches_per_step * args.steps),
        "tensor_frac_of_peak": round(tflop_per_step * world * args.steps / secs / (pk["bf16"] * world), 4),
        "tensor_frac_of_burst_peak": round(tflop_per_step * world * args.steps / secs / (pk["bf16_burst"] * world), 4),
        "roofline": roofline, "kernels": kernels, "kernel_timing": pinfo, "clocks": clk, "decode": decode,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, sample, cores = cpu_reference_forward_seconds(cid, B, 3, 1, budget_s=25.0)
        line["cpu_baseline"] = {"value": frames_per_step / (sec * NUM_INFERENCE_STEPS), "unit": "frames/s",
                                "cores": cores, "kind": "port", "sample": sample}
        # what the reference actually deploys: the same forward in eager torch on this GPU (own process)
        line["torch_eager_gpu"] = torch_eager_gpu_leg(cid, B)
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
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json config index (default 2: the configuration the metric is quoted on)")
    ap.add_argument("--clips-per-gpu", type=int, default=0,
                    help="clips per pipeline call and GPU (0 = the config's default: 1, config 5: 2)")
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
