"""Config-2 throughput with B clips per pipeline call on one B200 (supplementary to bench.py, whose line keeps B = 1).

Question for the next round: does batching independent clips into one forward raise frames/s?  The GEMMs see M = B x 3226
rows (the short-K attention-out projection gets 4 instead of 2 tiles per SM pair, so its exposed fill / last epilogue
amortise), attention gets B x 390 CTAs (same 12 % wave quantisation), the AdaLN weight stream is shared by the batch.

    python tools/bench_batch.py [B=2] [steps=3]        (under gpurun)
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from bench import FRAMES_PER_CLIP, FWD_TFLOP, NUM_INFERENCE_STEPS, config2, init_weights_, peaks  # noqa: E402
from orv_b200 import (CogVideoXDPMScheduler, CogVideoXImageToVideoPipelineTraj,  # noqa: E402
                      CogVideoXTransformer3DModelTraj)
from orv_b200.models.pipeline_control import default_vae_config  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda", 0)
    with torch.device(dev):
        model = CogVideoXTransformer3DModelTraj(**config2())
    init_weights_(model, seed=0)
    model = model.to(torch.bfloat16).eval()
    model.action_embed.mask = False
    pipe = CogVideoXImageToVideoPipelineTraj(None, None, default_vae_config(), model,
                                             CogVideoXDPMScheduler(timestep_spacing="trailing"))
    g = torch.Generator().manual_seed(1)
    image = torch.randn(B, 32, 1, 40, 60, generator=g).bfloat16().to(dev)
    text = (torch.randn(B, 226, 4096, generator=g) * 0.2).bfloat16().to(dev)
    act = ((torch.rand(B, 16, 7, generator=g) * 2 - 1) * torch.tensor([20.0] * 6 + [1.0])).bfloat16().to(dev)

    def run(seed):
        return pipe(image=image, prompt=[""] * B, prompt_embeds=text, height=320, width=480, num_frames=17,
                    num_inference_steps=NUM_INFERENCE_STEPS, guidance_scale=1.0,
                    generator=torch.Generator().manual_seed(seed), controls_or_guidances={"actions": act},
                    output_type="latent", return_dict=False)[0]

    for i in range(3):
        run(40 + i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = run(100 + i)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) / 1e3
    pk = peaks()
    print(json.dumps({
        "workload": f"config 2, {B} clips per pipeline call, {NUM_INFERENCE_STEPS} DPM-trailing iterations",
        "clips_per_call": B, "calls": steps, "out_shape": list(out.shape),
        "frames_per_s": B * steps * FRAMES_PER_CLIP / secs, "ms_per_call": secs / steps * 1e3,
        "ms_per_clip": secs / steps / B * 1e3,
        "tensor_frac_of_peak": NUM_INFERENCE_STEPS * FWD_TFLOP * B * steps / secs / pk["bf16"]}))


if __name__ == "__main__":
    main()
