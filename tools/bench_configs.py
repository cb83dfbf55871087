"""Forward time of the other BASELINE.json configs (3: 2B + depth/label conditions, 4: CogVideoX1.5-5B dims with a
CFG pair, 5: 2B multiview V=3) at full size on one B200 — supplementary to bench.py (whose line is config 2).
Run under gpurun:  python tools/bench_configs.py [3 4 5]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import init_weights_, peaks  # noqa: E402
from orv_b200 import CogVideoXTransformer3DModelTraj  # noqa: E402
from orv_b200.models.embeddings import get_3d_rotary_pos_embed  # noqa: E402

BASE2B = dict(num_attention_heads=30, attention_head_dim=64, in_channels=32, out_channels=16, num_layers=30,
              modulate_encoder_hidden_states=True, text_embed_dim=4096, max_text_seq_length=226, time_embed_dim=512, patch_size=2)
CASES = {
    3: dict(cfg=dict(BASE2B, sample_width=60, sample_height=40, sample_frames=17, visual_guidance=True, num_control_blocks=2),
            B=1, F=5, H=40, W=60, controls=True, tflop=11.023, clips=1, name="2B + depth/label conditions, 17x320x480"),
    4: dict(cfg=dict(BASE2B, num_attention_heads=48, num_layers=42, sample_width=60, sample_height=40, sample_frames=21,
                     patch_size_t=2, use_rotary_positional_embeddings=True, ofs_embed_dim=512, patch_bias=False),
            B=2, F=6, H=40, W=60, rope=True, ofs=2.0, n_actions=20, tflop=2 * 21.404, clips=1,
            name="CogVideoX1.5-5B dims, CFG pair (batch 2), 6 latent frames"),
    5: dict(cfg=dict(BASE2B, sample_width=48, sample_height=32, sample_frames=17, visual_guidance=True, num_control_blocks=2,
                     multiview=True, max_n_view=3),
            B=1, F=5, H=32, W=48, controls=True, views=3, tflop=33.633, clips=1, name="2B multiview V=3, 256x384 + conditions"),
}
dev = torch.device("cuda")
pk = peaks()
for cid in [int(a) for a in sys.argv[1:]] or [3, 4, 5]:
    c = CASES[cid]
    cfg = c["cfg"]
    V = c.get("views", 1)
    with torch.device(dev):
        model = CogVideoXTransformer3DModelTraj(**cfg)
    init_weights_(model, 0)
    model = model.to(torch.bfloat16).eval()
    model.action_embed.mask = False
    gen = torch.Generator(device=dev).manual_seed(1)
    lat_shape = (c["B"], c["F"] * V, cfg["in_channels"], c["H"], c["W"])
    hs = torch.randn(lat_shape, generator=gen, device=dev).bfloat16()
    text = (torch.randn(c["B"], 226, 4096, generator=gen, device=dev) * 0.2).bfloat16()
    cg = {"actions": ((torch.rand(c["B"], c.get("n_actions", 16), 7, generator=gen, device=dev) * 2 - 1) * 20).bfloat16()}
    if c.get("controls"):
        cg["depths"] = torch.randn(lat_shape, generator=gen, device=dev).bfloat16()
        cg["labels"] = torch.randn(lat_shape, generator=gen, device=dev).bfloat16()
    rope = None
    if c.get("rope"):  # CogVideoX1.5 'slice' grid over (F / p_t, H / 2, W / 2) tokens
        gh, gw, pt = c["H"] // 2, c["W"] // 2, cfg["patch_size_t"]
        cos, sin = get_3d_rotary_pos_embed(64, None, (gh, gw), (c["F"] + pt - 1) // pt, grid_type="slice",
                                           max_size=(cfg["sample_height"] // 2, cfg["sample_width"] // 2))
        rope = (cos.to(dev).float().contiguous(), sin.to(dev).float().contiguous())
    ofs = torch.tensor([c["ofs"]], device=dev) if "ofs" in c else None
    t = torch.full((c["B"],), 499, device=dev)
    call = lambda: model(hs, text, cg, t, ofs=ofs, image_rotary_emb=rope, return_dict=False, num_views=V)[0]  # noqa: E731
    with torch.no_grad():
        for _ in range(4):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for _ in range(n):
            call()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = c["tflop"] / (ms * 1e-3)
    print(json.dumps({"config": cid, "name": c["name"], "ms_per_forward": round(ms, 3), "tflops": round(tf, 1),
                      "frac_of_sustained_peak": round(tf / pk["bf16"], 4), "forward_tflop": c["tflop"],
                      "frames_per_s_at_50_steps": round(16 * V * c["clips"] / (50 * ms * 1e-3), 2)}), flush=True)
    del model
    torch.cuda.empty_cache()
    time.sleep(0.5)
