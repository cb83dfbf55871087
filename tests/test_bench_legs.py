"""CPU: the baseline legs of bench.py that cannot be exercised on the GPU from here."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


_TINY = dict(name="tiny test geometry", model=dict(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=16,
                                                  num_layers=2, sample_width=12, sample_height=8, sample_frames=17,
                                                  modulate_encoder_hidden_states=True, text_embed_dim=64,
                                                  max_text_seq_length=8, time_embed_dim=64, patch_size=2),
             px=(64, 96), views=1, controls=False, cfg_pair=False, tflop=1e-3, frames=16, clips=1)


def test_torch_eager_and_reference_legs_run_the_restated_forward(capsys, monkeypatch):
    import bench
    monkeypatch.setitem(bench.CONFIGS, 99, _TINY)
    args = types.SimpleNamespace(steps=1, warmup=1, config=99, clips_per_gpu=2, gpus=1)
    line = bench.run_torch_eager(args, device="cpu")
    printed = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert printed["impl"] == "torch-eager" and printed["unit"] == "frames/s"
    assert line["value"] > 0 and abs(line["value"] - 2 * 16 / (50 * line["ms_per_forward"] * 1e-3)) < 1e-6 * line["value"]
    monkeypatch.delenv("RANK", raising=False)
    bench.run_reference(args)
    ref = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["cpu_baseline"]["kind"] == "port" and ref["value"] > 0
    assert ref["config"]["workload"] == bench.workload_name(99, 2) and ref["e2e"]["value"] == ref["value"]


def test_every_baseline_config_builds_oracle_inputs_of_the_right_geometry():
    import bench
    want = {2: (1, 5, 40, 60, 1), 3: (1, 5, 40, 60, 1), 4: (2, 6, 40, 60, 1), 5: (2, 15, 32, 48, 3)}
    for cid, (B, F, h, w, V) in want.items():
        import torch
        from oracle import flat_oracle as O
        c = bench.CONFIGS[cid]
        cfg = O.default_config(**c["model"])
        pt = cfg["patch_size_t"] or 1
        assert (-(-5 // pt) * pt * V, c["px"][0] // 8, c["px"][1] // 8, c["views"]) == (F, h, w, V)
        assert (c["clips"] * (2 if c["cfg_pair"] else 1)) == B
    f = bench.class_flops(bench.CONFIGS[2]["model"], 1, 3226, 1, 226, 600, 5)
    total = 30 * sum(f.values())
    assert abs(total / 1e12 - 10.968) < 0.06  # 10.968 TFLOP per config-2 forward, 99.9 % of it in these classes


def test_device_time_attribution_charges_completion_to_completion_intervals():
    import bench
    classes = [2, 3, 4]  # ln, qkv, attention
    evs = []
    t = 100.0
    for it in range(3):
        if it == 0:
            evs.append((t, t + 5, "skinny_linear_kernel"))  # first iteration: one extra kernel -> segment is skipped
            t += 6
        evs += [(t, t + 10, "ln_ab_kernel"), (t + 8, t + 40, "gemm2_bf16_kernel<3>"),  # PDL overlap: starts 2 us early
                (t + 45, t + 100, "attention_kernel"), (t + 101, t + 104, "sampler_step_kernel")]
        t += 300  # host gap between iterations must not be charged to anything
    evs.append((0.0, 1.0, "at::native::copy_kernel"))  # caller filters torch kernels; a stray one before is harmless
    acc, raw, samp, spans, nf = bench.attribute_device_time([e for e in evs if "at::" not in e[2]], classes)
    assert nf == 2
    assert acc["ln_modulate"] == [20.0, 2] and acc["gemm_qkv"] == [60.0, 2] and acc["attention"] == [120.0, 2]
    assert raw["gemm_qkv"] == 64.0 and spans == [100.0, 100.0] and samp == 8.0
    assert bench.attribute_device_time(evs[:3], classes) is None


def test_position_tables_are_cached_per_device_and_dtype():
    import torch
    from oracle import flat_oracle as O
    cfg = O.default_config(num_attention_heads=2, attention_head_dim=64, sample_width=12, sample_height=8,
                           max_text_seq_length=8)
    like32, like16 = torch.zeros(1), torch.zeros(1, dtype=torch.bfloat16)
    a = O._cached_table("joint", cfg, (3, 8, 12), like32)
    assert O._cached_table("joint", cfg, (3, 8, 12), like32) is a
    assert torch.equal(a, O.joint_pos_embedding(cfg, 3, 8, 12))
    b = O._cached_table("joint", cfg, (3, 8, 12), like16)
    assert b.dtype == torch.bfloat16 and torch.equal(b, a.to(torch.bfloat16))
    assert O._cached_table("joint", cfg, (2, 8, 12), like32).shape[0] != a.shape[0]
