"""CPU: the baseline legs of bench.py that cannot be exercised on the GPU from here."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_torch_eager_leg_runs_the_restated_forward(capsys):
    import bench
    args = types.SimpleNamespace(steps=1, warmup=1)
    tiny = dict(num_attention_heads=2, num_layers=2, sample_width=12, sample_height=8, text_embed_dim=64,
                max_text_seq_length=8)
    line = bench.run_torch_eager(args, device="cpu", cfg_over=tiny)
    printed = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert printed["impl"] == "torch-eager" and printed["unit"] == "frames/s"
    assert line["value"] > 0 and abs(line["value"] - 16 / (50 * line["ms_per_forward"] * 1e-3)) < 1e-6 * line["value"]


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
