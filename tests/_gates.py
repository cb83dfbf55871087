"""Forward-level parity gates shared by the GPU tests: a mean-abs gate alone is blind to "a few rows come out wrong"
(the failure mode of a mis-synchronised attention tile), so every comparison also bounds the largest element error
and the worst row."""
import torch


def forward_errors(out: torch.Tensor, ref: torch.Tensor):
    """(mean-abs error / mean-abs ref, max-abs error / max-abs ref, worst row: ||err||_2 / ||ref||_2 over rows of the
    last dimension; rows whose reference norm is below 10 % of the median row norm are measured against the median)."""
    a, b = out.detach().float().cpu(), ref.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.isfinite(a).all()
    err = (a - b).abs()
    mean_rel = (err.mean() / b.abs().mean()).item()
    max_rel = (err.max() / b.abs().max()).item()
    e2 = (a - b).reshape(-1, a.shape[-1]).norm(dim=1)
    r2 = b.reshape(-1, b.shape[-1]).norm(dim=1)
    floor = 0.1 * r2.median()
    row_rel = (e2 / torch.maximum(r2, floor)).max().item()
    return mean_rel, max_rel, row_rel


def assert_forward_close(out, ref, mean_tol, max_tol, row_tol, what=""):
    mean_rel, max_rel, row_rel = forward_errors(out, ref)
    print(f"{what}: mean-rel {mean_rel:.3e} (< {mean_tol:g})  max-abs/max {max_rel:.3e} (< {max_tol:g})  "
          f"worst row {row_rel:.3e} (< {row_tol:g})")
    assert mean_rel < mean_tol, (what, "mean", mean_rel)
    assert max_rel < max_tol, (what, "max", max_rel)
    assert row_rel < row_tol, (what, "row", row_rel)
    return mean_rel, max_rel, row_rel
