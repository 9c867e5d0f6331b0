"""Helpers shared by the GPU parity tests (oracle = torch-CPU float64, see oracle/)."""
import numpy as np
import torch

from oracle import x3d_oracle as O


def dev():
    return torch.device("cuda", 0)


def bf16_round(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.asarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def to_dev(a: np.ndarray, dtype=torch.float32) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev()).to(dtype).contiguous()


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.float32).cpu().numpy()


def rel_err(got: np.ndarray, want: np.ndarray) -> float:
    """max |got-want| / max |want|  -- the 'relative on logits' measure of north_star."""
    return float(np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


def assert_close(got, want, dtype, what=""):
    """fp32 storage: 1e-5 of the tensor's scale; bf16 storage: one bf16 ulp (2^-8) of each value
    plus 2^-9 of the scale for accumulated input rounding."""
    want = np.asarray(want, np.float64)
    got = np.asarray(got, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(np.abs(want).max(), 1e-30)
    if dtype == torch.float32:
        tol = 2e-5 * scale + 1e-5 * np.abs(want)
    else:
        tol = 2.0 ** -9 * scale + 2.0 ** -8 * np.abs(want)
    bad = np.abs(got - want) > tol
    assert not bad.any(), (f"{what}: {bad.sum()} / {bad.size} elements off; max err "
                           f"{np.abs(got - want).max():.3e} at scale {scale:.3e}")


def ncdhw64(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.asarray(a, np.float64)).permute(0, 4, 1, 2, 3).contiguous()
