"""CPU restatement of the weight-only storage formats declared in include/esmk.h (TEST INFRASTRUCTURE: only
tests/ may import this).  The reference delegates quantised storage to bitsandbytes (esme/esm.py:414-446,
`Linear4bit` / `Linear8bitLt`), a third-party dependency that is absent from /root/reference and from this image,
so there is no reference output to pin against: PARITY UNPINNED.  What is restated here is bitsandbytes' published
fp4 codebook and block layout (blocksize 64, absmax scaling, even element in the high nibble) without its
second-level quantisation of the absmax values, and row-wise absmax int8."""
import numpy as np
import torch

FP4_TABLE = np.array([0.0, 5.208333333e-03, 0.66666667, 1.0, 0.33333333, 0.5, 0.16666667, 0.25], dtype=np.float32)


def _bf16_to_f32(w: torch.Tensor) -> np.ndarray:
    return w.detach().float().cpu().numpy()


def _f32_to_bf16(x: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).bfloat16()


def q4_codes(x: np.ndarray) -> np.ndarray:
    """Nearest fp4 code of fp32 values in [-1, 1]: decision tree over the midpoints of the table values."""
    a = np.abs(x)
    c = np.where(a > np.float32(0.29166667),
                 np.where(a > np.float32(0.583333), np.where(a > np.float32(0.8333333), 3, 2),
                          np.where(a > np.float32(0.4166667), 5, 4)),
                 np.where(a > np.float32(0.0859375), np.where(a > np.float32(0.20833333), 7, 6),
                          np.where(a > np.float32(0.00260417), 1, 0))).astype(np.uint8)
    return c | np.where(x < 0, 8, 0).astype(np.uint8)


def q4_quantize(w: torch.Tensor):
    """bf16 [N,K] -> (uint8 [N*K/2, 1], fp32 absmax [N*K/64])."""
    v = _bf16_to_f32(w).reshape(-1, 64)
    absmax = np.abs(v).max(axis=1).astype(np.float32)
    inv = np.where(absmax > 0, np.float32(1.0) / np.where(absmax > 0, absmax, 1).astype(np.float32), 0).astype(np.float32)
    codes = q4_codes((v * inv[:, None]).astype(np.float32)).reshape(-1, 2)
    packed = ((codes[:, 0] << 4) | codes[:, 1]).astype(np.uint8)
    return torch.from_numpy(packed.reshape(-1, 1)), torch.from_numpy(absmax)


def q4_dequantize(packed: torch.Tensor, absmax: torch.Tensor, N: int, K: int) -> torch.Tensor:
    p = packed.cpu().numpy().reshape(-1)
    codes = np.stack((p >> 4, p & 15), axis=1).reshape(-1, 64)
    mag = FP4_TABLE[codes & 7]
    val = np.where(codes & 8, -mag, mag).astype(np.float32) * absmax.cpu().numpy().astype(np.float32)[:, None]
    return _f32_to_bf16(val.reshape(N, K))


def q8_quantize(w: torch.Tensor):
    """bf16 [N,K] -> (int8 [N,K], fp32 scale [N] = row absmax / 127)."""
    v = _bf16_to_f32(w)
    absmax = np.abs(v).max(axis=1).astype(np.float32)
    inv = np.where(absmax > 0, np.float32(127.0) / np.where(absmax > 0, absmax, 1).astype(np.float32), 0).astype(np.float32)
    q = np.clip(np.rint((v * inv[:, None]).astype(np.float32)), -127, 127).astype(np.int8)
    return torch.from_numpy(q), torch.from_numpy((absmax / np.float32(127.0)).astype(np.float32))


def q8_dequantize(q: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    return _f32_to_bf16(q.cpu().numpy().astype(np.float32) * scale.cpu().numpy().astype(np.float32)[:, None])
