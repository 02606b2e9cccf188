"""Host-side D3PM constants for the CUDA reverse step.

The reference keeps three dense (S, K, K) fp16 tensors (``ar_discrete.py:257-277``: one-step
matrices, their fp16 chain product ``q_mats`` and the transposes, 3 x 210 MB at S=100, K=1025) and
indexes them with one-hot matmuls.  Both transition families have rank-structured matrices, so the
kernels only need a few scalars per timestep.  To keep those scalars equal to what the reference
holds (the fp16 chain product drifts from the analytic value — row sums reach 1.002 — and
``eps=1e-6`` makes that drift observable), they are read out of the same fp16 chain product,
computed here once with a running (K, K) matrix instead of being re-derived from alpha-bar.
Absorbing: every entry of the product has at most two non-zero terms, so the product is exactly
rank-structured and the scalars are bit-identical to the dense entries.  Uniform: K-term sums, so
the dense product carries up to 2 diagonal / 3 off-diagonal values per t that differ by one fp16
ulp; the scalars are representative entries (``tests/test_host_cpu.py`` bounds the gap), and the
bit-exact parity path reads the dense table instead (``dense_log_qbar``).
"""
from __future__ import annotations

import functools

import numpy as np
import torch

from ..b200 import lib as L

EPS = 1.0e-6  # ar_discrete.py:276


def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> torch.Tensor:
    """Cosine schedule with the reference's grid: ``steps`` points spread over [0, steps]
    (not [0, 1]) before normalising by ``steps`` (ar_discrete.py:286-304)."""
    steps = timesteps + 1
    grid = np.linspace(0, steps, steps)
    abar = np.cos(((grid / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    abar = abar / abar[0]
    betas = np.clip(1 - (abar[1:] / abar[:-1]), a_min=0, a_max=0.999)
    return torch.from_numpy(betas)


def _onestep(beta_t: torch.Tensor, K: int, transition: str) -> torch.Tensor:
    if transition == "absorbing":      # (1-b) I + b 1 e_m^T, float64 then fp16 (ar_discrete.py:315-334,269)
        b = float(beta_t.numpy())
        q = torch.zeros(K, K, dtype=torch.float64)
        q.fill_diagonal_(1.0 - b)
        q[:, K // 2] += b
        return q.to(torch.float16)
    if transition == "uniform":        # fp16 from the start (ar_discrete.py:308-313)
        q = torch.full((K, K), beta_t / K).to(torch.float16)
        q.fill_diagonal_(1.0 - beta_t * (K - 1) / K)
        return q
    raise ValueError(f"unknown transition {transition!r} (expected 'absorbing' or 'uniform')")


@functools.lru_cache(maxsize=16)
def scalar_table(timesteps: int, K: int, transition: str) -> torch.Tensor:
    """float32 (S, TAB_STRIDE) table for vb200_q_sample / vb200_posterior_sample_from_logits."""
    betas = cosine_beta_schedule(timesteps + 1).to(torch.float16)     # ar_discrete.py:257
    m = K // 2
    a = 0 if m != 0 else 1
    b = 1 if m != 1 else 2
    tab = torch.zeros(timesteps, L.TAB_STRIDE, dtype=torch.float32)
    cum = None
    for t in range(timesteps):
        one = _onestep(betas[t], K, transition)
        cum = one if cum is None else torch.tensordot(cum, one, dims=[[1], [0]])  # fp16, :270-274
        tab[t, L.TAB_ONE_KEEP] = one[a, a]
        tab[t, L.TAB_ONE_OFF] = one[a, b]
        tab[t, L.TAB_ONE_ABSORB] = one[a, m]
        tab[t, L.TAB_ONE_BOTH] = one[m, m]
        picks = torch.stack([cum[a, a], cum[a, b], cum[a, m], cum[m, m]])
        tab[t, L.TAB_CUM_KEEP:L.TAB_CUM_BOTH + 1] = picks.float()
        # q_sample logits are formed in fp16: log(fp16(q) + eps) (ar_discrete.py:482)
        tab[t, L.TAB_LOG_KEEP:L.TAB_LOG_BOTH + 1] = torch.log(picks + EPS).float()
    return tab


@functools.lru_cache(maxsize=2)
def dense_log_qbar(timesteps: int, K: int, transition: str) -> torch.Tensor:
    """fp16 (S, K, K) ``log(Qbar_t + eps)`` — the logits ``q_sample`` forms (ar_discrete.py:482) from
    the fp16 chain product (:270-275), for ``vb200_q_sample_dense``.  Only the *uniform* transition
    needs it, and only for bit-exact parity runs with supplied uniforms: its K-term fp16 sums are not
    rank-structured to the last bit (up to 2 diagonal and 3 off-diagonal values per t, placed by the
    summation order of the GEMM), so `scalar_table` holds representative values there, exact ones
    for absorbing (<= 2 non-zero terms per entry)."""
    betas = cosine_beta_schedule(timesteps + 1).to(torch.float16)
    out = torch.empty(timesteps, K, K, dtype=torch.float16)
    cum = None
    for t in range(timesteps):
        one = _onestep(betas[t], K, transition)
        cum = one if cum is None else torch.tensordot(cum, one, dims=[[1], [0]])
        out[t] = torch.log(cum + EPS)
    return out


def betas_fp16(timesteps: int) -> torch.Tensor:
    return cosine_beta_schedule(timesteps + 1).to(torch.float16)
