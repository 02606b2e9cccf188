"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

CPU restatement (torch-CPU / numpy, dense K x K matrices, fp16 tables exactly as the
reference keeps them) of the D3PM algebra in the reference file
``vall_e/vall_e/ar_discrete.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module; the
product path (``tts-with-diffusion-model_b200/``) never does.

Pinning: the reference ships no tests or golden vectors, so this restatement
is pinned against outputs of the reference's own functions run in the build container —
``tests/golden/make_golden.py`` imports ``/root/reference/vall_e/vall_e/ar_discrete.py`` and
writes ``tests/golden/d3pm_*.npz``; ``tests/test_oracle_golden.py`` checks this file against
them (integers bit-exact, fp16 tables bit-exact).

The class count ``K`` and the absorbing index ``K // 2`` are parameters here (the reference
hard-codes 1025 / 512, ``ar_discrete.py:255,309,328,332``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1.0e-6  # ar_discrete.py:276


def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> torch.Tensor:
    """ar_discrete.py:286-304.  Note the reference's ``linspace(0, steps, steps)`` (sic)."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    acp = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    acp = acp / acp[0]
    betas = 1 - (acp[1:] / acp[:-1])
    return torch.from_numpy(np.clip(betas, a_min=0, a_max=0.999))


def absorbing_onestep(beta_t: float, K: int) -> torch.Tensor:
    """ar_discrete.py:315-334: Q_t = (1-b) I + b 1 e_m^T with m = K // 2 (float64)."""
    mat = np.diag(np.full((K,), 1.0 - beta_t, dtype=np.float64), k=0)
    mat[:, K // 2] += beta_t
    return torch.from_numpy(mat)


def uniform_onestep(beta_t: torch.Tensor, K: int) -> torch.Tensor:
    """ar_discrete.py:308-313 (``create_transition_matrix``; built directly in fp16)."""
    mat = torch.full((K, K), beta_t / K).to(torch.float16)
    idx = torch.arange(K)
    mat[idx, idx] = 1.0 - beta_t * (K - 1) / K
    return mat


class D3PM:
    """Holds the reference's tensors: ``betas`` (fp16), ``q_onestep_mats``, ``q_mats``,
    ``transpose_q_onestep_mats`` (all fp16, shape (S, K, K)) — ar_discrete.py:257-277."""

    def __init__(self, timesteps: int, K: int, transition: str = "absorbing"):
        self.timesteps = timesteps
        self.K = K
        self.transition = transition
        self.eps = EPS
        self.betas = cosine_beta_schedule(timesteps + 1).to(torch.float16)  # :257
        if transition == "absorbing":
            one = [absorbing_onestep(self.betas[t].numpy(), K) for t in range(timesteps)]
        elif transition == "uniform":
            one = [uniform_onestep(self.betas[t], K) for t in range(timesteps)]
        else:
            raise ValueError(transition)
        self.q_onestep_mats = torch.stack(one, dim=0).to(torch.float16)  # :268-269
        q = self.q_onestep_mats[0]
        qs = [q]
        for t in range(1, timesteps):  # :270-274, fp16 chain product
            q = torch.tensordot(q, self.q_onestep_mats[t], dims=[[1], [0]])
            qs.append(q)
        self.q_mats = torch.stack(qs, dim=0).to(torch.float16)
        self.transpose_q_onestep_mats = torch.transpose(self.q_onestep_mats, 1, 2).contiguous()

    # ---- ar_discrete.py:337-345
    def _at(self, a, t, x):
        B, W = x.shape
        a_t = torch.index_select(a, dim=0, index=t)
        onehot = F.one_hot(x.view(B, -1).to(torch.int64), num_classes=self.K).to(torch.float16)
        return torch.matmul(onehot, a_t).view(B, W, self.K)

    # ---- ar_discrete.py:377-400
    def _at_onehot(self, a, t, x):
        B, W, _ = x.shape
        a_t = torch.index_select(a, dim=0, index=t)
        return torch.matmul(x.view(B, -1, self.K), a_t).view(B, W, self.K)

    # ---- ar_discrete.py:489-502
    def q_probs(self, x_start, t):
        return self._at(self.q_mats, t, x_start)

    # ---- ar_discrete.py:467-487, with the uniforms an argument instead of torch.rand
    def q_sample(self, x_start, t, mask, noise):
        logits = torch.log(self.q_probs(x_start, t) + self.eps)
        noise = torch.clamp(noise, min=torch.finfo(noise.dtype).tiny, max=1.0)
        g = -torch.log(-torch.log(noise))
        return torch.argmax(logits + g, dim=-1) * mask

    # ---- ar_discrete.py:347-375 (x_start_logits=True branch is the one p_sample uses)
    def q_posterior_logits(self, x_start_logits, x_t, t):
        fact1 = self._at(self.transpose_q_onestep_mats, t, x_t)
        t_1 = torch.where(t == 0, t, t - 1)
        fact2 = self._at_onehot(self.q_mats, t_1, F.softmax(x_start_logits, dim=-1))
        out = torch.log(fact1 + self.eps) + torch.log(fact2 + self.eps)
        tb = torch.reshape(t, [out.shape[0]] + [1] * (out.dim() - 1))
        return torch.where(tb == 0, x_start_logits, out)

    # ---- ar_discrete.py:401-420, uniforms as an argument; ``greedy`` drops the Gumbel term
    def p_sample(self, model_logits, t, x, noise=None, greedy=False):
        pred = model_logits
        tb = torch.reshape(t, [pred.shape[0]] + [1] * (pred.dim() - 1))
        post = torch.where(tb == 0, pred, self.q_posterior_logits(pred, x, t))
        if greedy:
            return torch.argmax(post, dim=-1), post
        nz = (t != 0).to(x.dtype).reshape(x.shape[0], *([1] * x.dim()))
        noise = torch.clamp(noise, min=torch.finfo(noise.dtype).tiny, max=1.0)
        g = -torch.log(-torch.log(noise))
        return torch.argmax(post + nz * g, dim=-1), post

    # ---- per-t scalar view of the fp16 tables (what a closed form needs)
    def scalars(self):
        """Returns dict of float32 arrays of length S read straight out of the fp16 tables.
        absorbing: onestep keep/absorb/both, cumulative keep/absorb/other/both
        uniform:   onestep diag/off,            cumulative diag/off"""
        K, m = self.K, self.K // 2
        a = 0 if m != 0 else 1          # some class that is not the mask
        b = 1 if m != 1 else 2          # another one
        one, cum = self.q_onestep_mats.float().numpy(), self.q_mats.float().numpy()
        if self.transition == "absorbing":
            return dict(
                one_keep=one[:, a, a], one_absorb=one[:, a, m], one_both=one[:, m, m],
                cum_keep=cum[:, a, a], cum_absorb=cum[:, a, m], cum_other=cum[:, a, b],
                cum_both=cum[:, m, m])
        return dict(one_diag=one[:, a, a], one_off=one[:, a, b],
                    cum_diag=cum[:, a, a], cum_off=cum[:, a, b])


def posterior_fp32(logits, x_t, t, d3pm: D3PM):
    """fp32 evaluation of the *same dense algorithm* (softmax @ Q̄_{t-1}, Q_t column) with the
    reference's fp16 tables upcast — the reference point for the CUDA closed form's KL test."""
    K = d3pm.K
    p0 = torch.softmax(logits.float(), dim=-1)
    B, W = x_t.shape
    t1 = torch.where(t == 0, t, t - 1)
    f2 = torch.matmul(p0, d3pm.q_mats.float()[t1])
    f1 = d3pm.transpose_q_onestep_mats.float()[t][torch.arange(B)[:, None], x_t.long()]
    return torch.log(f1 + EPS) + torch.log(f2 + EPS)
