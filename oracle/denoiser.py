"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

fp32 torch-CPU restatement of the reference denoiser forward, ``vall_e/vall_e/base.py``
(non-causal AdaLN configuration = ``vall_e/vall_e/nar.py:8-26``), written functionally over a
state dict with the reference's key names, plus the one glue step SURVEY.md §7.1 adds for the
D3PM denoiser (``time_emb`` row added to the response rows; AdaLN table indexed by timestep;
classifier with 8 K-way heads).  Padded dense (B, T_max, d) tensors, explicit masks and the
materialised (b, i, j, h) attention scores are kept on purpose — this follows the reference
algorithm, not the product's packed layout.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm
may import this.  Pinned against the reference modules themselves by
``tests/golden/make_golden.py`` -> ``tests/golden/denoiser_*.npz`` (``test_oracle_golden.py``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def sinusoidal_pe(n: int, d_model: int) -> torch.Tensor:
    """base.py:38-89: pe[p] = [sin(p w_i) || cos(p w_i)], w_i = exp(-ln(1e4) i / (d/2))."""
    d_half = d_model // 2
    omega = torch.exp(-math.log(1e4) * (torch.arange(d_half, dtype=torch.float32) / d_half))
    x = omega[None, :] * torch.arange(n)[:, None]
    return torch.cat([x.sin(), x.cos()], dim=-1)


def multi_embedding(weight: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """base.py:255-274: one-hot(x) (t l' k), zero-pad levels to l, einsum 'l k d, n l k -> n d'."""
    L, K, _ = weight.shape
    oh = F.one_hot(x, num_classes=K)
    oh = F.pad(oh, (0, 0, 0, L - oh.shape[1])).to(weight)
    return torch.einsum("l k d, n l k -> n d", weight, oh)


def join(parts, sep):
    """base.py:277-286."""
    ret = parts[0]
    for p in parts[1:]:
        ret = torch.cat((ret, sep[None], p), dim=0)
    return ret


def adaln(x, emb_weight, l, eps=1e-5, k=0.1, c=2):
    """base.py:145-158."""
    logg, beta = emb_weight[l].unsqueeze(1).chunk(2, dim=-1)
    h = F.layer_norm(x, x.shape[-1:], eps=eps)
    h = c * (1 - (k * h)) * h
    return logg.exp() * h + beta


def attention(x, m, w_qkv, w_out, b_out, n_heads, casual=False):
    """base.py:103-133."""
    b, t, d = x.shape
    dh = d // n_heads
    q, k, v = F.linear(x, w_qkv).chunk(3, dim=-1)
    q, k, v = (z.reshape(b, t, n_heads, dh) for z in (q, k, v))
    e = torch.einsum("b i h d, b j h d -> b i j h", q, k) * dh ** -0.5
    kpm = m.unsqueeze(1) * m.unsqueeze(2)
    if casual:
        kpm = kpm.squeeze(-1).tril().unsqueeze(-1)
    e = e.masked_fill(kpm == 0, -torch.finfo(e.dtype).max)
    a = e.softmax(dim=2)
    o = torch.einsum("b i j h, b j h d -> b i h d", a, v).flatten(-2)
    return F.linear(o, w_out, b_out) * m


def block(x, m, l, sd, prefix, n_heads, norm_type="adaln", casual=False):
    """base.py:184-194 (PrenormResidual) x2, base.py:221-234 (Block)."""
    def norm(z, which):
        if norm_type == "adaln":
            return adaln(z, sd[f"{prefix}.{which}.norm.emb.weight"], l)
        return F.layer_norm(z, z.shape[-1:], sd[f"{prefix}.{which}.norm.weight"],
                            sd[f"{prefix}.{which}.norm.bias"], eps=1e-5)
    a = attention(norm(x, "attn") * m, m,
                  sd[f"{prefix}.attn.block.to_qkv.weight"],
                  sd[f"{prefix}.attn.block.to_out.weight"],
                  sd[f"{prefix}.attn.block.to_out.bias"], n_heads, casual)
    x = (x + a) * m
    h = norm(x, "ffn") * m
    h = F.linear(h, sd[f"{prefix}.ffn.block.0.weight"], sd[f"{prefix}.ffn.block.0.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[f"{prefix}.ffn.block.3.weight"], sd[f"{prefix}.ffn.block.3.bias"])
    return (x + h) * m


def base_forward_logits(sd, text_list, proms_list, resps_list, levels, n_heads, n_layers,
                        time_t=None, norm_type="adaln", casual=False, return_hidden=False):
    """base.py:427-443 up to (and including) ``h = classifier(x) * m`` and the un-padding.

    ``levels``: LongTensor (b,) — the AdaLN row per utterance (reference: quant level;
    D3PM glue: timestep).  ``time_t``: optional LongTensor (b,), adds ``time_emb.weight[t]`` to
    the response rows.  Returns list of (T_b, n_out) logits rows (all positions).
    """
    d = sd["sep"].shape[0]
    x_list = []
    for i, (tx, pr, rs) in enumerate(zip(text_list, proms_list, resps_list)):
        te = sd["text_emb.weight"][tx]
        pe = multi_embedding(sd["proms_emb.weight"], pr)
        re_ = multi_embedding(sd["resps_emb.weight"], rs)
        if time_t is not None:
            re_ = re_ + sd["time_emb.weight"][time_t[i]][None]
        x_list.append(join((te, pe, re_), sd["sep"]))
    lens = [len(z) for z in x_list]
    T = max(lens)
    x = torch.stack([F.pad(z, (0, 0, 0, T - len(z))) for z in x_list])
    m = (torch.arange(T)[None, :] < torch.tensor(lens)[:, None]).float().unsqueeze(-1)
    x = x + sinusoidal_pe(T, d)[None]
    hidden = []
    for i in range(n_layers):
        x = block(x, m, levels, sd, f"blocks.{i}", n_heads, norm_type, casual)
        if return_hidden:
            hidden.append(x)
    h = F.linear(x, sd["classifier.weight"], sd["classifier.bias"]) * m
    out = [hi[:li] for hi, li in zip(h, lens)]
    if return_hidden:
        return out, [[hh[b, :lens[b]] for b in range(len(lens))] for hh in hidden]
    return out


def diffusion_logits(sd, text_list, proms_list, xt_list, t, n_heads, n_layers, n_levels=8):
    """D3PM denoiser logits for the response rows: list of (T_r, n_levels, K)."""
    rows = base_forward_logits(sd, text_list, proms_list, xt_list, t, n_heads, n_layers, time_t=t)
    K = sd["classifier.weight"].shape[0] // n_levels
    return [r[-len(x):].reshape(len(x), n_levels, K) for r, x in zip(rows, xt_list)]


def random_state_dict(n_tokens, d_model, n_layers, n_adaln_rows, n_resp_levels=8, n_out=None,
                      seed=0, adaln_std=0.02, time_rows=None, bf16_round=True):
    """Random weights with the reference's initialisers (base.py:253,339, nn.Linear/Embedding
    defaults), AdaLN table perturbed to N(0, adaln_std) so the path is exercised.
    ``bf16_round`` makes every value bf16-representable so oracle and CUDA see identical weights."""
    g = torch.Generator().manual_seed(seed)
    d = d_model

    def lin(o, i, bias=True):
        bound = 1 / math.sqrt(i)
        w = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        b = (torch.rand(o, generator=g) * 2 - 1) * bound if bias else None
        return w, b

    sd = {
        "sep": torch.randn(d, generator=g),
        "text_emb.weight": torch.randn(n_tokens, d, generator=g),
        "proms_emb.weight": torch.randn(8, n_tokens, d, generator=g),
        "resps_emb.weight": torch.randn(n_resp_levels, n_tokens, d, generator=g),
    }
    if time_rows:
        sd["time_emb.weight"] = torch.randn(time_rows, d, generator=g)
    for i in range(n_layers):
        p = f"blocks.{i}"
        sd[f"{p}.attn.block.to_qkv.weight"], _ = lin(3 * d, d, bias=False)
        sd[f"{p}.attn.block.to_out.weight"], sd[f"{p}.attn.block.to_out.bias"] = lin(d, d)
        sd[f"{p}.attn.norm.emb.weight"] = torch.randn(n_adaln_rows, 2 * d, generator=g) * adaln_std
        sd[f"{p}.ffn.block.0.weight"], sd[f"{p}.ffn.block.0.bias"] = lin(4 * d, d)
        sd[f"{p}.ffn.block.3.weight"], sd[f"{p}.ffn.block.3.bias"] = lin(d, 4 * d)
        sd[f"{p}.ffn.norm.emb.weight"] = torch.randn(n_adaln_rows, 2 * d, generator=g) * adaln_std
    n_out = n_out or n_tokens
    sd["classifier.weight"], sd["classifier.bias"] = lin(n_out, d)
    if bf16_round:
        sd = {k: v.bfloat16().float() for k, v in sd.items()}
    return sd
