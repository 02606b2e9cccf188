"""Host-side mirror of the reference transformer ``vall_e/vall_e/base.py``.

The classes below keep the reference's names, constructor arguments and *state-dict keys*
(``sep, text_emb.weight, proms_emb.weight, resps_emb.weight, blocks.{i}.attn.block.to_qkv.weight,
blocks.{i}.attn.block.to_out.{weight,bias}, blocks.{i}.attn.norm.emb.weight,
blocks.{i}.ffn.block.{0,3}.{weight,bias}, blocks.{i}.ffn.norm.emb.weight, classifier.{weight,bias}``)
so checkpoints — including whole-module pickles written by the reference's ``export.py`` — load
unchanged.  They are parameter containers: the arithmetic of their reference ``forward`` methods
(base.py:103-133, 145-158, 184-194, 221-234, 255-274) runs in libvalle_b200.so, driven by
``vall_e.b200.engine.DenoiserEngine``.  ``Base.forward`` keeps the reference call convention
(lists of per-utterance tensors in, sampled tokens out; base.py:403-499) for inference; the
training branch (``targ_list``) and the causal AR decode are out of scope.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor, nn
from torch.distributions import Categorical

from ..b200 import lib as L
from ..b200.engine import BatchLayout, DenoiserEngine, PackedWeights, check_ids


class SinusodialEmbedding(nn.Module):
    """Parameter-free; ``omega`` is a non-persistent buffer as in base.py:38-46."""

    def __init__(self, d_model):
        super().__init__()
        assert d_model % 2 == 0, "Only support even d_model."
        self.d_model = d_model
        d_half = d_model // 2
        omega = torch.exp(-math.log(1e4) * (torch.arange(d_half, dtype=torch.float32) / d_half))
        self.register_buffer("omega", omega, persistent=False)


class Attention(nn.Module):
    """Weights of base.py:92-101: ``to_qkv`` (3d, d) without bias, ``to_out`` (d, d) with bias."""

    def __init__(self, d_model, n_heads, casual):
        super().__init__()
        assert d_model % n_heads == 0
        self.casual = casual
        self.n_heads = n_heads
        self.scale = (d_model // n_heads) ** -0.5
        self.to_qkv = nn.Linear(d_model, d_model * 3, bias=False)
        self.to_out = nn.Linear(d_model, d_model)


class AdaLN(nn.Module):
    """``emb`` (n_levels, 2d) = [log gamma | beta], zero-initialised (base.py:136-143)."""

    def __init__(self, d_model, n_levels, eps=1e-5, k=0.1, c=2):
        super().__init__()
        self.eps, self.k, self.c = eps, k, c
        self.emb = nn.Embedding(n_levels, d_model * 2)
        nn.init.zeros_(self.emb.weight)


class PrenormResidual(nn.Module):
    """``block`` + ``norm`` (+ dropout, identity at inference) — base.py:161-182."""

    def __init__(self, block, d_model, p_dropout, requires_mask=False, norm_type="ln", n_levels=None):
        super().__init__()
        self.block = block
        self.requires_mask = requires_mask
        self.norm_type = norm_type
        if norm_type == "ln":
            self.norm = nn.LayerNorm(d_model)
        elif norm_type == "adaln":
            assert n_levels is not None
            self.norm = AdaLN(d_model, n_levels)
        else:
            raise NotImplementedError(norm_type)
        self.dropout = nn.Dropout(p_dropout)


class Block(nn.Sequential):
    """attn + ffn sub-layers with the reference's child names (base.py:197-219)."""

    def __init__(self, d_model, n_heads, p_dropout, casual, norm_type, n_levels):
        super().__init__()
        self.attn = PrenormResidual(Attention(d_model, n_heads, casual), d_model=d_model, p_dropout=p_dropout,
                                    requires_mask=True, norm_type=norm_type, n_levels=n_levels)
        ffn = nn.Sequential(nn.Linear(d_model, d_model * 4), nn.GELU(), nn.Dropout(p_dropout),
                            nn.Linear(d_model * 4, d_model))
        self.ffn = PrenormResidual(ffn, d_model=d_model, p_dropout=p_dropout, norm_type=norm_type,
                                   n_levels=n_levels)


class Embedding(nn.Embedding):
    """Text table (base.py:237-241); rows are gathered by vb200_embed_gather."""


class MultiEmbedding(nn.Module):
    """(levels, tokens, d) table whose per-level rows are summed (base.py:244-253)."""

    def __init__(self, max_n_levels, n_tokens, token_dim):
        super().__init__()
        self.max_n_levels = max_n_levels
        self.n_tokens = n_tokens
        self.weight = nn.Parameter(torch.randn(max_n_levels, n_tokens, token_dim))


class Base(nn.Module):
    @property
    def casual(self) -> bool:
        raise NotImplementedError

    @property
    def n_resp_levels(self) -> int:
        raise NotImplementedError

    @property
    def use_stop_token(self) -> bool:
        raise NotImplementedError

    @property
    def norm_type(self):
        raise NotImplementedError

    @property
    def n_prom_levels(self) -> int:
        return 8

    @property
    def resp_loss_only(self):
        raise NotImplementedError

    # The two places the D3PM glue class differs from the reference constructor
    @property
    def n_norm_levels(self) -> int:
        """rows of every AdaLN table (reference: n_resp_levels, base.py:348)."""
        return self.n_resp_levels

    def _n_classifier_out(self, n_resp_tokens: int) -> int:
        return n_resp_tokens

    def __init__(self, n_tokens: int, d_model: int = 512, n_heads: int = 8, n_layers: int = 12,
                 p_dropout: float = 0.1):
        super().__init__()
        self.n_tokens = n_tokens
        n_resp_tokens = n_tokens + (1 if self.use_stop_token else 0)
        self.text_emb = Embedding(n_tokens, d_model)
        self.proms_emb = MultiEmbedding(self.n_prom_levels, n_tokens, d_model)
        self.resps_emb = MultiEmbedding(self.n_resp_levels, n_resp_tokens, d_model)
        self.sin_emb = SinusodialEmbedding(d_model)
        self.sep = nn.Parameter(torch.randn(d_model))
        self.blocks = nn.ModuleList([
            Block(d_model=d_model, n_heads=n_heads, p_dropout=p_dropout, casual=self.casual,
                  norm_type=self.norm_type, n_levels=self.n_norm_levels)
            for _ in range(n_layers)
        ])
        self.classifier = nn.Linear(d_model, self._n_classifier_out(n_resp_tokens))

    @property
    def stop_token(self):
        if not self.use_stop_token:
            raise ValueError("Not using stop token!")
        return self.n_tokens

    @property
    def ignore_index(self):
        return -100

    # ------------------------------------------------------------------ engine plumbing
    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_b200", None)
        return state

    def _load_from_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_b200", None)          # weights changed: repack on next use
        return super()._load_from_state_dict(*args, **kwargs)

    @property
    def n_heads(self) -> int:
        return self.blocks[0].attn.block.n_heads

    def engine(self) -> DenoiserEngine:
        """Packs the current parameters for the kernels (cached; call ``refresh_engine`` after
        modifying parameters in place)."""
        dev = self.sep.device
        eng = self.__dict__.get("_b200")
        if eng is None or eng.w.device != dev:
            if dev.type != "cuda":
                raise L.VB200Error("the denoiser runs on CUDA only (sm_100a kernels, no CPU fallback); "
                                   "move the model with .to('cuda')")
            if self.casual or self.use_stop_token:
                raise NotImplementedError("causal / stop-token (autoregressive) models are outside the "
                                          "B200 hot path; only the non-causal denoiser is implemented")
            weights = PackedWeights(self.state_dict(), self.n_heads, len(self.blocks), self.norm_type, dev)
            eng = DenoiserEngine(weights)
            self.__dict__["_b200"] = eng
        return eng

    def refresh_engine(self):
        self.__dict__.pop("_b200", None)

    def _logits(self, text_list, proms_list, resps_list, levels: Tensor, use_time: bool,
                logits_dtype=torch.float32):
        """Response-row logits, list of (t'', n_out) — base.py:427-443 + :491 on the packed layout."""
        eng = self.engine()
        lay = BatchLayout(text_list, proms_list, [len(r) for r in resps_list], eng.w.device,
                          n_text=eng.w.text_w.shape[0], n_codes=eng.w.K)
        ws = eng.workspace(lay, logits_dtype=logits_dtype)
        resp = torch.cat([r.reshape(len(r), -1) for r in resps_list]).to(eng.w.device, torch.int32).contiguous()
        check_ids(resp, eng.w.K, "resps_list")
        lv = levels.to(eng.w.device, torch.int32).contiguous()
        logits = eng.forward(lay, ws, resp, lv, use_time=use_time)
        return lay.split_resp(logits)

    def forward(self, text_list: list[Tensor], proms_list: list[Tensor], resps_list: list[Tensor],
                targ_list: list[Tensor] | None = None, quant_levels: Tensor | None = None,
                shift_targ_list: bool = False, return_all_resp: bool = False,
                sampling_temperature: float = 1.0):
        """Same arguments and return value as the reference ``Base.forward`` (base.py:403-499),
        inference branch: tokens sampled from ``Categorical(logits / temperature)`` for every
        response position (``return_all_resp``) or for the last one."""
        if targ_list is not None:
            raise NotImplementedError("loss computation (training) is outside the B200 inference path")
        if quant_levels is None:
            if self.norm_type == "adaln":
                raise ValueError("quant_levels is required for AdaLN models")
            quant_levels = torch.zeros(len(text_list), dtype=torch.long)
        logits = self._logits(text_list, proms_list, resps_list, quant_levels, use_time=False)
        if return_all_resp:
            return [Categorical(logits=h / sampling_temperature).sample() for h in logits]
        last = torch.stack([h[-1] for h in logits])
        return Categorical(logits=last / sampling_temperature).sample()
