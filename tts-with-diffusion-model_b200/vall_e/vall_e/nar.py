"""NAR: the non-causal AdaLN configuration of ``Base`` (reference ``vall_e/vall_e/nar.py:8-26``)
and its level-by-level inference loop (``nar.py:76-99``), running on the B200 kernels."""
from __future__ import annotations

import torch
from torch import Tensor

from .base import Base


class NAR(Base):
    @property
    def n_resp_levels(self):
        return 7

    @property
    def casual(self):
        return False

    @property
    def use_stop_token(self):
        return False

    @property
    def norm_type(self):
        return "adaln"

    @property
    def resp_loss_only(self):
        return True

    def forward(self, text_list: list[Tensor], proms_list: list[Tensor], resps_list: list[Tensor],
                sampling_temperature: float = 0.2):
        """resps_list: [t'' l]; with l < 8 known levels, fills levels l .. 7 one forward pass per
        level (AdaLN row = level being predicted - 1) and returns [t'' 8].  Errors as the reference:
        ``ValueError`` when utterances carry different numbers of levels (nar.py:44-47)."""
        n_levels_set = {r.shape[-1] for r in resps_list}
        if len(n_levels_set) > 1:
            raise ValueError(f"Please give only one level, got {n_levels_set}.")
        n_levels = next(iter(n_levels_set))
        if n_levels == self.n_resp_levels + 1:
            raise NotImplementedError("NAR training step (8 given levels) is outside the B200 inference path")
        device = text_list[0].device
        prev_list = resps_list
        while True:
            level = prev_list[0].shape[-1] - 1
            if level >= self.n_resp_levels:
                break
            quant_levels = torch.full((len(text_list),), level, device=device)
            resp_list = super().forward(text_list, proms_list, prev_list, return_all_resp=True,
                                        shift_targ_list=False, quant_levels=quant_levels,
                                        sampling_temperature=sampling_temperature)
            prev_list = [torch.cat([rs, r.to(rs).unsqueeze(-1)], dim=-1) for rs, r in zip(prev_list, resp_list)]
        return prev_list
