"""The few helpers of the reference's ``vall_e/utils/utils.py`` the inference path touches
(``to_device`` :reference utils.py, ``load_state_dict_non_strict`` utils.py:55-75).  Training-side
helpers (logging, diagnostics, artifacts) are out of scope."""
from __future__ import annotations

from typing import Any, Callable

import torch
from torch import Tensor


def tree_map(fn: Callable, x: Any, only_tensor: bool = True):
    if isinstance(x, dict):
        return {k: tree_map(fn, v, only_tensor) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(tree_map(fn, v, only_tensor) for v in x)
    if not only_tensor or isinstance(x, Tensor):
        return fn(x)
    return x


def to_device(x: Any, device):
    return tree_map(lambda t: t.to(device), x)


def load_state_dict_non_strict(model: torch.nn.Module, state_dict: dict, logger=None):
    """Loads the entries whose name AND shape agree; reports the rest (reference utils.py:55-75)."""
    own = model.state_dict()
    usable = {k: v for k, v in state_dict.items() if k in own and own[k].shape == v.shape}
    if logger is not None:
        extra = set(state_dict) - set(usable)
        missing = set(own) - set(usable)
        if extra:
            logger.warning(f"Extra parameters are found. Provided but not required parameters: \n{extra}.")
        if missing:
            logger.warning(f"Some parameters are missing. Required but not provided parameters: \n{missing}.")
    model.load_state_dict(usable, strict=False)
