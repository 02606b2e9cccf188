"""Model/data config with the reference's schema (``vall_e/config.py:10-96``) plus the two keys
the D3PM denoiser needs: ``n_steps`` and ``transition``.  ``model`` accepts the
reference's names (``ar*``, ``nar*``) and ``diffusion[-quarter|-half]``.  ``cfg`` is created at
import from the command line, as in the reference (config.py:96)."""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import cached_property
from pathlib import Path

from .utils.config import Config as ConfigBase


@dataclass(frozen=True)
class Config(ConfigBase):
    data_root: Path = Path("data")
    data_dirs: list[Path] = field(default_factory=lambda: [])

    @property
    def sample_rate(self):
        return 24_000

    p_additional_prompt: float = 0.8
    max_prompts: int = 6

    max_num_val: int = 20
    max_val_ar_steps: int = 300

    token_dim: int = 256
    num_tokens: int = 1024

    nj: int = 8
    batch_size: int = 32
    eval_batch_size: int = 32
    warmup_min_lr: float = 1e-9
    warmup_max_lr: float = 1e-5
    dis_warmup_max_lr: float = 7e-5
    warmup_num_steps: int = 100
    max_iter: int = 1_000_000
    gradient_clipping: float = 1
    eval_every: int = 2_000
    save_ckpt_every: int = 2_000

    model: str = "ar-quarter"
    spkr_name_getter: str = "lambda p: p.parts[-1]"

    min_phones: int = 10
    max_phones: int = 50

    use_fp16: bool = True
    gradient_accumulation_steps: int = 1
    sampling_temperature: float = 1.0

    cache_dataloader: bool = False

    # --- D3PM denoiser (new keys)
    n_steps: int = 50              # diffusion timesteps S; the reverse loop runs t = S-1 .. 1
    transition: str = "absorbing"  # "absorbing" (ar_discrete.py:315-334) | "uniform" (:308-313)

    @cached_property
    def get_spkr(self):
        return eval(self.spkr_name_getter)


cfg = Config.from_cli()

if __name__ == "__main__":
    print(cfg)
