"""Model factory with the reference's names and sizes (``vall_e/vall_e/__init__.py:7-59``):
``ar*`` -> AR, ``nar*`` -> NAR, ``diffusion*`` -> the D3PM denoising sampler; ``-quarter``
(d=256, 4 heads), ``-half`` (d=512, 8 heads), default full (d=1024, 16 heads), 12 layers each."""
from ..config import cfg
from .ar import AR
from .diffusion import Diffusion
from .nar import NAR

_SIZES = {"-quarter": dict(d_model=256, n_heads=4, n_layers=12),
          "-half": dict(d_model=512, n_heads=8, n_layers=12),
          "": dict(d_model=1024, n_heads=16, n_layers=12)}


def get_model(name: str):
    name = name.lower()
    if name.startswith("ar"):
        Model, extra = AR, {}
    elif name.startswith("nar"):
        Model, extra = NAR, {}
    elif name.startswith("diffusion"):
        Model, extra = Diffusion, dict(n_steps=cfg.n_steps, transition=cfg.transition)
    else:
        raise ValueError("Model name should start with AR, NAR or diffusion.")
    for suffix in ("-quarter", "-half"):
        if suffix in name:
            return Model(cfg.num_tokens, **_SIZES[suffix], **extra)
    if name not in ["ar", "nar", "diffusion"]:
        raise NotImplementedError(name)
    return Model(cfg.num_tokens, **_SIZES[""], **extra)
