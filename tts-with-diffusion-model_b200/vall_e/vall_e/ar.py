"""AR: import-path placeholder for the reference's causal VALL-E AR model
(``vall_e/vall_e/ar.py:86-169``) so pickled ``zoo/ar.pt`` checkpoints resolve their class.

The autoregressive decode loop is not part of the non-autoregressive codec-token path this
repo accelerates (OUT OF SCOPE); calling it raises.  The D3PM helpers that the
reference pasted (unwired) into its ``ar.py`` (:170-440) live in ``diffusion.py`` / ``d3pm.py``.
"""
from __future__ import annotations

from .base import Base


class AR(Base):
    @property
    def n_resp_levels(self):
        return 1

    @property
    def casual(self):
        return True

    @property
    def use_stop_token(self):
        return True

    @property
    def norm_type(self):
        return "ln"

    @property
    def resp_loss_only(self):
        return False

    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            "causal autoregressive decoding is outside the B200 denoising-sampler path; run the AR "
            "stage with the reference implementation, or use model: diffusion")
