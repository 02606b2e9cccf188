"""``vall_e.emb`` — the EnCodec (``qnt``) and g2p front / back ends of ``python -m vall_e``.

These stay on the REFERENCE implementation (BASELINE.json: "EnCodec encode/decode in ``vall_e/emb``
stays on the reference path and is excluded from timing"), and they are not vendored: this package
extends its module search path with the ``vall_e/emb`` directory of a checkout of the reference, so
``from vall_e.emb import qnt, g2p`` executes the reference's own ``qnt.py`` / ``g2p.py``
(``emb/qnt.py:18-76``, ``emb/g2p.py:12-28``) as sub-modules of THIS package — their
``from ..config import cfg`` resolves to this package's config (same schema, ``sample_rate`` 24 kHz).

The checkout is found through ``VALL_E_REF`` (the repository root, its ``vall_e`` package or the
``emb`` directory itself), else next to this repository (``../reference``, ``/root/reference``).
Their third-party imports (encodec, torchaudio, soundfile, g2p_en) must be installed.
"""
import os
from pathlib import Path


def _candidates():
    env = os.environ.get("VALL_E_REF")
    roots = [Path(env)] if env else []
    here = Path(__file__).resolve()
    roots += [here.parents[3].parent / "reference", Path("/root/reference")]
    for r in roots:
        for sub in ("vall_e/emb", "emb", "."):
            yield (r / sub).resolve()


def reference_emb_dir():
    """Directory holding the reference's qnt.py and g2p.py, or None."""
    for d in _candidates():
        if (d / "qnt.py").is_file() and (d / "g2p.py").is_file() and d != Path(__file__).resolve().parent:
            return d
    return None


_dir = reference_emb_dir()
if _dir is not None:
    __path__.append(str(_dir))      # noqa: F821  (package attribute)


def __getattr__(name):
    if name in ("qnt", "g2p"):
        if reference_emb_dir() is None:
            raise ImportError(
                "vall_e.emb needs the reference's vall_e/emb/{qnt,g2p}.py (EnCodec / g2p are outside the "
                "accelerated path and are not vendored): set VALL_E_REF to a checkout of "
                "csulb-datascience/TTS-with-Diffusion-model")
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
