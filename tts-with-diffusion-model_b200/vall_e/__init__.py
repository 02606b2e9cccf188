"""B200-native drop-in for the D3PM denoising-sampler path of csulb-datascience/TTS-with-Diffusion-model.

Import paths mirror the reference package (``vall_e.vall_e.{base,nar,ar}``, ``vall_e.config``,
``python -m vall_e``) so pickled checkpoints and YAML configs keep working; the compute runs in
``vall_e.b200`` (ctypes -> libvalle_b200.so, hand-written sm_100a CUDA).
"""
