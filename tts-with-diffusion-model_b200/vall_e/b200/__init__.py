"""B200 (sm_100a) backend: ctypes binding of libvalle_b200.so and the denoiser engine."""
