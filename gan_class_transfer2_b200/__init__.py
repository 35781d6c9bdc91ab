"""B200-native training step for relgukxilef/GAN-Class-Transfer2's train.py.

  _lib    ctypes binding of libgct2_b200.so (C ABI: include/gct2_b200.h; CUDA sources: csrc/)
  ops     tensor-level wrappers over the C ABI
  engine  HBM-resident state + launch sequence of one training step (optionally a CUDA graph, optionally data parallel)
  train   drop-in host surface with train.py's names: Denoiser, Trainer, WarmUp, alpha_dash, optimizer, ...
"""
__all__ = ["_lib", "ops", "engine", "train"]
