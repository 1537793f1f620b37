"""fsnet_b200 -- B200-native (sm_100a) implementation of the FSNet self-supervised depth training step.

Package layout (only what the hot path needs):
  csrc/          hand-written CUDA kernels + the C ABI (include/fsnet_b200.h)
  _lib.py        ctypes binding of that ABI (fails loudly when the library or a GPU is missing)
  functional.py  torch.autograd Functions over the ABI
  networks/      host-side mirror of the reference's module surface (same class names / state-dict keys)
  utils/, hooks/, data/   builder, config loader, training hook, synthetic triplet dataset
"""
__version__ = "0.1.0"
