"""pytest plumbing for running fsnet_b200's Python side against the SIMT-emulated library (tests/host_emulation/simt.h) on CPU
tensors: the ctypes binding is pointed at the emulated .so, host pointers are accepted, the stream is null."""
import ctypes

import torch

from . import emulate


def install(monkeypatch):
    from fsnet_b200 import _lib
    lib = ctypes.CDLL(emulate.build_simt_library())
    lib.fsnet_last_error.restype = ctypes.c_char_p
    lib.fsnet_abi_version.restype = ctypes.c_int
    monkeypatch.setattr(_lib, "_lib", lib)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))      # _lib refuses host tensors: there is no CPU path
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)    # staging tables of FusedAdam
    return lib
