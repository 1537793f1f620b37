"""ctypes binding of libfsnet_b200.so (C ABI declared in include/fsnet_b200.h).

There is no CPU fallback: every call needs CUDA tensors and the built library; anything else raises.
"""
import ctypes
import os
import re
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSNET_B200_LIB") or os.path.join(_HERE, "lib", "libfsnet_b200.so")   # override: kernel-variant experiments
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fsnet_b200.h")

class WeightDesc(ctypes.Structure):
    """fsnet_weight_desc of include/fsnet_b200.h."""
    _fields_ = [("w", ctypes.c_void_p), ("fwd_hi", ctypes.c_void_p), ("fwd_lo", ctypes.c_void_p), ("dgrad_hi", ctypes.c_void_p),
                ("cout", ctypes.c_int), ("cin", ctypes.c_int), ("kh", ctypes.c_int), ("kw", ctypes.c_int),
                ("cout_pad", ctypes.c_int), ("cin_pad", ctypes.c_int)]


class WgradDesc(ctypes.Structure):
    """fsnet_wgrad_desc of include/fsnet_b200.h."""
    _fields_ = [("acc_off", ctypes.c_longlong), ("grad_off", ctypes.c_longlong), ("cout", ctypes.c_int), ("cin", ctypes.c_int),
                ("kh", ctypes.c_int), ("kw", ctypes.c_int), ("cout_pad", ctypes.c_int), ("cin_pad", ctypes.c_int)]


class Peer(ctypes.Structure):
    """fsnet_peer of include/fsnet_b200.h."""
    _fields_ = [("bufs", ctypes.c_void_p), ("flags", ctypes.c_void_p), ("rank", ctypes.c_int), ("world", ctypes.c_int),
                ("slot_off", ctypes.c_longlong), ("flag_off", ctypes.c_int), ("seq", ctypes.c_void_p)]


_lock = threading.Lock()
_lib = None
launch_count = 0          # entry-point calls issued through this binding
kernel_launches = 0       # CUDA kernels those calls launched (bench.py reports it as gpu_launches)
KERNELS_PER_ENTRY = {"fsnet_smooth_fwd": 2, "fsnet_mei_lut": 2, "fsnet_adam_step": 2}          # everything else launches exactly one kernel
_profiled = {}            # entry name -> list of (start_event, end_event) while profiling is on


def reset_counters():
    global launch_count, kernel_launches
    launch_count = 0
    kernel_launches = 0


_profile_tags = {}         # entry name -> function(args) -> tag stored with each timed launch


def profile_entry(name, on=True, tag=None):
    """Bracket every call of `name` with CUDA events on the launching stream (bench.py's roofline leg).
    ``tag(args)`` (optional) labels each launch, e.g. with the number of tensor-core products of fsnet_conv."""
    if on:
        _profiled[name] = []
        if tag is not None:
            _profile_tags[name] = tag
    else:
        _profiled.pop(name, None)
        _profile_tags.pop(name, None)


def profile_results(name, with_tags=False):
    """Per-launch durations in microseconds (synchronises)."""
    torch.cuda.synchronize()
    rows = _profiled.get(name, [])
    if with_tags:
        return [(r[0].elapsed_time(r[1]) * 1e3, r[2]) for r in rows]
    return [r[0].elapsed_time(r[1]) * 1e3 for r in rows]


class FsnetError(RuntimeError):
    pass


class View(ctypes.Structure):
    """fsnet_view of include/fsnet_b200.h."""
    _fields_ = [("ptr", ctypes.c_void_p), ("n", ctypes.c_int), ("h", ctypes.c_int), ("w", ctypes.c_int), ("c", ctypes.c_int),
                ("ring", ctypes.c_int), ("c_total", ctypes.c_int), ("c_off", ctypes.c_int)]


def declared_symbols():
    """Entry points declared in the public header (used by the CPU-side ABI test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(fsnet_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing=True):
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise FsnetError(f"{LIB_PATH} is missing: run `python -m fsnet_b200.build` (no CPU fallback exists)")
            from . import build as _build
            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        lib.fsnet_last_error.restype = ctypes.c_char_p
        lib.fsnet_abi_version.restype = ctypes.c_int
        _lib = lib
        return lib


def _ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, torch.Tensor):
        if not t.is_cuda:
            raise FsnetError("fsnet_b200 kernels take CUDA tensors only (there is no CPU path)")
        if not t.is_contiguous():
            raise FsnetError("fsnet_b200 kernels take contiguous tensors")
        return ctypes.c_void_p(t.data_ptr())
    raise TypeError(type(t))


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Invoke an int-returning entry point; tensors -> device pointers, ints/floats by value.
    Python floats are passed as C float, ints as C int.  The current torch stream is appended."""
    global launch_count, kernel_launches
    lib = load()
    fn = getattr(lib, name)
    cargs = []
    for a in args:
        if a is None or isinstance(a, torch.Tensor):
            cargs.append(_ptr(a))
        elif hasattr(a, "t") and isinstance(getattr(a, "t"), torch.Tensor):   # dense, non-'contiguous' (channels_last)
            if not a.t.is_cuda:
                raise FsnetError("fsnet_b200 kernels take CUDA tensors only (there is no CPU path)")
            cargs.append(ctypes.c_void_p(a.t.data_ptr()))
        elif isinstance(a, (View, Peer)):
            cargs.append(ctypes.byref(a))
        elif isinstance(a, bool):
            cargs.append(ctypes.c_int(int(a)))
        elif isinstance(a, int):
            cargs.append(ctypes.c_int(a))
        elif isinstance(a, float):
            cargs.append(ctypes.c_float(a))
        elif isinstance(a, (ctypes.c_void_p, ctypes.c_uint, ctypes.c_size_t, ctypes.c_longlong, ctypes.c_double)):
            cargs.append(a)
        else:
            raise TypeError(f"{name}: unsupported argument type {type(a)}")
    cargs.append(stream_ptr())
    rec = _profiled.get(name)
    if rec is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = fn(*cargs)
        ev1.record()
        tagger = _profile_tags.get(name)
        rec.append((ev0, ev1, tagger(args) if tagger else None))
    else:
        rc = fn(*cargs)
    launch_count += 1
    kernel_launches += KERNELS_PER_ENTRY.get(name, 1)
    if rc != 0:
        raise FsnetError(f"{name} failed ({rc}): {lib.fsnet_last_error().decode()}")
    return rc
