"""Where the network modules' arithmetic lives.

The modules of this package (ResNet, DepthDecoder, PoseDecoder, ConvBnReLU ...) only hold parameters in the reference's names /
state-dict layout.  Their arithmetic is the tcgen05 executor: ``ResNet.forward`` returns a deferred handle and the consuming head
runs encoder + decoder as ONE autograd node of hand-written kernels (fsnet_b200/engine.py).  There is no second back-end in the
product and no CPU path: CPU tensors raise.

``COMPARATOR`` is a test / tooling hook: ``tools/torch_reference_backend.enable()`` plugs stock-PyTorch forwards into the same
module tree (``bench.py --backend torch``, feature-map comparisons in the tests).  It is None unless a tool sets it.
"""
COMPARATOR = None          # set by tools/torch_reference_backend.enable(); never by product code

NO_CPU = ("fsnet_b200 has no CPU / eager-PyTorch path: the network runs as tcgen05 kernels on CUDA tensors (fsnet_b200/engine.py). "
          "For a stock-PyTorch comparison call tools/torch_reference_backend.enable() first.")


def tc_available() -> bool:
    """True when the tcgen05 convolution kernels are built into the library."""
    try:
        from . import ops_tc  # noqa: F401
        return ops_tc.available()
    except ImportError:
        return False


def precision_note() -> str:
    if COMPARATOR is None:
        return "tcgen05 bf16x3 split operands, fp32 accumulate"
    return "cuDNN fp32 (comparison path, tools/torch_reference_backend.py)"


def set_backend(name: str) -> None:
    """Kept for callers of round 1: "tc" is the only product back-end (clears a comparator a tool may have plugged in)."""
    global COMPARATOR
    if name != "tc":
        raise ValueError("the product has one back-end (tcgen05); the cuDNN comparator is tools/torch_reference_backend.enable()")
    COMPARATOR = None
