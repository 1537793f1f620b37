"""Glue between the reference-shaped modules and the tcgen05 executor (fsnet_b200/engine.py).

``ResNet.forward`` does not run anything: it returns ``LazyFeatures``.  The head that consumes them
(DepthDecoder / PoseDecoder) runs encoder + decoder as ONE autograd node.  Code that really wants the five
feature tensors (``feats[i]``, ``len(feats)``, iteration) gets them: the list materialises itself from the
executor's planes on first access -- as detached EXPORT copies (no autograd, BatchNorm statistics untouched);
the heads ignore those copies and always execute the network from the image.
"""
import torch

from .. import _lib


def available() -> bool:
    try:
        lib = _lib.load()
    except Exception:  # noqa: BLE001
        return False
    return hasattr(lib, "fsnet_conv") and hasattr(lib, "fsnet_conv_wgrad")


class LazyFeatures(list):
    """Deferred output of ``ResNet.forward`` on the tcgen05 path."""

    def __init__(self, backbone, image):
        super().__init__()
        self.backbone, self.image = backbone, image
        self.materialized = False

    def _materialize(self):
        if not self.materialized:
            from ..engine import Tape, resnet_forward
            with torch.no_grad():
                tape = Tape({}, self.backbone.training, False)
                # statistics must not be updated twice: run on a throw-away copy of the running buffers
                bufs = [b for _, b in self.backbone.named_buffers()]
                saved = [b.clone() for b in bufs]
                acts = resnet_forward(tape, self.backbone, self.image, False)
                for b, v in zip(bufs, saved):
                    b.copy_(v)
            super().extend(a.planes.to_float()[:, :a.c].contiguous() for a in acts)
            self.materialized = True

    def __getitem__(self, i):
        self._materialize()
        return super().__getitem__(i)

    def __iter__(self):
        self._materialize()
        return super().__iter__()

    def __len__(self):
        return 1 + self.backbone.num_stages


def runner_for(decoder, backbone):
    from ..engine import Runner
    cache = decoder.__dict__.setdefault("_tc_runners", {})
    r = cache.get(id(backbone))
    if r is None:
        r = Runner(backbone, decoder)
        cache[id(backbone)] = r
    return r
