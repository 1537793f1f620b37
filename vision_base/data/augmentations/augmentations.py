"""Reference-compatible dotted names; the implementation lives in fsnet_b200."""
from fsnet_b200.data.augmentations import (ConvertColor, ConvertToFloat, ConvertToTensor, Copy, EmptyAug, Normalize,  # noqa: F401
                                           RandomBrightness, RandomContrast, RandomMirror, RandomSaturation,
                                           RandomWarpAffine, Resize)
