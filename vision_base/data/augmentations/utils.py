"""Reference-compatible dotted name; the implementation lives in fsnet_b200."""
from fsnet_b200.data.augmentations import flip_relative_pose  # noqa: F401
