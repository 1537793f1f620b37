"""Reference-compatible dotted name; the implementation lives in fsnet_b200."""
from fsnet_b200.networks.meta_archs import DistillWPoseMeta, MonoDepthMeta, MonoDepthWPose  # noqa: F401
