"""Reference-compatible dotted name; the implementation lives in fsnet_b200."""
from fsnet_b200.networks.loss_head import FishEyeDecoder, MonoDepth2Decoder  # noqa: F401
