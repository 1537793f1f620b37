"""Reference-compatible dotted name; the implementation lives in fsnet_b200."""
from fsnet_b200.networks.depth_decoder import (DepthDecoder, MultiChannelDepthDecoder,  # noqa: F401
                                               MultiChannelDepthDecoderUncertain)
