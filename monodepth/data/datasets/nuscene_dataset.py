"""Reference-compatible dotted name; the implementation lives in fsnet_b200."""
from fsnet_b200.data.nuscenes_json import NusceneJsonDataset  # noqa: F401
