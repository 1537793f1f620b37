"""Reference-compatible dotted name; the implementation lives in fsnet_b200."""
from fsnet_b200.utils.config import (cfg_from_file, update_cfg, find_object, set_random_seed,  # noqa: F401
                                     get_num_parameters)
