"""Reference-compatible dotted name; the implementation lives in fsnet_b200."""
from fsnet_b200.utils.logger import AverageMeter, LogImageStruct, LossLogger, styling_git_info  # noqa: F401
