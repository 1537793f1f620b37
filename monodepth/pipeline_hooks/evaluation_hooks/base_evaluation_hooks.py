"""Reference-compatible dotted names; the implementation lives in fsnet_b200."""
from fsnet_b200.hooks.evaluation import KittiEvaluationHook, FastNuscEvaluationHook  # noqa: F401
