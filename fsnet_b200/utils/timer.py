"""``@profile`` and ETA ``Timer`` (reference: vision_base/utils/timer.py:5-78)."""
import os
import time
from functools import wraps

import torch


def profile(name, profile_start=0, profile_end=1):
    """When env DEBUGGING is 1/true, time calls [0, profile_end) with a device synchronise on both
    sides and print from call ``profile_start`` on."""
    def decorator(func):
        func.call_times = 0

        @wraps(func)
        def wrapped(*args, **kwargs):
            on = os.environ.get("DEBUGGING", "").lower() in ("1", "true")
            if on and func.call_times < profile_end:
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                t0 = time.time()
                result = func(*args, **kwargs)
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                if func.call_times >= profile_start:
                    print(f"{name} takes {time.time() - t0} seconds at call time {func.call_times}")
                func.call_times += 1
                return result
            return func(*args, **kwargs)
        return wrapped
    return decorator


class Timer:
    def __init__(self):
        self.init = time.time()

    def compute_eta(self, current_iter, total_iter):
        elapsed = time.time() - self.init
        eta = elapsed / max(current_iter, 1) * (total_iter - current_iter)
        h, rem = divmod(int(eta), 3600)
        m, s = divmod(rem, 60)
        return f"{h}h:{m}m:{s}s"
