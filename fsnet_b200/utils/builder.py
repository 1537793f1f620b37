"""String -> object plugin builder and its combinators (reference: vision_base/utils/builder.py:5-72)."""
from typing import Callable, Dict, List

import numpy as np

from .config import find_object


def build(name, *args, **kwargs):
    return find_object(name)(*args, **kwargs)


class _Chain(object):
    def __init__(self, cfg_list: List[Dict], **common_keywords):
        self.children: List[Callable] = [build(**{**common_keywords, **item}) for item in cfg_list]

    @staticmethod
    def _run(children, *args, **kwargs):
        result = None
        for i, child in enumerate(children):
            if i == 0:
                result = child(*args, **kwargs)
            elif isinstance(result, tuple):
                result = child(*result)
            else:
                result = child(result)
        return result


class Sequential(_Chain):
    """Children run in order, each fed the previous result (tuples are splatted)."""

    def __call__(self, *args, **kwargs):
        return self._run(self.children, *args, **kwargs)


class Shuffle(_Chain):
    """As Sequential but in a fresh random order on every call (np.random.permutation)."""

    def draw(self):
        return np.random.permutation(len(self.children))

    def __call__(self, *args, **kwargs):
        return self._run([self.children[i] for i in self.draw()], *args, **kwargs)


class Parallel(_Chain):
    """Every child gets the same inputs; results are returned as a list."""

    def __call__(self, *args, **kwargs):
        return [child(*args, **kwargs) for child in self.children]
