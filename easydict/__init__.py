"""Minimal stand-in for the ``easydict`` package (absent from this image, no network).

The reference's configs are Python files that build ``EasyDict`` trees
(``configs/kitti_wpose_example:1-5``); ``cfg_from_file`` asserts the type.  Only the behaviour
those configs and ``update_cfg`` rely on is provided: attribute <-> item aliasing and recursive
conversion of nested dicts (also inside lists / tuples).
"""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        if d is None:
            d = {}
        if kwargs:
            d = dict(d, **kwargs)
        for k, v in d.items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError:
            raise AttributeError(k)

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k, v in d.items():
            self[k] = v


__all__ = ["EasyDict"]
