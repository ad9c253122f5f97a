"""Minimal attribute-dict with the semantics the reference config relies on.

The reference builds its configuration with the third-party ``easydict``
package (``/root/reference/config/indoor_config.py:1``), which is not installed
in this image.  The hot path needs exactly three behaviours from it:
attribute access (``cfg.patch.num_fps``), item access (``cfg["data"]["dataset"]``,
``/root/reference/models/BUFFERX.py:156``) and ``dict.get`` with a default
(``cfg.match.get("enable_early_exit", True)``, ``BUFFERX.py:296``).  Nested plain
dicts are converted recursively on assignment, like easydict does.
"""


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        src = dict(d or {})
        src.update(kw)
        for k, v in src.items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, cls):
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
        except KeyError as e:
            raise AttributeError(k) from e

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def update(self, other=None, **kw):
        src = dict(other or {})
        src.update(kw)
        for k, v in src.items():
            self[k] = v

    def copy(self):
        return EasyDict(dict(self))
