"""Key/value hyper-parameter bag with the reference's parsing contract
(utils/hparams.py:48-63): a dict, a python-literal string, or a flat
``[k, v, k, v, ...]`` list whose values go through ``ast.literal_eval``."""
import ast


class HParams(object):
    def __init__(self, **kv):
        object.__setattr__(self, "_kv", dict(kv))

    def __getattr__(self, key):            # missing keys read as None, like the reference
        return object.__getattribute__(self, "_kv").get(key)

    __getitem__ = __getattr__

    def __setattr__(self, key, value):
        self._kv[key] = value

    def update(self, other):
        self._kv.update(other)
        return self

    def parse(self, values):
        if isinstance(values, dict):
            return self.update(values)
        if isinstance(values, (list, tuple, set)):
            values = list(values)
            out = {}
            for k, v in zip(values[::2], values[1::2]):
                try:
                    out[k] = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    out[k] = v
            return self.update(out)
        return self.update(ast.literal_eval(values))

    def values(self):
        return self._kv

    def __str__(self):
        return str(self._kv)
