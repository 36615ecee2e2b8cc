"""Name -> object lookup with the reference's contract (utils/generic_utils.py:43-84):
case-insensitive member lookup in a module; classes are instantiated with the parsed
``params`` list, module-level instances (``raw``, ``simple_char_parser``) are returned as-is.
Reference module names ('preprocessing.audio', 'core.models', ...) map onto this package."""
import importlib
import inspect

from .hparams import HParams

_ALIASES = {"preprocessing.audio": "asr_study_b200.preprocessing.audio",
            "preprocessing.text": "asr_study_b200.preprocessing.text",
            "core.models": "asr_study_b200.core.models",
            "core.layers": "asr_study_b200.core.layers"}


def inspect_module(module):
    mod = importlib.import_module(_ALIASES.get(module, module))
    return {k: v for k, v in inspect.getmembers(mod)
            if getattr(v, "__module__", None) == mod.__name__ or
            (not inspect.ismodule(v) and not inspect.isclass(v) and not inspect.isfunction(v)
             and getattr(type(v), "__module__", None) == mod.__name__)}


def get_from_module(module, name, params=None):
    if name is None or str(name).lower() == "none":
        return None
    members = {k.lower().strip(): v for k, v in inspect_module(module).items()}
    key = str(name).lower().strip()
    if key not in members:
        raise KeyError("%s not found in %s.\n Valid values are: %s" % (name, module, ", ".join(sorted(members))))
    member = members[key]
    if member and params is not None and inspect.isclass(member):
        return member(**HParams().parse(params).values())
    return member
