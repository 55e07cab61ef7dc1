"""pyrallis: `wrap()` builds the decorated function's config dataclass from `--key value` pairs; `field` -> dataclasses."""
import dataclasses
import functools
import inspect
import sys
import typing


def field(default=dataclasses.MISSING, default_factory=dataclasses.MISSING, is_mutable=False, **k):
    if default is not dataclasses.MISSING and (is_mutable or isinstance(default, (list, dict))):
        d = default
        return dataclasses.field(default_factory=lambda: type(d)(d))
    if default_factory is not dataclasses.MISSING:
        return dataclasses.field(default_factory=default_factory)
    return dataclasses.field(default=default)


def _cast(text, current):
    if isinstance(current, bool):
        return text.lower() in ("1", "true", "yes")
    if isinstance(current, int) and not isinstance(current, bool):
        return int(text)
    if isinstance(current, float):
        return float(text)
    if isinstance(current, (list, tuple)):
        import ast
        return type(current)(ast.literal_eval(text))
    return text


def wrap(*a, **k):
    def deco(fn):
        cls = next(iter(typing.get_type_hints(fn).values()), None) or next(iter(inspect.signature(fn).parameters.values())).annotation

        @functools.wraps(fn)
        def run(argv=None):
            argv = sys.argv[1:] if argv is None else argv
            cfg = cls()
            for key, val in zip(argv[::2], argv[1::2]):
                name = key.lstrip("-")
                setattr(cfg, name, _cast(val, getattr(cfg, name)))
            return fn(cfg)
        return run
    return deco
