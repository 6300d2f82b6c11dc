"""String/type keyed operator registry: the drop-in boundary of the reference
(stgp/dispatch.py:133-189).  `@dispatch(*keys)` registers a function (its __name__ is the first
key), `evoke(name, *keys)` returns the registered callable.

When the reference package itself is importable the b200 backends register straight into ITS
registry; this module is the stand-alone equivalent used by the host-side
mirror in this repo, with the same calling convention.
"""
import inspect


class DispatchNotFound(Exception):
    pass


_REGISTRY = []  # list of (keys tuple, obj)


def _key_name(k):
    if isinstance(k, str):
        return k
    if inspect.isclass(k):
        return k.__name__
    return type(k).__name__


def _matches(reg_key, call_key):
    """A registered key matches a call key if the names agree or the call key's type derives
    from the registered class."""
    if isinstance(reg_key, str) or isinstance(call_key, str):
        return _key_name(reg_key) == _key_name(call_key)
    call_cls = call_key if inspect.isclass(call_key) else type(call_key)
    reg_cls = reg_key if inspect.isclass(reg_key) else type(reg_key)
    return issubclass(call_cls, reg_cls)


def _specificity(reg_keys):
    score = 0
    for k in reg_keys:
        if inspect.isclass(k):
            score += len(k.__mro__)
    return score


def dispatch(*keys):
    def deco(obj):
        full = ((obj.__name__,) + keys) if inspect.isfunction(obj) else keys
        _REGISTRY.append((full, obj))
        return obj
    return deco


def evoke(*keys):
    best, best_score = None, -1
    for reg_keys, obj in _REGISTRY:
        if len(reg_keys) != len(keys):
            continue
        if all(_matches(r, c) for r, c in zip(reg_keys, keys)):
            s = _specificity(reg_keys)
            if s > best_score:
                best, best_score = obj, s
    if best is None:
        raise DispatchNotFound("Cannot evoke %r" % (keys,))
    return best
