"""No-op stand-in for gin-config (decorators only). Test infrastructure."""


def _passthrough(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]

    def deco(obj):
        return obj
    return deco


register = _passthrough
configurable = _passthrough
REQUIRED = object()
