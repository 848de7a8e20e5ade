"""Name -> constructor registries (model/registry.py:3-5, utils/registry.py:35)."""


class Registry(dict):
    def register(self, name, fn=None):
        if fn is not None:
            self[name] = fn
            return fn

        def deco(f):
            self[name] = f
            return f
        return deco


BACKBONES = Registry()
HEADS = Registry()
PREDICTOR = Registry()
