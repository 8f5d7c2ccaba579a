"""timm.models.registry.register_model: records the factory and returns it unchanged."""
_model_entrypoints = {}


def register_model(fn):
    _model_entrypoints[fn.__name__] = fn
    return fn
