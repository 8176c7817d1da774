from . import helpers, layers, registry, vision_transformer  # noqa: F401
