from . import base, config, misc  # noqa: F401
