from . import base, utils  # noqa: F401
