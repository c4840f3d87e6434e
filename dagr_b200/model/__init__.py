from .dagr import DAGR  # noqa: F401
