from .dla_dcn import build_backbone  # noqa: F401  (model/backbone/__init__.py:1)
