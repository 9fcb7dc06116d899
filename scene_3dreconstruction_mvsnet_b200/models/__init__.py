"""Host-side mirror of the reference's `models` package (models/__init__.py:1)."""
from .mvsnet import MVSNet, mvsnet_loss  # noqa: F401
