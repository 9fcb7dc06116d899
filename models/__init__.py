"""Drop-in `models` package: put this repository first on PYTHONPATH and the reference's
train.py / eval.py / evalDTU.py (`from models import *`, reference train.py:15) pick up the B200
implementation.  See INTEGRATION.md."""
from scene_3dreconstruction_mvsnet_b200.models import MVSNet, mvsnet_loss  # noqa: F401
from scene_3dreconstruction_mvsnet_b200.models import module, mvsnet  # noqa: F401

__all__ = ["MVSNet", "mvsnet_loss"]
