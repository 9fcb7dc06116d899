from scene_3dreconstruction_mvsnet_b200.models.mvsnet import *  # noqa: F401,F403
from scene_3dreconstruction_mvsnet_b200.models.mvsnet import MVSNet, mvsnet_loss, FeatureNet, CostRegNet  # noqa: F401
