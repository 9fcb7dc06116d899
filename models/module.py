from scene_3dreconstruction_mvsnet_b200.models.module import *  # noqa: F401,F403
from scene_3dreconstruction_mvsnet_b200.models.module import homo_warping, depth_regression  # noqa: F401
