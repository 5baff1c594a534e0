from .two_columns import FirstTwoColunmnsPoseParameters, pose_to_d9, r6d2mat  # noqa: F401
from .axis_rotation import AxisRotationPoseParameters  # noqa: F401
from .quaternion import QuaternionsPoseParameters  # noqa: F401
