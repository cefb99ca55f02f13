"""Import shim: `from nets.ColorHandPose3DNetwork import ColorHandPose3DNetwork` (run.py:26, eval2d.py:35)."""
from hand3d_b200.nets.ColorHandPose3DNetwork import ColorHandPose3DNetwork  # noqa: F401
