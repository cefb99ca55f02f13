"""Import shim: `from nets.PosePriorNetwork import PosePriorNetwork` (eval3d.py:38)."""
from hand3d_b200.nets.PosePriorNetwork import PosePriorNetwork  # noqa: F401
