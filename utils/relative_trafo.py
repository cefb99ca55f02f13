"""Import shim: `from utils.relative_trafo import *` (nets/PosePriorNetwork.py:25 of the reference)."""
from hand3d_b200.utils.relative_trafo import bone_rel_trafo_inv, kinematic_chain_dict, kinematic_chain_list  # noqa: F401
