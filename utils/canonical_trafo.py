from hand3d_b200.utils.canonical_trafo import canonical_trafo, flip_right_hand  # noqa: F401  (import shim)
