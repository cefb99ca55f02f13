"""TEST INFRASTRUCTURE ONLY (see oracle/hand3d_oracle.py).  PARITY UNPINNED: the reference has no
golden vectors and TensorFlow 1.3 cannot run in this image; pinned by hand-derived KATs only."""
