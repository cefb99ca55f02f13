from hand3d_b200.data.BinaryDbReader import BinaryDbReader  # noqa: F401  (import shim: `from data.BinaryDbReader import *`)
