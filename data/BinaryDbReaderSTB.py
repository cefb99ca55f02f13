from hand3d_b200.data.BinaryDbReader import BinaryDbReaderSTB  # noqa: F401  (import shim: `from data.BinaryDbReaderSTB import *`)
