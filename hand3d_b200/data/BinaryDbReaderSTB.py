"""Import shim: the reference keeps the STB reader in its own module (data/BinaryDbReaderSTB.py)."""
from .BinaryDbReader import BinaryDbReaderSTB  # noqa: F401
