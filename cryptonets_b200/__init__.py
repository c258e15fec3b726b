"""cryptonets_b200: B200-native BFV engine behind the CryptoNets IFactory/IVector/IMatrix plugin API.

The arithmetic lives in libcnhe.so (hand-written sm_100a CUDA behind the C ABI of include/cnhe.h); this package is
the thin host-side mirror of the reference's C# interfaces.  Nothing here falls back to the CPU."""
from ._lib import CnheError, lib  # noqa: F401
