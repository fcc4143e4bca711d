"""B200-native per-frame surfel fusion hot path (DenseSurfelMapping drop-in).

The product is the C-ABI shared library built from ``csrc/`` (see ``include/dsm.h``);
this Python package is only the thin host-side harness around it (ctypes binding,
synthetic frame generator, element dtypes) used by tests and ``bench.py``.
"""
from .elements import SEED_DTYPE, SURFEL_DTYPE, num_seeds  # noqa: F401
