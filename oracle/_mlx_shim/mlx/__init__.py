"""TEST INFRASTRUCTURE ONLY -- numpy stand-in for the third-party `mlx` package.

The reference (Acelogic/LTX-2-MLX) delegates all arithmetic to `mlx==0.30.1`
(uv.lock:537-538), which is not installed here and cannot be installed (no
network, no Apple/Linux wheel in the wheelhouse).  This shim restates the
*published semantics* of exactly the mlx primitives the reference's hot path
calls (see SURVEY.md section 8(c) for the call sites) on top of numpy, so that
the reference's OWN Python code (op order, layouts, constants, reshapes) can be
executed in this container by `tests/golden/make_golden.py` to produce golden
vectors.  It is never imported by the product package, bench.py's GPU arm or
the C-ABI library.
"""
from . import core  # noqa: F401
from . import nn  # noqa: F401
