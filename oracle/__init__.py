"""CPU oracle for the per-frame hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  ``librir_b200`` never does: the product has no CPU fallback.
"""
from .oracle import *  # noqa: F401,F403
