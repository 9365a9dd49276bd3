"""CPU oracle for the short-term / mid-term feature path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product (``pyaudioanalysis_b200``) never imports
this package and fails loudly when its CUDA library is missing.
"""
