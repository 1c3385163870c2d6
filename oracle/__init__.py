"""CPU oracle for the stain hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline
legs of ``bench.py`` may import it, and only as the checker / the reported CPU baseline.
"""
