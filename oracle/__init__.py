"""CPU oracle (test infrastructure only; parity unpinned -- see maf_numpy.py)."""
