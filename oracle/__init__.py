"""CPU oracle (test infrastructure only -- see reference_numpy.py header)."""
