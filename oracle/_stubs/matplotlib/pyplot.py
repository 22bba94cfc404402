"""Empty stub of `matplotlib.pyplot` (TEST INFRASTRUCTURE ONLY)."""
