"""Stub of `matplotlib` (TEST INFRASTRUCTURE ONLY): flow_vis.py:1 imports pyplot at module import."""
