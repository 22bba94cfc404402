"""Stub of `skimage` (TEST INFRASTRUCTURE ONLY) -- see transform.py."""
