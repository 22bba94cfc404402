"""Stub of `colorama` for running the reference offline (TEST INFRASTRUCTURE ONLY).

The reference's logger (src/logger.py:5) imports colorama for terminal colours; the real package is
not installed and there is no network.  Colour codes become empty strings.
"""


class _Codes:
    def __getattr__(self, name):
        return ""


Fore = _Codes()
Back = _Codes()
Style = _Codes()


def init(*args, **kwargs):
    return None
