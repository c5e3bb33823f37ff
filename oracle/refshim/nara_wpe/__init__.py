"""Stand-in for the un-vendored ``nara_wpe`` (setup.py:142 pins >=0.0.6).
Backed by the oracle's restatement (oracle/gss_oracle.py); WPE parity is
therefore UNPINNED (see oracle/__init__.py)."""
from . import wpe, utils  # noqa: F401
