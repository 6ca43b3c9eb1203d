"""Test-infrastructure shim: the reference imports matplotlib.pyplot at module scope (RVMC.py:7)
but only touches it inside plotting helpers that are out of scope."""
