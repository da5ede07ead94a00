"""gr_dvbt_b200 — B200 (sm_100a) receive hot path behind gr-dvbt's block API.

The product is the C-ABI shared library gr_dvbt_b200/libdvbt_b200.so (include/dvbt_b200.h)
plus the gr::block shims in gr_dvbt_b200/shim/.  This Python package is only the thin
ctypes host layer that tests and bench.py use; it mirrors the reference's block
constructors (dvbt.viterbi_decoder(...), ...) and has no CPU fallback: importing
`gr_dvbt_b200.capi` fails loudly if the library is missing.
"""
from .capi import lib, DvbtError  # noqa: F401
from .blocks import *  # noqa: F401,F403
