"""TEST INFRASTRUCTURE ONLY: CPU oracle for the gr-dvbt receive hot path.

oracle.port     - ctypes binding of the plain-C restatement (oracle/libdvbt_oracle.so)
oracle.refchain - ctypes driver of the reference's own sources compiled verbatim
                  (oracle/_ref/libdvbt_ref.so; only where it was built)
Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
"""
