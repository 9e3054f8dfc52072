"""CPU oracle for the DEP-GAN hot path -- TEST INFRASTRUCTURE ONLY (parity unpinned, see depgan_oracle.py).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
