"""CPU oracle for the DEP-GAN hot path -- TEST INFRASTRUCTURE ONLY (pinned to the executed reference source through oracle/keras_shim.py, see depgan_oracle.py).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
