"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatements of the reference algorithms).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
Nothing under pointcloudmatters_b200/ imports this package.
"""
