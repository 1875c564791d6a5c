"""ORACLE — test infrastructure only (CPU restatements of the reference's hot path).
Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg only."""
