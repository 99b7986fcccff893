"""oracle/ -- TEST INFRASTRUCTURE ONLY (see oracle/chroma_oracle.c header).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arms may import this package."""
