"""CPU oracle (test infrastructure only -- see oracle/stheno_oracle.py header)."""
