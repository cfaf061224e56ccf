"""CPU oracle (test infrastructure only - see oracle/nem_oracle.py header)."""
