"""CPU oracle (test infrastructure only -- see vismem_oracle.py header)."""
