"""CPU oracle package -- test infrastructure only (see equiprop_oracle.py header)."""
