"""CPU oracle package (test infrastructure only -- see oracle/sf_oracle.h)."""
