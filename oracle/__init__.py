"""CPU oracle for the -icl hot path. Test infrastructure: see oracle/icl_oracle.py."""
