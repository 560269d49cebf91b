"""CPU oracle for the TPDM hot path -- test infrastructure only (see sd3_oracle.py header)."""
