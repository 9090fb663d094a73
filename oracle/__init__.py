"""ORACLE - test infrastructure only (see oracle/lidar4d_oracle.py header)."""
