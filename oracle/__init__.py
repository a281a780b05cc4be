"""CPU oracle for the dispersion forward path.  TEST INFRASTRUCTURE ONLY (see oracle.py)."""
