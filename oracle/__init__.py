"""CPU oracle for the cnn-gp Gram recursion.  TEST INFRASTRUCTURE ONLY (see oracle.py)."""
