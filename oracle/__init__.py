"""CPU checkers for the parity tests.  TEST INFRASTRUCTURE ONLY -- see oracle/README.md."""
