"""TEST INFRASTRUCTURE: the checkers (see oracle/checker.py)."""
