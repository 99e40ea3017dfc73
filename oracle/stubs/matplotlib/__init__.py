"""Import stub for the reference's `import matplotlib.pyplot as plt` (rkf45.py:2).
Test infrastructure only."""
