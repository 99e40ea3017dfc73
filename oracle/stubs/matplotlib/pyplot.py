"""Empty stand-in for matplotlib.pyplot (only imported, never called on the RK45 path)."""
