"""Empty stand-in: R/src/assembly/ScatterFormData.py:2 imports matplotlib.pyplot and never uses it."""
