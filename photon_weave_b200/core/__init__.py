"""
Array-only seam (reference: photon_weave/core/__init__.py:1-6): functions here
"operate purely on arrays plus lightweight metadata".
"""
