"""
photon_weave_b200 -- B200 (sm_100a) implementation of photon_weave's
state-evolution hot path behind the reference's array-level seam
(``photon_weave.core.adapters``).  CUDA only; no CPU fallback.
"""

__version__ = "0.1.0"
