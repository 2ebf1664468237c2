"""
TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the photon_weave
state-evolution hot path.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the *checker* (or the timed CPU
baseline), never as the thing being measured or shipped.  The product package
``photon_weave_b200`` never imports this package and fails loudly when its CUDA
library is missing.

Why a restatement and not the reference itself: the reference's arithmetic
lives in un-vendored third-party code (jax 0.8.1 / jaxlib 0.8.1 /
opt_einsum 3.4.0, pinned in /root/reference/poetry.lock) that is not installed
in this image and cannot be (no network).  The restatement follows the
reference's own files line by line (each function cites file:line) and is
pinned against every golden vector / known-answer test the reference's test
suite holds for this path (see ``tests/test_oracle_golden.py`` and
``tests/golden/``), plus the Random123 known-answer vectors for threefry2x32.

Parity status: PINNED for the seam-level behaviours the reference tests assert
(kron-equivalence, literal arrays, seed -> outcome KATs).  Bulk random states at
the BASELINE shapes are unpinned by the reference itself (its tests never exceed
16 amplitudes); there the oracle is its own authority and is cross-checked by an
independent dense ``kron`` construction at small sizes.
"""

from . import einsum_strings, kernels_np, adapters_np, rng_threefry, operators, layout_np  # noqa: F401
