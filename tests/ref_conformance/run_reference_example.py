"""
TEST INFRASTRUCTURE ONLY.  Runs one of the REFERENCE'S example scripts (read in place from
/root/reference/examples) against photon_weave_b200, with the same aliasing as
run_reference_tests.py.  Usage:
    python tests/ref_conformance/run_reference_example.py <script.py> [--backend host|cuda]
The script's module globals are returned to the caller of ``run`` for checks.
"""

import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference_tests as rrt  # noqa: E402  (sets up sys.path for the package and fake jax)

sys.path.insert(0, os.path.join(HERE, "fake_mpl"))


def run(script: str, backend: str = "host"):
    rrt._jax_style_updates()
    if not any(isinstance(f, rrt._Alias) for f in sys.meta_path):
        sys.meta_path.insert(0, rrt._Alias())
    if backend == "host":
        import host_shim

        with host_shim.host_backend():
            return runpy.run_path(script, run_name="__main__")
    return runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    backend = "host"
    argv = sys.argv[1:]
    if "--backend" in argv:
        i = argv.index("--backend")
        backend = argv[i + 1]
        del argv[i:i + 2]
    g = run(argv[0], backend)
    print("ok:", argv[0])
