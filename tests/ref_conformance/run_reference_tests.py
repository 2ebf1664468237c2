"""
TEST INFRASTRUCTURE ONLY.  Runs the REFERENCE'S OWN test-suite (/root/reference/tests, read in
place, never copied) against photon_weave_b200:

* ``photon_weave.*`` imports are aliased to ``photon_weave_b200.*`` (same module objects),
* ``jax`` is the stand-in under fake_jax/ (input construction / comparisons only),
* with ``--backend host`` (default without a GPU) the package's ``core.ffi`` is routed to the
  NumPy oracle (tests/host_shim.py), so the run checks the HOST logic of the containers;
  with ``--backend cuda`` the real libpwb200 kernels run.

Usage: python tests/ref_conformance/run_reference_tests.py [--backend host|cuda] [pytest args]
Prints pytest's summary; exit code is pytest's.
"""

import importlib
import importlib.abc
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF_TESTS = "/root/reference/tests"

sys.path.insert(0, os.path.join(HERE, "fake_jax"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

_RENAMES = {  # module names the reference spells differently
    "photon_weave.operation.helpers.fock_dimension_esitmation":
        "photon_weave_b200.operation.helpers.fock_dimension_estimation",
}


class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name == "photon_weave" or name.startswith("photon_weave."):
            return importlib.util.spec_from_loader(name, self)
        return None

    def create_module(self, spec):
        real = _RENAMES.get(spec.name, "photon_weave_b200" + spec.name[len("photon_weave"):])
        return importlib.import_module(real)

    def exec_module(self, module):
        pass


def _jax_style_updates():
    """``x.at[idx].set(v)`` on arrays the package hands back (functional update, jax idiom)."""
    import torch

    class _At:
        def __init__(self, t):
            self.t = t

        def __getitem__(self, idx):
            self.idx = idx
            return self

        def set(self, value):
            out = self.t.clone()
            out[self.idx] = value
            return out

    torch.Tensor.at = property(lambda self: _At(self))


def main(argv):
    _jax_style_updates()
    backend = "host"
    if "--backend" in argv:
        i = argv.index("--backend")
        backend = argv[i + 1]
        del argv[i:i + 2]
    sys.meta_path.insert(0, _Alias())
    import pytest

    args = ["-q", "-p", "no:cacheprovider", "--rootdir", REF_TESTS,
            "-c", os.devnull] + list(argv)
    if not any(a.startswith("/") or a.startswith("tests") for a in argv):
        args.append(REF_TESTS)
    if backend == "host":
        import host_shim

        with host_shim.host_backend():
            return pytest.main(args)
    return pytest.main(args)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
