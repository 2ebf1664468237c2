"""
The XLA typed-FFI shim (photon_weave_b200/csrc/pw_xla_ffi.cc) cannot be built in this image
(no jaxlib headers), so it is at least TYPE-CHECKED: compiled with -fsyntax-only against a
stand-in header (tests/xla_ffi_stub) whose XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts that
every handler is callable with exactly the argument list its binding declares; and the symbol
table of core/jax_ffi.py must name exactly the handlers the source defines.
"""

import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SRC = os.path.join(ROOT, "photon_weave_b200", "csrc", "pw_xla_ffi.cc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_shim_type_checks_against_the_ffi_surface():
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "xla_ffi_stub"),
           "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, SRC]
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA headers not found")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]


def test_registration_table_names_every_handler():
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", open(SRC).read()))
    table = open(os.path.join(ROOT, "photon_weave_b200", "core", "jax_ffi.py")).read()
    named = set(re.findall(r'"pw_\w+":\s*"(\w+)"', table))
    assert defined and defined == named
