"""The C++ host layer above the C-ABI (include/modelardb_cuda.hpp: the reference's function and operator names in the
language class of the reference) built with g++ and run against the oracle on a GPU box (tests/cpp/host_api_test.cc)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from modelardb_rs_b200 import _native
    from oracle import mdb_oracle
    mdb_oracle.lib()  # builds oracle/libmdb_oracle.so if needed
    exe = str(tmp_path / "host_api_test")
    pkg, orc = os.path.dirname(_native.LIB_PATH), os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "host_api_test.cc"), "-o", exe,
                           "-L" + pkg, "-lmodelardb_cuda", "-L" + orc, "-lmdb_oracle", "-pthread",
                           "-Wl,-rpath," + pkg, "-Wl,-rpath," + orc])
    return exe


def test_cpp_host_layer_compiles_and_links(tmp_path):
    """CPU: the header is valid C++17 and every C-ABI symbol it uses resolves against the built library."""
    _build(tmp_path)


@pytest.mark.gpu
def test_cpp_host_layer_matches_oracle(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout, out.stdout
