"""The C++ host layer above the C-ABI (include/modelardb_cuda.hpp: the reference's function and operator names in the
language class of the reference) built with g++ and run against the oracle on a GPU box (tests/cpp/host_api_test.cc).

Every section of the program (including HOST_API_EXTENDED: GridStream's predicate / time range / limit, GROUP BY tags,
batched finished buffers) runs on the GPU against the CUDA library, and on the CPU with the oracle behind the C-ABI entry
points (tests/cpp/cabi_on_oracle.cc), which checks the header's own host logic under ASan / UBSan."""
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
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DHOST_API_EXTENDED", os.path.join(ROOT, "tests", "cpp", "host_api_test.cc"), "-o", exe,
                           "-L" + pkg, "-lmodelardb_cuda", "-L" + orc, "-lmdb_oracle", "-pthread",
                           "-Wl,-rpath," + pkg, "-Wl,-rpath," + orc])
    return exe


def test_cpp_host_layer_compiles_and_links(tmp_path):
    """CPU: the header is valid C++17 and every C-ABI symbol it uses resolves against the built library."""
    _build(tmp_path)


def test_cpp_host_layer_logic_on_the_oracle(tmp_path):
    """CPU: the same program with tests/cpp/cabi_on_oracle.cc (the C-ABI entry points implemented on the oracle) linked
    instead of the CUDA library, under AddressSanitizer and UBSan: the C++ operators' own logic -- leftovers, slicing,
    tags, predicate, time range, limit, grouping by tags, batching of finished buffers -- without a GPU."""
    from oracle import mdb_oracle
    mdb_oracle.lib()
    exe = str(tmp_path / "host_api_on_oracle")
    orc = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-g", "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                           "-DHOST_API_EXTENDED",
                           os.path.join(ROOT, "tests", "cpp", "host_api_test.cc"), os.path.join(ROOT, "tests", "cpp", "cabi_on_oracle.cc"),
                           "-o", exe, "-L" + orc, "-lmdb_oracle", "-pthread", "-Wl,-rpath," + orc])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout and "runtime error" not in out.stderr, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_host_layer_matches_oracle(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout, out.stdout
