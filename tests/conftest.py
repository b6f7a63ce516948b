import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def harness_lib():
    """tests/native/vm_harness.cpp compiled for the host: the product's rule compiler, prefilter
    tables and Pike-VM source running on the CPU (test-only; never linked into the product)."""
    import ctypes as C
    out_dir = os.path.join(ROOT, "tests", "native", "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libvmharness.so")
    csrc = os.path.join(ROOT, "vainplex_openclaw_b200", "csrc")
    srcs = [os.path.join(ROOT, "tests", "native", "vm_harness.cpp"), os.path.join(csrc, "rulec.cpp"),
            os.path.join(csrc, "ruleset_image.cpp")]
    deps = srcs + [os.path.join(csrc, h) for h in ("pike_vm.h", "bitprog.h", "gram_filter.h", "rulec.h", "kernels.h", "ruleset_image.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-o", lib] + srcs)
    L = C.CDLL(lib)
    L.harness_create.restype = C.c_void_p
    L.harness_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    L.harness_destroy.argtypes = [C.c_void_p]
    L.harness_info.argtypes = [C.c_void_p, C.c_void_p]
    L.harness_rule_error.restype = C.c_char_p
    L.harness_rule_error.argtypes = [C.c_void_p, C.c_uint32]
    L.harness_rule_nfactors.argtypes = [C.c_void_p, C.c_uint32]
    L.harness_find_all.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.harness_test.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32]
    L.harness_candidates.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32]
    L.harness_policy_hits.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
    L.harness_policy_hits_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.harness_batch_rates.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
    L.harness_l1_factor_counts.restype = C.c_uint32
    L.harness_l1_factor_counts.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_void_p]
    L.harness_dump_factors.argtypes = [C.c_void_p]
    L.harness_bitprog_stats.argtypes = [C.c_void_p]
    L.harness_bitprog_eligible.argtypes = [C.c_void_p]
    L.harness_bitprog_eligible.restype = C.c_uint32
    return L
