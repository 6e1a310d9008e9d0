"""csrc/xla_ffi_shim.cc -- the only path from JAX to the library -- cannot be built against the real
jaxlib / nanobind headers in this image (neither is installed).  This test compiles
it against a minimal mock of the `xla::ffi` typed-binder and nanobind surface it uses
(tests/mock_xla/include, test infrastructure) and RUNS its 18 handlers on the CPU: registration
names, attribute schema and order (ref lib/jax_finufft_gpu.cc:28-60, 356-422), float vs double
`eps`, operand-count check, error forwarding.  The real recipe stays INTEGRATION.md section 2."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "jax_finufft_b200")


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_shim_compiles_and_its_handlers_decode_the_reference_schema(tmp_path):
    assert os.path.exists(os.path.join(LIBDIR, "libb200nufft.so")), "build the library first (__graft_entry__.build())"
    exe = str(tmp_path / "shim_driver")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-DB2N_BUILD_XLA_SHIM", "-I", os.path.join(ROOT, "tests", "mock_xla", "include"),
           "-I", cuda_inc, os.path.join(ROOT, "tests", "mock_xla", "shim_driver.cc"),
           os.path.join(LIBDIR, "csrc", "xla_ffi_shim.cc"), "-L", LIBDIR, "-lb200nufft", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-2000:]
    lines = run.stdout.strip().splitlines()
    assert lines[-1] == "ALL OK" and sum(ln.startswith("ok nufft") for ln in lines) == 18, run.stdout
