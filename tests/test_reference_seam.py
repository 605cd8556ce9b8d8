"""The reference's own orchestrator against our seam header: KMerCounter.cpp (dispatchWork,
Start: KMerCounter.cpp:51-89,108-191) must compile unchanged against host/GPUHandler.h -- the
three calls and every GPUStream field it touches.  The reference's sources are copied to a
temporary directory for the compile only (nothing of them enters the repository); TBB, an
un-vendored dependency of the reference, is stubbed (tests/stubs/tbb).  Skipped where
/root/reference does not exist (the GPU box)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("KC_REFERENCE_DIR", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "KMerCounter.cpp")), reason="reference sources not present")
def test_reference_kmercounter_compiles_against_our_gpuhandler_h(tmp_path):
    for name in os.listdir(REF):
        if name.endswith((".h", ".cpp")) and name != "GPUHandler.h":
            shutil.copy(os.path.join(REF, name), tmp_path / name)
    shutil.copy(os.path.join(ROOT, "kmer-counter_b200", "host", "GPUHandler.h"), tmp_path / "GPUHandler.h")
    cmd = ["g++", "-std=c++11", "-w", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", str(tmp_path),
           str(tmp_path / "KMerCounter.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # and it links against the shim: compile to an object and check the seam symbols it wants are the ones the shim exports
    obj = tmp_path / "KMerCounter.o"
    subprocess.run(["g++", "-std=c++11", "-w", "-c", "-fPIC", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", str(tmp_path),
                    str(tmp_path / "KMerCounter.cpp"), "-o", str(obj)], check=True)
    want = subprocess.run(["nm", "-C", "--undefined-only", str(obj)], capture_output=True, text=True, check=True).stdout
    shim = os.path.join(ROOT, "kmer-counter_b200", "host", "libkc_shim.so")
    have = subprocess.run(["nm", "-C", "-D", "--defined-only", shim], capture_output=True, text=True, check=True).stdout
    for sym in ("PrepareGPU(", "FreeGPU(", "processKMers("):
        w = [l.split(" U ")[-1].strip() for l in want.splitlines() if sym in l]
        assert w, sym
        for sig in w:
            assert sig in have, "the shim does not export %s" % sig
