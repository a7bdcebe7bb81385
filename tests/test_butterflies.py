"""Host check of the in-register butterflies (csrc/fft_core.cuh) against the DFT definition -- no GPU needed.

The butterflies are __host__ __device__; csrc/test_butterflies.cu runs every radix the pass kernels use
(2, 3, 4, 6, 8, 12, 16 and the twiddled radix 8) in both directions and both precisions."""
import os
import shutil
import subprocess

import pytest

from tests.conftest import ROOT

CSRC = os.path.join(ROOT, "circulantpreconditioner_b200", "csrc")


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc missing")
def test_butterflies_match_dft_definition():
    subprocess.check_call(["make", "-s", "-C", CSRC, "test_butterflies"])
    out = subprocess.run([os.path.join(CSRC, "test_butterflies")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL PASSED" in out.stdout
