"""Host-side pieces of the path against the LIVE reference (CPU; only where /root/reference is mounted -- the GPU box
skips it and relies on the committed fixtures)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/qiskit_dynamics"), reason="reference not mounted")
def test_signals_and_step_grids_against_the_live_reference():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_live_reference_worker.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "LIVE_REFERENCE_OK" in res.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference/qiskit_dynamics"), reason="reference not mounted")
def test_oracle_equals_the_live_reference_on_the_gpu_fuzz_cases():
    """Closes the chain for the 52 randomised cases of tests/test_fuzz_gpu.py: there GPU == oracle (on the B200), here
    oracle == the unmodified reference (same seeds, same case generator)."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_live_reference_fuzz_worker.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "LIVE_REFERENCE_FUZZ_OK cases=52" in res.stdout
