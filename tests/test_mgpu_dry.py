"""The multi-GPU scripts' logic on the CPU (MGPU_DRY=1: one gloo rank, the oracle behind the same API): names, API usage and every
expectation at world size 1.  The scripts proper run under torchrun on 2-4 GPUs (tests/test_multigpu.py, tests/test_zzz_mgpu_next.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script", ["mgpu_core.py", "mgpu_remove.py", "mgpu_distribute.py", "mgpu_agentstate.py", "mgpu_purge.py", "mgpu_sir.py", "mgpu_gol.py", "mgpu_moveto.py"])
def test_mgpu_script_dry_run(oracle, script):
    env = dict(os.environ, MGPU_DRY="1", MGPU_N="4000", MGPU_L="301")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank 0/1: ok" in r.stdout
