"""The multi-rank device path over NCCL (the gloo rigs never execute the NCCL branches): torchrun with one rank
per GPU, force and list parity against the oracle's restatement of the reference flow.  Needs at least two GPUs; on the
one-GPU test box it is skipped (profiles/r2_nccl_parity_n*.json hold the runs made with gpurun --gpus 2 / 4)."""
import json
import os
import subprocess
import sys

import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_multirank_parity(world, tmp_path):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / "verdict.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "tools", "nccl_parity.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    v = json.loads(out.read_text())
    assert v["ok"], v
