"""On-hardware data-parallel equivalence: needs >= 2 GPUs on the box (`gpurun --gpus 2`), skipped
otherwise.  tests/dp_worker.py runs under torch.distributed.run with NCCL."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_data_parallel_step_matches_local_shards_and_replicas_agree():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    print(line)
    # the all-reduced SUM of the ranks' gradients (fine slice reduced early under the coarse backward, the rest after it)
    # is the sum of the same shards' gradients computed locally: two addends commute, so bit for bit
    assert line["grad_max_abs_diff_vs_local_sum"] == 0.0, line
    assert line["replica_max_abs_diff"] == 0.0 and line["params_moved"] > 0, line
