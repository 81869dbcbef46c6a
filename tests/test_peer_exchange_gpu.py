"""2-GPU check of the fused output exchange (in-kernel peer stores + flag wait).  Skipped on single-GPU boxes; the
same script was run at 2 and 8 GPUs during the round (`tools/peer_check.py`)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs of one NVLink domain")
@pytest.mark.parametrize("fold_wait", ["0", "1"])
def test_fused_exchange_two_gpus(fold_wait):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, FOLD_WAIT=fold_wait)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "peer_check.py")],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("fused exchange OK") == 2
    assert r.stdout.count("step exchange OK") == 2
