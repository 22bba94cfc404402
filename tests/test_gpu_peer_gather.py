"""Row N2: the collection of the stitched flows on rank 0 through a CUDA-IPC-mapped buffer (sharding.PeerGatherBuffer,
tif_ipc_*).  Two processes; on a 1-GPU box they share the device (the IPC mapping works the same), on a multi-GPU box
each takes its own and the stores cross NVLink."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_items,mode", [(6, "fused"), (5, "fused"), (5, "copy")])
def test_peer_store_gather(n_items, mode):
    script = os.path.join(ROOT, "tests", "_peer_gather_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + n_items + (10 if mode == "copy" else 0)), script, str(n_items), mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PEER_GATHER_OK" in res.stdout
