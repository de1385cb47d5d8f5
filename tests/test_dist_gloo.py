"""N>1 host logic on CPU: world_size 2, gloo backend, rendezvous on 127.0.0.1."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world2_gloo_partition_and_rendezvous():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29471", os.path.join(ROOT, "tests", "dist_gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert "GLOO_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]
