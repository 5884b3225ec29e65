"""The real multi-rank path on hardware: one process per GPU under torchrun, the in-library exchange step
(peer push over NVLink peer memory, or ncclSend/ncclRecv, or the all-gather) — the sharded result must equal the undivided
map bit for bit and the checksum must equal the single-GPU run's.  Skipped with fewer than two visible devices."""
import json
import os
import subprocess
import sys

import pytest

from common import ROOT

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _bench(n, workload, env=None, port=29631):
    base = [sys.executable]
    if n > 1:
        base += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                 "--master-port", str(port)]
    cmd = base + [os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--steps", "3", "--warmup", "3", "--workload", workload,
                  "--no-cpu-baseline"]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, **(env or {})), capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith('{"metric"')]
    assert r.returncode == 0 and lines, (r.stdout + r.stderr)[-3000:]
    return json.loads(lines[-1])


PEER, SENDRECV, ALLGATHER = {}, {"AMB_HALO_PEER": "0"}, {"AMB_HALO_EXCHANGE": "1"}


@pytest.mark.parametrize("workload,modes", [("joint_1k", (PEER, SENDRECV, ALLGATHER)), ("dsm_256_holes", (PEER,))])
def test_every_exchange_mode_equals_the_undivided_map(workload, modes):
    if _devices() < 2:
        pytest.skip("needs two visible GPUs")
    one = _bench(1, workload)
    port = 29631
    for env in modes:   # peer push, ncclSend/ncclRecv, all-gather
        port += 1
        two = _bench(2, workload, env, port)
        assert two["n_gpus"] == 2
        assert two.get("sharded_equals_undivided") is True, (env, two)
        assert two["checksum"] == one["checksum"], (env, two["checksum"], one["checksum"])
        expect = "peer push" if env is PEER else ("ncclSend/ncclRecv" if env is SENDRECV else "ncclAllGather")
        assert expect in two["run"]["sharding"], two["run"]["sharding"]
