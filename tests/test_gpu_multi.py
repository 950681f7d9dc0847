"""GPU, >= 2 devices: `torchrun -m`-style launch of the CLI must write the BED the single-process
reference wrote (DNMs shard by kid or, for a single family, by genomic slice; rank 0 gathers)."""
import os
import socket
import subprocess
import sys

import pytest

from tests.golden_util import CASES, load_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rows(text):
    from tests.test_gpu_golden import _bed_rows
    rows = _bed_rows(text)
    return rows[:1] + sorted(rows[1:])          # header, then records (gather order is by rank)


@pytest.mark.parametrize("name", CASES)
def test_cli_on_two_gpus_matches_reference(tmp_path, name):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "out.bed")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "cli_torchrun.py"), "--case", name, "--out", out]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    _, want = load_case(name)
    for kind in ("strict", "ambiguous"):
        assert _rows(open(out + "." + kind).read()) == _rows(want["cli"][kind]), (name, kind)
