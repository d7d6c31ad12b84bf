"""The product's multi-GPU path on hardware: 2 ranks, one GPU each, through the C ABI's own communicator
(rtu_comm_*) and collectives (rtu_reduce_resolve, rtu_gather_resolve).  SURVEY 8d: "the spp-sliced image for G GPUs must
equal the 1-GPU image up to FP32 summation order".  Skips on a single-GPU box; bench.py repeats the check at every N>1
("multi_gpu_parity" in its JSON line)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_image_equals_one_gpu_image(rtu, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    worker = os.path.join(ROOT, "tests", "multi_gpu_worker.py")
    idfile = str(tmp_path / "comm_id")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", idfile, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert open(str(tmp_path / "result.txt")).read() == "OK", "\n".join(outs)
