"""N>1 decomposition on CPU: world_size-2 gloo processes, each rendering its spp slice (or row range) with the
C oracle, one reduce(sum) to rank 0 -- the same plan bench.py runs on N GPUs with NCCL."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, SCENES

W, H, SPP = 96, 72, 4


def _worker(rank, world, port, mode, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    import rtu_b200 as R
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hs = R.HostScene(os.path.join(SCENES, "Project4.xml"))
    kw = dict(width=W, height=H, spp=SPP, pattern=R.PATTERN_REFERENCE, mode=R.MODE_WHITTED)
    if mode == "spp":
        per = SPP // world
        p = R.default_params(sample_begin=rank * per, sample_end=(rank + 1) * per, **kw)
    else:
        rows = [0, 40, H]
        p = R.default_params(row_begin=rows[rank], row_end=rows[rank + 1], **kw)
    o = oracle_py.render(hs.desc, params=p, want=("rgb",), threads=2)
    part = torch.from_numpy(o["rgb"].astype(np.float32) if mode == "spp" else o["rgb"])
    if mode == "rows":  # rows outside the range were never written: they are zero
        pass
    dist.reduce(part, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, part.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["spp", "rows"])
def test_two_rank_slices_reduce_to_the_full_frame(rtu, tmp_path, mode):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    out = str(tmp_path / ("reduced_%s.npy" % mode))
    port = 29600 + (os.getpid() % 300) + (0 if mode == "spp" else 1)
    mp.spawn(_worker, args=(2, port, mode, out), nprocs=2, join=True)
    got = np.load(out)
    hs = rtu.HostScene(os.path.join(SCENES, "Project4.xml"))
    full = oracle_py.render(hs.desc, width=W, height=H, spp=SPP, pattern=rtu.PATTERN_REFERENCE, mode=rtu.MODE_WHITTED, want=("rgb",))["rgb"]
    assert np.allclose(got, full, rtol=1e-5, atol=1e-7)
    assert float(np.abs(full).max()) > 0.01
