"""One rank of the 2-GPU parity test (tests/test_multi_gpu.py): python multi_gpu_worker.py RANK WORLD IDFILE OUTDIR.

The ranks meet through the file system (rank 0 writes the 128-byte communicator id), so the test exercises the library's
own communicator and collectives and nothing else.  Each rank renders its share through the C ABI, the library brings the
image together on rank 0 (rtu_reduce_resolve for spp slices, rtu_gather_resolve for row ranges), and rank 0 compares it
with the frame it renders alone."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python"))
import rtu_b200 as R  # noqa: E402


def main():
    rank, world, idfile, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    if rank == 0:
        with open(idfile + ".tmp", "wb") as f:
            f.write(R.comm_unique_id())
        os.rename(idfile + ".tmp", idfile)
    t0 = time.time()
    while not os.path.exists(idfile):
        if time.time() - t0 > 120:
            raise SystemExit("rank %d: no communicator id after 120 s" % rank)
        time.sleep(0.02)
    uid = open(idfile, "rb").read()
    ctx = R.Context(rank)
    comm = R.Comm(ctx, uid, rank, world)
    failures = []
    for scene, mode, size, spp in (("Teapot/scene2.xml", R.MODE_WHITTED, (480, 270), 8), ("Project11/scene.xml", R.MODE_PATH, (200, 150), 16)):
        hs = R.HostScene(os.path.join(R.SCENES, scene))
        sc = R.Scene(ctx, hs.desc)
        kw = dict(width=size[0], height=size[1], spp=spp, pattern=R.PATTERN_REFERENCE, mode=mode, shade_bounces=5, gi_bounces=4, seed=9)
        per = spp // world
        # ---- spp slices, one ncclReduce of the RGB planes
        p = R.default_params(sample_begin=rank * per, sample_end=(rank + 1) * per, **kw)
        sc.render_device(p, 0, clear=True)
        red = sc.reduce_resolve(comm, p, 0, root=0, want=("rgb", "rgb8", "z"))
        # ---- row ranges, gathered
        rows = [size[1] * r // world // 4 * 4 for r in range(world)] + [size[1]]
        p2 = R.default_params(row_begin=rows[rank], row_end=rows[rank + 1], **kw)
        sc.render_device(p2, 0, clear=True)
        gat = sc.gather_resolve(comm, p2, 0, root=0, want=("rgb", "rgb8"))
        if rank == 0:
            full = sc.render(R.default_params(**kw), want=("rgb", "rgb8", "z"))
            for name, got in (("reduce", red), ("gather", gat)):
                a, b = got["rgb"].astype(np.float64), full["rgb"].astype(np.float64)
                rel = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-2)
                d8 = np.abs(got["rgb8"].astype(np.int32) - full["rgb8"].astype(np.int32)).max()
                print("%s %s: max rel %.3g, max RGB8 diff %d" % (scene, name, rel.max(), d8))
                if not (rel.max() <= 1e-4 and d8 <= 1):
                    failures.append("%s %s: max rel %.3g, RGB8 %d" % (scene, name, rel.max(), d8))
            if not np.array_equal(red["z"].view("u4"), full["z"].view("u4")):
                failures.append("%s: z of the reduced frame differs" % scene)
            if not float(np.abs(full["rgb"]).max()) > 0.01:
                failures.append("%s: empty frame" % scene)
        ctx.synchronize()
        sc.close()
        hs.close()
    comm.close()
    ctx.close()
    if rank == 0:
        with open(os.path.join(outdir, "result.txt"), "w") as f:
            f.write("\n".join(failures) if failures else "OK")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
