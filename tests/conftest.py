import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "raytracer-utah_b200", "python"))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES = os.path.join(ROOT, "scenes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files if k != "meta_json"}
    meta = json.loads(bytes(z["meta_json"]).decode())
    return d, meta


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        return np.array_equal(a.view("u4"), b.astype(a.dtype).view("u4"))
    return np.array_equal(a, b)


@pytest.fixture(scope="session")
def rtu():
    import rtu_b200
    rtu_b200.lib()  # fails loudly if librtu_b200.so is not built
    return rtu_b200


@pytest.fixture(scope="session")
def gpu_ctx(rtu):
    ctx = rtu.Context(0)
    yield ctx
    ctx.close()


def single_object_scene(rtu, kind, mesh_from=None):
    """root (identity) + one identity node holding a unit sphere / plane / the first mesh of a loaded scene."""
    R = rtu
    nodes = (R.Node * 2)()
    ident = (1, 0, 0, 0, 1, 0, 0, 0, 1)
    for i, n in enumerate(nodes):
        n.tm[:] = ident
        n.itm[:] = ident
        n.pos[:] = (0, 0, 0)
        n.parent = -1 if i == 0 else 0
        n.kind = R.OBJ_NONE if i == 0 else kind
        n.mesh = -1
        n.material = -1
    d = R.SceneDesc()
    d.nodes = C.cast(nodes, C.POINTER(R.Node))
    d.n_nodes = 2
    keep = [nodes]
    if kind == R.OBJ_MESH:
        nodes[1].mesh = 0
        d.meshes = mesh_from.desc.meshes
        d.n_meshes = 1
        keep.append(mesh_from)
    d.camera.width = 4
    d.camera.height = 4
    d.camera.fov = 40
    d.camera.focaldist = 1
    d.camera.dir[:] = (0, 0, -1)
    d.camera.up[:] = (0, 1, 0)
    d._keep = keep
    return d


def synthetic_scene(name, meta):
    """Regenerates a section-8d synthetic scene (tools/make_synthetic.py, deterministic) and checks that it is the
    file the reference produced the fixture from."""
    import hashlib
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synthetic
    rel = make_synthetic.ensure((name,))[name]
    for f, want in meta["sha256"].items():
        got = hashlib.sha256(open(os.path.join(SCENES, "synthetic", f), "rb").read()).hexdigest()
        assert got == want, "generated %s differs from the one the fixture was made from" % f
    return os.path.join(SCENES, rel)


STOCHASTIC_CASES = ["p9_dof", "teapot1_soft", "p10", "p11_glossy_soft", "p11_glossy"]


def check_stochastic(img, g, meta):
    """A Monte-Carlo image against the reference's 256-spp Whitted mean of the same scene (tests/golden/stochastic_*.npz,
    two rand() seeds).  The streams are unrelated, so the comparison is statistical, with the reference's own seed-to-seed
    difference as yardstick: image mean within 1 %, RMSE <= 1.5 x the seed-to-seed RMSE, and - so that a bias hidden in the
    per-pixel noise cannot pass - the same bar on 8x8 block means against the mean of the two reference runs."""
    a = np.asarray(img, np.float64)
    r0, r1 = g["rgb0"].astype(np.float64), g["rgb1"].astype(np.float64)
    if "crop" in meta and list(meta["crop"]) != [0, 0, meta["width"], meta["height"]] and a.shape != r0.shape:
        x0, y0, x1, y1 = meta["crop"]
        a = a[y0:y1, x0:x1]
    assert a.shape == r0.shape
    assert np.isfinite(a).all()
    assert abs(a.mean() - r0.mean()) <= 0.01 * r0.mean() + abs(r0.mean() - r1.mean()), (a.mean(), r0.mean(), r1.mean())
    noise = np.sqrt(np.mean((r0 - r1) ** 2))
    rmse = np.sqrt(np.mean((a - r0) ** 2))
    assert noise > 0
    assert rmse <= 1.5 * noise, "RMSE %.4g vs the reference's seed-to-seed RMSE %.4g" % (rmse, noise)

    def pool(x):
        h, w = (x.shape[0] // 8) * 8, (x.shape[1] // 8) * 8
        return x[:h, :w].reshape(h // 8, 8, w // 8, 8, 3).mean(axis=(1, 3))

    def trimmed_rms(x):
        # rare "fireflies" (one sample in millions: a back-face Fresnel factor of ~30, mtlFunctions.cpp:236-237) put a whole
        # run's squared error into one block on either side; the 2 % largest blocks are left out of both sums
        e = np.sort((x ** 2).sum(axis=2).ravel())
        return np.sqrt(e[:max(1, int(0.98 * e.size))].mean())

    pn = trimmed_rms(pool(r0 - r1))
    pe = trimmed_rms(pool(a - 0.5 * (r0 + r1)))
    assert pe <= 1.5 * pn + 1e-6, "block-mean RMSE %.4g vs the reference's %.4g" % (pe, pn)
    return rmse, noise, pe, pn
