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
