"""TEST INFRASTRUCTURE: ctypes access to oracle/liboracle.so (the C restatement of the reference).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "raytracer-utah_b200", "python"))
import rtu_b200 as R  # struct definitions only (include/rtu.h mirrors)

LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


class OracleStats(C.Structure):
    _fields_ = [("trace_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("box_tests", C.c_uint64),
                ("tri_tests", C.c_uint64), ("node_visits", C.c_uint64), ("seconds", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle/liboracle.so not built: make -C oracle")
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_bvh_box.restype = C.c_float
        _lib.oracle_bvh_box.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        _lib.oracle_box_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
    return _lib


def box_intersect(rays, boxes, tmax):
    L = lib()
    rays = np.ascontiguousarray(rays, "f4"); boxes = np.ascontiguousarray(boxes, "f4")
    hit = np.zeros(len(rays), "i4"); tb = np.zeros(len(rays), "f4")
    for i in range(len(rays)):
        hit[i] = L.oracle_box_intersect(rays[i].ctypes.data, boxes[i].ctypes.data, float(tmax[i]))
        tb[i] = L.oracle_bvh_box(rays[i].ctypes.data, boxes[i].ctypes.data, float(tmax[i]))
    return hit, tb


def trace(desc, rays):
    L = lib()
    rays = np.ascontiguousarray(rays, R.RAY_DTYPE)
    hits = np.zeros(rays.shape[0], R.HIT_DTYPE)
    rc = L.oracle_trace(C.byref(desc), C.c_void_p(rays.ctypes.data), C.c_int64(rays.shape[0]), C.c_void_p(hits.ctypes.data))
    assert rc == 0
    return hits


def shadow_trace(desc, rays, t_max):
    L = lib()
    rays = np.ascontiguousarray(rays, R.RAY_DTYPE)
    t_max = np.ascontiguousarray(t_max, "f4")
    occ = np.zeros(rays.shape[0], "u1")
    rc = L.oracle_shadow_trace(C.byref(desc), C.c_void_p(rays.ctypes.data), C.c_void_p(t_max.ctypes.data), C.c_int64(rays.shape[0]), C.c_void_p(occ.ctypes.data))
    assert rc == 0
    return occ


def shade(desc, rays, hits, bounces=5):
    L = lib()
    rays = np.ascontiguousarray(rays, R.RAY_DTYPE)
    hits = np.ascontiguousarray(hits, R.HIT_DTYPE)
    rgb = np.zeros((rays.shape[0], 3), "f4")
    rc = L.oracle_shade(C.byref(desc), C.c_void_p(rays.ctypes.data), C.c_void_p(hits.ctypes.data), C.c_int64(rays.shape[0]), C.c_int(bounces), C.c_void_p(rgb.ctypes.data))
    assert rc == 0
    return rgb


def render(desc, width=0, height=0, spp=1, pattern=R.PATTERN_CENTER, mode=R.MODE_WHITTED, shade_bounces=5,
           crop=None, threads=None, want=("rgb", "rgb8"), params=None):
    """Render on the CPU with the C restatement; returns (buffers, stats dict incl. seconds)."""
    import time
    L = lib()
    p = params if params is not None else R.Params()
    if params is None:
        p.width, p.height, p.spp, p.pattern, p.mode, p.shade_bounces, p.gi_bounces = width, height, spp, pattern, mode, shade_bounces, 4
    w = p.width or desc.camera.width
    h = p.height or desc.camera.height
    spec = {"rgb8": ((h, w, 3), "u1", R.u8), "rgb": ((h, w, 3), "f4", R.f32), "z": ((h, w), "f4", R.f32),
            "z8": ((h, w), "u1", R.u8), "node_id": ((h, w), "i4", R.i32), "face_id": ((h, w), "i4", R.i32)}
    img = R.Image()
    bufs = {}
    for k in want:
        shp, dt, ct = spec[k]
        bufs[k] = np.zeros(shp, dt)
        setattr(img, k, bufs[k].ctypes.data_as(C.POINTER(ct)))
    st = OracleStats()
    cr = (C.c_int * 4)(*crop) if crop is not None else None
    if threads is None:
        threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    rc = L.oracle_render(C.byref(desc), C.byref(p), C.byref(img), cr, C.c_int(threads), C.byref(st))
    dt = time.perf_counter() - t0
    assert rc == 0
    stats = {k: getattr(st, k) for k, _ in OracleStats._fields_}
    stats["seconds"] = dt
    stats["threads"] = threads
    bufs["stats"] = stats
    return bufs


def sample_texcolor(desc, texcolor, uvw):
    L = lib()
    uvw = np.ascontiguousarray(uvw, "f4")
    out = np.zeros_like(uvw)
    L.oracle_sample_texcolor(C.byref(desc), C.byref(texcolor), C.c_void_p(uvw.ctypes.data), C.c_int64(len(uvw)), C.c_void_p(out.ctypes.data))
    return out


def sample_environment(desc, dirs):
    L = lib()
    dirs = np.ascontiguousarray(dirs, "f4")
    out = np.zeros_like(dirs)
    L.oracle_sample_environment(C.byref(desc), C.c_void_p(dirs.ctypes.data), C.c_int64(len(dirs)), C.c_void_p(out.ctypes.data))
    return out


PHOTON_DTYPE = np.dtype([("position", "<f4", 3), ("power", "<f4"), ("color", "u1", 3), ("plane_dirz", "u1"), ("dir_x", "<i2"), ("dir_y", "<i2")])
assert PHOTON_DTYPE.itemsize == 24


def balance_photons(photons):
    """cyPhotonMap::PrepareForIrradianceEstimation; returns n+1 records (index 0 unused)."""
    L = lib()
    photons = np.ascontiguousarray(photons, PHOTON_DTYPE)
    out = np.zeros(photons.shape[0] + 1, PHOTON_DTYPE)
    rc = L.oracle_balance_photons(C.c_void_p(photons.ctypes.data), C.c_uint32(photons.shape[0]), C.c_void_p(out.ctypes.data))
    assert rc == 0
    return out


def estimate_irradiance(balanced, pos, normal, radius, ellipticity):
    """cyPhotonMap::EstimateIrradiance<100> on a balanced map (n+1 records)."""
    L = lib()
    balanced = np.ascontiguousarray(balanced, PHOTON_DTYPE)
    pos = np.ascontiguousarray(pos, "f4")
    normal = None if normal is None else np.ascontiguousarray(normal, "f4")
    nq = pos.shape[0]
    irrad = np.zeros((nq, 3), "f4"); direction = np.zeros((nq, 3), "f4"); found = np.zeros(nq, "i4")
    rc = L.oracle_estimate_irradiance(C.c_void_p(balanced.ctypes.data), C.c_uint32(balanced.shape[0] - 1), C.c_void_p(pos.ctypes.data),
                                      C.c_void_p(None if normal is None else normal.ctypes.data), C.c_int64(nq), C.c_float(radius), C.c_float(ellipticity),
                                      C.c_void_p(irrad.ctypes.data), C.c_void_p(direction.ctypes.data), C.c_void_p(found.ctypes.data))
    assert rc == 0
    return irrad, direction, found


_photon_keepalive = None


def set_photon_map(balanced, radius=1.0, ellipticity=0.5):
    """The map RTU_MODE_PHOTON of oracle.render uses (balanced: n+1 records)."""
    global _photon_keepalive
    L = lib()
    _photon_keepalive = np.ascontiguousarray(balanced, PHOTON_DTYPE)
    L.oracle_set_photon_map(C.c_void_p(_photon_keepalive.ctypes.data), C.c_uint32(_photon_keepalive.shape[0] - 1), C.c_float(radius), C.c_float(ellipticity))
