"""ctypes binding of include/rtu.h (librtu_b200.so) for the tests and bench.py.

Not the product: the product is the C ABI.  This module only mirrors the structs and wraps
the calls with numpy buffers so that the Python parity tests read like calls into the
reference (Trace / ShadowTrace / Shade / Render).  It never computes anything itself and
fails loudly if the shared library is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(os.path.dirname(_HERE))          # raytracer-utah_b200/
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_PATH = os.environ.get("RTU_B200_LIB") or os.path.join(PKG_ROOT, "librtu_b200.so")  # override: tuning variants only
SCENES = os.path.join(REPO_ROOT, "scenes")
BIGFLOAT = np.float32(1.0e30)

OBJ_NONE, OBJ_SPHERE, OBJ_PLANE, OBJ_MESH = 0, 1, 2, 3
TEX_NULL, TEX_CHECKER, TEX_FILE = 0, 1, 2
LIGHT_AMBIENT, LIGHT_DIRECT, LIGHT_POINT = 0, 1, 2
MODE_PRIMARY, MODE_WHITTED, MODE_PATH, MODE_PHOTON, MODE_PHOTON_GATHER = 0, 1, 2, 3, 4
PATTERN_CENTER, PATTERN_REFERENCE = 0, 1
FLAG_CULL_NULL_SHADOW_RAYS = 1
FLAG_CULL_ZERO_WEIGHT_RAYS = 2
FLAG_TIME_KERNELS = 4
FLAG_REFERENCE_WALK = 8
MESH_DEVICE_BVH = 1
LOAD_DEVICE_BVH = 1

f32, i32, u32, u8 = C.c_float, C.c_int32, C.c_uint32, C.c_uint8
PF, PU, PB = C.POINTER(f32), C.POINTER(u32), C.POINTER(u8)


class Node(C.Structure):
    _fields_ = [("tm", f32 * 9), ("itm", f32 * 9), ("pos", f32 * 3), ("parent", i32),
                ("kind", i32), ("mesh", i32), ("material", i32)]


class Mesh(C.Structure):
    _fields_ = [("v", PF), ("nv", u32), ("vn", PF), ("nvn", u32), ("vt", PF), ("nvt", u32),
                ("f", PU), ("fn", PU), ("ft", PU), ("nf", u32),
                ("bvh_boxes", PF), ("bvh_data", PU), ("bvh_nodes", u32), ("bvh_elements", PU),
                ("bound_min", f32 * 3), ("bound_max", f32 * 3),
                ("occ_nodes", PF), ("occ_slots", PU), ("occ_n_nodes", u32), ("occ_root", u32), ("flags", u32)]


class TexMap(C.Structure):
    _fields_ = [("kind", i32), ("itm", f32 * 9), ("pos", f32 * 3), ("color1", f32 * 3),
                ("color2", f32 * 3), ("rgb8", PB), ("width", i32), ("height", i32)]


class TexColor(C.Structure):
    _fields_ = [("color", f32 * 3), ("texmap", i32)]


class Material(C.Structure):
    _fields_ = [("diffuse", TexColor), ("specular", TexColor), ("reflection", TexColor),
                ("refraction", TexColor), ("emission", TexColor), ("glossiness", f32),
                ("absorption", f32 * 3), ("ior", f32), ("reflection_glossiness", f32),
                ("refraction_glossiness", f32)]


class Light(C.Structure):
    _fields_ = [("kind", i32), ("intensity", f32 * 3), ("v", f32 * 3), ("size", f32)]


class Camera(C.Structure):
    _fields_ = [("pos", f32 * 3), ("dir", f32 * 3), ("up", f32 * 3), ("fov", f32),
                ("focaldist", f32), ("dof", f32), ("width", i32), ("height", i32)]


class LightMask(C.Structure):
    _fields_ = [("node", i32), ("light", i32), ("rec", f32 * 24), ("bits", C.POINTER(u32)),
                ("cell_start", C.POINTER(u32)), ("items", C.POINTER(u32)), ("n_items", u32)]


class SceneDesc(C.Structure):
    _fields_ = [("camera", Camera),
                ("nodes", C.POINTER(Node)), ("n_nodes", i32),
                ("meshes", C.POINTER(Mesh)), ("n_meshes", i32),
                ("materials", C.POINTER(Material)), ("n_materials", i32),
                ("lights", C.POINTER(Light)), ("n_lights", i32),
                ("texmaps", C.POINTER(TexMap)), ("n_texmaps", i32),
                ("background", TexColor), ("environment", TexColor),
                ("light_masks", C.POINTER(LightMask)), ("n_light_masks", i32)]


class Params(C.Structure):
    _fields_ = [("width", i32), ("height", i32), ("spp", i32), ("sample_begin", i32),
                ("sample_end", i32), ("pattern", i32), ("mode", i32), ("shade_bounces", i32),
                ("gi_bounces", i32), ("row_begin", i32), ("row_end", i32), ("flags", u32),
                ("seed", C.c_uint64), ("adaptive_min_spp", i32), ("adaptive_step", i32), ("adaptive_target", f32), ("reserved_", i32)]


class Image(C.Structure):
    _fields_ = [("rgb8", PB), ("rgb", PF), ("z", PF), ("z8", PB),
                ("node_id", C.POINTER(i32)), ("face_id", C.POINTER(i32)), ("sample_count", PB)]


class KernelStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("box_tests", C.c_uint64), ("tri_tests", C.c_uint64),
                ("node_visits", C.c_uint64), ("launches", C.c_uint64), ("ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [("trace_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("box_tests", C.c_uint64),
                ("tri_tests", C.c_uint64), ("node_visits", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("device_ms", C.c_double), ("primary_wave", KernelStats), ("secondary_waves", KernelStats),
                ("shadow_waves", KernelStats), ("shade_kernels", KernelStats), ("scene_device_bytes", C.c_uint64),
                ("pixel_samples", C.c_uint64), ("bvh_build_ms", C.c_double), ("queue_retries", C.c_uint64)]

    def as_dict(self):
        out = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            out[k] = v.as_dict() if isinstance(v, KernelStats) else v
        return out


RAY_DTYPE = np.dtype([("p", "<f4", 3), ("dir", "<f4", 3)])
HIT_DTYPE = np.dtype([("z", "<f4"), ("p", "<f4", 3), ("N", "<f4", 3), ("uvw", "<f4", 3),
                      ("node", "<i4"), ("face", "<i4"), ("front", "<i4")])

_lib = None


# cyPhotonMap::Photon, 24 bytes (include/rtu.h rtu_photon)
PHOTON_DTYPE = np.dtype([("position", "<f4", 3), ("power", "<f4"), ("color", "u1", 3), ("plane_dirz", "u1"), ("dir_x", "<i2"), ("dir_y", "<i2")])


class PhotonParams(C.Structure):
    _fields_ = [("map_size", u32), ("max_bounce", u32), ("est_radius", f32), ("ellipticity", f32), ("seed", C.c_uint64)]


class PhotonStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("from_light", C.c_uint64), ("stored", C.c_uint64), ("trace_rays", C.c_uint64),
                ("scale_factor", f32), ("emit_ms", f32), ("build_ms", f32), ("device_build", u32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RtuError(RuntimeError):
    pass


def lib():
    """Loads librtu_b200.so; there is no fallback of any kind."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RtuError("librtu_b200.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.rtu_last_error.restype = C.c_char_p
        L.rtu_host_scene_desc.restype = C.POINTER(SceneDesc)
        L.rtu_host_scene_desc.argtypes = [C.c_void_p]
        L.rtu_host_load_xml.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.rtu_host_scene_destroy.argtypes = [C.c_void_p]
        L.rtu_host_scene_destroy.restype = None
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RtuError("%s failed (status %d): %s" % (what, rc, lib().rtu_last_error().decode("utf-8", "replace")))


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


class HostScene:
    """rtu_host_load_xml: our replacement for LoadScene(const char*) (xmlload.cpp:64)."""

    def __init__(self, xml_path, asset_root=None, flags=0):
        L = lib()
        self._h = C.c_void_p()
        root = asset_root if asset_root is not None else SCENES
        L.rtu_host_load_xml_ex.argtypes = [C.c_char_p, C.c_char_p, u32, C.POINTER(C.c_void_p)]
        _check(L.rtu_host_load_xml_ex(os.fsencode(xml_path), os.fsencode(root), flags, C.byref(self._h)), "rtu_host_load_xml_ex")
        self.warnings = L.rtu_last_error().decode("utf-8", "replace")
        self.desc = L.rtu_host_scene_desc(self._h).contents

    def close(self):
        if self._h:
            lib().rtu_host_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # numpy views for the loader-parity tests
    def nodes(self):
        d = self.desc
        n = d.n_nodes
        out = {"tm": np.zeros((n, 9), "f4"), "itm": np.zeros((n, 9), "f4"), "pos": np.zeros((n, 3), "f4"),
               "meta": np.zeros((n, 4), "i4")}
        for i in range(n):
            nd = d.nodes[i]
            out["tm"][i] = np.frombuffer(nd.tm, "f4")
            out["itm"][i] = np.frombuffer(nd.itm, "f4")
            out["pos"][i] = np.frombuffer(nd.pos, "f4")
            out["meta"][i] = (nd.parent, nd.kind, nd.mesh, nd.material)
        return out

    def mesh(self, i):
        m = self.desc.meshes[i]

        def arr(p, n, dt):
            if not p or n == 0:
                return np.zeros((0, 3), dt)
            return np.ctypeslib.as_array(p, shape=(n * 3,)).view(dt).reshape(n, 3).copy()

        return {"v": arr(m.v, m.nv, "f4"), "vn": arr(m.vn, m.nvn, "f4"), "vt": arr(m.vt, m.nvt, "f4"),
                "f": arr(m.f, m.nf, "u4"), "fn": arr(m.fn, m.nf, "u4"), "ft": arr(m.ft, m.nf, "u4"),
                "bvh_boxes": np.ctypeslib.as_array(m.bvh_boxes, shape=(m.bvh_nodes * 6,)).reshape(-1, 6).copy() if m.bvh_boxes else np.zeros((0, 6), "f4"),
                "bvh_data": np.ctypeslib.as_array(m.bvh_data, shape=(m.bvh_nodes,)).copy() if m.bvh_data else np.zeros(0, "u4"),
                "bvh_elements": np.ctypeslib.as_array(m.bvh_elements, shape=(m.nf,)).copy() if m.bvh_elements else np.zeros(0, "u4"),
                "bound": np.concatenate([np.frombuffer(m.bound_min, "f4"), np.frombuffer(m.bound_max, "f4")])}

    def camera(self):
        c = self.desc.camera
        return np.array(list(c.pos) + list(c.dir) + list(c.up) + [c.fov, c.focaldist, c.dof, c.width, c.height], "f4")

    def materials(self):
        d = self.desc
        out = np.zeros((d.n_materials, 22), "f4")
        for i in range(d.n_materials):
            m = d.materials[i]
            out[i] = (list(m.diffuse.color) + list(m.specular.color) + list(m.reflection.color) +
                      list(m.refraction.color) + list(m.emission.color) + list(m.absorption) +
                      [m.glossiness, m.ior, m.reflection_glossiness, m.refraction_glossiness])
        return out

    def lights(self):
        d = self.desc
        out = np.zeros((d.n_lights, 8), "f4")
        for i in range(d.n_lights):
            l = d.lights[i]
            out[i] = [l.kind] + list(l.intensity) + list(l.v) + [l.size]
        return out


def default_params(**kw):
    p = Params()
    lib().rtu_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Context:
    def __init__(self, device=0, stream=None):
        L = lib()
        self._h = C.c_void_p()
        L.rtu_context_create.argtypes = [i32, C.c_void_p, C.POINTER(C.c_void_p)]
        _check(L.rtu_context_create(device, C.c_void_p(stream or 0), C.byref(self._h)), "rtu_context_create")

    def selftest_division(self, numerators_per_divisor=64, seed=1):
        """rtu_selftest_division: (pairs tested, bitwise mismatches) of the hoisted-reciprocal quotient vs a / b."""
        L = lib()
        t, m = C.c_uint64(0), C.c_uint64(0)
        L.rtu_selftest_division.argtypes = [C.c_void_p, u32, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _check(L.rtu_selftest_division(self._h, numerators_per_divisor, seed, C.byref(t), C.byref(m)), "rtu_selftest_division")
        return t.value, m.value

    def synchronize(self):
        L = lib()
        L.rtu_synchronize.argtypes = [C.c_void_p]
        _check(L.rtu_synchronize(self._h), "rtu_synchronize")

    def close(self):
        if self._h:
            L = lib()
            L.rtu_context_destroy.argtypes = [C.c_void_p]
            L.rtu_context_destroy.restype = None
            L.rtu_context_destroy(self._h)
            self._h = C.c_void_p()


PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64)
ERR_UNSUPPORTED = 5
ERR_CANCELLED = 6


class Job:
    """rtu_job: a frame (rtu_render_async) or a PNG file (rtu_write_png_async) in flight on a worker thread."""

    def __init__(self, buffers=None):
        self._h = C.c_void_p()
        self.buffers = buffers

    def progress(self):
        """(pixels_done, pixels_total, finished) - numRenderedPixels / IsRenderDone of RenderImage (scene.h:585-588)."""
        L = lib()
        d, t, f = C.c_int64(0), C.c_int64(0), i32(0)
        L.rtu_job_progress.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(i32)]
        _check(L.rtu_job_progress(self._h, C.byref(d), C.byref(t), C.byref(f)), "rtu_job_progress")
        return d.value, t.value, bool(f.value)

    def cancel(self):
        L = lib()
        L.rtu_job_cancel.argtypes = [C.c_void_p]
        L.rtu_job_cancel.restype = None
        L.rtu_job_cancel(self._h)

    def wait(self):
        """Joins the worker; raises RtuError if the job failed or was cancelled.  Returns the frame's buffers."""
        L = lib()
        L.rtu_job_wait.argtypes = [C.c_void_p]
        rc = L.rtu_job_wait(self._h)
        self.status = rc
        _check(rc, "rtu_job_wait")
        return self.buffers

    def close(self):
        if self._h:
            L = lib()
            L.rtu_job_destroy.argtypes = [C.c_void_p]
            L.rtu_job_destroy.restype = None
            L.rtu_job_destroy(self._h)
            self._h = C.c_void_p()


def write_png_async(path, pixels):
    """rtu_write_png_async: the encode runs on a worker thread; the pixels are copied before this returns."""
    L = lib()
    px = np.ascontiguousarray(pixels, "u1")
    ch = 1 if px.ndim == 2 else px.shape[2]
    job = Job()
    L.rtu_write_png_async.argtypes = [C.c_char_p, C.c_void_p, i32, i32, i32, C.POINTER(C.c_void_p)]
    _check(L.rtu_write_png_async(os.fsencode(path), px.ctypes.data, px.shape[1], px.shape[0], ch, C.byref(job._h)), "rtu_write_png_async")
    return job


COMM_ID_BYTES = 128


def comm_unique_id():
    """rtu_comm_unique_id: 128 bytes rank 0 hands to the other ranks (any transport) before rtu_comm_create."""
    L = lib()
    buf = (u8 * COMM_ID_BYTES)()
    _check(L.rtu_comm_unique_id(buf), "rtu_comm_unique_id")
    return bytes(buf)


class Comm:
    """rtu_comm: this rank's end of the NCCL communicator the multi-GPU entry points use (one context per rank)."""

    def __init__(self, ctx, unique_id, rank, world):
        L = lib()
        self.rank, self.world = rank, world
        self._h = C.c_void_p()
        buf = (u8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        L.rtu_comm_create.argtypes = [C.c_void_p, C.c_void_p, i32, i32, C.POINTER(C.c_void_p)]
        _check(L.rtu_comm_create(ctx._h, buf, rank, world, C.byref(self._h)), "rtu_comm_create")

    def close(self):
        if self._h:
            L = lib()
            L.rtu_comm_destroy.argtypes = [C.c_void_p]
            L.rtu_comm_destroy.restype = None
            L.rtu_comm_destroy(self._h)
            self._h = C.c_void_p()


class Scene:
    """Device-resident scene (rtu_scene_upload) with the batched operator calls."""

    def __init__(self, ctx, desc):
        L = lib()
        self.ctx = ctx
        self.desc = desc
        self._h = C.c_void_p()
        L.rtu_scene_upload.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]
        _check(L.rtu_scene_upload(ctx._h, C.byref(desc), C.byref(self._h)), "rtu_scene_upload")

    def close(self):
        if self._h:
            L = lib()
            L.rtu_scene_destroy.argtypes = [C.c_void_p]
            L.rtu_scene_destroy.restype = None
            L.rtu_scene_destroy(self._h)
            self._h = C.c_void_p()

    def trace(self, rays):
        """Trace(ray, &rootNode, HitInfo()) for every ray (RenderFunctions.cpp:181)."""
        L = lib()
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        L.rtu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _check(L.rtu_trace(self._h, rays.ctypes.data, rays.shape[0], hits.ctypes.data), "rtu_trace")
        return hits

    def shadow_trace(self, rays, t_max):
        """GenLight::Shadow's ShadowTrace with h.z = t_max (lightFunctions.cpp:27-37)."""
        L = lib()
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        t_max = np.ascontiguousarray(t_max, "f4")
        occ = np.zeros(rays.shape[0], "u1")
        L.rtu_shadow_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _check(L.rtu_shadow_trace(self._h, rays.ctypes.data, t_max.ctypes.data, rays.shape[0], occ.ctypes.data), "rtu_shadow_trace")
        return occ

    def shade(self, rays, hits, bounces=5):
        """hit.node->GetMaterial()->Shade(ray, hit, lights, bounces) (mtlFunctions.cpp:120)."""
        L = lib()
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.ascontiguousarray(hits, HIT_DTYPE)
        rgb = np.zeros((rays.shape[0], 3), "f4")
        L.rtu_shade.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, i32, C.c_void_p]
        _check(L.rtu_shade(self._h, rays.ctypes.data, hits.ctypes.data, rays.shape[0], bounces, rgb.ctypes.data), "rtu_shade")
        return rgb

    def camera_rays(self, params, sample=0):
        """The camera ray Render() builds for one sample of every pixel (RenderFunctions.cpp:78-97)."""
        L = lib()
        w, h = self._dims(params)
        rays = np.zeros(w * h, RAY_DTYPE)
        L.rtu_camera_rays.argtypes = [C.c_void_p, C.POINTER(Params), i32, C.c_void_p]
        _check(L.rtu_camera_rays(self._h, C.byref(params), sample, rays.ctypes.data), "rtu_camera_rays")
        return rays

    def _dims(self, params):
        w = params.width or self.desc.camera.width
        h = params.height or self.desc.camera.height
        return max(w, 1), max(h, 1)  # invalid sizes are rejected by the library, not here

    def render(self, params, want=("rgb8", "rgb", "z", "z8", "node_id", "face_id"), out=None):
        """rtu_render: whole frame, host buffers out (the e2e path).  out: optional dict of caller-owned arrays (e.g.
        page-locked ones) to write into instead of fresh numpy arrays."""
        L = lib()
        bufs, img = self._image(params, want, out)
        L.rtu_render.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Image)]
        _check(L.rtu_render(self._h, C.byref(params), C.byref(img)), "rtu_render")
        return bufs

    def render_device(self, params, d_accum=0, clear=True):
        L = lib()
        L.rtu_render_device.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, i32]
        _check(L.rtu_render_device(self._h, C.byref(params), C.c_void_p(d_accum), 1 if clear else 0), "rtu_render_device")

    def _image(self, params, want, out=None):
        w, h = self._dims(params)
        spec = {"rgb8": ((h, w, 3), "u1", u8), "rgb": ((h, w, 3), "f4", f32), "z": ((h, w), "f4", f32),
                "z8": ((h, w), "u1", u8), "node_id": ((h, w), "i4", i32), "face_id": ((h, w), "i4", i32),
                "sample_count": ((h, w), "u1", u8)}
        bufs, img = {}, Image()
        for k in want:
            shp, dt, ct = spec[k]
            if out is not None and k in out:
                a = out[k]
                if a.shape != shp or a.dtype != np.dtype(dt) or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError("out[%r] must be a C-contiguous %s array of shape %s" % (k, dt, shp))
                bufs[k] = a
            else:
                bufs[k] = np.zeros(shp, dt)
            setattr(img, k, _ptr(bufs[k], ct))
        return bufs, img

    def reduce_resolve(self, comm, params, d_accum=0, root=0, want=("rgb8",), out=None):
        """rtu_reduce_resolve: spp slices -> one ncclReduce of the RGB sums onto `root` -> resolve -> host buffers.
        Collective; returns the buffers on the root, None elsewhere."""
        L = lib()
        L.rtu_reduce_resolve.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, i32, C.POINTER(Image)]
        if comm.rank == root:
            bufs, img = self._image(params, want, out)
            _check(L.rtu_reduce_resolve(self._h, comm._h, C.byref(params), C.c_void_p(d_accum), root, C.byref(img)), "rtu_reduce_resolve")
            return bufs
        _check(L.rtu_reduce_resolve(self._h, comm._h, C.byref(params), C.c_void_p(d_accum), root, None), "rtu_reduce_resolve")
        return None

    def gather_resolve(self, comm, params, d_accum=0, root=0, want=("rgb8",), out=None):
        """rtu_gather_resolve: row ranges -> every rank resolves its rows and sends them to `root`."""
        L = lib()
        L.rtu_gather_resolve.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, i32, C.POINTER(Image)]
        if comm.rank == root:
            bufs, img = self._image(params, want, out)
            _check(L.rtu_gather_resolve(self._h, comm._h, C.byref(params), C.c_void_p(d_accum), root, C.byref(img)), "rtu_gather_resolve")
            return bufs
        _check(L.rtu_gather_resolve(self._h, comm._h, C.byref(params), C.c_void_p(d_accum), root, None), "rtu_gather_resolve")
        return None

    def render_async(self, params, want=("rgb8",), out=None, progress=None):
        """rtu_render_async (BeginRender): returns a Job at once; `progress(pixels_done, pixels_total)` runs on the worker
        thread after every slice, when the buffers hold the mean over what is done so far."""
        L = lib()
        bufs, img = self._image(params, want, out)
        job = Job(bufs)
        job._keep = (img, params)
        cb = PROGRESS_FN(lambda user, done, total: progress(done, total)) if progress is not None else C.cast(None, PROGRESS_FN)
        job._cb = cb
        L.rtu_render_async.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Image), PROGRESS_FN, C.c_void_p, C.POINTER(C.c_void_p)]
        _check(L.rtu_render_async(self._h, C.byref(params), C.byref(img), cb, None, C.byref(job._h)), "rtu_render_async")
        return job

    def resolve(self, params, d_accum=0, want=("rgb8", "rgb")):
        L = lib()
        bufs, img = self._image(params, want)
        L.rtu_resolve.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.POINTER(Image)]
        _check(L.rtu_resolve(self._h, C.byref(params), C.c_void_p(d_accum), C.byref(img)), "rtu_resolve")
        return bufs

    # ---- photon map
    def photon_map_generate(self, **kw):
        """GeneratePhotonMap() (RenderFunctions.cpp:341-392) on the device; returns the emission statistics."""
        L = lib()
        pp = photon_params(**kw)
        st = PhotonStats()
        L.rtu_photon_map_generate.argtypes = [C.c_void_p, C.POINTER(PhotonParams), C.POINTER(PhotonStats)]
        _check(L.rtu_photon_map_generate(self._h, C.byref(pp), C.byref(st)), "rtu_photon_map_generate")
        return st.as_dict()

    def photon_map_set(self, photons, **kw):
        """Installs caller photons (any order): PrepareForIrradianceEstimation + upload."""
        L = lib()
        photons = np.ascontiguousarray(photons, PHOTON_DTYPE)
        pp = photon_params(**kw)
        L.rtu_photon_map_set.argtypes = [C.c_void_p, C.c_void_p, u32, C.POINTER(PhotonParams)]
        _check(L.rtu_photon_map_set(self._h, photons.ctypes.data, photons.shape[0], C.byref(pp)), "rtu_photon_map_set")

    def photon_map_info(self):
        """(number of photons, whether the kd-tree was balanced on the device)."""
        L = lib()
        n, dev = u32(0), u32(0)
        L.rtu_photon_map_info.argtypes = [C.c_void_p, C.POINTER(u32), C.POINTER(u32)]
        _check(L.rtu_photon_map_info(self._h, C.byref(n), C.byref(dev)), "rtu_photon_map_info")
        return n.value, bool(dev.value)

    def photon_map_get(self):
        """The balanced map as cyPhotonMap stores it (photons[1..n])."""
        L = lib()
        n = u32(0)
        L.rtu_photon_map_get.argtypes = [C.c_void_p, C.c_void_p, u32, C.POINTER(u32)]
        _check(L.rtu_photon_map_get(self._h, None, 0, C.byref(n)), "rtu_photon_map_get")
        out = np.zeros(n.value, PHOTON_DTYPE)
        if n.value:
            _check(L.rtu_photon_map_get(self._h, out.ctypes.data, n.value, C.byref(n)), "rtu_photon_map_get")
        return out

    def estimate_irradiance(self, pos, normal, radius=1.0, ellipticity=0.5):
        """cyPhotonMap::EstimateIrradiance<100> for a batch of points (cyPhotonMap.h:276-323)."""
        L = lib()
        pos = np.ascontiguousarray(pos, "f4")
        nq = pos.shape[0]
        nrm = None if normal is None else np.ascontiguousarray(normal, "f4")
        irrad = np.zeros((nq, 3), "f4"); direction = np.zeros((nq, 3), "f4"); found = np.zeros(nq, "i4")
        L.rtu_estimate_irradiance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, f32, f32, C.c_void_p, C.c_void_p, C.c_void_p]
        _check(L.rtu_estimate_irradiance(self._h, pos.ctypes.data, None if nrm is None else nrm.ctypes.data, nq, radius, ellipticity,
                                         irrad.ctypes.data, direction.ctypes.data, found.ctypes.data), "rtu_estimate_irradiance")
        return irrad, direction, found

    def stats(self):
        L = lib()
        s = Stats()
        L.rtu_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        _check(L.rtu_get_stats(self._h, C.byref(s)), "rtu_get_stats")
        return s.as_dict()


def photon_params(**kw):
    L = lib()
    pp = PhotonParams()
    L.rtu_photon_params_default.argtypes = [C.POINTER(PhotonParams)]
    L.rtu_photon_params_default.restype = None
    L.rtu_photon_params_default(C.byref(pp))
    for k, v in kw.items():
        setattr(pp, k, v)
    return pp


def balance_photons(photons):
    """rtu_host_balance_photons: cyPhotonMap::PrepareForIrradianceEstimation on the host; n+1 records, [0] unused."""
    L = lib()
    photons = np.ascontiguousarray(photons, PHOTON_DTYPE)
    out = np.zeros(photons.shape[0] + 1, PHOTON_DTYPE)
    L.rtu_host_balance_photons.argtypes = [C.c_void_p, u32, C.c_void_p]
    _check(L.rtu_host_balance_photons(photons.ctypes.data, photons.shape[0], out.ctypes.data), "rtu_host_balance_photons")
    return out


def build_light_mask(desc, node, light):
    """rtu_host_build_light_mask: (rec[20] as f4, bits[256, 256] as bool) or None when the pair gets no mask."""
    L = lib()
    rec = np.zeros(24, "f4")
    bits = np.zeros(2048, "u4")
    L.rtu_host_build_light_mask.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    rc = L.rtu_host_build_light_mask(C.byref(desc), node, light, rec.ctypes.data, bits.ctypes.data)
    if rc == ERR_UNSUPPORTED:
        return None
    _check(rc, "rtu_host_build_light_mask")
    cells = np.unpackbits(bits.view("u1"), bitorder="little").reshape(256, 256).astype(bool)
    return rec, cells


def build_bvh(v, f, max_per_leaf=4):
    """rtu_host_build_bvh: cyBVH::Build on caller arrays."""
    L = lib()
    v = np.ascontiguousarray(v, "f4")
    f = np.ascontiguousarray(f, "u4")
    nf = f.shape[0]
    boxes = np.zeros((max(2 * nf, 2), 6), "f4")
    data = np.zeros(max(2 * nf, 2), "u4")
    elem = np.zeros(max(nf, 1), "u4")
    n = u32(0)
    L.rtu_host_build_bvh.argtypes = [C.c_void_p, u32, C.c_void_p, u32, u32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(u32)]
    _check(L.rtu_host_build_bvh(v.ctypes.data, v.shape[0], f.ctypes.data, nf, max_per_leaf,
                                boxes.ctypes.data, data.ctypes.data, elem.ctypes.data, C.byref(n)), "rtu_host_build_bvh")
    return boxes[:n.value].copy(), data[:n.value].copy(), elem[:nf].copy()


def write_png(path, pixels):
    L = lib()
    px = np.ascontiguousarray(pixels, "u1")
    ch = 1 if px.ndim == 2 else px.shape[2]
    L.rtu_write_png.argtypes = [C.c_char_p, C.c_void_p, i32, i32, i32]
    _check(L.rtu_write_png(os.fsencode(path), px.ctypes.data, px.shape[1], px.shape[0], ch), "rtu_write_png")
