#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_harness).

Runs only in the build container (it needs /root/reference to have been compiled by
`make -C oracle/ref`).  The fixtures are what pins parity: every number in them was computed by
the reference's own Trace / IntersectRay / Shade / LoadScene code.  The GPU box
never regenerates them.

  kat_<prim>.npz        seeded random rays through Sphere/Plane/Box/BVHBox/TriObj::IntersectRay
  loader_<scene>.npz    what LoadScene leaves in the globals (transforms, camera, BVH, ...)
  primary_<scene>.npz   pixel-centre Trace(): z, node, face, front, p, N, uvw
  whitted_<scene>.npz   Trace + Shade(ray,h,lights,5): linear RGB, RGB8, ray counts
  tex_<scene>.npz       TexturedColor::Sample / SampleEnvironment on seeded inputs
  stochastic_<scene>.npz  256-spp Whitted means of the scenes with depth of field / soft lights / glossy lobes, two seeds
"""
import glob
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
SCENES = os.path.join(ROOT, "scenes")
OUT = os.path.join(ROOT, "tests", "golden")

# scene -> (tag, low-res size for the committed fixtures)
PRIMARY = {
    "Project1Example.xml": ("p1example", (240, 180)),
    "Project1Test.xml": ("p1test", (200, 150)),
    "Project4.xml": ("p4", (240, 180)),
    "Project5/scene.xml": ("p5", (240, 180)),            # instanced meshes under a group node (depth 2)
    "Project5/scene-low.xml": ("p5low", (240, 180)),
    "Project7/scene.xml": ("p7", (240, 180)),
    "Project11/scene.xml": ("p11", (240, 180)),
    "Project13/scene.xml": ("p13", (240, 180)),
    "Teapot/scene.xml": ("teapot1", (480, 270)),
    "Teapot/scene2.xml": ("teapot2", (480, 270)),
    "ObjMtl/scene.xml": ("objmtl", (320, 240)),           # OBJ with usemtl / mtllib -> MultiMtl, PPM texture
}
# deterministic Whitted scenes: no soft lights, no glossy reflection/refraction, no DOF
WHITTED = {
    "Project2.xml": ("p2", (240, 180), 1),
    "Project3Simple.xml": ("p3simple", (240, 180), 1),
    "Project3Box.xml": ("p3box", (240, 180), 1),
    "Project4.xml": ("p4", (240, 180), 1),
    "Project5/scene.xml": ("p5", (240, 180), 1),
    "Project7/scene.xml": ("p7", (240, 180), 1),
    "Project11/scene.xml": ("p11", (240, 180), 1),
    "Project13/scene.xml": ("p13", (240, 180), 1),
    "Teapot/scene2.xml": ("teapot2", (480, 270), 1),
    "ObjMtl/scene.xml": ("objmtl", (320, 240), 1),
}
# multi-sample (reference Halton pattern) fixtures
WHITTED_SPP = {
    "Project4.xml": ("p4_spp4", (200, 150), 4),
    "Teapot/scene2.xml": ("teapot2_spp4", (240, 135), 4),
}
# full-size headline fixtures (BASELINE configs 2 and 3)
FULL = {
    "Project1Example.xml": ("p1example_full", (800, 600)),   # BASELINE config 1 at native size
    "Project4.xml": ("p4_full", (800, 600)),
    "Teapot/scene2.xml": ("teapot2_1080p", (1920, 1080)),
}
LOADER = ["ObjMtl/scene.xml", "Project1Example.xml", "Project4.xml", "Project5/scene.xml", "Project7/scene.xml", "Project9/scene.xml",
          "Project10/scene.xml", "Project11/scene_86.xml", "Project13/scene.xml", "Teapot/scene.xml", "Teapot/scene2.xml"]
# generated scenes of section 8d (tools/make_synthetic.py): 1 M-triangle mesh, flat lists of spheres
SYNTHETIC = {"grid1M": (240, 135), "spheres_100": (240, 135), "spheres_1000": (240, 135), "dupmesh": (480, 270), "manymtl": (240, 135)}
TEX = {"Project7/scene.xml": "p7", "Project9/scene.xml": "p9", "Project10/scene.xml": "p10"}
# stochastic branches of Shade / Render (rand()-driven in the reference): thin-lens depth of field (RenderFunctions.cpp:88-97),
# soft shadows = one disk sample of radius `size` per shadow ray (lightFunctions.cpp:39-84), glossy reflection / refraction
# lobes = SampleSphere offsets of the normal (mtlFunctions.cpp:163-165, 225-227, 276-278).  Converged means at 256 spp of
# the reference's sample pattern, two rand() seeds each: the second one is the noise yardstick.
STOCHASTIC = {
    "Project9/scene.xml": ("p9_dof", (160, 120), 256, None),
    "Teapot/scene.xml": ("teapot1_soft", (640, 360), 256, (272, 196, 432, 292)),   # the teapot and its soft shadows; the rest is background
    "Project10/scene.xml": ("p10", (160, 120), 256, None),
    "Project11/scene_glossy_soft.xml": ("p11_glossy_soft", (160, 120), 256, None),
    "Project11/scene_glossy.xml": ("p11_glossy", (160, 120), 256, None),
}


def run(scene, mode, prefix, *extra):
    cmd = [HARNESS, os.path.join(SCENES, scene), "--root", SCENES, "--mode", mode, "--out", prefix] + [str(e) for e in extra]
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
    line = [l for l in r.stderr.strip().splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def collect(prefix):
    out = {}
    for f in sorted(glob.glob(prefix + "_*.npy")):
        out[os.path.basename(f)[len(os.path.basename(prefix)) + 1:-4]] = np.load(f)
    return out


def save(name, arrays, meta):
    arrays = dict(arrays)
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-34s %8.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def main():
    if not os.path.exists(HARNESS):
        sys.exit("build the reference harness first: make -C oracle/ref")
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    with tempfile.TemporaryDirectory() as tmp:
        def want(kind):
            return not only or kind in only

        if want("kat"):
            pre = os.path.join(tmp, "kat")
            meta = run("Teapot/scene2.xml", "kat", pre, "--n", 16384, "--seed", 7)
            arrs = collect(pre)
            for prim in ("sphere", "plane", "mesh", "box"):
                sub = {k[len(prim) + 1:]: v for k, v in arrs.items() if k.startswith(prim + "_")}
                save("kat_" + prim, sub, dict(meta, scene="Teapot/scene2.xml", prim=prim))
        if want("loader"):
            for sc in LOADER:
                pre = os.path.join(tmp, "ld")
                for f in glob.glob(pre + "_*"):
                    os.remove(f)
                meta = run(sc, "dump", pre)
                arrs = collect(pre)
                if sc not in ("Teapot/scene2.xml", "Project5/scene.xml", "ObjMtl/scene.xml"):   # the same teapot.obj everywhere else
                    arrs = {k: v for k, v in arrs.items() if not k.startswith("mesh")}
                save("loader_" + sc.replace("/", "_").replace(".xml", ""), arrs, dict(meta, scene=sc))
        if want("primary"):
            for sc, (tag, (w, h)) in PRIMARY.items():
                pre = os.path.join(tmp, "pr")
                meta = run(sc, "primary", pre, "--width", w, "--height", h, "--threads", 8)
                save("primary_" + tag, collect(pre), dict(meta, scene=sc))
        if want("whitted"):
            for table, pattern in ((WHITTED, "center"), (WHITTED_SPP, "ref")):
                for sc, (tag, (w, h), spp) in table.items():
                    pre = os.path.join(tmp, "wh")
                    meta = run(sc, "whitted", pre, "--width", w, "--height", h, "--spp", spp, "--pattern", pattern, "--threads", 8)
                    save("whitted_" + tag, collect(pre), dict(meta, scene=sc))
        if want("full"):
            for sc, (tag, (w, h)) in FULL.items():
                pre = os.path.join(tmp, "fp")
                meta = run(sc, "primary", pre, "--width", w, "--height", h, "--threads", 8)
                a = collect(pre)
                save("primary_" + tag, {k: a[k] for k in ("z", "node", "face", "front")}, dict(meta, scene=sc))
                if sc == "Project1Example.xml":   # no materials: Shade() would dereference NULL
                    continue
                pre = os.path.join(tmp, "fw")
                meta = run(sc, "whitted", pre, "--width", w, "--height", h, "--threads", 8)
                save("whitted_" + tag, collect(pre), dict(meta, scene=sc))
        if want("synthetic"):
            import hashlib
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import make_synthetic
            paths = make_synthetic.ensure(tuple(SYNTHETIC))
            for name, (w, h) in SYNTHETIC.items():
                pre = os.path.join(tmp, "sy")
                meta = run(paths[name], "primary", pre, "--width", w, "--height", h, "--threads", 8)
                a = collect(pre)
                arrs = {k: a[k] for k in ("z", "node", "face", "front")}
                pre = os.path.join(tmp, "sw")
                meta2 = run(paths[name], "whitted", pre, "--width", w, "--height", h, "--threads", 8)
                arrs["rgb"] = collect(pre)["rgb"]
                digest = {}
                for f in (name + ".xml", name + ".obj"):
                    fp = os.path.join(SCENES, "synthetic", f)
                    if os.path.exists(fp):
                        digest[f] = hashlib.sha256(open(fp, "rb").read()).hexdigest()
                save("synthetic_" + name, arrs, dict(meta, scene=paths[name], sha256=digest, whitted=meta2))
        if want("photonkat"):
            pre = os.path.join(tmp, "pk")
            meta = run("Project13/scene.xml", "photonkat", pre, "--n", 20000, "--seed", 3)
            save("kat_photonmap", collect(pre), meta)
        if want("photon"):
            # GeneratePhotonMap() is rand()-driven: the fixture keeps distributional summaries of two runs (the second
            # one is the noise yardstick) plus a low-resolution PhotonMapping() image
            stats = {}
            for run_i, seed in enumerate((5, 6)):
                pre = os.path.join(tmp, "ph%d" % run_i)
                meta = run("Project13/scene.xml", "photon", pre, "--width", 160, "--height", 120, "--threads", 8, "--seed", seed)
                a = collect(pre)
                ph = a["photons_balanced"].reshape(-1, 24)
                pos = ph[:, :12].copy().view("<f4")
                pw = ph[:, 12:16].copy().view("<f4")[:, 0]
                col = ph[:, 16:19].astype("f4") / 255.0
                H, edges = np.histogramdd(pos, bins=(8, 8, 8), range=((-32, 32), (-32, 32), (-20, 44)))
                Pw, _ = np.histogramdd(pos, bins=(8, 8, 8), range=((-32, 32), (-32, 32), (-20, 44)), weights=pw)
                stats["hist%d" % run_i] = H.astype("f8")
                stats["power%d" % run_i] = Pw
                stats["mean_color%d" % run_i] = (col * pw[:, None]).sum(0) / pw.sum()
                stats["rgb%d" % run_i] = a["rgb"]
                stats["irrad%d" % run_i] = a["irrad"]
                stats["node%d" % run_i] = a["node"]
                stats["meta%d" % run_i] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
            save("photon_Project13", stats, dict(scene="Project13/scene.xml", width=160, height=120, bins=[8, 8, 8],
                                                 range=[[-32, 32], [-32, 32], [-20, 44]]))
        if want("stochastic"):
            for sc, (tag, (w, h), spp, crop) in STOCHASTIC.items():
                arrs, metas = {}, []
                for i, seed in enumerate((101, 202)):
                    pre = os.path.join(tmp, "st%d" % i)
                    # one thread: rand() is process-wide state, so a multi-threaded run is not reproducible
                    meta = run(sc, "whitted", pre, "--width", w, "--height", h, "--spp", spp, "--pattern", "ref", "--threads", 1, "--seed", seed,
                               *(("--crop",) + crop if crop else ()))
                    arrs["rgb%d" % i] = collect(pre)["rgb"]
                    metas.append(meta)
                save("stochastic_" + tag, arrs, dict(metas[0], scene=sc, seeds=[101, 202], trace_rays=[m["trace_rays"] for m in metas],
                                                     shadow_rays=[m["shadow_rays"] for m in metas]))
        if want("tex"):
            for sc, tag in TEX.items():
                pre = os.path.join(tmp, "tx")
                for f in glob.glob(pre + "_*"):
                    os.remove(f)
                meta = run(sc, "tex", pre, "--n", 4096, "--seed", 11)
                a = collect(pre)
                save("tex_" + tag, {k[4:]: v for k, v in a.items()}, dict(meta, scene=sc))


if __name__ == "__main__":
    main()
