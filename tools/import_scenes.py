#!/usr/bin/env python3
"""Import the reference's scene *assets* (input data, not source code) into scenes/.

The reference's XML scene files hard-code the author's absolute paths
(/Users/Peter/GitRepos/RayTracer-Utah/SceneFiles/...), so they cannot
be used verbatim anywhere.  This script rewrites that prefix to a path relative to the
scene root (scenes/), and copies the OBJ meshes and the two PNG textures the scenes
reference.  Both the reference harness (oracle/ref) and our loader resolve relative asset
paths against the scene root, so both sides see exactly the same bytes.

Files that are missing from the reference checkout (Teapot/ink.png, Teapot/ink2.png; see
.MISSING_LARGE_BLOBS) stay missing: a failed texture load samples black on both sides
(xmlload.cpp:542-546, scene.h:382).

Run once in the build container (needs /root/reference); the GPU box only ever sees scenes/.
"""
import os
import shutil
import sys

REF = os.environ.get("RTU_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "SceneFiles")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes")
PREFIX = "/Users/Peter/GitRepos/RayTracer-Utah/SceneFiles/"

COPY_EXT = (".obj", ".png")
SKIP = {"inkoriginal.png", "ink2original.png"}  # 6 MB plates that no scene references by that name


def main():
    if not os.path.isdir(SRC):
        sys.exit("reference SceneFiles not found at %s" % SRC)
    n = 0
    for root, _dirs, files in os.walk(SRC):
        rel = os.path.relpath(root, SRC)
        out = os.path.join(DST, rel) if rel != "." else DST
        os.makedirs(out, exist_ok=True)
        for f in sorted(files):
            s = os.path.join(root, f)
            d = os.path.join(out, f)
            if f.endswith(".xml"):
                txt = open(s, "r", encoding="utf-8", errors="replace").read()
                txt = txt.replace(PREFIX, "")
                open(d, "w", encoding="utf-8").write(txt)
                n += 1
            elif f.endswith(COPY_EXT) and f not in SKIP:
                shutil.copyfile(s, d)
                n += 1
    print("imported %d files into %s" % (n, DST))


if __name__ == "__main__":
    main()
