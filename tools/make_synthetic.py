#!/usr/bin/env python3
"""Synthetic scenes of named shape, written in the reference's own formats so that the
reference loader / cyBVH builder ingest them too.  The camera is deliberately off the grid's symmetry plane: with x = 0 the reference sample pattern (offset 0 for
sample 0) sends a whole pixel column exactly along a grid line, and each such ray walks ~5000 boxes / 4000 triangles
(reference behaviour, reproduced, but a 40 ms serial tail that says nothing about throughput).
Deterministic (seed 0x5EED); generated on demand into
scenes/synthetic/ (git-ignored: grid1M.obj is ~45 MB).

  grid1M      1000 x 500 quads (1 000 000 triangles after fan triangulation) of a displaced height field
  soup1M      1 000 000 random small triangles in [-10,10]^3 (side ~0.05)
  spheres_N   N random spheres, flat scene graph (N = 100, 1000, 10000)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "scenes", "synthetic")
SEED = 0x5EED

CAMERA = """  <camera>
    <position x="1.7" y="-%(dist)s" z="%(height)s"/>
    <target x="0.3" y="0" z="0"/>
    <up x="0" y="0" z="1"/>
    <fov value="40"/>
    <width value="%(w)d"/>
    <height value="%(h)d"/>
  </camera>
"""
MATERIALS = """    <material type="blinn" name="matte">
      <diffuse r="0.7" g="0.75" b="0.6"/>
      <specular value="0.3"/>
      <glossiness value="20"/>
    </material>
    <material type="blinn" name="mirror">
      <diffuse value="0.1"/>
      <specular value="0.8"/>
      <glossiness value="60"/>
      <reflection value="0.7"/>
    </material>
    <material type="blinn" name="glass">
      <diffuse value="0"/>
      <specular value="0.9"/>
      <glossiness value="80"/>
      <refraction value="0.9" index="1.5"/>
    </material>
    <light type="ambient" name="amb"><intensity value="0.15"/></light>
    <light type="direct" name="sun"><intensity value="0.8"/><direction x="-0.4" y="0.5" z="-1"/></light>
    <light type="point" name="lamp"><intensity value="300"/><position x="6" y="-8" z="20"/></light>
"""


def write_grid(path, nx=1000, ny=500):
    rng = np.random.default_rng(SEED)
    xs = np.linspace(-10, 10, nx + 1)
    ys = np.linspace(-5, 5, ny + 1)
    X, Y = np.meshgrid(xs, ys)
    Z = 0.6 * np.sin(X * 1.3) * np.cos(Y * 1.7) + 0.25 * np.sin(X * 5.1 + Y * 3.3) + 0.02 * rng.standard_normal(X.shape)
    with open(path, "w") as f:
        f.write("# grid1M: %d x %d quads\n" % (nx, ny))
        for x, y, z in zip(X.ravel(), Y.ravel(), Z.ravel()):
            f.write("v %.4f %.4f %.4f\n" % (x, y, z))
        for x, y in zip(X.ravel(), Y.ravel()):
            f.write("vt %.4f %.4f 0.0000\n" % ((x + 10) / 20, (y + 5) / 10))
        w = nx + 1
        for j in range(ny):
            row = []
            for i in range(nx):
                a = j * w + i + 1
                row.append("f %d/%d %d/%d %d/%d %d/%d\n" % (a, a, a + 1, a + 1, a + 1 + w, a + 1 + w, a + w, a + w))
            f.write("".join(row))


def write_soup(path, n=1000000):
    rng = np.random.default_rng(SEED + 1)
    c = rng.uniform(-10, 10, (n, 3))
    v = c[:, None, :] + rng.uniform(-0.05, 0.05, (n, 3, 3))
    with open(path, "w") as f:
        f.write("# soup1M: %d random triangles\n" % n)
        f.write("vt 0.0 0.0 0.0\nvt 1.0 0.0 0.0\nvt 0.0 1.0 0.0\n")
        lines = []
        for t in range(n):
            for k in range(3):
                lines.append("v %.4f %.4f %.4f\n" % tuple(v[t, k]))
            if len(lines) > 300000:
                f.write("".join(lines)); lines = []
        f.write("".join(lines)); lines = []
        for t in range(n):
            a = 3 * t + 1
            lines.append("f %d/1 %d/2 %d/3\n" % (a, a + 1, a + 2))
            if len(lines) > 300000:
                f.write("".join(lines)); lines = []
        f.write("".join(lines))


def write_dupmesh(path, n=24):
    """A bumpy n x n quad patch whose every triangle exists twice (same three vertices): every hit is an exact tie in z,
    so the winner is decided by the reference's visiting order alone."""
    rng = np.random.default_rng(SEED + 7)
    xs = np.linspace(-6, 6, n + 1)
    X, Y = np.meshgrid(xs, xs)
    Z = 0.8 * np.sin(X * 0.9) * np.cos(Y * 0.7) + 0.05 * rng.standard_normal(X.shape)
    with open(path, "w") as f:
        f.write("# dupmesh: %d x %d quads, every triangle twice\n" % (n, n))
        for x, y, z in zip(X.ravel(), Y.ravel(), Z.ravel()):
            f.write("v %.4f %.4f %.4f\n" % (x, y, z))
        for x, y in zip(X.ravel(), Y.ravel()):
            f.write("vt %.4f %.4f 0.0000\n" % ((x + 6) / 12, (y + 6) / 12))
        w = n + 1
        tris = []
        for j in range(n):
            for i in range(n):
                a = j * w + i + 1
                tris.append((a, a + 1, a + 1 + w))
                tris.append((a, a + 1 + w, a + w))
        order = list(range(len(tris))) + list(rng.permutation(len(tris)))   # the copies come in a shuffled order
        for t in order:
            a, b, c = tris[t]
            f.write("f %d/%d %d/%d %d/%d\n" % (a, a, b, b, c, c))


def write_mesh_scene(path, obj, w, h, dist, height):
    with open(path, "w") as f:
        f.write("<xml>\n  <scene>\n    <background r=\"0.05\" g=\"0.06\" b=\"0.09\"/>\n    <environment value=\"0.2\"/>\n")
        f.write("    <object type=\"obj\" name=\"synthetic/%s\" material=\"matte\"/>\n" % obj)
        f.write("    <object type=\"sphere\" name=\"ball\" material=\"mirror\"><scale value=\"1.5\"/><translate x=\"-3\" y=\"-1\" z=\"2.5\"/></object>\n")
        f.write("    <object type=\"sphere\" name=\"lens\" material=\"glass\"><scale value=\"1.2\"/><translate x=\"3\" y=\"-2\" z=\"2.2\"/></object>\n")
        f.write(MATERIALS)
        f.write("  </scene>\n" + CAMERA % dict(dist=dist, height=height, w=w, h=h) + "</xml>\n")


def write_spheres(path, n, w=1920, h=1080):
    rng = np.random.default_rng(SEED + n)
    with open(path, "w") as f:
        f.write("<xml>\n  <scene>\n    <background r=\"0.05\" g=\"0.06\" b=\"0.09\"/>\n    <environment value=\"0.2\"/>\n")
        f.write("    <object type=\"plane\" name=\"floor\" material=\"matte\"><scale value=\"40\"/><translate z=\"-6\"/></object>\n")
        mats = ["matte", "matte", "matte", "mirror", "glass"]
        r = 4.0 / n ** (1 / 3)
        for i in range(n):
            c = rng.uniform(-8, 8, 3)
            s = r * rng.uniform(0.5, 1.5)
            f.write("    <object type=\"sphere\" name=\"s%d\" material=\"%s\"><scale value=\"%.4f\"/><translate x=\"%.4f\" y=\"%.4f\" z=\"%.4f\"/></object>\n"
                    % (i, mats[i % 5], s, c[0], c[1], c[2] * 0.6))
        f.write(MATERIALS)
        f.write("  </scene>\n" + CAMERA % dict(dist=34, height=12, w=w, h=h) + "</xml>\n")


def write_manymtl(path, n_mtl=48, n_light=24, w=1920, h=1080):
    """More lights and materials than the shading kernel stages in shared memory (16 / 32): 96 spheres over a floor,
    every sphere with its own Blinn material (every fifth a mirror, every seventh glass), point lights on a ring."""
    rng = np.random.default_rng(SEED + 4242)
    with open(path, "w") as f:
        f.write("<xml>\n  <scene>\n    <background r=\"0.05\" g=\"0.06\" b=\"0.09\"/>\n    <environment value=\"0.2\"/>\n")
        f.write("    <object type=\"plane\" name=\"floor\" material=\"m0\"><scale value=\"40\"/><translate z=\"-6\"/></object>\n")
        for i in range(96):
            c = rng.uniform(-8, 8, 3)
            f.write("    <object type=\"sphere\" name=\"s%d\" material=\"m%d\"><scale value=\"%.4f\"/><translate x=\"%.4f\" y=\"%.4f\" z=\"%.4f\"/></object>\n"
                    % (i, i % n_mtl, rng.uniform(0.6, 1.4), c[0], c[1], c[2] * 0.6))
        for i in range(n_mtl):
            d = rng.uniform(0.1, 0.9, 3)
            f.write("    <material type=\"blinn\" name=\"m%d\">\n      <diffuse r=\"%.3f\" g=\"%.3f\" b=\"%.3f\"/>\n      <specular value=\"%.3f\"/>\n      <glossiness value=\"%d\"/>\n"
                    % (i, d[0], d[1], d[2], rng.uniform(0.1, 0.8), int(rng.integers(5, 120))))
            if i % 5 == 3:
                f.write("      <reflection value=\"%.3f\"/>\n" % rng.uniform(0.3, 0.8))
            if i % 7 == 5:
                f.write("      <refraction index=\"%.3f\" value=\"%.3f\"/>\n" % (rng.uniform(1.2, 1.8), rng.uniform(0.4, 0.8)))
            f.write("    </material>\n")
        f.write("    <light type=\"ambient\" name=\"amb\"><intensity value=\"0.1\"/></light>\n")
        for i in range(n_light - 1):
            a = 2 * np.pi * i / (n_light - 1)
            f.write("    <light type=\"point\" name=\"l%d\"><intensity value=\"%.3f\"/><position x=\"%.4f\" y=\"%.4f\" z=\"%.4f\"/></light>\n"
                    % (i, rng.uniform(30, 80), 14 * np.cos(a), 14 * np.sin(a), rng.uniform(8, 16)))
        f.write("  </scene>\n" + CAMERA % dict(dist=34, height=12, w=w, h=h) + "</xml>\n")


def ensure(names=("grid1M", "spheres_100", "spheres_1000", "spheres_10000")):
    """Creates the requested synthetic scenes if missing; returns {name: xml path relative to scenes/}."""
    os.makedirs(OUT, exist_ok=True)
    out = {}
    for n in names:
        xml = os.path.join(OUT, n + ".xml")
        if n in ("grid1M", "soup1M", "dupmesh"):
            obj = os.path.join(OUT, n + ".obj")
            if not os.path.exists(obj):
                {"grid1M": write_grid, "soup1M": write_soup, "dupmesh": write_dupmesh}[n](obj)
            if not os.path.exists(xml):
                write_mesh_scene(xml, n + ".obj", 3840, 2160, 22 if n != "soup1M" else 40, 12)
        elif n == "manymtl":
            if not os.path.exists(xml):
                write_manymtl(xml)
        elif n.startswith("spheres_"):
            if not os.path.exists(xml):
                write_spheres(xml, int(n.split("_")[1]))
        else:
            raise ValueError(n)
        out[n] = "synthetic/%s.xml" % n
    return out


if __name__ == "__main__":
    print(ensure(tuple(sys.argv[1:]) or ("grid1M", "soup1M", "spheres_100", "spheres_1000", "spheres_10000")))
