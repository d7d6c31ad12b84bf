// Light masks: where a mesh can cast a hard shadow at all, as seen from one light (device_scene.h LightMask).
//
// Every shadow ray of a point light without size ends in the light (lightFunctions.cpp:76-78), every shadow ray of a
// directional light runs against its direction (lights.h:48).  Seen from the light such a ray is ONE POINT: of the
// perspective image of the scene around the axis light -> mesh centre, or of the orthographic image along the direction.  A
// triangle can only stop the ray when its own image covers that point.  The mask is a RTU_MASK_RES^2 bitmap over the image
// of the mesh; a cell is set when a triangle's image touches it or one of its 8 neighbours (rasterised row by row from
// the triangle's edges, then dilated: the mask follows the silhouette, not the triangles' rectangles).  The any-hit kernel
// looks a ray up after the mesh's bound-box gate (objFunctions.cpp:337) and skips the walk when the cell is clear.
//
// Only the boolean of ShadowTrace is observable, so the mask has to be conservative and nothing else.  What protects it:
//  * images are evaluated here in double from the float vertices; the device evaluates the ray's image in float.  A mask is
//    only built when a cell is at least 100 x the bound of that rounding (err_u below) ...
//  * ... and when the deviation the device tolerates in "this ray ends in the light" (1e-5 of the distance) / "this ray is
//    parallel to the direction" (sin < 4e-6), carried to the mesh, is below a quarter of a cell.  Both bounds grow with the
//    distance of the ray's origin: `lim` is the largest distance (1-norm) up to which they hold; farther rays are walked.
//  * every vertex has to lie in front of the light (perspective) by a clear margin; otherwise there is no mask.
// The same for camera rays (`eye`): without depth of field they all START in one point, the image is the perspective one around
// the axis eye -> mesh, `lim` the distance of a ray's origin from that point up to which it counts as the eye.
// Anything unusual (NaN, a flat image, a light inside the mesh's box, more than RTU_MASKS_PER_NODE hard lights) gives no
// mask and the rays are walked as before.
// Light lists (build_light_mask, second half): the same image also says WHICH triangles a ray of a cell can meet; with the
// triangles of every cell listed in the order of their distance from the light, a shadow ray tests a handful of them and
// walks no hierarchy at all.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "host_scene.h"

namespace rtu {

namespace {

const int RES = 256; // RTU_MASK_RES of csrc/device_scene.h (checked there by a static_assert on the record size only)
const double EPS = 5.96e-8; // 2^-24

inline double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

void basis(const double *a, double *e1, double *e2)
{
    int k = std::fabs(a[0]) <= std::fabs(a[1]) ? (std::fabs(a[0]) <= std::fabs(a[2]) ? 0 : 2) : (std::fabs(a[1]) <= std::fabs(a[2]) ? 1 : 2);
    double t[3] = {0, 0, 0};
    t[k] = 1;
    e1[0] = a[1] * t[2] - a[2] * t[1]; e1[1] = a[2] * t[0] - a[0] * t[2]; e1[2] = a[0] * t[1] - a[1] * t[0];
    double n = std::sqrt(dot(e1, e1));
    for (int i = 0; i < 3; i++) e1[i] /= n;
    e2[0] = a[1] * e1[2] - a[2] * e1[1]; e2[1] = a[2] * e1[0] - a[0] * e1[2]; e2[2] = a[0] * e1[1] - a[1] * e1[0];
    n = std::sqrt(dot(e2, e2));
    for (int i = 0; i < 3; i++) e2[i] /= n;
}

// set cells ix0 .. ix1 of row iy
inline void set_span(uint32_t *row, int ix0, int ix1)
{
    for (int w = ix0 >> 5; w <= ix1 >> 5; w++) {
        const int b0 = std::max(ix0 - w * 32, 0), b1 = std::min(ix1 - w * 32, 31);
        row[w] |= (0xffffffffu >> (31 - b1)) & (0xffffffffu << b0);
    }
}

inline bool span_set(const uint32_t *row, int ix0, int ix1)
{
    for (int w = ix0 >> 5; w <= ix1 >> 5; w++) {
        const int b0 = std::max(ix0 - w * 32, 0), b1 = std::min(ix1 - w * 32, 31);
        const uint32_t mask = (0xffffffffu >> (31 - b1)) & (0xffffffffu << b0);
        if ((row[w] & mask) != mask) return false;
    }
    return true;
}

inline int ifloor(double x) { return (int)(x + 16.0) - 16; } // -16 < x < 2^31 - 16 (callers clamp)

// The cells a triangle (cell coordinates) touches: row by row, the extent in x of the triangle's part inside the row
// [iy, iy+1] (the ends of its edges clipped to the row; the intersection is convex).  The ends are widened by 1e-6 of a cell:
// the arithmetic here is double, its rounding is far below that.
struct Rows {
    int iy0, iy1;           // rows the triangle reaches (clamped to the bitmap); iy0 > iy1: none
    int rx0, rx1;           // columns of its rectangle
    int lo[RES], hi[RES];   // per row iy the cells lo[iy] .. hi[iy]; hi < lo: none
};

void tri_rect(const double (*p)[2], Rows &R)
{
    const double big = RES + 8.0;
    double y0 = std::min(p[0][1], std::min(p[1][1], p[2][1])), y1 = std::max(p[0][1], std::max(p[1][1], p[2][1]));
    double xa = std::min(p[0][0], std::min(p[1][0], p[2][0])), xb = std::max(p[0][0], std::max(p[1][0], p[2][0]));
    y0 = std::max(y0, -8.0); y1 = std::min(y1, big); xa = std::max(xa, -8.0); xb = std::min(xb, big);
    R.iy0 = std::max(0, ifloor(y0 - 1e-6)); R.iy1 = std::min(RES - 1, ifloor(y1 + 1e-6));
    R.rx0 = std::max(0, ifloor(xa - 1e-6)); R.rx1 = std::min(RES - 1, ifloor(xb + 1e-6));
    if (R.rx0 > R.rx1) R.iy1 = R.iy0 - 1;
}

void tri_rows(const double (*p)[2], Rows &R)
{
    const double big = RES + 8.0;
    // per edge: x(t), y(t) = a + t (b - a); the part of it inside a row lo <= y <= hi is t in [max(0, ta), min(1, tb)]
    double ax[3], ay[3], dx[3], dy[3], inv[3];
    for (int k = 0; k < 3; k++) {
        ax[k] = p[k][0]; ay[k] = p[k][1];
        dx[k] = p[(k + 1) % 3][0] - ax[k]; dy[k] = p[(k + 1) % 3][1] - ay[k];
        inv[k] = dy[k] != 0.0 ? 1.0 / dy[k] : 0.0;
    }
    for (int iy = R.iy0; iy <= R.iy1; iy++) {
        const double lo = iy - 1e-6, hi = iy + 1.0 + 1e-6;
        double x0 = 1e300, x1 = -1e300;
        for (int k = 0; k < 3; k++) {
            double ta, tb;
            if (dy[k] != 0.0) {
                const double t0 = (lo - ay[k]) * inv[k], t1 = (hi - ay[k]) * inv[k];
                ta = std::max(0.0, std::min(t0, t1)); tb = std::min(1.0, std::max(t0, t1));
            } else {
                ta = 0.0; tb = (ay[k] >= lo && ay[k] <= hi) ? 1.0 : -1.0;
            }
            if (ta <= tb) {
                const double xa2 = ax[k] + ta * dx[k], xb2 = ax[k] + tb * dx[k];
                x0 = std::min(x0, std::min(xa2, xb2)); x1 = std::max(x1, std::max(xa2, xb2));
            }
        }
        R.lo[iy] = 1; R.hi[iy] = 0;
        if (!(x0 <= x1)) continue; // the triangle does not reach this row
        x0 = std::max(x0, -8.0); x1 = std::min(x1, big);
        R.lo[iy] = std::max(0, ifloor(x0 - 1e-6)); R.hi[iy] = std::min(RES - 1, ifloor(x1 + 1e-6));
    }
}

void raster(const double (*p)[2], uint32_t *bits, Rows &R)
{
    tri_rect(p, R);
    if (R.iy0 > R.iy1) return;
    {   // every cell of the triangle's rectangle already set (by its neighbours, by the other side of a closed surface)?
        bool all = true;
        for (int iy = R.iy0; iy <= R.iy1 && all; iy++) all = span_set(bits + (size_t)iy * (RES / 32), R.rx0, R.rx1);
        if (all) return;
    }
    if (R.rx1 - R.rx0 <= 1 && R.iy1 - R.iy0 <= 1) { // up to 2 x 2 cells: the rectangle
        for (int iy = R.iy0; iy <= R.iy1; iy++) set_span(bits + (size_t)iy * (RES / 32), R.rx0, R.rx1);
        return;
    }
    tri_rows(p, R);
    for (int iy = R.iy0; iy <= R.iy1; iy++)
        if (R.lo[iy] <= R.hi[iy]) set_span(bits + (size_t)iy * (RES / 32), R.lo[iy], R.hi[iy]);
}

// The cells whose 3 x 3 neighbourhood the triangle touches (what dilate() makes of its own cells), row by row: fn(iy, ix0, ix1)
template <class F> void dilated_rows(const double (*p)[2], Rows &R, F fn)
{
    tri_rect(p, R);
    if (R.iy0 > R.iy1) return;
    tri_rows(p, R);
    for (int r = std::max(0, R.iy0 - 1); r <= std::min(RES - 1, R.iy1 + 1); r++) {
        int lo = RES, hi = -1;
        for (int q = std::max(R.iy0, r - 1); q <= std::min(R.iy1, r + 1); q++)
            if (R.lo[q] <= R.hi[q]) { lo = std::min(lo, R.lo[q]); hi = std::max(hi, R.hi[q]); }
        if (hi < lo) continue;
        fn(r, std::max(0, lo - 1), std::min(RES - 1, hi + 1));
    }
}

// every cell takes the OR of its 3 x 3 neighbourhood: the margin of one cell on every side that covers the device's float
// evaluation of a ray's image and the tolerances of its ray classification (see the head of this file)
void dilate(std::vector<uint32_t> &bits)
{
    const int W = RES / 32;
    std::vector<uint32_t> h(bits.size());
    for (int y = 0; y < RES; y++)
        for (int w = 0; w < W; w++) {
            const uint32_t c = bits[y * W + w];
            uint32_t v = c | (c << 1) | (c >> 1);
            if (w > 0) v |= bits[y * W + w - 1] >> 31;
            if (w + 1 < W) v |= bits[y * W + w + 1] << 31;
            h[y * W + w] = v;
        }
    for (int y = 0; y < RES; y++)
        for (int w = 0; w < W; w++)
            bits[y * W + w] = h[y * W + w] | (y > 0 ? h[(y - 1) * W + w] : 0u) | (y + 1 < RES ? h[(y + 1) * W + w] : 0u);
}

} // namespace

// The light in the node's coordinates: p' = itm (p - pos), d' = itm d, level by level (column-major itm).  false: a light that
// gets no mask (ambient, soft), or not finite.
static bool light_in_node(const rtu_node *const *chain, int n_chain, const rtu_light &light, double *L)
{
    const bool point = light.kind == RTU_LIGHT_POINT;
    if (!point && light.kind != RTU_LIGHT_DIRECT) return false;
    if (point && !(light.size == 0.f)) return false; // a soft light's rays end on its disk
    for (int k = 0; k < 3; k++) L[k] = light.v[k];
    for (int c = 0; c < n_chain; c++) {
        const rtu_node &t = *chain[c];
        double q[3] = {L[0], L[1], L[2]};
        if (point) for (int k = 0; k < 3; k++) q[k] -= (double)t.pos[k];
        for (int r = 0; r < 3; r++) L[r] = (double)t.itm[r] * q[0] + (double)t.itm[3 + r] * q[1] + (double)t.itm[6 + r] * q[2];
    }
    return std::isfinite(L[0]) && std::isfinite(L[1]) && std::isfinite(L[2]);
}

// rec: the 24 words of a device LightMask (bits / lists offsets left 0).  Returns false when no mask can be given.
// chain: the nodes from the root down to the mesh node (ToNodeCoords is applied in that order, RenderFunctions.cpp:186).
bool build_light_mask(const rtu_node *const *chain, int n_chain, const rtu_mesh &m, const rtu_light &light, float *rec,
                      std::vector<uint32_t> *bits, bool eye, LightLists *lists)
{
    if (lists) { lists->cell_start.clear(); lists->items.clear(); }
    if (m.nf == 0 || !m.v || !m.f) return false;
    const bool point = light.kind == RTU_LIGHT_POINT;
    double L[3];
    if (!light_in_node(chain, n_chain, light, L)) return false;
    double a[3], e1[3], e2[3];
    if (point) {
        for (int k = 0; k < 3; k++) a[k] = 0.5 * ((double)m.bound_min[k] + (double)m.bound_max[k]) - L[k];
    } else {
        for (int k = 0; k < 3; k++) a[k] = L[k];
    }
    const double an = std::sqrt(dot(a, a));
    if (!(an > 0.0) || !std::isfinite(an)) return false;
    for (int k = 0; k < 3; k++) a[k] /= an;
    basis(a, e1, e2);

    // images of the vertices
    std::vector<double> img((size_t)m.nv * 2), dep(m.nv, 0.0);
    std::vector<uint8_t> used(m.nv, 0);
    for (size_t k = 0; k < (size_t)m.nf * 3; k++) {
        if (m.f[k] >= m.nv) return false;
        used[m.f[k]] = 1;
    }
    double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300}, dmin = 1e300, vmax1 = 0;
    for (uint32_t i = 0; i < m.nv; i++) {
        if (!used[i]) continue;
        double w[3] = {(double)m.v[3 * i], (double)m.v[3 * i + 1], (double)m.v[3 * i + 2]};
        vmax1 = std::max(vmax1, std::fabs(w[0]) + std::fabs(w[1]) + std::fabs(w[2]));
        double u, v;
        if (point) {
            for (int k = 0; k < 3; k++) w[k] -= L[k];
            const double depth = dot(w, a);
            if (!(depth > 0.0)) return false;
            dmin = std::min(dmin, depth);
            dep[i] = depth;
            u = dot(w, e1) / depth;
            v = dot(w, e2) / depth;
        } else {
            dep[i] = dot(w, a); // along the light's direction: a point between p and the light has a smaller one
            u = dot(w, e1);
            v = dot(w, e2);
        }
        if (!(std::isfinite(u) && std::isfinite(v))) return false;
        img[2 * i] = u; img[2 * i + 1] = v;
        lo[0] = std::min(lo[0], u); hi[0] = std::max(hi[0], u);
        lo[1] = std::min(lo[1], v); hi[1] = std::max(hi[1], v);
    }
    if (!(hi[0] > lo[0]) || !(hi[1] > lo[1])) return false; // flat image
    const double cell[2] = {(hi[0] - lo[0]) / (RES - 2), (hi[1] - lo[1]) / (RES - 2)}; // the image spans cells 1 .. RES-2
    const double cmin = std::min(cell[0], cell[1]);
    const double umax = std::max(std::max(std::fabs(lo[0]), std::fabs(hi[0])), std::max(std::fabs(lo[1]), std::fabs(hi[1]))) + 2 * std::max(cell[0], cell[1]);
    const double l1 = std::fabs(L[0]) + std::fabs(L[1]) + std::fabs(L[2]);
    double lim;
    if (point) {
        if (!(dmin > 1e-3 * an) || umax > 4.0) return false; // the light is in or next to the mesh: too wide an image
        // float evaluation of u = (w.e1)/(w.a), w = p - L, for an origin whose image is inside the bitmap: a component of w is off
        // by eps (|w_k| + |L_k|), a dot product by eps (5 |w|_1 + |L|_1) incl. the rounding of the axes, |w|_1 <= sqrt(3) depth
        // sqrt(1 + 2 umax^2), depth >= dmin for every origin behind a triangle; twice that
        const double err_u = 2 * EPS * ((1.0 + umax) * (9.0 * std::sqrt(1.0 + 2.0 * umax * umax) + l1 / dmin) + umax);
        if (!(100 * err_u <= cmin)) return false;
        // a ray that passes the light at 1e-5 |w|_1 is off by that much at the mesh, i.e. by (1 + umax) / depth of it in u
        lim = cmin * dmin / (4 * 1.0e-5 * (1.0 + umax));
        if (!eye && !(lim > 2 * an)) return false;
        // eye mask: rays START in the point; an origin off by t moves the image of a triangle by t (1 + umax) / depth of it
        if (eye) lim = cmin * dmin / (4 * (1.0 + umax));
    } else {
        // u = p.e1: rounding 4 eps |p|_1; a direction off by sin = 4e-6 drifts by that times the way to the mesh (<= |p|_1 + vmax1)
        lim = std::min(cmin / (400 * EPS), cmin / (4 * 4.0e-6) - vmax1);
        if (!(lim > 2 * vmax1)) return false;
    }
    if (!std::isfinite(lim)) return false;
    lim = std::min(lim, 1.0e30);

    const double s[2] = {1.0 / cell[0], 1.0 / cell[1]};
    const double o[2] = {lo[0] - cell[0], lo[1] - cell[1]};
    bits->assign((size_t)RES * RES / 32, 0u);
    Rows R;
    auto image_of = [&](uint32_t f, double (*p)[2]) {
        for (int k = 0; k < 3; k++) {
            const uint32_t i = m.f[3 * f + k];
            p[k][0] = (img[2 * i] - o[0]) * s[0];
            p[k][1] = (img[2 * i + 1] - o[1]) * s[1];
        }
    };
    for (uint32_t f = 0; f < m.nf; f++) {
        double p[3][2];
        image_of(f, p);
        raster(p, bits->data(), R);
    }
    dilate(*bits);

    // Light lists: per cell the triangles whose image touches its 3 x 3 neighbourhood, as (slot in DMesh::tris, least depth of the
    // triangle) pairs in the order of that depth.  A recognised shadow ray tests the triangles of its cell that begin before
    // its own origin instead of walking the mesh's hierarchy (k_shadow_wave); every triangle the exact test could accept is
    // among them by the argument that makes the bitmap conservative.  Only where the lists stay short (triangles that are
    // not much smaller than a cell) and the mesh has a slot order the host knows.
    double zmargin = 0.0;
    if (lists && !eye && m.nf <= (1u << 24)) {
        std::vector<uint32_t> slot_of(m.nf);
        bool have = true;
        if (m.flags & RTU_MESH_DEVICE_BVH) for (uint32_t f = 0; f < m.nf; f++) slot_of[f] = f;
        else if (m.bvh_elements) {
            for (uint32_t k = 0; k < m.nf && have; k++) { have = m.bvh_elements[k] < m.nf; if (have) slot_of[m.bvh_elements[k]] = k; }
        } else have = false;
        std::vector<uint32_t> count((size_t)RES * RES + 1, 0u);
        uint64_t total = 0;
        for (uint32_t f = 0; f < m.nf && have; f++) {
            double p[3][2];
            image_of(f, p);
            dilated_rows(p, R, [&](int iy, int ix0, int ix1) { for (int ix = ix0; ix <= ix1; ix++) count[(size_t)iy * RES + ix + 1]++; total += (uint64_t)(ix1 - ix0 + 1); });
        }
        size_t cells_set = 0;
        for (size_t c = 1; c < count.size(); c++) cells_set += count[c] != 0;
        if (have && total > 0 && total <= 24 * (uint64_t)cells_set && total < (1u << 26)) {
            for (size_t c = 1; c < count.size(); c++) count[c] += count[c - 1];
            lists->cell_start = count; // RES * RES + 1 offsets
            std::vector<uint32_t> fill(count.begin(), count.end() - 1);
            lists->items.assign((size_t)total * 2, 0u);
            double zabs = 0;
            for (uint32_t f = 0; f < m.nf; f++) {
                double p[3][2];
                image_of(f, p);
                const double z = std::min(dep[m.f[3 * f]], std::min(dep[m.f[3 * f + 1]], dep[m.f[3 * f + 2]]));
                zabs = std::max(zabs, std::fabs(z));
                float zf = (float)z;
                if ((double)zf > z) zf = std::nextafterf(zf, -3.0e38f); // rounded down
                uint32_t zb;
                memcpy(&zb, &zf, 4);
                dilated_rows(p, R, [&](int iy, int ix0, int ix1) {
                    for (int ix = ix0; ix <= ix1; ix++) {
                        const size_t at = fill[(size_t)iy * RES + ix]++;
                        lists->items[2 * at] = slot_of[f];
                        lists->items[2 * at + 1] = zb;
                    }
                });
            }
            // per cell in the order of depth (ties by slot: the same lists whatever the order of the faces)
            std::vector<std::pair<float, uint32_t>> seg;
            for (size_t c = 0; c + 1 < lists->cell_start.size(); c++) {
                const size_t b0 = lists->cell_start[c], b1 = lists->cell_start[c + 1];
                if (b1 - b0 < 2) continue;
                seg.clear();
                for (size_t k = b0; k < b1; k++) { float zf; memcpy(&zf, &lists->items[2 * k + 1], 4); seg.emplace_back(zf, lists->items[2 * k]); }
                std::sort(seg.begin(), seg.end());
                for (size_t k = b0; k < b1; k++) { lists->items[2 * k] = seg[k - b0].second; memcpy(&lists->items[2 * k + 1], &seg[k - b0].first, 4); }
            }
            // the device's depth of a ray's origin: a dot product of (p - L) / p with a unit axis in float, rounding
            // <= eps (5 |p - L|_1 + |L|_1) (see err_u); |p - L|_1 <= lim by the lookup's own check.  Far more than that:
            zmargin = 1.0e-4 * (zabs + l1 + vmax1 + 1.0e-30) + 64 * EPS * lim;
        }
    }

    memset(rec, 0, 24 * sizeof(float));
    const int32_t kind = eye ? RTU_MASK_EYE : light.kind;
    for (int k = 0; k < 3; k++) { rec[k] = (float)L[k]; rec[4 + k] = (float)a[k]; rec[8 + k] = (float)e1[k]; rec[12 + k] = (float)e2[k]; }
    memcpy(&rec[3], &kind, 4);
    rec[7] = (float)o[0];
    rec[11] = (float)o[1];
    rec[15] = (float)s[0];
    rec[16] = (float)s[1];
    rec[18] = (float)lim;
    rec[19] = (float)zmargin; // words 20 / 21: where the lists are (rtu_scene_upload)
    return true;
}

// budget: triangles this call may rasterise (a bitmap costs a mesh's faces once, ~0.15 us each; its lists 8 times that).  The
// loader spends what a scene takes; rtu_scene_upload, for a description that comes without masks, stays within a few tenths
// of a second and leaves the pairs beyond that without mask / without lists.
void collect_light_masks(const rtu_scene_desc &d, std::vector<rtu_light_mask> *out, std::vector<OwnedMask> *own, uint64_t budget)
{
    out->clear();
    own->clear();
    if (const char *e = getenv("RTU_LIGHT_MASKS")) if (atoi(e) == 0) return;
    bool lists_on = true;
    if (const char *e = getenv("RTU_LIGHT_LISTS")) lists_on = atoi(e) != 0;
    int hard = 0;
    for (int l = 0; l < d.n_lights; l++)
        hard += d.lights[l].kind == RTU_LIGHT_DIRECT || (d.lights[l].kind == RTU_LIGHT_POINT && d.lights[l].size == 0.f);
    // the lookup steps through a node's masks: scenes with many hard lights get none for them
    const bool lights_on = hard > 0 && hard <= 4, eye_on = !(d.camera.dof > 0.f);
    if (!lights_on && !eye_on) return;
    rtu_light eye;
    memset(&eye, 0, sizeof eye);
    eye.kind = RTU_LIGHT_POINT;
    memcpy(eye.v, d.camera.pos, sizeof eye.v);
    own->reserve(1024); // pointers into it are handed out
    for (int i = 0; i < d.n_nodes && out->size() + 5 <= 1024; i++) {
        if (d.nodes[i].kind != RTU_OBJ_MESH || d.nodes[i].mesh < 0 || d.nodes[i].mesh >= d.n_meshes) continue;
        const rtu_node *chain[64];
        int n_chain = 0;
        for (int a = i; a >= 0 && n_chain < 64; a = d.nodes[a].parent) n_chain++;
        if (n_chain >= 64) continue;
        for (int a = i, k = n_chain - 1; k >= 0; a = d.nodes[a].parent, k--) chain[k] = &d.nodes[a];
        for (int l = 0; l <= d.n_lights; l++) { // the eye's mask comes last
            const bool is_eye = l == d.n_lights;
            if (is_eye ? !eye_on : !lights_on) continue;
            const rtu_light &lt = is_eye ? eye : d.lights[l];
            double L[3];
            if (!light_in_node(chain, n_chain, lt, L)) continue;
            const rtu_light_mask *pre = nullptr;
            for (int k = 0; k < d.n_light_masks && d.light_masks && !pre; k++) {
                const rtu_light_mask &c = d.light_masks[k];
                int32_t kind;
                memcpy(&kind, &c.rec[3], 4);
                if (c.node == i && c.light == (is_eye ? -1 : l) && c.bits && kind == (is_eye ? (int32_t)RTU_MASK_EYE : lt.kind) &&
                    c.rec[0] == (float)L[0] && c.rec[1] == (float)L[1] && c.rec[2] == (float)L[2])
                    pre = &c;
            }
            if (pre) {
                out->push_back(*pre);
                if (!lists_on) { out->back().cell_start = nullptr; out->back().items = nullptr; out->back().n_items = 0; }
                continue;
            }
            rtu_light_mask lm;
            memset(&lm, 0, sizeof lm);
            OwnedMask om;
            const uint64_t nf = d.meshes[d.nodes[i].mesh].nf;
            if (nf > budget) continue;
            const bool with_lists = lists_on && !is_eye && 9 * nf <= budget;
            if (!build_light_mask(chain, n_chain, d.meshes[d.nodes[i].mesh], lt, lm.rec, &om.bits, is_eye, with_lists ? &om.lists : nullptr)) continue;
            budget -= with_lists ? 9 * nf : nf;
            own->push_back(std::move(om));
            const OwnedMask &k = own->back();
            lm.node = i;
            lm.light = is_eye ? -1 : l;
            lm.bits = k.bits.data();
            if (!k.lists.cell_start.empty()) {
                lm.cell_start = k.lists.cell_start.data();
                lm.items = k.lists.items.data();
                lm.n_items = (uint32_t)(k.lists.items.size() / 2);
            }
            out->push_back(lm);
        }
    }
}

} // namespace rtu
