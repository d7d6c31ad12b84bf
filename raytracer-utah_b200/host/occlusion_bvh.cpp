// The any-hit ("occlusion") hierarchy of a mesh: what ShadowTrace walks on the device instead of the cyBVH.
//
// Only the boolean of a shadow ray is observable (GenLight::Shadow, lightFunctions.cpp:27-37), so the any-hit kernel is free
// to FIND candidate triangles through its own acceleration structure as long as the answer stays the reference's: every
// candidate still goes through the exact triangle test, and a triangle that accepts is only believed after the exact slab
// tests of its ancestors in the cyBVH (the reference reaches a triangle only through them, objFunctions.cpp:346-395; see
// ref_reaches() in csrc/intersect.cuh).  That leaves the shape of the search structure open, and the cyBVH - midpoint splits
// of the widest axis, cyBVH.h:295-328 - is a poor one for rays that skim a mesh.  This file builds a binned-SAH binary
// hierarchy over the same triangles (Wald 2007: 16 bins per axis, cost = 1 + (A_l N_l + A_r N_r) / A), leaves of up to 4
// triangles, and collapses it into 4-wide nodes (device_scene.h OccNode: the boxes of four children as six float4 + four
// child words = one 128-byte line per visit).  The any-hit kernel is bound by the latency of one pool iteration (shared
// memory pop -> node fetch -> tests -> push), not by its arithmetic: a 4-wide node halves the iterations per ray.
//
// Boxes are the exact float bounds of the triangles' vertices; the conservative margin that covers the rounding of a
// triangle's computed hit point is applied per ray on the device (occ_setup in csrc/intersect.cuh).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <vector>

#include "host_scene.h"

namespace rtu {

namespace {

struct Box3 {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = 3.0e38f; hi[k] = -3.0e38f; } }
    void grow(const float *p) { for (int k = 0; k < 3; k++) { if (p[k] < lo[k]) lo[k] = p[k]; if (p[k] > hi[k]) hi[k] = p[k]; } }
    void grow(const Box3 &b) { for (int k = 0; k < 3; k++) { if (b.lo[k] < lo[k]) lo[k] = b.lo[k]; if (b.hi[k] > hi[k]) hi[k] = b.hi[k]; } }
    double half_area() const
    {
        double d[3] = {(double)hi[0] - lo[0], (double)hi[1] - lo[1], (double)hi[2] - lo[2]};
        if (d[0] < 0 || d[1] < 0 || d[2] < 0) return 0.0;
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
    }
};

struct Prim {
    Box3 box;
    float c[3];      // centroid of the box
    uint32_t slot;   // cyBVH leaf-order slot of the triangle
};

struct TempNode {
    Box3 box;
    int left = -1, right = -1; // children (internal) ...
    uint32_t first = 0, count = 0; // ... or the leaf's range of `order`
};

struct Builder {
    std::vector<Prim> prims;
    std::vector<TempNode> nodes;
    static constexpr int BINS = 16;
    uint32_t MAX_LEAF = 4;   // (RTU_OCC_MAX_LEAF / RTU_OCC_TRAV_COST: tuning experiments, profiles/README.md)
    double TRAV = 1.0;
    Builder()
    {
        if (const char *e = getenv("RTU_OCC_MAX_LEAF")) { int v = atoi(e); if (v >= 1 && v <= 8) MAX_LEAF = (uint32_t)v; }
        if (const char *e = getenv("RTU_OCC_TRAV_COST")) { double v = atof(e); if (v > 0) TRAV = v; }
    }

    int build(uint32_t b, uint32_t e)
    {
        const int me = (int)nodes.size();
        nodes.push_back(TempNode());
        Box3 box, cbox;
        box.reset();
        cbox.reset();
        for (uint32_t i = b; i < e; i++) { box.grow(prims[i].box); cbox.grow(prims[i].c); }
        nodes[me].box = box;
        const uint32_t n = e - b;
        auto make_leaf = [&]() { nodes[me].first = b; nodes[me].count = n; return me; };
        if (n == 1) return make_leaf();
        // best binned split over the three axes
        double best_cost = 1e300;
        int best_axis = -1, best_bin = -1;
        const double area = box.half_area();
        for (int axis = 0; axis < 3; axis++) {
            const float lo = cbox.lo[axis], ext = cbox.hi[axis] - cbox.lo[axis];
            if (!(ext > 0.f)) continue;
            Box3 bb[BINS];
            uint32_t cnt[BINS];
            for (int k = 0; k < BINS; k++) { bb[k].reset(); cnt[k] = 0; }
            const float scale = (float)BINS / ext;
            for (uint32_t i = b; i < e; i++) {
                int k = (int)((prims[i].c[axis] - lo) * scale);
                k = k < 0 ? 0 : (k >= BINS ? BINS - 1 : k);
                bb[k].grow(prims[i].box);
                cnt[k]++;
            }
            double right_area[BINS];
            uint32_t right_cnt[BINS];
            Box3 acc;
            acc.reset();
            uint32_t c = 0;
            for (int k = BINS - 1; k > 0; k--) { acc.grow(bb[k]); c += cnt[k]; right_area[k] = acc.half_area(); right_cnt[k] = c; }
            acc.reset();
            c = 0;
            for (int k = 0; k < BINS - 1; k++) {
                acc.grow(bb[k]);
                c += cnt[k];
                if (c == 0 || right_cnt[k + 1] == 0) continue;
                const double cost = TRAV + (acc.half_area() * c + right_area[k + 1] * right_cnt[k + 1]) / (area > 0 ? area : 1.0);
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = k; }
            }
        }
        if (n <= MAX_LEAF && (best_axis < 0 || (double)n <= best_cost)) return make_leaf();
        uint32_t mid;
        if (best_axis < 0) {
            mid = b + n / 2; // all centroids coincide: halve by index
        } else {
            const float lo = cbox.lo[best_axis], scale = (float)BINS / (cbox.hi[best_axis] - cbox.lo[best_axis]);
            auto it = std::partition(prims.begin() + b, prims.begin() + e, [&](const Prim &p) {
                int k = (int)((p.c[best_axis] - lo) * scale);
                k = k < 0 ? 0 : (k >= BINS ? BINS - 1 : k);
                return k <= best_bin;
            });
            mid = (uint32_t)(it - prims.begin());
            if (mid == b || mid == e) mid = b + n / 2;
        }
        const int l = build(b, mid);
        const int r = build(mid, e);
        nodes[me].left = l;
        nodes[me].right = r;
        return me;
    }
};

} // namespace

void build_occlusion_bvh(const float *v, const uint32_t *f, const uint32_t *elements, uint32_t nf, OccBvh *out)
{
    out->nodes.clear();
    out->slots.clear();
    out->root = 0;
    if (nf == 0) return;
    Builder B;
    B.prims.resize(nf);
    for (uint32_t s = 0; s < nf; s++) {
        const uint32_t face = elements[s];
        Prim &p = B.prims[s];
        p.box.reset();
        for (int k = 0; k < 3; k++) p.box.grow(v + (size_t)f[(size_t)face * 3 + k] * 3);
        for (int k = 0; k < 3; k++) p.c[k] = 0.5f * p.box.lo[k] + 0.5f * p.box.hi[k];
        p.slot = s;
    }
    B.nodes.reserve((size_t)nf * 2);
    B.build(0, nf);
    out->slots.resize(nf);
    for (uint32_t i = 0; i < nf; i++) out->slots[i] = B.prims[i].slot;
    const uint32_t NONE = 0x7fffffffu;
    auto leaf_word = [&](const TempNode &t) { return 0x80000000u | ((t.count - 1u) << 28) | t.first; };
    if (B.nodes[0].left < 0) { out->root = leaf_word(B.nodes[0]); return; }
    // Collapse the binary tree into 4-wide nodes: a wide node starts with the two children of a binary node and keeps
    // replacing its largest internal child by that child's two children until it has four (or only leaves are left).
    // Breadth-first numbering: the top levels come first and stay cache resident.
    struct Wide { int child[4]; int n; };
    std::vector<Wide> wide;
    std::vector<int> wide_of_binary; // binary node index -> wide node index (for children that stay internal)
    wide_of_binary.assign(B.nodes.size(), -1);
    std::vector<int> queue;          // binary nodes that become wide nodes, in BFS order
    queue.push_back(0);
    wide_of_binary[0] = 0;
    for (size_t h = 0; h < queue.size(); h++) {
        const TempNode &t = B.nodes[queue[h]];
        Wide w;
        w.child[0] = t.left; w.child[1] = t.right; w.n = 2;
        while (w.n < 4) {
            int pick = -1;
            double best = -1.0;
            for (int k = 0; k < w.n; k++) {
                const TempNode &c = B.nodes[w.child[k]];
                if (c.left < 0) continue;
                const double a = c.box.half_area();
                if (a > best) { best = a; pick = k; }
            }
            if (pick < 0) break;
            const TempNode &c = B.nodes[w.child[pick]];
            w.child[pick] = c.left;
            w.child[w.n++] = c.right;
        }
        for (int k = 0; k < w.n; k++)
            if (B.nodes[w.child[k]].left >= 0) {
                wide_of_binary[w.child[k]] = (int)queue.size();
                queue.push_back(w.child[k]);
            }
        wide.push_back(w);
    }
    // device layout (device_scene.h OccNode, 32 words): lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] child[4] pad[4];
    // an unused child slot has an inverted box (never hit) and the word NONE
    out->nodes.assign(wide.size() * 32, 0.f);
    for (size_t i = 0; i < wide.size(); i++) {
        float *N = &out->nodes[i * 32];
        uint32_t words[8] = {NONE, NONE, NONE, NONE, 0u, 0u, 0u, 0u};
        for (int k = 0; k < 4; k++) {
            if (k < wide[i].n) {
                const TempNode &c = B.nodes[wide[i].child[k]];
                for (int a = 0; a < 3; a++) { N[a * 4 + k] = c.box.lo[a]; N[12 + a * 4 + k] = c.box.hi[a]; }
                words[k] = c.left < 0 ? leaf_word(c) : (uint32_t)wide_of_binary[wide[i].child[k]];
            } else {
                for (int a = 0; a < 3; a++) { N[a * 4 + k] = 3.0e38f; N[12 + a * 4 + k] = -3.0e38f; }
            }
        }
        memcpy(N + 24, words, sizeof words);
    }
    out->root = 0; // wide node 0
}

} // namespace rtu
