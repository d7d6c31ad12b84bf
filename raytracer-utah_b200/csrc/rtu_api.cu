// C ABI (include/rtu.h) device side: context, scene packing + upload, wave orchestration.
// There is no CPU fallback: every entry point that renders or traces needs a CUDA device and
// returns RTU_ERR_NO_DEVICE / RTU_ERR_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../host/hmath.h"
#include "../host/host_scene.h"
#include "rtu_internal.h"
#include "rtu_objects.h"

using rtu::V3;

namespace {

thread_local size_t g_upload_bytes = 0; // host->device bytes of the scene upload in progress
std::atomic<uint64_t> g_scene_serial{1}; // identifies an uploaded scene in the per-context frame-setup cache

template <class T> int dev_upload(const std::vector<T> &h, T **d, cudaStream_t st, std::vector<void *> &owned)
{
    *d = nullptr;
    if (h.empty()) return RTU_OK;
    g_upload_bytes += h.size() * sizeof(T);
    CU(cudaMallocAsync((void **)d, h.size() * sizeof(T), st)); // scene buffers come from the device's stream-ordered pool
    owned.push_back(*d);
    CU(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return RTU_OK;
}

} // namespace

namespace {

void kt_begin(rtu_context *c, int cls)
{
    c->cls_launches[cls]++;
    if (!c->kt_on) return;
    if (c->kt_used * 2 + 2 > c->kt_ev.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        c->kt_ev.push_back(a);
        c->kt_ev.push_back(b);
        c->kt_cls.push_back(cls);
    }
    c->kt_cls[c->kt_used] = cls;
    cudaEventRecord(c->kt_ev[c->kt_used * 2], c->stream);
}
void kt_end(rtu_context *c)
{
    if (!c->kt_on) return;
    cudaEventRecord(c->kt_ev[c->kt_used * 2 + 1], c->stream);
    c->kt_used++;
}
void kt_reset(rtu_context *c, bool on)
{
    c->kt_on = on;
    c->kt_used = 0;
    c->cls_launches[0] = c->cls_launches[1] = c->cls_launches[2] = c->cls_launches[3] = 0;
}

int free_list(std::vector<void *> &v)
{
    for (void *p : v) cudaFree(p);
    v.clear();
    return 0;
}

// scene buffers: freed in stream order (no device-wide synchronisation; the pool keeps the memory for the next scene)
int free_list_async(std::vector<void *> &v, cudaStream_t st)
{
    for (void *p : v) cudaFreeAsync(p, st);
    v.clear();
    return 0;
}

int ensure_scratch(rtu_context *c, size_t q_cap, size_t shadow_cap)
{
    if (q_cap <= c->q_cap && shadow_cap <= c->shadow_cap) return RTU_OK;
    CU(cudaStreamSynchronize(c->stream));
    free_list(c->scratch);
    c->q_cap = c->shadow_cap = 0;
    memset(&c->wb, 0, sizeof c->wb); // nothing may point into the freed block if an allocation below fails
    auto alloc = [&](void **p, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) c->scratch.push_back(*p);
        else if (e == cudaErrorMemoryAllocation) { c->scratch_oom = true; cudaGetLastError(); free_list(c->scratch); }
        return e;
    };
    c->scratch_oom = false;
    unsigned *counts = nullptr;
    CU(alloc((void **)&counts, 8 * sizeof(unsigned)));
    CU(cudaMemsetAsync(counts, 0, 8 * sizeof(unsigned), c->stream));
    for (int i = 0; i < 2; i++) {
        CU(alloc((void **)&c->wb.q[i].o, q_cap * sizeof(float4)));
        CU(alloc((void **)&c->wb.q[i].d, q_cap * sizeof(float4)));
        CU(alloc((void **)&c->wb.q[i].w, q_cap * sizeof(float4)));
        CU(alloc((void **)&c->wb.q[i].path, q_cap * sizeof(uint32_t)));
        c->wb.q[i].count = counts + i;
        c->wb.q[i].cap = (uint32_t)q_cap;
        CU(alloc((void **)&c->wb.aux[i].a, q_cap * sizeof(float4)));
        CU(alloc((void **)&c->wb.aux[i].b, q_cap * sizeof(float4)));
        c->wb.aux[i].count = counts + 2 + i;
        c->wb.aux[i].cap = (uint32_t)q_cap;
    }
    CU(alloc((void **)&c->wb.shadow.o, shadow_cap * sizeof(float4)));
    CU(alloc((void **)&c->wb.shadow.d, shadow_cap * sizeof(float4)));
    CU(alloc((void **)&c->wb.shadow.c, shadow_cap * sizeof(float4)));
    c->wb.shadow.count = counts + 4;
    c->wb.shadow.cap = (uint32_t)shadow_cap;
    CU(alloc((void **)&c->wb.hits.a, q_cap * sizeof(float4)));
    CU(alloc((void **)&c->wb.hits.b, q_cap * sizeof(float4)));
    c->wb.hits.count = counts + 5;
    c->wb.hits.cap = (uint32_t)q_cap;
    c->wb.gi_count = counts + 6;
    CU(alloc((void **)&c->wb.park, (size_t)XP_MAX_WARPS * (XP_JOBS + XP_RES) * 3 * sizeof(float4)));
    CU(alloc((void **)&c->wb.counters, sizeof(DCounters)));
    CU(cudaMemsetAsync(c->wb.counters, 0, sizeof(DCounters), c->stream));
    c->q_cap = q_cap;
    c->shadow_cap = shadow_cap;
    return RTU_OK;
}

int ensure_gi(rtu_context *c, size_t n)
{
    if (n <= c->gi_n) return RTU_OK;
    CU(cudaStreamSynchronize(c->stream));
    if (c->gi) cudaFree(c->gi);
    c->gi = nullptr;
    c->gi_n = 0;
    CU(cudaMalloc((void **)&c->gi, n * sizeof(float4)));
    c->gi_n = n;
    return RTU_OK;
}

int ensure_work(rtu_context *c, size_t n)
{
    if (n > c->work_n) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->work) cudaFree(c->work);
        c->work = nullptr;
        CU(cudaMalloc((void **)&c->work, n * sizeof(unsigned)));
        c->work_n = n;
    }
    CU(cudaMemsetAsync(c->work, 0, c->work_n * sizeof(unsigned), c->stream));
    return RTU_OK;
}

// ---- camera frame: CalculateImageOrigin / CalculateCurrentPoint (RenderFunctions.cpp:243-269)
void make_camera(const rtu_camera &c, int W, int H, DCamera *out)
{
    V3 pos(c.pos[0], c.pos[1], c.pos[2]), dir(c.dir[0], c.dir[1], c.dir[2]), up(c.up[0], c.up[1], c.up[2]);
    float d = c.focaldist;
    float actualHeight = (float)(tan((c.fov / 2) * M_PI / 180.0) * 2 * d);
    float actualWidth = ((float)W / (float)H) * actualHeight;
    V3 nd = rtu::normalized(dir), nu = rtu::normalized(up);
    V3 right = rtu::normalized(rtu::cross(nd, nu));
    V3 topCenter = pos + d * nd + (actualHeight / 2) * nu;
    V3 origin = topCenter - (actualWidth / 2) * right;
    V3 u = right * (actualWidth / (float)W);
    V3 v = (-1.0f * nu) * (actualHeight / (float)H);
    for (int k = 0; k < 3; k++) {
        out->pos[k] = pos[k]; out->origin[k] = origin[k]; out->u[k] = u[k]; out->v[k] = v[k];
        out->lens_x[k] = right[k]; out->lens_y[k] = up[k];
    }
    out->dof = c.dof;
    out->width = W;
    out->height = H;
    out->inv_w = 1.0f / W;
    out->inv_h = 1.0f / H;
}

float halton(int index, int base) // scene.h:130-139
{
    float r = 0;
    float f = 1.0f / (float)base;
    for (int i = index; i > 0; i /= base) {
        r += f * (i % base);
        f /= (float)base;
    }
    return r;
}

// ---- mesh packing
// The ray-independent part of IntersectTriangle (objFunctions.cpp:259-300) and the winner's attributes (:317-320) of one
// face, evaluated with the reference's float operations; `slot`: where the record goes (cyBVH leaf order, or face order)
int triangle_records(const rtu_mesh &m, uint32_t face, TriRec &T, TriShade &Sh)
{
    auto P3 = [&](const float *a, uint32_t i) { return V3(a[(size_t)i * 3], a[(size_t)i * 3 + 1], a[(size_t)i * 3 + 2]); };
    uint32_t i0 = m.f[face * 3], i1 = m.f[face * 3 + 1], i2 = m.f[face * 3 + 2];
    if (i0 >= m.nv || i1 >= m.nv || i2 >= m.nv) { rtu::set_error("rtu_scene_upload: vertex index out of range"); return RTU_ERR_INVALID; }
    V3 A = P3(m.v, i0), B = P3(m.v, i1), C = P3(m.v, i2);
    V3 N = rtu::normalized(rtu::cross(B - A, C - A));                       // objFunctions.cpp:263
    float ax = fabsf(N.x), ay = fabsf(N.y), az = fabsf(N.z);
    float mxy = (ax < ay) ? ay : ax;                                        // std::max (:274)
    float mx = (mxy < az) ? az : mxy;
    unsigned axis = mx == ax ? 0u : (mx == ay ? 1u : (mx == az ? 2u : 3u)); // :278-295
    float Au, Av, Bu, Bv, Cu, Cv;
    if (axis == 0) { Au = A.y; Av = A.z; Bu = B.y; Bv = B.z; Cu = C.y; Cv = C.z; }
    else if (axis == 1) { Au = A.x; Av = A.z; Bu = B.x; Bv = B.z; Cu = C.x; Cv = C.z; }
    else { Au = A.x; Av = A.y; Bu = B.x; Bv = B.y; Cu = C.x; Cv = C.y; }
    T.nx = N.x; T.ny = N.y; T.nz = N.z;
    T.ax = A.x; T.ay = A.y; T.az = A.z;
    T.cau = Cu - Au; T.cav = Cv - Av; T.bau = Bu - Au; T.bav = Bv - Av;
    float cr = (-T.cav) * T.bau + T.cau * T.bav;                            // Point2::Cross (cyPoint.h:248)
    T.area = (float)((double)cr / 2.0);                                     // :298
    uint32_t fb = (face & 0x3fffffffu) | (axis << 30);
    memcpy(&T.fbits, &fb, 4);
    memset(&Sh, 0, sizeof Sh);
    for (int k = 0; k < 3; k++) {
        uint32_t vi = m.f[face * 3 + k];
        uint32_t ni = m.fn[face * 3 + k];
        if (ni >= m.nvn) { rtu::set_error("rtu_scene_upload: normal index out of range"); return RTU_ERR_INVALID; }
        for (int c = 0; c < 3; c++) { Sh.v[k * 3 + c] = m.v[(size_t)vi * 3 + c]; Sh.vn[k * 3 + c] = m.vn[(size_t)ni * 3 + c]; }
        if (m.ft && m.vt) {
            uint32_t ti = m.ft[face * 3 + k];
            if (ti >= m.nvt) { rtu::set_error("rtu_scene_upload: texture index out of range"); return RTU_ERR_INVALID; }
            for (int c = 0; c < 3; c++) Sh.vt[k * 3 + c] = m.vt[(size_t)ti * 3 + c];
        }
    }
    return RTU_OK;
}

bool coords_fine(const float *bmin, const float *bmax, const float *extra, size_t n_extra)
{
    auto fine = [](float v) { return v == 0.f || (std::fabs(v) >= 1.4551915228366852e-11f && std::isfinite(v)); };
    bool ok = true;
    for (size_t i = 0; i < n_extra && ok; i++) ok = fine(extra[i]);
    for (int k = 0; k < 3; k++) ok = ok && fine(bmin[k]) && fine(bmax[k]);
    return ok;
}

// RTU_MESH_DEVICE_BVH: triangle records in FACE order (a triangle's slot is its face index), the hierarchy built on the device
int pack_mesh_device_bvh(const rtu_mesh &m, DMesh *out, cudaStream_t st, std::vector<void *> &owned, double *build_ms)
{
    if (m.nf == 0) { out->empty = 1; return RTU_OK; }
    if (!m.v || !m.f || !m.vn || !m.fn) { rtu::set_error("rtu_scene_upload: mesh with missing arrays"); return RTU_ERR_INVALID; }
    if (m.nf > (1u << 24)) { rtu::set_error("rtu_scene_upload: RTU_MESH_DEVICE_BVH handles up to 2^24 triangles per mesh"); return RTU_ERR_UNSUPPORTED; }
    std::vector<TriRec> tris(m.nf);
    std::vector<TriShade> shade(m.nf);
    int rc;
    for (uint32_t face = 0; face < m.nf; face++)
        if ((rc = triangle_records(m, face, tris[face], shade[face]))) return rc;
    TriRec *dt = nullptr;
    TriShade *ds = nullptr;
    if ((rc = dev_upload(tris, &dt, st, owned))) return rc;
    if ((rc = dev_upload(shade, &ds, st, owned))) return rc;
    // vertices and faces go up for the builder only
    float *dv = nullptr;
    uint32_t *df = nullptr;
    CU(cudaMallocAsync((void **)&dv, sizeof(float) * 3 * (size_t)m.nv, st));
    cudaError_t e = cudaMallocAsync((void **)&df, sizeof(uint32_t) * 3 * (size_t)m.nf, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv, m.v, sizeof(float) * 3 * (size_t)m.nv, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(df, m.f, sizeof(uint32_t) * 3 * (size_t)m.nf, cudaMemcpyHostToDevice, st);
    g_upload_bytes += sizeof(float) * 3 * (size_t)m.nv + sizeof(uint32_t) * 3 * (size_t)m.nf;
    float ms = 0.f;
    rc = RTU_OK;
    if (e == cudaSuccess) rc = rtu_lbvh_build(st, dv, df, m.nf, m.bound_min, m.bound_max, dt, owned, &out->occ_nodes, &out->occ_tris, &out->occ_root, &ms);
    cudaFreeAsync(dv, st);
    if (df) cudaFreeAsync(df, st);
    CU(e);
    if (rc) return rc;
    CU(cudaStreamSynchronize(st)); // host vectors go out of scope
    if (build_ms) *build_ms += ms;
    out->pairs = nullptr;
    out->tris = dt;
    out->shade = ds;
    out->tri_up = nullptr;
    out->root = 0x7fffffffu;
    out->n_pairs = 0;
    out->n_tris = m.nf;
    out->no_ref = 1u;
    out->nested = 0u;
    out->coords_ok = coords_fine(out->bmin, out->bmax, nullptr, 0) ? 1u : 0u;
    float sc = 0.f;
    for (int k = 0; k < 3; k++) sc = std::max(sc, std::max(std::fabs(m.bound_min[k]), std::fabs(m.bound_max[k])));
    out->occ_scale = sc;
    return RTU_OK;
}

int pack_mesh(const rtu_mesh &m, DMesh *out, cudaStream_t st, std::vector<void *> &owned, double *build_ms)
{
    memset(out, 0, sizeof *out);
    for (int k = 0; k < 3; k++) { out->bmin[k] = m.bound_min[k]; out->bmax[k] = m.bound_max[k]; }
    if (m.flags & RTU_MESH_DEVICE_BVH) return pack_mesh_device_bvh(m, out, st, owned, build_ms);
    if (m.nf == 0 || m.bvh_nodes < 2) { out->empty = 1; return RTU_OK; }
    if (!m.v || !m.f || !m.vn || !m.fn || !m.bvh_boxes || !m.bvh_data || !m.bvh_elements) {
        rtu::set_error("rtu_scene_upload: mesh with missing arrays");
        return RTU_ERR_INVALID;
    }
    const uint32_t LEAF = 0x80000000u;
    // BFS numbering of internal nodes -> pair records; top levels end up first (smem staging)
    std::vector<uint32_t> pair_of(m.bvh_nodes, 0xffffffffu), order;
    auto check_node = [&](uint32_t n) { return n >= 1 && n < m.bvh_nodes; };
    if (!(m.bvh_data[1] & LEAF)) { pair_of[1] = 0; order.push_back(1); }
    for (size_t h = 0; h < order.size(); h++) {
        uint32_t n = order[h];
        uint32_t c1 = m.bvh_data[n] & 0x7fffffffu;
        if (!check_node(c1) || !check_node(c1 + 1)) { rtu::set_error("rtu_scene_upload: BVH child index out of range"); return RTU_ERR_INVALID; }
        for (uint32_t c = c1; c <= c1 + 1; c++)
            if (!(m.bvh_data[c] & LEAF)) {
                if (pair_of[c] != 0xffffffffu) { rtu::set_error("rtu_scene_upload: BVH is not a tree"); return RTU_ERR_INVALID; }
                pair_of[c] = (uint32_t)order.size();
                order.push_back(c);
            }
    }
    auto child_word = [&](uint32_t c, uint32_t *w) -> bool {
        uint32_t dw = m.bvh_data[c];
        if (dw & LEAF) {
            uint32_t off = dw & 0x0fffffffu, cnt = ((dw >> 28) & 7u) + 1u;
            if ((uint64_t)off + cnt > m.nf) return false;
            *w = dw;
        } else {
            *w = pair_of[c];
        }
        return true;
    };
    std::vector<BvhPair> pairs(order.size());
    for (size_t i = 0; i < order.size(); i++) {
        uint32_t c1 = m.bvh_data[order[i]] & 0x7fffffffu;
        BvhPair &P = pairs[i];
        memcpy(P.b1, m.bvh_boxes + (size_t)c1 * 6, 6 * sizeof(float));
        memcpy(P.b2, m.bvh_boxes + (size_t)(c1 + 1) * 6, 6 * sizeof(float));
        if (!child_word(c1, &P.c1) || !child_word(c1 + 1, &P.c2)) { rtu::set_error("rtu_scene_upload: BVH leaf range out of bounds"); return RTU_ERR_INVALID; }
        P.up = 0xffffffffu;
        P.pad = 0;
    }
    // ancestor links: pair i holds the boxes of the two children of internal node order[i]; a child that is internal has its
    // own pair, whose `up` points back here; a child that is a leaf hands the link to its triangles (tri_up)
    std::vector<uint32_t> tri_up(m.nf, 0xffffffffu);
    for (size_t i = 0; i < order.size(); i++) {
        const uint32_t c1 = m.bvh_data[order[i]] & 0x7fffffffu;
        for (uint32_t k = 0; k < 2; k++) {
            const uint32_t c = c1 + k, link = (uint32_t)i | (k << 31);
            const uint32_t dw = m.bvh_data[c];
            if (dw & LEAF) {
                const uint32_t off = dw & 0x0fffffffu, cnt = ((dw >> 28) & 7u) + 1u;
                for (uint32_t t = 0; t < cnt; t++) tri_up[off + t] = link;
            } else {
                pairs[pair_of[c]].up = link;
            }
        }
    }
    if (!child_word(1, &out->root)) { rtu::set_error("rtu_scene_upload: BVH root leaf out of bounds"); return RTU_ERR_INVALID; }
    // depth bound for the traversal stack (depth-first, both children pushed: depth+1 entries)
    {
        std::vector<std::pair<uint32_t, int>> st2;
        st2.push_back({1u, 1});
        int maxd = 1;
        while (!st2.empty()) {
            auto [n, dpt] = st2.back();
            st2.pop_back();
            maxd = std::max(maxd, dpt);
            if (!(m.bvh_data[n] & LEAF)) {
                uint32_t c1 = m.bvh_data[n] & 0x7fffffffu;
                st2.push_back({c1, dpt + 1});
                st2.push_back({c1 + 1, dpt + 1});
            }
        }
        if (maxd + 2 > RTU_STACK) { rtu::set_error("rtu_scene_upload: BVH deeper than the traversal stack"); return RTU_ERR_UNSUPPORTED; }
    }
    std::vector<TriRec> tris(m.nf);
    std::vector<TriShade> shade(m.nf);
    for (uint32_t slot = 0; slot < m.nf; slot++) {
        uint32_t face = m.bvh_elements[slot];
        if (face >= m.nf) { rtu::set_error("rtu_scene_upload: BVH element out of range"); return RTU_ERR_INVALID; }
        int trc = triangle_records(m, face, tris[slot], shade[slot]);
        if (trc) return trc;
    }
    // does every box contain its children's boxes?  cyBVH::Build's do (a node's box is the min / max over its elements)
    bool nested = true;
    for (size_t i = 0; i < pairs.size() && nested; i++) {
        const BvhPair &P = pairs[i];
        if (P.up == 0xffffffffu) continue; // the root's own box is never tested (objFunctions.cpp:343)
        const BvhPair &U = pairs[P.up & 0x7fffffffu];
        const float *outer = (P.up >> 31) ? U.b2 : U.b1;
        for (const float *inner : {P.b1, P.b2})
            for (int k = 0; k < 3 && nested; k++) nested = inner[k] >= outer[k] && inner[3 + k] <= outer[3 + k]; // (false for NaN)
    }
    // any-hit hierarchy: the caller's (built once at load time by rtu_host_load_xml / rtu_host_build_occlusion_bvh), else built here
    rtu::OccBvh built;
    const float *onodes = m.occ_nodes;
    const uint32_t *oslots = m.occ_slots;
    uint32_t on_nodes = m.occ_n_nodes, oroot = m.occ_root;
    if (!oslots) {
        rtu::build_occlusion_bvh(m.v, m.f, m.bvh_elements, m.nf, &built);
        onodes = built.nodes.data();
        oslots = built.slots.data();
        on_nodes = (uint32_t)(built.nodes.size() / 32);
        oroot = built.root;
    }
    static_assert(sizeof(OccNode) == 128, "any-hit node layout");
    std::vector<OccNode> occ(on_nodes);
    if (on_nodes) {
        if (!onodes) { rtu::set_error("rtu_scene_upload: occ_n_nodes without occ_nodes"); return RTU_ERR_INVALID; }
        memcpy(occ.data(), onodes, (size_t)on_nodes * sizeof(OccNode));
    }
    {   // validate what the device will index with: child words of the hierarchy and the slot permutation
        auto word_ok = [&](uint32_t w) {
            if (w == 0x7fffffffu) return true; // unused child slot
            if (w & LEAF) return (uint64_t)(w & 0x0fffffffu) + (((w >> 28) & 7u) + 1u) <= m.nf;
            return w < on_nodes;
        };
        bool ok = word_ok(oroot) && oroot != 0x7fffffffu && on_nodes < (1u << 27);
        for (uint32_t i = 0; i < on_nodes && ok; i++)
            for (int k = 0; k < 4; k++) ok = ok && word_ok(occ[i].child[k]);
        std::vector<unsigned char> seen(m.nf, 0);
        for (uint32_t i = 0; i < m.nf && ok; i++) { ok = oslots[i] < m.nf && !seen[oslots[i]]; if (ok) seen[oslots[i]] = 1; }
        if (!ok) { rtu::set_error("rtu_scene_upload: malformed any-hit hierarchy (occ_nodes / occ_slots)"); return RTU_ERR_INVALID; }
    }
    std::vector<TriRec> otris(m.nf);
    for (uint32_t i = 0; i < m.nf; i++) {
        otris[i] = tris[oslots[i]];
        uint32_t fb;
        memcpy(&fb, &otris[i].fbits, 4);
        fb = (fb & 0xc0000000u) | oslots[i]; // keep the projection axis, carry the cyBVH slot instead of the face id
        memcpy(&otris[i].fbits, &fb, 4);
    }
    BvhPair *dp = nullptr;
    OccNode *dop = nullptr;
    TriRec *dt = nullptr, *dot = nullptr;
    TriShade *ds = nullptr;
    uint32_t *dup = nullptr;
    int rc;
    if ((rc = dev_upload(pairs, &dp, st, owned))) return rc;
    if ((rc = dev_upload(tris, &dt, st, owned))) return rc;
    if ((rc = dev_upload(shade, &ds, st, owned))) return rc;
    if ((rc = dev_upload(occ, &dop, st, owned))) return rc;
    if ((rc = dev_upload(otris, &dot, st, owned))) return rc;
    if ((rc = dev_upload(tri_up, &dup, st, owned))) return rc;
    CU(cudaStreamSynchronize(st)); // host vectors go out of scope
    out->pairs = dp;
    out->tris = dt;
    out->shade = ds;
    out->occ_nodes = dop;
    out->occ_tris = dot;
    out->tri_up = dup;
    out->occ_root = oroot;
    out->nested = nested ? 1u : 0u;
    {
        float sc = 0.f;
        for (int k = 0; k < 3; k++) sc = std::max(sc, std::max(std::fabs(m.bound_min[k]), std::fabs(m.bound_max[k])));
        out->occ_scale = sc;
    }
    out->n_pairs = (uint32_t)pairs.size();
    out->n_tris = m.nf;
    {   // box coordinates that are 0 or >= 2^-36 in magnitude: then "bound - origin" is 0 or >= 2^-60 for every such
        // origin, which is what the hoisted division needs (intersect.cuh, mesh_invdir)
        auto fine = [](float v) { return v == 0.f || (std::fabs(v) >= 1.4551915228366852e-11f && std::isfinite(v)); };
        bool ok = true;
        const float *pf = reinterpret_cast<const float *>(pairs.data());
        for (size_t i = 0; i < pairs.size() && ok; i++)
            for (int k = 0; k < 12; k++) ok = ok && fine(pf[i * (sizeof(BvhPair) / 4) + k]);
        for (int k = 0; k < 3; k++) ok = ok && fine(out->bmin[k]) && fine(out->bmax[k]);
        out->coords_ok = ok ? 1u : 0u;
    }
    return RTU_OK;
}

DTexColor pack_tc(const rtu_texcolor &t, int n_texmaps)
{
    DTexColor d;
    d.c[0] = t.color[0]; d.c[1] = t.color[1]; d.c[2] = t.color[2];
    d.map = (t.texmap >= 0 && t.texmap < n_texmaps) ? t.texmap : -1;
    return d;
}

} // namespace

extern "C" {

void rtu_params_default(rtu_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->spp = 1;
    p->pattern = RTU_PATTERN_CENTER;
    p->mode = RTU_MODE_WHITTED;
    p->shade_bounces = 5; // RenderFunctions.cpp:134
    p->gi_bounces = 4;    // RenderFunctions.cpp:31
}

static int context_create_impl(int32_t device, void *stream, rtu_context **out)
{
    if (!out) { rtu::set_error("rtu_context_create: null out"); return RTU_ERR_INVALID; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        rtu::set_error(std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
        return RTU_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { rtu::set_error("rtu_context_create: bad device index"); return RTU_ERR_INVALID; }
    CU(cudaSetDevice(device));
    {   // scene buffers are stream-ordered allocations: keep freed blocks in the pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    rtu_context *c = new rtu_context;
    c->device = device;
    c->stream = (cudaStream_t)stream;
    int sm_count = 0;
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sm_count < 1) sm_count = 148;
    c->cfg.sm_count = sm_count;
    c->cfg.blocks_per_sm = 2;
    c->cfg.threads = 256;
    {   // 2^28 primary rays per chunk (80 GB of queues for two shadow lights) where the device is a 180 GB part with most of it free
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b >= (size_t)140 << 30) c->chunk_rays = (size_t)1 << 28;
    }
    if (const char *s = getenv("RTU_CHUNK_RAYS")) { long long v = atoll(s); if (v >= 1024) { c->chunk_rays = (size_t)v; c->chunk_rays_set = true; } }
    if (const char *s = getenv("RTU_QUEUE_FACTOR")) { double v = atof(s); if (v >= 1.0 && v <= 8.0) c->queue_factor = v; }
    if (const char *s = getenv("RTU_BLOCKS_PER_SM")) { int v = atoi(s); if (v >= 1 && v <= 8) c->cfg.blocks_per_sm = v; }
    memset(&c->wb, 0, sizeof c->wb);
    auto bail = [&](cudaError_t e, const char *what) {
        rtu::set_error(std::string(what) + ": " + cudaGetErrorString(e));
        rtu_context_destroy(c);
        return RTU_ERR_CUDA;
    };
    cudaError_t e2;
    if ((e2 = cudaMalloc((void **)&c->zmm, (4 + 2 * rtu_context::WAVE_LOG_MAX) * sizeof(unsigned))) != cudaSuccess) return bail(e2, "cudaMalloc");
    c->d_wave_log = c->zmm + 4;
    if ((e2 = cudaEventCreateWithFlags(&c->wave_ev, cudaEventDisableTiming)) != cudaSuccess) return bail(e2, "cudaEventCreate");
    if (const char *s = getenv("RTU_TAIL_RAYS")) { long long v = atoll(s); if (v >= 0) c->tail_rays = (size_t)v; }
    if ((e2 = cudaEventCreate(&c->ev0)) != cudaSuccess) return bail(e2, "cudaEventCreate");
    if ((e2 = cudaEventCreate(&c->ev1)) != cudaSuccess) return bail(e2, "cudaEventCreate");
    if ((e2 = cudaHostAlloc((void **)&c->h_flag, 64 + rtu_context::WAVE_LOG_MAX * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess) return bail(e2, "cudaHostAlloc");
    c->h_wave_log = c->h_flag + 16;
    *out = c;
    return RTU_OK;
}

void rtu_context_destroy(rtu_context *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_list(c->scratch);
    if (c->work) cudaFree(c->work);
    if (c->zmm) cudaFree(c->zmm);
    if (c->gi) cudaFree(c->gi);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->wave_ev) cudaEventDestroy(c->wave_ev);
    for (cudaEvent_t e : c->kt_ev) cudaEventDestroy(e);
    void *fbp[] = {c->fb.accum, c->fb.accum2, c->fb.d_rgb, c->fb.d_rgb8, c->fb.d_z, c->fb.d_z8, c->fb.d_node, c->fb.d_face, c->fb.d_offsets};
    for (void *p : fbp) if (p) cudaFree(p);
    if (c->fb.d_tile) cudaFree(c->fb.d_tile);
    c->stage_off.release();
    c->stage_tile.release();
    c->stage_photon_raw.release();
    c->stage_photon_bal.release();
    for (void *q : {(void *)c->d_tile_done, (void *)c->d_tile_samples, (void *)c->d_n_active, (void *)c->d_scount}) if (q) cudaFree(q);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    delete c;
}

int rtu_synchronize(rtu_context *c)
{
    if (!c) { rtu::set_error("rtu_synchronize: null context"); return RTU_ERR_INVALID; }
    CU(cudaStreamSynchronize(c->stream));
    return RTU_OK;
}

static int scene_upload_impl(rtu_context *c, const rtu_scene_desc *d, rtu_scene **out)
{
    if (!c || !d || !out) { rtu::set_error("rtu_scene_upload: null argument"); return RTU_ERR_INVALID; }
    if (d->n_nodes < 1 || !d->nodes) { rtu::set_error("rtu_scene_upload: scene has no root node"); return RTU_ERR_INVALID; }
    if (d->n_meshes < 0 || d->n_materials < 0 || d->n_lights < 0 || d->n_texmaps < 0 || (d->n_meshes > 0 && !d->meshes) ||
        (d->n_materials > 0 && !d->materials) || (d->n_lights > 0 && !d->lights) || (d->n_texmaps > 0 && !d->texmaps)) {
        rtu::set_error("rtu_scene_upload: a non-empty array of the scene description is NULL");
        return RTU_ERR_INVALID;
    }
    CU(cudaSetDevice(c->device));
    std::unique_ptr<rtu_scene> sc(new rtu_scene);
    sc->ctx = c;
    sc->cam = d->camera;
    g_upload_bytes = 0;
    memset(&sc->S, 0, sizeof sc->S);
    int rc;
    auto fail = [&](int code) { free_list_async(sc->owned, c->stream); return code; };
    auto cuda_fail = [&](cudaError_t e, const char *what) {
        rtu::set_error(std::string(what) + ": " + cudaGetErrorString(e));
        return fail(RTU_ERR_CUDA);
    };
    sc->serial = g_scene_serial.fetch_add(1);

    // nodes
    std::vector<DNode> nodes(d->n_nodes);
    int flat = 1;
    for (int i = 0; i < d->n_nodes; i++) {
        const rtu_node &n = d->nodes[i];
        DNode &o = nodes[i];
        memset(&o, 0, sizeof o);
        memcpy(o.itm, n.itm, sizeof o.itm);
        memcpy(o.pos, n.pos, sizeof o.pos);
        memcpy(o.tm, n.tm, sizeof o.tm);
        if ((i == 0) != (n.parent < 0) || n.parent >= i) { rtu::set_error("rtu_scene_upload: nodes must be in pre-order with node 0 as the only root"); return fail(RTU_ERR_INVALID); }
        o.parent = n.parent;
        o.depth = i == 0 ? 0 : nodes[n.parent].depth + 1;
        if (o.depth >= RTU_MAX_DEPTH) { rtu::set_error("rtu_scene_upload: scene graph deeper than RTU_MAX_DEPTH"); return fail(RTU_ERR_UNSUPPORTED); }
        if (i > 0 && n.parent != 0) flat = 0;
        o.kind = n.kind;
        o.mesh = n.mesh;
        o.material = (n.material >= 0 && n.material < d->n_materials) ? n.material : -1;
        if (n.kind == RTU_OBJ_MESH && (n.mesh < 0 || n.mesh >= d->n_meshes)) { rtu::set_error("rtu_scene_upload: bad mesh index"); return fail(RTU_ERR_INVALID); }
        if (n.kind < 0 || n.kind > 3) { rtu::set_error("rtu_scene_upload: bad object kind"); return fail(RTU_ERR_INVALID); }
    }
    if (d->nodes[0].kind != RTU_OBJ_NONE) { rtu::set_error("rtu_scene_upload: the root node cannot hold an object (rootNode never does)"); return fail(RTU_ERR_UNSUPPORTED); }
    // light masks (device_scene.h LightMask, host/light_mask.cpp): per mesh node, one per light that casts hard shadows
    std::vector<LightMask> light_masks;
    std::vector<uint32_t> mask_bits;
    uint32_t *dmask_lists = nullptr;
    {
        static_assert(sizeof(LightMask) == 24 * sizeof(float), "host/light_mask.cpp writes the record as 24 words");
        std::vector<rtu_light_mask> got;
        std::vector<rtu::OwnedMask> own;
        std::vector<const rtu_light_mask *> list_src; // masks with light lists, in the order of their device offsets
        size_t list_words = 0;
        // the caller's where they fit the scene (rtu_host_load_xml builds them), else built here within ~0.3 s of host time
        rtu::collect_light_masks(*d, &got, &own, 2u << 20);
        for (const rtu_light_mask &g : got) { // ordered by node, a node's lights in order, its eye mask last
            DNode &o = nodes[g.node];
            if (o.mask_count == 0) o.mask_first = (int32_t)light_masks.size();
            if (g.light < 0) o.mask_count |= RTU_MASK_HAS_EYE;
            else o.mask_count++;
            LightMask lm;
            memcpy(&lm, g.rec, sizeof lm);
            lm.bits = (uint32_t)mask_bits.size();
            mask_bits.insert(mask_bits.end(), g.bits, g.bits + RTU_MASK_RES * RTU_MASK_RES / 32);
            lm.cells = 0xffffffffu;
            lm.items = 0;
            const uint32_t n_tris = d->meshes[d->nodes[g.node].mesh].nf;
            if (g.cell_start && g.items && g.n_items > 0 && g.light >= 0 && g.cell_start[RTU_MASK_RES * RTU_MASK_RES] == g.n_items &&
                list_words + (size_t)RTU_MASK_RES * RTU_MASK_RES + 2 + 2 * (size_t)g.n_items < 0xfffffff0u) {
                // light lists: the offsets of the cells, then the (slot, depth) pairs; validated like everything the device indexes
                // with, copied to the device straight from the caller's arrays (several MB: no second copy on the host)
                bool ok = true;
                for (int cidx = 0; cidx < RTU_MASK_RES * RTU_MASK_RES; cidx++) ok &= g.cell_start[cidx] <= g.cell_start[cidx + 1];
                for (uint32_t k = 0; k < g.n_items; k++) ok &= g.items[2 * k] < n_tris;
                if (ok) {
                    lm.cells = (uint32_t)list_words;
                    list_words += (size_t)RTU_MASK_RES * RTU_MASK_RES + 1;
                    list_words += list_words & 1u; // the pairs are read as uint2
                    lm.items = (uint32_t)list_words;
                    list_words += 2 * (size_t)g.n_items;
                    list_src.push_back(&g);
                }
            }
            light_masks.push_back(lm);
        }
        if (list_words) {
            g_upload_bytes += list_words * sizeof(uint32_t);
            cudaError_t e = cudaMallocAsync((void **)&dmask_lists, list_words * sizeof(uint32_t), c->stream);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(light lists)");
            sc->owned.push_back(dmask_lists);
            size_t li = 0;
            for (const LightMask &lm : light_masks) {
                if (lm.cells == 0xffffffffu) continue;
                const rtu_light_mask &g = *list_src[li++];
                if ((e = cudaMemcpyAsync(dmask_lists + lm.cells, g.cell_start, ((size_t)RTU_MASK_RES * RTU_MASK_RES + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess ||
                    (e = cudaMemcpyAsync(dmask_lists + lm.items, g.items, 2 * (size_t)g.n_items * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess)
                    return cuda_fail(e, "cudaMemcpyAsync(light lists)");
            }
        }
    }
    DNode *dn = nullptr;
    if ((rc = dev_upload(nodes, &dn, c->stream, sc->owned))) return fail(rc);
    LightMask *dmasks = nullptr;
    uint32_t *dmask_bits = nullptr;
    if (!light_masks.empty()) {
        if ((rc = dev_upload(light_masks, &dmasks, c->stream, sc->owned))) return fail(rc);
        if ((rc = dev_upload(mask_bits, &dmask_bits, c->stream, sc->owned))) return fail(rc);
    }
    // bounding spheres for the conservative cull (traverse.cuh), in root space, evaluated in double
    std::vector<float4> bounds(d->n_nodes);
    for (int i = 0; i < d->n_nodes; i++) {
        const rtu_node &n = d->nodes[i];
        bounds[i] = make_float4(0, 0, 0, -1.f);
        if (n.kind == RTU_OBJ_NONE) continue;
        double lo[3] = {-1, -1, -1}, hi[3] = {1, 1, 1};
        if (n.kind == RTU_OBJ_PLANE) { lo[2] = hi[2] = 0; }
        if (n.kind == RTU_OBJ_MESH) {
            const rtu_mesh &m = d->meshes[n.mesh];
            if (m.nf == 0 || (m.bvh_nodes < 2 && !(m.flags & RTU_MESH_DEVICE_BVH))) { bounds[i].w = -2.f; continue; }
            for (int k = 0; k < 3; k++) { lo[k] = m.bound_min[k]; hi[k] = m.bound_max[k]; }
        }
        auto to_root = [&](double p[3]) { // FromNodeCoords chain up to (not including) the root
            for (int a = i; a > 0; a = d->nodes[a].parent) {
                const rtu_node &t = d->nodes[a];
                double q[3];
                for (int r = 0; r < 3; r++) q[r] = (double)t.tm[r] * p[0] + (double)t.tm[3 + r] * p[1] + (double)t.tm[6 + r] * p[2] + (double)t.pos[r];
                p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
            }
        };
        double ctr[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
        to_root(ctr);
        double r2 = 0;
        bool finite = std::isfinite(ctr[0]) && std::isfinite(ctr[1]) && std::isfinite(ctr[2]);
        for (int corner = 0; corner < 8 && finite; corner++) {
            double p[3] = {(corner & 1) ? hi[0] : lo[0], (corner & 2) ? hi[1] : lo[1], (corner & 4) ? hi[2] : lo[2]};
            to_root(p);
            double dx = p[0] - ctr[0], dy = p[1] - ctr[1], dz = p[2] - ctr[2];
            double q = dx * dx + dy * dy + dz * dz;
            if (!std::isfinite(q)) finite = false;
            r2 = std::max(r2, q);
        }
        {   // the same 8 corners, kept for the per-frame image-space footprint of the object (setup_frame)
            rtu_scene::Footprint fp;
            fp.finite = finite;
            for (int corner = 0; corner < 8; corner++) {
                double p[3] = {(corner & 1) ? hi[0] : lo[0], (corner & 2) ? hi[1] : lo[1], (corner & 4) ? hi[2] : lo[2]};
                to_root(p);
                for (int k = 0; k < 3; k++) fp.c[corner][k] = p[k];
            }
            fp.node = i;
            sc->footprints.push_back(fp);
        }
        if (!finite) { bounds[i].w = 3.0e38f; continue; } // never culls
        double rad = std::sqrt(r2) * 1.002 + 1e-5 * (std::fabs(ctr[0]) + std::fabs(ctr[1]) + std::fabs(ctr[2]) + 1.0);
        bounds[i] = make_float4((float)ctr[0], (float)ctr[1], (float)ctr[2], (float)std::min(rad * rad, 3.0e38));
    }
    sc->n_obj = 0;
    for (int i = 0; i < d->n_nodes; i++) if (d->nodes[i].kind != RTU_OBJ_NONE) sc->n_obj++;
    {   // the bounding spheres are in root space; camera rays are in world space: the same thing for the identity root of every XML scene
        const rtu_node &r = d->nodes[0];
        const float ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        sc->root_identity = memcmp(r.itm, ident, sizeof ident) == 0 && r.pos[0] == 0.f && r.pos[1] == 0.f && r.pos[2] == 0.f;
    }
    float4 *db = nullptr;
    if ((rc = dev_upload(bounds, &db, c->stream, sc->owned))) return fail(rc);
    // top-level hierarchy over the bounding spheres for scenes with many nodes (device_scene.h TopNode)
    std::vector<TopNode> top_nodes;
    std::vector<int32_t> top_items, obj_rank(d->n_nodes, 0);
    {
        int rank = 0;
        for (int i = 0; i < d->n_nodes; i++) { if (d->nodes[i].kind != RTU_OBJ_NONE) rank++; obj_rank[i] = rank; }
        // below that, stepping through every node in lock-step is cheaper; scenes with meshes stay with the kernels that pool the
        // mesh walks per warp much longer (100 spheres at 1080p 16 spp: 24.0 ms in lock-step, 12.9 ms searched)
        int top_min = d->n_meshes > 0 ? 256 : 32;
        if (const char *e = getenv("RTU_TOP_MIN")) top_min = atoi(e);
        std::vector<int> objs;
        bool usable = top_min > 0 && rank >= top_min;
        for (int i = 0; i < d->n_nodes && usable; i++) {
            if (bounds[i].w < 0.f) continue;            // no object / never hit: nothing to nominate
            if (!(bounds[i].w < 1.0e37f)) usable = false; // a node that must never be culled: keep the linear visit
            objs.push_back(i);
        }
        if (usable && !objs.empty()) {
            struct Box { float lo[3], hi[3]; };
            // the box a ray has to cross to pass the object's own bound-box gate: the world-space extent of that bound
            // box (tighter than the bounding sphere, which is what the per-node cull uses), inflated by 1e-4
            std::vector<int> fp_of(d->n_nodes, -1);
            for (size_t f = 0; f < sc->footprints.size(); f++) fp_of[sc->footprints[f].node] = (int)f;
            auto box_of = [&](int i) {
                Box b;
                const rtu_scene::Footprint &fp = sc->footprints[fp_of[i]];
                for (int k = 0; k < 3; k++) {
                    double lo = fp.c[0][k], hi = fp.c[0][k];
                    for (int c8 = 1; c8 < 8; c8++) { lo = std::min(lo, fp.c[c8][k]); hi = std::max(hi, fp.c[c8][k]); }
                    double m = 1e-4 * (hi - lo) + 1e-5 * (std::fabs(lo) + std::fabs(hi) + 1.0);
                    b.lo[k] = (float)(lo - m);
                    b.hi[k] = (float)(hi + m);
                }
                return b;
            };
            std::vector<Box> obj_box(d->n_nodes);
            for (int i : objs) obj_box[i] = box_of(i);
            const bool sah = !getenv("RTU_TOP_MEDIAN"); // (A/B: median splits along the longest axis of the centres)
            int leaf_max = 1; // (measured 1 / 2 / 3 / 4 / 6 / 8: 14.9 / 15.2 / 15.6 / 15.9 / 16.4 / 16.7 ms per 4-spp frame of 10 000 spheres)
            if (const char *e = getenv("RTU_TOP_LEAF")) leaf_max = std::min(std::max(atoi(e), 1), 16);
            std::function<int(int, int, int)> build = [&](int first, int last, int depth) -> int { // objs[first, last)
                int me = (int)top_nodes.size();
                top_nodes.push_back(TopNode());
                Box bb = obj_box[objs[first]];
                float cmin[3], cmax[3];
                {
                    const float c0[3] = {bounds[objs[first]].x, bounds[objs[first]].y, bounds[objs[first]].z};
                    for (int k = 0; k < 3; k++) cmin[k] = cmax[k] = c0[k];
                }
                for (int j = first; j < last; j++) {
                    const Box &b = obj_box[objs[j]];
                    const float cc[3] = {bounds[objs[j]].x, bounds[objs[j]].y, bounds[objs[j]].z};
                    for (int k = 0; k < 3; k++) {
                        bb.lo[k] = std::min(bb.lo[k], b.lo[k]); bb.hi[k] = std::max(bb.hi[k], b.hi[k]);
                        cmin[k] = std::min(cmin[k], cc[k]); cmax[k] = std::max(cmax[k], cc[k]);
                    }
                }
                TopNode n;
                for (int k = 0; k < 3; k++) { n.lo[k] = bb.lo[k]; n.hi[k] = bb.hi[k]; }
                if (last - first <= leaf_max) {
                    n.a = -((int)top_items.size() + 1);
                    n.b = last - first;
                    for (int j = first; j < last; j++) top_items.push_back(objs[j]);
                    top_nodes[me] = n;
                    return me;
                }
                int mid = -1;
                if (sah && depth < 28) { // (deeper: median splits, so that the searches' 40-entry stacks always suffice)
                    // binned surface-area heuristic over the three axes (16 bins of the centres' extent): the search opens
                    // a box with a probability proportional to its area
                    const int NB = 16;
                    double best = 1e300;
                    int best_axis = -1, best_bin = -1;
                    auto area = [](const Box &b) {
                        const double x = (double)b.hi[0] - b.lo[0], y = (double)b.hi[1] - b.lo[1], z = (double)b.hi[2] - b.lo[2];
                        return x * y + y * z + z * x;
                    };
                    auto bin_of = [&](int obj, int axis) {
                        const float c = axis == 0 ? bounds[obj].x : (axis == 1 ? bounds[obj].y : bounds[obj].z);
                        int bi = (int)((c - cmin[axis]) / (cmax[axis] - cmin[axis]) * NB);
                        return std::min(std::max(bi, 0), NB - 1);
                    };
                    for (int axis = 0; axis < 3; axis++) {
                        if (!(cmax[axis] > cmin[axis])) continue;
                        Box bbx[NB];
                        int cnt[NB] = {0};
                        for (int j = first; j < last; j++) {
                            const int bi = bin_of(objs[j], axis);
                            const Box &b = obj_box[objs[j]];
                            if (cnt[bi]++ == 0) bbx[bi] = b;
                            else for (int k = 0; k < 3; k++) { bbx[bi].lo[k] = std::min(bbx[bi].lo[k], b.lo[k]); bbx[bi].hi[k] = std::max(bbx[bi].hi[k], b.hi[k]); }
                        }
                        double right_area[NB];
                        int right_cnt[NB];
                        Box acc;
                        int c = 0;
                        for (int bi = NB - 1; bi > 0; bi--) {
                            if (cnt[bi]) {
                                if (c == 0) acc = bbx[bi];
                                else for (int k = 0; k < 3; k++) { acc.lo[k] = std::min(acc.lo[k], bbx[bi].lo[k]); acc.hi[k] = std::max(acc.hi[k], bbx[bi].hi[k]); }
                                c += cnt[bi];
                            }
                            right_area[bi] = c ? area(acc) : 0.0;
                            right_cnt[bi] = c;
                        }
                        c = 0;
                        for (int bi = 0; bi < NB - 1; bi++) { // split between bin bi and bi + 1
                            if (cnt[bi]) {
                                if (c == 0) acc = bbx[bi];
                                else for (int k = 0; k < 3; k++) { acc.lo[k] = std::min(acc.lo[k], bbx[bi].lo[k]); acc.hi[k] = std::max(acc.hi[k], bbx[bi].hi[k]); }
                                c += cnt[bi];
                            }
                            if (c == 0 || right_cnt[bi + 1] == 0) continue;
                            const double cost = area(acc) * c + right_area[bi + 1] * right_cnt[bi + 1];
                            if (cost < best) { best = cost; best_axis = axis; best_bin = bi; }
                        }
                    }
                    if (best_axis >= 0) {
                        auto it = std::partition(objs.begin() + first, objs.begin() + last, [&](int o) { return bin_of(o, best_axis) <= best_bin; });
                        mid = (int)(it - objs.begin());
                        if (mid <= first || mid >= last) mid = -1;
                    }
                }
                if (mid < 0) {
                    int axis = 0;
                    if (cmax[1] - cmin[1] > cmax[axis] - cmin[axis]) axis = 1;
                    if (cmax[2] - cmin[2] > cmax[axis] - cmin[axis]) axis = 2;
                    mid = (first + last) / 2;
                    std::nth_element(objs.begin() + first, objs.begin() + mid, objs.begin() + last, [&](int a, int b) {
                        const float ca[3] = {bounds[a].x, bounds[a].y, bounds[a].z}, cb[3] = {bounds[b].x, bounds[b].y, bounds[b].z};
                        return ca[axis] < cb[axis];
                    });
                }
                n.a = build(first, mid, depth + 1);
                n.b = build(mid, last, depth + 1);
                top_nodes[me] = n;
                return me;
            };
            build(0, (int)objs.size(), 0);
        }
    }
    TopNode *dtop = nullptr;
    int32_t *dtop_items = nullptr, *drank = nullptr;
    float4 *dtop_bounds = nullptr;
    std::vector<float4> top_bounds(top_items.size());
    for (size_t k = 0; k < top_items.size(); k++) top_bounds[k] = bounds[top_items[k]];
    if ((rc = dev_upload(top_nodes, &dtop, c->stream, sc->owned))) return fail(rc);
    if ((rc = dev_upload(top_items, &dtop_items, c->stream, sc->owned))) return fail(rc);
    if ((rc = dev_upload(top_bounds, &dtop_bounds, c->stream, sc->owned))) return fail(rc);
    if ((rc = dev_upload(obj_rank, &drank, c->stream, sc->owned))) return fail(rc);
    // meshes
    std::vector<DMesh> meshes(d->n_meshes);
    for (int i = 0; i < d->n_meshes; i++)
        if ((rc = pack_mesh(d->meshes[i], &meshes[i], c->stream, sc->owned, &sc->bvh_build_ms))) return fail(rc);
    DMesh *dm = nullptr;
    if ((rc = dev_upload(meshes, &dm, c->stream, sc->owned))) return fail(rc);
    // textures (pixel arrays de-duplicated by host pointer)
    std::vector<DTexMap> tms(d->n_texmaps);
    std::vector<std::pair<const uint8_t *, uint8_t *>> pix;
    for (int i = 0; i < d->n_texmaps; i++) {
        const rtu_texmap &t = d->texmaps[i];
        DTexMap &o = tms[i];
        memset(&o, 0, sizeof o);
        o.kind = t.kind;
        memcpy(o.itm, t.itm, sizeof o.itm);
        memcpy(o.pos, t.pos, sizeof o.pos);
        memcpy(o.c1, t.color1, sizeof o.c1);
        memcpy(o.c2, t.color2, sizeof o.c2);
        o.width = t.width;
        o.height = t.height;
        if (t.kind == RTU_TEX_FILE) {
            if (!t.rgb8 || t.width <= 0 || t.height <= 0) { o.width = o.height = 0; o.rgb8 = nullptr; continue; }
            uint8_t *dp = nullptr;
            for (auto &pr : pix) if (pr.first == t.rgb8) dp = pr.second;
            if (!dp) {
                size_t bytes = (size_t)t.width * t.height * 3;
                cudaError_t te = cudaMallocAsync((void **)&dp, bytes, c->stream);
                if (te != cudaSuccess) return cuda_fail(te, "rtu_scene_upload: texture allocation");
                sc->owned.push_back(dp);
                te = cudaMemcpyAsync(dp, t.rgb8, bytes, cudaMemcpyHostToDevice, c->stream);
                if (te != cudaSuccess) return cuda_fail(te, "rtu_scene_upload: texture copy");
                pix.push_back({t.rgb8, dp});
                g_upload_bytes += bytes;
            }
            o.rgb8 = dp;
        }
    }
    DTexMap *dt = nullptr;
    if ((rc = dev_upload(tms, &dt, c->stream, sc->owned))) return fail(rc);
    // materials
    std::vector<DMaterial> mats(d->n_materials);
    bool any_refl = false, any_refr = false;
    for (int i = 0; i < d->n_materials; i++) {
        const rtu_material &m = d->materials[i];
        DMaterial &o = mats[i];
        // Kr / Kt are colour x texture sample: a black colour can never spawn a ray (mtlFunctions.cpp:160, 273)
        any_refl = any_refl || m.reflection.color[0] != 0.f || m.reflection.color[1] != 0.f || m.reflection.color[2] != 0.f;
        any_refr = any_refr || m.refraction.color[0] != 0.f || m.refraction.color[1] != 0.f || m.refraction.color[2] != 0.f;
        o.diffuse = pack_tc(m.diffuse, d->n_texmaps);
        o.specular = pack_tc(m.specular, d->n_texmaps);
        o.reflection = pack_tc(m.reflection, d->n_texmaps);
        o.refraction = pack_tc(m.refraction, d->n_texmaps);
        o.glossiness = m.glossiness;
        memcpy(o.absorption, m.absorption, sizeof o.absorption);
        o.ior = m.ior;
        o.refl_gloss = m.reflection_glossiness;
        o.refr_gloss = m.refraction_glossiness;
    }
    DMaterial *dmt = nullptr;
    if ((rc = dev_upload(mats, &dmt, c->stream, sc->owned))) return fail(rc);
    // lights
    std::vector<DLight> lts(d->n_lights);
    sc->n_shadow_lights = 0;
    for (int i = 0; i < d->n_lights; i++) {
        const rtu_light &l = d->lights[i];
        DLight &o = lts[i];
        o.kind = l.kind;
        memcpy(o.I, l.intensity, sizeof o.I);
        memcpy(o.v, l.v, sizeof o.v);
        o.size = l.size;
        if (l.kind != RTU_LIGHT_AMBIENT) sc->n_shadow_lights++;
        if (i == 0) { sc->h_light0_kind = l.kind; memcpy(sc->h_light0_I, l.intensity, sizeof sc->h_light0_I); }
    }
    DLight *dl = nullptr;
    if ((rc = dev_upload(lts, &dl, c->stream, sc->owned))) return fail(rc);
    {
        cudaError_t se = cudaStreamSynchronize(c->stream);
        if (se != cudaSuccess) return cuda_fail(se, "rtu_scene_upload: cudaStreamSynchronize");
    }
    sc->tree_waves = any_refr ? 2 : (any_refl ? 1 : 0);

    DScene &S = sc->S;
    S.nodes = dn;
    S.bounds = db;
    S.n_nodes = d->n_nodes;
    S.flat = flat;
    S.n_obj = sc->n_obj;
    S.n_top = (int)top_nodes.size();
    S.top = dtop;
    S.top_items = dtop_items;
    S.top_bounds = dtop_bounds;
    S.light_masks = dmasks;
    S.mask_bits = dmask_bits;
    S.mask_lists = dmask_lists;
    S.scan_min = 12; // tools/scan_sweep.sh: Teapot 1080p 6.53 / 6.34 / 6.30 / 6.34 / 7.30 ms of shadow waves per 128 spp at 1 / 8 / 12 / 16 / never
    if (const char *e = getenv("RTU_SCAN_MIN")) S.scan_min = atoi(e);
    S.obj_rank = drank;
    S.any_no_ref = 0;
    for (int m = 0; m < d->n_meshes; m++) if (d->meshes[m].flags & RTU_MESH_DEVICE_BVH) S.any_no_ref = 1;
    S.pool_ok = d->n_meshes > 0 ? 1 : 0; // without meshes nothing would ever be pooled
    for (int m = 0; m < d->n_meshes; m++)
        if (d->meshes[m].nf > (1u << 24) || d->meshes[m].bvh_nodes >= (1u << 27)) S.pool_ok = 0;
    S.meshes = dm;
    S.materials = dmt;
    S.n_materials = d->n_materials;
    S.lights = dl;
    S.n_lights = d->n_lights;
    S.n_shadow_lights = sc->n_shadow_lights;
    S.texmaps = dt;
    S.background = pack_tc(d->background, d->n_texmaps);
    S.environment = pack_tc(d->environment, d->n_texmaps);
    memcpy(S.cam_pos, d->camera.pos, sizeof S.cam_pos);
    sc->device_bytes = g_upload_bytes;
    *out = sc.release();
    return RTU_OK;
}

void rtu_scene_destroy(rtu_scene *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    free_list_async(s->owned, s->ctx->stream); // after everything already queued on the context's stream
    if (s->d_photons) cudaFreeAsync(s->d_photons, s->ctx->stream);
    if (s->d_knn) cudaFreeAsync(s->d_knn, s->ctx->stream);
    delete s;
}

} // extern "C"

namespace {

int frame_dims(const rtu_scene *s, const rtu_params *p, int *W, int *H)
{
    if (p->width < 0 || p->height < 0) { rtu::set_error("negative image size"); return RTU_ERR_INVALID; }
    *W = p->width > 0 ? p->width : s->cam.width;
    *H = p->height > 0 ? p->height : s->cam.height;
    if (*W <= 0 || *H <= 0 || (long long)*W * *H > (1ll << 28)) { rtu::set_error("bad image size"); return RTU_ERR_INVALID; }
    return RTU_OK;
}

int ensure_image(rtu_scene *s, size_t npix)
{
    if (npix <= s->ctx->fb.img_n) return RTU_OK;
    CU(cudaStreamSynchronize(s->ctx->stream));
    if (s->ctx->fb.d_rgb) cudaFree(s->ctx->fb.d_rgb);
    if (s->ctx->fb.d_rgb8) cudaFree(s->ctx->fb.d_rgb8);
    if (s->ctx->fb.d_z) cudaFree(s->ctx->fb.d_z);
    if (s->ctx->fb.d_z8) cudaFree(s->ctx->fb.d_z8);
    if (s->ctx->fb.d_node) cudaFree(s->ctx->fb.d_node);
    if (s->ctx->fb.d_face) cudaFree(s->ctx->fb.d_face);
    s->ctx->fb.d_rgb = nullptr; s->ctx->fb.d_rgb8 = nullptr; s->ctx->fb.d_z = nullptr; s->ctx->fb.d_z8 = nullptr; s->ctx->fb.d_node = nullptr; s->ctx->fb.d_face = nullptr;
    s->ctx->fb.img_n = 0;
    CU(cudaMalloc((void **)&s->ctx->fb.d_rgb, npix * 3 * sizeof(float)));
    CU(cudaMalloc((void **)&s->ctx->fb.d_rgb8, npix * 3));
    CU(cudaMalloc((void **)&s->ctx->fb.d_z, npix * sizeof(float)));
    CU(cudaMalloc((void **)&s->ctx->fb.d_z8, npix));
    CU(cudaMalloc((void **)&s->ctx->fb.d_node, npix * sizeof(int)));
    CU(cudaMalloc((void **)&s->ctx->fb.d_face, npix * sizeof(int)));
    s->ctx->fb.img_n = npix;
    return RTU_OK;
}

int ensure_accum(rtu_scene *s, size_t npix)
{
    if (npix <= s->ctx->fb.accum_n) return RTU_OK;
    CU(cudaStreamSynchronize(s->ctx->stream));
    if (s->ctx->fb.accum) cudaFree(s->ctx->fb.accum);
    s->ctx->fb.accum = nullptr;
    s->ctx->fb.accum_n = 0;
    CU(cudaMalloc((void **)&s->ctx->fb.accum, npix * sizeof(float4)));
    s->ctx->fb.accum_n = npix;
    return RTU_OK;
}

// Image-space footprint of every object for one camera: the convex hull of the projected corners of its bound box
// (clipped against the eye plane), as a pixel-space bounding box plus outward half-planes.  A camera ray can only pass an
// object's bound-box gate (objFunctions.cpp:17,109,337) where its image-plane point lies inside that hull.
// Returns false when the camera is degenerate (then no tile is ever called empty).
bool image_footprints(const rtu_scene *s, const DCamera &C, std::vector<TileObject> *objs, std::vector<float4> *edges)
{
    // image-plane point of pixel coordinates (fi,fj): origin + fi u + fj v; ray direction D = w0 + fi u + fj v with
    // w0 = origin - eye.  A world point X lies on that ray iff X - eye = t (w0 + fi u + fj v), t > 0: solve the 3x3 system
    double M[3][3], inv[3][3];
    for (int r = 0; r < 3; r++) { M[r][0] = (double)C.origin[r] - (double)C.pos[r]; M[r][1] = C.u[r]; M[r][2] = C.v[r]; }
    double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                 M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    if (!std::isfinite(det) || std::fabs(det) < 1e-300) return false;
    inv[0][0] = (M[1][1] * M[2][2] - M[1][2] * M[2][1]) / det; inv[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) / det; inv[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) / det;
    inv[1][0] = (M[1][2] * M[2][0] - M[1][0] * M[2][2]) / det; inv[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) / det; inv[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) / det;
    inv[2][0] = (M[1][0] * M[2][1] - M[1][1] * M[2][0]) / det; inv[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) / det; inv[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) / det;
    const double len0 = std::sqrt(M[0][0] * M[0][0] + M[1][0] * M[1][0] + M[2][0] * M[2][0]);
    // nothing nearer than 1e-6 along a ray can be hit (the primitives' own epsilons are 1e-5 and 1e-3): clip there
    const double t_near = 1e-6 / std::max(len0, 1e-30) * 0.5;
    static const int E[12][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    for (const rtu_scene::Footprint &fp : s->footprints) {
        TileObject o;
        o.first_edge = (int)edges->size();
        o.n_edges = 0;
        o.lo[0] = o.lo[1] = -3.0e38f; o.hi[0] = o.hi[1] = 3.0e38f; // everywhere
        if (!fp.finite) { objs->push_back(o); continue; }
        double q[8][3];
        bool ok = true;
        for (int c = 0; c < 8; c++) {
            double d[3] = {fp.c[c][0] - (double)C.pos[0], fp.c[c][1] - (double)C.pos[1], fp.c[c][2] - (double)C.pos[2]};
            for (int r = 0; r < 3; r++) q[c][r] = inv[r][0] * d[0] + inv[r][1] * d[1] + inv[r][2] * d[2];
            if (!std::isfinite(q[c][0]) || !std::isfinite(q[c][1]) || !std::isfinite(q[c][2])) ok = false;
        }
        if (!ok) { objs->push_back(o); continue; }
        std::vector<std::pair<double, double>> pts;
        for (int c = 0; c < 8; c++) if (q[c][0] >= t_near) pts.push_back({q[c][1] / q[c][0], q[c][2] / q[c][0]});
        for (auto &e : E) {
            const double *a = q[e[0]], *b = q[e[1]];
            if ((a[0] - t_near) * (b[0] - t_near) < 0) {
                double l = (t_near - a[0]) / (b[0] - a[0]);
                pts.push_back({(a[1] + l * (b[1] - a[1])) / t_near, (a[2] + l * (b[2] - a[2])) / t_near});
            }
        }
        if (pts.empty()) { o.lo[0] = o.lo[1] = 1.f; o.hi[0] = o.hi[1] = -1.f; objs->push_back(o); continue; } // entirely behind the eye
        double lo[2] = {pts[0].first, pts[0].second}, hi[2] = {lo[0], lo[1]}, big = 0;
        for (auto &p : pts) {
            lo[0] = std::min(lo[0], p.first); hi[0] = std::max(hi[0], p.first);
            lo[1] = std::min(lo[1], p.second); hi[1] = std::max(hi[1], p.second);
            big = std::max(big, std::max(std::fabs(p.first), std::fabs(p.second)));
        }
        if (!std::isfinite(big) || big > 1e7) { objs->push_back(o); continue; } // touches the eye plane: treat as everywhere
        const double margin = 0.05 + 1e-5 * big; // pixels: far above the float error of ray construction and slab tests
        o.lo[0] = (float)(lo[0] - margin); o.lo[1] = (float)(lo[1] - margin);
        o.hi[0] = (float)(hi[0] + margin); o.hi[1] = (float)(hi[1] + margin);
        // convex hull (monotone chain), counter-clockwise
        std::sort(pts.begin(), pts.end());
        pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
        std::vector<std::pair<double, double>> h(2 * pts.size() + 2);
        int k = 0;
        auto cross = [](const std::pair<double, double> &O, const std::pair<double, double> &A, const std::pair<double, double> &B) {
            return (A.first - O.first) * (B.second - O.second) - (A.second - O.second) * (B.first - O.first);
        };
        for (size_t i = 0; i < pts.size(); i++) { while (k >= 2 && cross(h[k - 2], h[k - 1], pts[i]) <= 0) k--; h[k++] = pts[i]; }
        for (size_t i = pts.size() - 1, t = k + 1; i > 0; i--) { while ((size_t)k >= t && cross(h[k - 2], h[k - 1], pts[i - 1]) <= 0) k--; h[k++] = pts[i - 1]; }
        int n = k - 1;
        if (n >= 3) {
            for (int i = 0; i < n; i++) {
                const auto &A = h[i], &B = h[(i + 1) % n];
                double ex = B.first - A.first, ey = B.second - A.second, l = std::sqrt(ex * ex + ey * ey);
                if (!(l > 0)) continue;
                double nx = ey / l, ny = -ex / l; // outward normal of a counter-clockwise polygon
                edges->push_back(make_float4((float)nx, (float)ny, (float)(nx * A.first + ny * A.second + margin + 1e-4 * big * 1e-3), 0.f));
                o.n_edges++;
            }
        }
        objs->push_back(o);
    }
    return true;
}

int setup_frame(rtu_scene *s, const rtu_params *p, FrameSetup *F, int *s_begin, int *s_end, bool *mask_launched)
{
    int W, H, rc;
    if ((rc = frame_dims(s, p, &W, &H))) return rc;
    if (p->spp < 1 || p->spp > (1 << 20)) { rtu::set_error("bad spp"); return RTU_ERR_INVALID; }
    if (p->pattern == RTU_PATTERN_CENTER && p->spp != 1) { rtu::set_error("RTU_PATTERN_CENTER needs spp == 1"); return RTU_ERR_INVALID; }
    if (p->mode != RTU_MODE_WHITTED && p->mode != RTU_MODE_PATH && p->mode != RTU_MODE_PHOTON && p->mode != RTU_MODE_PHOTON_GATHER) { rtu::set_error("bad render mode"); return RTU_ERR_INVALID; }
    if (p->mode == RTU_MODE_PHOTON_GATHER && (p->gi_bounces < 1 || p->gi_bounces > 16)) { rtu::set_error("gi_bounces out of range (1..16)"); return RTU_ERR_INVALID; }
    if ((p->mode == RTU_MODE_PHOTON || p->mode == RTU_MODE_PHOTON_GATHER) && !s->d_photons) { rtu::set_error("RTU_MODE_PHOTON needs a photon map: call rtu_photon_map_generate or rtu_photon_map_set first"); return RTU_ERR_INVALID; }
    if (p->mode == RTU_MODE_PATH && (p->gi_bounces < 0 || p->gi_bounces > 6)) { rtu::set_error("gi_bounces out of range (0..6)"); return RTU_ERR_INVALID; }
    if (p->shade_bounces < 0 || p->shade_bounces > 15) { rtu::set_error("shade_bounces out of range"); return RTU_ERR_INVALID; }
    make_camera(s->cam, W, H, &F->cam);
    F->spp = p->spp;
    F->row_begin = 0;
    F->row_end = H;
    if (p->row_begin != 0 || p->row_end != 0) { F->row_begin = p->row_begin; F->row_end = p->row_end; }
    if (F->row_begin < 0 || F->row_end > H || F->row_begin >= F->row_end) { rtu::set_error("bad row range"); return RTU_ERR_INVALID; }
    *s_begin = 0;
    *s_end = p->spp;
    if (p->sample_begin != 0 || p->sample_end != 0) { *s_begin = p->sample_begin; *s_end = p->sample_end; }
    if (*s_begin < 0 || *s_end > p->spp || *s_begin >= *s_end) { rtu::set_error("bad sample range"); return RTU_ERR_INVALID; }
    F->mode = p->mode;
    F->shade_bounces = p->shade_bounces;
    F->gi_bounces = p->gi_bounces;
    F->flags = p->flags;
    F->seed = make_uint2((unsigned)(p->seed & 0xffffffffu), (unsigned)(p->seed >> 32));
    // sub-pixel offsets of every sample (RenderFunctions.cpp:71,81-85,96); uploaded only when the pattern changes
    rtu_context *c = s->ctx;
    if (c->off_spp != p->spp || c->off_pattern != p->pattern) {
        c->off_spp = c->off_pattern = -1;
        c->tile_key = rtu_context::TileKey();
        std::vector<float2> &off = c->h_offsets;
        off.resize(p->spp);
        if (p->pattern == RTU_PATTERN_CENTER) {
            off[0] = make_float2(0.5f, 0.5f);
        } else {
            float pixelIncrement = 1.0 / p->spp;
            for (int i = 0; i < p->spp; i++) {
                float cur = i * pixelIncrement;
                off[i] = make_float2(cur + halton(i, 4), cur + halton(i, 5));
            }
        }
        if ((size_t)p->spp > c->fb.offsets_n) {
            CU(cudaStreamSynchronize(c->stream));
            if (c->fb.d_offsets) cudaFree(c->fb.d_offsets);
            c->fb.d_offsets = nullptr;
            c->fb.offsets_n = 0;
            CU(cudaMalloc((void **)&c->fb.d_offsets, (size_t)p->spp * sizeof(float2)));
            c->fb.offsets_n = p->spp;
        }
        void *stage = c->stage_off.acquire(off.size() * sizeof(float2));
        if (!stage) { rtu::set_error("page-locked staging allocation failed"); return RTU_ERR_CUDA; }
        memcpy(stage, off.data(), off.size() * sizeof(float2));
        CU(cudaMemcpyAsync(c->fb.d_offsets, stage, off.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        c->stage_off.submitted(c->stream);
        c->off_spp = p->spp;
        c->off_pattern = p->pattern;
    }
    const std::vector<float2> &off = c->h_offsets;
    F->sample_offsets = c->fb.d_offsets;
    // tiles of the primary wave that no camera ray of the rendered samples can leave with a hit
    F->tile_empty = nullptr;
    F->n_empty_tiles = nullptr;
    F->n_tiles = 0;
    F->n_obj = s->n_obj;
    F->tile_done = c->adaptive_on ? c->d_tile_done : nullptr;
    F->half_split = c->adaptive_on ? (unsigned)((size_t)W * H) : 0u;
    *mask_launched = false;
    if (!getenv("RTU_NO_TILE_MASK") && s->cam.dof <= 0.f && s->root_identity) {
        rtu_context::TileKey key;
        key.scene = s->serial; key.W = W; key.H = H; key.row0 = F->row_begin; key.row1 = F->row_end;
        key.s0 = *s_begin; key.s1 = *s_end; key.spp = p->spp; key.pattern = p->pattern;
        if (c->tile_mask && key == c->tile_key) { // same scene, camera, size and samples as the last frame: the mask is still on the device
            F->tile_empty = c->tile_mask;
            F->n_empty_tiles = c->tile_count;
            F->n_tiles = c->tile_total;
            return RTU_OK;
        }
        c->tile_mask = nullptr;
        float ox0 = off[*s_begin].x, ox1 = ox0, oy0 = off[*s_begin].y, oy1 = oy0;
        for (int i = *s_begin; i < *s_end; i++) {
            ox0 = std::min(ox0, off[i].x); ox1 = std::max(ox1, off[i].x);
            oy0 = std::min(oy0, off[i].y); oy1 = std::max(oy1, off[i].y);
        }
        std::vector<TileObject> objs;
        std::vector<float4> edges;
        if (image_footprints(s, F->cam, &objs, &edges)) {
            size_t tiles = (size_t)((W + 7) / 8) * (size_t)((F->row_end - F->row_begin + 3) / 4);
            const size_t objs_b = ((objs.size() * sizeof(TileObject) + 15) / 16) * 16, edges_b = edges.size() * sizeof(float4);
            size_t need = tiles + objs_b + edges_b + 128;
            if (need > c->fb.tile_n) {
                CU(cudaStreamSynchronize(c->stream));
                if (c->fb.d_tile) cudaFree(c->fb.d_tile);
                c->fb.d_tile = nullptr;
                c->fb.tile_n = 0;
                CU(cudaMalloc((void **)&c->fb.d_tile, need));
                c->fb.tile_n = need;
            }
            // layout: [objects][edges][mask][count]; objects and edges go up in one copy from page-locked staging
            unsigned char *base = c->fb.d_tile;
            TileObject *d_objs = (TileObject *)base;
            float4 *d_edges = (float4 *)(base + objs_b);
            unsigned char *d_mask = base + objs_b + edges_b;
            unsigned *d_count = (unsigned *)(base + ((objs_b + edges_b + tiles + 15) / 16) * 16);
            if (objs_b + edges_b > 0) {
                unsigned char *stage = (unsigned char *)c->stage_tile.acquire(objs_b + edges_b);
                if (!stage) { rtu::set_error("page-locked staging allocation failed"); return RTU_ERR_CUDA; }
                if (!objs.empty()) memcpy(stage, objs.data(), objs.size() * sizeof(TileObject));
                if (!edges.empty()) memcpy(stage + objs_b, edges.data(), edges_b);
                CU(cudaMemcpyAsync(base, stage, objs_b + edges_b, cudaMemcpyHostToDevice, c->stream));
                c->stage_tile.submitted(c->stream);
            }
            launch_tile_mask(c->stream, *F, d_objs, (int)objs.size(), d_edges, ox0, ox1, oy0, oy1, d_mask, d_count);
            *mask_launched = true;
            F->tile_empty = d_mask;
            F->n_empty_tiles = d_count;
            F->n_tiles = (unsigned)tiles;
            c->tile_key = key;
            c->tile_mask = d_mask;
            c->tile_count = d_count;
            c->tile_total = (unsigned)tiles;
            if (getenv("RTU_TILE_DEBUG")) {
                std::vector<unsigned char> hm(tiles);
                cudaMemcpy(hm.data(), d_mask, tiles, cudaMemcpyDeviceToHost);
                size_t e = 0;
                for (unsigned char v : hm) e += v;
                fprintf(stderr, "tile mask: %zu of %zu tiles empty, %zu objects, %zu hull edges, offsets [%g,%g]x[%g,%g]\n", e, tiles, objs.size(), edges.size(), ox0, ox1, oy0, oy1);
                for (auto &o : objs) fprintf(stderr, "  object bbox [%g,%g]x[%g,%g] edges %d\n", o.lo[0], o.hi[0], o.lo[1], o.hi[1], o.n_edges);
            }
        }
    }
    return RTU_OK;
}

// the waves that follow a first wave whose output is in q[out_q]
int wave_count(const FrameSetup &F, int tree_waves)
{
    // Waves one Shade tree can last: none when no material of the scene reflects or refracts (nothing is ever spawned),
    // one per bounce with mirrors only; with refraction a Fresnel ray starts one wave after its sibling at every level.
    // Every GI vertex restarts a Shade tree.
    const int tree = tree_waves == 0 ? 0 : (tree_waves == 1 ? F.shade_bounces : 2 * F.shade_bounces + 1);
    return tree + (F.mode == RTU_MODE_PATH ? F.gi_bounces : 0);
}

// wave_log: where the rays entering each wave are noted (NULL: nowhere); tail_w0 >= 0: waves [tail_w0, n_waves) as ONE
// cooperative launch that stops at the first empty wave (k_tail_waves)
int run_waves(rtu_scene *s, const FrameSetup &F, float4 *accum, int out_q, size_t *work_i, int tree_waves, unsigned *wave_log = nullptr,
              int tail_w0 = -1)
{
    rtu_context *c = s->ctx;
    int n_waves = wave_count(F, tree_waves);
    kt_begin(c, 2);
    const bool refwalk = (F.flags & RTU_FLAG_REFERENCE_WALK) != 0;
    launch_shadow_wave(c->cfg, c->stream, s->S, c->wb, accum, c->work + (*work_i)++, refwalk);
    kt_end(c);
    s->launches++;
    int in_q = out_q;
    for (int w = 0; w < n_waves; w++) {
        if (w == tail_w0) {
            if (launch_tail_waves(c->cfg, c->stream, s->S, F, c->wb, in_q, w, n_waves, accum, c->work + *work_i, wave_log)) {
                *work_i += 3 * (size_t)(n_waves - w);
                s->launches++;
                c->tail_launches++;
                break;
            }
            cudaGetLastError(); // no cooperative launch on this device: one by one
        }
        launch_reset_counts(c->stream, c->wb.q[1 - in_q].count, c->wb.aux[1 - in_q].count, c->wb.shadow.count, c->wb.hits.count,
                            wave_log ? wave_log + w : nullptr, c->wb.q[in_q].count);
        kt_begin(c, 1);
        launch_extend_queue(c->cfg, c->stream, s->S, F, c->wb, in_q, accum, c->work + (*work_i)++);
        kt_end(c);
        kt_begin(c, 3);
        launch_shade_queue(c->cfg, c->stream, s->S, F, c->wb, in_q, accum, c->work + (*work_i)++);
        kt_end(c);
        kt_begin(c, 2);
        launch_shadow_wave(c->cfg, c->stream, s->S, c->wb, accum, c->work + (*work_i)++, refwalk);
        kt_end(c);
        s->launches += 4;
        in_q = 1 - in_q;
    }
    CU(cudaGetLastError());
    return RTU_OK;
}

DPhotonMap photon_map_of(const rtu_scene *s)
{
    DPhotonMap PM;
    PM.map = s->d_photons;
    PM.knn_nodes = s->d_knn;
    PM.knn_dir = s->d_knn ? s->d_knn + 3 * ((size_t)s->n_photons + 1) : nullptr;
    PM.knn_pw = s->d_knn ? s->d_knn + 4 * ((size_t)s->n_photons + 1) : nullptr;
    PM.n = (int)s->n_photons;
    PM.half = (int)s->n_photons / 2 - 1; // halfStoredPhotons (cyPhotonMap.h:227)
    PM.radius = s->photon_params.est_radius;
    PM.norm_scale = s->photon_params.ellipticity == 1.f ? 0.f : 1.f / s->photon_params.ellipticity - 1.f;
    return PM;
}

int check_overflow(rtu_scene *s, DCounters *host)
{
    rtu_context *c = s->ctx;
    CU(cudaMemcpyAsync(host, c->wb.counters, sizeof(DCounters), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (host->overflow) {
        rtu::set_error("ray queue overflow: lower RTU_CHUNK_RAYS");
        return RTU_ERR_UNSUPPORTED;
    }
    return RTU_OK;
}

} // namespace

extern "C" {

static int render_device_once(rtu_scene *s, const rtu_params *p, float4 *accum, int32_t clear_accum)
{
    rtu_context *c = s->ctx;
    FrameSetup F;
    int s0, s1, rc;
    bool mask_launched = false;
    if ((rc = setup_frame(s, p, &F, &s0, &s1, &mask_launched))) return rc;
    int W = F.cam.width, H = F.cam.height;
    size_t npix = (size_t)W * H;
    int rows = F.row_end - F.row_begin;
    size_t per_sample = (size_t)((W + 7) / 8) * ((rows + 3) / 4) * 32;
    size_t chunk_samples, chunk_cap;
    for (;;) {
        size_t chunk_rays = c->chunk_rays;
        if (s->chunk_limit && s->chunk_limit < chunk_rays) chunk_rays = s->chunk_limit;
        if (F.mode == RTU_MODE_PATH && !c->chunk_rays_set) chunk_rays = std::min<size_t>(chunk_rays, (size_t)1 << 27); // (+192 B of GI record per entry; no gain measured)
        chunk_samples = std::max<size_t>(1, chunk_rays / per_sample);
        chunk_samples = std::min<size_t>(chunk_samples, (size_t)(s1 - s0)); // queues are sized for what this call renders
        chunk_cap = std::max(per_sample, chunk_samples * per_sample);
        if (chunk_cap >= (1ull << 31)) { rtu::set_error("image too large for one wave; use row ranges"); return RTU_ERR_UNSUPPORTED; }
        // Queue capacities: one entry per primary ray of a chunk.  A hit can spawn up to 3 rays, so no
        // fixed factor is a bound; overflow is detected on the device (DCounters::overflow) and the frame is
        // rendered again with larger queues / smaller chunks (render_checked).
        s->last_chunk_cap = chunk_cap;
        size_t q_cap = (size_t)((double)chunk_cap * c->queue_factor * s->queue_boost);
        if (q_cap < chunk_cap) q_cap = chunk_cap;
        size_t sh_cap = q_cap * (size_t)std::max(1, s->n_shadow_lights);
        rc = ensure_scratch(c, q_cap, sh_cap);
        if (rc == RTU_OK) break;
        // the device cannot spare that much: smaller chunks (for this context from now on) rather than no frame
        if (!c->scratch_oom || chunk_samples <= 1 || c->chunk_rays <= (1u << 22)) return rc;
        c->chunk_rays = std::min(c->chunk_rays, chunk_cap) / 2;
    }
    size_t n_chunks = ((size_t)(s1 - s0) + chunk_samples - 1) / chunk_samples;
    size_t launches_per_chunk = 3 + 3 * (size_t)wave_count(F, s->tree_waves);
    if ((rc = ensure_work(c, n_chunks * launches_per_chunk + 8))) return rc;
    const bool path_mode = F.mode == RTU_MODE_PATH;
    // primary hits of one chunk: what the per-hit buffers (GI records, photon queries) are sized for - not the hit queue's capacity,
    // which remembers the largest frame this context ever rendered
    const size_t frame_hits = std::min<size_t>(c->wb.hits.cap, chunk_cap);
    float4 *target = accum; // the array ray slots index: pixels, or GI records folded into pixels per chunk
    if (path_mode) {
        if ((rc = ensure_gi(c, frame_hits * (size_t)(2 * (F.gi_bounces + 1) + 2)))) return rc;
        target = c->gi;
    }
    c->adaptive_totals_valid = false;
    c->pixel_samples = (uint64_t)W * (uint64_t)rows * (uint64_t)(s1 - s0);
    CU(cudaEventRecord(c->ev0, c->stream));
    CU(cudaMemsetAsync(c->wb.counters, 0, sizeof(DCounters), c->stream));
    if (clear_accum) CU(cudaMemsetAsync(accum, 0, npix * (c->adaptive_on ? 2 : 1) * sizeof(float4), c->stream));
    s->launches = mask_launched ? 1 : 0; // k_tile_mask ran in setup_frame
    kt_reset(c, (p->flags & RTU_FLAG_TIME_KERNELS) != 0);
    // the deep waves of a frame this shape held few rays the last time: one launch for all of them
    rtu_context::WaveKey wkey;
    wkey.W = W; wkey.rows = rows; wkey.mode = F.mode; wkey.shade_bounces = F.shade_bounces; wkey.gi_bounces = F.gi_bounces;
    wkey.n_waves = wave_count(F, s->tree_waves); wkey.n_nodes = s->S.n_nodes; wkey.chunk_samples = chunk_samples;
    const bool log_waves = wkey.n_waves > 0 && wkey.n_waves <= rtu_context::WAVE_LOG_MAX;
    int tail_w0 = -1;
    if (log_waves && c->tail_rays > 0 && !(p->flags & (RTU_FLAG_TIME_KERNELS | RTU_FLAG_REFERENCE_WALK)) && c->wave_log_valid &&
        c->wave_key == wkey && cudaEventQuery(c->wave_ev) == cudaSuccess) {
        tail_w0 = wkey.n_waves;
        while (tail_w0 > 0 && c->h_wave_log[tail_w0 - 1] <= c->tail_rays) tail_w0--;
        if (tail_w0 >= wkey.n_waves) tail_w0 = -1;
    }
    if (log_waves && !(p->flags & (RTU_FLAG_TIME_KERNELS | RTU_FLAG_REFERENCE_WALK)) && getenv("RTU_TAIL_FORCE")) tail_w0 = 0; // (tests)
    cudaGetLastError(); // (cudaErrorNotReady of the query above is not an error)
    if (getenv("RTU_TAIL_DEBUG") && c->wave_log_valid && c->wave_key == wkey) {
        fprintf(stderr, "wave log:");
        for (int w = 0; w < wkey.n_waves; w++) fprintf(stderr, " %u", c->h_wave_log[w]);
        fprintf(stderr, " -> tail from wave %d (%llu tail launches so far)\n", tail_w0, (unsigned long long)c->tail_launches);
    }
    size_t wi = 0;
    for (int a = s0; a < s1; a += (int)chunk_samples) {
        int b = std::min<int>(s1, a + (int)chunk_samples);
        unsigned *wave_log = log_waves ? c->d_wave_log + (a == s0 ? 0 : rtu_context::WAVE_LOG_MAX) : nullptr;
        launch_reset_counts(c->stream, c->wb.q[0].count, c->wb.aux[0].count, c->wb.shadow.count, c->wb.hits.count);
        kt_begin(c, 0);
        launch_extend_primary(c->cfg, c->stream, s->S, F, a, b, c->wb, accum, target, c->work + wi++);
        kt_end(c);
        if (F.mode == RTU_MODE_PHOTON) { // PhotonMapping(ray, hInfo) per hit; no secondary or shadow rays (bounceCount 0)
            kt_begin(c, 3);
            CU(launch_photon_shade(c->stream, s->S, F, a, c->wb, (unsigned)frame_hits, photon_map_of(s), accum));
            kt_end(c);
            s->launches += 3 + 3 * (((unsigned)frame_hits + (1u << 20) - 1) >> 20);
            continue;
        }
        kt_begin(c, 3);
        launch_shade_primary(c->cfg, c->stream, s->S, F, a, c->wb, 0, target, c->work + wi++);
        if (F.mode == RTU_MODE_PHOTON_GATHER) { // + MonteCarloPhoton per primary hit (the hit queue is still intact)
            CU(launch_photon_gather(c->stream, s->S, F, a, c->wb, (unsigned)frame_hits, photon_map_of(s), accum));
            s->launches += 3 + 2 * (unsigned)((frame_hits * (size_t)std::max(1, F.gi_bounces) + (1u << 20) - 1) >> 20);
        }
        kt_end(c);
        s->launches += 3;
        if ((rc = run_waves(s, F, target, 0, &wi, s->tree_waves, wave_log, tail_w0))) return rc;
        if (path_mode) {
            launch_gi_combine(c->stream, c->gi, c->wb.gi_count, (unsigned)frame_hits, F.gi_bounces, accum);
            s->launches++;
        }
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    s->timed = true;
    c->wave_log_valid = false;
    if (log_waves && F.mode != RTU_MODE_PHOTON) {
        CU(cudaMemcpyAsync(c->h_wave_log, c->d_wave_log, rtu_context::WAVE_LOG_MAX * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaEventRecord(c->wave_ev, c->stream));
        c->wave_key = wkey;
        c->wave_log_valid = true;
    }
    CU(cudaGetLastError());
    return RTU_OK;
}

// After a queue overflow: more queue entries per primary ray (up to ~64 GB of queues), after that smaller chunks; the scene
// remembers the setting for its next frames.  Returns RTU_OK when another attempt makes sense.
static int grow_queues(rtu_scene *s, int attempt)
{
    rtu_context *c = s->ctx;
    const double bytes_per_entry = 300.0 * std::max(1, s->n_shadow_lights);
    const double doubled = (double)s->last_chunk_cap * c->queue_factor * s->queue_boost * 2.0 * bytes_per_entry;
    if (attempt >= 10) { rtu::set_error("ray queue overflow: lower RTU_CHUNK_RAYS or raise RTU_QUEUE_FACTOR"); return RTU_ERR_UNSUPPORTED; }
    if (doubled <= 64e9 && s->queue_boost < 16.0) s->queue_boost *= 2.0;
    else if (s->last_chunk_cap > 65536) s->chunk_limit = s->last_chunk_cap / 2;
    else { rtu::set_error("ray queue overflow: raise RTU_QUEUE_FACTOR"); return RTU_ERR_UNSUPPORTED; }
    c->queue_retries++;
    return RTU_OK;
}

// the overflow flag of the frame just enqueued, read through the page-locked word (waits for the frame)
static int frame_overflowed(rtu_context *c, bool *overflow)
{
    CU(cudaMemcpyAsync(c->h_flag, &c->wb.counters->overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *overflow = *c->h_flag != 0;
    return RTU_OK;
}


// Renders into `accum`; the queues hold one entry per primary ray of a chunk, and a scene whose hits spawn more rays than
// that (glossy rooms in RTU_MODE_PATH) overflows them: the device flags it and the frame is rendered again with larger
// queues.  That needs a frame that starts from zero, so a call that ADDS to the caller's accumulator (clear_accum == 0: the
// spp-sliced / row-sliced path) renders into the context's second accumulator and adds that to the caller's once it is
// known to be complete: an overflow never leaves a caller's accumulator half written.
// `out`: resolve + device->host copies enqueued behind the frame, covered by the same wait as the overflow flag.
static int render_checked(rtu_scene *s, const rtu_params *p, float *d_accum, int32_t clear_accum, rtu_image *out);
} // extern "C"
int rtu_render_checked(rtu_scene *s, const rtu_params *p, float *d_accum, int32_t clear_accum, rtu_image *out) { return render_checked(s, p, d_accum, clear_accum, out); }
extern "C" {
static int render_checked(rtu_scene *s, const rtu_params *p, float *d_accum, int32_t clear_accum, rtu_image *out)
{
    if (!s || !p) { rtu::set_error("rtu_render_device: null argument"); return RTU_ERR_INVALID; }
    rtu_context *c = s->ctx;
    CU(cudaSetDevice(c->device));
    int W, H, rc;
    if ((rc = frame_dims(s, p, &W, &H))) return rc;
    const size_t npix = (size_t)W * H * (c->adaptive_on ? 2 : 1); // adaptive frames: even and odd samples side by side
    float4 *dst = (float4 *)d_accum;
    if (!dst) {
        if ((rc = ensure_accum(s, npix))) return rc;
        dst = c->fb.accum;
    }
    float4 *render_into = dst;
    if (!clear_accum) {
        if (npix > c->fb.accum2_n) {
            CU(cudaStreamSynchronize(c->stream));
            if (c->fb.accum2) cudaFree(c->fb.accum2);
            c->fb.accum2 = nullptr;
            c->fb.accum2_n = 0;
            CU(cudaMalloc((void **)&c->fb.accum2, npix * sizeof(float4)));
            c->fb.accum2_n = npix;
        }
        render_into = c->fb.accum2;
    }
    for (int attempt = 0;; attempt++) {
        if ((rc = render_device_once(s, p, render_into, 1))) return rc;
        if (out && clear_accum && (rc = rtu_resolve_enqueue(s, p, dst, out))) return rc; // optimistic: dropped if the frame overflowed
        bool overflow = false;
        if ((rc = frame_overflowed(c, &overflow))) return rc;
        if (!overflow) break;
        if ((rc = grow_queues(s, attempt))) return rc;
    }
    if (!clear_accum) {
        launch_accum_add(c->stream, dst, render_into, npix);
        s->launches++;
        if (out) {
            if ((rc = rtu_resolve_enqueue(s, p, dst, out))) return rc;
            CU(cudaStreamSynchronize(c->stream));
        }
    }
    CU(cudaGetLastError());
    return RTU_OK;
}

// Adaptive sampling (rtu_params::adaptive_min_spp > 0): passes over the tiles that have not converged yet.
// Every pass is an ordinary frame over a range of sample indices of the spp-sample pattern, added to the split accumulator;
// after it k_adaptive_update retires the tiles whose worst pixel is at or below the target and counts the rest.
static int render_adaptive(rtu_scene *s, const rtu_params *p, rtu_image *out)
{
    rtu_context *c = s->ctx;
    CU(cudaSetDevice(c->device));
    int W, H, rc;
    if ((rc = frame_dims(s, p, &W, &H))) return rc;
    if (p->mode != RTU_MODE_WHITTED && p->mode != RTU_MODE_PATH) { rtu::set_error("adaptive sampling: RTU_MODE_WHITTED and RTU_MODE_PATH only"); return RTU_ERR_UNSUPPORTED; }
    if (p->sample_begin || p->sample_end || p->row_begin || p->row_end) { rtu::set_error("adaptive sampling renders whole frames (no sample / row ranges)"); return RTU_ERR_INVALID; }
    if (p->pattern != RTU_PATTERN_REFERENCE || p->spp < 2 || p->adaptive_min_spp < 2 || p->adaptive_min_spp > p->spp || !(p->adaptive_target >= 0.f)) {
        rtu::set_error("adaptive sampling: needs RTU_PATTERN_REFERENCE, 2 <= adaptive_min_spp <= spp and adaptive_target >= 0");
        return RTU_ERR_INVALID;
    }
    const size_t tiles = (size_t)((W + 7) / 8) * ((H + 3) / 4), npix = (size_t)W * H;
    if (tiles > c->adaptive_tiles) {
        CU(cudaStreamSynchronize(c->stream));
        for (void *q : {(void *)c->d_tile_done, (void *)c->d_tile_samples, (void *)c->d_n_active}) if (q) cudaFree(q);
        c->d_tile_done = nullptr; c->d_tile_samples = nullptr; c->d_n_active = nullptr;
        c->adaptive_tiles = 0;
        CU(cudaMalloc((void **)&c->d_tile_done, tiles));
        CU(cudaMalloc((void **)&c->d_tile_samples, tiles * sizeof(int)));
        CU(cudaMalloc((void **)&c->d_n_active, sizeof(unsigned)));
        c->adaptive_tiles = tiles;
    }
    CU(cudaMemsetAsync(c->d_tile_done, 0, tiles, c->stream));
    CU(cudaMemsetAsync(c->d_tile_samples, 0, tiles * sizeof(int), c->stream));
    const int step = p->adaptive_step > 0 ? p->adaptive_step : 8;
    struct Off { rtu_context *c; ~Off() { c->adaptive_on = false; } } off{c}; // every exit leaves the context in fixed-spp mode
    c->adaptive_on = true;
    uint64_t spent = 0, active_tiles = tiles, launches = 0;
    double device_ms = 0;
    rtu_stats total;
    memset(&total, 0, sizeof total);
    for (int n = 0; n < p->spp && active_tiles > 0;) {
        const int n_next = n == 0 ? p->adaptive_min_spp : std::min(p->spp, n + step);
        rtu_params q = *p;
        q.sample_begin = n;
        q.sample_end = n_next;
        q.adaptive_min_spp = 0;
        if ((rc = render_checked(s, &q, nullptr, n == 0 ? 1 : 0, nullptr))) return rc;
        launch_adaptive_update(c->stream, c->fb.accum, W, H, n_next, p->spp, p->adaptive_target, c->d_tile_done, c->d_tile_samples, c->d_n_active);
        CU(cudaMemcpyAsync(c->h_flag + 1, c->d_n_active, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        spent += active_tiles * 32ull * (uint64_t)(n_next - n); // (tiles at the image border count whole)
        active_tiles = c->h_flag[1];
        rtu_stats st;
        if (rtu_get_stats(s, &st) == RTU_OK) { // the counters are per call: add the passes up
            total.trace_rays += st.trace_rays; total.shadow_rays += st.shadow_rays; total.box_tests += st.box_tests;
            total.tri_tests += st.tri_tests; total.node_visits += st.node_visits;
            device_ms += st.device_ms;
            launches += st.kernel_launches + 1;
        }
        n = n_next;
    }
    c->adaptive_on = false;
    c->pixel_samples = spent;
    c->adaptive_totals = total;
    c->adaptive_totals.device_ms = device_ms;
    c->adaptive_totals.kernel_launches = launches;
    c->adaptive_totals_valid = true;
    // the pixel's colour: mean over its own sample count
    if ((rc = ensure_image(s, npix))) return rc;
    if (out->sample_count && npix > c->scount_n) {
        if (c->d_scount) cudaFree(c->d_scount);
        c->d_scount = nullptr;
        c->scount_n = 0;
        CU(cudaMalloc((void **)&c->d_scount, npix));
        c->scount_n = npix;
    }
    if (out->rgb || out->rgb8 || out->sample_count) {
        launch_resolve_adaptive(c->stream, c->fb.accum, W, H, c->d_tile_samples, out->rgb ? c->fb.d_rgb : nullptr, out->rgb8 ? c->fb.d_rgb8 : nullptr,
                                out->sample_count ? c->d_scount : nullptr);
        if (out->rgb) CU(cudaMemcpyAsync(out->rgb, c->fb.d_rgb, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (out->rgb8) CU(cudaMemcpyAsync(out->rgb8, c->fb.d_rgb8, npix * 3, cudaMemcpyDeviceToHost, c->stream));
        if (out->sample_count) CU(cudaMemcpyAsync(out->sample_count, c->d_scount, npix, cudaMemcpyDeviceToHost, c->stream));
    }
    rtu_image rest = *out;
    rest.rgb = nullptr; rest.rgb8 = nullptr; rest.sample_count = nullptr;
    if (rest.z || rest.z8 || rest.node_id || rest.face_id)
        if ((rc = rtu_resolve_enqueue(s, p, nullptr, &rest))) return rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return RTU_OK;
}

int rtu_render_device(rtu_scene *s, const rtu_params *p, float *d_accum, int32_t clear_accum)
{
    if (p && p->adaptive_min_spp > 0) { rtu::set_error("rtu_render_device: adaptive sampling goes through rtu_render / rtu_render_async"); return RTU_ERR_UNSUPPORTED; }
    return rtu::guarded("rtu_render_device", [&]() -> int { return render_checked(s, p, d_accum, clear_accum, nullptr); });
}

} // extern "C"

int rtu_frame_dims(const rtu_scene *s, const rtu_params *p, int *W, int *H) { return frame_dims(s, p, W, H); }
int rtu_ensure_image(rtu_scene *s, size_t npix) { return ensure_image(s, npix); }

// accum -> device images -> host buffers, enqueued on the context's stream (no wait)

int rtu_resolve_enqueue(rtu_scene *s, const rtu_params *p, const float4 *accum, rtu_image *out)
{
    rtu_context *c = s->ctx;
    int W, H, rc;
    if ((rc = frame_dims(s, p, &W, &H))) return rc;
    if (p->spp < 1) { rtu::set_error("rtu_resolve: spp must be at least 1"); return RTU_ERR_INVALID; }
    size_t npix = (size_t)W * H;
    if ((rc = ensure_image(s, npix))) return rc;
    if (out->rgb || out->rgb8) {
        if (!accum) { rtu::set_error("rtu_resolve: nothing rendered yet"); return RTU_ERR_INVALID; }
        if (accum == c->fb.accum && npix > c->fb.accum_n) { rtu::set_error("rtu_resolve: the last frame rendered into the internal accumulator was smaller than this image"); return RTU_ERR_INVALID; }
        launch_resolve(c->stream, accum, (int)npix, 0.f, p->spp, out->rgb ? c->fb.d_rgb : nullptr, out->rgb8 ? c->fb.d_rgb8 : nullptr);
        if (out->rgb) CU(cudaMemcpyAsync(out->rgb, c->fb.d_rgb, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (out->rgb8) CU(cudaMemcpyAsync(out->rgb8, c->fb.d_rgb8, npix * 3, cudaMemcpyDeviceToHost, c->stream));
    }
    if (out->z || out->z8 || out->node_id || out->face_id) {
        DCamera cam;
        make_camera(s->cam, W, H, &cam);
        // visibility at pixel centres: the z the reference meant to store (SURVEY A-3); always the whole image, whatever
        // row range the frame's samples covered
        if (!c->wb.counters) { if ((rc = ensure_scratch(c, 1024, 1024))) return rc; }
        launch_primary_ids(c->cfg, c->stream, s->S, cam, c->fb.d_z, c->fb.d_node, c->fb.d_face, c->wb.counters);
        if (out->z8) launch_zimage(c->stream, c->fb.d_z, (int)npix, c->zmm, c->fb.d_z8);
        if (out->z) CU(cudaMemcpyAsync(out->z, c->fb.d_z, npix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (out->z8) CU(cudaMemcpyAsync(out->z8, c->fb.d_z8, npix, cudaMemcpyDeviceToHost, c->stream));
        if (out->node_id) CU(cudaMemcpyAsync(out->node_id, c->fb.d_node, npix * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        if (out->face_id) CU(cudaMemcpyAsync(out->face_id, c->fb.d_face, npix * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    return RTU_OK;
}

extern "C" {

int rtu_resolve(rtu_scene *s, const rtu_params *p, const float *d_accum, rtu_image *out)
{
    if (!s || !p || !out) { rtu::set_error("rtu_resolve: null argument"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_resolve", [&]() -> int {
        rtu_context *c = s->ctx;
        CU(cudaSetDevice(c->device));
        int rc = rtu_resolve_enqueue(s, p, d_accum ? (const float4 *)d_accum : c->fb.accum, out);
        if (rc) return rc;
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        return RTU_OK;
    });
}

int rtu_render(rtu_scene *s, const rtu_params *p, rtu_image *out)
{
    if (!s || !p || !out) { rtu::set_error("rtu_render: null argument"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_render", [&]() -> int {
        int rc;
        if (p->mode == RTU_MODE_PRIMARY) {
            rtu_image o = *out;
            o.rgb = nullptr;
            o.rgb8 = nullptr;
            if ((rc = ensure_scratch(s->ctx, 1024, 1024))) return rc;
            CU(cudaMemsetAsync(s->ctx->wb.counters, 0, sizeof(DCounters), s->ctx->stream));
            CU(cudaEventRecord(s->ctx->ev0, s->ctx->stream));
            rc = rtu_resolve_enqueue(s, p, nullptr, &o);
            CU(cudaEventRecord(s->ctx->ev1, s->ctx->stream));
            s->timed = true;
            s->launches = o.z8 ? 3 : 1;
            if (rc) return rc;
            CU(cudaStreamSynchronize(s->ctx->stream));
            CU(cudaGetLastError());
            return RTU_OK;
        }
        if (p->adaptive_min_spp > 0) return render_adaptive(s, p, out);
        // frame, resolve and device->host copies are enqueued together: one wait covers the images and the overflow flag
        return render_checked(s, p, nullptr, 1, out);
    });
}

int rtu_get_stats(const rtu_scene *s, rtu_stats *out)
{
    if (!s || !out) { rtu::set_error("rtu_get_stats: null argument"); return RTU_ERR_INVALID; }
    rtu_context *c = s->ctx;
    memset(out, 0, sizeof *out);
    out->scene_device_bytes = s->device_bytes;
    out->bvh_build_ms = s->bvh_build_ms;
    out->queue_retries = c->queue_retries;
    out->pixel_samples = c->pixel_samples;
    if (!c->wb.counters) return RTU_OK; // nothing has run on this context yet
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    DCounters hc;
    CU(cudaMemcpy(&hc, c->wb.counters, sizeof hc, cudaMemcpyDeviceToHost));
    rtu_kernel_stats *ks[4] = {&out->primary_wave, &out->secondary_waves, &out->shadow_waves, &out->shade_kernels};
    out->shade_kernels.launches = c->cls_launches[3];
    for (int k = 0; k < 3; k++) {
        const DCounterBlock &b = hc.k[k];
        out->trace_rays += b.trace_rays;
        out->shadow_rays += b.shadow_rays;
        out->box_tests += b.box_tests;
        out->tri_tests += b.tri_tests;
        out->node_visits += b.node_visits;
        ks[k]->rays = b.trace_rays + b.shadow_rays;
        ks[k]->box_tests = b.box_tests;
        ks[k]->tri_tests = b.tri_tests;
        ks[k]->node_visits = b.node_visits;
        ks[k]->launches = c->cls_launches[k];
    }
    for (size_t i = 0; i < c->kt_used; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->kt_ev[i * 2], c->kt_ev[i * 2 + 1]) == cudaSuccess) ks[c->kt_cls[i]]->ms += ms;
    }
    out->scene_device_bytes = s->device_bytes;
    out->queue_retries = c->queue_retries;
    out->bvh_build_ms = s->bvh_build_ms;
    out->kernel_launches = s->launches;
    if (s->timed) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) out->device_ms = ms;
    }
    if (c->adaptive_totals_valid) { // the last call was an adaptive frame: its passes added up
        const rtu_stats &t = c->adaptive_totals;
        out->trace_rays = t.trace_rays; out->shadow_rays = t.shadow_rays; out->box_tests = t.box_tests; out->tri_tests = t.tri_tests;
        out->node_visits = t.node_visits; out->device_ms = t.device_ms; out->kernel_launches = t.kernel_launches;
    }
    if (hc.overflow) { rtu::set_error("ray queue overflow: lower RTU_CHUNK_RAYS"); return RTU_ERR_UNSUPPORTED; }
    return RTU_OK;
}

static int trace_impl(rtu_scene *s, const rtu_ray *rays, int64_t n, rtu_hit *hits)
{
    if (!s || (n > 0 && (!rays || !hits)) || n < 0) { rtu::set_error("rtu_trace: bad argument"); return RTU_ERR_INVALID; }
    if (n == 0) return RTU_OK;
    if (n >= (1ll << 30)) { rtu::set_error("rtu_trace: batch too large"); return RTU_ERR_UNSUPPORTED; }
    rtu_context *c = s->ctx;
    CU(cudaSetDevice(c->device));
    // The operator runs the kernel the frames run (k_extend_pool on the meshes' own hierarchies).  RTU_TRACE=reference walks
    // the cyBVH with the reference's tests inside the same kernel, RTU_TRACE=exact selects the plain per-lane walk of the cyBVH
    // in the reference's order (k_trace_batch): the tests use both as cross-checks.
    const char *sel = getenv("RTU_TRACE");
    const bool exact = sel && sel[0] == 'e', refwalk = sel && sel[0] == 'r';
    int rc;
    if ((rc = ensure_scratch(c, std::max<size_t>(c->q_cap, exact ? 1024 : (size_t)n), std::max<size_t>(c->shadow_cap, 1024)))) return rc;
    if ((rc = ensure_work(c, 8))) return rc;
    if ((rc = ensure_accum(s, 1))) return rc;
    rtu_ray *dr = nullptr;
    rtu_hit *dh = nullptr;
    CU(cudaMalloc((void **)&dr, n * sizeof(rtu_ray)));
    cudaError_t e = cudaMalloc((void **)&dh, n * sizeof(rtu_hit));
    if (e != cudaSuccess) { cudaFree(dr); CU(e); }
    e = cudaMemcpyAsync(dr, rays, n * sizeof(rtu_ray), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->wb.counters, 0, sizeof(DCounters), c->stream);
    if (e == cudaSuccess) {
        cudaEventRecord(c->ev0, c->stream);
        if (exact) {
            launch_trace_batch(c->cfg, c->stream, s->S, dr, n, dh, c->wb.counters);
        } else {
            launch_reset_counts(c->stream, c->wb.q[0].count, c->wb.aux[0].count, c->wb.shadow.count, c->wb.hits.count);
            launch_trace_batch_wave(c->cfg, c->stream, s->S, dr, n, dh, c->wb, c->fb.accum, c->work, refwalk);
        }
        cudaEventRecord(c->ev1, c->stream);
        s->timed = true;
        s->launches = exact ? 1 : 4;
        e = cudaMemcpyAsync(hits, dh, n * sizeof(rtu_hit), cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(dr);
    cudaFree(dh);
    CU(e);
    return RTU_OK;
}

static int shadow_trace_impl(rtu_scene *s, const rtu_ray *rays, const float *t_max, int64_t n, uint8_t *occluded)
{
    if (!s || (n > 0 && (!rays || !t_max || !occluded)) || n < 0) { rtu::set_error("rtu_shadow_trace: bad argument"); return RTU_ERR_INVALID; }
    if (n == 0) return RTU_OK;
    if (n >= (1ll << 30)) { rtu::set_error("rtu_shadow_trace: batch too large"); return RTU_ERR_UNSUPPORTED; }
    rtu_context *c = s->ctx;
    CU(cudaSetDevice(c->device));
    // The operator runs the kernel the frames run (k_shadow_wave: pooled walks of the any-hit hierarchy); RTU_SHADOW_TRACE=exact
    // selects the plain per-lane walk of the cyBVH instead (k_shadow_batch), which the tests use as the cross-check.
    const char *sel = getenv("RTU_SHADOW_TRACE");
    const bool wave = !(sel && sel[0] == 'e');
    int rc;
    if ((rc = ensure_scratch(c, std::max<size_t>(c->q_cap, 1024), std::max<size_t>(c->shadow_cap, wave ? (size_t)n : 1024)))) return rc;
    if ((rc = ensure_work(c, 8))) return rc;
    rtu_ray *dr = nullptr;
    float *dt = nullptr;
    unsigned char *docc = nullptr;
    float4 *acc = nullptr;
    CU(cudaMalloc((void **)&dr, n * sizeof(rtu_ray)));
    cudaError_t e = cudaMalloc((void **)&dt, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void **)&docc, n);
    if (e == cudaSuccess && wave) e = cudaMalloc((void **)&acc, n * sizeof(float4));
    if (e == cudaSuccess && wave) e = cudaMemsetAsync(acc, 0, n * sizeof(float4), c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dr, rays, n * sizeof(rtu_ray), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dt, t_max, n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->wb.counters, 0, sizeof(DCounters), c->stream);
    if (e == cudaSuccess) {
        cudaEventRecord(c->ev0, c->stream);
        if (wave) launch_shadow_batch_wave(c->cfg, c->stream, s->S, dr, dt, n, docc, c->wb, acc, c->work);
        else launch_shadow_batch(c->cfg, c->stream, s->S, dr, dt, n, docc, c->wb.counters);
        cudaEventRecord(c->ev1, c->stream);
        s->timed = true;
        s->launches = wave ? 3 : 1;
        e = cudaMemcpyAsync(occluded, docc, n, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(dr);
    if (dt) cudaFree(dt);
    if (docc) cudaFree(docc);
    if (acc) cudaFree(acc);
    CU(e);
    return RTU_OK;
}

static int selftest_division_impl(rtu_context *c, uint32_t numerators_per_divisor, uint64_t seed, uint64_t *tested, uint64_t *mismatches)
{
    if (!c || !tested || !mismatches) { rtu::set_error("rtu_selftest_division: null argument"); return RTU_ERR_INVALID; }
    CU(cudaSetDevice(c->device));
    unsigned long long *d = nullptr;
    CU(cudaMalloc((void **)&d, 2 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), c->stream);
    unsigned long long h[2] = {0, 0};
    if (e == cudaSuccess) {
        launch_selftest_div(c->stream, numerators_per_divisor, seed, d, d + 1);
        e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d);
    CU(e);
    *mismatches = h[0];
    *tested = h[1];
    return RTU_OK;
}

static int camera_rays_impl(rtu_scene *s, const rtu_params *p, int32_t sample, rtu_ray *rays)
{
    if (!s || !p || !rays) { rtu::set_error("rtu_camera_rays: null argument"); return RTU_ERR_INVALID; }
    rtu_context *c = s->ctx;
    CU(cudaSetDevice(c->device));
    int W, H, rc;
    if ((rc = frame_dims(s, p, &W, &H))) return rc;
    if (sample < 0 || sample >= p->spp) { rtu::set_error("rtu_camera_rays: bad sample index"); return RTU_ERR_INVALID; }
    float ox = 0.5f, oy = 0.5f;
    if (p->pattern == RTU_PATTERN_REFERENCE) {
        float pixelIncrement = 1.0 / p->spp;
        float cur = sample * pixelIncrement;
        ox = cur + halton(sample, 4);
        oy = cur + halton(sample, 5);
    }
    DCamera cam;
    make_camera(s->cam, W, H, &cam);
    size_t n = (size_t)W * H;
    rtu_ray *dr = nullptr;
    CU(cudaMalloc((void **)&dr, n * sizeof(rtu_ray)));
    launch_camera_rays(c->stream, cam, ox, oy, dr);
    cudaError_t e = cudaMemcpyAsync(rays, dr, n * sizeof(rtu_ray), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(dr);
    CU(e);
    return RTU_OK;
}

static int shade_impl(rtu_scene *s, const rtu_ray *rays, const rtu_hit *hits, int64_t n, int32_t bounces, float *rgb)
{
    if (!s || (n > 0 && (!rays || !hits || !rgb)) || n < 0 || bounces < 0 || bounces > 15) { rtu::set_error("rtu_shade: bad argument"); return RTU_ERR_INVALID; }
    if (n == 0) return RTU_OK;
    if (n >= (1ll << 27)) { rtu::set_error("rtu_shade: batch too large"); return RTU_ERR_UNSUPPORTED; }
    rtu_context *c = s->ctx;
    CU(cudaSetDevice(c->device));
    int rc;
    size_t q_cap = std::max<size_t>((size_t)n * 2, 1024);
    size_t sh_cap = q_cap * (size_t)std::max(1, s->n_shadow_lights);
    if ((rc = ensure_scratch(c, q_cap, sh_cap))) return rc;
    if ((rc = ensure_work(c, 3 + 3 * (size_t)(2 * bounces + 1) + 8))) return rc;
    FrameSetup F;
    memset(&F, 0, sizeof F);
    F.cam.width = (int)n;
    F.cam.height = 1;
    F.spp = 1;
    F.shade_bounces = bounces;
    F.row_end = 1;
    F.mode = RTU_MODE_WHITTED;
    rtu_ray *dr = nullptr;
    rtu_hit *dh = nullptr;
    float4 *acc = nullptr;
    std::vector<float4> host_acc((size_t)n);
    CU(cudaMalloc((void **)&dr, n * sizeof(rtu_ray)));
    cudaError_t e = cudaMalloc((void **)&dh, n * sizeof(rtu_hit));
    if (e == cudaSuccess) e = cudaMalloc((void **)&acc, n * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dr, rays, n * sizeof(rtu_ray), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dh, hits, n * sizeof(rtu_hit), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(acc, 0, n * sizeof(float4), c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->wb.counters, 0, sizeof(DCounters), c->stream);
    rc = RTU_OK;
    if (e == cudaSuccess) {
        size_t wi = 0;
        s->launches = 2;
        launch_reset_counts(c->stream, c->wb.q[0].count, c->wb.aux[0].count, c->wb.shadow.count, nullptr);
        launch_shade_batch(c->cfg, c->stream, s->S, F, dr, dh, n, c->wb, 0, acc);
        rc = run_waves(s, F, acc, 0, &wi, s->tree_waves);
        if (rc == RTU_OK) e = cudaMemcpyAsync(host_acc.data(), acc, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(dr);
    if (dh) cudaFree(dh);
    if (acc) cudaFree(acc);
    if (rc) return rc;
    CU(e);
    for (int64_t i = 0; i < n; i++) { rgb[i * 3] = host_acc[i].x; rgb[i * 3 + 1] = host_acc[i].y; rgb[i * 3 + 2] = host_acc[i].z; }
    DCounters hc;
    return check_overflow(s, &hc);
}

} // extern "C"


// ---------------------------------------------------------------------------------------------- photon map
namespace {

// Balances `n` raw photons into the scene's map and builds the estimate's walk records.  d_raw: the photons on the device;
// h_raw: the same on the host, or NULL (then they are read back if the host has to balance).  The device build
// (photon_build.cu) is the reference's tree whenever no median ties with a neighbour; with ties it says so and the host build
// (byte-identical to the reference in every case) takes over.  RTU_PHOTON_BUILD=host skips the device build.
int install_photon_map(rtu_scene *s, const rtu_photon *d_raw, const rtu_photon *h_raw, uint32_t n, const rtu_photon_params &pp)
{
    rtu_context *c = s->ctx;
    if (s->d_photons) { cudaFreeAsync(s->d_photons, c->stream); s->d_photons = nullptr; s->n_photons = 0; }
    if (s->d_knn) { cudaFreeAsync(s->d_knn, c->stream); s->d_knn = nullptr; }
    CU(cudaMallocAsync((void **)&s->d_photons, sizeof(rtu_photon) * ((size_t)n + 1), c->stream));
    CU(cudaMallocAsync((void **)&s->d_knn, sizeof(float4) * 5 * ((size_t)n + 1), c->stream));
    const char *how = getenv("RTU_PHOTON_BUILD");
    unsigned tie = 1;
    if (!(how && !strcmp(how, "host"))) CU(launch_photon_balance(c->stream, d_raw, n, s->d_photons, &tie));
    s->photon_device_build = tie == 0;
    if (tie) {
        // page-locked staging kept by the context: fresh pageable vectors of this size cost more (first-touch faults, slow copies)
        // than the balancing's top level
        if (!h_raw) {
            rtu_photon *raw = (rtu_photon *)c->stage_photon_raw.acquire(sizeof(rtu_photon) * (size_t)n);
            if (!raw) { rtu::set_error("page-locked staging allocation failed"); return RTU_ERR_CUDA; }
            CU(cudaMemcpyAsync(raw, d_raw, sizeof(rtu_photon) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            h_raw = raw;
        }
        rtu_photon *balanced = (rtu_photon *)c->stage_photon_bal.acquire(sizeof(rtu_photon) * ((size_t)n + 1));
        if (!balanced) { rtu::set_error("page-locked staging allocation failed"); return RTU_ERR_CUDA; }
        int rc = rtu_host_balance_photons(h_raw, n, balanced);
        if (rc) return rc;
        CU(cudaMemcpyAsync(s->d_photons, balanced, sizeof(rtu_photon) * ((size_t)n + 1), cudaMemcpyHostToDevice, c->stream));
        c->stage_photon_bal.submitted(c->stream);
    }
    CU(launch_knn_build(c->stream, s->d_photons, (int)n, (int)n / 2 - 1, s->d_knn, s->d_knn + 3 * ((size_t)n + 1), s->d_knn + 4 * ((size_t)n + 1)));
    s->n_photons = n;
    s->photon_params = pp;
    return RTU_OK;
}

} // namespace

extern "C" {

static int photon_map_set_impl(rtu_scene *s, const rtu_photon *photons, uint32_t n, const rtu_photon_params *params)
{
    if (!s || (n && !photons)) { rtu::set_error("rtu_photon_map_set: null argument"); return RTU_ERR_INVALID; }
    CU(cudaSetDevice(s->ctx->device));
    rtu_photon_params pp;
    rtu_photon_params_default(&pp);
    if (params) pp = *params;
    rtu_context *c = s->ctx;
    rtu_photon *d_raw = nullptr;
    if (n) {
        CU(cudaMallocAsync((void **)&d_raw, sizeof(rtu_photon) * (size_t)n, c->stream));
        cudaError_t e = cudaMemcpyAsync(d_raw, photons, sizeof(rtu_photon) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { cudaFreeAsync(d_raw, c->stream); CU(e); }
    }
    int rc = install_photon_map(s, d_raw, photons, n, pp);
    if (d_raw) cudaFreeAsync(d_raw, c->stream);
    return rc;
}

int rtu_photon_map_info(const rtu_scene *s, uint32_t *n_photons, uint32_t *device_build)
{
    if (!s) { rtu::set_error("rtu_photon_map_info: null scene"); return RTU_ERR_INVALID; }
    if (n_photons) *n_photons = s->d_photons ? s->n_photons : 0;
    if (device_build) *device_build = s->d_photons && s->photon_device_build ? 1u : 0u;
    return RTU_OK;
}

static int photon_map_get_impl(rtu_scene *s, rtu_photon *out, uint32_t cap, uint32_t *n)
{
    if (!s || !n) { rtu::set_error("rtu_photon_map_get: null argument"); return RTU_ERR_INVALID; }
    *n = s->n_photons;
    if (!out || !s->d_photons) return RTU_OK;
    CU(cudaSetDevice(s->ctx->device));
    uint32_t m = std::min(cap, s->n_photons);
    CU(cudaMemcpyAsync(out, s->d_photons + 1, sizeof(rtu_photon) * (size_t)m, cudaMemcpyDeviceToHost, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    return RTU_OK;
}

static int estimate_irradiance_impl(rtu_scene *s, const float *pos, const float *normal, int64_t n, float radius, float ellipticity,
                            float *irrad, float *direction, int32_t *found)
{
    if (!s || !pos || !irrad || !direction || n < 0) { rtu::set_error("rtu_estimate_irradiance: bad argument"); return RTU_ERR_INVALID; }
    if (!s->d_photons) { rtu::set_error("rtu_estimate_irradiance: the scene has no photon map"); return RTU_ERR_INVALID; }
    if (n == 0) return RTU_OK;
    rtu_context *c = s->ctx;
    CU(cudaSetDevice(c->device));
    float *dpos = nullptr, *dn = nullptr, *dirr = nullptr, *ddir = nullptr;
    int *dfound = nullptr;
    size_t b3 = (size_t)n * 3 * sizeof(float);
    cudaError_t e = cudaMallocAsync((void **)&dpos, b3, c->stream);
    if (e == cudaSuccess && normal) e = cudaMallocAsync((void **)&dn, b3, c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dirr, b3, c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&ddir, b3, c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dfound, (size_t)n * sizeof(int), c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpos, pos, b3, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && normal) e = cudaMemcpyAsync(dn, normal, b3, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        DPhotonMap PM = photon_map_of(s);
        float norm_scale = ellipticity == 1.f ? 0.f : 1.f / ellipticity - 1.f;
        e = launch_estimate(c->stream, PM, dpos, dn, n, radius, norm_scale, dirr, ddir, dfound);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(irrad, dirr, b3, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(direction, ddir, b3, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && found) e = cudaMemcpyAsync(found, dfound, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    for (void *q : {(void *)dpos, (void *)dn, (void *)dirr, (void *)ddir, (void *)dfound}) if (q) cudaFreeAsync(q, c->stream);
    CU(e);
    return RTU_OK;
}

static int photon_map_generate_impl(rtu_scene *s, const rtu_photon_params *params, rtu_photon_stats *stats)
{
    if (!s) { rtu::set_error("rtu_photon_map_generate: null scene"); return RTU_ERR_INVALID; }
    rtu_context *c = s->ctx;
    CU(cudaSetDevice(c->device));
    rtu_photon_params pp;
    rtu_photon_params_default(&pp);
    if (params) pp = *params;
    if (pp.map_size == 0 || pp.map_size >= (1u << 28) || pp.max_bounce == 0 || pp.max_bounce > 100) { rtu::set_error("bad photon parameters"); return RTU_ERR_INVALID; }
    // "TODO: Randomly decide on light source" (RenderFunctions.cpp:348): the reference always emits from lights[0]
    // and casts it to PointLight
    if (s->S.n_lights < 1 || s->h_light0_kind != RTU_LIGHT_POINT) { rtu::set_error("photon emission needs lights[0] to be a point light (RenderFunctions.cpp:349)"); return RTU_ERR_UNSUPPORTED; }
    int rc;
    if (!c->wb.counters) { if ((rc = ensure_scratch(c, 1024, 1024))) return rc; }
    if ((rc = ensure_work(c, 4096))) return rc;
    const unsigned batch = 1u << 19;
    const unsigned cap = pp.map_size;
    rtu_photon *staging = nullptr, *d_map = nullptr;
    unsigned char *d_counts = nullptr;
    unsigned *d_offsets = nullptr;
    std::vector<unsigned char> counts(batch);
    std::vector<unsigned> offsets(batch);
    cudaError_t e = cudaMallocAsync((void **)&staging, sizeof(rtu_photon) * (size_t)batch * pp.max_bounce, c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_map, sizeof(rtu_photon) * (size_t)cap, c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_counts, batch, c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_offsets, sizeof(unsigned) * batch, c->stream);
    auto cleanup = [&]() {
        for (void *q : {(void *)staging, (void *)d_map, (void *)d_counts, (void *)d_offsets}) if (q) cudaFreeAsync(q, c->stream);
    };
    if (e != cudaSuccess) { cleanup(); CU(e); }
    CU(cudaMemsetAsync(c->wb.counters, 0, sizeof(DCounters), c->stream));
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, c->stream);
    uint64_t paths = 0, from_light = 0, stored = 0;
    uint2 seed = make_uint2((unsigned)pp.seed, (unsigned)(pp.seed >> 32));
    size_t wi = 0;
    while (stored < cap) {
        if (wi >= 4096 || paths > (1ull << 40)) { cleanup(); rtu::set_error("photon emission does not converge (no photon surface reachable?)"); return RTU_ERR_UNSUPPORTED; }
        // size the batch from the yield so far, so that little is traced past the point where the map is full
        unsigned n_batch = batch;
        if (paths > 0 && stored > 0) {
            double need = (double)(cap - stored) * (double)paths / (double)stored * 1.02 + 1024.0;
            if (need < (double)batch) n_batch = (unsigned)need;
        }
        cudaMemsetAsync(c->work + wi, 0, sizeof(unsigned), c->stream);
        launch_photon_emit(c->cfg, c->stream, s->S, paths, n_batch, (int)pp.max_bounce, seed, 0, staging, d_counts, c->wb.counters, c->work + wi);
        wi++;
        e = cudaMemcpyAsync(counts.data(), d_counts, n_batch, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { cleanup(); CU(e); }
        // the sequential loop of GeneratePhotonMap: path k runs while the map is not yet full (RenderFunctions.cpp:346)
        unsigned cut = n_batch;
        for (unsigned k = 0; k < n_batch; k++) {
            if (stored >= cap) { cut = k; break; }
            offsets[k] = (unsigned)stored;
            if (counts[k] & 0x80u) from_light++;
            stored += counts[k] & 0x7fu; // photons past the capacity are dropped by AddPhoton
        }
        if (stored > cap) stored = cap;
        paths += cut;
        e = cudaMemcpyAsync(d_offsets, offsets.data(), sizeof(unsigned) * cut, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { cleanup(); CU(e); }
        launch_photon_compact(c->stream, staging, d_counts, d_offsets, n_batch, cut, (int)pp.max_bounce, d_map, cap);
        if (from_light == 0 && paths >= (8ull << 20)) { cleanup(); rtu::set_error("photon emission: no photon hits the scene"); return RTU_ERR_UNSUPPORTED; }
    }
    // scaleFactor = (lights[0] intensity / photonFromLight).Gray() (RenderFunctions.cpp:384)
    float fl = (float)(int)from_light;
    float scale = ((s->h_light0_I[0] / fl + s->h_light0_I[1] / fl) + s->h_light0_I[2] / fl) / 3.0f;
    launch_photon_scale(c->stream, d_map, cap, scale);
    cudaEventRecord(ev1, c->stream);
    e = cudaStreamSynchronize(c->stream);
    DCounters hc;
    if (e == cudaSuccess) e = cudaMemcpy(&hc, c->wb.counters, sizeof hc, cudaMemcpyDeviceToHost);
    float emit_ms = 0.f;
    cudaEventElapsedTime(&emit_ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (e != cudaSuccess) { cleanup(); CU(e); }
    auto t0 = std::chrono::steady_clock::now();
    rc = install_photon_map(s, d_map, nullptr, cap, pp);
    if (rc == RTU_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = RTU_ERR_CUDA;
    cleanup();
    if (rc) return rc;
    float build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (stats) {
        stats->paths = paths;
        stats->from_light = from_light;
        stats->stored = cap;
        stats->trace_rays = hc.k[0].trace_rays;
        stats->scale_factor = scale;
        stats->emit_ms = emit_ms;
        stats->build_ms = build_ms;
        stats->device_build = s->photon_device_build ? 1u : 0u;
    }
    return RTU_OK;
}

// The entry points above that build std containers, behind the exception barrier of the ABI (a bad_alloc becomes an error code)
int rtu_context_create(int32_t device, void *stream, rtu_context **out)
{
    return rtu::guarded("rtu_context_create", [&]() -> int { return context_create_impl(device, stream, out); });
}

int rtu_scene_upload(rtu_context *c, const rtu_scene_desc *d, rtu_scene **out)
{
    return rtu::guarded("rtu_scene_upload", [&]() -> int { return scene_upload_impl(c, d, out); });
}

int rtu_trace(rtu_scene *s, const rtu_ray *rays, int64_t n, rtu_hit *hits)
{
    return rtu::guarded("rtu_trace", [&]() -> int { return trace_impl(s, rays, n, hits); });
}

int rtu_shadow_trace(rtu_scene *s, const rtu_ray *rays, const float *t_max, int64_t n, uint8_t *occluded)
{
    return rtu::guarded("rtu_shadow_trace", [&]() -> int { return shadow_trace_impl(s, rays, t_max, n, occluded); });
}

int rtu_selftest_division(rtu_context *c, uint32_t numerators_per_divisor, uint64_t seed, uint64_t *tested, uint64_t *mismatches)
{
    return rtu::guarded("rtu_selftest_division", [&]() -> int { return selftest_division_impl(c, numerators_per_divisor, seed, tested, mismatches); });
}

int rtu_camera_rays(rtu_scene *s, const rtu_params *p, int32_t sample, rtu_ray *rays)
{
    return rtu::guarded("rtu_camera_rays", [&]() -> int { return camera_rays_impl(s, p, sample, rays); });
}

int rtu_shade(rtu_scene *s, const rtu_ray *rays, const rtu_hit *hits, int64_t n, int32_t bounces, float *rgb)
{
    return rtu::guarded("rtu_shade", [&]() -> int { return shade_impl(s, rays, hits, n, bounces, rgb); });
}

int rtu_photon_map_set(rtu_scene *s, const rtu_photon *photons, uint32_t n, const rtu_photon_params *params)
{
    return rtu::guarded("rtu_photon_map_set", [&]() -> int { return photon_map_set_impl(s, photons, n, params); });
}

int rtu_photon_map_get(rtu_scene *s, rtu_photon *out, uint32_t cap, uint32_t *n)
{
    return rtu::guarded("rtu_photon_map_get", [&]() -> int { return photon_map_get_impl(s, out, cap, n); });
}

int rtu_estimate_irradiance(rtu_scene *s, const float *pos, const float *normal, int64_t n, float radius, float ellipticity,
                            float *irrad, float *direction, int32_t *found)
{
    return rtu::guarded("rtu_estimate_irradiance", [&]() -> int { return estimate_irradiance_impl(s, pos, normal, n, radius, ellipticity, irrad, direction, found); });
}

int rtu_photon_map_generate(rtu_scene *s, const rtu_photon_params *params, rtu_photon_stats *stats)
{
    return rtu::guarded("rtu_photon_map_generate", [&]() -> int { return photon_map_generate_impl(s, params, stats); });
}

} // extern "C"
