// OBJ loading and BVH construction for TriObj meshes.
//
// Both must reproduce the reference's results exactly, because hit ids are defined by them:
//  * the OBJ parser follows the tokenisation rules of cyTriMesh::LoadFromFileObj
//    (cyTriMesh.h:263-438): whitespace runs collapse to one space, '#' comments only at line
//    start, n-gons are fanned (v0, v[k-1], v[k]), '-' makes every index of that vertex relative;
//  * the BVH follows cyBVH::Build (cyBVH.h:122-142): split at the mid-point of the widest box
//    axis with the two fall-back axes (MeanSplit :295-328), forced halving above 8 elements
//    (:249-254), children stored adjacently, child1's subtree numbered before child2's
//    (ConvertTempData :281-291).  Here the tree is numbered in the same recursion that
//    splits it, so no temporary node tree is built.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hmath.h"
#include "host_scene.h"

namespace rtu {

namespace {

// One logical OBJ line, normalised like cyTriMesh's Buffer::ReadLine (cyTriMesh.h:278-302).
struct LineReader {
    const char *p, *end;
    char buf[1024];
    int len = 0;
    LineReader(const char *b, const char *e) : p(b), end(e) {}
    bool eof() const { return p >= end; }
    int get() { return p < end ? (unsigned char)*p++ : -1; }

    int next()
    {
        int c = get();
        while (c >= 0) {
            while (c >= 0 && isspace(c)) c = get();
            if (c == '#') {
                while (c >= 0 && c != '\n' && c != '\r' && c != '\0') c = get();
            } else {
                break;
            }
        }
        int i = 0;
        bool inspace = false;
        while (i < 1024 - 1) {
            if (c < 0 || c == '\n' || c == '\r' || c == '\0') break;
            if (isspace(c)) {
                inspace = true;
            } else {
                if (inspace) buf[i++] = ' ';
                inspace = false;
                buf[i++] = (char)c;
            }
            c = get();
        }
        buf[i] = '\0';
        len = i;
        return i;
    }
    bool is_cmd(const char *cmd) const
    {
        size_t n = strlen(cmd);
        if (strncmp(buf, cmd, n) != 0) return false;
        return buf[n] == '\0' || buf[n] == ' ';
    }
    // sscanf(data+2, "%f %f %f") with the target zeroed first (cyTriMesh.h:304)
    void read3(float out[3]) const
    {
        out[0] = out[1] = out[2] = 0.f;
        if (len < 2) return;
        const char *s = buf + 2;
        for (int k = 0; k < 3; k++) {
            char *e;
            float f = strtof(s, &e);
            if (e == s) return;
            out[k] = f;
            s = e;
        }
    }
};

} // namespace

bool load_obj_mesh(const char *path, HostMesh *m, std::string *err)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        if (err) *err = std::string("cannot open OBJ file ") + path;
        return false;
    }
    std::vector<char> text;
    {
        char tmp[1 << 16];
        size_t n;
        while ((n = fread(tmp, 1, sizeof tmp, fp)) > 0) text.insert(text.end(), tmp, tmp + n);
        fclose(fp);
    }
    m->name = path;
    m->v.clear(); m->vn.clear(); m->vt.clear(); m->f.clear(); m->fn.clear(); m->ft.clear();
    bool has_tex = false, has_nrm = false;
    LineReader lr(text.data(), text.data() + text.size());
    while (int rb = lr.next()) {
        float t[3];
        if (lr.is_cmd("v")) {
            lr.read3(t);
            m->v.insert(m->v.end(), t, t + 3);
        } else if (lr.is_cmd("vt")) {
            lr.read3(t);
            m->vt.insert(m->vt.end(), t, t + 3);
            has_tex = true;
        } else if (lr.is_cmd("vn")) {
            lr.read3(t);
            m->vn.insert(m->vn.end(), t, t + 3);
            has_nrm = true;
        } else if (lr.is_cmd("f")) {
            // corner state machine of cyTriMesh.h:379-438
            uint32_t fv[3] = {0, 0, 0}, ftx[3] = {0, 0, 0}, fnr[3] = {0, 0, 0};
            int corner = -1;
            bool inspace = true, negative = false;
            int type = 0;
            uint32_t index = 0;
            auto emit = [&]() {
                m->f.insert(m->f.end(), fv, fv + 3);
                if (has_tex) m->ft.insert(m->ft.end(), ftx, ftx + 3);
                if (has_nrm) m->fn.insert(m->fn.end(), fnr, fnr + 3);
            };
            for (int i = 2; i < rb; i++) {
                char ch = lr.buf[i];
                if (ch == ' ') { inspace = true; continue; }
                if (inspace) {
                    inspace = false;
                    negative = false;
                    type = 0;
                    index = 0;
                    if (corner < 2) {
                        corner++;
                    } else { // 4th, 5th ... corner: close the previous triangle, keep v0 and the last corner
                        emit();
                        fv[1] = fv[2]; ftx[1] = ftx[2]; fnr[1] = fnr[2];
                    }
                }
                if (ch == '/') { type++; index = 0; }
                if (ch == '-') negative = true;
                if (ch >= '0' && ch <= '9') {
                    index = index * 10 + (uint32_t)(ch - '0');
                    if (type == 0) fv[corner] = negative ? (uint32_t)(m->v.size() / 3) - index : index - 1;
                    else if (type == 1) { ftx[corner] = negative ? (uint32_t)(m->vt.size() / 3) - index : index - 1; has_tex = true; }
                    else if (type == 2) { fnr[corner] = negative ? (uint32_t)(m->vn.size() / 3) - index : index - 1; has_nrm = true; }
                }
            }
            emit();
        }
        // usemtl / mtllib: every node in the shipped scenes names its material in the XML, so
        // TriObj::Load is called with loadMtl=false (xmlload.cpp:204) and these are ignored.
        if (lr.eof()) break;
    }
    if (m->f.empty()) { // cyTriMesh.h:455: nothing is allocated, the mesh stays empty
        m->v.clear(); m->vn.clear(); m->vt.clear();
        return true;
    }
    uint32_t nf = m->nf();
    if (!m->vt.empty()) m->ft.resize((size_t)nf * 3, 0); else m->ft.clear();
    if (!m->vn.empty()) m->fn.resize((size_t)nf * 3, 0); else m->fn.clear();
    // validate indices so the device never reads out of bounds
    auto check = [&](const std::vector<uint32_t> &idx, size_t n, const char *what) {
        for (uint32_t i : idx)
            if (i >= n) { if (err) *err = std::string("OBJ ") + path + ": " + what + " index out of range"; return false; }
        return true;
    };
    if (!check(m->f, m->v.size() / 3, "vertex") || !check(m->ft, m->vt.size() / 3, "texture") ||
        !check(m->fn, m->vn.size() / 3, "normal"))
        return false;
    if (m->vn.empty()) compute_vertex_normals(m); // objects.h:56
    compute_bounds(m);                             // objects.h:57
    build_bvh(m->v.data(), m->f.data(), nf, 4, &m->bvh_boxes, &m->bvh_data, &m->bvh_elements); // objects.h:58
    build_occlusion_bvh(m->v.data(), m->f.data(), m->bvh_elements.data(), nf, &m->occ);       // any-hit hierarchy (not in the reference)
    return true;
}

// cyTriMesh::ComputeNormals (cyTriMesh.h:248-261)
void compute_vertex_normals(HostMesh *m)
{
    size_t nv = m->v.size() / 3;
    uint32_t nf = m->nf();
    std::vector<V3> acc(nv);
    const float *v = m->v.data();
    auto P = [&](uint32_t i) { return V3(v[i * 3], v[i * 3 + 1], v[i * 3 + 2]); };
    for (uint32_t i = 0; i < nf; i++) {
        uint32_t a = m->f[i * 3], b = m->f[i * 3 + 1], c = m->f[i * 3 + 2];
        V3 n = cross(P(b) - P(a), P(c) - P(a));
        acc[a] = acc[a] + n;
        acc[b] = acc[b] + n;
        acc[c] = acc[c] + n;
    }
    m->vn.resize(nv * 3);
    for (size_t i = 0; i < nv; i++) {
        V3 n = normalized(acc[i]);
        m->vn[i * 3] = n.x; m->vn[i * 3 + 1] = n.y; m->vn[i * 3 + 2] = n.z;
    }
    m->fn = m->f;
}

// cyTriMesh::ComputeBoundingBox (cyTriMesh.h:229-246)
void compute_bounds(HostMesh *m)
{
    size_t nv = m->v.size() / 3;
    if (nv == 0) {
        m->bound_min[0] = m->bound_min[1] = m->bound_min[2] = 1;
        m->bound_max[0] = m->bound_max[1] = m->bound_max[2] = 0;
        return;
    }
    for (int k = 0; k < 3; k++) m->bound_min[k] = m->bound_max[k] = m->v[k];
    for (size_t i = 1; i < nv; i++)
        for (int k = 0; k < 3; k++) {
            float c = m->v[i * 3 + k];
            if (m->bound_min[k] > c) m->bound_min[k] = c;
            if (m->bound_max[k] < c) m->bound_max[k] = c;
        }
}

namespace {

struct BvhBuilder {
    const float *v;
    const uint32_t *f;
    uint32_t max_leaf;
    std::vector<float> &boxes;
    std::vector<uint32_t> &data;
    std::vector<uint32_t> &elem;
    uint32_t next_free = 2; // node 0 unused, root = 1 (cyBVH.h:76,140)

    static const uint32_t kLeafBit = 0x80000000u;   // cyBVH.h:53
    static const int kOffsetBits = 28;              // 32-1-3 (cyBVH.h:56)
    static const uint32_t kHardMax = 8;             // CY_BVH_MAX_ELEMENT_COUNT

    void tri_box(uint32_t face, float b[6]) const // BVHTriMesh::GetElementBounds (cyBVH.h:356-368)
    {
        const float *p0 = v + 3 * (size_t)f[face * 3];
        b[0] = b[3] = p0[0]; b[1] = b[4] = p0[1]; b[2] = b[5] = p0[2];
        for (int j = 1; j < 3; j++) {
            const float *p = v + 3 * (size_t)f[face * 3 + j];
            for (int k = 0; k < 3; k++) {
                if (b[k] > p[k]) b[k] = p[k];
                if (b[k + 3] < p[k]) b[k + 3] = p[k];
            }
        }
    }
    float tri_center(uint32_t face, int dim) const // cyBVH.h:371-375
    {
        return (v[3 * (size_t)f[face * 3] + dim] + v[3 * (size_t)f[face * 3 + 1] + dim] + v[3 * (size_t)f[face * 3 + 2] + dim]) / 3.0f;
    }
    static void grow(float a[6], const float b[6]) // BVH::Box::operator+= (cyBVH.h:184)
    {
        for (int i = 0; i < 3; i++) {
            if (a[i] > b[i]) a[i] = b[i];
            if (a[i + 3] < b[i + 3]) a[i + 3] = b[i + 3];
        }
    }
    static void empty_box(float b[6])
    {
        b[0] = b[1] = b[2] = 1e30f;
        b[3] = b[4] = b[5] = -1e30f;
    }
    void ensure(uint32_t node)
    {
        if (data.size() <= node) {
            data.resize((size_t)node + 1, 0);
            boxes.resize(((size_t)node + 1) * 6, 0.f);
        }
    }

    // MeanSplit (cyBVH.h:295-328): returns the size of the first part, 0 if no split
    uint32_t mean_split(uint32_t *e, uint32_t count, const float box[6]) const
    {
        if (count <= max_leaf) return 0;
        float ext[3] = {box[3] - box[0], box[4] - box[1], box[5] - box[2]};
        int order[3];
        order[0] = ext[0] >= ext[1] ? (ext[0] >= ext[2] ? 0 : 2) : (ext[1] >= ext[2] ? 1 : 2);
        order[1] = (order[0] + 1) % 3;
        order[2] = (order[0] + 2) % 3;
        if (ext[order[1]] < ext[order[2]]) { int t = order[1]; order[1] = order[2]; order[2] = t; }
        for (int s = 0; s < 3; s++) {
            int dim = order[s];
            float mid = 0.5f * (box[dim] + box[dim + 3]);
            uint32_t lo = 0, hi = count;
            while (lo < hi) {
                if (tri_center(e[lo], dim) <= mid) {
                    lo++;
                } else {
                    hi--;
                    uint32_t t = e[lo]; e[lo] = e[hi]; e[hi] = t;
                }
            }
            if (lo < count && lo > 0) return lo;
        }
        return 0;
    }

    void build(uint32_t node, uint32_t offset, uint32_t count, const float box[6])
    {
        ensure(node);
        uint32_t *e = elem.data() + offset;
        uint32_t n1 = mean_split(e, count, box);
        if (n1 == 0 || n1 >= count) {
            if (count > kHardMax) {
                n1 = count / 2; // cyBVH.h:251-253
            } else {
                for (int k = 0; k < 6; k++) boxes[(size_t)node * 6 + k] = box[k];
                data[node] = (offset & ((1u << kOffsetBits) - 1)) | ((count - 1) << kOffsetBits) | kLeafBit;
                return;
            }
        }
        float b1[6], b2[6], t[6];
        empty_box(b1);
        empty_box(b2);
        for (uint32_t i = 0; i < n1; i++) { tri_box(e[i], t); grow(b1, t); }
        for (uint32_t i = n1; i < count; i++) { tri_box(e[i], t); grow(b2, t); }
        uint32_t c = next_free;
        next_free += 2;
        ensure(c + 1);
        for (int k = 0; k < 6; k++) boxes[(size_t)node * 6 + k] = box[k];
        data[node] = c & (kLeafBit - 1);
        build(c, offset, n1, b1);
        build(c + 1, offset + n1, count - n1, b2);
    }
};

} // namespace

void build_bvh(const float *v, const uint32_t *f, uint32_t nf, uint32_t max_per_leaf,
               std::vector<float> *boxes, std::vector<uint32_t> *data, std::vector<uint32_t> *elements)
{
    boxes->clear();
    data->clear();
    elements->clear();
    if (nf == 0) return; // cyBVH.h:125
    if (max_per_leaf > 8) max_per_leaf = 8;
    elements->resize(nf);
    for (uint32_t i = 0; i < nf; i++) (*elements)[i] = i;
    boxes->reserve((size_t)nf * 6);
    data->reserve(nf);
    BvhBuilder b{v, f, max_per_leaf, *boxes, *data, *elements};
    float root[6], t[6];
    BvhBuilder::empty_box(root);
    for (uint32_t i = 0; i < nf; i++) { b.tri_box(i, t); BvhBuilder::grow(root, t); }
    b.build(1, 0, nf, root);
}

} // namespace rtu
