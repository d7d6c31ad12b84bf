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

bool load_obj_mesh(const char *path, HostMesh *m, std::string *err, bool load_mtl, bool build_hierarchies)
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
    m->mtls.clear(); m->mcfc.clear();
    bool has_tex = false, has_nrm = false;
    // usemtl / mtllib state (cyTriMesh.h:326-358): materials in order of first use, per face the material current when the
    // face was read, per material its first face and how many faces were read while it was current
    struct MtlData { std::string name; uint32_t first_face = 0, face_count = 0; };
    std::vector<MtlData> mtl_data;
    std::vector<std::string> mtl_files;
    std::vector<int> face_mtl;
    int current_mtl = -1;
    auto mtl_index = [&](const char *name) { for (size_t i = 0; i < mtl_data.size(); i++) if (mtl_data[i].name == name) return (int)i; return -1; };
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
            const size_t faces_before = m->f.size() / 3;
            int corner = -1;
            bool inspace = true, negative = false;
            int type = 0;
            uint32_t index = 0;
            auto emit = [&]() {
                m->f.insert(m->f.end(), fv, fv + 3);
                if (has_tex) m->ft.insert(m->ft.end(), ftx, ftx + 3);
                if (has_nrm) m->fn.insert(m->fn.end(), fnr, fnr + 3);
                face_mtl.push_back(current_mtl);
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
            if (current_mtl >= 0) mtl_data[current_mtl].face_count += (uint32_t)(m->f.size() / 3 - faces_before);
        } else if (load_mtl) {
            // only when the node names no material in the XML: TriObj::Load(name, mtlName==NULL) (xmlload.cpp:204)
            if (lr.is_cmd("usemtl")) {                      // MtlList::CreateMtl(buffer.Data(7), _f.size())
                const char *name = lr.len > 7 ? lr.buf + 7 : "";
                if (name[0] == '\0') {
                    current_mtl = mtl_data.empty() ? -1 : 0; // CreateMtl("") returns 0
                } else {
                    int i = mtl_index(name);
                    if (i < 0) {
                        MtlData d;
                        d.name = name;
                        d.first_face = (uint32_t)(m->f.size() / 3);
                        mtl_data.push_back(d);
                        i = (int)mtl_data.size() - 1;
                    }
                    current_mtl = i;
                }
            }
            if (lr.is_cmd("mtllib") && lr.len > 7) mtl_files.push_back(lr.buf + 7);
        }
        if (lr.eof()) break;
    }
    if (m->f.empty()) { // cyTriMesh.h:455: nothing is allocated, the mesh stays empty
        m->v.clear(); m->vn.clear(); m->vt.clear();
        return true;
    }
    uint32_t nf = m->nf();
    if (!m->vt.empty()) m->ft.resize((size_t)nf * 3, 0); else m->ft.clear();
    if (!m->vn.empty()) m->fn.resize((size_t)nf * 3, 0); else m->fn.clear();
    if (!mtl_data.empty()) {
        // faces grouped by material, materials in order of first use, faces without a material last (cyTriMesh.h:468-493;
        // the scan of a material starts at its first face and ends after face_count matches)
        std::vector<uint32_t> f2, ft2, fn2;
        f2.reserve(m->f.size());
        auto take = [&](uint32_t i) {
            f2.insert(f2.end(), m->f.begin() + (size_t)i * 3, m->f.begin() + (size_t)i * 3 + 3);
            if (!m->ft.empty()) ft2.insert(ft2.end(), m->ft.begin() + (size_t)i * 3, m->ft.begin() + (size_t)i * 3 + 3);
            if (!m->fn.empty()) fn2.insert(fn2.end(), m->fn.begin() + (size_t)i * 3, m->fn.begin() + (size_t)i * 3 + 3);
        };
        m->mcfc.assign(mtl_data.size(), 0);
        for (size_t k = 0; k < mtl_data.size(); k++) {
            for (uint32_t i = mtl_data[k].first_face, j = 0; j < mtl_data[k].face_count && i < nf; i++)
                if (face_mtl[i] == (int)k) { take(i); j++; }
            m->mcfc[k] = (int)(f2.size() / 3);
        }
        if (f2.size() / 3 < nf)
            for (uint32_t i = 0; i < nf; i++)
                if (face_mtl[i] < 0) take(i);
        if (f2.size() == m->f.size()) { // (always, unless a material's faces were not all found by its scan)
            m->f.swap(f2);
            if (!m->ft.empty()) m->ft.swap(ft2);
            if (!m->fn.empty()) m->fn.swap(fn2);
        }
        // the .mtl libraries, looked up next to the OBJ (cyTriMesh.h:499-544)
        m->mtls.resize(mtl_data.size());
        std::string dir;
        {
            const char *e1 = strrchr(path, '\\'), *e2 = e1 ? e1 : strrchr(path, '/');
            if (e2) dir.assign(path, e2 - path + 1);
        }
        for (const std::string &lib : mtl_files) {
            FILE *mf = fopen((dir + lib).c_str(), "rb");
            if (!mf) continue; // "ERROR: Cannot open file": the materials keep their defaults
            std::vector<char> mt;
            char tmp[1 << 14];
            size_t n;
            while ((n = fread(tmp, 1, sizeof tmp, mf)) > 0) mt.insert(mt.end(), tmp, tmp + n);
            fclose(mf);
            LineReader ml(mt.data(), mt.data() + mt.size());
            int id = -1;
            auto float3 = [&](float out[3]) { // Buffer::ReadFloat3: one value fills all three
                out[0] = out[1] = out[2] = 0.f;
                if (ml.len < 2) return;
                int got = sscanf(ml.buf + 2, "%f %f %f", &out[0], &out[1], &out[2]);
                if (got == 1) out[2] = out[1] = out[0];
            };
            auto copy_from = [&](int start) { // Buffer::Copy: skip blanks from `start`
                if (ml.len < start) return std::string();
                const char *q = ml.buf + start;
                while (*q != '\0' && *q <= ' ') q++;
                return std::string(q);
            };
            while (ml.next()) {
                if (ml.is_cmd("newmtl")) {
                    id = ml.len > 7 ? mtl_index(ml.buf + 7) : -1;
                    if (id >= 0) m->mtls[id].name = copy_from(7);
                } else if (id >= 0) {
                    ObjMtl &M = m->mtls[id];
                    if (ml.is_cmd("Ka")) float3(M.Ka);
                    else if (ml.is_cmd("Kd")) float3(M.Kd);
                    else if (ml.is_cmd("Ks")) float3(M.Ks);
                    else if (ml.is_cmd("Tf")) float3(M.Tf);
                    else if (ml.is_cmd("Ns")) { if (ml.len >= 2) sscanf(ml.buf + 2, "%f", &M.Ns); }
                    else if (ml.is_cmd("Ni")) { if (ml.len >= 2) sscanf(ml.buf + 2, "%f", &M.Ni); }
                    else if (ml.is_cmd("illum")) { if (ml.len >= 5) sscanf(ml.buf + 5, "%d", &M.illum); }
                    else if (ml.is_cmd("map_Kd")) { M.map_Kd = copy_from(7); M.has_map_Kd = true; }
                    else if (ml.is_cmd("map_Ks")) { M.map_Ks = copy_from(7); M.has_map_Ks = true; }
                }
                if (ml.eof()) break;
            }
        }
    }
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
    m->device_bvh = !build_hierarchies;
    if (!build_hierarchies) return true; // RTU_LOAD_DEVICE_BVH: the device builds the only hierarchy at upload
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
