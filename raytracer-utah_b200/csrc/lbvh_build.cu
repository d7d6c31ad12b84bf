// Device-side hierarchy builder: an LBVH over the triangles of a mesh, built entirely on the GPU and emitted
// in the layout the wave kernels walk (device_scene.h OccNode, 4-wide; leaves of up to 4 triangles).
//
// The reference builds its cyBVH on one host thread (cyBVH.h:122-142, 242-328: pointer-chasing TempNodes, seconds for a
// million triangles) and every answer of the default path is defined against that tree.  A mesh flagged
// RTU_MESH_DEVICE_BVH carries no cyBVH at all: this builder is its only hierarchy, both closest-hit and any-hit walks run
// on it with the exact triangle test, and the results are the reference's except where the reference's own tree decides -
// two triangles at exactly the same distance (here the lower face index wins instead of the first one visited).
//
// Pipeline (Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees and k-d trees"):
//   k_lbvh_morton   30-bit Morton code of every triangle's box centre inside the mesh's bound box
//   cub radix sort  (code, face) pairs - a library sort, it is plumbing of a build step, not a render kernel
//   k_lbvh_tree     one thread per internal node: its key range from the common-prefix function, the split, its children
//   k_lbvh_fit      bottom-up boxes: every leaf climbs, the second arrival at a node (atomic counter) merges and goes on
//   k_lbvh_wide     one thread per internal node: the 4-wide node made of its grandchildren (a child that is a single
//                   triangle or covers at most 4 of them stays a leaf); only nodes reachable from node 0 are ever read
//   k_lbvh_gather   triangle records in sorted order for the leaves (face index in the record's low bits)
#include <cub/device/device_radix_sort.cuh>

#include "rtu_objects.h"

namespace {

__device__ __forceinline__ unsigned expand10(unsigned v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// vertices of face f -> box
__device__ __forceinline__ void face_box(const float *v, const unsigned *f, unsigned face, float lo[3], float hi[3])
{
#pragma unroll
    for (int k = 0; k < 3; k++) { lo[k] = 3.0e38f; hi[k] = -3.0e38f; }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *p = v + (size_t)f[(size_t)face * 3 + c] * 3;
#pragma unroll
        for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], p[k]); hi[k] = fmaxf(hi[k], p[k]); }
    }
}

__global__ void k_lbvh_morton(const float *v, const unsigned *f, unsigned nf, float3 bmin, float3 inv_ext, unsigned *codes, unsigned *faces)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    float lo[3], hi[3];
    face_box(v, f, i, lo, hi);
    const float cx = (0.5f * (lo[0] + hi[0]) - bmin.x) * inv_ext.x, cy = (0.5f * (lo[1] + hi[1]) - bmin.y) * inv_ext.y,
                cz = (0.5f * (lo[2] + hi[2]) - bmin.z) * inv_ext.z;
    const unsigned x = (unsigned)fminf(fmaxf(cx * 1024.f, 0.f), 1023.f), y = (unsigned)fminf(fmaxf(cy * 1024.f, 0.f), 1023.f),
                   z = (unsigned)fminf(fmaxf(cz * 1024.f, 0.f), 1023.f);
    codes[i] = (expand10(x) << 2) | (expand10(y) << 1) | expand10(z);
    faces[i] = i;
}

// length of the common prefix of keys i and j (the index breaks ties between equal codes); -1 outside [0, n)
__device__ __forceinline__ int lcp(const unsigned *codes, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned a = codes[i], b = codes[j];
    if (a != b) return __clz(a ^ b);
    return 32 + __clz((unsigned)i ^ (unsigned)j);
}

// child encoding inside the builder: >= 0 internal node index, < 0 leaf: -(sorted position + 1)
__global__ void k_lbvh_tree(const unsigned *codes, int n, int2 *children, int2 *ranges, int *parent_of_internal, int *parent_of_leaf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = lcp(codes, n, i, i + 1) - lcp(codes, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lcp(codes, n, i, i - d);
    int lmax = 2;
    while (lcp(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lcp(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lcp(codes, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lcp(codes, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    int2 c;
    c.x = first == gamma ? -(gamma + 1) : gamma;
    c.y = last == gamma + 1 ? -(gamma + 2) : gamma + 1;
    children[i] = c;
    ranges[i] = make_int2(first, last);
    if (c.x >= 0) parent_of_internal[c.x] = i; else parent_of_leaf[gamma] = i;
    if (c.y >= 0) parent_of_internal[c.y] = i; else parent_of_leaf[gamma + 1] = i;
    if (i == 0) parent_of_internal[0] = -1;
}

struct Box6 { float lo[3], hi[3]; };

__global__ void k_lbvh_fit(const float *v, const unsigned *f, const unsigned *faces, int n, const int2 *children, const int *parent_of_internal,
                           const int *parent_of_leaf, Box6 *leaf_box, Box6 *node_box, unsigned *arrived)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Box6 b;
    face_box(v, f, faces[i], b.lo, b.hi);
    leaf_box[i] = b;
    if (n == 1) return;
    int node = parent_of_leaf[i];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&arrived[node], 1u) == 0u) return; // the first arrival leaves the node to its sibling's subtree
        const int2 c = children[node];
        // boxes written by other SMs: read around L1 (a neighbouring entry may have pulled a stale line in earlier)
        Box6 l, r;
        {
            const float *pl = reinterpret_cast<const float *>(c.x >= 0 ? &node_box[c.x] : &leaf_box[-c.x - 1]);
            const float *pr = reinterpret_cast<const float *>(c.y >= 0 ? &node_box[c.y] : &leaf_box[-c.y - 1]);
#pragma unroll
            for (int k = 0; k < 3; k++) { l.lo[k] = __ldcg(pl + k); l.hi[k] = __ldcg(pl + 3 + k); r.lo[k] = __ldcg(pr + k); r.hi[k] = __ldcg(pr + 3 + k); }
        }
        Box6 u;
#pragma unroll
        for (int k = 0; k < 3; k++) { u.lo[k] = fminf(l.lo[k], r.lo[k]); u.hi[k] = fmaxf(l.hi[k], r.hi[k]); }
        node_box[node] = u;
        node = parent_of_internal[node];
    }
}

// child word of a builder child for the wide layout: a subtree of at most 4 triangles becomes one leaf
__device__ __forceinline__ unsigned wide_word(int c, const int2 *ranges)
{
    if (c < 0) return 0x80000000u | (unsigned)(-c - 1);
    const int2 r = ranges[c];
    if (r.y - r.x + 1 <= 4) return 0x80000000u | ((unsigned)(r.y - r.x) << 28) | (unsigned)r.x;
    return (unsigned)c;
}

__device__ __forceinline__ Box6 child_box(int c, const Box6 *leaf_box, const Box6 *node_box) { return c >= 0 ? node_box[c] : leaf_box[-c - 1]; }

__global__ void k_lbvh_wide(int n, const int2 *children, const int2 *ranges, const Box6 *leaf_box, const Box6 *node_box, OccNode *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int kids[4];
    int nk = 0;
    const int2 c = children[i];
    for (int side = 0; side < 2; side++) {
        const int ch = side ? c.y : c.x;
        const bool expand = ch >= 0 && (ranges[ch].y - ranges[ch].x + 1) > 4;
        if (expand) { kids[nk++] = children[ch].x; kids[nk++] = children[ch].y; }
        else kids[nk++] = ch;
    }
    OccNode N;
    for (int k = 0; k < 4; k++) {
        if (k < nk) {
            const Box6 b = child_box(kids[k], leaf_box, node_box);
            N.lox[k] = b.lo[0]; N.loy[k] = b.lo[1]; N.loz[k] = b.lo[2];
            N.hix[k] = b.hi[0]; N.hiy[k] = b.hi[1]; N.hiz[k] = b.hi[2];
            N.child[k] = wide_word(kids[k], ranges);
        } else {
            N.lox[k] = N.loy[k] = N.loz[k] = 3.0e38f;
            N.hix[k] = N.hiy[k] = N.hiz[k] = -3.0e38f;
            N.child[k] = 0x7fffffffu;
        }
        N.pad[k] = 0;
    }
    out[i] = N;
}

__global__ void k_lbvh_gather(const TriRec *by_face, const unsigned *faces, unsigned nf, TriRec *sorted)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const unsigned face = faces[i];
    TriRec T = by_face[face];
    const unsigned fb = (__float_as_uint(T.fbits) & 0xc0000000u) | face; // projection axis | slot (= face index in this mode)
    T.fbits = __uint_as_float(fb);
    sorted[i] = T;
}

} // namespace

// by_face: the mesh's TriRec array in face order (device).  Outputs are stream-ordered allocations pushed onto `owned`.
int rtu_lbvh_build(cudaStream_t st, const float *d_v, const unsigned *d_f, unsigned nf, const float bmin[3], const float bmax[3],
                   const TriRec *by_face, std::vector<void *> &owned, const OccNode **nodes_out, const TriRec **tris_out, uint32_t *root_out,
                   float *build_ms)
{
    *nodes_out = nullptr;
    *tris_out = nullptr;
    *root_out = 0x7fffffffu;
    if (nf == 0) return RTU_OK;
    if (nf >= (1u << 27)) { rtu::set_error("device BVH: too many triangles"); return RTU_ERR_UNSUPPORTED; }
    const int n = (int)nf;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (build_ms) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    unsigned *codes = nullptr, *faces = nullptr, *codes2 = nullptr, *faces2 = nullptr, *arrived = nullptr;
    int2 *children = nullptr, *ranges = nullptr;
    int *par_int = nullptr, *par_leaf = nullptr;
    Box6 *leaf_box = nullptr, *node_box = nullptr;
    void *tmp = nullptr;
    OccNode *nodes = nullptr;
    TriRec *sorted = nullptr;
    std::vector<void *> scratch;
    auto alloc = [&](void **p, size_t bytes) { cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, st); if (e == cudaSuccess) scratch.push_back(*p); return e; };
    auto cleanup = [&]() { for (void *p : scratch) cudaFreeAsync(p, st); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); };
#define LB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rtu::set_error(std::string("device BVH build: ") + #call + ": " + cudaGetErrorString(e_)); cleanup(); return RTU_ERR_CUDA; } } while (0)
    LB(alloc((void **)&codes, sizeof(unsigned) * nf));
    LB(alloc((void **)&faces, sizeof(unsigned) * nf));
    LB(alloc((void **)&codes2, sizeof(unsigned) * nf));
    LB(alloc((void **)&faces2, sizeof(unsigned) * nf));
    LB(alloc((void **)&children, sizeof(int2) * nf));
    LB(alloc((void **)&ranges, sizeof(int2) * nf));
    LB(alloc((void **)&par_int, sizeof(int) * nf));
    LB(alloc((void **)&par_leaf, sizeof(int) * nf));
    LB(alloc((void **)&leaf_box, sizeof(Box6) * nf));
    LB(alloc((void **)&node_box, sizeof(Box6) * nf));
    LB(alloc((void **)&arrived, sizeof(unsigned) * nf));
    float3 lo = make_float3(bmin[0], bmin[1], bmin[2]);
    float3 inv = make_float3(bmax[0] > bmin[0] ? 1.f / (bmax[0] - bmin[0]) : 0.f, bmax[1] > bmin[1] ? 1.f / (bmax[1] - bmin[1]) : 0.f,
                             bmax[2] > bmin[2] ? 1.f / (bmax[2] - bmin[2]) : 0.f);
    const unsigned T = 256, G = (nf + T - 1) / T;
    size_t tmp_bytes = 0;
    LB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, codes, codes2, faces, faces2, n, 0, 30, st));
    LB(alloc(&tmp, tmp_bytes));
    LB(cudaMallocAsync((void **)&sorted, sizeof(TriRec) * nf, st));
    owned.push_back(sorted);
    if (nf > 4) {
        LB(cudaMallocAsync((void **)&nodes, sizeof(OccNode) * (nf - 1), st));
        owned.push_back(nodes);
    }
    // the build proper (the allocations above come from the stream-ordered pool; growing it the first time is the driver's time)
    if (build_ms) cudaEventRecord(e0, st);
    k_lbvh_morton<<<G, T, 0, st>>>(d_v, d_f, nf, lo, inv, codes, faces);
    LB(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, codes, codes2, faces, faces2, n, 0, 30, st));
    // leaves' triangle records in sorted order (kept: the walk reads them)
    k_lbvh_gather<<<G, T, 0, st>>>(by_face, faces2, nf, sorted);
    *tris_out = sorted;
    if (nf <= 4) {
        *root_out = 0x80000000u | ((nf - 1u) << 28); // one leaf
    } else {
        LB(cudaMemsetAsync(arrived, 0, sizeof(unsigned) * nf, st));
        k_lbvh_tree<<<G, T, 0, st>>>(codes2, n, children, ranges, par_int, par_leaf);
        k_lbvh_fit<<<G, T, 0, st>>>(d_v, d_f, faces2, n, children, par_int, par_leaf, leaf_box, node_box, arrived);
        k_lbvh_wide<<<G, T, 0, st>>>(n, children, ranges, leaf_box, node_box, nodes);
        *nodes_out = nodes;
        *root_out = 0u;
    }
    LB(cudaGetLastError());
    if (build_ms) {
        cudaEventRecord(e1, st);
        LB(cudaEventSynchronize(e1));
        cudaEventElapsedTime(build_ms, e0, e1);
    }
    cleanup();
#undef LB
    return RTU_OK;
}
