// sm_100a kernels of the photon path (reference: RenderFunctions.cpp:341-413,
// mtlFunctions.cpp:19-118, lightFunctions.cpp:19-25, cyPhotonMap.h:104-424).
//
//   k_photon_emit     one thread per photon PATH (numbered; path i owns Philox stream i): RandomPhoton, Trace, up to
//                     max_bounce x RandomPhotonBounce; the photons the path would store go to a per-path staging slot
//   k_photon_compact  staging -> map, at offsets from the scan of the per-path counts (path order = the order of the
//                     reference's sequential loop, so the map holds exactly the photons that loop would have stored)
//   k_photon_scale    ScalePhotonPowers
//   k_estimate        EstimateIrradiance<100> per query point (batched operator)
//   k_photon_shade    PhotonMapping(ray, hInfo) per primary hit of the hit queue (RTU_MODE_PHOTON)
//
// The gather walks the kd-tree in the reference's order (near child, far child if still in range, then the node
// itself) with the reference's 100-entry max-heap, so the photons it returns and the order they are summed in are
// the reference's: irradiance and direction are bit-identical to cyPhotonMap's on the same map.
#include <algorithm>
#include <initializer_list>
#include "rtu_internal.h"
#include "shade.cuh"
#include "camera.cuh"

#ifndef GATHER_MIN_CTAS
#define GATHER_MIN_CTAS 16 // resident 128-thread CTAs per SM the estimate kernels are compiled for
#endif

#ifndef GATHER_STEPS
#define GATHER_STEPS 8 // tree steps a lane may take before the warp turns to the heap updates
#endif

#define WAVE_THREADS_PHOTON 256

#define PHOTON_K 100 // photonSampleSize (RenderFunctions.cpp:33)

struct PhotonRec {
    float x, y, z, power;
    unsigned packed0; // color r,g,b, plane_dirz
    unsigned packed1; // dir_x | dir_y << 16
};

__device__ __forceinline__ PhotonRec load_photon(const rtu_photon *map, int index)
{
    const uint2 *q = reinterpret_cast<const uint2 *>(map + index);
    uint2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    PhotonRec p;
    p.x = __uint_as_float(a.x); p.y = __uint_as_float(a.y);
    p.z = __uint_as_float(b.x); p.power = __uint_as_float(b.y);
    p.packed0 = c.x; p.packed1 = c.y;
    return p;
}

// Photon::GetDirection (cyPhotonMap.h:158-181) with its `dirY-dirY` slip: z = isqrt(0x3FFF0001 - dirX^2)
__device__ __forceinline__ void photon_direction(const PhotonRec &p, float &dx, float &dy, float &dz)
{
    int ix = (int)(short)(p.packed1 & 0xffffu), iy = (int)(short)(p.packed1 >> 16);
    dx = (float)ix / 32767.0f;
    dy = (float)iy / 32767.0f;
    int xy2 = ix * ix + iy - iy;
    if (xy2 > 0x3FFF0001) xy2 = 0x3FFF0001;
    // the reference extracts floor(sqrt(0x3FFF0001 - xy2)) bit by bit (16 rounds); the same integer from one sqrtf
    // and a correction step (equal for all 32 768 reachable arguments, tests/test_host.py)
    const int v = 0x3FFF0001 - xy2;
    int z = (int)sqrtf((float)v);
    while (z * z > v) z--;
    while ((z + 1) * (z + 1) <= v) z++;
    dz = (float)z / 32767.0f;
    if ((p.packed0 >> 24) & 0x8u) dz = -dz;
}

// The reference's heap arrays dist2[] / index[] as one array of (distance, photon) pairs: the two children of a heap slot
// are 16 adjacent, aligned bytes, so a sift step is ONE 128-bit local-memory load instead of two dependent round trips
// (distances, then the index to move).  r2 = the current search radius^2 (the reference's maxDist2, its slot 0).
struct __align__(8) HeapEnt {
    float d2;
    int idx;
};
struct Gather {
    __align__(16) HeapEnt e[PHOTON_K + 2]; // slots 1..PHOTON_K; slot PHOTON_K+1 is only ever loaded, never used
    int found;
    float r2;
};
__device__ __forceinline__ void heap_children(const Gather &G, int j, HeapEnt &a, HeapEnt &b)
{
    const float4 v = *reinterpret_cast<const float4 *>(&G.e[j]); // j is even
    a.d2 = v.x; a.idx = __float_as_int(v.y);
    b.d2 = v.z; b.idx = __float_as_int(v.w);
}

// LocatePhotons's per-node part (cyPhotonMap.h:368-423), in two halves so that the lanes of a warp can run the second one
// together: gather_test() decides whether photon `index` enters the heap (and with which squared distance),
// gather_insert() is the heap update.
__device__ __forceinline__ bool gather_test(const rtu_photon *map, int index, float qx, float qy, float qz, bool has_n, float nx,
                                            float ny, float nz, float norm_scale, const Gather &G, float &dist2)
{
    const PhotonRec p = load_photon(map, index);
    float fx = p.x - qx, fy = p.y - qy, fz = p.z - qz;
    dist2 = dot3(fx, fy, fz, fx, fy, fz);
    if (!(dist2 < G.r2)) return false;
    if (has_n) {
        float dx, dy, dz;
        photon_direction(p, dx, dy, dz);
        if (dot3(dx, dy, dz, nx, ny, nz) >= 0.f) return false;
        if (norm_scale > 0.f) {
            float perp = dot3(fx, fy, fz, nx, ny, nz);
            float s = perp * norm_scale;
            fx = fx + nx * s; fy = fy + ny * s; fz = fz + nz * s;
            dist2 = dot3(fx, fy, fz, fx, fy, fz);
            if (dist2 >= G.r2) return false;
        }
    }
    return true;
}

__device__ __forceinline__ void gather_insert(int index, float dist2, Gather &G)
{
    if (G.found < PHOTON_K) {
        G.found++;
        G.e[G.found].d2 = dist2;
        G.e[G.found].idx = index;
        if (G.found == PHOTON_K) { // build the max-heap (:385-401)
            const int half = PHOTON_K >> 1;
            for (int k = half; k >= 1; k--) {
                int parent = k;
                const HeapEnt t = G.e[k];
                while (parent <= half) {
                    int j = parent + parent;
                    HeapEnt a, b;
                    heap_children(G, j, a, b);
                    if (j < PHOTON_K && a.d2 < b.d2) { j++; a = b; }
                    if (t.d2 >= a.d2) break;
                    G.e[parent] = a;
                    parent = j;
                }
                G.e[parent] = t;
            }
        }
    } else { // replace the farthest (:403-418)
        int parent = 1, j = 2;
        float top = dist2; // what ends up in slot 1
        while (j <= PHOTON_K) {
            HeapEnt a, b;
            heap_children(G, j, a, b);
            if (j < PHOTON_K && a.d2 < b.d2) { j++; a = b; }
            if (dist2 > a.d2) break;
            G.e[parent] = a;
            if (parent == 1) top = a.d2;
            parent = j;
            j <<= 1;
        }
        G.e[parent].idx = index;
        G.e[parent].d2 = dist2;
        G.r2 = top; // maxDist2 = dist2[1]
    }
}

// EstimateIrradiance<100>(irrad, direction, radius, pos, normal, ellipticity, FILTER_TYPE_CONSTANT) (:276-323)
__device__ __noinline__ void estimate_irradiance(const DPhotonMap &PM, float qx, float qy, float qz, bool has_n, float nx, float ny,
                                                 float nz, float radius, float norm_scale, Col &irrad, float &ox, float &oy,
                                                 float &oz, int &found)
{
    Gather G;
    G.found = 0;
    G.r2 = radius * radius;
    if (PM.n > 0) {
        // explicit form of the recursion: frame = node | state << 28; state 0 = entered, 1 = near child done, 2 = both done
        unsigned frame[32];
        float fdist[32];
        int top = 0;
        frame[0] = 1u;
        fdist[0] = 0.f;
        // Lanes walk until they hold a photon that enters the heap (at most GATHER_STEPS tree steps per round), then the
        // lanes that hold one update their heaps side by side; each lane still sees its photons in LocatePhotons order.
        bool pending = false;
        int pidx = 0;
        float pd2 = 0.f;
        while (top >= 0 || pending) {
#pragma unroll 1
            for (int step = 0; step < GATHER_STEPS && top >= 0 && !pending; step++) {
                unsigned f = frame[top];
                int index = (int)(f & 0x0fffffffu);
                unsigned state = f >> 28;
                if (state == 0) {
                    if (index < PM.half) {
                        const PhotonRec p = load_photon(PM.map, index);
                        unsigned axis = (p.packed0 >> 24) & 0x3u;
                        float dist = (axis == 0 ? qx : (axis == 1 ? qy : qz)) - (axis == 0 ? p.x : (axis == 1 ? p.y : p.z));
                        fdist[top] = dist;
                        frame[top] = (unsigned)index | (1u << 28);
                        int near_child = dist > 0 ? 2 * index + 1 : 2 * index;
                        top++;
                        frame[top] = (unsigned)near_child;
                        continue;
                    }
                    state = 2;
                }
                if (state == 1) {
                    float dist = fdist[top];
                    frame[top] = (unsigned)index | (2u << 28);
                    if (dist * dist < G.r2) {
                        int far_child = dist > 0 ? 2 * index : 2 * index + 1;
                        top++;
                        frame[top] = (unsigned)far_child;
                        continue;
                    }
                }
                pending = gather_test(PM.map, index, qx, qy, qz, has_n, nx, ny, nz, norm_scale, G, pd2);
                pidx = index;
                top--;
            }
            if (pending) {
                gather_insert(pidx, pd2, G);
                pending = false;
            }
        }
    }
    irrad = mk(0, 0, 0);
    ox = oy = oz = 0.f;
    for (int i = 1; i <= G.found; i++) {
        const PhotonRec p = load_photon(PM.map, G.e[i].idx);
        Col pw = mk((float)(p.packed0 & 0xffu) / 255.0f, (float)((p.packed0 >> 8) & 0xffu) / 255.0f, (float)((p.packed0 >> 16) & 0xffu) / 255.0f) * p.power;
        const float filter = 1.f;
        irrad = irrad + pw * filter;
        float dx, dy, dz;
        photon_direction(p, dx, dy, dz);
        float w = filter * p.power;
        ox = ox + dx * w; oy = oy + dy * w; oz = oz + dz * w;
    }
    if (G.found > 0) {
        float area = 3.14159274101257324f * G.r2; // (float)M_PI
        if (area > 0.f) {
            const float inv = 1.0f / area;
            irrad = irrad * inv;
        }
        norm3(ox, oy, oz);
    }
    found = G.found;
}

// The exact, one-lane-per-query form (walk + heap in local memory).  It serves the queries the two-phase path below hands
// back (candidate list full).
__global__ void __launch_bounds__(128, GATHER_MIN_CTAS)
k_estimate_list(DPhotonMap PM, const float *pos, const float *normal, const unsigned *ids, const unsigned *n_ids, float radius,
                float norm_scale, float *irrad, float *direction, int *found)
{
    const unsigned n = *n_ids;
    const bool has_n = normal != nullptr;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const size_t i = ids[k];
        Col e;
        float dx, dy, dz;
        int f;
        float nx = has_n ? normal[i * 3] : 0.f, ny = has_n ? normal[i * 3 + 1] : 0.f, nz = has_n ? normal[i * 3 + 2] : 0.f;
        estimate_irradiance(PM, pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2], has_n, nx, ny, nz, radius, norm_scale, e, dx, dy, dz, f);
        irrad[i * 3] = e.r; irrad[i * 3 + 1] = e.g; irrad[i * 3 + 2] = e.b;
        direction[i * 3] = dx; direction[i * 3 + 1] = dy; direction[i * 3 + 2] = dz;
        if (found) found[i] = f;
    }
}

// ------------------------------------------------------------------ walk records and per-photon tables (built once per map)
// k_knn_candidates walks a copy of the kd-tree whose nodes also know what lies below them: the bounding box of the positions
// and of the decoded directions (components in 1/127 steps, rounded outwards) of the node and everything the reference's
// walk can reach through it.  A subtree none of whose photons can pass the node tests (cyPhotonMap.h:368-383) is never
// entered; what the reference would have done inside it changes nothing (it inserts nothing there).  Three float4 per node:
//   (x, y, z, packed colour / plane bits)   (box min xyz, direction-box min as 3 signed bytes)   (box max xyz, direction-box max)
// Two tables hold what the estimate needs of a photon, computed with the reference's own expressions once instead of per use:
// its decoded direction (GetDirection) and its power as a colour (GetPower).
__device__ __forceinline__ unsigned pack_s8x3(int a, int b, int c) { return (unsigned)(a & 0xff) | ((unsigned)(b & 0xff) << 8) | ((unsigned)(c & 0xff) << 16); }
__device__ __forceinline__ int s8_of(unsigned w, int k) { return (int)(signed char)((w >> (8 * k)) & 0xffu); }

__global__ void k_knn_init(const rtu_photon *map, int n, float4 *nodes, float4 *dir, float4 *pw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i > n) return;
    const PhotonRec p = load_photon(map, i);
    float dx, dy, dz;
    photon_direction(p, dx, dy, dz);
    dir[i] = make_float4(dx, dy, dz, 0.f);
    const Col c = mk((float)(p.packed0 & 0xffu) / 255.0f, (float)((p.packed0 >> 8) & 0xffu) / 255.0f, (float)((p.packed0 >> 16) & 0xffu) / 255.0f) * p.power;
    pw[i] = make_float4(c.r, c.g, c.b, p.power);
    const int lx = (int)floorf(dx * 127.0f), ly = (int)floorf(dy * 127.0f), lz = (int)floorf(dz * 127.0f);
    const int hx = (int)ceilf(dx * 127.0f), hy = (int)ceilf(dy * 127.0f), hz = (int)ceilf(dz * 127.0f);
    nodes[3 * (size_t)i] = make_float4(p.x, p.y, p.z, __uint_as_float(p.packed0));
    nodes[3 * (size_t)i + 1] = make_float4(p.x, p.y, p.z, __uint_as_float(pack_s8x3(lx, ly, lz)));
    nodes[3 * (size_t)i + 2] = make_float4(p.x, p.y, p.z, __uint_as_float(pack_s8x3(hx, hy, hz)));
}

// nodes [first, last) of one tree level take in their children's boxes (deeper levels are done)
__global__ void k_knn_merge(float4 *nodes, int first, int last)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= last) return;
    float4 lo = nodes[3 * (size_t)i + 1], hi = nodes[3 * (size_t)i + 2];
    unsigned dl = __float_as_uint(lo.w), dh = __float_as_uint(hi.w);
    int l[3] = {s8_of(dl, 0), s8_of(dl, 1), s8_of(dl, 2)}, h[3] = {s8_of(dh, 0), s8_of(dh, 1), s8_of(dh, 2)};
    for (int c = 2 * i; c <= 2 * i + 1; c++) {
        const float4 cl = nodes[3 * (size_t)c + 1], ch = nodes[3 * (size_t)c + 2];
        lo.x = fminf(lo.x, cl.x); lo.y = fminf(lo.y, cl.y); lo.z = fminf(lo.z, cl.z);
        hi.x = fmaxf(hi.x, ch.x); hi.y = fmaxf(hi.y, ch.y); hi.z = fmaxf(hi.z, ch.z);
        const unsigned cdl = __float_as_uint(cl.w), cdh = __float_as_uint(ch.w);
        for (int k = 0; k < 3; k++) { l[k] = min(l[k], s8_of(cdl, k)); h[k] = max(h[k], s8_of(cdh, k)); }
    }
    lo.w = __uint_as_float(pack_s8x3(l[0], l[1], l[2]));
    hi.w = __uint_as_float(pack_s8x3(h[0], h[1], h[2]));
    nodes[3 * (size_t)i + 1] = lo;
    nodes[3 * (size_t)i + 2] = hi;
}

cudaError_t launch_knn_build(cudaStream_t st, const rtu_photon *map, int n, int half, float4 *nodes, float4 *dir, float4 *pw)
{
    if (n <= 0) return cudaSuccess;
    k_knn_init<<<(n + 255) / 256, 256, 0, st>>>(map, n, nodes, dir, pw);
    // only nodes below `half` have children the walk enters (cyPhotonMap.h:354); levels from the deepest such node up
    int top = 0;
    while ((2 << top) <= half - 1) top++; // level of node half - 1
    for (int lv = top; lv >= 0 && half > 1; lv--) {
        const int first = 1 << lv, last = std::min(2 << lv, half);
        if (last > first) k_knn_merge<<<(last - first + 255) / 256, 256, 0, st>>>(nodes, first, last);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------ EstimateIrradiance<100> in two phases
// LocatePhotons (cyPhotonMap.h:350-424) interleaves a kd-tree walk with updates of a 100-entry max-heap whose root is the
// search radius.  One lane per query doing both is bound by the heap (816 bytes of local memory per lane, every sift
// level a scattered 16-byte access through L1/L2) and by lanes waiting for the longest walk of their warp.  Here:
//
//   k_knn_candidates   persistent lanes, one query each, the next one as soon as theirs ends.  A lane walks the tree in
//                      the reference's order (near child, far child if in range, then the node) WITHOUT a heap: it keeps
//                      a 32-bin histogram (4 bins per octave of squared distance, 32 bytes of shared memory) of the
//                      photons it accepted, and prunes with the upper edge of the bin that holds the 100th smallest -
//                      a radius that is never below the reference's.  Its walk therefore contains the reference's walk
//                      as a subsequence, and every photon the reference inserts is written, in order, to the query's
//                      candidate list (distance, index); photons in between have distances >= the reference's radius
//                      at that moment.  The walk is stackless (children of i are 2i, 2i+1; the split plane is re-read
//                      on the way up).
//   k_knn_replay       one lane per query replays its list against the real heap: entries whose distance is not below
//                      the current radius are exactly those the reference never inserted; the others go through the
//                      reference's insert (append, heapify at 100, replace-the-root sift).  The heaps are columns of
//                      shared memory (bank = lane: conflict free), the lists are interleaved by lane so that a warp
//                      reads 256 contiguous bytes per step, and the 32 queries of a warp heapify at the same step
//                      (the first 100 list entries are always the first 100 inserts).  Then the sum over the heap in
//                      heap order like EstimateIrradiance: irradiance, direction, radius bit for bit.
//
// The histogram follows the heap's one irregularity: the 101st insert replaces the root even when it is farther than
// the root (the radius is still the caller's until then), so at that moment one photon leaves the top occupied bin.
#define KNN_CAP 1024u        // list entries per query; a longer list sends the query to k_estimate_list (none on Project13)
#define KNN_THREADS 256
#define KNN_CHUNK (1u << 20) // queries per pass (8 GB of lists)
#define KNN_REPLAY_WARPS 11
#ifndef KNN_WALK
#define KNN_WALK 2 // node visits per lane and round of k_knn_candidates
#endif
#ifndef KNN_BOX_SHIFT
#define KNN_BOX_SHIFT 0 // inner nodes below half >> this are box-tested on arrival
#endif
#ifndef KNN_CTAS
#define KNN_CTAS 5
#endif

__device__ __forceinline__ size_t knn_entry(unsigned q, unsigned i) { return ((size_t)(q >> 5) * KNN_CAP + i) * 32u + (q & 31u); }

__global__ void __launch_bounds__(KNN_THREADS, KNN_CTAS)
k_knn_candidates(DPhotonMap PM, const float *pos, const float *normal, const unsigned *n_ptr, unsigned n_mult, unsigned n_max,
                 unsigned q0, unsigned chunk, float radius, float norm_scale, uint2 *lists, unsigned *len, unsigned *work,
                 unsigned *fb_count, unsigned *fb_ids)
{
    __shared__ unsigned hist[8 * KNN_THREADS]; // word w of this lane's 32 byte counters: hist[w * KNN_THREADS + tid]
    const unsigned tid = threadIdx.x, lane = tid & 31u, FULL = 0xffffffffu;
    unsigned n = n_max;
    if (n_ptr) { const unsigned long long m = (unsigned long long)__ldg(n_ptr) * n_mult; if (m < n) n = (unsigned)m; }
    n = n > q0 ? n - q0 : 0u;
    if (n > chunk) n = chunk;
    const float r2_0 = radius * radius;
    const int key_top = (int)(__float_as_uint(r2_0) >> 21);
    const bool can_shrink = key_top >= 32 && key_top < (0x7f800000 >> 21);
    const bool has_n = normal != nullptr;
    const float kcoef = norm_scale > 0.f ? (2.f * norm_scale + norm_scale * norm_scale) * 0.9999f : 0.f;
    bool active = false, drained = false;
    unsigned q = 0, cur = 0, up = 0, cnt = 0, below = 0, flag = 0;
    int B = 31;
    float bound = r2_0, qx = 0.f, qy = 0.f, qz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
    unsigned cidx = 0;
    float4 cn0 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (;;) {
        const unsigned want = __ballot_sync(FULL, !active && !drained);
        if (want) {
            unsigned base = 0;
            const unsigned leader = __ffs(want) - 1;
            if (lane == leader) base = atomicAdd(work, (unsigned)__popc(want));
            base = __shfl_sync(FULL, base, leader);
            if (!active && !drained) {
                q = base + __popc(want & ((1u << lane) - 1u));
                if (q >= n) drained = true;
                else {
                    const size_t g = (size_t)q0 + q;
                    qx = pos[g * 3]; qy = pos[g * 3 + 1]; qz = pos[g * 3 + 2];
                    if (has_n) { nx = normal[g * 3]; ny = normal[g * 3 + 1]; nz = normal[g * 3 + 2]; }
                    cnt = 0; below = 0; flag = 0; B = 31; bound = r2_0; cur = 1; up = 0;
#pragma unroll
                    for (int w = 0; w < 8; w++) hist[w * KNN_THREADS + tid] = 0u;
                    if (PM.n <= 0 || !(qx == qx)) len[q] = 0u; // no map / a slot without a query (NaN)
                    else active = true;
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (__all_sync(FULL, drained)) break;
            continue;
        }
        // ---- walk: up to KNN_WALK node visits per lane, all lanes step together; a lane stops at a node that passes the
        // Euclidean test (the rest of that node's work is the second half of the round)
        bool pending = false;
        float fx = 0.f, fy = 0.f, fz = 0.f, d2 = 0.f;
#pragma unroll 1
        for (int it = 0; it < KNN_WALK; it++) {
            const bool go = active && !pending && cur != 0u;
            if (!__any_sync(FULL, go)) break;
            if (go) {
                const float4 *rec = PM.knn_nodes + 3 * (size_t)cur;
                // the last inner node this lane read is kept: coming back from a leaf to its parent needs no load
                float4 n0 = cn0;
                if (up == 0u || cur != cidx) n0 = __ldg(rec);
                if ((int)cur < PM.half) { cidx = cur; cn0 = n0; }
                bool process = true;
                if (up == 0u && (int)cur < (PM.half >> KNN_BOX_SHIFT)) { // arriving at an inner node from above: can anything at or below it pass the node tests?
                    // (a leaf's box is the photon itself: the node test below says the same)
                    const float4 lo = __ldg(rec + 1), hi = __ldg(rec + 2);
                    const float ax = lo.x - qx, ay = lo.y - qy, az = lo.z - qz, bx = hi.x - qx, by = hi.y - qy, bz = hi.z - qz;
                    // per component the photon's |p - q| is at least this, and rounding keeps the order: no margin needed
                    const float mx = fmaxf(fmaxf(ax, -bx), 0.f), my = fmaxf(fmaxf(ay, -by), 0.f), mz = fmaxf(fmaxf(az, -bz), 0.f);
                    float lb = dot3(mx, my, mz, mx, my, mz);
                    bool cull = false;
                    if (has_n) {
                        // the plane offset (p - q).n over the box, and the stretch it adds: |f + n s|^2 = |f|^2 + perp^2 (2 ns + ns^2 |n|^2)
                        const float t0 = nx * ax, t1 = nx * bx, t2 = ny * ay, t3 = ny * by, t4 = nz * az, t5 = nz * bz;
                        const float plo = (fminf(t0, t1) + fminf(t2, t3)) + fminf(t4, t5), phi = (fmaxf(t0, t1) + fmaxf(t2, t3)) + fmaxf(t4, t5);
                        const float mag = (fmaxf(fabsf(t0), fabsf(t1)) + fmaxf(fabsf(t2), fabsf(t3))) + fmaxf(fabsf(t4), fabsf(t5));
                        float pm = fmaxf(fmaxf(plo, -phi), 0.f) - 4e-7f * mag;
                        pm = fmaxf(pm, 0.f);
                        lb = lb + kcoef * (pm * pm);
                        // every direction in the subtree fails `direction . normal < 0`?
                        const unsigned dl = __float_as_uint(lo.w), dh = __float_as_uint(hi.w);
                        const float u0 = nx * (float)s8_of(dl, 0), u1 = nx * (float)s8_of(dh, 0);
                        const float u2 = ny * (float)s8_of(dl, 1), u3 = ny * (float)s8_of(dh, 1);
                        const float u4 = nz * (float)s8_of(dl, 2), u5 = nz * (float)s8_of(dh, 2);
                        cull = (fminf(u0, u1) + fminf(u2, u3)) + fminf(u4, u5) > 0.02f; // (in 1/127 units: > 1.6e-4)
                    }
                    cull = cull || lb > bound * 1.00001f;
                    if (cull) { up = cur; cur >>= 1; process = false; }
                }
                if (process && (up != 0u || (int)cur < PM.half)) {
                    const unsigned axis = (__float_as_uint(n0.w) >> 24) & 0x3u;
                    const float dist = (axis == 0 ? qx : (axis == 1 ? qy : qz)) - (axis == 0 ? n0.x : (axis == 1 ? n0.y : n0.z));
                    const unsigned near_child = dist > 0 ? 2u * cur + 1u : 2u * cur;
                    if (up == 0u) { cur = near_child; process = false; }
                    else if (up == near_child && dist * dist < bound) { cur = near_child ^ 1u; up = 0u; process = false; }
                }
                if (process) { // the node itself (cyPhotonMap.h:368-383) against `bound` instead of the heap's radius
                    fx = n0.x - qx; fy = n0.y - qy; fz = n0.z - qz;
                    d2 = dot3(fx, fy, fz, fx, fy, fz);
                    if (d2 < bound) pending = true;
                    else { up = cur; cur >>= 1; }
                }
            }
        }
        if (pending) {
            bool take = true;
            unsigned mark = 0u;
            if (has_n) {
                const float4 dv = __ldg(PM.knn_dir + cur); // Photon::GetDirection, decoded once per map
                if (dot3(dv.x, dv.y, dv.z, nx, ny, nz) >= 0.f) take = false;
                else if (norm_scale > 0.f) {
                    const float perp = dot3(fx, fy, fz, nx, ny, nz);
                    const float s = perp * norm_scale;
                    fx = fx + nx * s; fy = fy + ny * s; fz = fz + nz * s;
                    const float d2e = dot3(fx, fy, fz, fx, fy, fz);
                    if (d2e < d2) mark = 0x80000000u; // the replay sees d2e; it implies the Euclidean test only when d2e >= d2
                    d2 = d2e;
                    if (d2 >= bound) take = false;
                }
            }
            if (take) {
                if (cnt < KNN_CAP) lists[knn_entry(q, cnt)] = make_uint2(__float_as_uint(d2), cur | mark);
                else flag = 1u;
                int c = 31 - (key_top - (int)(__float_as_uint(d2) >> 21));
                if (c < 0) c = 0;
                if (cnt == 100u) { // the 101st insert evicts the root whatever its own distance
                    int t = 31;
                    for (; t > 0; t--) if ((hist[(t >> 2) * KNN_THREADS + tid] >> ((t & 3) * 8)) & 0xffu) break;
                    hist[(t >> 2) * KNN_THREADS + tid] -= 1u << ((t & 3) * 8);
                    if (t < B) below--;
                }
                cnt++;
                const unsigned hw = hist[(c >> 2) * KNN_THREADS + tid];
                if (c < B) { hist[(c >> 2) * KNN_THREADS + tid] = hw + (1u << ((c & 3) * 8)); below++; }
                else if (((hw >> ((c & 3) * 8)) & 0xffu) < 0xffu) hist[(c >> 2) * KNN_THREADS + tid] = hw + (1u << ((c & 3) * 8));
                if (can_shrink && cnt > 100u)
                    while (below >= 100u) {
                        B--;
                        below -= (hist[(B >> 2) * KNN_THREADS + tid] >> ((B & 3) * 8)) & 0xffu;
                        bound = __uint_as_float((unsigned)(key_top - 31 + B + 1) << 21);
                    }
            }
            up = cur;
            cur >>= 1;
        }
        if (active && cur == 0u) {
            len[q] = cnt | (flag << 31);
            if (flag) fb_ids[atomicAdd(fb_count, 1u)] = q0 + q;
            active = false;
        }
    }
}

// cyPhotonMap's insert on shared-memory columns: D[k * 32] / I[k * 32] = heap slot k + 1 of this lane.  The heap holds a
// photon as its position in the query's candidate list (16 bits: 6 bytes per entry, 11 warps per SM instead of 8).
struct ColumnHeap {
    float *D;
    unsigned short *I;
    __device__ __forceinline__ float &d(int slot) { return D[(slot - 1) * 32]; }
    __device__ __forceinline__ unsigned short &i(int slot) { return I[(slot - 1) * 32]; }
};

__global__ void __launch_bounds__(KNN_REPLAY_WARPS * 32, 1)
k_knn_replay(DPhotonMap PM, const float *pos, const unsigned *n_ptr, unsigned n_mult, unsigned n_max, unsigned q0, unsigned chunk, float radius, const uint2 *lists,
             const unsigned *len, unsigned *work, float *irrad, float *direction, int *found_out)
{
    extern __shared__ __align__(16) unsigned char knn_raw[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, FULL = 0xffffffffu;
    ColumnHeap Hp;
    Hp.D = reinterpret_cast<float *>(knn_raw + (size_t)warp * (PHOTON_K * 32 * 6)) + lane;
    Hp.I = reinterpret_cast<unsigned short *>(knn_raw + (size_t)warp * (PHOTON_K * 32 * 6) + PHOTON_K * 32 * 4) + lane;
    unsigned n = n_max;
    if (n_ptr) { const unsigned long long m = (unsigned long long)__ldg(n_ptr) * n_mult; if (m < n) n = (unsigned)m; }
    n = n > q0 ? n - q0 : 0u;
    if (n > chunk) n = chunk;
    const unsigned groups = (n + 31u) >> 5;
    for (;;) {
        unsigned g = 0;
        if (lane == 0) g = atomicAdd(work, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= groups) break;
        const unsigned q = g * 32u + lane;
        unsigned L = q < n ? len[q] : 0x80000000u;
        const bool skip = (L >> 31) != 0u; // handed to k_estimate_list (or past the end)
        L = skip ? 0u : L;
        const uint2 *row = lists + knn_entry(q, 0);
        int nfound = 0;
        float r2 = radius * radius;
        // the first 100 list entries are the first 100 inserts (the radius is still the caller's)
        const unsigned head = L < (unsigned)PHOTON_K ? L : (unsigned)PHOTON_K;
        const unsigned head_max = __reduce_max_sync(FULL, head);
        for (unsigned i = 0; i < head_max; i++)
            if (i < head) {
                const uint2 e = row[(size_t)i * 32u];
                Hp.d((int)i + 1) = __uint_as_float(e.x);
                Hp.i((int)i + 1) = (unsigned short)i; // (both tests were made against the caller's radius)
            }
        nfound = (int)head;
        if (head == (unsigned)PHOTON_K) { // build the max-heap (:385-401)
            const int half = PHOTON_K >> 1;
            for (int k = half; k >= 1; k--) {
                int parent = k;
                const float td = Hp.d(k);
                const unsigned short ti = Hp.i(k);
                while (parent <= half) {
                    int j = parent + parent;
                    float a = Hp.d(j);
                    if (j < PHOTON_K) { const float b = Hp.d(j + 1); if (a < b) { j++; a = b; } }
                    if (td >= a) break;
                    Hp.d(parent) = a;
                    Hp.i(parent) = Hp.i(j);
                    parent = j;
                }
                Hp.d(parent) = td;
                Hp.i(parent) = ti;
            }
        }
        const unsigned Lmax = __reduce_max_sync(FULL, L);
        uint2 e[4], en[4]; // the next four entries are in flight while these four are sifted
#pragma unroll
        for (unsigned u = 0; u < 4u; u++) en[u] = (PHOTON_K + u < L) ? row[(size_t)(PHOTON_K + u) * 32u] : make_uint2(0x7f800000u, 0u);
        for (unsigned i0 = PHOTON_K; i0 < Lmax; i0 += 4u) {
#pragma unroll
            for (unsigned u = 0; u < 4u; u++) {
                e[u] = en[u];
                en[u] = (i0 + 4u + u < L) ? row[(size_t)(i0 + 4u + u) * 32u] : make_uint2(0x7f800000u, 0u);
            }
#pragma unroll
            for (unsigned u = 0; u < 4u; u++) {
                const float dist2 = __uint_as_float(e[u].x);
                bool in = i0 + u < L && dist2 < r2;
                if (in && (e[u].y >> 31)) { // rounding put the stretched distance below the Euclidean one: the first test (:368) again
                    const PhotonRec p = load_photon(PM.map, (int)(e[u].y & 0x7fffffffu));
                    const size_t g3 = ((size_t)q0 + q) * 3;
                    const float fx = p.x - pos[g3], fy = p.y - pos[g3 + 1], fz = p.z - pos[g3 + 2];
                    in = dot3(fx, fy, fz, fx, fy, fz) < r2;
                }
                if (in) { // replace the farthest (:403-418)
                    int parent = 1, j = 2;
                    float top = dist2;
                    while (j <= PHOTON_K) {
                        float a = Hp.d(j);
                        if (j < PHOTON_K) { const float b = Hp.d(j + 1); if (a < b) { j++; a = b; } }
                        if (dist2 > a) break;
                        Hp.d(parent) = a;
                        Hp.i(parent) = Hp.i(j);
                        if (parent == 1) top = a;
                        parent = j;
                        j <<= 1;
                    }
                    Hp.d(parent) = dist2;
                    Hp.i(parent) = (unsigned short)(i0 + u);
                    r2 = top;
                }
            }
        }
        if (skip) continue;
        // EstimateIrradiance's sum over the heap, in heap order (:300-323)
        Col irr = mk(0, 0, 0);
        float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll 4
        for (int i = 1; i <= nfound; i++) {
            const int k = (int)(row[(size_t)Hp.i(i) * 32u].y & 0x7fffffffu);
            const float4 pwv = __ldg(PM.knn_pw + k), dv = __ldg(PM.knn_dir + k); // GetPower, GetDirection
            const float filter = 1.f;
            irr = irr + mk(pwv.x, pwv.y, pwv.z) * filter;
            const float w = filter * pwv.w;
            ox = ox + dv.x * w; oy = oy + dv.y * w; oz = oz + dv.z * w;
        }
        if (nfound > 0) {
            const float area = 3.14159274101257324f * r2;
            if (area > 0.f) {
                const float inv = 1.0f / area;
                irr = irr * inv;
            }
            norm3(ox, oy, oz);
        }
        const size_t o = ((size_t)q0 + q) * 3;
        irrad[o] = irr.r; irrad[o + 1] = irr.g; irrad[o + 2] = irr.b;
        direction[o] = ox; direction[o + 1] = oy; direction[o + 2] = oz;
        if (found_out) found_out[(size_t)q0 + q] = nfound;
    }
}

// PhotonMapping(ray, hInfo) (RenderFunctions.cpp:394-413) once the estimate is known: the estimate becomes a PhotonLight
// (lights.h:61-74: Illuminate = intensity, Direction = direction, not ambient, no shadow ray) and the hit is shaded with it
// alone, bounceCount 0.  A hit without photons in range has direction 0/0 = NaN and shades to NaN, as in the reference.
__device__ __forceinline__ Col photon_light_shade(const DScene &S, int material, int front, float u, float v, float w, float px,
                                                  float py, float pz, float hnx, float hny, float hnz, Col e, float dx, float dy,
                                                  float dz)
{
    norm3(dx, dy, dz); // PhotonLight::SetDirection normalises once more (lights.h:70)
    Col out = mk(0, 0, 0);
    if (material < 0) {
        out = mk(1, 1, 1);
    } else if (front) {
        const DMaterial &M = S.materials[material];
        Col Kd = texcolor_sample(S, M.diffuse, u, v, w);
        Col Ks = texcolor_sample(S, M.specular, u, v, w);
        float vx = S.cam_pos[0] - px, vy = S.cam_pos[1] - py, vz = S.cam_pos[2] - pz; // mtlFunctions.cpp:137
        norm3(vx, vy, vz);
        float lx = -dx, ly = -dy, lz = -dz;
        norm3(lx, ly, lz);
        float hx = vx + lx, hy = vy + ly, hz = vz + lz;
        norm3(hx, hy, hz);
        float ndl = dot3(hnx, hny, hnz, lx, ly, lz);
        float ndh = dot3(hnx, hny, hnz, hx, hy, hz);
        if (ndl < 0.f) ndl = 0.f;
        if (ndh < 0.f) ndh = 0.f;
        out = (e * ndl) * (Kd + Ks * powf(ndh, M.glossiness));
    }
    return out;
}

__device__ __forceinline__ void hit_of_queue(const DScene &S, const FrameSetup &F, int s0, const HitQueue &hq, unsigned h, HitRec &H,
                                             Ray &ray, int &pixel, int &x, int &y, int &s, int &W)
{
    float4 a = hq.a[h], b = hq.b[h];
    Best B;
    B.z = a.x; B.node = __float_as_int(a.y); B.front = __float_as_int(a.z); B.slot = __float_as_int(a.w);
    B.bc1 = b.x; B.bc2 = b.y; B.bc3 = b.z;
    unsigned idx = __float_as_uint(b.w);
    PrimaryMap pm; // the primary ray of work item idx (same mapping as k_extend<primary>)
    pm.init(F);
    pm.decode(idx, s0, s, x, y);
    W = pm.W;
    pixel = y * pm.W + x;
    ray = primary_ray(F, s, x, y, pixel);
    finalize_hit(S, ray, B, H);
}

// RTU_MODE_PHOTON: PhotonMapping per primary hit of the hit queue, in three steps: the hits' positions and normals become
// queries, the two-phase estimate runs over them, the hits are shaded with their estimates.
__global__ void __launch_bounds__(128)
k_photon_queries(DScene S, FrameSetup F, int s0, HitQueue hq, float *qpos, float *qnormal)
{
    unsigned total = *hq.count;
    if (total > hq.cap) total = hq.cap;
    unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= total) return;
    HitRec H;
    Ray ray;
    int pixel, x, y, s, W;
    hit_of_queue(S, F, s0, hq, h, H, ray, pixel, x, y, s, W);
    qpos[(size_t)h * 3] = H.px; qpos[(size_t)h * 3 + 1] = H.py; qpos[(size_t)h * 3 + 2] = H.pz;
    qnormal[(size_t)h * 3] = H.nx; qnormal[(size_t)h * 3 + 1] = H.ny; qnormal[(size_t)h * 3 + 2] = H.nz;
}

__global__ void __launch_bounds__(128)
k_photon_shade(DScene S, FrameSetup F, int s0, HitQueue hq, const float *irrad, const float *direction, float4 *accum)
{
    unsigned total = *hq.count;
    if (total > hq.cap) total = hq.cap;
    unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= total) return;
    HitRec H;
    Ray ray;
    int pixel, x, y, s, W;
    hit_of_queue(S, F, s0, hq, h, H, ray, pixel, x, y, s, W);
    const size_t o = (size_t)h * 3;
    Col out = photon_light_shade(S, H.material, H.front, H.u, H.v, H.w, H.px, H.py, H.pz, H.nx, H.ny, H.nz,
                                 mk(irrad[o], irrad[o + 1], irrad[o + 2]), direction[o], direction[o + 1], direction[o + 2]);
    float *acc = reinterpret_cast<float *>(accum + pixel);
    atomicAdd(acc, out.r);
    atomicAdd(acc + 1, out.g);
    atomicAdd(acc + 2, out.b);
}

// RTU_MODE_PHOTON_GATHER: MonteCarloPhoton(hInfo, x, y, 1) per primary hit (RenderFunctions.cpp:416-451), added to the
// Whitted radiance the other kernels produce.  Up to gi_bounces cosine samples of the hemisphere of the SAME first hit; the
// HitInfo of the sample rays is never reset, so a later sample only finds what is nearer than the previous sample's hit;
// a sample that finds nothing adds the background and ends the loop; the sum is divided by the samples taken.
// k_gather_trace traces the samples and turns every sample hit into a query (slot hit * gi_bounces + sample; unused slots
// hold NaN), the two-phase estimate runs over the slots, k_gather_sum shades the sample hits in order and adds the mean.
struct GatherPoint {
    float u, v, w;
    int material_front; // material << 1 | front
};

__global__ void __launch_bounds__(128, 4)
k_gather_trace(DScene S, FrameSetup F, int s0, HitQueue hq, float *qpos, float *qnormal, GatherPoint *pts, unsigned char *meta,
               DCounters *counters)
{
    Tally tl = {0, 0, 0, 0, 0};
    unsigned total = *hq.count;
    if (total > hq.cap) total = hq.cap;
    unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < total) {
        HitRec H;
        Ray ray;
        int pixel, x, y, s, W;
        hit_of_queue(S, F, s0, hq, h, H, ray, pixel, x, y, s, W);
        Rng rng;
        rng.key = F.seed; rng.pixel = 0x50474154u; rng.path = primary_path(pixel, s); rng.dim = 0;
        Best Bs;
        Bs.z = RTU_BIG; Bs.node = -1; Bs.front = 1; Bs.slot = 0; Bs.bc1 = Bs.bc2 = Bs.bc3 = 0.f;
        int actual = 0, missed = 0;
        const float qnan = __int_as_float(0x7fc00000);
        for (int bnc = 0; bnc < F.gi_bounces; bnc++) {
            const size_t o = ((size_t)h * F.gi_bounces + bnc) * 3;
            if (missed) { qpos[o] = qnan; continue; }
            Ray sr;
            sr.px = H.px; sr.py = H.py; sr.pz = H.pz;
            sample_hemi_cos(rng, H.nx, H.ny, H.nz, sr.dx, sr.dy, sr.dz);
            norm3(sr.dx, sr.dy, sr.dz);
            actual++;
            tl.trace++;
            if (scene_hit<false, true>(S, sr, Bs, tl, false)) {
                HitRec Hs;
                finalize_hit(S, sr, Bs, Hs);
                qpos[o] = Hs.px; qpos[o + 1] = Hs.py; qpos[o + 2] = Hs.pz;
                qnormal[o] = Hs.nx; qnormal[o + 1] = Hs.ny; qnormal[o + 2] = Hs.nz;
                GatherPoint gp;
                gp.u = Hs.u; gp.v = Hs.v; gp.w = Hs.w;
                gp.material_front = (Hs.material << 1) | (Hs.front ? 1 : 0);
                pts[(size_t)h * F.gi_bounces + bnc] = gp;
            } else {
                qpos[o] = qnan;
                missed = 1;
            }
        }
        meta[h] = (unsigned char)(actual | (missed << 7));
    }
    // the sample rays are booked with the secondary class
    DCounterBlock *c = &counters->k[1];
    unsigned t = tl.trace, bx = tl.box, r = tl.tri, nn = tl.node;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t += __shfl_xor_sync(0xffffffffu, t, o);
        bx += __shfl_xor_sync(0xffffffffu, bx, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if ((threadIdx.x & 31) == 0 && t) {
        atomicAdd(&c->trace_rays, (unsigned long long)t);
        atomicAdd(&c->box_tests, (unsigned long long)bx);
        atomicAdd(&c->tri_tests, (unsigned long long)r);
        atomicAdd(&c->node_visits, (unsigned long long)nn);
    }
}

__global__ void __launch_bounds__(128)
k_gather_sum(DScene S, FrameSetup F, int s0, HitQueue hq, const float *qpos, const float *qnormal, const GatherPoint *pts,
             const unsigned char *meta, const float *irrad, const float *direction, float4 *accum)
{
    unsigned total = *hq.count;
    if (total > hq.cap) total = hq.cap;
    unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= total) return;
    // the pixel of work item idx
    const unsigned idx = __float_as_uint(hq.b[h].w);
    PrimaryMap pm;
    pm.init(F);
    int s, x, y;
    pm.decode(idx, s0, s, x, y);
    const int pixel = y * pm.W + x;
    const int actual = meta[h] & 0x7f, missed = meta[h] >> 7;
    Col sum = mk(0, 0, 0);
    for (int bnc = 0; bnc < actual - missed; bnc++) {
        const size_t k = (size_t)h * F.gi_bounces + bnc, o = k * 3;
        const GatherPoint gp = pts[k];
        sum = sum + photon_light_shade(S, gp.material_front >> 1, gp.material_front & 1, gp.u, gp.v, gp.w, qpos[o], qpos[o + 1], qpos[o + 2],
                                       qnormal[o], qnormal[o + 1], qnormal[o + 2], mk(irrad[o], irrad[o + 1], irrad[o + 2]),
                                       direction[o], direction[o + 1], direction[o + 2]);
    }
    if (missed) sum = sum + background_sample(S, x, y, pm.W, F.cam.height);
    if (actual > 0) {
        float n = (float)actual;
        float *acc = reinterpret_cast<float *>(accum + pixel);
        atomicAdd(acc, sum.r / n);
        atomicAdd(acc + 1, sum.g / n);
        atomicAdd(acc + 2, sum.b / n);
    }
}

// ------------------------------------------------------------------ emission
// Photon::SetDirection / SetPower (cyPhotonMap.h:142-156)
__device__ __forceinline__ void encode_photon(rtu_photon *dst, float px, float py, float pz, float dx, float dy, float dz, Col c)
{
    float power = c.r;
    if (power < c.g) power = c.g;
    if (power < c.b) power = c.b;
    Col q = mk(c.r / power, c.g / power, c.b / power);
    int r = (int)(q.r * 255), g = (int)(q.g * 255), b = (int)(q.b * 255); // Color24::FloatToByte (cyColor.h:245-246)
    r = r < 0 ? 0 : (r > 255 ? 255 : r);
    g = g < 0 ? 0 : (g > 255 ? 255 : g);
    b = b < 0 ? 0 : (b > 255 ? 255 : b);
    int ix = (int)(dx * 32767.0f), iy = (int)(dy * 32767.0f); // short(dir * 0x7FFF)
    unsigned plane = dz > 0 ? 0u : 0x8u;
    uint2 *o = reinterpret_cast<uint2 *>(dst);
    o[0] = make_uint2(__float_as_uint(px), __float_as_uint(py));
    o[1] = make_uint2(__float_as_uint(pz), __float_as_uint(power));
    o[2] = make_uint2((unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16) | (plane << 24),
                      ((unsigned)ix & 0xffffu) | (((unsigned)iy & 0xffffu) << 16));
}

__device__ __forceinline__ float gray(Col c) { return ((c.r + c.g) + c.b) / 3.0f; } // Color::Gray (cyColor.h:83)

__global__ void __launch_bounds__(WAVE_THREADS_PHOTON, 2)
k_photon_emit(DScene S, unsigned long long path0, unsigned n_paths, int max_bounce, uint2 seed, int light, rtu_photon *staging,
              unsigned char *counts, DCounters *counters, unsigned *work)
{
    Tally tl = {0, 0, 0, 0, 0};
    const unsigned lane = threadIdx.x & 31u;
    const DLight L = S.lights[light];
    // Every lane carries one photon path and advances it by one segment per round: [bounce off the last hit] -> Trace ->
    // [store].  A lane whose path ended takes the next path number at the top of the round, so lanes stay busy although
    // path lengths differ (mean 2.4 segments, maximum max_bounce + 1), and the traversal code exists once.
    bool alive = false, drained = false, first = false;
    unsigned k = 0, stored = 0, flag = 0;
    int bounce = 0;
    Rng rng;
    Ray ray;
    Best B;
    HitRec H;
    Col outgoing = mk(0, 0, 0), incoming = mk(0, 0, 0);
    for (;;) {
        const unsigned want = __ballot_sync(0xffffffffu, !alive && !drained);
        if (want) {
            unsigned base = 0;
            const unsigned leader = __ffs(want) - 1;
            if (lane == leader) base = atomicAdd(work, (unsigned)__popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!alive && !drained) {
                k = base + __popc(want & ((1u << lane) - 1u));
                if (k >= n_paths) drained = true;
                else {
                    const unsigned long long path = path0 + k;
                    rng.key = seed; rng.pixel = 0x9407u ^ (unsigned)(path >> 32); rng.path = (unsigned)path; rng.dim = 0;
                    // PointLight::RandomPhoton (lightFunctions.cpp:19-25)
                    ray.px = L.v[0]; ray.py = L.v[1]; ray.pz = L.v[2];
                    sample_ball(rng, 1.0f, ray.dx, ray.dy, ray.dz);
                    norm3(ray.dx, ray.dy, ray.dz);
                    B.z = RTU_BIG; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f;
                    stored = 0; flag = 0; bounce = 0;
                    alive = true; first = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, alive)) break;
        const bool was_alive = alive;
        if (alive && !first) {
            // one iteration of the bounce loop (RenderFunctions.cpp:359-381) up to its Trace()
            incoming = outgoing;
            if (H.material < 0) alive = false; // the reference would dereference a NULL material
            else {
                const DMaterial &M = S.materials[H.material];
                // MtlBlinn::RandomPhotonBounce (mtlFunctions.cpp:19-118)
                Col Kd = texcolor_sample(S, M.diffuse, H.u, H.v, H.w);
                Col Ks = texcolor_sample(S, M.specular, H.u, H.v, H.w);
                Col Kt = texcolor_sample(S, M.refraction, H.u, H.v, H.w);
                float dG = gray(Kd), sG = gray(Ks), rG = gray(Kt);
                float sum = (dG + sG) + rG;
                dG = dG / sum; sG = sG / sum; rG = rG / sum;
                float4 u = rng.next4();
                float pick = u.x;
                Ray nr;
                nr.px = H.px; nr.py = H.py; nr.pz = H.pz;
                nr.dx = nr.dy = nr.dz = 0.f;
                if (pick > dG) {
                    if (pick > dG + sG) {
                        if (pick > (dG + sG) + rG) alive = false; // absorbed
                        else {
                            float ox, oy, oz;
                            sample_ball(rng, M.refr_gloss, ox, oy, oz);
                            float snx = ((H.px + H.nx) + ox) - H.px, sny = ((H.py + H.ny) + oy) - H.py, snz = ((H.pz + H.nz) + oz) - H.pz;
                            norm3(snx, sny, snz);
                            float cos1 = dot3(snx, sny, snz, -ray.dx, -ray.dy, -ray.dz);
                            float sin1 = (float)sqrt(1.0 - (double)cos1 * (double)cos1);
                            if (sin1 > 1) sin1 = 1.0f;
                            if (sin1 < -1) sin1 = -1.0f;
                            if (cos1 > 1) cos1 = 1.0f;
                            if (cos1 < -1) cos1 = -1.0f;
                            float n1 = M.ior, n2 = 1.0f;
                            if (H.front) { n1 = 1.0f; n2 = M.ior; }
                            float sin2 = (n1 / n2) * sin1;
                            float cos2 = sqrtf(1 - sin2 * sin2);
                            if (cos2 > 1) cos2 = 1.0f;
                            float cx = sny * (-ray.dz) - snz * (-ray.dy), cy = snz * (-ray.dx) - snx * (-ray.dz), cz = snx * (-ray.dy) - sny * (-ray.dx);
                            norm3(cx, cy, cz);
                            float svx = sny * cz - snz * cy, svy = snz * cx - snx * cz, svz = snx * cy - sny * cx;
                            norm3(svx, svy, svz);
                            nr.dx = (-snx) * cos2 + svx * sin2; nr.dy = (-sny) * cos2 + svy * sin2; nr.dz = (-snz) * cos2 + svz * sin2;
                            norm3(nr.dx, nr.dy, nr.dz);
                            float w = rG / 1.0f;
                            outgoing = outgoing * mk(Kt.r / w, Kt.g / w, Kt.b / w);
                        }
                    } else {
                        sample_ball(rng, 1.0f, nr.dx, nr.dy, nr.dz); // unnormalised point of the unit ball (SURVEY A-16)
                        float w = sG / 1.0f;
                        outgoing = outgoing * mk(Ks.r / w, Ks.g / w, Ks.b / w);
                    }
                } else {
                    sample_ball(rng, 1.0f, nr.dx, nr.dy, nr.dz);
                    float w = dG / 1.0f;
                    outgoing = outgoing * mk(Kd.r / w, Kd.g / w, Kd.b / w);
                }
                ray = nr;
            }
        }
        if (alive) {
            // Trace(r, &rootNode, hInfo); after the first segment into the SAME HitInfo: its z still holds the previous
            // segment's length, so only nearer hits are found (SURVEY A-16)
            tl.trace++;
            if (!scene_hit<false, true>(S, ray, B, tl, false)) alive = false;
            else {
                finalize_hit(S, ray, B, H); // serves the store below and the next round's bounce
                if (first) {
                    flag = 0x80u; // photonFromLight++ (RenderFunctions.cpp:357)
                    outgoing = mk(L.I[0], L.I[1], L.I[2]);
                    first = false;
                    if (max_bounce <= 0) alive = false;
                } else {
                    const int mtl = __ldg(&S.nodes[B.node].material);
                    if (mtl < 0) alive = false;
                    else {
                        const DMaterial &M2 = S.materials[mtl];
                        if (gray(mk(M2.diffuse.c[0], M2.diffuse.c[1], M2.diffuse.c[2])) > 0.f) { // IsPhotonSurface (materials.h:48)
                            float dx = ray.dx, dy = ray.dy, dz = ray.dz;
                            norm3(dx, dy, dz);
                            encode_photon(staging + (size_t)k * (size_t)max_bounce + stored, H.px, H.py, H.pz, dx, dy, dz, incoming);
                            stored++;
                        }
                        if (++bounce >= max_bounce) alive = false;
                    }
                }
            }
        }
        if (was_alive && !alive) counts[k] = (unsigned char)(stored | flag);
    }
    // emission rays are booked with the primary class
    DCounterBlock *c = &counters->k[0];
    unsigned t = tl.trace, b = tl.box, r = tl.tri, nn = tl.node;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t += __shfl_xor_sync(0xffffffffu, t, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if (lane == 0) {
        atomicAdd(&c->trace_rays, (unsigned long long)t);
        atomicAdd(&c->box_tests, (unsigned long long)b);
        atomicAdd(&c->tri_tests, (unsigned long long)r);
        atomicAdd(&c->node_visits, (unsigned long long)nn);
    }
}

// staging -> map: path k's photons go to offsets[k] .. ; photons past the map's capacity are dropped (AddPhoton
// returns false, cyPhotonMap.h:192) and paths at or after `cut` never ran in the sequential loop.
__global__ void k_photon_compact(const rtu_photon *staging, const unsigned char *counts, const unsigned *offsets, unsigned n_paths,
                                 unsigned cut, int max_bounce, rtu_photon *map1, unsigned cap)
{
    unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_paths || k >= cut) return;
    unsigned cnt = counts[k] & 0x7fu, off = offsets[k];
    for (unsigned j = 0; j < cnt; j++) {
        if (off + j >= cap) break;
        const uint2 *s = reinterpret_cast<const uint2 *>(staging + (size_t)k * (size_t)max_bounce + j);
        uint2 *d = reinterpret_cast<uint2 *>(map1 + off + j);
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
    }
}

__global__ void k_photon_scale(rtu_photon *map1, unsigned n, float scale)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) map1[i].power *= scale;
}

// ------------------------------------------------------------------ launch wrappers
static int knn_sm_count()
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        sms = n;
    }
    return sms;
}

// The two-phase estimate over queries [0, n) with n = min(*n_ptr * n_mult, n_max) (n_ptr == NULL: n_max), in passes of
// KNN_CHUNK queries.  Scratch comes from the device's stream-ordered pool and goes back to it behind the last kernel.
static cudaError_t run_estimate(cudaStream_t st, const DPhotonMap &PM, const float *pos, const float *normal, const unsigned *n_ptr,
                                unsigned n_mult, unsigned n_max, float radius, float norm_scale, float *irrad, float *direction,
                                int *found)
{
    if (n_max == 0) return cudaSuccess;
    static bool attr_set = false;
    const size_t replay_smem = (size_t)KNN_REPLAY_WARPS * 32 * PHOTON_K * 6;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_knn_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)replay_smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    // the candidate lists are the big allocation (8 KB per query of a chunk): smaller chunks when the device cannot spare it
    unsigned chunk = (unsigned)std::min<size_t>(n_max, KNN_CHUNK);
    uint2 *lists = nullptr;
    unsigned *len = nullptr, *ctr = nullptr, *fb_ids = nullptr;
    cudaError_t e;
    for (;;) {
        e = cudaMallocAsync((void **)&lists, (((size_t)chunk + 31) / 32) * 32 * (size_t)KNN_CAP * sizeof(uint2), st);
        if (e != cudaErrorMemoryAllocation || chunk <= 32768u) break;
        cudaGetLastError();
        lists = nullptr;
        chunk /= 2;
    }
    const unsigned chunks = (n_max + chunk - 1) / chunk;
    const size_t per_chunk = chunk;
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&len, per_chunk * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&ctr, (2 * (size_t)chunks + 1) * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&fb_ids, (size_t)n_max * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctr, 0, (2 * (size_t)chunks + 1) * sizeof(unsigned), st);
    if (e == cudaSuccess) {
        const int sms = knn_sm_count();
        unsigned *fb_count = ctr + 2 * (size_t)chunks;
        for (unsigned c = 0; c < chunks; c++) {
            const unsigned q0 = c * chunk;
            k_knn_candidates<<<sms * KNN_CTAS, KNN_THREADS, 0, st>>>(PM, pos, normal, n_ptr, n_mult, n_max, q0, chunk, radius, norm_scale, lists, len,
                                                             ctr + 2 * c, fb_count, fb_ids);
            k_knn_replay<<<sms, KNN_REPLAY_WARPS * 32, replay_smem, st>>>(PM, pos, n_ptr, n_mult, n_max, q0, chunk, radius, lists, len, ctr + 2 * c + 1,
                                                                         irrad, direction, found);
        }
        k_estimate_list<<<sms * 2, 128, 0, st>>>(PM, pos, normal, fb_ids, fb_count, radius, norm_scale, irrad, direction, found);
        e = cudaGetLastError();
    }
    for (void *q : {(void *)lists, (void *)len, (void *)ctr, (void *)fb_ids}) if (q) cudaFreeAsync(q, st);
    return e;
}

cudaError_t launch_estimate(cudaStream_t st, const DPhotonMap &PM, const float *pos, const float *normal, long long n, float radius,
                            float norm_scale, float *irrad, float *direction, int *found)
{
    if (n <= 0) return cudaSuccess;
    if (n > 0xffffffffll) return cudaErrorInvalidValue;
    return run_estimate(st, PM, pos, normal, nullptr, 1u, (unsigned)n, radius, norm_scale, irrad, direction, found);
}

cudaError_t launch_photon_shade(cudaStream_t st, const DScene &S, const FrameSetup &F, int s0, const WaveBuffers &B, unsigned max_hits,
                                const DPhotonMap &PM, float4 *accum)
{
    if (max_hits == 0) return cudaSuccess;
    float *buf = nullptr; // qpos | qnormal | irrad | direction
    const size_t n3 = (size_t)max_hits * 3;
    cudaError_t e = cudaMallocAsync((void **)&buf, 4 * n3 * sizeof(float), st);
    if (e != cudaSuccess) return e;
    k_photon_queries<<<(max_hits + 127) / 128, 128, 0, st>>>(S, F, s0, B.hits, buf, buf + n3);
    e = run_estimate(st, PM, buf, buf + n3, B.hits.count, 1u, max_hits, PM.radius, PM.norm_scale, buf + 2 * n3, buf + 3 * n3, nullptr);
    if (e == cudaSuccess) {
        k_photon_shade<<<(max_hits + 127) / 128, 128, 0, st>>>(S, F, s0, B.hits, buf + 2 * n3, buf + 3 * n3, accum);
        e = cudaGetLastError();
    }
    cudaFreeAsync(buf, st);
    return e;
}

cudaError_t launch_photon_gather(cudaStream_t st, const DScene &S, const FrameSetup &F, int s0, const WaveBuffers &B, unsigned max_hits,
                                 const DPhotonMap &PM, float4 *accum)
{
    if (max_hits == 0 || F.gi_bounces <= 0) return cudaSuccess;
    if (F.gi_bounces > 127 || (unsigned long long)max_hits * (unsigned)F.gi_bounces > 0xffffffffull) return cudaErrorInvalidValue;
    const unsigned nq = max_hits * (unsigned)F.gi_bounces;
    const size_t n3 = (size_t)nq * 3;
    float *buf = nullptr; // qpos | qnormal | irrad | direction
    GatherPoint *pts = nullptr;
    unsigned char *meta = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&buf, 4 * n3 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&pts, (size_t)nq * sizeof(GatherPoint), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&meta, max_hits, st);
    if (e == cudaSuccess) {
        k_gather_trace<<<(max_hits + 127) / 128, 128, 0, st>>>(S, F, s0, B.hits, buf, buf + n3, pts, meta, B.counters);
        e = run_estimate(st, PM, buf, buf + n3, B.hits.count, (unsigned)F.gi_bounces, nq, PM.radius, PM.norm_scale, buf + 2 * n3, buf + 3 * n3, nullptr);
    }
    if (e == cudaSuccess) {
        k_gather_sum<<<(max_hits + 127) / 128, 128, 0, st>>>(S, F, s0, B.hits, buf, buf + n3, pts, meta, buf + 2 * n3, buf + 3 * n3, accum);
        e = cudaGetLastError();
    }
    for (void *q : {(void *)buf, (void *)pts, (void *)meta}) if (q) cudaFreeAsync(q, st);
    return e;
}

void launch_photon_emit(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, unsigned long long path0, unsigned n_paths,
                        int max_bounce, uint2 seed, int light, rtu_photon *staging, unsigned char *counts, DCounters *counters,
                        unsigned *work_counter)
{
    static int occ = 0;
    if (occ == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_photon_emit, WAVE_THREADS_PHOTON, 0) != cudaSuccess || n < 1) n = 1;
        occ = n;
    }
    k_photon_emit<<<cfg.sm_count * occ, WAVE_THREADS_PHOTON, 0, st>>>(S, path0, n_paths, max_bounce, seed, light, staging, counts,
                                                                     counters, work_counter);
}

void launch_photon_compact(cudaStream_t st, const rtu_photon *staging, const unsigned char *counts, const unsigned *offsets,
                           unsigned n_paths, unsigned cut, int max_bounce, rtu_photon *map1, unsigned cap)
{
    if (n_paths == 0) return;
    k_photon_compact<<<(n_paths + 255) / 256, 256, 0, st>>>(staging, counts, offsets, n_paths, cut, max_bounce, map1, cap);
}

void launch_photon_scale(cudaStream_t st, rtu_photon *map1, unsigned n, float scale)
{
    if (n == 0) return;
    k_photon_scale<<<(n + 255) / 256, 256, 0, st>>>(map1, n, scale);
}
