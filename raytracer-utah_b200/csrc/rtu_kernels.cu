// sm_100a kernels of the render path: a wavefront of intersect / shade / any-hit kernels.
//
//   k_extend_pool<primary>   persistent threads; builds the camera rays of a chunk of (sample, pixel) work items on the
//                            fly and finds their closest hit (Trace).  32 rays step through the node list; rays that
//                            enter a mesh are parked and walked as one pool of (ray, node) items per warp.  Misses add
//                            the background in place, hits are compacted into the hit queue (32 B records).  Tiles
//                            that k_tile_mask found empty are not traced at all.
//   k_extend_pool<queue>     same for the rays of the previous wave's queue (reflection / refraction / Fresnel / GI
//                            rays); misses add the environment term their parent would have added
//   k_shade<..>              one MtlBlinn::Shade step per compacted hit: evaluates the hit record, appends shadow rays
//                            (radiance-if-unoccluded) and the next wave's secondary rays; GI records in RTU_MODE_PATH.
//                            One inlined shade_hit() site, CTA rounds of 256 hits behind a barrier (instruction cache),
//                            lights / materials in shared memory, 3 CTAs per SM
//   k_shadow_wave            any-hit (ShadowTrace) over the shadow queue with pooled mesh walks; adds the unoccluded
//                            contributions to the accumulator
//   k_extend / k_shadow_wave_simple      the same waves with one walk per lane (meshes beyond the pool's item encoding)
//   k_extend_top / k_shadow_wave_top     scenes with hundreds of objects: nodes nominated through a top-level hierarchy
//   k_primary_ids    pixel-centre visibility (ids + z)        k_trace_batch / k_shadow_batch / k_shade_first: the
//   k_gi_combine     folds the GI records into the pixels     batched operators; k_resolve / k_zminmax / k_zimage
//
// Traversal and shading are separate kernels so that the traversal kernels stay small (2 resident CTAs of 256 threads
// per SM, measured best) and shading runs on dense warps of hits only.
// Work distribution replaces PixelIterator's atomic ticket counter (PixelIterator.h:25-38): each warp takes 32
// consecutive tickets from a global counter until the wave is drained (256 per atomic in a primary wave whose tiles are
// mostly empty, 256 per CTA in k_shade).  Grids are SM-count multiples (148 x resident
// CTAs per SM, from the occupancy API); CTAs stay resident for the whole wave.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cooperative_groups.h>

#include "rtu_internal.h"
#include "shade.cuh"
#include "camera.cuh"

#define WAVE_THREADS 256
#define TICKET_BLOCK 256u // work items a warp takes from the global counter at a time (pooled kernels)
#ifndef EXT_BLOCKS
#define EXT_BLOCKS 2 // resident CTAs per SM the traversal kernels are compiled for (measured best of 2/3/4 on B200)
#endif

__device__ __forceinline__ void flush_tally(const Tally &tl, DCounters *cc, int cls)
{
    DCounterBlock *c = &cc->k[cls];
    unsigned t = tl.trace, s = tl.shadow, b = tl.box, r = tl.tri, n = tl.node;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t += __shfl_xor_sync(0xffffffffu, t, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (t) atomicAdd(&c->trace_rays, (unsigned long long)t);
        if (s) atomicAdd(&c->shadow_rays, (unsigned long long)s);
        if (b) atomicAdd(&c->box_tests, (unsigned long long)b);
        if (r) atomicAdd(&c->tri_tests, (unsigned long long)r);
        if (n) atomicAdd(&c->node_visits, (unsigned long long)n);
    }
}

// What happens to a ray once Trace() is over: hits are compacted into the hit queue, misses add what the recursion adds.
template <bool PRIMARY>
__device__ __forceinline__ void extend_finish(const DScene &S, const FrameSetup &F, const PrimaryMap &pm, const RayQueue &in,
                                              const AuxPool &inaux, const HitQueue &hq, float4 *accum, float4 *target,
                                              DCounters *counters, unsigned idx, const Ray &ray, int pixel, int px, int py,
                                              const Best &B)
{
    if (B.node >= 0) {
        unsigned slot = warp_alloc(hq.count, true);
        if (slot >= hq.cap) { counters->overflow = 1; return; }
        hq.a[slot] = make_float4(B.z, __int_as_float(B.node), __int_as_float(B.front), __int_as_float(B.slot));
        hq.b[slot] = make_float4(B.bc1, B.bc2, B.bc3, __uint_as_float(idx));
        return;
    }
    // what the recursion adds when Trace() misses
    Col c = mk(0, 0, 0);
    if (PRIMARY) {
        c = background_sample(S, px, py, pm.W, F.cam.height);                     // RenderFunctions.cpp:145
        accum_add(accum, pixel, c);
    } else {
        float4 w = in.w[idx];
        int slot = __float_as_int(in.o[idx].w);
        int kind, bounce, tree, gidepth, mtl;
        unpack_meta(__float_as_uint(in.d[idx].w), kind, bounce, tree, gidepth, mtl);
        int aux = __float_as_int(w.w);
        Col Wt = mk(w.x, w.y, w.z);
        if (kind == RK_GI) {
            // MonteCarlo(): the sample ray left the scene, c = environment (RenderFunctions.cpp:575);
            // slot = first entry of the GI record, its last entry holds (c, index of the vertex that was missed)
            Col e = environment_sample(S, ray.dx, ray.dy, ray.dz);
            target[slot + 2 * (F.gi_bounces + 1)] = make_float4(e.r, e.g, e.b, (float)gidepth);
            return;
        }
        if (kind == RK_REFRACT) {
            c = Wt * environment_sample(S, ray.dx, ray.dy, ray.dz);               // mtlFunctions.cpp:267
        } else if (kind == RK_REFLECT || kind == RK_FRESNEL) {
            Col wm = Wt;
            if (aux >= 0) { float4 a = inaux.a[aux]; wm = mk(a.x, a.y, a.z); }
            c = wm * environment_sample(S, ray.dx, ray.dy, ray.dz);               // :250, :289
        }
        accum_add(target, slot + tree, c); // environment terms do not scale with the ambient light: slot+1 in tree 1
    }
}

// A tile none of whose camera rays can reach any object (FrameSetup::tile_empty): no ray is built.  The work item of the
// chunk's FIRST sample does the tile's pixel for all `ns` samples of the chunk: it books what Trace() books for rays whose every
// node fails its bound-box gate and adds the background (RenderFunctions.cpp:145; a function of the pixel only) once per
// sample, in order, like the reference's sample loop; the items of the other samples return at once.  Nobody else touches the
// accumulator of such a pixel while the primary wave runs, so the sum is kept in registers (no atomics, one texture fetch).
__device__ __forceinline__ void primary_miss_fast(const DScene &S, const FrameSetup &F, const PrimaryMap &pm, float4 *accum, int x,
                                                  int y, int s, int s0, int ns, Tally &tl)
{
    if (s != s0) return;
    tl.trace += (unsigned)ns;
    tl.node += (unsigned)(ns * F.n_obj);
    tl.box += (unsigned)(ns * F.n_obj);
    const Col c = background_sample(S, x, y, pm.W, F.cam.height);
    // adaptive frames keep even and odd samples apart (FrameSetup::half_split): samples s0 .. s0+ns-1 alternate, so
    // ns - ns/2 of them go where s0's parity says and ns/2 to the other half
    const int n_same = F.half_split ? ns - ns / 2 : ns, n_other = F.half_split ? ns / 2 : 0;
    const int base = half_slot(F, y * pm.W + x, F.half_split ? s0 : 0);
    float4 a = accum[base];
#pragma unroll 1
    for (int i = 0; i < n_same; i++) {
        a.x += c.r; a.y += c.g; a.z += c.b;
    }
    accum[base] = a;
    if (n_other) {
        const int other = half_slot(F, y * pm.W + x, s0 + 1);
        float4 b = accum[other];
#pragma unroll 1
        for (int i = 0; i < n_other; i++) {
            b.x += c.r; b.y += c.g; b.z += c.b;
        }
        accum[other] = b;
    }
}

// ------------------------------------------------------------------ closest hit
// Plain version: every lane walks its own meshes (used when a mesh does not fit the pooled kernel's item encoding).
// (FAST: meshes through their 4-wide hierarchies, one lane per ray - the tail kernel below; the stand-alone kernel keeps
// the cyBVH walk whose counters book the reference's work)
template <bool PRIMARY, bool FAST>
__device__ __forceinline__ void extend_body(const DScene &S, const FrameSetup &F, int s0, int s1, const RayQueue &in, const AuxPool &inaux,
                                            const HitQueue &hq, float4 *accum, float4 *target, DCounters *counters, unsigned *work)
{
    // accum: the pixel accumulator (primary misses add the background there)
    // target: the array the rays' slots index: == accum for Whitted frames, the GI records in RTU_MODE_PATH
    Tally tl = {0, 0, 0, 0, 0};
    const unsigned lane = threadIdx.x & 31u;
    PrimaryMap pm;
    pm.init(F);
    unsigned total;
    if (PRIMARY) total = pm.perSample * (unsigned)(s1 - s0);
    else { total = *in.count; if (total > in.cap) total = in.cap; }

    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        unsigned idx = base + lane;
        if (idx >= total) continue;

        Ray ray;
        int pixel = 0, x = 0, y = 0;
        if (PRIMARY) {
            int s;
            if (!pm.decode(idx, s0, s, x, y)) continue;
            if (F.tile_done && F.tile_done[pm.tile_of(idx)]) continue; // adaptive sampling: the tile has converged
            if (F.tile_empty && F.tile_empty[pm.tile_of(idx)]) { primary_miss_fast(S, F, pm, accum, x, y, s, s0, s1 - s0, tl); continue; }
            pixel = y * pm.W + x;
            ray = primary_ray(F, s, x, y, pixel);
            pixel = half_slot(F, pixel, s); // from here on: the accumulator slot
        } else {
            float4 o = in.o[idx], d = in.d[idx];
            ray.px = o.x; ray.py = o.y; ray.pz = o.z;
            ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
        }
        Best B;
        B.z = RTU_BIG; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f;
        tl.trace++;
        scene_hit<false, FAST>(S, ray, B, tl, PRIMARY);

        extend_finish<PRIMARY>(S, F, pm, in, inaux, hq, accum, target, counters, idx, ray, pixel, x, y, B);
    }
    flush_tally(tl, counters, PRIMARY ? 0 : 1);
}

template <bool PRIMARY>
__global__ void __launch_bounds__(WAVE_THREADS, EXT_BLOCKS)
k_extend(DScene S, FrameSetup F, int s0, int s1, RayQueue in, AuxPool inaux, HitQueue hq, float4 *accum, float4 *target,
         DCounters *counters, unsigned *work)
{
    extend_body<PRIMARY, false>(S, F, s0, s1, in, inaux, hq, accum, target, counters, work);
}

// Pooled version (see k_shadow_wave below for the idea): a ray that reaches a mesh whose bound box it enters is
// parked; 32 parked rays are walked as ONE pool of (ray, node) items per warp, 32 items per iteration, leaves in a
// second pool.  Closest hit does not depend on the order in which boxes are opened - the reference never prunes by the
// current z (objFunctions.cpp:358-359), so the boxes and triangles tested are the same set, and so are the counters -
// EXCEPT when two triangles report exactly the same distance: the reference keeps the one it visits first.  Distances
// are merged with a 64-bit atomicMin on (z bits, triangle); an equal z from another triangle flags the ray, and a
// flagged ray is re-walked in the reference's order (bvh_walk) at the end of the batch.  After its mesh a ray resumes
// its node loop with the updated HitInfo, so the order of nodes (stale-z sphere returns, ties between objects) is the
// reference's too.
#ifndef XP_POOL
#define XP_POOL 512
#endif
#define XP_LEAF 192 // one iteration of the 4-wide walk can add 128 leaves while 31 wait
static_assert(XP_POOL >= 160, "an iteration of the 4-wide walk expands up to 32 items into 128");

struct XpWarp {
    float4 o[32];                  // mesh-local origin, z of the HitInfo when the mesh was entered
    float4 d[32];                  // mesh-local direction, InvDir::ok
    float4 y[32];                  // hoisted reciprocals
    float4 ci[32], cn[32], cf[32]; // OccRay: 1/d, near-plane offsets, far-plane offsets (OCC)
    const DMesh *mesh[32];
    const OccNode *nodes[32];      // OCC: the mesh's 4-wide hierarchy; tris = its triangle records in ITS leaf order
    const BvhPair *pairs[32];      // !OCC: the cyBVH; tris = the cyBVH's triangle records
    const TriRec *tris[32];
    unsigned long long zkey[32];   // (closest z so far) << 32 | triangle, merged with atomicMin
    unsigned pool[XP_POOL];
    unsigned leaf[XP_LEAF];
    unsigned tie;                  // bit s: slot s saw two triangles at the same z
};

__device__ __forceinline__ void tri_load(const TriRec *p, TriRec &T)
{
    const float4 *q = reinterpret_cast<const float4 *>(p);
    float4 x = __ldg(q), y = __ldg(q + 1), w4 = __ldg(q + 2);
    T.nx = x.x; T.ny = x.y; T.nz = x.z; T.ax = x.w;
    T.ay = y.x; T.az = y.y; T.area = y.z; T.fbits = y.w;
    T.cau = w4.x; T.cav = w4.y; T.bau = w4.z; T.bav = w4.w;
}

__device__ __forceinline__ bool ref_reaches_lane(const DMesh &M, unsigned slot, const Ray &r, bool inv_ok, const float4 yv, Tally &tl)
{
    InvDir I;
    I.yx = yv.x; I.yy = yv.y; I.yz = yv.z; I.ok = inv_ok;
    return ref_reaches(M, slot, r, I, tl);
}

// Rare paths of the pooled closest-hit kernel as real functions: the kernel is ~100 KB of code with them inlined at every use
// and a quarter of its stall samples were instruction fetch.
struct RefWalkArgs {
    float px, py, pz, dx, dy, dz, yx, yy, yz;
    int ok;
    float z, bc1, bc2, bc3;
    int front, slot;
};
static __device__ __noinline__ bool walk_reference_ni(const DMesh *M, RefWalkArgs *a, Tally *tl)
{
    Ray r;
    r.px = a->px; r.py = a->py; r.pz = a->pz; r.dx = a->dx; r.dy = a->dy; r.dz = a->dz;
    InvDir I;
    I.yx = a->yx; I.yy = a->yy; I.yz = a->yz; I.ok = a->ok != 0;
    return bvh_walk<false>(M->pairs, M->tris, M->root, r, I, a->z, a->front, a->slot, a->bc1, a->bc2, a->bc3, *tl);
}
static __device__ __noinline__ bool ref_reaches_ni(const DMesh *M, unsigned slot, const RefWalkArgs *a, Tally *tl)
{
    Ray r;
    r.px = a->px; r.py = a->py; r.pz = a->pz; r.dx = a->dx; r.dy = a->dy; r.dz = a->dz;
    InvDir I;
    I.yx = a->yx; I.yy = a->yy; I.yz = a->yz; I.ok = a->ok != 0;
    return ref_reaches(*M, slot, r, I, *tl);
}
// Light lists (host/light_mask.cpp): the triangles of one mask cell in the order of their least depth from the light, as (slot in
// DMesh::tris, depth) pairs.  The exact triangle test on those that begin before the ray's origin; an accepting one is
// confirmed like every candidate of the any-hit search (ref_reaches, else the exact walk of the cyBVH decides).
static __device__ __noinline__ bool light_list_occludes_ni(const uint32_t *items, unsigned n, float zcut, const DMesh *M, const RefWalkArgs *a,
                                                          float t_max, Tally *tl)
{
    Ray r;
    r.px = a->px; r.py = a->py; r.pz = a->pz; r.dx = a->dx; r.dy = a->dy; r.dz = a->dz;
    const uint2 *e = reinterpret_cast<const uint2 *>(items);
    if (n == 0u) return false;
    uint2 it = __ldg(e);
    if (__uint_as_float(it.y) > zcut) return false; // this one and all behind it begin beyond the origin
    const float4 *q = reinterpret_cast<const float4 *>(M->tris + it.x);
    float4 x = __ldg(q), y = __ldg(q + 1), w4 = __ldg(q + 2);
    bool have1 = n > 1u;
    uint2 it1 = have1 ? __ldg(e + 1) : make_uint2(0u, 0u);
    for (unsigned k = 0;; k++) {
        // entry -> record -> test is a chain of two dependent loads per triangle, and the scan is bound by their latency (L2: the
        // lists of a mask are 3 MB).  Two steps ahead: while triangle k is tested, the record of entry k+1 (whose entry came
        // with the last step) and entry k+2 are in flight.
        const bool more = have1 && !(__uint_as_float(it1.y) > zcut);
        float4 nx = x, ny = y, nw = w4;
        if (more) {
            const float4 *nq = reinterpret_cast<const float4 *>(M->tris + it1.x);
            nx = __ldg(nq); ny = __ldg(nq + 1); nw = __ldg(nq + 2);
        }
        const bool have2 = more && k + 2u < n;
        const uint2 it2 = have2 ? __ldg(e + k + 2u) : make_uint2(0u, 0u);
        TriRec T;
        T.nx = x.x; T.ny = x.y; T.nz = x.z; T.ax = x.w;
        T.ay = y.x; T.az = y.y; T.area = y.z; T.fbits = y.w;
        T.cau = w4.x; T.cav = w4.y; T.bau = w4.z; T.bav = w4.w;
        tl->tri++;
        float z = t_max, b1, b2, b3;
        int fr;
        if (tri_hit(T, r, z, fr, b1, b2, b3)) {
            InvDir I;
            I.yx = a->yx; I.yy = a->yy; I.yz = a->yz; I.ok = a->ok != 0;
            if (ref_reaches(*M, it.x, r, I, *tl)) return true;
            return bvh_walk_any_fallback(*M, r, I, t_max, *tl);
        }
        if (!more) return false;
        it = it1; x = nx; y = ny; w4 = nw;
        it1 = it2; have1 = have2;
    }
}
static __device__ __noinline__ bool occ_walk_ni(const DMesh *M, unsigned start, const RefWalkArgs *a, const OccRay *oc, float t_max, Tally *tl)
{
    Ray r;
    r.px = a->px; r.py = a->py; r.pz = a->pz; r.dx = a->dx; r.dy = a->dy; r.dz = a->dz;
    InvDir I;
    I.yx = a->yx; I.yy = a->yy; I.yz = a->yz; I.ok = a->ok != 0;
    return occ_walk(*M, start, r, I, *oc, t_max, *tl);
}
static __device__ __noinline__ void occ_walk_closest_ni(const DMesh *M, unsigned start, const RefWalkArgs *a, const OccRay *oc, float z_in,
                                                        unsigned long long *zkey, unsigned *tie, unsigned tie_bit, Tally *tl)
{
    Ray r;
    r.px = a->px; r.py = a->py; r.pz = a->pz; r.dx = a->dx; r.dy = a->dy; r.dz = a->dz;
    InvDir I;
    I.yx = a->yx; I.yy = a->yy; I.yz = a->yz; I.ok = a->ok != 0;
    occ_walk_closest(*M, start, r, I, *oc, z_in, zkey, tie, tie_bit, *tl);
}

template <bool PRIMARY, bool FLAT, bool OCC>
__global__ void __launch_bounds__(WAVE_THREADS, EXT_BLOCKS)
k_extend_pool(DScene S, FrameSetup F, int s0, int s1, RayQueue in, AuxPool inaux, HitQueue hq, float4 *accum, float4 *target,
              DCounters *counters, unsigned *work, float4 *park)
{
    extern __shared__ __align__(16) unsigned char xp_raw[];
    XpWarp &W = reinterpret_cast<XpWarp *>(xp_raw)[threadIdx.x >> 5];
    Tally tl = {0, 0, 0, 0, 0};
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u, FULL = 0xffffffffu, NONE = 0x7fffffffu;
    // this warp's lists of parked / resuming rays: 3 float4 per entry = (idx, node, z, hit node | front, tri, bc1, bc2 | bc3)
    float4 *jobs = park + (size_t)(blockIdx.x * (WAVE_THREADS / 32) + (threadIdx.x >> 5)) * (size_t)((XP_JOBS + XP_RES) * 3);
    float4 *res = jobs + XP_JOBS * 3;
    PrimaryMap pm;
    pm.init(F);
    unsigned total;
    if (PRIMARY) total = pm.perSample * (unsigned)(s1 - s0);
    else { total = *in.count; if (total > in.cap) total = in.cap; }
    unsigned njobs = 0, nres = 0; // warp-uniform
    unsigned t_next = 0, t_end = 0; // this warp's block of work items: large waves take TICKET_BLOCK items per atomic,
    unsigned t_block;               // small ones fewer, so that every resident warp still gets several blocks
    {
        // blocks only pay where most work items are nearly free (frames whose tiles are mostly empty, FrameSetup::n_empty_tiles);
        // heavy items are better handed out 32 at a time
        const unsigned want = (PRIMARY && F.n_empty_tiles && 2u * __ldg(F.n_empty_tiles) > F.n_tiles) ? 256u : 32u;
        const unsigned per_warp = total / (gridDim.x * (WAVE_THREADS / 32) * 8u);
        t_block = per_warp >= want ? want : (per_warp < 32u ? 32u : (per_warp & ~31u));
    }
    bool drained = false;
    // (see the ticket code below) no adaptive frame: its converged tiles are looked up per ticket
    bool sparse = PRIMARY && F.tile_empty && !F.tile_done && F.n_empty_tiles && 2u * __ldg(F.n_empty_tiles) > F.n_tiles;
    unsigned live = 0, live_base = 0, sparse_tickets = 8;
    if (sparse) { // about four tickets with work per atomic: more would leave the last warps with long tails
        const unsigned n_live = F.n_tiles - __ldg(F.n_empty_tiles);
        sparse_tickets = 8u * n_live < F.n_tiles ? 32u : (4u * n_live < F.n_tiles ? 16u : 8u);
        // a small wave (few samples) has less than a block per resident warp: the plain tickets serve it better
        const unsigned per_warp = (total >> 5) / (gridDim.x * (WAVE_THREADS / 32));
        if (per_warp < 4u * sparse_tickets) sparse = false;
    }
    DNode root;
    load_node(S.nodes, root);

    for (;;) {
        if (njobs >= 32u || (njobs > 0u && nres == 0u && drained)) {
            // ------------------------------------------------------------ one batch of mesh walks
            const unsigned take = njobs < 32u ? njobs : 32u;
            njobs -= take;
            unsigned rootw = NONE, idx = 0, node = 0;
            bool direct = false;
            Best B;
            B.z = RTU_BIG; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f;
            if (lane < take) {
                const float4 e0 = jobs[(njobs + lane) * 3], e1 = jobs[(njobs + lane) * 3 + 1], e2 = jobs[(njobs + lane) * 3 + 2];
                idx = __float_as_uint(e0.x); node = __float_as_uint(e0.y);
                B.z = e0.z; B.node = __float_as_int(e0.w);
                B.front = __float_as_int(e1.x); B.slot = __float_as_int(e1.y); B.bc1 = e1.z; B.bc2 = e1.w; B.bc3 = e2.x;
                Ray ray;
                if (PRIMARY) {
                    int s, x, y;
                    pm.decode(idx, s0, s, x, y);
                    ray = primary_ray(F, s, x, y, y * pm.W + x);
                } else {
                    float4 o = in.o[idx], d = in.d[idx];
                    ray.px = o.x; ray.py = o.y; ray.pz = o.z;
                    ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
                }
                DNode nd;
                load_node(S.nodes + node, nd);
                const Ray lr = FLAT ? to_node(nd.itm, nd.pos, to_node(root.itm, root.pos, ray)) : local_ray_of(S, (int)node, ray, nullptr);
                const DMesh &M = S.meshes[nd.mesh];
                const InvDir I = mesh_invdir(M, lr);
                W.o[lane] = make_float4(lr.px, lr.py, lr.pz, B.z);
                W.d[lane] = make_float4(lr.dx, lr.dy, lr.dz, I.ok ? 1.f : 0.f);
                W.y[lane] = make_float4(I.yx, I.yy, I.yz, 0.f);
                W.mesh[lane] = &M;
                W.zkey[lane] = ((unsigned long long)__float_as_uint(B.z) << 32) | 0xffffffffull;
                if constexpr (OCC) {
                    const OccRay oc = occ_setup(lr, M.occ_scale, B.z);
                    W.ci[lane] = make_float4(oc.ix, oc.iy, oc.iz, 0.f);
                    W.cn[lane] = make_float4(oc.nx, oc.ny, oc.nz, 0.f);
                    W.cf[lane] = make_float4(oc.fx, oc.fy, oc.fz, 0.f);
                    W.nodes[lane] = M.occ_nodes;
                    W.tris[lane] = M.occ_tris;
                    rootw = M.occ_root;
                    if (!oc.ok) { direct = true; rootw = NONE; } // non-finite / huge components: walked in the reference's order below
                } else {
                    W.pairs[lane] = M.pairs;
                    W.tris[lane] = M.tris;
                    rootw = M.root;
                }
            }
            unsigned bi = __ballot_sync(FULL, rootw < NONE), bl = __ballot_sync(FULL, rootw > NONE);
            if (rootw < NONE) W.pool[__popc(bi & lt)] = (lane << 27) | rootw;
            if (rootw > NONE) W.leaf[__popc(bl & lt)] = (lane << 27) | (((rootw >> 28) & 7u) << 24) | (rootw & 0x00ffffffu);
            unsigned pool_n = __popc(bi), leaf_n = __popc(bl);
            const unsigned dmask = __ballot_sync(FULL, direct);
            if (lane == 0) W.tie = dmask; // a ray that does not use the hierarchy takes the path of a tie: the reference-order walk
            __syncwarp();
            for (;;) {
                const bool do_leaf = leaf_n >= 32u || (pool_n == 0u && leaf_n > 0u);
                if (!do_leaf && pool_n == 0u) break;
                if (do_leaf) {
                    const unsigned n = leaf_n < 32u ? leaf_n : 32u;
                    leaf_n -= n;
                    if (lane < n) {
                        const unsigned it = W.leaf[leaf_n + lane], sl = it >> 27;
                        const float4 o = W.o[sl], d = W.d[sl];
                        Ray r;
                        r.px = o.x; r.py = o.y; r.pz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                        const TriRec *tris = W.tris[sl];
                        const unsigned first = it & 0x00ffffffu, cnt = ((it >> 24) & 7u) + 1u;
                        if constexpr (OCC) {
                            const float4 yv = W.y[sl];
                            InvDir I;
                            I.yx = yv.x; I.yy = yv.y; I.yz = yv.z; I.ok = d.w != 0.f;
                            const DMesh *M = W.mesh[sl];
                            for (unsigned i = 0; i < cnt; i++) occ_candidate_closest(*M, tris + first + i, r, I, o.w, &W.zkey[sl], &W.tie, 1u << sl, tl);
                        } else
                        for (unsigned i = 0; i < cnt; i++) {
                            TriRec T;
                            tri_load(tris + first + i, T);
                            tl.tri++;
                            // gate with the closest distance any lane has found so far (<=: equal distances are looked at below)
                            float z = __uint_as_float((unsigned)(*(volatile unsigned long long *)&W.zkey[sl] >> 32)), b1, b2, b3;
                            int fr;
                            if (tri_hit<true>(T, r, z, fr, b1, b2, b3) && z < o.w) {
                                const unsigned long long key = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned long long)(first + i);
                                const unsigned long long old = atomicMin(&W.zkey[sl], key);
                                if ((unsigned)(old >> 32) == __float_as_uint(z) && (unsigned)old != 0xffffffffu && (unsigned)old != first + i)
                                    atomicOr(&W.tie, 1u << sl);
                            }
                        }
                    }
                    __syncwarp();
                    continue;
                }
                if constexpr (OCC) {
                // one item = one 4-wide node; boxes entered beyond the best z so far are skipped
                const bool finish = pool_n > XP_POOL - 128u; // no room to expand 32 items: walk them to the end instead
                const unsigned n = pool_n < 32u ? pool_n : 32u;
                pool_n -= n;
                unsigned sl = 0, hit = 0;
                uint4 ch = make_uint4(NONE, NONE, NONE, NONE);
                if (lane < n) {
                    const unsigned it = W.pool[pool_n + lane];
                    sl = it >> 27;
                    const float4 ci = W.ci[sl], cn = W.cn[sl], cf = W.cf[sl];
                    OccRay oc;
                    oc.ix = ci.x; oc.iy = ci.y; oc.iz = ci.z;
                    oc.nx = cn.x; oc.ny = cn.y; oc.nz = cn.z;
                    oc.fx = cf.x; oc.fy = cf.y; oc.fz = cf.z;
                    oc.tlim = __uint_as_float((unsigned)(*(volatile unsigned long long *)&W.zkey[sl] >> 32));
                    oc.ok = true;
                    if (finish) {
                        const float4 o = W.o[sl], d = W.d[sl], yv = W.y[sl];
                        RefWalkArgs a;
                        a.px = o.x; a.py = o.y; a.pz = o.z; a.dx = d.x; a.dy = d.y; a.dz = d.z;
                        a.yx = yv.x; a.yy = yv.y; a.yz = yv.z; a.ok = d.w != 0.f;
                        Tally t2 = {0, 0, 0, 0, 0};
                        const OccRay oc2 = oc;
                        occ_walk_closest_ni(W.mesh[sl], it & 0x07ffffffu, &a, &oc2, o.w, &W.zkey[sl], &W.tie, 1u << sl, &t2);
                        tl.box += t2.box; tl.tri += t2.tri;
                    } else {
                        hit = occ_node(oc, W.nodes[sl] + (it & 0x07ffffffu), ch);
                        tl.box += 4;
                    }
                }
                const unsigned w0 = (hit & 1u) ? ch.x : NONE, w1 = (hit & 2u) ? ch.y : NONE, w2 = (hit & 4u) ? ch.z : NONE, w3 = (hit & 8u) ? ch.w : NONE;
                // positions from eight ballots (independent of each other) instead of a five-step shuffle scan (a dependent chain)
                const unsigned i0 = __ballot_sync(FULL, w0 < NONE), i1 = __ballot_sync(FULL, w1 < NONE), i2 = __ballot_sync(FULL, w2 < NONE), i3 = __ballot_sync(FULL, w3 < NONE);
                const unsigned l0 = __ballot_sync(FULL, w0 > NONE), l1 = __ballot_sync(FULL, w1 > NONE), l2 = __ballot_sync(FULL, w2 > NONE), l3 = __ballot_sync(FULL, w3 > NONE);
                unsigned pi = pool_n + __popc(i0 & lt) + __popc(i1 & lt) + __popc(i2 & lt) + __popc(i3 & lt);
                unsigned li = leaf_n + __popc(l0 & lt) + __popc(l1 & lt) + __popc(l2 & lt) + __popc(l3 & lt);
                const unsigned tot = (__popc(i0) + __popc(i1) + __popc(i2) + __popc(i3)) | ((__popc(l0) + __popc(l1) + __popc(l2) + __popc(l3)) << 16);
                const unsigned tag = sl << 27;
#define RTU_PUSH(WORD)                                                                                                           \
                if (WORD < NONE) { W.pool[pi++] = tag | WORD; }                                  \
                else if (WORD > NONE) { W.leaf[li++] = tag | (((WORD >> 28) & 7u) << 24) | (WORD & 0x00ffffffu); }
                RTU_PUSH(w3)
                RTU_PUSH(w2)
                RTU_PUSH(w1)
                RTU_PUSH(w0)
#undef RTU_PUSH
                pool_n += tot & 0xffffu;
                leaf_n += tot >> 16;
#ifdef RTU_DEBUG_BOUNDS
                if (pool_n > XP_POOL || leaf_n > XP_LEAF) counters->overflow = 0xBAD1;
#endif
                __syncwarp();
                } else {
                const bool finish = pool_n > XP_POOL - 64u; // no room to expand 32 items: walk them to the end instead
                const unsigned n = pool_n < 32u ? pool_n : 32u;
                pool_n -= n;
                unsigned c1 = NONE, c2 = NONE, sl = 0;
                if (lane < n) {
                    const unsigned it = W.pool[pool_n + lane];
                    sl = it >> 27;
                    const float4 o = W.o[sl], d = W.d[sl], yv = W.y[sl];
                    Ray r;
                    r.px = o.x; r.py = o.y; r.pz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                    InvDir I;
                    I.yx = yv.x; I.yy = yv.y; I.yz = yv.z; I.ok = d.w != 0.f;
                    const BvhPair *pairs = W.pairs[sl];
                    if (finish) {
                        float z = o.w, b1, b2, b3;
                        int fr, tslot = -1;
                        if (bvh_walk<false>(pairs, W.tris[sl], it & 0x07ffffffu, r, I, z, fr, tslot, b1, b2, b3, tl)) {
                            const unsigned long long key = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned long long)(unsigned)tslot;
                            const unsigned long long old = atomicMin(&W.zkey[sl], key);
                            if ((unsigned)(old >> 32) == __float_as_uint(z) && (unsigned)old != 0xffffffffu && (unsigned)old != (unsigned)tslot)
                                atomicOr(&W.tie, 1u << sl);
                        }
                    } else {
                        float4 a, b, c, dd;
                        load_pair(pairs + (it & 0x07ffffffu), a, b, c, dd);
                        float e1, e2;
                        bool h1 = slab_fast(r, I, a.x, a.y, a.z, a.w, b.x, b.y, RTU_BIG, e1);
                        bool h2 = slab_fast(r, I, b.z, b.w, c.x, c.y, c.z, c.w, RTU_BIG, e2);
                        tl.box += 2;
                        if (h1) c1 = __float_as_uint(dd.x);
                        if (h2) c2 = __float_as_uint(dd.y);
                    }
                }
                const unsigned n2 = __ballot_sync(FULL, c2 < NONE), n1 = __ballot_sync(FULL, c1 < NONE);
                const unsigned l2 = __ballot_sync(FULL, c2 > NONE), l1 = __ballot_sync(FULL, c1 > NONE);
                if (c2 < NONE) W.pool[pool_n + __popc(n2 & lt)] = (sl << 27) | c2;
                if (c1 < NONE) W.pool[pool_n + __popc(n2) + __popc(n1 & lt)] = (sl << 27) | c1;
                if (c2 > NONE) W.leaf[leaf_n + __popc(l2 & lt)] = (sl << 27) | (((c2 >> 28) & 7u) << 24) | (c2 & 0x00ffffffu);
                if (c1 > NONE) W.leaf[leaf_n + __popc(l2) + __popc(l1 & lt)] = (sl << 27) | (((c1 >> 28) & 7u) << 24) | (c1 & 0x00ffffffu);
                pool_n += __popc(n2) + __popc(n1);
                leaf_n += __popc(l2) + __popc(l1);
#ifdef RTU_DEBUG_BOUNDS
                if (pool_n > XP_POOL || leaf_n > XP_LEAF) counters->overflow = 0xBAD1;
#endif
                __syncwarp();
                }
            }
            // every slot's ray takes the closest triangle into its HitInfo and goes on with the node behind the mesh
            if (lane < take) {
                const unsigned long long key = *(volatile unsigned long long *)&W.zkey[lane];
                const float4 o = W.o[lane], d = W.d[lane];
                Ray r;
                r.px = o.x; r.py = o.y; r.pz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                const DMesh *M = W.mesh[lane];
                // two triangles at the same distance: the first one in the reference's visiting order wins (this walk repeats
                // tests that are already booked); the winner's leaf box rejects the ray in the reference's own test: the
                // reference-order walk decides.  (no cyBVH: the lower face index has won the atomicMin)
                const bool tied = ((*(volatile unsigned *)&W.tie >> lane) & 1u) && !M->no_ref;
                bool rewalk = tied;
                RefWalkArgs a;
                {
                    const float4 yv = W.y[lane];
                    a.px = r.px; a.py = r.py; a.pz = r.pz; a.dx = r.dx; a.dy = r.dy; a.dz = r.dz;
                    a.yx = yv.x; a.yy = yv.y; a.yz = yv.z; a.ok = d.w != 0.f;
                }
                // (the confirmation runs once per ray that found a triangle: inlined, a call here cost 8 % of the primary wave)
                if (!rewalk && (unsigned)key != 0xffffffffu && OCC) rewalk = !ref_reaches_lane(*M, (unsigned)key, r, d.w != 0.f, W.y[lane], tl);
                if (rewalk) {
                    a.z = B.z; a.front = B.front; a.slot = B.slot; a.bc1 = B.bc1; a.bc2 = B.bc2; a.bc3 = B.bc3;
                    Tally t2 = {0, 0, 0, 0, 0}; // (its address is taken: kept apart from the kernel's own tally)
                    if (walk_reference_ni(M, &a, &t2)) {
                        B.z = a.z; B.front = a.front; B.slot = a.slot; B.bc1 = a.bc1; B.bc2 = a.bc2; B.bc3 = a.bc3;
                        B.node = (int)node;
                    }
                    if (!tied) { tl.box += t2.box; tl.tri += t2.tri; }
                } else if ((unsigned)key != 0xffffffffu) {
                    TriRec T;
                    tri_load(M->tris + (unsigned)key, T); // the key carries the cyBVH slot in both walks
                    float z = RTU_BIG;
                    tri_hit(T, r, z, B.front, B.bc1, B.bc2, B.bc3); // front / barycentrics of the winner; z is the merged one
                    B.z = __uint_as_float((unsigned)(key >> 32));
                    B.slot = (int)(unsigned)key;
                    B.node = (int)node;
                }
                float4 *e = res + (nres + lane) * 3;
                e[0] = make_float4(__uint_as_float(idx), __uint_as_float(node + 1u), B.z, __int_as_float(B.node));
                e[1] = make_float4(__int_as_float(B.front), __int_as_float(B.slot), B.bc1, B.bc2);
                e[2] = make_float4(B.bc3, 0.f, 0.f, 0.f);
            }
            nres += take;
#ifdef RTU_DEBUG_BOUNDS
            if (nres > XP_RES) counters->overflow = 0xBAD2;
#endif
            __syncwarp();
            continue;
        }
        // -------------------------------------------------------------------- node loop of 32 rays
        if (nres == 0u && drained) break; // (njobs == 0 here)
        // resuming rays have short node loops left, fresh rays long ones: a pass takes 32 of one kind, not a mix
        const unsigned k = (nres >= 32u || drained) ? (nres < 32u ? nres : 32u) : 0u;
        unsigned fresh = (k == 0u && !drained) ? 32u : 0u;
        unsigned base = 0;
        if (PRIMARY && fresh && sparse) {
            // a frame of mostly empty tiles: 8 to 32 tickets (= tiles of one sample) per atomic; every lane looks at one ticket's
            // tile, the warp fills the empty tiles that fall to it (the tickets of the chunk's first sample) and then
            // takes the tickets that hold work one pass at a time
            while (live == 0u && !drained) {
                unsigned blk = 0;
                if (lane == 0) blk = atomicAdd(work, 32u * sparse_tickets);
                blk = __shfl_sync(FULL, blk, 0);
                if (blk >= total) { drained = true; break; }
                const unsigned mine = blk + 32u * lane;
                const bool valid = lane < sparse_tickets && mine < total;
                const bool empty = valid && F.tile_empty[pm.tile_of(mine)];
                live = __ballot_sync(FULL, valid && !empty);
                live_base = blk;
                unsigned fill = __ballot_sync(FULL, empty && mine < pm.perSample);
                while (fill) {
                    const unsigned b = __ffs(fill) - 1u;
                    fill &= fill - 1u;
                    int s, x, y;
                    if (pm.decode(blk + 32u * b + lane, s0, s, x, y)) primary_miss_fast(S, F, pm, accum, x, y, s, s0, s1 - s0, tl);
                }
            }
            if (live) {
                base = live_base + 32u * (__ffs(live) - 1u);
                live &= live - 1u;
            } else fresh = 0;
        } else if (fresh) {
            // work items come in blocks of TICKET_BLOCK (8 tiles / 8 x 32 queue entries) per atomic: a wave of mostly empty
            // tiles would otherwise spend its time waiting for the single work counter
            if (t_next >= t_end) {
                if (lane == 0) t_next = atomicAdd(work, t_block);
                t_next = __shfl_sync(FULL, t_next, 0);
                t_end = t_next + t_block;
            }
            base = t_next;
            t_next += 32u;
            if (base >= total) { drained = true; fresh = 0; }
        }
        unsigned idx = 0;
        int i0 = 1;
        bool have = false;
        Best B;
        B.z = RTU_BIG; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f;
        if (lane < k) {
            const float4 *e = res + (nres - k + lane) * 3;
            const float4 e0 = e[0], e1 = e[1], e2 = e[2];
            idx = __float_as_uint(e0.x); i0 = (int)__float_as_uint(e0.y);
            B.z = e0.z; B.node = __float_as_int(e0.w);
            B.front = __float_as_int(e1.x); B.slot = __float_as_int(e1.y); B.bc1 = e1.z; B.bc2 = e1.w; B.bc3 = e2.x;
            have = true;
        } else if (lane - k < fresh) {
            idx = base + (lane - k);
            have = idx < total;
        }
        nres -= k;
        Ray ray;
        int pixel = 0, x = 0, y = 0;
        if (have) {
            if (PRIMARY) {
                int s;
                have = pm.decode(idx, s0, s, x, y);
                pixel = y * pm.W + x;
                if (have && i0 == 1 && F.tile_done && F.tile_done[pm.tile_of(idx)]) have = false; // adaptive sampling: converged tile
                if (have && i0 == 1 && !sparse && F.tile_empty && F.tile_empty[pm.tile_of(idx)]) { // warp-uniform: a ticket is one tile
                    primary_miss_fast(S, F, pm, accum, x, y, s, s0, s1 - s0, tl);
                    have = false;
                }
                if (have) ray = primary_ray(F, s, x, y, pixel);
                pixel = half_slot(F, pixel, s); // from here on: the accumulator slot
            } else {
                float4 o = in.o[idx], d = in.d[idx];
                ray.px = o.x; ray.py = o.y; ray.pz = o.z;
                ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
            }
        }
        int parked = 0;
        if (have) {
            if (i0 == 1) tl.trace++;
            const Ray r0 = to_node(root.itm, root.pos, ray);
            const float dd = dot3(r0.dx, r0.dy, r0.dz, r0.dx, r0.dy, r0.dz);
            Ray lvl[FLAT ? 1 : RTU_MAX_DEPTH];
            if (!FLAT) {
                lvl[0] = r0;
                if (i0 > 1 && i0 < S.n_nodes) local_ray_of(S, __ldg(&S.nodes[i0].parent), ray, lvl);
            }
            for (int i = i0; i < S.n_nodes; i++) {
                DNode nd;
                Ray lr;
                if (FLAT) {
                    const float4 bs = __ldg(&S.bounds[i]);
                    if (bs.w < 0.f && bs.w > -1.5f) continue; // no object
                    if (bound_culled(bs, r0, dd)) { tl.node++; tl.box++; continue; }
                    load_node(S.nodes + i, nd);
                    lr = to_node(nd.itm, nd.pos, r0);
                } else {
                    load_node(S.nodes + i, nd);
                    lr = to_node(nd.itm, nd.pos, lvl[nd.depth - 1]);
                    lvl[nd.depth] = lr;
                    if (nd.kind == 0) continue;
                    if (bound_culled(__ldg(&S.bounds[i]), r0, dd)) { tl.node++; tl.box++; continue; }
                }
                if (nd.kind == 3) { // TriObj::IntersectRay up to its bound-box gate (objFunctions.cpp:337)
                    const DMesh &M = S.meshes[nd.mesh];
                    tl.node++;
                    if (M.empty) continue;
                    tl.box++;
                    const InvDir I = mesh_invdir(M, lr);
                    float te;
                    if (!slab_fast(lr, I, M.bmin[0], M.bmin[1], M.bmin[2], M.bmax[0], M.bmax[1], M.bmax[2], RTU_BIG, te)) continue;
                    if (PRIMARY && OCC && eye_mask_rejects(S, nd, lr)) continue; // inside the box, beside the mesh's silhouette from the eye
                    parked = i;
                    break;
                }
                sphere_or_plane_hit(nd, i, lr, B, tl);
            }
            if (!parked) extend_finish<PRIMARY>(S, F, pm, in, inaux, hq, accum, target, counters, idx, ray, pixel, x, y, B);
        }
        const unsigned m = __ballot_sync(FULL, parked != 0);
        if (parked) {
            float4 *e = jobs + (njobs + __popc(m & lt)) * 3;
            e[0] = make_float4(__uint_as_float(idx), __uint_as_float((unsigned)parked), B.z, __int_as_float(B.node));
            e[1] = make_float4(__int_as_float(B.front), __int_as_float(B.slot), B.bc1, B.bc2);
            e[2] = make_float4(B.bc3, 0.f, 0.f, 0.f);
        }
        njobs += __popc(m);
#ifdef RTU_DEBUG_BOUNDS
        if (njobs > XP_JOBS) counters->overflow = 0xBAD3;
#endif
        __syncwarp();
    }
    flush_tally(tl, counters, PRIMARY ? 0 : 1);
}

// ------------------------------------------------------------------ shade the compacted hits
#ifndef SHADE_BLOCKS
#define SHADE_BLOCKS 3 // resident CTAs per SM of the shading kernel (80 registers; measured best of 2/3/4 over Teapot, Project10, Project11)
#endif
#define SHADE_SMEM_LIGHTS 16     // scenes with no more lights / materials than this shade out of shared-memory copies
#define SHADE_SMEM_MATERIALS 32
template <bool PRIMARY>
__device__ __forceinline__ void shade_body(DScene S, const FrameSetup &F, int s0, const RayQueue &in, const AuxPool &inaux, const HitQueue &hq,
                                           const WaveOut &O, unsigned *work, unsigned *gi_count)
{
    PrimaryMap pm;
    pm.init(F);
    unsigned total = *hq.count;
    if (total > hq.cap) total = hq.cap;
    const bool path_mode = F.mode == RTU_MODE_PATH;
    const int gi_end = 2 * (F.gi_bounces + 1); // GI record: A_0, D_0, ..., A_K, D_K, End, pad (.x = pixel)
    if (PRIMARY && path_mode && blockIdx.x == 0 && threadIdx.x == 0) *gi_count = total; // one GI record per primary hit
    ShadeParams SP;
    SP.flags = F.flags;
    SP.seed = F.seed;
    // lights and materials are read all through the shading code: out of shared memory they never wait on L2
    __shared__ DLight s_lights[SHADE_SMEM_LIGHTS];
    __shared__ DMaterial s_materials[SHADE_SMEM_MATERIALS];
    if (S.n_lights <= SHADE_SMEM_LIGHTS) {
        const int words = S.n_lights * (int)(sizeof(DLight) / 4);
#pragma unroll 1
        for (int i = threadIdx.x; i < words; i += WAVE_THREADS) ((unsigned *)s_lights)[i] = ((const unsigned *)S.lights)[i];
        S.lights = s_lights;
    }
    if (S.n_materials <= SHADE_SMEM_MATERIALS) {
        const int words = S.n_materials * (int)(sizeof(DMaterial) / 4);
#pragma unroll 1
        for (int i = threadIdx.x; i < words; i += WAVE_THREADS) ((unsigned *)s_materials)[i] = ((const unsigned *)S.materials)[i];
        S.materials = s_materials;
    }
    // The shading code is long and nearly loop-free, so the kernel is bound by instruction fetch: the warps of a CTA
    // therefore take their hits together and start every round at a barrier, which keeps them within an instruction-cache
    // window of each other (one warp's misses are the others' hits).
    __shared__ unsigned cta_base;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) cta_base = atomicAdd(work, (unsigned)WAVE_THREADS);
        __syncthreads();
        if (cta_base >= total) break;
        const unsigned h = cta_base + threadIdx.x;
        if (h >= total) continue;
        float4 ha = hq.a[h], hb = hq.b[h];
        Best B;
        B.z = ha.x; B.node = __float_as_int(ha.y); B.front = __float_as_int(ha.z); B.slot = __float_as_int(ha.w);
        B.bc1 = hb.x; B.bc2 = hb.y; B.bc3 = hb.z;
        unsigned idx = __float_as_uint(hb.w);

        Ray ray;
        Col Wt;
        int pixel, kind, bounce, mtl, aux, tree = 0, gidepth = 0;
        unsigned path;
        if (PRIMARY) {
            int s, x, y;
            pm.decode(idx, s0, s, x, y);
            pixel = y * pm.W + x;
            ray = primary_ray(F, s, x, y, pixel);
            path = primary_path(pixel, s);
            pixel = half_slot(F, pixel, s); // from here on: the accumulator slot (adaptive frames: odd samples in the second half)
            Wt = mk(1.f, 1.f, 1.f);
            kind = RK_PRIMARY; bounce = F.shade_bounces; mtl = 0; aux = -1;
        } else {
            float4 o = in.o[idx], d = in.d[idx], w = in.w[idx];
            path = in.path[idx];
            ray.px = o.x; ray.py = o.y; ray.pz = o.z;
            ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
            pixel = __float_as_int(o.w);
            unpack_meta(__float_as_uint(d.w), kind, bounce, tree, gidepth, mtl);
            Wt = mk(w.x, w.y, w.z);
            aux = __float_as_int(w.w);
        }
        HitRec H;
        finalize_hit(S, ray, B, H);
        // shade_hit() is inlined exactly once: a GI vertex runs it twice (t = 0, 1), every other ray kind once, all lanes of
        // the warp together.  (Three inlined copies made the kernel ~300 KB of SASS and bound by instruction fetch.)
        const bool gi = path_mode && (kind == RK_PRIMARY || kind == RK_GI);
        int n_shade = 1, k = 0, rec = 0;
        Col Ws = Wt;
        if (gi) {
            // A vertex h_k of the GI chain (RenderFunctions.cpp:129-135 for the camera hit, :565-570 for a
            // MonteCarlo() sample hit).  L(h_k) = Shade(h_k, lights) + Shade(h_k, {Ambient c_k}) where c_k is what
            // MonteCarlo(h_k) returns.  The second Shade is linear in c_k, so its c_k-factor A_k and everything that
            // does not depend on c_k (D_k) are accumulated in two slots of the sample's GI record and folded by
            // k_gi_combine once all waves are done: L_k = D_k + A_k * L_{k+1}.
            k = kind == RK_PRIMARY ? 0 : gidepth;
            if (kind == RK_PRIMARY) {
                rec = (int)h * (gi_end + 2); // stride kept even: a vertex's two slots are one aligned 32-byte sector
                // no zero fill: every vertex stores both of its slots when it is shaded (fresh_slot), k_gi_combine reads the
                // slots of shaded vertices only, and the End slot is always written (terminal vertex or missed sample ray)
                O.accum[rec + gi_end + 1] = make_float4(__int_as_float(pixel), 0.f, 0.f, 0.f); // the record's pixel, in its pad slot
            } else {
                rec = pixel;
            }
            n_shade = 2;
        } else if (kind == RK_REFRACT) {
            float4 a = inaux.a[aux], b = inaux.b[aux];
            Col Kt = mk(a.x, a.y, a.z);
            float Fr = a.w;
            Col ab = mk(1.f, 1.f, 1.f);
            if (!H.front) {                                                        // Beer absorption on exit (:258-262)
                const DMaterial &PM = S.materials[mtl];
                ab = mk(expf((-H.z) * PM.absorption[0]), expf((-H.z) * PM.absorption[1]), expf((-H.z) * PM.absorption[2]));
            }
            Ws = (Wt * (ab * Kt)) * (float)(1.0 - (double)Fr);                     // :264
            // the Fresnel mirror ray exists only because the refracted ray hit (:234-251)
            Col Wf = Wt * Fr;
            Col WfKt = Wf * Kt;
            if (!((F.flags & 2u) && !nonblack(Wf))) {
                unsigned na = warp_alloc(O.aux.count, true);
                if (na >= O.aux.cap) O.counters->overflow = 1;
                else {
                    O.aux.a[na] = make_float4(Wf.r, Wf.g, Wf.b, 0.f);
                    O.aux.b[na] = make_float4(0, 0, 0, 0);
                    push_ray(O, ray.px, ray.py, ray.pz, b.x, b.y, b.z, WfKt, pixel, pack_meta(RK_FRESNEL, bounce, mtl, tree), (int)na,
                             child_path(path, 4u));
                }
            }
            if ((F.flags & 2u) && !nonblack(Ws)) n_shade = 0;
        }
        // the queue slot of a GI vertex's sample ray is reserved before the vertex is shaded: the atomic's round trip hides
        // behind shade_hit()
        const bool gi_ray = gi && k < F.gi_bounces;
        SlotTicket tk;
        if (gi_ray) tk.issue(O.next.count, 1u);
#pragma unroll 1
        for (int t = 0; t < n_shade; t++) {
            const Col w = gi ? mk(1.f, 1.f, 1.f) : Ws;
            const int sb = gi ? F.shade_bounces : bounce;
            const int target = gi ? rec + 2 * k + (t ? 0 : 1) : pixel;
            const unsigned sp = gi ? child_path(path, 8u + (unsigned)t) : path;
            shade_hit(S, SP, O, ray.dx, ray.dy, ray.dz, H, w, sb, target, sp, gi ? t : tree, gi);
        }
        if (gi) {
            if (gi_ray) {
                Rng rng;
                rng.key = F.seed; rng.pixel = 0x61u; rng.path = path; rng.dim = 0;
                float ox, oy, oz;
                sample_hemi_cos(rng, H.nx, H.ny, H.nz, ox, oy, oz);                                   // :561
                norm3(ox, oy, oz);                                                                    // :562
                const unsigned slot = tk.slot();
                if (slot >= O.next.cap) O.counters->overflow = 1;
                else {
                    O.next.o[slot] = make_float4(H.px, H.py, H.pz, __int_as_float(rec));
                    O.next.d[slot] = make_float4(ox, oy, oz, __uint_as_float(pack_meta(RK_GI, F.shade_bounces, 0, 0, k + 1)));
                    O.next.w[slot] = make_float4(1.f, 1.f, 1.f, __int_as_float(-1));
                    O.next.path[slot] = child_path(path, 6u);
                }
            } else {
                O.accum[rec + gi_end] = make_float4(0.1f, 0.1f, 0.1f, (float)(k + 1));               // :584
            }
        }
    }
}

template <bool PRIMARY>
__global__ void __launch_bounds__(WAVE_THREADS, SHADE_BLOCKS)
k_shade(DScene S, FrameSetup F, int s0, RayQueue in, AuxPool inaux, HitQueue hq, WaveOut O, unsigned *work, unsigned *gi_count)
{
    shade_body<PRIMARY>(S, F, s0, in, inaux, hq, O, work, gi_count);
}

// ------------------------------------------------------------------ any hit
template <bool FAST>
__device__ __forceinline__ void shadow_body(const DScene &S, const ShadowQueue &Q, float4 *accum, DCounters *counters, unsigned *work)
{
    Tally tl = {0, 0, 0, 0, 0};
    const unsigned lane = threadIdx.x & 31u;
    unsigned total = *Q.count;
    if (total > Q.cap) total = Q.cap;
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        unsigned idx = base + lane;
        if (idx >= total) continue;
        float4 o = Q.o[idx], d = Q.d[idx];
        Ray ray;
        ray.px = o.x; ray.py = o.y; ray.pz = o.z;
        ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
        Best B;
        B.z = d.w; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f; // h.z = t_max (lightFunctions.cpp:29)
        tl.shadow++;
        bool occ = scene_hit<true, FAST>(S, ray, B, tl, false);
        if (occ && B.z > 0.0f) continue;                                                // :31-35
        float4 c = Q.c[idx];
        accum_add(accum, __float_as_int(o.w), mk(c.x, c.y, c.z));
    }
    flush_tally(tl, counters, 2);
}

__global__ void __launch_bounds__(WAVE_THREADS, EXT_BLOCKS)
k_shadow_wave_simple(DScene S, ShadowQueue Q, float4 *accum, DCounters *counters, unsigned *work)
{
    shadow_body<false>(S, Q, accum, counters, work);
}

// Scenes with a top-level hierarchy (hundreds to thousands of nodes): one lane per ray, each on its own ordered, pruned
// search (scene_hit_bvh).  The k_*_top kernels further down keep the visit in scene order for RTU_FLAG_REFERENCE_WALK.
#ifndef BVH_BLOCKS
#define BVH_BLOCKS 3
#endif
template <bool PRIMARY>
__global__ void __launch_bounds__(WAVE_THREADS, BVH_BLOCKS)
k_extend_bvh(DScene S, FrameSetup F, int s0, int s1, RayQueue in, AuxPool inaux, HitQueue hq, float4 *accum, float4 *target,
             DCounters *counters, unsigned *work)
{
    extend_body<PRIMARY, true>(S, F, s0, s1, in, inaux, hq, accum, target, counters, work);
}

__global__ void __launch_bounds__(WAVE_THREADS, BVH_BLOCKS)
k_shadow_wave_bvh(DScene S, ShadowQueue Q, float4 *accum, DCounters *counters, unsigned *work)
{
    shadow_body<true>(S, Q, accum, counters, work);
}

// ------------------------------------------------------------------ the tail of a frame in one launch
// Deep waves hold few rays (the first bounces' survivors), yet each costs four launches whose floor - launch, ticket
// atomics, a chain of dependent loads through an almost empty pipeline - is ~15 us a piece: on 1-spp frames that is most
// of the frame.  k_tail_waves runs waves [w0, n_waves) as ONE cooperative launch of persistent threads: per wave
// reset / extend / shade / any-hit with grid-wide barriers in between (the three bodies are the stand-alone kernels' own
// code), and it stops at the first wave whose ray queue is empty.  It is chosen by the host from what the same frame did the
// last time (WaveLog); whatever the waves turn out to hold, the result is the ordinary one.
struct TailArgs {
    RayQueue q[2];
    AuxPool aux[2];
    ShadowQueue shadow;
    HitQueue hits;
    DCounters *counters;
    unsigned *gi_count;
    unsigned *work;      // three counters per wave
    unsigned *wave_log;  // rays that entered each wave (read back for the next frame's choice)
    float4 *target;
    int in_q, w0, n_waves;
};

__global__ void __launch_bounds__(WAVE_THREADS, EXT_BLOCKS)
k_tail_waves(DScene S, FrameSetup F, TailArgs A)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    int in_q = A.in_q;
    for (int w = A.w0; w < A.n_waves; w++) {
        const unsigned n_in = *(volatile unsigned *)A.q[in_q].count;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (A.wave_log)
                for (int k = w; k < (n_in ? w + 1 : A.n_waves); k++) A.wave_log[k] = n_in; // (an empty wave: so are all after it)
            *A.q[1 - in_q].count = 0;
            *A.aux[1 - in_q].count = 0;
            *A.shadow.count = 0;
            *A.hits.count = 0;
        }
        if (n_in == 0u) break; // (every thread of the grid reads the same word: nobody is left waiting at a barrier)
        grid.sync();
        unsigned *work = A.work + 3 * (w - A.w0);
        extend_body<false, true>(S, F, 0, 0, A.q[in_q], A.aux[in_q], A.hits, A.target, A.target, A.counters, work);
        grid.sync();
        WaveOut O;
        O.next = A.q[1 - in_q];
        O.aux = A.aux[1 - in_q];
        O.shadow = A.shadow;
        O.accum = A.target;
        O.counters = A.counters;
        shade_body<false>(S, F, 0, A.q[in_q], A.aux[in_q], A.hits, O, work + 1, A.gi_count);
        grid.sync();
        shadow_body<true>(S, A.shadow, A.target, A.counters, work + 2);
        grid.sync(); // the next wave's reset must not overtake this wave's readers
        in_q = 1 - in_q;
    }
}


// Flat scenes (every object a child of the root): any-hit with the mesh walks pooled per warp.
//
// Inside a BVH walk the lanes of a warp drift apart: rays that skim the teapot walk ten times longer than rays that
// are stopped by its first triangle, and in the plain kernel above 8 of 32 lanes are active per instruction of the
// walk (profiles/).  Only the boolean of a shadow ray is observable and it does not depend on the order in which
// the boxes of a mesh are opened, so here a warp walks the meshes of 32 rays as ONE pool of (ray, node) items in
// shared memory: every iteration 32 items are popped, each lane tests the two child boxes of its item and pushes the
// children that were hit; leaves go to a second pool whose triangles are tested 32 leaves at a time.  A ray is
// occluded as soon as one of its triangles accepts (its remaining items are dropped when popped); rays whose items
// run out resume their node loop after the mesh, so every ray still visits its nodes in the reference's order
// (RenderFunctions.cpp:224-240) and the first object that reports a hit still decides.  If the pool is about to
// overflow, the popped items are walked to the end by their lanes (bvh_walk) instead of being expanded.
#define SP_JOBS 64   // parked (ray, mesh node) entries waiting for a batch
#define SP_RES 96    // rays that missed their mesh and resume the node loop behind it
#ifndef SP_POOL
#define SP_POOL 512  // internal-node items: slot << 27 | pair index
#endif
#define SP_LEAF 192  // leaf items: slot << 27 | (count - 1) << 24 | first triangle (one iteration can add 128, 31 may wait)
#ifndef SHADOW_BLOCKS
#define SHADOW_BLOCKS EXT_BLOCKS
#endif
#ifndef SHADOW_PRUNE_TMAX
#define SHADOW_PRUNE_TMAX 1 // any-hit walks skip boxes entered beyond the light (result-neutral, see k_shadow_wave)
#endif

// (OCC: the hoisted reciprocals are only needed where a stopped ray is confirmed and are recomputed there, and one pointer
// table serves both hierarchies: 7.8 KB per warp instead of 8.6, so that two CTAs leave 124 KB instead of 92 KB of L1)
template <bool OCC> struct SpWarpT {
    float4 o[32];                // mesh-local origin, t_max
    float4 d[32];                // mesh-local direction, InvDir::ok
    float4 y[OCC ? 1 : 32];      // hoisted reciprocals (exact slab tests: cyBVH walk, ref_reaches)
    float4 ci[32];               // OccRay: 1/d, tlim
    float4 cn[32];               // OccRay: near-plane offsets
    float4 cf[32];               // OccRay: far-plane offsets
    const DMesh *mesh[32];
    unsigned hitslot[32];        // cyBVH slot of the triangle that (tentatively) occludes the slot's ray
    union {
        const OccNode *nodes[32]; // the hierarchy the pool walks: the mesh's 4-wide any-hit hierarchy (OCC) ...
        const BvhPair *pairs[32]; // ... or its cyBVH in the reference's own tests (RTU_FLAG_REFERENCE_WALK)
    };
    const TriRec *tris[32];
    unsigned idx[32], node[32];  // shadow-queue entry and mesh node of the slot
    // the slot's hoisted reciprocals
    __device__ __forceinline__ float4 recip(unsigned sl, const float4 d) const
    {
        if constexpr (OCC) {
            const InvDir I = make_invdir(d.x, d.y, d.z);
            return make_float4(I.yx, I.yy, I.yz, 0.f);
        } else return y[sl];
    }
    unsigned pool[SP_POOL];
    unsigned leaf[SP_LEAF];
    uint2 jobs[SP_JOBS];
    uint2 res[SP_RES];
    unsigned occl;               // bit s: the ray in slot s is occluded
};

template <bool FLAT, bool OCC>
__global__ void __launch_bounds__(WAVE_THREADS, SHADOW_BLOCKS)
k_shadow_wave(DScene S, ShadowQueue Q, float4 *accum, DCounters *counters, unsigned *work)
{
    extern __shared__ __align__(16) unsigned char sp_raw[];
    SpWarpT<OCC> &W = reinterpret_cast<SpWarpT<OCC> *>(sp_raw)[threadIdx.x >> 5];
    Tally tl = {0, 0, 0, 0, 0};
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u, FULL = 0xffffffffu, NONE = 0x7fffffffu;
    unsigned total = *Q.count;
    if (total > Q.cap) total = Q.cap;
    unsigned njobs = 0, nres = 0; // warp-uniform
    unsigned t_next = 0, t_end = 0; // this warp's block of work items: large waves take TICKET_BLOCK items per atomic,
    unsigned t_block;               // small ones fewer, so that every resident warp still gets several blocks
    t_block = 32u;
    bool drained = false;         // warp-uniform: no tickets left
    DNode root;
    load_node(S.nodes, root);
    for (;;) {
        if (njobs >= 32u || (njobs > 0u && nres == 0u && drained)) {
            // ------------------------------------------------------------ one batch of mesh walks
            const unsigned take = njobs < 32u ? njobs : 32u;
            njobs -= take;
            unsigned rootw = NONE;
            bool decided = false; // occluded before the pool starts (rays that do not use the hierarchy)
            if (lane < take) {
                uint2 j = W.jobs[njobs + lane];
                const bool scan = (j.y >> 31) != 0u;
                j.y &= 0x7fffffffu;
                float4 o = Q.o[j.x], d = Q.d[j.x];
                Ray ray;
                ray.px = o.x; ray.py = o.y; ray.pz = o.z;
                ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
                DNode nd;
                load_node(S.nodes + j.y, nd);
                const Ray lr = FLAT ? to_node(nd.itm, nd.pos, to_node(root.itm, root.pos, ray)) : local_ray_of(S, (int)j.y, ray, nullptr);
                const DMesh &M = S.meshes[nd.mesh];
                const InvDir I = mesh_invdir(M, lr);
                W.o[lane] = make_float4(lr.px, lr.py, lr.pz, d.w);
                W.d[lane] = make_float4(lr.dx, lr.dy, lr.dz, I.ok ? 1.f : 0.f);
                if constexpr (!OCC) W.y[lane] = make_float4(I.yx, I.yy, I.yz, 0.f);
                W.idx[lane] = j.x;
                W.node[lane] = j.y;
                if constexpr (OCC) {
                W.hitslot[lane] = 0xffffffffu;
                bool scanned = false;
                if (scan) {
                    // a shadow ray of a light with light lists (host/light_mask.cpp): the exact test on the triangles of its mask
                    // cell that begin before its origin decides; the hierarchy is not walked.  (The lookup is the node loop's,
                    // on the same local ray: it says 2 again.)
                    unsigned it0, it1;
                    float zcut;
                    if (light_mask_lookup(S, nd, lr, d.w, it0, it1, zcut) == 2) {
                        RefWalkArgs a;
                        a.px = lr.px; a.py = lr.py; a.pz = lr.pz; a.dx = lr.dx; a.dy = lr.dy; a.dz = lr.dz;
                        a.yx = I.yx; a.yy = I.yy; a.yz = I.yz; a.ok = I.ok;
                        Tally t2 = {0, 0, 0, 0, 0};
                        decided = light_list_occludes_ni(S.mask_lists + it0, (it1 - it0) >> 1, zcut, &M, &a, d.w, &t2);
                        tl.box += t2.box; tl.tri += t2.tri;
                        scanned = true;
                    }
                }
                if (!scanned) {
                const OccRay oc = occ_setup(lr, M.occ_scale, d.w);
                W.ci[lane] = make_float4(oc.ix, oc.iy, oc.iz, oc.tlim);
                W.cn[lane] = make_float4(oc.nx, oc.ny, oc.nz, 0.f);
                W.cf[lane] = make_float4(oc.fx, oc.fy, oc.fz, 0.f);
                W.mesh[lane] = &M;
                W.nodes[lane] = M.occ_nodes;
                W.tris[lane] = M.occ_tris;
                rootw = M.occ_root;
                if (!oc.ok) { // non-finite / huge components: the exact walk of the cyBVH decides
                    decided = bvh_walk_any_fallback(M, lr, I, d.w, tl);
                    rootw = NONE;
                }
                }
                } else {
                W.pairs[lane] = M.pairs;
                W.tris[lane] = M.tris;
                rootw = M.root;
                }
            }
            unsigned bi = __ballot_sync(FULL, rootw < NONE), bl = __ballot_sync(FULL, rootw > NONE);
            if (rootw < NONE) W.pool[__popc(bi & lt)] = (lane << 27) | rootw;
            if (rootw > NONE) W.leaf[__popc(bl & lt)] = (lane << 27) | (((rootw >> 28) & 7u) << 24) | (rootw & 0x00ffffffu);
            unsigned pool_n = __popc(bi), leaf_n = __popc(bl);
            const unsigned dec = __ballot_sync(FULL, decided);
            if (lane == 0) W.occl = dec;
            __syncwarp();
            for (;;) {
                const bool do_leaf = leaf_n >= 32u || (pool_n == 0u && leaf_n > 0u);
                if (!do_leaf && pool_n == 0u) break;
                const unsigned occl = *(volatile unsigned *)&W.occl;
                if (do_leaf) {
                    const unsigned n = leaf_n < 32u ? leaf_n : 32u;
                    leaf_n -= n;
                    if (lane < n) {
                        const unsigned it = W.leaf[leaf_n + lane], sl = it >> 27;
                        if (!((occl >> sl) & 1u)) {
                            const float4 o = W.o[sl], d = W.d[sl];
                            Ray r;
                            r.px = o.x; r.py = o.y; r.pz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                            const TriRec *tris = W.tris[sl];
                            const unsigned first = it & 0x00ffffffu, cnt = ((it >> 24) & 7u) + 1u;
                            if constexpr (OCC) {
                            for (unsigned i = 0; i < cnt; i++) {
                                const unsigned cs = occ_candidate(tris + first + i, r, o.w, tl);
                                if (cs != 0xffffffffu) { W.hitslot[sl] = cs; atomicOr(&W.occl, 1u << sl); break; } // tentative: confirmed below
                            }
                            } else {
                            for (unsigned i = 0; i < cnt; i++) {
                                const float4 *q = reinterpret_cast<const float4 *>(tris + first + i);
                                float4 x = __ldg(q), yv = __ldg(q + 1), w4 = __ldg(q + 2);
                                TriRec T;
                                T.nx = x.x; T.ny = x.y; T.nz = x.z; T.ax = x.w;
                                T.ay = yv.x; T.az = yv.y; T.area = yv.z; T.fbits = yv.w;
                                T.cau = w4.x; T.cav = w4.y; T.bau = w4.z; T.bav = w4.w;
                                tl.tri++;
                                float z = o.w, b1, b2, b3;
                                int fr;
                                if (tri_hit(T, r, z, fr, b1, b2, b3)) { atomicOr(&W.occl, 1u << sl); break; }
                            }
                            }
                        }
                    }
                    __syncwarp();
                    continue;
                }
                if constexpr (OCC) {
                // one item = one 4-wide node: up to four children go back to the pools
                const bool finish = pool_n > SP_POOL - 128u; // no room to expand 32 items: walk them to the end instead
                const unsigned n = pool_n < 32u ? pool_n : 32u;
                pool_n -= n;
                unsigned sl = 0, hit = 0;
                uint4 ch = make_uint4(NONE, NONE, NONE, NONE);
                if (lane < n) {
                    const unsigned it = W.pool[pool_n + lane];
                    sl = it >> 27;
                    if (!((occl >> sl) & 1u)) {
                        const float4 ci = W.ci[sl], cn = W.cn[sl], cf = W.cf[sl];
                        OccRay oc;
                        oc.ix = ci.x; oc.iy = ci.y; oc.iz = ci.z; oc.tlim = ci.w;
                        oc.nx = cn.x; oc.ny = cn.y; oc.nz = cn.z;
                        oc.fx = cf.x; oc.fy = cf.y; oc.fz = cf.z;
                        oc.ok = true;
                        if (finish) {
                            const float4 o = W.o[sl], d = W.d[sl], yv = W.recip(sl, d);
                            RefWalkArgs a;
                            a.px = o.x; a.py = o.y; a.pz = o.z; a.dx = d.x; a.dy = d.y; a.dz = d.z;
                            a.yx = yv.x; a.yy = yv.y; a.yz = yv.z; a.ok = d.w != 0.f;
                            Tally t2 = {0, 0, 0, 0, 0};
                            const OccRay oc2 = oc;
                            if (occ_walk_ni(W.mesh[sl], it & 0x07ffffffu, &a, &oc2, o.w, &t2)) atomicOr(&W.occl, 1u << sl);
                            tl.box += t2.box; tl.tri += t2.tri;
                        } else {
                            hit = occ_node(oc, W.nodes[sl] + (it & 0x07ffffffu), ch);
                            tl.box += 4;
                        }
                    }
                }
                // every lane pushes its hit children: internal nodes to the item pool, leaves to the leaf pool; positions from
                // one packed warp scan (internal count in the low half, leaf count in the high half)
                const unsigned w0 = (hit & 1u) ? ch.x : NONE, w1 = (hit & 2u) ? ch.y : NONE, w2 = (hit & 4u) ? ch.z : NONE, w3 = (hit & 8u) ? ch.w : NONE;
                // positions from eight ballots (independent of each other) instead of a five-step shuffle scan (a dependent chain)
                const unsigned i0 = __ballot_sync(FULL, w0 < NONE), i1 = __ballot_sync(FULL, w1 < NONE), i2 = __ballot_sync(FULL, w2 < NONE), i3 = __ballot_sync(FULL, w3 < NONE);
                const unsigned l0 = __ballot_sync(FULL, w0 > NONE), l1 = __ballot_sync(FULL, w1 > NONE), l2 = __ballot_sync(FULL, w2 > NONE), l3 = __ballot_sync(FULL, w3 > NONE);
                unsigned pi = pool_n + __popc(i0 & lt) + __popc(i1 & lt) + __popc(i2 & lt) + __popc(i3 & lt);
                unsigned li = leaf_n + __popc(l0 & lt) + __popc(l1 & lt) + __popc(l2 & lt) + __popc(l3 & lt);
                const unsigned tot = (__popc(i0) + __popc(i1) + __popc(i2) + __popc(i3)) | ((__popc(l0) + __popc(l1) + __popc(l2) + __popc(l3)) << 16);
                const unsigned tag = sl << 27;
#define RTU_PUSH(WORD)                                                                                                           \
                if (WORD < NONE) { W.pool[pi++] = tag | WORD; }                                  \
                else if (WORD > NONE) { W.leaf[li++] = tag | (((WORD >> 28) & 7u) << 24) | (WORD & 0x00ffffffu); }
                RTU_PUSH(w3)
                RTU_PUSH(w2)
                RTU_PUSH(w1)
                RTU_PUSH(w0)
#undef RTU_PUSH
                pool_n += tot & 0xffffu;
                leaf_n += tot >> 16;
#ifdef RTU_DEBUG_BOUNDS
                if (pool_n > SP_POOL || leaf_n > SP_LEAF) counters->overflow = 0xBAD4;
#endif
                __syncwarp();
                } else {
                const bool finish = pool_n > SP_POOL - 64u; // no room to expand 32 items: walk them to the end instead
                const unsigned n = pool_n < 32u ? pool_n : 32u;
                pool_n -= n;
                unsigned c1 = NONE, c2 = NONE, sl = 0;
                if (lane < n) {
                    const unsigned it = W.pool[pool_n + lane];
                    sl = it >> 27;
                    if (!((occl >> sl) & 1u)) {
                        const BvhPair *pairs = W.pairs[sl];
                        const float4 o = W.o[sl], d = W.d[sl], yv = W.recip(sl, d);
                        Ray r;
                        r.px = o.x; r.py = o.y; r.pz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                        InvDir I;
                        I.yx = yv.x; I.yy = yv.y; I.yz = yv.z; I.ok = d.w != 0.f;
                        if (finish) {
                            float z = o.w, b1, b2, b3;
                            int fr, tslot;
                            if (bvh_walk<true>(pairs, W.tris[sl], it & 0x07ffffffu, r, I, z, fr, tslot, b1, b2, b3, tl)) atomicOr(&W.occl, 1u << sl);
                        } else {
                            float4 a, b, c, dd;
                            load_pair(pairs + (it & 0x07ffffffu), a, b, c, dd);
                            float e1, e2;
                            bool h1 = slab_fast(r, I, a.x, a.y, a.z, a.w, b.x, b.y, RTU_BIG, e1);
                            bool h2 = slab_fast(r, I, b.z, b.w, c.x, c.y, c.z, c.w, RTU_BIG, e2);
                            tl.box += 2;
#if SHADOW_PRUNE_TMAX
                            // Only the boolean is observable: a box the ray enters well beyond the light holds no triangle
                            // that can pass `t < t_max` (objFunctions.cpp:270; the reference walks such boxes because it
                            // hands BIGFLOAT to BVHBoxIntersection, :358-359).  "Well beyond": 1 % + 0.01, orders of
                            // magnitude above the rounding of a triangle's own t against its box's tEntry.
                            const float tlim = o.w * 1.01f + 0.01f;
                            h1 = h1 && !(e1 > tlim);
                            h2 = h2 && !(e2 > tlim);
#endif
                            if (h1) c1 = __float_as_uint(dd.x);
                            if (h2) c2 = __float_as_uint(dd.y);
                        }
                    }
                }
                // push the children that were hit: child 2 below child 1, so that child 1 is popped first
                const unsigned n2 = __ballot_sync(FULL, c2 < NONE), n1 = __ballot_sync(FULL, c1 < NONE);
                const unsigned l2 = __ballot_sync(FULL, c2 > NONE), l1 = __ballot_sync(FULL, c1 > NONE);
                if (c2 < NONE) W.pool[pool_n + __popc(n2 & lt)] = (sl << 27) | c2;
                if (c1 < NONE) W.pool[pool_n + __popc(n2) + __popc(n1 & lt)] = (sl << 27) | c1;
                if (c2 > NONE) W.leaf[leaf_n + __popc(l2 & lt)] = (sl << 27) | (((c2 >> 28) & 7u) << 24) | (c2 & 0x00ffffffu);
                if (c1 > NONE) W.leaf[leaf_n + __popc(l2) + __popc(l1 & lt)] = (sl << 27) | (((c1 >> 28) & 7u) << 24) | (c1 & 0x00ffffffu);
                pool_n += __popc(n2) + __popc(n1);
                leaf_n += __popc(l2) + __popc(l1);
#ifdef RTU_DEBUG_BOUNDS
                if (pool_n > SP_POOL || leaf_n > SP_LEAF) counters->overflow = 0xBAD4;
#endif
                __syncwarp();
                }
            }
            // rays that were not stopped by their mesh go on with the node behind it
            const unsigned occl = *(volatile unsigned *)&W.occl;
            bool stopped = lane < take && ((occl >> lane) & 1u);
            if constexpr (OCC) {
                // one lane per ray: the triangle that stopped the ray is confirmed by the exact box test of its cyBVH leaf; a
                // rejection (hit point within rounding of the box) hands the ray to the exact walk of the cyBVH
                if (stopped && W.hitslot[lane] != 0xffffffffu) {
                    const float4 o = W.o[lane], d = W.d[lane], yv = W.recip(lane, d);
                    RefWalkArgs a;
                    a.px = o.x; a.py = o.y; a.pz = o.z; a.dx = d.x; a.dy = d.y; a.dz = d.z;
                    a.yx = yv.x; a.yy = yv.y; a.yz = yv.z; a.ok = d.w != 0.f;
                    Tally t2 = {0, 0, 0, 0, 0};
                    const bool reached = ref_reaches_ni(W.mesh[lane], W.hitslot[lane], &a, &t2);
                    tl.box += t2.box;
                    if (!reached) {
                        Ray r;
                        r.px = o.x; r.py = o.y; r.pz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                        InvDir I;
                        I.yx = yv.x; I.yy = yv.y; I.yz = yv.z; I.ok = d.w != 0.f;
                        stopped = bvh_walk_any_fallback(*W.mesh[lane], r, I, o.w, tl);
                    }
                }
            }
            const bool go_on = lane < take && !stopped;
            const unsigned m = __ballot_sync(FULL, go_on);
            if (go_on) W.res[nres + __popc(m & lt)] = make_uint2(W.idx[lane], W.node[lane] + 1u);
            nres += __popc(m);
#ifdef RTU_DEBUG_BOUNDS
            if (nres > SP_RES) counters->overflow = 0xBAD5;
#endif
            __syncwarp();
            continue;
        }
        // -------------------------------------------------------------------- node loop of 32 rays
        if (nres == 0u && drained) break; // (njobs == 0 here)
        // resuming rays have short node loops left, fresh rays long ones: a pass takes 32 of one kind, not a mix
        const unsigned k = (nres >= 32u || drained) ? (nres < 32u ? nres : 32u) : 0u;
        unsigned fresh = (k == 0u && !drained) ? 32u : 0u;
        unsigned base = 0;
        if (fresh) {
            // work items come in blocks of TICKET_BLOCK (8 tiles / 8 x 32 queue entries) per atomic: a wave of mostly empty
            // tiles would otherwise spend its time waiting for the single work counter
            if (t_next >= t_end) {
                if (lane == 0) t_next = atomicAdd(work, t_block);
                t_next = __shfl_sync(FULL, t_next, 0);
                t_end = t_next + t_block;
            }
            base = t_next;
            t_next += 32u;
            if (base >= total) { drained = true; fresh = 0; }
        }
        unsigned idx = 0;
        int i0 = 1;
        bool have = false;
        if (lane < k) {
            uint2 e = W.res[nres - k + lane];
            idx = e.x; i0 = (int)e.y;
            have = true;
        } else if (lane - k < fresh) {
            idx = base + (lane - k);
            have = idx < total;
            if (have) tl.shadow++;
        }
        nres -= k;
        int park = 0; // mesh node whose walk this lane's ray has to wait for
        unsigned scan = 0; // ... bit 31: not a walk, the scan of a light list (light_mask_lookup said 2)
        if (have) {
            float4 o = Q.o[idx], d = Q.d[idx];
            Ray ray;
            ray.px = o.x; ray.py = o.y; ray.pz = o.z;
            ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
            Best B;
            B.z = d.w; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f; // h.z = t_max (lightFunctions.cpp:29)
            const Ray r0 = to_node(root.itm, root.pos, ray);
            const float dd = dot3(r0.dx, r0.dy, r0.dz, r0.dx, r0.dy, r0.dz);
            bool occ = false, mesh_occ = false; // (mesh_occ: stopped by a triangle of a light list)
            Ray lvl[FLAT ? 1 : RTU_MAX_DEPTH]; // hierarchies: the ray at every depth of the current branch
            if (!FLAT) {
                lvl[0] = r0;
                if (i0 > 1 && i0 < S.n_nodes) local_ray_of(S, __ldg(&S.nodes[i0].parent), ray, lvl); // resumed behind a mesh
            }
            for (int i = i0; i < S.n_nodes; i++) {
                DNode nd;
                Ray lr;
                if (FLAT) {
                    const float4 bs = __ldg(&S.bounds[i]);
                    if (bs.w < 0.f && bs.w > -1.5f) continue; // no object
                    if (bound_culled(bs, r0, dd)) { tl.node++; tl.box++; continue; }
                    load_node(S.nodes + i, nd);
                    lr = to_node(nd.itm, nd.pos, r0);
                } else {
                    load_node(S.nodes + i, nd);
                    lr = to_node(nd.itm, nd.pos, lvl[nd.depth - 1]);
                    lvl[nd.depth] = lr; // children need it even when this node's own object is culled
                    if (nd.kind == 0) continue;
                    if (bound_culled(__ldg(&S.bounds[i]), r0, dd)) { tl.node++; tl.box++; continue; }
                }
                if (nd.kind == 3) { // TriObj::IntersectRay up to its bound-box gate (objFunctions.cpp:337)
                    const DMesh &M = S.meshes[nd.mesh];
                    tl.node++;
                    if (M.empty) continue;
                    tl.box++;
                    const InvDir I = mesh_invdir(M, lr);
                    float te;
                    if (!slab_fast(lr, I, M.bmin[0], M.bmin[1], M.bmin[2], M.bmax[0], M.bmax[1], M.bmax[2], RTU_BIG, te)) continue;
                    if constexpr (OCC) {
                        // the light's mask: beside the mesh's silhouette from its light -> skipped; with light lists the ray
                        // tests the few triangles of its cell that lie before its origin instead of walking the hierarchy
                        unsigned it0, it1;
                        float zcut;
                        const int lk = light_mask_lookup(S, nd, lr, d.w, it0, it1, zcut);
                        if (lk == 1) continue;
                        if (lk == 2) {
                            // Its list is scanned here when enough lanes of the warp arrived together with one (rays of one
                            // surface: a scan is 2-10 triangle tests), else with the next batch of mesh jobs, where all lanes are busy.
                            if (__popc(__activemask()) < S.scan_min) scan = 0x80000000u;
                            else {
                                RefWalkArgs a;
                                a.px = lr.px; a.py = lr.py; a.pz = lr.pz; a.dx = lr.dx; a.dy = lr.dy; a.dz = lr.dz;
                                a.yx = I.yx; a.yy = I.yy; a.yz = I.yz; a.ok = I.ok;
                                Tally t2 = {0, 0, 0, 0, 0};
                                const bool stop = light_list_occludes_ni(S.mask_lists + it0, (it1 - it0) >> 1, zcut, &M, &a, d.w, &t2);
                                tl.box += t2.box; tl.tri += t2.tri;
                                if (stop) { mesh_occ = true; break; }
                                continue;
                            }
                        }
                    }
                    park = i;
                    break;
                }
                if (sphere_or_plane_hit(nd, i, lr, B, tl)) { occ = true; break; }
            }
            if (!park && !mesh_occ && !(occ && B.z > 0.0f)) {                             // :31-35
                float4 c = Q.c[idx];
                accum_add(accum, __float_as_int(o.w), mk(c.x, c.y, c.z));
            }
        }
        const unsigned m = __ballot_sync(FULL, park != 0);
        if (park) W.jobs[njobs + __popc(m & lt)] = make_uint2(idx, (unsigned)park | scan);
        njobs += __popc(m);
#ifdef RTU_DEBUG_BOUNDS
        if (njobs > SP_JOBS) counters->overflow = 0xBAD6;
#endif
        __syncwarp();
    }
    flush_tally(tl, counters, 2);
}

__global__ void k_reset_counts(unsigned *a, unsigned *b, unsigned *c, unsigned *d, unsigned *log_dst, const unsigned *log_src)
{
    if (log_dst) *log_dst = *log_src; // rays entering the wave that is about to start (WaveLog)
    if (a) *a = 0;
    if (b) *b = 0;
    if (c) *c = 0;
    if (d) *d = 0;
}

__global__ void __launch_bounds__(WAVE_THREADS, EXT_BLOCKS)
k_primary_ids(DScene S, DCamera C, float *z, int *node, int *face, DCounters *counters)
{
    Tally tl = {0, 0, 0, 0, 0};
    int npix = C.width * C.height;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        int y = p / C.width, x = p - y * C.width;
        Ray ray = camera_ray(C, x, y, 0.5f, 0.5f, nullptr);
        Best B;
        B.z = RTU_BIG; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f;
        tl.trace++;
        scene_hit<false>(S, ray, B, tl);
        if (z) z[p] = B.z;
        if (node) node[p] = B.node;
        if (face) {
            int f = -1;
            if (B.node >= 0 && S.nodes[B.node].kind == 3) {
                const DMesh &M = S.meshes[S.nodes[B.node].mesh];
                f = (int)(((unsigned)__float_as_int(M.tris[B.slot].fbits)) & 0x3fffffffu);
            }
            face[p] = f;
        }
    }
    flush_tally(tl, counters, 0);
}

__global__ void __launch_bounds__(WAVE_THREADS, 2)
k_trace_batch(DScene S, const rtu_ray *rays, long long n, rtu_hit *hits, DCounters *counters)
{
    Tally tl = {0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Ray ray;
        ray.px = rays[i].p[0]; ray.py = rays[i].p[1]; ray.pz = rays[i].p[2];
        ray.dx = rays[i].dir[0]; ray.dy = rays[i].dir[1]; ray.dz = rays[i].dir[2];
        Best B;
        B.z = RTU_BIG; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f;
        tl.trace++;
        scene_hit<false>(S, ray, B, tl);
        HitRec H;
        finalize_hit(S, ray, B, H);
        rtu_hit o;
        o.z = H.z;
        o.p[0] = H.px; o.p[1] = H.py; o.p[2] = H.pz;
        o.N[0] = H.nx; o.N[1] = H.ny; o.N[2] = H.nz;
        o.uvw[0] = H.u; o.uvw[1] = H.v; o.uvw[2] = H.w;
        o.node = H.node; o.face = H.face; o.front = H.front;
        hits[i] = o;
    }
    flush_tally(tl, counters, 0);
}

__global__ void __launch_bounds__(WAVE_THREADS, EXT_BLOCKS)
k_shadow_batch(DScene S, const rtu_ray *rays, const float *tmax, long long n, unsigned char *occ, DCounters *counters)
{
    Tally tl = {0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Ray ray;
        ray.px = rays[i].p[0]; ray.py = rays[i].p[1]; ray.pz = rays[i].p[2];
        ray.dx = rays[i].dir[0]; ray.dy = rays[i].dir[1]; ray.dz = rays[i].dir[2];
        Best B;
        B.z = tmax[i]; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f;
        tl.shadow++;
        bool h = scene_hit<true>(S, ray, B, tl);
        occ[i] = (h && B.z > 0.0f) ? 1 : 0;
    }
    flush_tally(tl, counters, 2);
}

// rtu_shadow_trace through the frame's own any-hit kernel: caller rays become shadow-queue entries whose contribution is 1
// for "pixel" i, so accum[i].x stays 0 exactly where ray i is occluded.
__global__ void k_fill_shadow_queue(const rtu_ray *rays, const float *tmax, long long n, ShadowQueue Q)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Q.o[i] = make_float4(rays[i].p[0], rays[i].p[1], rays[i].p[2], __int_as_float((int)i));
        Q.d[i] = make_float4(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2], tmax[i]);
        Q.c[i] = make_float4(1.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *Q.count = (unsigned)n;
}
__global__ void k_occluded_from_accum(const float4 *accum, long long n, unsigned char *occ)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        occ[i] = accum[i].x == 0.f ? 1 : 0;
}

// rtu_trace through the frame's own closest-hit kernel: caller rays become queue entries of a kind whose miss adds nothing
// (RK_TIR with zero weight); the compacted hits are expanded to HitInfo records afterwards, every other ray keeps the
// record of a miss.
__global__ void k_fill_ray_queue(const rtu_ray *rays, long long n, RayQueue Q, rtu_hit *out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Q.o[i] = make_float4(rays[i].p[0], rays[i].p[1], rays[i].p[2], __int_as_float((int)i));
        Q.d[i] = make_float4(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2], __uint_as_float(pack_meta(RK_TIR, 0, 0)));
        Q.w[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        Q.path[i] = 0u;
        rtu_hit o; // HitInfo::Init (scene.h:162)
        o.z = RTU_BIG;
        o.p[0] = o.p[1] = o.p[2] = 0.f;
        o.N[0] = o.N[1] = o.N[2] = 0.f;
        o.uvw[0] = o.uvw[1] = o.uvw[2] = 0.5f;
        o.node = -1; o.face = -1; o.front = 1;
        out[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *Q.count = (unsigned)n;
}
__global__ void __launch_bounds__(WAVE_THREADS, 2)
k_hits_to_records(DScene S, RayQueue Q, HitQueue hq, rtu_hit *out)
{
    unsigned total = *hq.count;
    if (total > hq.cap) total = hq.cap;
    for (unsigned h = blockIdx.x * blockDim.x + threadIdx.x; h < total; h += gridDim.x * blockDim.x) {
        const float4 ha = hq.a[h], hb = hq.b[h];
        Best B;
        B.z = ha.x; B.node = __float_as_int(ha.y); B.front = __float_as_int(ha.z); B.slot = __float_as_int(ha.w);
        B.bc1 = hb.x; B.bc2 = hb.y; B.bc3 = hb.z;
        const unsigned idx = __float_as_uint(hb.w);
        const float4 o = Q.o[idx], d = Q.d[idx];
        Ray ray;
        ray.px = o.x; ray.py = o.y; ray.pz = o.z; ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
        HitRec H;
        finalize_hit(S, ray, B, H);
        rtu_hit r;
        r.z = H.z;
        r.p[0] = H.px; r.p[1] = H.py; r.p[2] = H.pz;
        r.N[0] = H.nx; r.N[1] = H.ny; r.N[2] = H.nz;
        r.uvw[0] = H.u; r.uvw[1] = H.v; r.uvw[2] = H.w;
        r.node = H.node; r.face = H.face; r.front = H.front;
        out[idx] = r;
    }
}

// First Shade() step on caller-provided hits (rtu_shade); pixel index = ray index.
__global__ void __launch_bounds__(WAVE_THREADS, 2)
k_shade_first(DScene S, FrameSetup F, const rtu_ray *rays, const rtu_hit *hits, long long n, WaveOut O)
{
    ShadeParams SP;
    SP.flags = F.flags;
    SP.seed = F.seed;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        rtu_hit h = hits[i];
        if (h.node < 0 || h.node >= S.n_nodes) continue;
        HitRec H;
        H.z = h.z;
        H.px = h.p[0]; H.py = h.p[1]; H.pz = h.p[2];
        H.nx = h.N[0]; H.ny = h.N[1]; H.nz = h.N[2];
        H.u = h.uvw[0]; H.v = h.uvw[1]; H.w = h.uvw[2];
        H.node = h.node; H.face = h.face; H.front = h.front;
        H.material = S.nodes[h.node].material;
        shade_hit(S, SP, O, rays[i].dir[0], rays[i].dir[1], rays[i].dir[2], H, mk(1.f, 1.f, 1.f), F.shade_bounces, (int)i, 0u);
    }
}

__global__ void k_camera_rays(DCamera C, float ox, float oy, rtu_ray *rays)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= C.width * C.height) return;
    int y = p / C.width, x = p - y * C.width;
    Ray r = camera_ray(C, x, y, ox, oy, nullptr);
    rtu_ray o;
    o.p[0] = r.px; o.p[1] = r.py; o.p[2] = r.pz;
    o.dir[0] = r.dx; o.dir[1] = r.dy; o.dir[2] = r.dz;
    rays[p] = o;
}

__global__ void k_resolve(const float4 *accum, int npix, int spp, float *rgb, unsigned char *rgb8)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float4 a = accum[p];
    float n = (float)spp;
    float r = a.x / n, g = a.y / n, b = a.z / n; // pixelValuesSum /= (float)maxSampleSize (RenderFunctions.cpp:152)
    if (rgb) { rgb[p * 3] = r; rgb[p * 3 + 1] = g; rgb[p * 3 + 2] = b; }
    if (rgb8) {
        float c[3] = {r, g, b};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float gmm = (float)pow((double)c[k], 1 / 2.2); // :155-157
            float s = gmm * 255;                           // Color24::FloatToByte (cyColor.h:245-246)
            int v = (s == s) ? (int)s : 0;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            rgb8[p * 3 + k] = (unsigned char)v;
        }
    }
}

// dst += src (a frame that was rendered on its own and is now known to be complete; see render_checked in rtu_api.cu)
__global__ void k_accum_add(float4 *dst, const float4 *src, size_t npix)
{
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
        float4 a = dst[p];
        const float4 b = src[p];
        a.x += b.x; a.y += b.y; a.z += b.z;
        dst[p] = a;
    }
}

// The RGB sums of the accumulator as three planes of npix floats: what a collective has to move (the .w lane carries nothing)
__global__ void k_pack_rgb(const float4 *accum, size_t npix, float *planes)
{
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
        const float4 a = accum[p];
        planes[p] = a.x;
        planes[npix + p] = a.y;
        planes[2 * npix + p] = a.z;
    }
}

// k_resolve on reduced planes (RenderFunctions.cpp:152-159); rows [row0,row1) only when a rank resolves its own tile rows
__global__ void k_resolve_planes(const float *planes, size_t npix, int spp, float *rgb, unsigned char *rgb8)
{
    size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float n = (float)spp;
    float c[3] = {planes[p] / n, planes[npix + p] / n, planes[2 * npix + p] / n};
    if (rgb) { rgb[p * 3] = c[0]; rgb[p * 3 + 1] = c[1]; rgb[p * 3 + 2] = c[2]; }
    if (rgb8) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float gmm = (float)pow((double)c[k], 1 / 2.2);
            float s = gmm * 255;
            int v = (s == s) ? (int)s : 0;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            rgb8[p * 3 + k] = (unsigned char)v;
        }
    }
}

// ---- adaptive sampling (minSampleSize / targetVariance / sampleIncrement, RenderFunctions.cpp:25-28, declared by
// the reference and never wired up).  One thread per 8x4 tile; see launch_adaptive_update.
__global__ void k_adaptive_update(const float4 *accum, int W, int H, int n_now, int max_spp, float target, unsigned char *tile_done,
                                  int *tile_samples, unsigned *n_active)
{
    const int tilesX = (W + 7) >> 3, tilesY = (H + 3) >> 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tilesX * tilesY) return;
    if (tile_done[t]) return;
    tile_samples[t] = n_now;
    const int tx = t % tilesX, ty = t / tilesX;
    const size_t npix = (size_t)W * H;
    const float nA = (float)(n_now - n_now / 2), nB = (float)(n_now / 2); // even / odd samples of [0, n_now)
    float worst = 0.f;
    if (n_now >= 2) {
        for (int j = 0; j < 4; j++)
            for (int i = 0; i < 8; i++) {
                const int x = tx * 8 + i, y = ty * 4 + j;
                if (x >= W || y >= H) continue;
                const float4 a = accum[(size_t)y * W + x], b = accum[npix + (size_t)y * W + x];
                const float dr = 0.5f * (a.x / nA - b.x / nB), dg = 0.5f * (a.y / nA - b.y / nB), db = 0.5f * (a.z / nA - b.z / nB);
                float v = fmaxf(fmaxf(dr * dr, dg * dg), db * db);
                if (!(v == v)) v = 0.f; // a NaN pixel never converges by waiting
                worst = fmaxf(worst, v);
            }
    } else {
        worst = 3.0e38f;
    }
    if (worst <= target || n_now >= max_spp) tile_done[t] = 1;
    else atomicAdd(n_active, 1u);
}

__global__ void k_resolve_adaptive(const float4 *accum, int W, int H, const int *tile_samples, float *rgb, unsigned char *rgb8,
                                   unsigned char *sample_count)
{
    const size_t npix = (size_t)W * H;
    const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const int n = tile_samples[(y >> 2) * ((W + 7) >> 3) + (x >> 3)];
    const float4 a = accum[p], b = accum[npix + p];
    const float fn = (float)(n > 0 ? n : 1);
    float c[3] = {(a.x + b.x) / fn, (a.y + b.y) / fn, (a.z + b.z) / fn};
    if (rgb) { rgb[p * 3] = c[0]; rgb[p * 3 + 1] = c[1]; rgb[p * 3 + 2] = c[2]; }
    if (rgb8) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float gmm = (float)pow((double)c[k], 1 / 2.2);
            float s = gmm * 255;
            int v = (s == s) ? (int)s : 0;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            rgb8[p * 3 + k] = (unsigned char)v;
        }
    }
    if (sample_count) sample_count[p] = (unsigned char)(n > 255 ? 255 : n);
}

__global__ void k_zminmax(const float *z, int npix, unsigned *mm)
{
    // zmin / zmax over hit pixels (scene.h:596-601); positive floats order like their bit patterns
    unsigned lo = 0x7f800000u, hi = 0u;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        float v = z[p];
        if (v == RTU_BIG) continue;
        float vc = v < 0.f ? 0.f : v;
        unsigned b = __float_as_uint(vc);
        lo = min(lo, b);
        hi = max(hi, b);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[0], lo);
        atomicMax(&mm[1], hi);
    }
}

__global__ void k_zimage(const float *z, int npix, const unsigned *mm, unsigned char *z8)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float zmin = __uint_as_float(mm[0]), zmax = __uint_as_float(mm[1]);
    if (zmin > RTU_BIG) zmin = RTU_BIG; // no hit at all: zmin stays BIGFLOAT (scene.h:596)
    float v = z[p];
    unsigned char o = 0;
    if (v != RTU_BIG) {
        float f = (zmax - v) / (zmax - zmin); // scene.h:605
        float s = f * 255;
        int c = (s == s) ? (int)s : 0;
        c = c < 0 ? 0 : (c > 255 ? 255 : c);
        o = (unsigned char)c;
    }
    z8[p] = o;
}

// ------------------------------------------------------------------ scenes with very many nodes
// Thousands of objects: stepping through every node per ray (what Trace() does) is the whole cost, and letting every
// lane walk the top-level hierarchy on its own serialises 32 different walks.  Here the warp nominates the nodes of
// its 32 rays together: a pool of (ray, hierarchy node) items in shared memory, 32 box tests per iteration, leaves
// append their object nodes to the ray's list.  Each lane then sorts its own list by node index and runs the usual
// per-node code over it (scene_hit_list), so results and counters are those of the linear visit.
#define TW_POOL 512
struct TopWarp {
    float4 p[32], d[32], inv[32];
    unsigned pool[TW_POOL];
    int count[32];
    int list[32][RTU_TOP_CAND];
};

// returns the number of nominees of this lane's ray (in W.list[lane], ascending), or -1 if they did not fit
__device__ __forceinline__ int warp_nominate(const DScene &S, TopWarp &W, const Ray &r0, bool have, unsigned lane)
{
    const unsigned lt = (1u << lane) - 1u, FULL = 0xffffffffu, NONE = 0x7fffffffu;
    W.p[lane] = make_float4(r0.px, r0.py, r0.pz, 0.f);
    W.d[lane] = make_float4(r0.dx, r0.dy, r0.dz, 0.f);
    W.inv[lane] = make_float4(1.f / r0.dx, 1.f / r0.dy, 1.f / r0.dz, 0.f);
    W.count[lane] = 0;
    const unsigned hv = __ballot_sync(FULL, have);
    if (have) W.pool[__popc(hv & lt)] = lane << 27; // hierarchy node 0 = root
    unsigned pool_n = __popc(hv);
    __syncwarp();
    while (pool_n > 0u) {
        const bool finish = pool_n > TW_POOL - 64u; // no room to expand: the popped items are finished by their lanes
        const unsigned n = pool_n < 32u ? pool_n : 32u;
        pool_n -= n;
        unsigned c1 = NONE, c2 = NONE, sl = 0;
        if (lane < n) {
            const unsigned it = W.pool[pool_n + lane];
            sl = it >> 27;
            const float4 p = W.p[sl], d = W.d[sl], iv = W.inv[sl];
            int stack[32];
            int top = 0;
            stack[0] = (int)(it & 0x07ffffffu);
            if (*(volatile int *)&W.count[sl] > RTU_TOP_CAND) top = -1; // the ray's list is already full: it will be visited linearly
            while (top >= 0) {
                const int ni = stack[top--];
                const float4 *q = reinterpret_cast<const float4 *>(S.top + ni);
                const float4 lo = __ldg(q), hi = __ldg(q + 1);
                if (!top_box_crossed(lo, hi, p.x, p.y, p.z, d.x, d.y, d.z, iv.x, iv.y, iv.z)) continue;
                const int a = __float_as_int(lo.w), b = __float_as_int(hi.w);
                if (a >= 0) {
                    if (!finish) { c1 = (unsigned)a; c2 = (unsigned)b; break; } // children go back to the pool
                    if (top + 2 >= 32) { atomicAdd(&W.count[sl], RTU_TOP_CAND + 1); break; }
                    stack[++top] = a;
                    stack[++top] = b;
                } else {
                    const int first = -a - 1;
                    for (int k = 0; k < b; k++) {
                        const int pos = atomicAdd(&W.count[sl], 1);
                        if (pos < RTU_TOP_CAND) W.list[sl][pos] = __ldg(&S.top_items[first + k]);
                    }
                }
            }
        }
        const unsigned b2 = __ballot_sync(FULL, c2 != NONE), b1 = __ballot_sync(FULL, c1 != NONE);
        if (c2 != NONE) W.pool[pool_n + __popc(b2 & lt)] = (sl << 27) | c2;
        if (c1 != NONE) W.pool[pool_n + __popc(b2) + __popc(b1 & lt)] = (sl << 27) | c1;
        pool_n += __popc(b2) + __popc(b1);
#ifdef RTU_DEBUG_BOUNDS
        if (pool_n > TW_POOL) asm volatile("trap;");
#endif
        __syncwarp();
    }
    int nc = W.count[lane];
    if (!have) return 0;
    if (nc > RTU_TOP_CAND) return -1;
    int *L = W.list[lane];
    sort_ascending(L, nc); // ascending node order
    return nc;
}

template <bool PRIMARY>
__global__ void __launch_bounds__(WAVE_THREADS, EXT_BLOCKS)
k_extend_top(DScene S, FrameSetup F, int s0, int s1, RayQueue in, AuxPool inaux, HitQueue hq, float4 *accum, float4 *target,
             DCounters *counters, unsigned *work)
{
    extern __shared__ __align__(16) unsigned char tw_raw[];
    TopWarp &W = reinterpret_cast<TopWarp *>(tw_raw)[threadIdx.x >> 5];
    Tally tl = {0, 0, 0, 0, 0};
    const unsigned lane = threadIdx.x & 31u;
    PrimaryMap pm;
    pm.init(F);
    unsigned total;
    if (PRIMARY) total = pm.perSample * (unsigned)(s1 - s0);
    else { total = *in.count; if (total > in.cap) total = in.cap; }
    DNode root;
    load_node(S.nodes, root);
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        const unsigned idx = base + lane;
        bool have = idx < total;
        Ray ray;
        ray.px = ray.py = ray.pz = 0.f; ray.dx = ray.dy = 0.f; ray.dz = 1.f;
        int pixel = 0, x = 0, y = 0;
        if (have) {
            if (PRIMARY) {
                int s;
                have = pm.decode(idx, s0, s, x, y);
                if (have && F.tile_done && F.tile_done[pm.tile_of(idx)]) have = false; // adaptive sampling: converged tile
                if (have && F.tile_empty && F.tile_empty[pm.tile_of(idx)]) { primary_miss_fast(S, F, pm, accum, x, y, s, s0, s1 - s0, tl); have = false; }
                pixel = y * pm.W + x;
                if (have) ray = primary_ray(F, s, x, y, pixel);
                pixel = half_slot(F, pixel, s); // from here on: the accumulator slot
            } else {
                float4 o = in.o[idx], d = in.d[idx];
                ray.px = o.x; ray.py = o.y; ray.pz = o.z;
                ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
            }
        }
        const Ray r0 = to_node(root.itm, root.pos, ray);
        const int nc = warp_nominate(S, W, r0, have, lane);
        if (have) {
            Best B;
            B.z = RTU_BIG; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f;
            tl.trace++;
            if (nc >= 0) scene_hit_list<false>(S, ray, r0, dot3(r0.dx, r0.dy, r0.dz, r0.dx, r0.dy, r0.dz), W.list[lane], nc, B, tl);
            else {
                bool done;
                scene_hit_long_list<false>(S, ray, B, tl, done);
                if (!done) scene_hit<false>(S, ray, B, tl, false);
            }
            extend_finish<PRIMARY>(S, F, pm, in, inaux, hq, accum, target, counters, idx, ray, pixel, x, y, B);
        }
        __syncwarp();
    }
    flush_tally(tl, counters, PRIMARY ? 0 : 1);
}

__global__ void __launch_bounds__(WAVE_THREADS, EXT_BLOCKS)
k_shadow_wave_top(DScene S, ShadowQueue Q, float4 *accum, DCounters *counters, unsigned *work)
{
    extern __shared__ __align__(16) unsigned char tw_raw[];
    TopWarp &W = reinterpret_cast<TopWarp *>(tw_raw)[threadIdx.x >> 5];
    Tally tl = {0, 0, 0, 0, 0};
    const unsigned lane = threadIdx.x & 31u;
    unsigned total = *Q.count;
    if (total > Q.cap) total = Q.cap;
    DNode root;
    load_node(S.nodes, root);
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        const unsigned idx = base + lane;
        const bool have = idx < total;
        Ray ray;
        ray.px = ray.py = ray.pz = 0.f; ray.dx = ray.dy = 0.f; ray.dz = 1.f;
        float4 o = make_float4(0, 0, 0, 0), d = make_float4(0, 0, 1, 0);
        if (have) {
            o = Q.o[idx]; d = Q.d[idx];
            ray.px = o.x; ray.py = o.y; ray.pz = o.z;
            ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
        }
        const Ray r0 = to_node(root.itm, root.pos, ray);
        const int nc = warp_nominate(S, W, r0, have, lane);
        if (have) {
            Best B;
            B.z = d.w; B.node = -1; B.front = 1; B.slot = 0; B.bc1 = B.bc2 = B.bc3 = 0.f; // h.z = t_max (lightFunctions.cpp:29)
            tl.shadow++;
            bool occ;
            if (nc >= 0) occ = scene_hit_list<true>(S, ray, r0, dot3(r0.dx, r0.dy, r0.dz, r0.dx, r0.dy, r0.dz), W.list[lane], nc, B, tl);
            else {
                bool done;
                occ = scene_hit_long_list<true>(S, ray, B, tl, done);
                if (!done) occ = scene_hit<true>(S, ray, B, tl, false);
            }
            if (!(occ && B.z > 0.0f)) {                                                 // :31-35
                float4 c = Q.c[idx];
                accum_add(accum, __float_as_int(o.w), mk(c.x, c.y, c.z));
            }
        }
        __syncwarp();
    }
    flush_tally(tl, counters, 2);
}

// ------------------------------------------------------------------ empty tiles of the primary wave
// One thread per 8x4-pixel tile.  The image-plane points of the tile's camera rays (all rendered samples: sub-pixel
// offsets in [ox0,ox1] x [oy0,oy1]) fill a rectangle in pixel coordinates; the tile is empty when that rectangle is
// separated from the footprint of every object (bounding boxes disjoint, or all four corners outside one hull edge).
__global__ void k_tile_mask(FrameSetup F, const TileObject *objs, int n_objs, const float4 *edges, float ox0, float ox1, float oy0,
                            float oy1, unsigned char *mask, unsigned *n_empty)
{
    const int tilesX = (F.cam.width + 7) >> 3, tilesY = (F.row_end - F.row_begin + 3) >> 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tilesX * tilesY) return;
    const int tx = t % tilesX, ty = t / tilesX;
    const float i0 = tx * 8 + ox0, i1 = tx * 8 + 7 + ox1;
    const float j0 = F.row_begin + ty * 4 + oy0, j1 = F.row_begin + ty * 4 + 3 + oy1;
    bool empty = true;
    for (int k = 0; k < n_objs && empty; k++) {
        const TileObject o = objs[k];
        if (i1 < o.lo[0] || i0 > o.hi[0] || j1 < o.lo[1] || j0 > o.hi[1]) continue;
        bool separated = false;
        for (int e = 0; e < o.n_edges && !separated; e++) {
            const float4 h = edges[o.first_edge + e];
            separated = h.x * i0 + h.y * j0 > h.z && h.x * i1 + h.y * j0 > h.z && h.x * i0 + h.y * j1 > h.z && h.x * i1 + h.y * j1 > h.z;
        }
        if (!separated) empty = false;
    }
    mask[t] = empty ? 1 : 0;
    const unsigned votes = __ballot_sync(__activemask(), empty);
    if ((threadIdx.x & 31) == 0 && votes) atomicAdd(n_empty, (unsigned)__popc(votes));
}

void launch_tile_mask(cudaStream_t st, const FrameSetup &F, const TileObject *objs, int n_objs, const float4 *edges, float ox0,
                      float ox1, float oy0, float oy1, unsigned char *mask, unsigned *n_empty)
{
    const int tiles = ((F.cam.width + 7) >> 3) * ((F.row_end - F.row_begin + 3) >> 2);
    cudaMemsetAsync(n_empty, 0, sizeof(unsigned), st);
    k_tile_mask<<<(tiles + 127) / 128, 128, 0, st>>>(F, objs, n_objs, edges, ox0, ox1, oy0, oy1, mask, n_empty);
}

// ------------------------------------------------------------------ launch wrappers
// grid = SM count x CTAs that are actually resident per SM for that kernel (occupancy API, cached)
template <class K> static int resident_grid(const LaunchCfg &cfg, K kernel, int *cache)
{
    if (*cache == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, WAVE_THREADS, 0) != cudaSuccess || n < 1) n = 1;
        if (cfg.blocks_per_sm > 0 && n > cfg.blocks_per_sm) n = cfg.blocks_per_sm;
        *cache = n;
    }
    return cfg.sm_count * *cache;
}

// the per-lane search kernels of many-node scenes: as many CTAs as fit (BVH_BLOCKS)
template <class K> static int bvh_grid(const LaunchCfg &cfg, K kernel, int *cache)
{
    if (*cache == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, WAVE_THREADS, 0) != cudaSuccess || n < 1) n = 1;
        if (n > BVH_BLOCKS) n = BVH_BLOCKS;
        *cache = n;
    }
    return cfg.sm_count * *cache;
}

static int top_mode()
{
    static int mode = -1;
    if (mode < 0) { // RTU_TOP_KERNEL=list selects the nominate / sort / visit-in-order kernels (A/B measurements)
        const char *e = getenv("RTU_TOP_KERNEL");
        mode = (e && e[0] == 'l') ? 0 : 1;
    }
    return mode;
}

// the shading kernel has its own residency (SHADE_BLOCKS), independent of the traversal kernels' cfg.blocks_per_sm
template <class K> static int shade_grid(const LaunchCfg &cfg, K kernel, int *cache)
{
    if (*cache == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, WAVE_THREADS, 0) != cudaSuccess || n < 1) n = 1;
        if (n > SHADE_BLOCKS) n = SHADE_BLOCKS;
        *cache = n;
    }
    return cfg.sm_count * *cache;
}

static WaveOut make_out(const WaveBuffers &B, int out_q, float4 *accum)
{
    WaveOut O;
    O.next = B.q[out_q];
    O.aux = B.aux[out_q];
    O.shadow = B.shadow;
    O.accum = accum;
    O.counters = B.counters;
    return O;
}

static int extend_mode()
{
    static int mode = -1;
    if (mode < 0) { // RTU_EXTEND_KERNEL=simple selects the plain kernel (A/B measurements)
        const char *e = getenv("RTU_EXTEND_KERNEL");
        mode = (e && e[0] == 's') ? 0 : 1;
    }
    return mode;
}

template <class K> static int pooled_grid(const LaunchCfg &cfg, K kernel, size_t smem, int *cache)
{
    if (*cache == 0) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, WAVE_THREADS, smem) != cudaSuccess || n < 1) n = 1;
        while (n > 1 && cfg.sm_count * n * (WAVE_THREADS / 32) > XP_MAX_WARPS) n--; // one parking region per resident warp
        if (cfg.blocks_per_sm > 0 && n > cfg.blocks_per_sm) n = cfg.blocks_per_sm;
        *cache = n;
    }
    return cfg.sm_count * *cache;
}

// <PRIMARY, FLAT, OCC>: OCC walks the meshes' 4-wide hierarchies with pruning; !OCC (RTU_FLAG_REFERENCE_WALK) walks the cyBVH
// with the reference's own tests and books exactly the work Trace() does
template <bool PRIMARY> static void launch_extend_pooled(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, int s0, int s1,
                                                         const RayQueue &q, const AuxPool &aux, const WaveBuffers &B, float4 *pixel_accum, float4 *accum,
                                                         unsigned *work_counter)
{
    static int occ[4] = {0, 0, 0, 0};
    const size_t smem = sizeof(XpWarp) * (WAVE_THREADS / 32);
    const bool ref = (F.flags & RTU_FLAG_REFERENCE_WALK) != 0 && !S.any_no_ref;
#define RTU_LAUNCH_XP(FLAT_, OCC_, SLOT)                                                                                                \
    k_extend_pool<PRIMARY, FLAT_, OCC_><<<pooled_grid(cfg, k_extend_pool<PRIMARY, FLAT_, OCC_>, smem, &occ[SLOT]), WAVE_THREADS, smem, st>>>( \
        S, F, s0, s1, q, aux, B.hits, pixel_accum, accum, B.counters, work_counter, B.park)
    if (S.flat && !ref) RTU_LAUNCH_XP(true, true, 0);
    else if (S.flat) RTU_LAUNCH_XP(true, false, 1);
    else if (!ref) RTU_LAUNCH_XP(false, true, 2);
    else RTU_LAUNCH_XP(false, false, 3);
#undef RTU_LAUNCH_XP
}

void launch_extend_primary(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, int s0, int s1,
                           const WaveBuffers &B, float4 *pixel_accum, float4 *accum, unsigned *work_counter)
{
    static int occ = 0;
    if (extend_mode() == 1 && S.pool_ok && S.n_top == 0) { // many-node scenes: plain kernels with top-level nomination
        launch_extend_pooled<true>(cfg, st, S, F, s0, s1, B.q[1], B.aux[1], B, pixel_accum, accum, work_counter);
        return;
    }
    if (S.n_top > 0 && extend_mode() == 1 && !(F.flags & RTU_FLAG_REFERENCE_WALK) && top_mode() == 1) {
        static int occ_b = 0;
        k_extend_bvh<true><<<bvh_grid(cfg, k_extend_bvh<true>, &occ_b), WAVE_THREADS, 0, st>>>(S, F, s0, s1, B.q[1], B.aux[1], B.hits, pixel_accum, accum,
                                                                                               B.counters, work_counter);
        return;
    }
    if (S.n_top > 0 && extend_mode() == 1) {
        static int occ_t = 0;
        const size_t tsmem = sizeof(TopWarp) * (WAVE_THREADS / 32);
        k_extend_top<true><<<pooled_grid(cfg, k_extend_top<true>, tsmem, &occ_t), WAVE_THREADS, tsmem, st>>>(
            S, F, s0, s1, B.q[1], B.aux[1], B.hits, pixel_accum, accum, B.counters, work_counter);
        return;
    }
    k_extend<true><<<resident_grid(cfg, k_extend<true>, &occ), WAVE_THREADS, 0, st>>>(S, F, s0, s1, B.q[1], B.aux[1], B.hits, pixel_accum,
                                                                                     accum, B.counters, work_counter);
}

void launch_shade_primary(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, int s0,
                          const WaveBuffers &B, int out_q, float4 *accum, unsigned *work_counter)
{
    static int occ = 0;
    WaveOut O = make_out(B, out_q, accum);
    k_shade<true><<<shade_grid(cfg, k_shade<true>, &occ), WAVE_THREADS, 0, st>>>(S, F, s0, B.q[1 - out_q], B.aux[1 - out_q], B.hits, O,
                                                                                   work_counter, B.gi_count);
}

void launch_extend_queue(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, const WaveBuffers &B,
                         int in_q, float4 *accum, unsigned *work_counter)
{
    static int occ = 0;
    if (extend_mode() == 1 && S.pool_ok && S.n_top == 0) { // many-node scenes: plain kernels with top-level nomination
        launch_extend_pooled<false>(cfg, st, S, F, 0, 0, B.q[in_q], B.aux[in_q], B, accum, accum, work_counter);
        return;
    }
    if (S.n_top > 0 && extend_mode() == 1 && !(F.flags & RTU_FLAG_REFERENCE_WALK) && top_mode() == 1) {
        static int occ_b = 0;
        k_extend_bvh<false><<<bvh_grid(cfg, k_extend_bvh<false>, &occ_b), WAVE_THREADS, 0, st>>>(S, F, 0, 0, B.q[in_q], B.aux[in_q], B.hits, accum, accum,
                                                                                                B.counters, work_counter);
        return;
    }
    if (S.n_top > 0 && extend_mode() == 1) {
        static int occ_t = 0;
        const size_t tsmem = sizeof(TopWarp) * (WAVE_THREADS / 32);
        k_extend_top<false><<<pooled_grid(cfg, k_extend_top<false>, tsmem, &occ_t), WAVE_THREADS, tsmem, st>>>(
            S, F, 0, 0, B.q[in_q], B.aux[in_q], B.hits, accum, accum, B.counters, work_counter);
        return;
    }
    k_extend<false><<<resident_grid(cfg, k_extend<false>, &occ), WAVE_THREADS, 0, st>>>(S, F, 0, 0, B.q[in_q], B.aux[in_q], B.hits, accum,
                                                                                       accum, B.counters, work_counter);
}

void launch_shade_queue(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, const WaveBuffers &B,
                        int in_q, float4 *accum, unsigned *work_counter)
{
    static int occ = 0;
    WaveOut O = make_out(B, 1 - in_q, accum);
    k_shade<false><<<shade_grid(cfg, k_shade<false>, &occ), WAVE_THREADS, 0, st>>>(S, F, 0, B.q[in_q], B.aux[in_q], B.hits, O, work_counter,
                                                                                     B.gi_count);
}

template <class K> static int shadow_grid(const LaunchCfg &cfg, K kernel, size_t smem, int *cache)
{
    if (*cache == 0) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, WAVE_THREADS, smem) != cudaSuccess || n < 1) n = 1;
        if (cfg.blocks_per_sm > 0 && n > cfg.blocks_per_sm) n = cfg.blocks_per_sm;
        *cache = n;
    }
    return cfg.sm_count * *cache;
}

void launch_shadow_wave(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const WaveBuffers &B, float4 *accum,
                        unsigned *work_counter, bool reference_walk)
{
    static int occ[4] = {0, 0, 0, 0}, occ_simple = 0, mode = -1;
    if (mode < 0) { // RTU_SHADOW_KERNEL=simple selects the plain kernel (A/B measurements)
        const char *e = getenv("RTU_SHADOW_KERNEL");
        mode = (e && e[0] == 's') ? 0 : 1;
    }
    if (S.any_no_ref) reference_walk = false; // a mesh without cyBVH has nothing else to walk
    if (mode == 1 && S.pool_ok && S.n_top == 0) {
        const size_t smem = sizeof(SpWarpT<false>) * (WAVE_THREADS / 32), smem_occ = sizeof(SpWarpT<true>) * (WAVE_THREADS / 32);
        // <FLAT, OCC>: OCC walks the meshes' any-hit hierarchies; the other instantiation walks the cyBVH with the reference's
        // own box tests (RTU_FLAG_REFERENCE_WALK: it books what ShadowTrace's walk tests, up to the any-hit early out)
        if (S.flat && !reference_walk) k_shadow_wave<true, true><<<shadow_grid(cfg, k_shadow_wave<true, true>, smem_occ, &occ[0]), WAVE_THREADS, smem_occ, st>>>(S, B.shadow, accum, B.counters, work_counter);
        else if (S.flat) k_shadow_wave<true, false><<<shadow_grid(cfg, k_shadow_wave<true, false>, smem, &occ[1]), WAVE_THREADS, smem, st>>>(S, B.shadow, accum, B.counters, work_counter);
        else if (!reference_walk) k_shadow_wave<false, true><<<shadow_grid(cfg, k_shadow_wave<false, true>, smem_occ, &occ[2]), WAVE_THREADS, smem_occ, st>>>(S, B.shadow, accum, B.counters, work_counter);
        else k_shadow_wave<false, false><<<shadow_grid(cfg, k_shadow_wave<false, false>, smem, &occ[3]), WAVE_THREADS, smem, st>>>(S, B.shadow, accum, B.counters, work_counter);
    } else if (S.n_top > 0 && mode == 1 && !reference_walk && top_mode() == 1) {
        static int occ_b = 0;
        k_shadow_wave_bvh<<<bvh_grid(cfg, k_shadow_wave_bvh, &occ_b), WAVE_THREADS, 0, st>>>(S, B.shadow, accum, B.counters, work_counter);
    } else if (S.n_top > 0 && mode == 1) {
        static int occ_t = 0;
        const size_t tsmem = sizeof(TopWarp) * (WAVE_THREADS / 32);
        k_shadow_wave_top<<<shadow_grid(cfg, k_shadow_wave_top, tsmem, &occ_t), WAVE_THREADS, tsmem, st>>>(S, B.shadow, accum, B.counters, work_counter);
    } else {
        k_shadow_wave_simple<<<resident_grid(cfg, k_shadow_wave_simple, &occ_simple), WAVE_THREADS, 0, st>>>(S, B.shadow, accum, B.counters, work_counter);
    }
}

// Folds the GI records of one chunk into the pixel accumulator: L = End; L = D_k + A_k * L for k = end-1 .. 0
// (the recursion of MonteCarlo(), RenderFunctions.cpp:454-591, unrolled; see k_shade).
__global__ void k_gi_combine(const float4 *gi, const unsigned *count, unsigned cap, int gi_bounces, float4 *accum)
{
    unsigned n = *count;
    if (n > cap) n = cap;
    const int end_slot = 2 * (gi_bounces + 1);
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const float4 *rec = gi + (size_t)r * (end_slot + 2);
        float4 e = rec[end_slot];
        int end = (int)e.w;
        if (end > gi_bounces + 1) end = gi_bounces + 1;
        Col L = mk(e.x, e.y, e.z);
        for (int k = end - 1; k >= 0; k--) {
            float4 a = rec[2 * k], d = rec[2 * k + 1];
            L = mk(d.x, d.y, d.z) + mk(a.x, a.y, a.z) * L;
        }
        accum_add(accum, __float_as_int(rec[end_slot + 1].x), L);
    }
}

void launch_gi_combine(cudaStream_t st, const float4 *gi, const unsigned *count, unsigned cap, int gi_bounces, float4 *accum)
{
    k_gi_combine<<<1184, 256, 0, st>>>(gi, count, cap, gi_bounces, accum);
}

void launch_reset_counts(cudaStream_t st, unsigned *a, unsigned *b, unsigned *c, unsigned *d, unsigned *log_dst, const unsigned *log_src)
{
    k_reset_counts<<<1, 1, 0, st>>>(a, b, c, d, log_dst, log_src);
}

// waves [w0, n_waves) of a chunk as one cooperative launch; false: the device cannot (the caller launches them one by one)
bool launch_tail_waves(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, const WaveBuffers &B, int in_q, int w0,
                       int n_waves, float4 *target, unsigned *work, unsigned *wave_log)
{
    static int occ = -1;
    if (occ < 0) {
        int n = 0, coop = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tail_waves, WAVE_THREADS, 0) != cudaSuccess || n < 1) n = 0;
        if (n > EXT_BLOCKS) n = EXT_BLOCKS;
        occ = n;
    }
    if (occ == 0) return false;
    TailArgs A;
    A.q[0] = B.q[0]; A.q[1] = B.q[1];
    A.aux[0] = B.aux[0]; A.aux[1] = B.aux[1];
    A.shadow = B.shadow;
    A.hits = B.hits;
    A.counters = B.counters;
    A.gi_count = B.gi_count;
    A.work = work;
    A.wave_log = wave_log;
    A.target = target;
    A.in_q = in_q; A.w0 = w0; A.n_waves = n_waves;
    DScene S2 = S;
    FrameSetup F2 = F;
    void *args[3] = {(void *)&S2, (void *)&F2, (void *)&A};
    return cudaLaunchCooperativeKernel((const void *)k_tail_waves, dim3(cfg.sm_count * occ), dim3(WAVE_THREADS), args, 0, st) == cudaSuccess;
}

void launch_primary_ids(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const DCamera &cam, float *z, int *node,
                        int *face, DCounters *counters)
{
    static int occ = 0;
    k_primary_ids<<<resident_grid(cfg, k_primary_ids, &occ), WAVE_THREADS, 0, st>>>(S, cam, z, node, face, counters);
}

void launch_trace_batch(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const rtu_ray *rays, long long n,
                        rtu_hit *hits, DCounters *counters)
{
    static int occ = 0;
    k_trace_batch<<<resident_grid(cfg, k_trace_batch, &occ), WAVE_THREADS, 0, st>>>(S, rays, n, hits, counters);
}

void launch_shadow_batch(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const rtu_ray *rays, const float *tmax,
                         long long n, unsigned char *occl, DCounters *counters)
{
    static int occ = 0;
    k_shadow_batch<<<resident_grid(cfg, k_shadow_batch, &occ), WAVE_THREADS, 0, st>>>(S, rays, tmax, n, occl, counters);
}

void launch_trace_batch_wave(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const rtu_ray *rays, long long n, rtu_hit *hits,
                             const WaveBuffers &B, float4 *scratch_accum, unsigned *work_counter, bool reference_walk)
{
    FrameSetup F;
    memset(&F, 0, sizeof F);
    F.flags = reference_walk ? RTU_FLAG_REFERENCE_WALK : 0u;
    k_fill_ray_queue<<<148 * 4, 256, 0, st>>>(rays, n, B.q[0], hits);
    launch_extend_queue(cfg, st, S, F, B, 0, scratch_accum, work_counter);
    k_hits_to_records<<<148 * 2, WAVE_THREADS, 0, st>>>(S, B.q[0], B.hits, hits);
}

void launch_shadow_batch_wave(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const rtu_ray *rays, const float *tmax,
                              long long n, unsigned char *occl, const WaveBuffers &B, float4 *accum, unsigned *work_counter)
{
    k_fill_shadow_queue<<<148 * 4, 256, 0, st>>>(rays, tmax, n, B.shadow);
    launch_shadow_wave(cfg, st, S, B, accum, work_counter, false);
    k_occluded_from_accum<<<148 * 4, 256, 0, st>>>(accum, n, occl);
}

void launch_shade_batch(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, const rtu_ray *rays,
                        const rtu_hit *hits, long long n, const WaveBuffers &B, int out_q, float4 *accum)
{
    static int occ = 0;
    WaveOut O = make_out(B, out_q, accum);
    k_shade_first<<<resident_grid(cfg, k_shade_first, &occ), WAVE_THREADS, 0, st>>>(S, F, rays, hits, n, O);
}

// Self-test of div_hoisted(): thread t owns divisor mantissa t (all 2^23 of them); for a spread of
// divisor exponents inside the window and `per_thread` hashed numerators (plus the all-zero /
// all-one mantissas) it compares the hoisted quotient with the compiler's IEEE `a / b`, bit for bit.
__global__ void k_selftest_div(unsigned per_thread, unsigned long long seed, unsigned long long *mismatch, unsigned long long *tested)
{
    unsigned man = blockIdx.x * blockDim.x + threadIdx.x;
    if (man >= (1u << 23)) return;
    const int bexp[9] = {-40, -20, -3, -1, 0, 1, 3, 20, 40};
    unsigned long long bad = 0, n = 0;
    unsigned long long h = seed ^ (0x9E3779B97F4A7C15ULL * (man + 1));
    for (int e = 0; e < 9; e++) {
        float b = __uint_as_float(((unsigned)(127 + bexp[e]) << 23) | man);
        if (e & 1) b = -b;
        float y = rcp_refined(b);
        for (unsigned k = 0; k < per_thread; k++) {
            h ^= h >> 12; h ^= h << 25; h ^= h >> 27;
            unsigned long long r = h * 2685821657736338717ULL;
            unsigned am = (unsigned)(r >> 41);
            if (k == 0) am = 0;
            if (k == 1) am = 0x7fffffu;
            if (k == 2) am = man;
            int ae = (int)((r >> 8) % 121u) - 60;
            float a = __uint_as_float(((unsigned)(127 + ae) << 23) | am | ((unsigned)(r & 1u) << 31));
            float q = div_hoisted(a, b, y);
            float d = a / b;
            bad += __float_as_uint(q) != __float_as_uint(d);
            n++;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(mismatch, bad);
        atomicAdd(tested, n);
    }
}

void launch_selftest_div(cudaStream_t st, unsigned per_thread, unsigned long long seed, unsigned long long *mismatch,
                         unsigned long long *tested)
{
    k_selftest_div<<<(1u << 23) / 256, 256, 0, st>>>(per_thread, seed, mismatch, tested);
}

void launch_camera_rays(cudaStream_t st, const DCamera &cam, float ox, float oy, rtu_ray *rays)
{
    int n = cam.width * cam.height;
    k_camera_rays<<<(n + 255) / 256, 256, 0, st>>>(cam, ox, oy, rays);
}

void launch_resolve(cudaStream_t st, const float4 *accum, int npix, float, int spp, float *rgb, unsigned char *rgb8)
{
    k_resolve<<<(npix + 255) / 256, 256, 0, st>>>(accum, npix, spp, rgb, rgb8);
}

void launch_accum_add(cudaStream_t st, float4 *dst, const float4 *src, size_t npix)
{
    k_accum_add<<<148 * 8, 256, 0, st>>>(dst, src, npix);
}

void launch_pack_rgb(cudaStream_t st, const float4 *accum, size_t npix, float *planes)
{
    k_pack_rgb<<<148 * 8, 256, 0, st>>>(accum, npix, planes);
}

void launch_resolve_planes(cudaStream_t st, const float *planes, size_t npix, int spp, float *rgb, unsigned char *rgb8)
{
    k_resolve_planes<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(planes, npix, spp, rgb, rgb8);
}

void launch_adaptive_update(cudaStream_t st, const float4 *accum, int W, int H, int n_now, int max_spp, float target, unsigned char *tile_done,
                            int *tile_samples, unsigned *n_active)
{
    const int tiles = ((W + 7) >> 3) * ((H + 3) >> 2);
    cudaMemsetAsync(n_active, 0, sizeof(unsigned), st);
    k_adaptive_update<<<(tiles + 127) / 128, 128, 0, st>>>(accum, W, H, n_now, max_spp, target, tile_done, tile_samples, n_active);
}

void launch_resolve_adaptive(cudaStream_t st, const float4 *accum, int W, int H, const int *tile_samples, float *rgb, unsigned char *rgb8,
                             unsigned char *sample_count)
{
    const size_t npix = (size_t)W * H;
    k_resolve_adaptive<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(accum, W, H, tile_samples, rgb, rgb8, sample_count);
}

void launch_zimage(cudaStream_t st, const float *z, int npix, unsigned *minmax_bits, unsigned char *z8)
{
    static const unsigned init[2] = {0x7f800000u, 0u};
    cudaMemcpyAsync(minmax_bits, init, sizeof init, cudaMemcpyHostToDevice, st);
    k_zminmax<<<296, 256, 0, st>>>(z, npix, minmax_bits);
    k_zimage<<<(npix + 255) / 256, 256, 0, st>>>(z, npix, minmax_bits, z8);
}
