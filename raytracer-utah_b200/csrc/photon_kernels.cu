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
#include "rtu_internal.h"
#include "shade.cuh"
#include "camera.cuh"

#ifndef PGATHER_MIN_CTAS
#define PGATHER_MIN_CTAS 16 // same for k_photon_gather (final gathering traces its sample rays in the kernel)
#endif

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

__global__ void __launch_bounds__(128, GATHER_MIN_CTAS)
k_estimate(DPhotonMap PM, const float *pos, const float *normal, long long n, float radius, float norm_scale, float *irrad,
           float *direction, int *found)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Col e;
    float dx, dy, dz;
    int f;
    bool has_n = normal != nullptr;
    float nx = has_n ? normal[i * 3] : 0.f, ny = has_n ? normal[i * 3 + 1] : 0.f, nz = has_n ? normal[i * 3 + 2] : 0.f;
    estimate_irradiance(PM, pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2], has_n, nx, ny, nz, radius, norm_scale, e, dx, dy, dz, f);
    irrad[i * 3] = e.r; irrad[i * 3 + 1] = e.g; irrad[i * 3 + 2] = e.b;
    direction[i * 3] = dx; direction[i * 3 + 1] = dy; direction[i * 3 + 2] = dz;
    if (found) found[i] = f;
}

// PhotonMapping(ray, hInfo) (RenderFunctions.cpp:394-413): the estimate becomes a PhotonLight (lights.h:61-74: Illuminate =
// intensity, Direction = direction, not ambient, no shadow ray) and the hit is shaded with it alone, bounceCount 0.
// A hit without photons in range has direction 0/0 = NaN and shades to NaN, as in the reference.
__device__ __forceinline__ Col photon_mapping(const DScene &S, const DPhotonMap &PM, const HitRec &H)
{
    Col e;
    float dx, dy, dz;
    int found;
    estimate_irradiance(PM, H.px, H.py, H.pz, true, H.nx, H.ny, H.nz, PM.radius, PM.norm_scale, e, dx, dy, dz, found);
    norm3(dx, dy, dz); // PhotonLight::SetDirection normalises once more (lights.h:70)
    Col out = mk(0, 0, 0);
    if (H.material < 0) {
        out = mk(1, 1, 1);
    } else if (H.front) {
        const DMaterial &M = S.materials[H.material];
        Col Kd = texcolor_sample(S, M.diffuse, H.u, H.v, H.w);
        Col Ks = texcolor_sample(S, M.specular, H.u, H.v, H.w);
        float vx = S.cam_pos[0] - H.px, vy = S.cam_pos[1] - H.py, vz = S.cam_pos[2] - H.pz; // mtlFunctions.cpp:137
        norm3(vx, vy, vz);
        float lx = -dx, ly = -dy, lz = -dz;
        norm3(lx, ly, lz);
        float hx = vx + lx, hy = vy + ly, hz = vz + lz;
        norm3(hx, hy, hz);
        float ndl = dot3(H.nx, H.ny, H.nz, lx, ly, lz);
        float ndh = dot3(H.nx, H.ny, H.nz, hx, hy, hz);
        if (ndl < 0.f) ndl = 0.f;
        if (ndh < 0.f) ndh = 0.f;
        out = (e * ndl) * (Kd + Ks * powf(ndh, M.glossiness));
    }
    return out;
}

// RTU_MODE_PHOTON: PhotonMapping per primary hit of the hit queue.
__global__ void __launch_bounds__(128, GATHER_MIN_CTAS)
k_photon_shade(DScene S, FrameSetup F, int s0, HitQueue hq, DPhotonMap PM, float4 *accum)
{
    unsigned total = *hq.count;
    if (total > hq.cap) total = hq.cap;
    unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= total) return;
    float4 a = hq.a[h], b = hq.b[h];
    Best B;
    B.z = a.x; B.node = __float_as_int(a.y); B.front = __float_as_int(a.z); B.slot = __float_as_int(a.w);
    B.bc1 = b.x; B.bc2 = b.y; B.bc3 = b.z;
    unsigned idx = __float_as_uint(b.w);
    // the primary ray of work item idx (same mapping as k_extend<primary>)
    PrimaryMap pm;
    pm.init(F);
    int s, x, y;
    pm.decode(idx, s0, s, x, y);
    int pixel = y * pm.W + x;
    Ray ray = primary_ray(F, s, x, y, pixel);
    HitRec H;
    finalize_hit(S, ray, B, H);
    Col out = photon_mapping(S, PM, H);
    float *acc = reinterpret_cast<float *>(accum + pixel);
    atomicAdd(acc, out.r);
    atomicAdd(acc + 1, out.g);
    atomicAdd(acc + 2, out.b);
}

// RTU_MODE_PHOTON_GATHER: MonteCarloPhoton(hInfo, x, y, 1) per primary hit (RenderFunctions.cpp:416-451), added to the
// Whitted radiance the other kernels produce.  Up to gi_bounces cosine samples of the hemisphere of the SAME first hit; the
// HitInfo of the sample rays is never reset, so a later sample only finds what is nearer than the previous sample's hit;
// a sample that finds nothing adds the background and ends the loop; the sum is divided by the samples taken.
__global__ void __launch_bounds__(128, PGATHER_MIN_CTAS)
k_photon_gather(DScene S, FrameSetup F, int s0, HitQueue hq, DPhotonMap PM, float4 *accum, DCounters *counters)
{
    Tally tl = {0, 0, 0, 0, 0};
    unsigned total = *hq.count;
    if (total > hq.cap) total = hq.cap;
    unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < total) {
        float4 a = hq.a[h], b = hq.b[h];
        Best B;
        B.z = a.x; B.node = __float_as_int(a.y); B.front = __float_as_int(a.z); B.slot = __float_as_int(a.w);
        B.bc1 = b.x; B.bc2 = b.y; B.bc3 = b.z;
        unsigned idx = __float_as_uint(b.w);
        PrimaryMap pm;
        pm.init(F);
        int s, x, y;
        pm.decode(idx, s0, s, x, y);
        int pixel = y * pm.W + x;
        Ray ray = primary_ray(F, s, x, y, pixel);
        HitRec H;
        finalize_hit(S, ray, B, H);
        Rng rng;
        rng.key = F.seed; rng.pixel = 0x50474154u; rng.path = primary_path(pixel, s); rng.dim = 0;
        Best Bs;
        Bs.z = RTU_BIG; Bs.node = -1; Bs.front = 1; Bs.slot = 0; Bs.bc1 = Bs.bc2 = Bs.bc3 = 0.f;
        Col sum = mk(0, 0, 0);
        int actual = 0;
        for (int bnc = 0; bnc < F.gi_bounces; bnc++) {
            Ray sr;
            sr.px = H.px; sr.py = H.py; sr.pz = H.pz;
            sample_hemi_cos(rng, H.nx, H.ny, H.nz, sr.dx, sr.dy, sr.dz);
            norm3(sr.dx, sr.dy, sr.dz);
            actual++;
            tl.trace++;
            if (scene_hit<false>(S, sr, Bs, tl, false)) {
                HitRec Hs;
                finalize_hit(S, sr, Bs, Hs);
                sum = sum + photon_mapping(S, PM, Hs);
            } else {
                sum = sum + background_sample(S, x, y, pm.W, F.cam.height);
                break;
            }
        }
        if (actual > 0) {
            float n = (float)actual;
            float *acc = reinterpret_cast<float *>(accum + pixel);
            atomicAdd(acc, sum.r / n);
            atomicAdd(acc + 1, sum.g / n);
            atomicAdd(acc + 2, sum.b / n);
        }
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
            if (!scene_hit<false>(S, ray, B, tl, false)) alive = false;
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
void launch_estimate(cudaStream_t st, const DPhotonMap &PM, const float *pos, const float *normal, long long n, float radius,
                     float norm_scale, float *irrad, float *direction, int *found)
{
    if (n <= 0) return;
    k_estimate<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(PM, pos, normal, n, radius, norm_scale, irrad, direction, found);
}

void launch_photon_shade(cudaStream_t st, const DScene &S, const FrameSetup &F, int s0, const WaveBuffers &B, unsigned max_hits,
                         const DPhotonMap &PM, float4 *accum)
{
    if (max_hits == 0) return;
    k_photon_shade<<<(max_hits + 127) / 128, 128, 0, st>>>(S, F, s0, B.hits, PM, accum);
}

void launch_photon_gather(cudaStream_t st, const DScene &S, const FrameSetup &F, int s0, const WaveBuffers &B, unsigned max_hits,
                          const DPhotonMap &PM, float4 *accum)
{
    if (max_hits == 0) return;
    k_photon_gather<<<(max_hits + 127) / 128, 128, 0, st>>>(S, F, s0, B.hits, PM, accum, B.counters);
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
