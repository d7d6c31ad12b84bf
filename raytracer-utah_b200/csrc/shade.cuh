// MtlBlinn::Shade, Light::Illuminate and texture sampling on the device.
//
// The recursive CPU formulation (mtlFunctions.cpp:120-298) becomes one non-recursive step per
// hit: direct light turns into shadow-queue entries carrying "radiance to add if unoccluded",
// and each reflection / refraction branch turns into a ray of the next wave carrying the
// product of all factors the recursion would have multiplied onto its result.  Geometry
// (directions, origins) is evaluated with the reference's operation order so that secondary
// rays hit the same primitives; colour factors are within float rounding of the reference.
#pragma once
#ifndef ACCUM_VECTOR_RED
#define ACCUM_VECTOR_RED 1
#endif

#include "intersect.cuh"

struct Col {
    float r, g, b;
};
__device__ __forceinline__ Col mk(float r, float g, float b) { Col c; c.r = r; c.g = g; c.b = b; return c; }
__device__ __forceinline__ Col operator*(Col a, Col b) { return mk(a.r * b.r, a.g * b.g, a.b * b.b); }
__device__ __forceinline__ Col operator*(Col a, float s) { return mk(a.r * s, a.g * s, a.b * s); }
__device__ __forceinline__ Col operator+(Col a, Col b) { return mk(a.r + b.r, a.g + b.g, a.b + b.b); }
__device__ __forceinline__ bool nonblack(Col a) { return a.r != 0.f || a.g != 0.f || a.b != 0.f; } // cyColor.h operator!=

// ------------------------------------------------------------------ counter-based RNG
// Philox4x32-10 (Salmon et al. 2011).  key = frame seed; counter = (pixel, path word, dimension, 0).
// The reference draws from libc rand() seeded by wall clock (RenderFunctions.cpp:60), so only
// the distributions, not the streams, are comparable.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int i = 0; i < 10; i++) {
        unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
struct Rng {
    uint2 key;
    unsigned pixel, path, dim;
    __device__ __forceinline__ float4 next4()
    {
        uint4 r = philox4x32(make_uint4(pixel, path, dim++, 0u), key);
        const float s = 1.0f / 4294967296.0f;
        // (x + 0.5) / 2^32 in (0,1): the reference's rand()/RAND_MAX is in [0,1]
        return make_float4((r.x + 0.5f) * s, (r.y + 0.5f) * s, (r.z + 0.5f) * s, (r.w + 0.5f) * s);
    }
};

// SampleSphere (RenderFunctions.cpp:282-302): uniform point in the ball of the given radius by
// rejection from the cube.  radius == 0 gives exactly (0,0,0), as the reference's rand()/inf does.
__device__ __forceinline__ void sample_ball(Rng &rng, float radius, float &x, float &y, float &z)
{
    x = y = z = 0.f;
    if (!(radius > 0.f)) return;
    for (int it = 0; it < 64; it++) {
        float4 u = rng.next4();
        x = -radius + u.x * (radius * 2);
        y = -radius + u.y * (radius * 2);
        z = -radius + u.z * (radius * 2);
        if (!(sqrtf(dot3(x, y, z, x, y, z)) > radius)) return;
    }
}

// SampleHemiSphereCosine(origin, normal, 1) (RenderFunctions.cpp:320-337): theta = acos(1-2u)/2, phi = 2 pi v,
// tangent frame built from normal x (u,u,u) (SURVEY A-14).  Returns the (unnormalised) offset.
__device__ __forceinline__ void sample_hemi_cos(Rng &rng, float nx, float ny, float nz, float &ox, float &oy, float &oz)
{
    float4 u = rng.next4();
    float sx = u.x, phi = u.y * 6.283185307179586f;
    float theta = 0.5f * acosf(1.f - 2.f * sx);
    float ax = ny * sx - nz * sx, ay = nz * sx - nx * sx, az = nx * sx - ny * sx; // N x (s,s,s)
    float l1 = sqrtf(dot3(ax, ay, az, ax, ay, az));
    ax /= l1; ay /= l1; az /= l1;
    float bx = ay * nz - az * ny, by = az * nx - ax * nz, bz = ax * ny - ay * nx; // v1 x N
    float l2 = sqrtf(dot3(bx, by, bz, bx, by, bz));
    bx /= l2; by /= l2; bz /= l2;
    float ct = cosf(theta), st = sinf(theta), cp = cosf(phi), sp = sinf(phi);
    ox = (nx * ct + ax * (st * cp)) + bx * (st * sp);
    oy = (ny * ct + ay * (st * cp)) + by * (st * sp);
    oz = (nz * ct + az * (st * cp)) + bz * (st * sp);
}

// ------------------------------------------------------------------ textures
// Texture::TileClamp (scene.h:355-365)
__device__ __forceinline__ float tile_clamp(float v)
{
    float u = v - (float)(int)v;
    if (u < 0) u += 1;
    return u;
}

// TextureMap::Sample -> TextureFile::Sample / TextureChecker::Sample (scene.h:382, texture.cpp:95-133)
// One copy per kernel (not inlined): the bilinear fetch with its twelve IEEE divisions is ~350 instructions, and the shading
// kernels are bound by instruction fetch once warps spread over many such copies.
static __device__ __noinline__ Col texmap_sample(const DTexMap &T, float u, float v, float w)
{
    if (T.kind == 0) return mk(0, 0, 0); // texture == NULL (failed load)
    // Transformation::TransformTo (scene.h:235)
    float qx = u - T.pos[0], qy = v - T.pos[1], qz = w - T.pos[2];
    float tu = qx * T.itm[0] + qy * T.itm[3] + qz * T.itm[6];
    float tv = qx * T.itm[1] + qy * T.itm[4] + qz * T.itm[7];
    float cu = tile_clamp(tu), cv = tile_clamp(tv);
    if (T.kind == 1) {
        if (cu <= 0.5f) return cv <= 0.5f ? mk(T.c1[0], T.c1[1], T.c1[2]) : mk(T.c2[0], T.c2[1], T.c2[2]);
        return cv <= 0.5f ? mk(T.c2[0], T.c2[1], T.c2[2]) : mk(T.c1[0], T.c1[1], T.c1[2]);
    }
    int W = T.width, H = T.height;
    if (W + H == 0) return mk(0, 0, 0);
    float x = W * cu, y = H * cv;
    int ix = (int)x, iy = (int)y;
    float fx = x - ix, fy = y - iy;
    if (ix < 0) ix -= (ix / W - 1) * W;
    if (ix >= W) ix -= (ix / W) * W;
    int ixp = ix + 1;
    if (ixp >= W) ixp -= W;
    if (iy < 0) iy -= (iy / H - 1) * H;
    if (iy >= H) iy -= (iy / H) * H;
    int iyp = iy + 1;
    if (iyp >= H) iyp -= H;
    const uint8_t *d = T.rgb8;
    const uint8_t *p00 = d + 3 * ((size_t)iy * W + ix), *p01 = d + 3 * ((size_t)iy * W + ixp);
    const uint8_t *p10 = d + 3 * ((size_t)iyp * W + ix), *p11 = d + 3 * ((size_t)iyp * W + ixp);
    float w00 = (1 - fx) * (1 - fy), w01 = fx * (1 - fy), w10 = (1 - fx) * fy, w11 = fx * fy;
    Col c;
    c.r = (__ldg(p00) / 255.0f) * w00 + (__ldg(p01) / 255.0f) * w01 + (__ldg(p10) / 255.0f) * w10 + (__ldg(p11) / 255.0f) * w11;
    c.g = (__ldg(p00 + 1) / 255.0f) * w00 + (__ldg(p01 + 1) / 255.0f) * w01 + (__ldg(p10 + 1) / 255.0f) * w10 + (__ldg(p11 + 1) / 255.0f) * w11;
    c.b = (__ldg(p00 + 2) / 255.0f) * w00 + (__ldg(p01 + 2) / 255.0f) * w01 + (__ldg(p10 + 2) / 255.0f) * w10 + (__ldg(p11 + 2) / 255.0f) * w11;
    return c;
}

// TexturedColor::Sample (scene.h:421)
__device__ __forceinline__ Col texcolor_sample(const DScene &S, const DTexColor &tc, float u, float v, float w)
{
    Col c = mk(tc.c[0], tc.c[1], tc.c[2]);
    if (tc.map < 0) return c;
    return c * texmap_sample(S.texmaps[tc.map], u, v, w);
}

// TexturedColor::SampleEnvironment (scene.h:425-431)
__device__ __forceinline__ Col environment_sample(const DScene &S, float dx, float dy, float dz)
{
    if (S.environment.map < 0) return mk(S.environment.c[0], S.environment.c[1], S.environment.c[2]);
    float z = asinf(-dz) / 3.14159265358979323846f + 0.5f;
    float den = fabsf(dx) + fabsf(dy);
    float x = dx / den, y = dy / den;
    float ax = 0.5f * x + (-0.5f) * y;
    float ay = 0.5f * x + 0.5f * y;
    return texcolor_sample(S, S.environment, 0.5f + ax * z, 0.5f + ay * z, 0.0f + 0.f * z);
}

// background.Sample(Point3((float)x/W, (float)y/H, 0)) (RenderFunctions.cpp:145,164)
__device__ __forceinline__ Col background_sample(const DScene &S, int x, int y, int W, int H)
{
    if (S.background.map < 0) return mk(S.background.c[0], S.background.c[1], S.background.c[2]);
    return texcolor_sample(S, S.background, (float)x / W, (float)y / H, 0.f);
}

// ------------------------------------------------------------------ queues
__device__ __forceinline__ unsigned warp_alloc(unsigned *counter, bool want)
{
    // warp-aggregated append: one atomic per warp, lanes get consecutive slots
    unsigned mask = __ballot_sync(__activemask(), want);
    if (!want) return 0xffffffffu;
    unsigned lane = threadIdx.x & 31u;
    unsigned leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

// Split-phase, warp-aggregated queue allocation.  issue(): one atomic per warp reserves n slots for every participating
// lane; its result stays in the leader's register.  slot(): fetched (one shuffle) only where the first slot is written,
// so the atomic's round trip to L2 overlaps whatever is computed in between.  Lane's k-th slot = slot() + k * stride.
struct SlotTicket {
    unsigned mask, leader, base_in_leader, rank, stride;
    __device__ __forceinline__ void issue(unsigned *counter, unsigned n)
    {
        mask = __activemask();
        const unsigned lane = threadIdx.x & 31u;
        leader = __ffs(mask) - 1;
        stride = __popc(mask);
        rank = __popc(mask & ((1u << lane) - 1u));
        base_in_leader = 0;
        // inline PTX: nvcc rewrites a plain atomicAdd() here into its own warp-aggregated form, whose shuffle consumes the
        // result at once and so waits out the whole round trip on the spot
        if (lane == leader)
            asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(base_in_leader) : "l"(counter), "r"(stride * n) : "memory");
    }
    // must be called by exactly the lanes that called issue() (still converged or reconverged)
    __device__ __forceinline__ unsigned slot() const { return __shfl_sync(mask, base_in_leader, leader) + rank; }
};

struct WaveOut {
    RayQueue next;
    AuxPool aux;
    ShadowQueue shadow;
    float4 *accum;        // W*H float4: rgb sum, .w unused here
    DCounters *counters;
};

// One 16-byte vector reduction per contribution (red.global.add.v4.f32) instead of three scalar ones: a third of the L2
// atomic traffic.  The .w lane only ever receives +0: no accumulator keeps anything but zero there (the pixel index of a GI
// record lives in the record's last slot, which is never added to).
__device__ __forceinline__ void accum_add(float4 *accum, int pixel, Col c)
{
#if ACCUM_VECTOR_RED
    if (c.r != 0.f || c.g != 0.f || c.b != 0.f) atomicAdd(accum + pixel, make_float4(c.r, c.g, c.b, 0.f));
#else
    float *a = reinterpret_cast<float *>(accum + pixel);
    if (c.r != 0.f) atomicAdd(a, c.r);
    if (c.g != 0.f) atomicAdd(a + 1, c.g);
    if (c.b != 0.f) atomicAdd(a + 2, c.b);
#endif
}

// ray meta word: kind (3 bits) | bounceCount left (4) | tree (1) | GI depth (3) | parent material (21)
//   tree 0: Shade(..., lights, ...)            -> radiance goes to the ray's target slot
//   tree 1: Shade(..., {AmbientLight c}, ...)  -> the factor of c goes to the target slot, the
//           c-independent environment terms to target+1 (RTU_MODE_PATH only; see k_shade)
__device__ __forceinline__ unsigned pack_meta(int kind, int bounce, int material, int tree = 0, int gidepth = 0)
{
    return (unsigned)kind | ((unsigned)bounce << 3) | ((unsigned)tree << 7) | ((unsigned)gidepth << 8) |
           ((unsigned)(material & 0x1fffff) << 11);
}
__device__ __forceinline__ void unpack_meta(unsigned meta, int &kind, int &bounce, int &tree, int &gidepth, int &material)
{
    kind = (int)(meta & 7u);
    bounce = (int)((meta >> 3) & 15u);
    tree = (int)((meta >> 7) & 1u);
    gidepth = (int)((meta >> 8) & 7u);
    material = (int)(meta >> 11);
}

__device__ __forceinline__ void push_ray(const WaveOut &O, float ox, float oy, float oz, float dx, float dy, float dz,
                                         Col w, int pixel, unsigned meta, int aux, unsigned path)
{
    unsigned slot = warp_alloc(O.next.count, true);
    if (slot >= O.next.cap) { O.counters->overflow = 1; return; }
    O.next.o[slot] = make_float4(ox, oy, oz, __int_as_float(pixel));
    O.next.d[slot] = make_float4(dx, dy, dz, __uint_as_float(meta));
    O.next.w[slot] = make_float4(w.r, w.g, w.b, __int_as_float(aux));
    O.next.path[slot] = path;
}

__device__ __forceinline__ void push_shadow(const WaveOut &O, float ox, float oy, float oz, float dx, float dy, float dz,
                                            float tmax, Col c, int pixel)
{
    unsigned slot = warp_alloc(O.shadow.count, true);
    if (slot >= O.shadow.cap) { O.counters->overflow = 1; return; }
    O.shadow.o[slot] = make_float4(ox, oy, oz, __int_as_float(pixel));
    O.shadow.d[slot] = make_float4(dx, dy, dz, tmax);
    O.shadow.c[slot] = make_float4(c.r, c.g, c.b, 0.f);
}

__device__ __forceinline__ void norm3(float &x, float &y, float &z)
{
    float len = sqrtf(dot3(x, y, z, x, y, z));
    x = x / len; y = y / len; z = z / len;
}

// RNG path word of a child ray: a mixing step keeps siblings, parents and other samples apart
__device__ __forceinline__ unsigned child_path(unsigned path, unsigned code)
{
    return (path ^ (code * 0x9E3779B9u)) * 0x85EBCA6Bu + 0xC2B2AE35u;
}

struct ShadeParams {
    unsigned flags;
    uint2 seed;
};

// One MtlBlinn::Shade(ray, hInfo, lights, bounce) step with incoming throughput Wt.
//   dirx..: ray.dir (world).  H: the hit.  bounce: bounceCount.  path: RNG path word of this ray.
// Direct light  -> shadow queue (mtlFunctions.cpp:125-155, lightFunctions.cpp:27-84, lights.h:32,48)
// Refraction    -> RK_REFRACT / RK_TIR rays (:160-270);  Reflection -> RK_REFLECT rays (:273-291)
//   pixel: the accumulator slot the radiance is added to (a pixel, or a slot of a GI record).
//   tree 1: the light list is the single AmbientLight MonteCarlo() built (RenderFunctions.cpp:587-590)
//           with unit intensity; the caller multiplies the accumulated factor by its intensity later.
__device__ __forceinline__ void shade_hit(const DScene &S, const ShadeParams &P, const WaveOut &O, float dirx, float diry,
                                          float dirz, const HitRec &H, Col Wt, int bounce, int pixel, unsigned path,
                                          int tree = 0, bool fresh_slot = false)
{
    if (H.material < 0) { // node without material: the reference would dereference NULL
        if (fresh_slot) {
            O.accum[pixel] = make_float4(Wt.r, Wt.g, Wt.b, 0.f);
        } else {
            accum_add(O.accum, pixel, Wt);
        }
        return;
    }
    const DMaterial &M = S.materials[H.material];
    Rng rng;
    rng.key = P.seed; rng.pixel = tree ? 0x7A11u : 0u; rng.path = path; rng.dim = 0;
    Col local = mk(0, 0, 0);
    Col Kd = mk(0, 0, 0);
    if (H.front) Kd = texcolor_sample(S, M.diffuse, H.u, H.v, H.w);
    if (H.front && tree) {
        local = Kd; // Kd * Illuminate() of the unit ambient light (:131-133)
    } else if (H.front) {
        Col Ks = texcolor_sample(S, M.specular, H.u, H.v, H.w);
        // viewDirection uses camera.pos, not the ray origin (mtlFunctions.cpp:137, SURVEY A-6)
        // every non-ambient light sends one shadow ray (unless null rays are culled): their queue slots are reserved
        // up front with one atomic per warp, whose latency then hides behind the Blinn terms
        const bool reserve = !(P.flags & 1u) && S.n_shadow_lights > 0;
        SlotTicket st;
        unsigned sslot = 0, sk = 0;
        if (reserve) st.issue(O.shadow.count, (unsigned)S.n_shadow_lights);
        float vx = S.cam_pos[0] - H.px, vy = S.cam_pos[1] - H.py, vz = S.cam_pos[2] - H.pz;
        norm3(vx, vy, vz);
        for (int i = 0; i < S.n_lights; i++) {
            const DLight &L = S.lights[i];
            if (L.kind == 0) {
                local = local + Kd * mk(L.I[0], L.I[1], L.I[2]);
                continue;
            }
            float lx, ly, lz;       // lightDirection = (-Direction(p)).GetNormalized()
            float sx, sy, sz, tmax; // shadow ray
            Col illum;
            if (L.kind == 1) {
                lx = -L.v[0]; ly = -L.v[1]; lz = -L.v[2];
                norm3(lx, ly, lz);
                sx = -L.v[0]; sy = -L.v[1]; sz = -L.v[2];
                tmax = RTU_BIG;
                illum = mk(L.I[0], L.I[1], L.I[2]);
            } else {
                float ex = H.px - L.v[0], ey = H.py - L.v[1], ez = H.pz - L.v[2]; // Direction(p) = (p-position).GetNormalized()
                norm3(ex, ey, ez);
                lx = -ex; ly = -ey; lz = -ez;
                norm3(lx, ly, lz);
                float tx = L.v[0], ty = L.v[1], tz = L.v[2]; // point on the light the shadow ray aims at
                if (L.size > 0.f) {
                    // one random point on the disk of radius `size` facing p (lightFunctions.cpp:43-62)
                    float4 u = rng.next4();
                    float sr = u.x * L.size, th = u.y * 6.283185307179586f;
                    float offx = sr * cosf(th), offy = sr * sinf(th);
                    float nx = L.v[0] - H.px, ny = L.v[1] - H.py, nz = L.v[2] - H.pz;
                    norm3(nx, ny, nz);
                    float ax = ny * 1.f - nz * 0.f, ay = nz * 0.f - nx * 1.f, az = nx * 0.f - ny * 0.f; // N x (0,0,1)
                    norm3(ax, ay, az);
                    float bx = ay * nz - az * ny, by = az * nx - ax * nz, bz = ax * ny - ay * nx;       // v1 x N
                    norm3(bx, by, bz);
                    tx = (L.v[0] + ax * offx) + bx * offy;
                    ty = (L.v[1] + ay * offx) + by * offy;
                    tz = (L.v[2] + az * offx) + bz * offy;
                }
                sx = tx - H.px; sy = ty - H.py; sz = tz - H.pz;
                norm3(sx, sy, sz);
                float qx = H.px - tx, qy = H.py - ty, qz = H.pz - tz;
                tmax = sqrtf(dot3(qx, qy, qz, qx, qy, qz));
                float wx = L.v[0] - H.px, wy = L.v[1] - H.py, wz = L.v[2] - H.pz;
                float inv = 1 / dot3(wx, wy, wz, wx, wy, wz); // 1/d^2 falloff (lightFunctions.cpp:83)
                illum = mk(L.I[0], L.I[1], L.I[2]) * inv;
            }
            float hx = vx + lx, hy = vy + ly, hz = vz + lz;
            norm3(hx, hy, hz);
            float ndl = dot3(H.nx, H.ny, H.nz, lx, ly, lz);
            float ndh = dot3(H.nx, H.ny, H.nz, hx, hy, hz);
            if (ndl < 0.f) ndl = 0.f;
            if (ndh < 0.f) ndh = 0.f;
            Col c = (illum * ndl) * (Kd + Ks * powf(ndh, M.glossiness)); // :152
            c = c * Wt;
            bool null = !nonblack(c) || !(c.r == c.r) || !(c.g == c.g) || !(c.b == c.b);
            if ((P.flags & 1u) && null) continue; // RTU_FLAG_CULL_NULL_SHADOW_RAYS
            if (reserve) {
                if (sk == 0) sslot = st.slot(); // first use of the reservation: all reserving lanes arrive here together
                const unsigned slot = sslot + sk * st.stride;
                sk++;
                if (slot >= O.shadow.cap) { O.counters->overflow = 1; continue; }
                O.shadow.o[slot] = make_float4(H.px, H.py, H.pz, __int_as_float(pixel));
                O.shadow.d[slot] = make_float4(sx, sy, sz, tmax);
                O.shadow.c[slot] = make_float4(c.r, c.g, c.b, 0.f);
            } else {
                push_shadow(O, H.px, H.py, H.pz, sx, sy, sz, tmax, c, pixel);
            }
        }
    }
    if (fresh_slot) {
        // a GI vertex's own slot still holds the zeros it was opened with (its shadow rays and child rays come later):
        // 0 + v = v, so the sum is stored; an atomic here would be a DRAM read-modify-write, the record left L2 waves ago
        const Col v = local * Wt;
        O.accum[pixel] = make_float4(v.r, v.g, v.b, 0.f);
    } else {
        accum_add(O.accum, pixel, local * Wt);
    }

    if (bounce <= 0) return;
    Col Kt = texcolor_sample(S, M.refraction, H.u, H.v, H.w);
    if (nonblack(Kt)) {
        // first glossiness sample: sampledNormal = ((p+N) + offset - p).GetNormalized()  (:163-165)
        float ox, oy, oz;
        sample_ball(rng, M.refr_gloss, ox, oy, oz);
        float snx = ((H.px + H.nx) + ox) - H.px, sny = ((H.py + H.ny) + oy) - H.py, snz = ((H.pz + H.nz) + oz) - H.pz;
        norm3(snx, sny, snz);
        float cos1 = dot3(snx, sny, snz, -dirx, -diry, -dirz);
        float sin1 = (float)sqrt(1.0 - (double)cos1 * (double)cos1); // sqrt(1-pow(cosTheta1,2)) in double (:169)
        if (sin1 > 1) sin1 = 1.0f;
        if (sin1 < -1) sin1 = -1.0f;
        if (cos1 > 1) cos1 = 1.0f;
        if (cos1 < -1) cos1 = -1.0f;
        float n1 = M.ior, n2 = 1.0f;
        if (H.front) { n1 = 1.0f; n2 = M.ior; }
        float sin2 = (n1 / n2) * sin1;
        float cos2 = sqrtf(1 - sin2 * sin2);
        if (cos2 > 1) cos2 = 1.0f;
        // SVector = N x ((N x -d).GetNormalized()), normalised (:203)
        float cx = sny * (-dirz) - snz * (-diry), cy = snz * (-dirx) - snx * (-dirz), cz = snx * (-diry) - sny * (-dirx);
        norm3(cx, cy, cz);
        float svx = sny * cz - snz * cy, svy = snz * cx - snx * cz, svz = snx * cy - sny * cx;
        norm3(svx, svy, svz);
        if (sin2 > 1) {
            // total internal reflection (:205-222); absorption uses an un-traced HitInfo, z = BIGFLOAT (SURVEY A-11)
            float k = 2 * dot3(dirx, diry, dirz, snx, sny, snz);
            float rx = dirx - k * snx, ry = diry - k * sny, rz = dirz - k * snz;
            norm3(rx, ry, rz);
            Col ab = mk(expf((-RTU_BIG) * M.absorption[0]), expf((-RTU_BIG) * M.absorption[1]), expf((-RTU_BIG) * M.absorption[2]));
            Col w = Wt * ab;
            if (!((P.flags & 2u) && !nonblack(w)))
                push_ray(O, H.px, H.py, H.pz, rx, ry, rz, w, pixel, pack_meta(RK_TIR, bounce - 1, H.material, tree), -1, child_path(path, 1u));
        } else {
            // second glossiness sample shadows the first for the refracted / mirror directions (:225-239)
            float o2x, o2y, o2z;
            sample_ball(rng, M.refr_gloss, o2x, o2y, o2z);
            float tnx = ((H.px + H.nx) + o2x) - H.px, tny = ((H.py + H.ny) + o2y) - H.py, tnz = ((H.pz + H.nz) + o2z) - H.pz;
            norm3(tnx, tny, tnz);
            float rdx = (-tnx) * cos2 + svx * sin2, rdy = (-tny) * cos2 + svy * sin2, rdz = (-tnz) * cos2 + svz * sin2;
            norm3(rdx, rdy, rdz);
            // Schlick (:236-237): pow(float,int) and the 1.0 literals promote to double
            double r0d = (double)((n1 - n2) / (n1 + n2));
            float R0 = (float)(r0d * r0d);
            float F = (float)((double)R0 + (1.0 - (double)R0) * pow(1.0 - (double)cos1, 5.0));
            float k = 2 * dot3(dirx, diry, dirz, tnx, tny, tnz);
            float mx = dirx - k * tnx, my = diry - k * tny, mz = dirz - k * tnz;
            norm3(mx, my, mz);
            unsigned a = warp_alloc(O.aux.count, true);
            if (a >= O.aux.cap) { O.counters->overflow = 1; }
            else {
                O.aux.a[a] = make_float4(Kt.r, Kt.g, Kt.b, F);
                O.aux.b[a] = make_float4(mx, my, mz, 0.f);
                push_ray(O, H.px, H.py, H.pz, rdx, rdy, rdz, Wt, pixel, pack_meta(RK_REFRACT, bounce - 1, H.material, tree), (int)a, child_path(path, 2u));
            }
        }
    }
    Col Kr = texcolor_sample(S, M.reflection, H.u, H.v, H.w);
    if (nonblack(Kr)) {
        float ox, oy, oz;
        sample_ball(rng, M.refl_gloss, ox, oy, oz);
        float snx = ((H.px + H.nx) + ox) - H.px, sny = ((H.py + H.ny) + oy) - H.py, snz = ((H.pz + H.nz) + oz) - H.pz;
        norm3(snx, sny, snz);
        float k = 2 * dot3(dirx, diry, dirz, snx, sny, snz);
        float rx = dirx - k * snx, ry = diry - k * sny, rz = dirz - k * snz;
        norm3(rx, ry, rz);
        int aux = -1;
        if (M.reflection.map >= 0) { // miss uses reflection.GetColor(), hit uses the sampled Kr (:286-289)
            unsigned a = warp_alloc(O.aux.count, true);
            if (a >= O.aux.cap) { O.counters->overflow = 1; return; }
            Col wm = Wt * mk(M.reflection.c[0], M.reflection.c[1], M.reflection.c[2]);
            O.aux.a[a] = make_float4(wm.r, wm.g, wm.b, 0.f);
            O.aux.b[a] = make_float4(0, 0, 0, 0);
            aux = (int)a;
        }
        push_ray(O, H.px, H.py, H.pz, rx, ry, rz, Wt * Kr, pixel, pack_meta(RK_REFLECT, bounce - 1, H.material, tree), aux, child_path(path, 3u));
    }
}
