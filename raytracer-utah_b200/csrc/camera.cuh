// Camera rays and the work-item -> (sample, pixel) mapping of the primary wave, shared by the wave kernels
// (rtu_kernels.cu) and the photon-mapping shade kernel (photon_kernels.cu).
#pragma once
#include "rtu_internal.h"
#include "shade.cuh"

// The lens point of a depth-of-field camera (RenderFunctions.cpp:272-283).  A real function: the counter-based generator and
// sinf / cosf are ~250 instructions, the camera ray is built at two places of the pooled primary kernel, and most scenes
// have no depth of field.
static __device__ __noinline__ void lens_sample(uint2 key, unsigned pixel, unsigned path, unsigned dim, float dof, float *lx, float *ly)
{
    Rng rng;
    rng.key = key; rng.pixel = pixel; rng.path = path; rng.dim = dim;
    float4 u = rng.next4();
    float th = u.y * 6.283185307179586f;
    float rad = sqrtf(u.x * dof * dof);
    *lx = rad * cosf(th);
    *ly = rad * sinf(th);
}

// Camera ray of pixel (x,y) with sub-pixel offset (ox,oy): RenderFunctions.cpp:88-97, 258-269
__device__ __forceinline__ Ray camera_ray(const DCamera &C, int x, int y, float ox, float oy, Rng *rng)
{
    float fi = (float)x + ox, fj = (float)y + oy;
    float cx = (C.origin[0] + fi * C.u[0]) + fj * C.v[0];
    float cy = (C.origin[1] + fi * C.u[1]) + fj * C.v[1];
    float cz = (C.origin[2] + fi * C.u[2]) + fj * C.v[2];
    Ray r;
    r.px = C.pos[0]; r.py = C.pos[1]; r.pz = C.pos[2];
    if (C.dof > 0.f && rng) {
        float lx, ly;
        lens_sample(rng->key, rng->pixel, rng->path, rng->dim, C.dof, &lx, &ly);
        r.px = (C.pos[0] + C.lens_y[0] * ly) + C.lens_x[0] * lx;
        r.py = (C.pos[1] + C.lens_y[1] * ly) + C.lens_x[1] * lx;
        r.pz = (C.pos[2] + C.lens_y[2] * ly) + C.lens_x[2] * lx;
    }
    r.dx = cx - r.px; r.dy = cy - r.py; r.dz = cz - r.pz;
    norm3(r.dx, r.dy, r.dz);
    return r;
}

// Work item -> (sample, pixel) of the primary wave: samples outermost, then 8x4 pixel tiles so
// that the 32 lanes of a warp trace a compact bundle of camera rays.
struct PrimaryMap {
    int W, row_begin, row_end, tilesX;
    unsigned perSample;
    unsigned m_ps, m_tx; // floor(2^32 / divisor) - 1: quotient estimates for the two divisions of decode()
    __device__ __forceinline__ void init(const FrameSetup &F)
    {
        W = F.cam.width;
        row_begin = F.row_begin;
        row_end = F.row_end;
        tilesX = (W + 7) >> 3;
        int tilesY = (row_end - row_begin + 3) >> 2;
        perSample = (unsigned)(tilesX * tilesY) * 32u;
        // 2^32 / d in double is off by less than 1/d from the true quotient, so its floor is exact; one less keeps the
        // estimate q' = umulhi(n, m) at or below the true quotient for every n
        m_ps = (unsigned)(4294967296.0 / (double)perSample) - 1u;
        m_tx = tilesX > 1 ? (unsigned)(4294967296.0 / (double)tilesX) - 1u : 0xfffffffeu;
    }
    // n / d and n % d from the precomputed estimate (a 32-bit division costs ~20 instructions, this one 5)
    static __device__ __forceinline__ unsigned divmod(unsigned n, unsigned d, unsigned m, unsigned &rem)
    {
        unsigned q = __umulhi(n, m);
        unsigned r = n - q * d;
        while (r >= d) { q++; r -= d; }
        rem = r;
        return q;
    }
    // the tile (index into FrameSetup::tile_empty) of a work item
    __device__ __forceinline__ unsigned tile_of(unsigned idx) const
    {
        unsigned t;
        divmod(idx, perSample, m_ps, t);
        return t >> 5;
    }
    __device__ __forceinline__ bool decode(unsigned idx, int s0, int &s, int &x, int &y) const
    {
        unsigned t;
        unsigned sl = divmod(idx, perSample, m_ps, t);
        s = s0 + (int)sl;
        unsigned tile = t >> 5, in5 = t & 31u, utx;
        unsigned uty = divmod(tile, (unsigned)tilesX, m_tx, utx);
        x = (int)utx * 8 + (int)(in5 & 7u);
        y = row_begin + (int)uty * 4 + (int)(in5 >> 3);
        return x < W && y < row_end;
    }
};

// RNG path word of a camera sample: a function of (pixel, sample) only, so random streams do not
// depend on chunking, queue order or the number of GPUs
__device__ __forceinline__ unsigned primary_path(int pixel, int s)
{
    unsigned h = (unsigned)pixel * 0x9E3779B9u ^ ((unsigned)s + 0x7F4A7C15u) * 0x85EBCA6Bu;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}

__device__ __forceinline__ Ray primary_ray(const FrameSetup &F, int s, int x, int y, int pixel)
{
    float2 off = __ldg(&F.sample_offsets[s]);
    Rng rng;
    rng.key = F.seed; rng.pixel = (unsigned)pixel; rng.path = (unsigned)s; rng.dim = 1000u;
    return camera_ray(F.cam, x, y, off.x, off.y, &rng);
}

