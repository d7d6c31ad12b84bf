// Device build of cyPhotonMap's left-balanced kd-tree (PhotonMap::PrepareForIrradianceEstimation / BalanceSegment,
// cyPhotonMap.h:207-296): heap order, slot 0 unused, splitting axis in the low two bits of plane_dirz.
//
// The reference recursion works on ranges [start, end] of one array: it moves the element of rank `median` along the
// widest axis of the range's box to position `median` (a Hoare selection), stores it at heap slot `index`, and recurses
// into [start, median-1] (slot 2 index) and [median+1, end] (slot 2 index + 1) with the box cut at the split value.
// WHICH photon has rank `median`, and which photons lie on either side of it, is a property of the set as long as the
// median's coordinate differs from its two neighbours' in sorted order - the order inside a range never matters.  So all
// ranges of one recursion level are processed together:
//
//   k_bal_keys     key(position i) = (start of i's range) << 32 | order-preserving bits of its coordinate on the range's axis
//   cub radix sort (a library sort: plumbing of a build step, like the LBVH builder's) - every range is sorted in place
//   k_bal_split    one thread per range: the median goes to its heap slot with the plane bits, one-element sides go to
//                  their slots as they are (BalanceSegment :276, :287), longer sides become the next level's ranges
//   k_bal_assign   every position learns its next range (the left side keeps the start, the right side starts at median+1)
//
// for floor(log2 n) + 1 levels.  If a median ties with a neighbour (or a coordinate is NaN) the result would depend on the
// selection's swap sequence: the flag is raised and the caller balances on the host (host/photon_host.cpp), byte-identical
// to the reference in every case.  Without ties the two builds agree byte for byte (tests: the reference's own 20 000-photon
// map and a 10^6-photon emission).
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <initializer_list>

#include "rtu_internal.h"

namespace {

struct BalMeta { // of the range that starts at this position
    unsigned end, index;
    float lo[3], hi[3];
};

__device__ __forceinline__ unsigned median_of(unsigned start, unsigned end) // cyPhotonMap.h:233-241
{
    const unsigned count = end - start + 1u;
    unsigned m = 1;
    while (4u * m <= count) m += m;
    if (3u * m <= count) return 2u * m + start - 1u;
    return end - m + 1u;
}

__device__ __forceinline__ int axis_of(const BalMeta &M) // widest extent; ties fall through exactly like :244-248
{
    const float ex = M.hi[0] - M.lo[0], ey = M.hi[1] - M.lo[1], ez = M.hi[2] - M.lo[2];
    int axis = 2;
    if (ex > ey) {
        if (ex > ez) axis = 0;
    } else if (ey > ez) {
        axis = 1;
    }
    return axis;
}

__device__ __forceinline__ unsigned ordered_bits(float v)
{
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// the reference seeds the box with the zeroed slot 0 of its vector (:212-213): the origin is always inside
__global__ void k_bal_bounds(const rtu_photon *raw, unsigned n, unsigned *mm)
{
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int k = 0; k < 3; k++) {
            const unsigned b = ordered_bits(raw[i].position[k]);
            lo[k] = min(lo[k], b);
            hi[k] = max(hi[k], b);
        }
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&mm[k], lo[k]);
            atomicMax(&mm[3 + k], hi[k]);
        }
    }
}

__global__ void k_bal_init(unsigned n, const unsigned *mm, unsigned *seg, unsigned *src, BalMeta *meta)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    seg[i] = i == 0 ? 0u : 1u;
    src[i] = i == 0 ? 0u : i - 1u;
    if (i == 1) {
        BalMeta M;
        M.end = n;
        M.index = 1;
        for (int k = 0; k < 3; k++) {
            const unsigned l = mm[k], h = mm[3 + k];
            M.lo[k] = __uint_as_float((l & 0x80000000u) ? (l & 0x7fffffffu) : ~l);
            M.hi[k] = __uint_as_float((h & 0x80000000u) ? (h & 0x7fffffffu) : ~h);
        }
        meta[1] = M;
    }
}

__global__ void k_bal_keys(const rtu_photon *raw, unsigned n, const unsigned *seg, const unsigned *src, const BalMeta *meta,
                           unsigned long long *keys)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const unsigned s = seg[i];
    if (s == 0u) { keys[i] = (unsigned long long)i << 32; return; } // placed (or slot 0): stays where it is
    const int axis = axis_of(meta[s]);
    keys[i] = ((unsigned long long)s << 32) | ordered_bits(raw[src[i]].position[axis]);
}

__device__ __forceinline__ void copy_photon(rtu_photon *dst, const rtu_photon *srcp, int axis_bits)
{
    const uint2 *q = reinterpret_cast<const uint2 *>(srcp);
    uint2 a = q[0], b = q[1], c = q[2];
    if (axis_bits >= 0) c.x = (c.x & ~(0xf7u << 24)) | ((unsigned)axis_bits << 24); // plane_dirz = (plane_dirz & 0x8) | axis
    uint2 *o = reinterpret_cast<uint2 *>(dst);
    o[0] = a; o[1] = b; o[2] = c;
}

__global__ void k_bal_split(const rtu_photon *raw, unsigned n, const unsigned *seg, const unsigned *src, const BalMeta *meta,
                            BalMeta *next, rtu_photon *out, unsigned *tie)
{
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0u || s > n || seg[s] != s) return;
    const BalMeta M = meta[s];
    const unsigned e = M.end, m = median_of(s, e);
    const int axis = axis_of(M);
    const float split = raw[src[m]].position[axis];
    bool bad = !(split == split);
    if (m > s) { const float v = raw[src[m - 1]].position[axis]; if (!(v < split)) bad = true; }
    if (m < e) { const float v = raw[src[m + 1]].position[axis]; if (!(v > split)) bad = true; }
    if (bad) *tie = 1u;
    copy_photon(out + M.index, raw + src[m], axis);
    if (m > s) {
        if (m - 1u == s) copy_photon(out + 2u * M.index, raw + src[s], -1);
        else {
            BalMeta L = M;
            L.end = m - 1u;
            L.index = 2u * M.index;
            L.hi[axis] = split;
            next[s] = L;
        }
    }
    if (m < e) {
        if (m + 1u == e) copy_photon(out + 2u * M.index + 1u, raw + src[e], -1);
        else {
            BalMeta R = M;
            R.index = 2u * M.index + 1u;
            R.lo[axis] = split;
            next[m + 1u] = R;
        }
    }
}

__global__ void k_bal_assign(unsigned n, unsigned *seg, const BalMeta *meta)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0u || i > n) return;
    const unsigned s = seg[i];
    if (s == 0u) return;
    const unsigned e = meta[s].end, m = median_of(s, e);
    unsigned ns;
    if (i == m) ns = 0u;
    else if (i < m) ns = (m - 1u == s) ? 0u : s;
    else ns = (m + 1u == e) ? 0u : m + 1u;
    seg[i] = ns;
}

} // namespace

// raw: n photons (device, any order); out: n + 1 records (device), out[1..n] the balanced heap; *tie_host = 1 when the
// build met a tie and `out` must not be used.  Synchronises the stream (the flag is needed on the host).
cudaError_t launch_photon_balance(cudaStream_t st, const rtu_photon *raw, unsigned n, rtu_photon *out, unsigned *tie_host)
{
    *tie_host = 0;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(rtu_photon), st);
    if (e != cudaSuccess || n == 0) return e;
    const size_t N = (size_t)n + 1;
    unsigned long long *keys[2] = {nullptr, nullptr};
    unsigned *src[2] = {nullptr, nullptr}, *seg = nullptr, *flags = nullptr; // flags: 6 bounds words + the tie flag
    BalMeta *meta[2] = {nullptr, nullptr};
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int end_bit = 33;
    while (end_bit < 64 && (N >> (end_bit - 32)) != 0) end_bit++;
    e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys[0], keys[1], src[0], src[1], (int)N, 0, end_bit, st);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaMallocAsync((void **)&keys[k], N * sizeof(unsigned long long), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&src[k], N * sizeof(unsigned), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&meta[k], (N + 1) * sizeof(BalMeta), st);
    }
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&seg, N * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&flags, 8 * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st);
    if (e == cudaSuccess) {
        // bounds start at the origin: ordered_bits(+0.0f) = 0x80000000 for both min and max
        const unsigned init[8] = {0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0u, 0u};
        e = cudaMemcpyAsync(flags, init, sizeof init, cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess) {
        const unsigned G = (unsigned)((N + 255) / 256);
        k_bal_bounds<<<std::min(G, 1184u), 256, 0, st>>>(raw, n, flags);
        k_bal_init<<<G, 256, 0, st>>>(n, flags, seg, src[0], meta[0]);
        int levels = 0;
        while ((1ull << levels) <= (unsigned long long)n) levels++; // floor(log2 n) + 1
        int cur = 0;
        for (int lv = 0; lv < levels && e == cudaSuccess; lv++) {
            k_bal_keys<<<G, 256, 0, st>>>(raw, n, seg, src[cur], meta[lv & 1], keys[0]);
            e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys[0], keys[1], src[cur], src[cur ^ 1], (int)N, 0, end_bit, st);
            cur ^= 1;
            k_bal_split<<<G, 256, 0, st>>>(raw, n, seg, src[cur], meta[lv & 1], meta[(lv & 1) ^ 1], out, flags + 6);
            k_bal_assign<<<G, 256, 0, st>>>(n, seg, meta[lv & 1]);
            // maps that tie mostly do so at once (photons on axis-aligned walls): look at the flag after the first levels
            // instead of building twenty levels of a tree nobody will use
            if ((lv == 1 || lv == 5) && lv + 1 < levels && e == cudaSuccess) {
                e = cudaMemcpyAsync(tie_host, flags + 6, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e == cudaSuccess && *tie_host) break;
            }
        }
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(tie_host, flags + 6, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
    }
    for (void *q : {(void *)keys[0], (void *)keys[1], (void *)src[0], (void *)src[1], (void *)meta[0], (void *)meta[1], (void *)seg, (void *)flags, tmp})
        if (q) cudaFreeAsync(q, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e;
}
