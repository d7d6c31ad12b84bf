// Launch wrappers implemented in rtu_kernels.cu and used by rtu_api.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/rtu.h"
#include "device_scene.h"

struct FrameSetup {
    DCamera cam;
    int spp;             // total samples per pixel of the frame (defines the reference pattern)
    const float2 *sample_offsets; // device, spp entries: pixel offsets of sample s
    int row_begin, row_end;
    int mode;
    int shade_bounces;
    int gi_bounces;
    unsigned flags;
    uint2 seed;
    // primary wave: tile_empty[tile] != 0 when no camera ray of that 8x4-pixel tile (any rendered sample) can reach the
    // bounding sphere of any object; such rays book n_obj culled nodes each and add the background (may be NULL)
    const unsigned char *tile_empty;
    int n_obj;
    // number of empty tiles (device word written by k_tile_mask) and of tiles: a warp of the primary wave takes 256 work
    // items per atomic when most tiles are empty, else 32.  Read on the device so that the host never waits for the mask.
    const unsigned *n_empty_tiles;
    unsigned n_tiles;
    // adaptive sampling (rtu_params::adaptive_min_spp > 0): tiles whose pixels have all reached the target variance are no
    // longer sampled (tile_done[tile] != 0), and the accumulator has two halves of npix entries - even samples go to the
    // first, odd samples to the second - whose difference is the variance estimate
    const unsigned char *tile_done;
    unsigned half_split; // 0, or npix: the offset of the odd samples' half
};

__host__ __device__ inline int half_slot(const FrameSetup &F, int pixel, int s) { return pixel + ((s & 1) ? (int)F.half_split : 0); }

struct LaunchCfg {
    int sm_count;
    int blocks_per_sm;
    int threads;
};

// rays a warp of the pooled closest-hit kernel can have parked (jobs) and resuming (res); 3 float4 per entry
#define XP_JOBS 64
#define XP_RES 96
#define XP_MAX_WARPS (148 * 4 * 8 * 2)

struct WaveBuffers {
    RayQueue q[2];
    AuxPool aux[2];
    ShadowQueue shadow;
    HitQueue hits;
    float4 *park;         // per-warp lists of rays parked at a mesh / resuming behind it (pooled closest-hit kernel)
    unsigned *gi_count;   // number of GI records (= primary hits) of the current chunk (RTU_MODE_PATH)
    unsigned *work;       // device work-fetch counters (one per launch slot)
    DCounters *counters;
};

// closest hit of generated primary rays: samples [s0,s1) of rows [row_begin,row_end) -> hit queue
void launch_extend_primary(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, int s0, int s1,
                           const WaveBuffers &B, float4 *pixel_accum, float4 *accum, unsigned *work_counter);
// shade the hit queue of a primary wave -> shadow queue + q[out_q]
void launch_shade_primary(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, int s0,
                          const WaveBuffers &B, int out_q, float4 *accum, unsigned *work_counter);
// closest hit of queue q[in_q] -> hit queue;  shade -> shadow queue + q[1-in_q]
void launch_extend_queue(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, const WaveBuffers &B,
                         int in_q, float4 *accum, unsigned *work_counter);
void launch_shade_queue(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, const WaveBuffers &B,
                        int in_q, float4 *accum, unsigned *work_counter);
// any-hit over the shadow queue, adds unoccluded contributions to accum
// reference_walk (RTU_FLAG_REFERENCE_WALK): mesh walks go through the cyBVH with the reference's own box tests instead of the
// meshes' any-hit hierarchies
void launch_shadow_wave(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const WaveBuffers &B, float4 *accum,
                        unsigned *work_counter, bool reference_walk = false);
void launch_gi_combine(cudaStream_t st, const float4 *gi, const unsigned *count, unsigned cap, int gi_bounces, float4 *accum);
void launch_reset_counts(cudaStream_t st, unsigned *a, unsigned *b, unsigned *c, unsigned *d, unsigned *log_dst = nullptr,
                         const unsigned *log_src = nullptr);
// waves [w0, n_waves) of a chunk as ONE cooperative launch (k_tail_waves); work: three zeroed counters per wave
bool launch_tail_waves(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, const WaveBuffers &B, int in_q, int w0,
                       int n_waves, float4 *target, unsigned *work, unsigned *wave_log);

// pixel-centre primary visibility: z / node / face per pixel (RTU_MODE_PRIMARY, ZBuffer.png)
void launch_primary_ids(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const DCamera &cam, float *z, int *node,
                        int *face, DCounters *counters);
// batched operators
void launch_trace_batch(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const rtu_ray *rays, long long n,
                        rtu_hit *hits, DCounters *counters);
void launch_shadow_batch(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const rtu_ray *rays, const float *tmax,
                         long long n, unsigned char *occ, DCounters *counters);
// Trace() of caller rays through the frame's own closest-hit kernel (launch_extend_queue); scratch_accum: one float4 (never written)
void launch_trace_batch_wave(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const rtu_ray *rays, long long n, rtu_hit *hits,
                             const WaveBuffers &B, float4 *scratch_accum, unsigned *work_counter, bool reference_walk);
// the same operator through the frame's own any-hit kernel (launch_shadow_wave): accum = n zeroed float4
void launch_shadow_batch_wave(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const rtu_ray *rays, const float *tmax,
                              long long n, unsigned char *occ, const WaveBuffers &B, float4 *accum, unsigned *work_counter);
// Shade(ray, hit, lights, bounces) for caller-provided hits: first step, then the usual waves
void launch_shade_batch(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, const FrameSetup &F, const rtu_ray *rays,
                        const rtu_hit *hits, long long n, const WaveBuffers &B, int out_q, float4 *accum);
void launch_camera_rays(cudaStream_t st, const DCamera &cam, float ox, float oy, rtu_ray *rays);
void launch_selftest_div(cudaStream_t st, unsigned per_thread, unsigned long long seed, unsigned long long *mismatch,
                         unsigned long long *tested);
// accum -> mean, gamma 1/2.2, Color24 (RenderFunctions.cpp:152-159)
void launch_resolve(cudaStream_t st, const float4 *accum, int npix, float inv_unused, int spp, float *rgb, unsigned char *rgb8);
void launch_accum_add(cudaStream_t st, float4 *dst, const float4 *src, size_t npix);
// multi-GPU: accumulator -> three planes of RGB sums (what the collective moves); resolve of reduced planes
void launch_pack_rgb(cudaStream_t st, const float4 *accum, size_t npix, float *planes);
void launch_resolve_planes(cudaStream_t st, const float *planes, size_t npix, int spp, float *rgb, unsigned char *rgb8);
// adaptive sampling: after a pass that brought the active tiles to n_now samples, per 8x4 tile the largest estimated variance of
// a pixel mean, ((A/nA - B/nB) / 2)^2 per channel; tiles at or below `target` (or at max_spp) are done.  *n_active counts the rest.
void launch_adaptive_update(cudaStream_t st, const float4 *accum, int W, int H, int n_now, int max_spp, float target, unsigned char *tile_done,
                            int *tile_samples, unsigned *n_active);
// mean over each pixel's OWN sample count, gamma, Color24; sample_count: RenderImage::sampleCount (scene.h:545), saturating at 255
void launch_resolve_adaptive(cudaStream_t st, const float4 *accum, int W, int H, const int *tile_samples, float *rgb, unsigned char *rgb8,
                             unsigned char *sample_count);
// RenderImage::ComputeZBufferImage (scene.h:590-612)
void launch_zimage(cudaStream_t st, const float *z, int npix, unsigned *minmax_bits, unsigned char *z8);

// photon path (photon_kernels.cu)
// device kd-tree build (photon_build.cu); *tie_host = 1: a median tied with a neighbour, `out` is not the reference's tree
cudaError_t launch_photon_balance(cudaStream_t st, const rtu_photon *raw, unsigned n, rtu_photon *out, unsigned *tie_host);
cudaError_t launch_knn_build(cudaStream_t st, const rtu_photon *map, int n, int half, float4 *nodes, float4 *dir, float4 *pw);
// (these three allocate their scratch from the stream-ordered pool and return the first CUDA error)
cudaError_t launch_estimate(cudaStream_t st, const DPhotonMap &PM, const float *pos, const float *normal, long long n, float radius,
                            float norm_scale, float *irrad, float *direction, int *found);
cudaError_t launch_photon_shade(cudaStream_t st, const DScene &S, const FrameSetup &F, int s0, const WaveBuffers &B, unsigned max_hits,
                                const DPhotonMap &PM, float4 *accum);
cudaError_t launch_photon_gather(cudaStream_t st, const DScene &S, const FrameSetup &F, int s0, const WaveBuffers &B, unsigned max_hits,
                                 const DPhotonMap &PM, float4 *accum);
void launch_photon_emit(const LaunchCfg &cfg, cudaStream_t st, const DScene &S, unsigned long long path0, unsigned n_paths,
                        int max_bounce, uint2 seed, int light, rtu_photon *staging, unsigned char *counts, DCounters *counters,
                        unsigned *work_counter);
void launch_photon_compact(cudaStream_t st, const rtu_photon *staging, const unsigned char *counts, const unsigned *offsets,
                           unsigned n_paths, unsigned cut, int max_bounce, rtu_photon *map1, unsigned cap);
void launch_photon_scale(cudaStream_t st, rtu_photon *map1, unsigned n, float scale);

// image-space footprint of one object: pixel-space bounding box and n_edges outward half-planes (nx, ny, c):
// a point p is outside when nx p.x + ny p.y > c
struct TileObject {
    float lo[2], hi[2];
    int first_edge, n_edges;
};
// tiles of the primary wave that no object's footprint touches (see FrameSetup::tile_empty)
void launch_tile_mask(cudaStream_t st, const FrameSetup &F, const TileObject *objs, int n_objs, const float4 *edges, float ox0,
                      float ox1, float oy0, float oy1, unsigned char *mask, unsigned *n_empty);
