// The opaque handles of include/rtu.h (one context per GPU / host thread, scenes owned by a context), shared by the
// translation units that implement the C ABI (rtu_api.cu: single-GPU path, rtu_multi.cu: collectives, rtu_async.cpp-style
// job control lives in rtu_api.cu as well).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../host/host_scene.h"
#include "rtu_internal.h"

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            rtu::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                         \
            return e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? RTU_ERR_NO_DEVICE : RTU_ERR_CUDA; \
        }                                                                                               \
    } while (0)

// Page-locked staging for small per-frame uploads: the copy is asynchronous and the host never waits for it; the buffer is
// only waited for when the NEXT upload wants to overwrite it (by then the copy is long done).
struct PinnedStage {
    void *host = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    bool inflight = false;
    void *acquire(size_t bytes)
    {
        if (inflight) { cudaEventSynchronize(done); inflight = false; }
        if (bytes > cap) {
            if (host) cudaFreeHost(host);
            host = nullptr;
            cap = 0;
            size_t want = bytes < 4096 ? 4096 : bytes * 2;
            if (cudaHostAlloc(&host, want, cudaHostAllocDefault) != cudaSuccess) { host = nullptr; return nullptr; }
            cap = want;
        }
        if (!done && cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        return host;
    }
    void submitted(cudaStream_t st) { cudaEventRecord(done, st); inflight = true; }
    void release()
    {
        if (inflight) cudaEventSynchronize(done);
        if (host) cudaFreeHost(host);
        if (done) cudaEventDestroy(done);
        host = nullptr; done = nullptr; cap = 0; inflight = false;
    }
};

struct rtu_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    LaunchCfg cfg;
    size_t chunk_rays = 1u << 27;  // primary rays per wave chunk (~300 B of queue space each): a 64-spp 1080p frame is one chunk;
                                   // 2^28 where the device has the memory (rtu_context_create), halved when an allocation fails
    bool chunk_rays_set = false;   // RTU_CHUNK_RAYS given: used as is
    bool scratch_oom = false;      // the last ensure_scratch failed for lack of memory
    double queue_factor = 1.0;
    // scratch (lazily sized)
    WaveBuffers wb;
    size_t q_cap = 0, shadow_cap = 0;
    std::vector<void *> scratch;
    unsigned *work = nullptr;
    size_t work_n = 0;
    unsigned *zmm = nullptr;
    float4 *gi = nullptr;  // GI records of the current chunk (RTU_MODE_PATH): one per primary hit
    size_t gi_n = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // optional per-launch timing (RTU_FLAG_TIME_KERNELS)
    std::vector<cudaEvent_t> kt_ev;
    std::vector<int> kt_cls;
    size_t kt_used = 0;
    bool kt_on = false;
    uint64_t cls_launches[4] = {0, 0, 0, 0}; // primary extend, queue extend, shadow, shade
    // frame buffers: owned by the context so that re-uploading a scene every frame (the e2e
    // path) does not re-allocate them
    struct FrameBuffers {
        float4 *accum = nullptr;
        size_t accum_n = 0;
        float4 *accum2 = nullptr;   // frame that is ADDED to a caller's accumulator once it is known to be complete
        size_t accum2_n = 0;
        float *d_rgb = nullptr;
        unsigned char *d_rgb8 = nullptr;
        float *d_z = nullptr;
        unsigned char *d_z8 = nullptr;
        int *d_node = nullptr, *d_face = nullptr;
        size_t img_n = 0;
        float2 *d_offsets = nullptr;
        size_t offsets_n = 0;
        unsigned char *d_tile = nullptr; // FrameSetup::tile_empty
        size_t tile_n = 0;
    } fb;
    // frame setup that is still valid on the device: a frame with the same key skips the uploads and k_tile_mask
    std::vector<float2> h_offsets;   // sub-pixel offsets of the pattern in d_offsets
    int off_spp = -1, off_pattern = -1;
    struct TileKey {
        uint64_t scene = 0;
        int W = 0, H = 0, row0 = 0, row1 = 0, s0 = 0, s1 = 0, spp = 0, pattern = 0;
        bool operator==(const TileKey &o) const { return scene == o.scene && W == o.W && H == o.H && row0 == o.row0 && row1 == o.row1 && s0 == o.s0 && s1 == o.s1 && spp == o.spp && pattern == o.pattern; }
    } tile_key;
    const unsigned char *tile_mask = nullptr; // device, valid for tile_key
    const unsigned *tile_count = nullptr;
    unsigned tile_total = 0;
    PinnedStage stage_off, stage_tile;
    PinnedStage stage_photon_raw, stage_photon_bal; // the host balancing's input and output (24 MB each for 10^6 photons)
    // adaptive sampling state of the frame in flight (rtu_params::adaptive_min_spp > 0): consulted by setup_frame
    unsigned char *d_tile_done = nullptr;
    int *d_tile_samples = nullptr;
    unsigned *d_n_active = nullptr;
    size_t adaptive_tiles = 0;
    bool adaptive_on = false;        // the frame being set up splits even / odd samples and skips finished tiles
    uint64_t pixel_samples = 0;      // camera samples of the last frame (rtu_stats::pixel_samples)
    rtu_stats adaptive_totals;       // counters summed over the passes of the last adaptive frame
    bool adaptive_totals_valid = false;
    unsigned char *d_scount = nullptr; // sample-count image
    size_t scount_n = 0;
    uint32_t *h_flag = nullptr;      // page-locked word the overflow flag is copied into
    uint64_t queue_retries = 0;      // frames re-rendered after a queue overflow (rtu_stats::queue_retries)
    // Rays that entered each secondary wave of the last frame's first chunk (WaveLog).  The next frame of the same shape runs
    // its small deep waves as one cooperative launch (k_tail_waves); a wrong guess costs time, never the result.
    enum { WAVE_LOG_MAX = 64 };
    unsigned *d_wave_log = nullptr;  // [0, 64): the first chunk's waves; [64, 128): scratch for the other chunks
    uint32_t *h_wave_log = nullptr;  // page-locked copy of the first half, readable once wave_ev has passed
    cudaEvent_t wave_ev = nullptr;
    struct WaveKey {
        int W = 0, rows = 0, mode = -1, shade_bounces = 0, gi_bounces = 0, n_waves = 0, n_nodes = 0;
        size_t chunk_samples = 0;
        bool operator==(const WaveKey &o) const
        {
            return W == o.W && rows == o.rows && mode == o.mode && shade_bounces == o.shade_bounces && gi_bounces == o.gi_bounces &&
                   n_waves == o.n_waves && n_nodes == o.n_nodes && chunk_samples == o.chunk_samples;
        }
    } wave_key;
    bool wave_log_valid = false;
    size_t tail_rays = 16384;        // waves at most this large go into the tail launch (RTU_TAIL_RAYS; 0: never)
    uint64_t tail_launches = 0;
};

struct rtu_scene {
    rtu_context *ctx = nullptr;
    DScene S;
    rtu_camera cam;
    std::vector<void *> owned;
    uint64_t serial = 0;          // unique per upload (frame-setup cache key)
    double bvh_build_ms = 0;      // device time of the LBVH builds of this upload (RTU_MESH_DEVICE_BVH meshes)
    int tree_waves = 2;           // 0: no material reflects or refracts, 1: mirrors only, 2: refraction present (wave_count)
    size_t device_bytes = 0;
    int n_shadow_lights = 0;
    uint64_t launches = 0;
    bool timed = false;
    // photon map (balanced, n+1 records, record 0 unused) and the parameters it was made with
    struct Footprint { double c[8][3]; bool finite; int node; }; // world-space corners of an object's bound box
    std::vector<Footprint> footprints;
    size_t chunk_limit = 0;       // set after a queue overflow: later frames of this scene start with smaller chunks ...
    double queue_boost = 1.0;     // ... or with more queue entries per primary ray
    size_t last_chunk_cap = 0;    // primary rays per chunk of the last frame
    bool root_identity = true;
    int n_obj = 0;                // nodes whose object Trace() tests (what a ray that misses everything books)
    int h_light0_kind = -1;       // lights[0]: the only light GeneratePhotonMap emits from
    float h_light0_I[3] = {0, 0, 0};
    rtu_photon *d_photons = nullptr;
    float4 *d_knn = nullptr;      // walk records (3 per node) + direction table + power table of the estimate, one allocation
    uint32_t n_photons = 0;
    bool photon_device_build = false; // the map in d_photons was balanced by photon_build.cu (no tie met)
    rtu_photon_params photon_params;
};


// shared between the translation units of the C ABI (defined in rtu_api.cu)
int rtu_frame_dims(const rtu_scene *s, const rtu_params *p, int *W, int *H);
int rtu_ensure_image(rtu_scene *s, size_t npix);
// accum -> device images -> host buffers, enqueued on the context's stream (no wait)
int rtu_resolve_enqueue(rtu_scene *s, const rtu_params *p, const float4 *accum, rtu_image *out);
// frame into d_accum (NULL: the context's accumulator) with the overflow check / retry; out != NULL: resolve + copies behind it
int rtu_render_checked(rtu_scene *s, const rtu_params *p, float *d_accum, int32_t clear_accum, rtu_image *out);
// LBVH over the triangles of a mesh, built on the device (lbvh_build.cu); outputs are pushed onto `owned`
int rtu_lbvh_build(cudaStream_t st, const float *d_v, const unsigned *d_f, unsigned nf, const float bmin[3], const float bmax[3],
                   const TriRec *by_face, std::vector<void *> &owned, const OccNode **nodes_out, const TriRec **tris_out, uint32_t *root_out,
                   float *build_ms);
