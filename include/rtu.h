/*
 * rtu.h -- C ABI of the B200-native render path for RayTracer-Utah scenes.
 *
 * This is the drop-in boundary for ONE path of the reference: the per-pixel render loop
 *   Render -> Trace/ShadowTrace -> Object::IntersectRay -> Material::Shade -> Light::Illuminate
 * (RenderFunctions.cpp:55-240, objFunctions.cpp:15-522, mtlFunctions.cpp:120-298,
 *  lightFunctions.cpp:27-84).  Everything is extern "C", plain pointers and sizes; no C++,
 * CUDA or torch types appear in a signature.  All functions return 0 on success and a
 * non-zero rtu_status otherwise; rtu_last_error() gives the message (the reference's render
 * path has no error reporting at all: void functions on detached threads, main.cpp:29-64).
 *
 * The host-side scene description below is a flattened, read-only mirror of the reference's
 * scene globals (main.cpp:17-27): what LoadScene() (xmlload.cpp:64) leaves in rootNode,
 * camera, materials, lights, objList, background, environment, textureList.  A reference
 * maintainer fills it by walking those globals (INTEGRATION.md shows the stub); our own
 * loader (rtu_host_load_xml) fills it from the same XML/OBJ/PNG files.
 *
 * Matrices are column-major float[9] exactly as cyMatrix3f::data (cyMatrix.h:290-294).
 */
#ifndef RTU_H_INCLUDED
#define RTU_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTU_BIGFLOAT 1.0e30f /* scene.h:55 */

typedef enum {
    RTU_OK = 0,
    RTU_ERR_INVALID = 1,   /* bad argument / malformed scene description */
    RTU_ERR_CUDA = 2,      /* CUDA runtime error (message has the cudaError string) */
    RTU_ERR_IO = 3,        /* file missing / unreadable / unparsable */
    RTU_ERR_NO_DEVICE = 4, /* no CUDA device: there is NO CPU fallback */
    RTU_ERR_UNSUPPORTED = 5,
    RTU_ERR_CANCELLED = 6  /* rtu_job_cancel (StopRender) ended the frame */
} rtu_status;

/* ------------------------------------------------------------------ scene description */

/* Object kinds: the closed world xmlload.cpp:193-199 can create. */
enum { RTU_OBJ_NONE = 0, RTU_OBJ_SPHERE = 1, RTU_OBJ_PLANE = 2, RTU_OBJ_MESH = 3 };

/* One scene-graph Node (scene.h:437-513).  Nodes are stored in PRE-ORDER, which is the order
 * Trace() visits them (RenderFunctions.cpp:181-213); node 0 is rootNode. */
typedef struct rtu_node {
    float tm[9];      /* Transformation::tm  (scene.h:226) */
    float itm[9];     /* Transformation::itm (scene.h:228) */
    float pos[3];     /* Transformation::pos (scene.h:227) */
    int32_t parent;   /* index of the parent node, -1 for the root */
    int32_t kind;     /* RTU_OBJ_* of Node::obj (scene.h:442) */
    int32_t mesh;     /* index into meshes[] when kind==RTU_OBJ_MESH, else -1 */
    int32_t material; /* index into materials[] of Node::mtl, -1 if none */
} rtu_node;

/* One TriObj (objects.h:46-66): cyTriMesh arrays (cyTriMesh.h:106-119) + its cyBVH
 * (cyBVH.h:187-203) exactly as built by bvh.SetMesh(this,4) (objects.h:58). */
typedef struct rtu_mesh {
    const float *v;     uint32_t nv;  /* 3 floats per vertex */
    const float *vn;    uint32_t nvn;
    const float *vt;    uint32_t nvt;
    const uint32_t *f;                /* nf x 3 vertex indices */
    const uint32_t *fn;               /* nf x 3 normal indices */
    const uint32_t *ft;               /* nf x 3 texture-vertex indices (may be NULL) */
    uint32_t nf;
    const float *bvh_boxes;           /* bvh_nodes x 6 (min xyz, max xyz); node 0 unused, root = 1 */
    const uint32_t *bvh_data;         /* bvh_nodes packed words: leaf bit 31, count-1 bits 28..30, offset / child index */
    uint32_t bvh_nodes;               /* number of entries INCLUDING the unused node 0 */
    const uint32_t *bvh_elements;     /* nf face ids in leaf order */
    float bound_min[3], bound_max[3]; /* cyTriMesh::boundMin/Max */
    /* Optional: the any-hit hierarchy of rtu_host_build_occlusion_bvh over the same triangles (what shadow rays walk on the
     * device; the cyBVH above still decides every answer).  All zero / NULL: rtu_scene_upload builds it itself. */
    const float *occ_nodes;           /* occ_n_nodes x 32 words: 4-wide nodes */
    const uint32_t *occ_slots;        /* nf indices into bvh_elements, in the hierarchy's leaf order */
    uint32_t occ_n_nodes;
    uint32_t occ_root;
    uint32_t flags;                   /* RTU_MESH_* */
} rtu_mesh;

/* rtu_mesh::flags.  RTU_MESH_DEVICE_BVH: the mesh carries no cyBVH (bvh_* and occ_* are ignored and may be NULL);
 * rtu_scene_upload builds an LBVH over its triangles on the device and every walk runs on that.  Results are the
 * reference's except where its own tree decides: of two triangles at exactly the same distance the lower face index wins
 * (the reference keeps the one its walk visits first), and RTU_FLAG_REFERENCE_WALK has nothing to walk. */
enum { RTU_MESH_DEVICE_BVH = 1 };

enum { RTU_TEX_NULL = 0, RTU_TEX_CHECKER = 1, RTU_TEX_FILE = 2 };

/* A TextureMap (scene.h:375-397) together with the Texture it points at
 * (texture.h:20-45).  RTU_TEX_NULL is a map whose texture failed to load: it samples black
 * (scene.h:382). */
typedef struct rtu_texmap {
    int32_t kind;
    float itm[9];       /* Transformation of the map: Sample() uses itm*(uvw-pos) (scene.h:235,382) */
    float pos[3];
    float color1[3];    /* TextureChecker */
    float color2[3];
    const uint8_t *rgb8; /* TextureFile::data, width*height*3, row 0 first */
    int32_t width, height;
} rtu_texmap;

/* TexturedColor (scene.h:405-433): colour, optionally multiplied by a texture map. */
typedef struct rtu_texcolor {
    float color[3];
    int32_t texmap; /* index into texmaps[], -1 when TexturedColor::map == NULL */
} rtu_texcolor;

/* MtlBlinn (materials.h:20-57).  A MultiMtl is represented by its sub-material 0, which is
 * the only one the reference can ever shade with (hInfo.mtlID is never set: SURVEY A-9). */
typedef struct rtu_material {
    rtu_texcolor diffuse, specular, reflection, refraction, emission;
    float glossiness;
    float absorption[3];
    float ior;
    float reflection_glossiness, refraction_glossiness;
} rtu_material;

enum { RTU_LIGHT_AMBIENT = 0, RTU_LIGHT_DIRECT = 1, RTU_LIGHT_POINT = 2 };

/* AmbientLight / DirectLight / PointLight (lights.h:28-99). */
typedef struct rtu_light {
    int32_t kind;
    float intensity[3];
    float v[3]; /* DirectLight::direction (normalised) or PointLight::position */
    float size; /* PointLight::size (soft shadow disk radius) */
} rtu_light;

/* Camera (scene.h:517-535) after LoadScene's fix-up (xmlload.cpp:109-126). */
typedef struct rtu_camera {
    float pos[3], dir[3], up[3];
    float fov, focaldist, dof;
    int32_t width, height;
} rtu_camera;

/* A light mask prebuilt by rtu_host_build_light_mask / rtu_host_load_xml for mesh node `node` and light `light` (-1: the camera).
 * It has to be rebuilt when the mesh, the transforms above the node or the light change. */
typedef struct rtu_light_mask {
    int32_t node, light;
    float rec[24];
    const uint32_t *bits;       /* 2048 words */
    /* Optional light lists (lights only; NULL / 0: none): per cell the triangles a shadow ray of that cell can meet, in the
     * order of their distance from the light.  Such a ray tests those instead of walking the mesh's hierarchy. */
    const uint32_t *cell_start; /* 256 * 256 + 1 offsets into items */
    const uint32_t *items;      /* 2 words per entry: the triangle (cyBVH slot; face index for RTU_MESH_DEVICE_BVH), its least depth (float) */
    uint32_t n_items;
} rtu_light_mask;

typedef struct rtu_scene_desc {
    rtu_camera camera;
    const rtu_node *nodes;         int32_t n_nodes;
    const rtu_mesh *meshes;        int32_t n_meshes;
    const rtu_material *materials; int32_t n_materials;
    const rtu_light *lights;       int32_t n_lights;
    const rtu_texmap *texmaps;     int32_t n_texmaps;
    rtu_texcolor background;       /* scene.h global `background`  */
    rtu_texcolor environment;      /* scene.h global `environment` */
    /* Optional (NULL / 0: rtu_scene_upload builds what it needs itself, about 1 ms per mesh node and light for 6 000 triangles):
     * prebuilt light masks.  An entry is used when the position / direction of its light in the node's coordinates still is the
     * one recorded in it; any other pair is built at upload. */
    const rtu_light_mask *light_masks; int32_t n_light_masks;
} rtu_scene_desc;

/* ------------------------------------------------------------------ batched operator I/O */

/* Ray (scene.h:59-68). */
typedef struct rtu_ray { float p[3]; float dir[3]; } rtu_ray;

/* HitInfo (scene.h:150-163) with the node pointer replaced by the pre-order node index and
 * the winning face id added (the reference does not record it). */
typedef struct rtu_hit {
    float z;
    float p[3];
    float N[3];
    float uvw[3];
    int32_t node;  /* -1 = no hit (then z == RTU_BIGFLOAT) */
    int32_t face;  /* winning triangle of a mesh hit, else -1 */
    int32_t front; /* HitInfo::front */
} rtu_hit;

/* ------------------------------------------------------------------ frame parameters */

enum {
    RTU_MODE_PRIMARY = 0, /* one Trace() per pixel centre: node/face ids + z only */
    RTU_MODE_WHITTED = 1, /* Trace + Shade(ray,h,lights,shade_bounces): RenderFunctions.cpp:135 alone */
    RTU_MODE_PATH = 2,    /* HEAD estimator: MonteCarlo GI list + lights (RenderFunctions.cpp:132-135) */
    RTU_MODE_PHOTON = 3,  /* PhotonMapping(ray, hInfo) per sample (RenderFunctions.cpp:141-142, 394-413); needs a photon map */
    RTU_MODE_PHOTON_GATHER = 4 /* Shade(ray,h,lights,bounces) + MonteCarloPhoton(h,x,y,1): direct light plus a final gather of
                                  gi_bounces cosine samples into the photon map (RenderFunctions.cpp:137-139, 416-451) */
};
enum {
    RTU_PATTERN_CENTER = 0,   /* pixel centre (0.5,0.5); spp must be 1 */
    RTU_PATTERN_REFERENCE = 1 /* s/spp + Halton(s,4), s/spp + Halton(s,5) (RenderFunctions.cpp:81-85,96) */
};
enum {
    RTU_FLAG_CULL_NULL_SHADOW_RAYS = 1, /* skip shadow rays whose contribution is exactly 0 (the reference traces them) */
    RTU_FLAG_CULL_ZERO_WEIGHT_RAYS = 2, /* skip secondary rays whose throughput is exactly 0 (e.g. absorbed TIR; the reference traces them) */
    RTU_FLAG_TIME_KERNELS = 4,          /* bracket every wave kernel with CUDA events (fills rtu_kernel_stats.ms) */
    RTU_FLAG_REFERENCE_WALK = 8         /* mesh walks through the cyBVH with the reference's own box / triangle tests and no pruning, so
                                           that the counters book exactly the work Trace() does (the figure SURVEY 8d builds the
                                           algorithmic bytes from).  Default: the meshes' own 4-wide hierarchies (same image, less work). */
};

typedef struct rtu_params {
    int32_t width, height;        /* 0 = use the scene camera's (xml <width>/<height>) */
    int32_t spp;                  /* samples per pixel; the reference hard-codes 1024 (RenderFunctions.cpp:27) */
    int32_t sample_begin;         /* this call renders samples [sample_begin, sample_end) of the spp ... */
    int32_t sample_end;           /* ... (spp-sliced multi-GPU); 0,0 = all */
    int32_t pattern;              /* RTU_PATTERN_* */
    int32_t mode;                 /* RTU_MODE_* */
    int32_t shade_bounces;        /* bounceCount passed to Shade; the reference hard-codes 5 (RenderFunctions.cpp:134) */
    int32_t gi_bounces;           /* monteCarloBounces = 4 (RenderFunctions.cpp:31) */
    int32_t row_begin, row_end;   /* image rows rendered by this call (tile-sliced multi-GPU); 0,0 = all */
    uint32_t flags;               /* RTU_FLAG_* */
    uint64_t seed;                /* counter-based RNG key (soft shadows, glossy, DOF, GI) */
    /* Adaptive sampling: the reference declares minSampleSize = 8, targetVariance = 0.005, sampleIncrement = 1
     * (RenderFunctions.cpp:25-28) and never uses them.  adaptive_min_spp > 0 turns it on for rtu_render / rtu_render_async
     * (RTU_MODE_WHITTED and RTU_MODE_PATH, whole frames): every pixel gets adaptive_min_spp samples of the spp-sample
     * pattern, then passes of adaptive_step more samples go to the 8x4-pixel tiles that still hold a pixel whose estimated
     * variance of the MEAN exceeds adaptive_target, until none is left or spp samples are spent.  The estimate comes from
     * two half-images (even / odd samples): ((mean_even - mean_odd) / 2)^2, largest channel.  A pixel's colour is the mean
     * over its own sample count (rtu_image::sample_count). */
    int32_t adaptive_min_spp;     /* 0 = fixed spp (the reference's behaviour) */
    int32_t adaptive_step;        /* samples per pass after the first; <= 0: 8 */
    float adaptive_target;        /* variance of the pixel mean to stop at */
    int32_t reserved_;
} rtu_params;

/* Host output buffers of one frame; any pointer may be NULL. */
typedef struct rtu_image {
    uint8_t *rgb8;    /* W*H*3  Result.png pixels: gamma 1/2.2 then Color24 (RenderFunctions.cpp:152-159) */
    float *rgb;       /* W*H*3  linear mean radiance (parity buffer, not in the reference) */
    float *z;         /* W*H    pixel-centre primary z, RTU_BIGFLOAT on miss (SURVEY A-3) */
    uint8_t *z8;      /* W*H    ZBuffer.png greys (RenderImage::ComputeZBufferImage, scene.h:590-612) */
    int32_t *node_id; /* W*H    primary hit node (pre-order index), -1 on miss */
    int32_t *face_id; /* W*H    primary hit face, -1 if not a mesh */
    uint8_t *sample_count; /* W*H  samples the pixel received, saturating at 255: RenderImage::sampleCount (scene.h:545,581);
                              constant without adaptive sampling */
} rtu_image;

/* Counters of the last render/trace call.  A "ray" is one root-level Trace or ShadowTrace
 * (SURVEY section 8d); box/tri/node counts feed the algorithmic-bytes roofline figure
 * 28*box_tests + 52*tri_tests + 48*node_visits. */
typedef struct rtu_kernel_stats {
    uint64_t rays;        /* rays this kernel class traced */
    uint64_t box_tests;   /* slab tests (object bound boxes + BVH child boxes) */
    uint64_t tri_tests;
    uint64_t node_visits; /* object nodes visited (one ToNodeCoords each) */
    uint64_t launches;
    double ms;            /* summed CUDA-event time of its launches; only with RTU_FLAG_TIME_KERNELS */
} rtu_kernel_stats;

typedef struct rtu_stats {
    uint64_t trace_rays;
    uint64_t shadow_rays;
    uint64_t box_tests;
    uint64_t tri_tests;
    uint64_t node_visits;
    uint64_t kernel_launches;
    double device_ms;               /* CUDA-event time of the device work of the last call */
    rtu_kernel_stats primary_wave;  /* k_extend<primary>: camera rays, closest hit */
    rtu_kernel_stats secondary_waves; /* k_extend<queue>: reflection / refraction / Fresnel rays, closest hit */
    rtu_kernel_stats shadow_waves;  /* k_shadow_wave: any-hit */
    rtu_kernel_stats shade_kernels; /* k_shade: MtlBlinn::Shade steps on the compacted hits (launches, ms only) */
    uint64_t scene_device_bytes;    /* bytes rtu_scene_upload copied host -> device */
    uint64_t pixel_samples;         /* camera samples the last frame spent (= pixels x spp without adaptive sampling) */
    double bvh_build_ms;            /* device time of the LBVH builds of the scene's upload (RTU_MESH_DEVICE_BVH meshes), else 0 */
    uint64_t queue_retries;         /* frames this context rendered again because a ray queue overflowed (the scene then
                                       remembers the larger queues, so a steady state shows no new retries) */
} rtu_stats;

typedef struct rtu_context rtu_context; /* one per GPU / host thread */
typedef struct rtu_scene rtu_scene;     /* device-resident scene, owned by a context */
typedef struct rtu_host_scene rtu_host_scene; /* result of our XML/OBJ/PNG loader (host memory) */

const char *rtu_last_error(void);
int rtu_version(void);

/* ------------------------------------------------------------------ scene front-end (host only, no CUDA)
 * Replaces LoadScene(const char*) (xmlload.cpp:64-131) + TriObj::Load (objects.h:52-60)
 * + TextureFile::Load (texture.cpp:57-91).  asset_root is prepended to relative object /
 * texture paths (the reference resolves them against the CWD). */
int rtu_host_load_xml(const char *xml_path, const char *asset_root, rtu_host_scene **out);
/* flags: RTU_LOAD_DEVICE_BVH skips both host hierarchy builds (cyBVH::Build is seconds for a million triangles) and marks
 * every mesh RTU_MESH_DEVICE_BVH. */
enum { RTU_LOAD_DEVICE_BVH = 1 };
int rtu_host_load_xml_ex(const char *xml_path, const char *asset_root, uint32_t flags, rtu_host_scene **out);
const rtu_scene_desc *rtu_host_scene_desc(const rtu_host_scene *hs);
void rtu_host_scene_destroy(rtu_host_scene *hs);
/* cyBVH build (cyBVH.h:122-142, BVHTriMesh 339-379) on caller arrays; fills caller-provided
 * boxes[(2*nf)*6], data[2*nf], elements[nf]; *n_nodes gets the entry count incl. node 0. */
int rtu_host_build_bvh(const float *v, uint32_t nv, const uint32_t *f, uint32_t nf,
                       uint32_t max_per_leaf, float *boxes, uint32_t *data, uint32_t *elements,
                       uint32_t *n_nodes);
/* Binned-SAH any-hit hierarchy over the triangles of a mesh, 4-wide nodes (not in the reference: ShadowTrace walks the
 * cyBVH there).  nodes: room for nf x 32 words, slots: nf words; *n_nodes < nf nodes are written. */
int rtu_host_build_occlusion_bvh(const float *v, uint32_t nv, const uint32_t *f, uint32_t nf, const uint32_t *bvh_elements,
                                 float *nodes, uint32_t *n_nodes, uint32_t *root, uint32_t *slots);
/* Light mask of mesh node `node` for light `light` (not in the reference: where, seen from a light that casts hard shadows,
 * the mesh can stop a shadow ray at all; rtu_scene_upload builds the same masks and the any-hit kernel skips the mesh's walk
 * for a ray whose cell is clear - lightFunctions.cpp:27-37 observes only the boolean).  bits: 2048 words
 * (256 x 256 cells, row = second image coordinate).  RTU_ERR_UNSUPPORTED: no mask for this pair (soft light, light inside the
 * mesh, margins not met); such rays are walked.  light = -1: the same for the camera rays of a camera without depth of field
 * (they all start in its position; the primary wave skips the mesh for a ray beside its silhouette).  rec: 24 words. */
int rtu_host_build_light_mask(const rtu_scene_desc *desc, int32_t node, int32_t light, float *rec, uint32_t *bits);
/* Result.png / ZBuffer.png writers (RenderImage::SaveImage/SaveZImage, scene.h:638-654). */
int rtu_write_png(const char *path, const uint8_t *pixels, int32_t width, int32_t height, int32_t channels);

/* ------------------------------------------------------------------ device path
 * stream: a cudaStream_t cast to void* (NULL = legacy default stream). */
int rtu_context_create(int32_t device, void *stream, rtu_context **out);
void rtu_context_destroy(rtu_context *ctx);

/* Pack + upload (H2D) a scene: flattens instances, re-lays the BVH out as 64-byte node
 * pairs and 64-byte triangle records, copies materials/lights/textures. */
int rtu_scene_upload(rtu_context *ctx, const rtu_scene_desc *desc, rtu_scene **out);
void rtu_scene_destroy(rtu_scene *scene);

/* Batched operators with HOST buffers (copies inside):
 *   rtu_trace        == Trace(ray, &rootNode, hInfo) with a fresh HitInfo   (RenderFunctions.cpp:181)
 *   rtu_shadow_trace == GenLight::Shadow's ShadowTrace with h.z = t_max     (lightFunctions.cpp:27-37)
 *   rtu_shade        == hit.node->GetMaterial()->Shade(ray, hit, lights, bounces) (mtlFunctions.cpp:120) */
int rtu_trace(rtu_scene *scene, const rtu_ray *rays, int64_t n, rtu_hit *hits);
int rtu_shadow_trace(rtu_scene *scene, const rtu_ray *rays, const float *t_max, int64_t n, uint8_t *occluded);
int rtu_shade(rtu_scene *scene, const rtu_ray *rays, const rtu_hit *hits, int64_t n, int32_t bounces, float *rgb);
/* The camera ray Render() builds for sample `sample` of every pixel (RenderFunctions.cpp:78-97),
 * rays[x + width*y]; lens offsets are 0 (dof is ignored here). */
int rtu_camera_rays(rtu_scene *scene, const rtu_params *params, int32_t sample, rtu_ray *rays);
/* Device self-test: the traversal kernels take the six quotients of a slab test through a
 * reciprocal hoisted out of the BVH loop; this compares that path bit for bit with the IEEE
 * division on every divisor mantissa x numerators_per_divisor x 9 exponents. */
int rtu_selftest_division(rtu_context *ctx, uint32_t numerators_per_divisor, uint64_t seed, uint64_t *tested, uint64_t *mismatches);

/* Frame level.  rtu_render: the whole Render() job with HOST output buffers (the e2e path).
 * rtu_render_device: same work, result left in device memory (accum: W*H float4 = sum of
 * radiance over the rendered samples in .xyz; .w stays 0); accum may be a caller-owned device
 * pointer or NULL for the internal one.  clear_accum == 0 ADDS the frame to what accum holds
 * (spp slices / row ranges rendered by separate calls); the frame is added only once it is known
 * to be complete, so a failed call never leaves accum half written.
 * rtu_resolve: accum -> mean over params->spp, gamma, Color24, z image; writes HOST buffers. */
void rtu_params_default(rtu_params *p);
int rtu_render(rtu_scene *scene, const rtu_params *params, rtu_image *out);
int rtu_render_device(rtu_scene *scene, const rtu_params *params, float *d_accum, int32_t clear_accum);
int rtu_resolve(rtu_scene *scene, const rtu_params *params, const float *d_accum, rtu_image *out);
int rtu_get_stats(const rtu_scene *scene, rtu_stats *out);
int rtu_synchronize(rtu_context *ctx);

/* ---- Non-blocking frame + progress (BeginRender / StopRender / numRenderedPixels: viewport.cpp:36,443,447, main.cpp:66-72,
 * scene.h:585-588) and the PNG writer that overlaps the next frame (RenderImage::SaveImage, scene.h:638-654).
 * rtu_render_async returns at once; a worker thread renders the frame in slices (groups of samples, or row blocks when
 * there are few samples).  After every slice - when a callback is given - and at the end, `out`'s buffers hold the mean over
 * what is done so far, the counter moves and the callback runs on the worker thread.  The scene's context must not be used by
 * other calls until rtu_job_wait returns.  pixels_done reaches width*height at the end like numRenderedPixels. */
typedef struct rtu_job rtu_job;
typedef void (*rtu_progress_fn)(void *user, int64_t pixels_done, int64_t pixels_total);
int rtu_render_async(rtu_scene *scene, const rtu_params *params, const rtu_image *out, rtu_progress_fn progress, void *user, rtu_job **job);
int rtu_job_progress(const rtu_job *job, int64_t *pixels_done, int64_t *pixels_total, int32_t *finished);
void rtu_job_cancel(rtu_job *job);   /* StopRender(): the frame ends after the slice in flight, rtu_job_wait returns RTU_ERR_CANCELLED */
int rtu_job_wait(rtu_job *job);      /* joins the worker; the job's status (message in rtu_last_error) */
void rtu_job_destroy(rtu_job *job);  /* waits if needed */
/* Encodes on a worker thread; the pixels are copied first, so the caller may render the next frame into the same buffer. */
int rtu_write_png_async(const char *path, const uint8_t *pixels, int32_t width, int32_t height, int32_t channels, rtu_job **job);

/* ---- Multi-GPU: one process or host thread per GPU, each with its own context.  The path shards by
 * independent units; the only data-path communication is the step that brings the partial images together on one rank.
 * The communicator wraps an NCCL communicator created from a 128-byte unique id: rank 0 calls rtu_comm_unique_id and hands
 * the bytes to the other ranks by whatever means the host program has (the reference has none: it is a single process). */
#define RTU_COMM_ID_BYTES 128
typedef struct rtu_comm rtu_comm;
int rtu_comm_unique_id(uint8_t id[RTU_COMM_ID_BYTES]);
int rtu_comm_create(rtu_context *ctx, const uint8_t id[RTU_COMM_ID_BYTES], int32_t rank, int32_t world, rtu_comm **out);
void rtu_comm_destroy(rtu_comm *comm);
/* spp slices: every rank rendered samples [sample_begin,sample_end) of ALL pixels (rtu_render_device) into d_accum (NULL =
 * the context's own accumulator).  RGB planes -> one ncclReduce(sum, FP32) onto `root` -> resolve (mean over params->spp,
 * gamma, Color24) -> the root's HOST buffers, all on the context's stream.  Collective: every rank calls it; `out` is only
 * read on the root, the other ranks return as soon as their part is enqueued. */
int rtu_reduce_resolve(rtu_scene *scene, rtu_comm *comm, const rtu_params *params, const float *d_accum, int32_t root, rtu_image *out);
/* row ranges: every rank rendered ALL samples of rows [row_begin,row_end) (disjoint, covering the image); each rank resolves
 * its rows and sends them to the root, which assembles Result.png's pixels.  No reduction. */
int rtu_gather_resolve(rtu_scene *scene, rtu_comm *comm, const rtu_params *params, const float *d_accum, int32_t root, rtu_image *out);

/* ---- Photon map (dead code at the reference's HEAD, main.cpp:31) -------------------------------
 * rtu_photon is cyPhotonMap::Photon (cyPhotonMap.h:47-66) byte for byte: a balanced array can be exchanged with the
 * reference.  plane_dirz: bits 0-1 splitting axis of the kd-tree node, bit 3 = direction z is negative. */
typedef struct rtu_photon {
    float position[3];
    float power;          /* max(r,g,b) of the photon power */
    uint8_t color[3];     /* Color24(power_rgb / power) */
    uint8_t plane_dirz;
    int16_t dir_x, dir_y; /* short(dir * 0x7FFF) */
} rtu_photon;

typedef struct rtu_photon_params {
    uint32_t map_size;    /* photonMapSize   = 1000000 (RenderFunctions.cpp:32) */
    uint32_t max_bounce;  /* photonMaxBounce = 10      (:34) */
    float est_radius;     /* photonEstRadius = 1       (:35) */
    float ellipticity;    /* photonEllipticity = 0.5   (:36); at most photonSampleSize = 100 photons per estimate (:33) */
    uint64_t seed;
} rtu_photon_params;

typedef struct rtu_photon_stats {
    uint64_t paths;            /* photon paths emitted (RandomPhoton calls) */
    uint64_t from_light;       /* photonFromLight: paths whose first segment hit something (:357) */
    uint64_t stored;           /* photons in the map */
    uint64_t trace_rays;       /* root-level Trace calls of the emission */
    float scale_factor;        /* (lights[0] intensity / photonFromLight).Gray()  (:384) */
    float emit_ms, build_ms;   /* device time of the emission; wall time of the kd-tree balancing (+ the estimate's walk records) */
    uint32_t device_build;     /* 1: the kd-tree was balanced on the device; 0: on the host (a median tied with a neighbour) */
} rtu_photon_stats;

void rtu_photon_params_default(rtu_photon_params *p);
/* GeneratePhotonMap(): emits from lights[0] (must be a point light) until map_size photons are stored, scales the
 * powers, balances the kd-tree (RenderFunctions.cpp:341-392).  Paths are numbered; path i draws from its own
 * Philox stream and the map holds exactly what the sequential loop over i = 0,1,2,... would have stored. */
int rtu_photon_map_generate(rtu_scene *scene, const rtu_photon_params *params, rtu_photon_stats *stats);
/* Installs caller photons (any order, e.g. the reference's own map): PrepareForIrradianceEstimation.  The tree is balanced on
 * the device where that is the reference's tree by construction (no median ties with a neighbour along its split axis), else
 * on the host; either way the array is byte-identical to cyPhotonMap's. */
int rtu_photon_map_set(rtu_scene *scene, const rtu_photon *photons, uint32_t n, const rtu_photon_params *params);
/* Number of photons of the scene's map (0: none) and whether the device balanced it. */
int rtu_photon_map_info(const rtu_scene *scene, uint32_t *n_photons, uint32_t *device_build);
/* Copies the balanced map out: out[i] = photons[i+1] of cyPhotonMap, *n = number of photons (cap = room in out). */
int rtu_photon_map_get(rtu_scene *scene, rtu_photon *out, uint32_t cap, uint32_t *n);
/* cyPhotonMap::EstimateIrradiance<100>(irrad, direction, radius, pos, &normal, ellipticity, FILTER_TYPE_CONSTANT)
 * for n query points (cyPhotonMap.h:276-323); found[i] = photons used (may be NULL). */
int rtu_estimate_irradiance(rtu_scene *scene, const float *pos, const float *normal, int64_t n, float radius, float ellipticity,
                            float *irrad, float *direction, int32_t *found);
/* Host only: cyPhotonMap::PrepareForIrradianceEstimation on caller arrays (cyPhotonMap.h:207-274).
 * in: n photons; out: n+1 records, out[0] unused, out[1..n] the left-balanced kd-tree in heap order. */
int rtu_host_balance_photons(const rtu_photon *in, uint32_t n, rtu_photon *out);

#ifdef __cplusplus
}
#endif
#endif /* RTU_H_INCLUDED */
