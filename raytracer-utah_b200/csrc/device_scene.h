// Device-side scene layout (HBM), shared by the packer (rtu_api.cu) and the kernels.
//
// Everything the traversal loop touches is a 16-byte-aligned record read with 128-bit loads:
//   DNode     (112 B)  one scene-graph node in pre-order: the "to" transform (itm,pos) used by
//                      Node::ToNodeCoords and the "from" transform (tm, itm^T) used by FromNodeCoords
//   BvhPair   ( 64 B)  the two adjacent children of one internal cyBVH node: both boxes + both
//                      child words, so one pop of an internal node costs exactly one 64-byte fetch
//                      (the reference reads 2 x 28 B nodes + the parent's child index)
//   TriRec    ( 48 B)  one triangle in LEAF order with everything ray-independent pre-evaluated on
//                      the host with the reference's float op order (unit normal, projection axis,
//                      projected edges, projected area): objFunctions.cpp:259-300
//   TriShade  (112 B)  the 3 positions / normals / texture vertices of the same slot, fetched only
//                      for the winning triangle (objFunctions.cpp:317-320)
#pragma once
#include <stdint.h>

#define RTU_MAX_DEPTH 8          // scene-graph nesting handled by the generic (non-flat) path
#ifndef RTU_STACK
#define RTU_STACK 64             // BVH traversal stack entries (reference: 100; depth-first needs depth+1)
#endif
#define RTU_KIND_BITS 3

struct __align__(16) DNode {
    float itm[9];   // ToNodeCoords:   p' = itm*(p-pos)                  scene.h:235,501-507
    float pos[3];
    float tm[9];    // FromNodeCoords: p  = tm*p + pos; N = norm(itm^T N) scene.h:236,242,508-512
    int32_t parent; // pre-order index, -1 root
    int32_t depth;  // 0 root
    int32_t kind;   // RTU_OBJ_*
    int32_t mesh;
    int32_t material;
    int32_t mask_first; // mesh nodes: this node's LightMasks in DScene::light_masks ...
    int32_t mask_count; // ... low byte: one per light that casts hard shadows; RTU_MASK_HAS_EYE: one more for the camera rays
};

// Where a mesh can shadow at all, as seen from one light (built by host/light_mask.cpp, which states the margins).  Every
// shadow ray of a point light without size ends in the light (lightFunctions.cpp:76-78), every shadow ray of a directional
// light runs against its direction (lights.h:48): seen from the light such a ray is one POINT of the perspective image around
// the axis light -> mesh / of the orthographic image along the direction.  The mask is a MASK_RES x MASK_RES bitmap over the
// image of the mesh, a bit set where the image of a triangle, grown by a cell, touches the cell.  A ray whose cell is clear
// cannot meet a triangle, whatever the hierarchy walk would find: the walk is skipped (any-hit only; result-neutral; in the
// node's local coordinates, where the walks happen).
#define RTU_MASK_RES 256
#define RTU_MASK_HAS_EYE 0x100
#define RTU_MASKS_PER_NODE 4 // scenes with more hard lights get no masks: the lookup steps through a node's masks
struct __align__(16) LightMask {
    float L[3];      // point light: position; directional light: direction (both node-local)
    int32_t kind;    // RTU_LIGHT_POINT / RTU_LIGHT_DIRECT / 3: rays that START in L (the camera without depth of field)
    float a[3];      // axis light -> mesh centre (point light only)
    float u0;
    float e1[3];
    float v0;
    float e2[3];
    float su;        // cells per unit of u
    float sv;
    uint32_t bits;   // first word of this mask's bitmap in DScene::mask_bits
    float lim;       // largest 1-norm of (origin - light) / of the origin for which the mask's margins hold; eye: |origin - L|_1
    float zmargin;   // light lists: slack of the depth cut (rounding of the origin's depth on the device)
    uint32_t cells;  // light lists: first of the RES * RES + 1 cell offsets in DScene::mask_lists; 0xffffffff: none
    uint32_t items;  // ... first word of the (slot, least depth) pairs
    float pad[2];
};

// child word: bit31 = leaf; leaf: bits 28..30 = count-1, bits 0..27 = first triangle slot
//             internal: index of that child's BvhPair
struct __align__(16) BvhPair {
    float b1[6];    // child 1 box: min xyz, max xyz
    float b2[6];    // child 2 box
    uint32_t c1, c2;
    uint32_t up;    // cyBVH pairs only: the pair that holds this node's own box | (1u << 31 if it is that pair's child 2);
                    // 0xffffffff for the root's pair (the root's box is never tested, objFunctions.cpp:343)
    uint32_t pad;
};

// One node of a mesh's any-hit hierarchy (host/occlusion_bvh.cpp): the boxes of up to four children, one coordinate of all
// four per float4, and their child words; an unused slot has an inverted box and the word 0x7fffffff.  One visit = one
// 128-byte line.
struct __align__(16) OccNode {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    uint32_t child[4];
    uint32_t pad[4];
};

struct __align__(16) TriRec {
    float nx, ny, nz, ax;       // unit geometric normal, A.x
    float ay, az, area, fbits;  // A.y, A.z, projected area of ABC (already /2), face id | axis<<30 as int bits
    float cau, cav, bau, bav;   // projected C-A and B-A
};

struct __align__(16) TriShade {
    float v[9];   // A, B, C
    float vn[9];  // normals at A, B, C
    float vt[9];  // texture vertices at A, B, C
    float pad;
};

struct DMesh {
    const BvhPair *pairs;
    const TriRec *tris;
    const TriShade *shade;
    uint32_t root;       // child word of the root node (leaf meshes have no pairs)
    uint32_t n_pairs;
    uint32_t n_tris;
    float bmin[3], bmax[3];
    uint32_t empty;      // no faces: cyTriMesh's "not ready" box never intersects
    uint32_t coords_ok;  // every box coordinate of the mesh is 0 or at least 2^-36 in magnitude (see mesh_invdir)
    // any-hit hierarchy (host/occlusion_bvh.cpp): 4-wide binned-SAH nodes over the same triangles, its triangle records in
    // ITS leaf order (fbits: projection axis << 30 | cyBVH slot), and per cyBVH slot the pair that holds the box of the
    // slot's leaf (| 1u << 31: child 2) - the start of the ancestor chain ref_reaches() climbs
    const OccNode *occ_nodes;
    const TriRec *occ_tris;
    const uint32_t *tri_up;
    uint32_t occ_root;
    uint32_t no_ref;     // RTU_MESH_DEVICE_BVH: no cyBVH (pairs, tri_up are NULL); occ_* is an LBVH built on the device, tris / shade
                         // are in FACE order and a triangle's slot is its face index
    uint32_t nested;     // every cyBVH box contains the boxes of its children (then a leaf's box implies its ancestors', ref_reaches)
    float occ_scale;     // largest |coordinate| of the mesh's bound box (scale of the per-ray conservative margin)
};

struct DTexMap {
    int32_t kind;
    float itm[9];
    float pos[3];
    float c1[3], c2[3];
    const uint8_t *rgb8;
    int32_t width, height;
};

struct DTexColor {
    float c[3];
    int32_t map; // -1 none
};

struct DMaterial {
    DTexColor diffuse, specular, reflection, refraction;
    float glossiness;
    float absorption[3];
    float ior;
    float refl_gloss, refr_gloss;
};

struct DLight {
    int32_t kind;
    float I[3];
    float v[3];
    float size;
};

// camera frame pre-evaluated on the host exactly like CalculateImageOrigin / CalculateCurrentPoint
// (RenderFunctions.cpp:243-269)
struct DCamera {
    float pos[3];
    float origin[3]; // top-left corner of the image plane
    float u[3], v[3]; // one-pixel steps
    float lens_x[3], lens_y[3]; // thin-lens basis (RenderFunctions.cpp:93)
    float dof;
    int32_t width, height;
    float inv_w, inv_h; // not used for parity-critical math
};

// Top-level hierarchy over the objects' bounding spheres (root space), used when a scene has many nodes: it only
// NOMINATES nodes for a ray (a superset of those the per-node sphere cull lets through); the nominees are then visited in
// node order by the usual code, so results and counters are those of the linear visit of every node.
struct TopNode {
    float lo[3];
    int32_t a;  // internal: left child; leaf: -(first item + 1)
    float hi[3];
    int32_t b;  // internal: right child; leaf: item count
};
#define RTU_TOP_CAND 64 // nominees per ray before the linear visit takes over

struct DScene {
    const DNode *nodes;
    const float4 *bounds; // per node: bounding sphere of the object's bound box in root space (xyz, r^2 inflated);
                          // w = -1: no object, w = -2: object that can never be hit (empty mesh)
    int32_t n_nodes;
    int32_t flat;        // 1: every object node hangs directly off the root
    int32_t n_obj;       // nodes that hold an object
    int32_t n_top;       // nodes of the top-level hierarchy (0: not built, every node is visited linearly)
    const TopNode *top;
    const int32_t *top_items;
    const float4 *top_bounds; // bounds[] of the items, in the order of top_items (no dependent load in a leaf)
    const LightMask *light_masks;
    const uint32_t *mask_bits;
    int32_t scan_min;           // lanes that must reach a light list together for it to be scanned in the node loop (else: with a batch)
    const uint32_t *mask_lists; // light lists of the masks that have them: cell offsets and (slot, depth) pairs (LightMask::cells / items)
    const int32_t *obj_rank; // per node: number of object nodes with index <= that node
    int32_t any_no_ref;  // some mesh has no cyBVH: RTU_FLAG_REFERENCE_WALK cannot be honoured
    int32_t pool_ok;     // 1: every mesh fits the item encoding of the pooled shadow kernel (<= 2^24 triangles, < 2^27 pairs)
    const DMesh *meshes;
    const DMaterial *materials;
    int32_t n_materials;
    const DLight *lights;
    int32_t n_lights;
    int32_t n_shadow_lights; // lights that send a shadow ray (everything but ambient)
    const DTexMap *texmaps;
    DTexColor background, environment;
    float cam_pos[3];    // Shade uses camera.pos for the view vector (mtlFunctions.cpp:137)
};

// --- wavefront queues (structure of float4 arrays: every lane's 16-byte store is contiguous with its neighbour's)
struct RayQueue {        // closest-hit rays of the next wave
    float4 *o;  // origin.xyz, pixel (int bits)
    float4 *d;  // dir.xyz, meta (int bits): kind | bounce<<3 | material<<8
    float4 *w;  // throughput rgb, aux index (int bits, -1 none)
    uint32_t *path; // RNG path word of the ray
    uint32_t *count;
    uint32_t cap;
};
struct AuxPool {         // extra payload of refracted / Fresnel rays
    float4 *a;  // Kt.rgb, Schlick F
    float4 *b;  // mirror direction at the parent hit, unused
    uint32_t *count;
    uint32_t cap;
};
struct HitQueue {        // compacted closest hits of one wave, consumed by k_shade
    float4 *a;  // z, node, front, triangle slot (int bits)
    float4 *b;  // bc1, bc2, bc3, index of the ray in the wave's input (uint bits)
    uint32_t *count;
    uint32_t cap;
};
struct ShadowQueue {     // any-hit rays
    float4 *o;  // origin.xyz, pixel
    float4 *d;  // dir.xyz, t_max
    float4 *c;  // contribution rgb if unoccluded
    uint32_t *count;
    uint32_t cap;
};

enum RayKind {
    RK_PRIMARY = 0,
    RK_REFRACT = 1,  // refracted ray: miss -> env(dir); hit -> Beer * Kt * (1-F), then spawns RK_FRESNEL
    RK_FRESNEL = 2,  // mirror ray spawned only after the refracted ray hit (mtlFunctions.cpp:234-251)
    RK_TIR = 3,      // total internal reflection: miss adds nothing (mtlFunctions.cpp:205-222)
    RK_REFLECT = 4,  // Kr mirror ray: miss -> env(dir)*Kr colour (mtlFunctions.cpp:273-290)
    RK_GI = 5        // cosine-hemisphere bounce of MonteCarlo() (RenderFunctions.cpp:561-575)
};

// The photon map (cyPhotonMap): n+1 records in heap order, record 0 unused; nodes below `half` are internal
// (halfStoredPhotons = n/2 - 1, cyPhotonMap.h:227,356)
struct DPhotonMap {
    const rtu_photon *map;
    const float4 *knn_nodes, *knn_dir, *knn_pw; // walk records and per-photon tables of the estimate (photon_kernels.cu)
    int n, half;
    float radius, norm_scale; // photonEstRadius; 1/ellipticity - 1 (0 when ellipticity == 1)
};

// counters per kernel class: 0 = primary wave (and the batched closest-hit operators),
// 1 = secondary (queue) waves, 2 = shadow waves (and the batched any-hit operator)
struct DCounterBlock {
    unsigned long long trace_rays, shadow_rays, box_tests, tri_tests, node_visits;
};
struct DCounters {
    DCounterBlock k[3];
    uint32_t overflow;
    uint32_t pad;
};
