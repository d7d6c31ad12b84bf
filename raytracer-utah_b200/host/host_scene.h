// Owning host-side containers behind the POD rtu_scene_desc of include/rtu.h.
#pragma once
#include <cstdint>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtu.h"

namespace rtu {

// The any-hit hierarchy of a mesh (occlusion_bvh.cpp): binned-SAH tree collapsed to 4-wide nodes, device layout
struct OccBvh {
    std::vector<float> nodes;    // 32 words per node: lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] child[4] pad[4]
    std::vector<uint32_t> slots; // leaf-ordered cyBVH slots (indices into bvh_elements)
    uint32_t root = 0;           // child word of the root: node 0, or a leaf word for a mesh of at most 4 triangles
};
void build_occlusion_bvh(const float *v, const uint32_t *f, const uint32_t *elements, uint32_t nf, OccBvh *out);

// Where a mesh can cast a hard shadow at all, seen from one light (light_mask.cpp): the 20 words of a device LightMask and its
// bitmap.  chain = the nodes from the root down to the mesh's node.  false: no mask for this pair, the rays are walked.
// eye: `light` is a point light standing in for the camera position; the mask is for rays that start there (kind RTU_MASK_EYE).
enum { RTU_MASK_EYE = 3 };
struct LightLists {
    std::vector<uint32_t> cell_start; // 256 * 256 + 1 offsets, empty: no lists for this pair
    std::vector<uint32_t> items;      // (slot, least depth) pairs, per cell in the order of depth
};
bool build_light_mask(const rtu_node *const *chain, int n_chain, const rtu_mesh &m, const rtu_light &light, float *rec,
                      std::vector<uint32_t> *bits, bool eye = false, LightLists *lists = nullptr);
// Every mask of a scene, in the order rtu_scene_upload lays them out (by node; per node the lights in their order, the
// camera last): entries of d.light_masks that still fit the scene are taken as they are, the others are built into `own`.
// RTU_LIGHT_MASKS=0 in the environment: none.
struct OwnedMask { std::vector<uint32_t> bits; LightLists lists; };
void collect_light_masks(const rtu_scene_desc &d, std::vector<rtu_light_mask> *out, std::vector<OwnedMask> *own, uint64_t budget);

// cyTriMesh::Mtl (cyTriMesh.h:74-103): one material of an OBJ's .mtl library, with the constructor's defaults
struct ObjMtl {
    std::string name;
    float Ka[3] = {0, 0, 0}, Kd[3] = {1, 1, 1}, Ks[3] = {0, 0, 0}, Tf[3] = {0, 0, 0};
    float Ns = 0, Ni = 1;
    int illum = 2;
    std::string map_Kd, map_Ks;
    bool has_map_Kd = false, has_map_Ks = false; // Str::data != nullptr: the command appeared
};

// cyTriMesh arrays + cyBVH arrays of one TriObj (objects.h:46-66)
struct HostMesh {
    std::string name;
    std::vector<ObjMtl> mtls;            // cyTriMesh::m, only when the OBJ was loaded with loadMtl (xmlload.cpp:204)
    std::vector<int> mcfc;               // material cumulative face count (faces are grouped by material, cyTriMesh.h:468-493)
    std::vector<float> v, vn, vt;        // xyz triples
    std::vector<uint32_t> f, fn, ft;     // index triples
    std::vector<float> bvh_boxes;        // (n_nodes) x 6, node 0 unused
    std::vector<uint32_t> bvh_data;
    std::vector<uint32_t> bvh_elements;
    OccBvh occ;                          // any-hit hierarchy over the same triangles (built at load time)
    bool device_bvh = false;             // RTU_LOAD_DEVICE_BVH: no host hierarchies, the device builds an LBVH at upload
    float bound_min[3] = {1, 1, 1}, bound_max[3] = {0, 0, 0}; // cyTriMesh.h:128 "not ready" box
    uint32_t nf() const { return (uint32_t)(f.size() / 3); }
};

struct HostTexture {
    std::string name;
    int kind = RTU_TEX_NULL;
    float color1[3] = {0, 0, 0}, color2[3] = {1, 1, 1};
    std::vector<uint8_t> rgb8;
    int width = 0, height = 0;
};

// Loads an OBJ with the parsing rules of cyTriMesh::LoadFromFileObj (cyTriMesh.h:263-547),
// then TriObj::Load's post-steps (objects.h:52-60): ComputeNormals if none, bounding box,
// BVH with max 4 elements per leaf.  Returns false if the file cannot be opened.
// load_mtl: TriObj::Load(name, loadMtl) - usemtl / mtllib are honoured only then (cyTriMesh.h:439-448, 499-544).
bool load_obj_mesh(const char *path, HostMesh *out, std::string *err, bool load_mtl = false, bool build_hierarchies = true);
void compute_vertex_normals(HostMesh *m);
void compute_bounds(HostMesh *m);
// cyBVH::Build (cyBVH.h:122-142) + BVHTriMesh element callbacks (:356-375)
void build_bvh(const float *v, const uint32_t *f, uint32_t nf, uint32_t max_per_leaf,
               std::vector<float> *boxes, std::vector<uint32_t> *data, std::vector<uint32_t> *elements);

bool decode_png_rgb8(const char *path, std::vector<uint8_t> *rgb, int *w, int *h, std::string *err);
bool decode_ppm_rgb8(const char *path, std::vector<uint8_t> *rgb, int *w, int *h, std::string *err); // LoadPPM, texture.cpp:32-55
bool encode_png(const char *path, const uint8_t *px, int w, int h, int channels, std::string *err);

void set_error(const std::string &msg);
const std::string &last_error();

// Body of an extern "C" entry point: no exception (std::bad_alloc from a vector sized by a file header, std::system_error
// from std::thread, ...) may cross the C ABI; it becomes a status and a message like every other failure.
template <class F> int guarded(const char *what, F body)
{
    try {
        return body();
    } catch (const std::bad_alloc &) {
        set_error(std::string(what) + ": out of host memory");
    } catch (const std::exception &e) {
        set_error(std::string(what) + ": " + e.what());
    } catch (...) {
        set_error(std::string(what) + ": unknown exception");
    }
    return RTU_ERR_INVALID;
}

} // namespace rtu

struct rtu_host_scene {
    rtu_scene_desc desc;
    std::vector<rtu_node> nodes;
    std::vector<std::string> node_names;
    std::vector<std::unique_ptr<rtu::HostMesh>> meshes;
    std::vector<rtu_mesh> mesh_descs;
    std::vector<rtu_material> materials;
    std::vector<std::string> material_names;
    std::vector<rtu_light> lights;
    std::vector<std::unique_ptr<rtu::HostTexture>> textures; // TextureList (scene.h:368)
    std::vector<rtu_texmap> texmaps;
    std::vector<rtu_light_mask> light_masks; // light_mask.cpp: built once here, rtu_scene_upload takes them from desc
    std::vector<rtu::OwnedMask> light_mask_data;
    void finalize(); // points desc at the vectors
};
