// The reference-side binding: what a RayTracer-Utah maintainer adds next to main.cpp to run the
// render path on the GPU.  It walks the reference's own globals (main.cpp:17-27) after LoadScene(),
// fills the flattened rtu_scene_desc of include/rtu.h, and replaces SpawnRenderThreads()
// (main.cpp:29-64) by rtu_scene_upload + rtu_render, writing into renderImage so that the
// reference's own SaveImage / ComputeZBufferImage / SaveZImage produce Result.png and ZBuffer.png.
//
// The library is dlopen()ed, so this file needs neither CUDA nor our headers' implementation at
// link time.  MtlBlinn / light / TriObj fields are private in the reference (materials.h:51-56,
// lights.h:38-97, objects.h:62-63); a maintainer would add accessors or a friend declaration, this
// example takes the same shortcut as the test harness (#define private public on those headers).
//
// Built into oracle/_ref/ref_harness (mode "gpu"); INTEGRATION.md quotes it.
#include "std_first.h"
#include <dlfcn.h>
#define private public
#define protected public
#include "ExternalLibrary/scene.h"
#include "ExternalLibrary/objects.h"
#include "ExternalLibrary/materials.h"
#include "ExternalLibrary/lights.h"
#include "ExternalLibrary/texture.h"
#undef private
#undef protected
#include "../../include/rtu.h"

extern RenderImage renderImage;
extern Camera camera;
extern Sphere theSphere;
extern Plane thePlane;
extern Node rootNode;
extern MaterialList materials;
extern LightList lights;
extern TexturedColor background;
extern TexturedColor environment;

namespace {

struct Packed {
    rtu_scene_desc desc;
    std::vector<rtu_node> nodes;
    std::vector<rtu_mesh> meshes;
    std::vector<const TriObj *> mesh_objs;
    std::vector<rtu_material> mats;
    std::vector<rtu_light> lts;
    std::vector<rtu_texmap> texmaps;
    std::vector<std::vector<float>> bvh_boxes;
    std::vector<std::vector<uint32_t>> bvh_data;
};

void copy3(float *d, const Point3 &p) { d[0] = p.x; d[1] = p.y; d[2] = p.z; }
void copyc(float *d, const Color &c) { d[0] = c.r; d[1] = c.g; d[2] = c.b; }

int pack_texmap(Packed &P, const TextureMap *tm)
{
    rtu_texmap m;
    memset(&m, 0, sizeof m);
    memcpy(m.itm, tm->GetInverseTransform().data, sizeof m.itm);
    copy3(m.pos, tm->GetPosition());
    m.kind = RTU_TEX_NULL;
    if (const TextureChecker *c = dynamic_cast<const TextureChecker *>(tm->texture)) {
        m.kind = RTU_TEX_CHECKER;
        copyc(m.color1, c->color1);
        copyc(m.color2, c->color2);
    } else if (const TextureFile *f = dynamic_cast<const TextureFile *>(tm->texture)) {
        m.kind = RTU_TEX_FILE;
        m.rgb8 = &f->data[0].r;
        m.width = f->width;
        m.height = f->height;
    }
    P.texmaps.push_back(m);
    return (int)P.texmaps.size() - 1;
}

rtu_texcolor pack_tc(Packed &P, const TexturedColor &t)
{
    rtu_texcolor o;
    copyc(o.color, t.GetColor());
    o.texmap = t.GetTexture() ? pack_texmap(P, t.GetTexture()) : -1;
    return o;
}

int pack_mesh(Packed &P, const TriObj *t)
{
    for (size_t i = 0; i < P.mesh_objs.size(); i++) if (P.mesh_objs[i] == t) return (int)i;
    rtu_mesh m;
    memset(&m, 0, sizeof m);
    m.v = &t->v[0].x;   m.nv = t->NV();
    m.vn = t->vn ? &t->vn[0].x : nullptr; m.nvn = t->NVN();
    m.vt = t->vt ? &t->vt[0].x : nullptr; m.nvt = t->NVT();
    m.f = t->f ? t->f[0].v : nullptr;
    m.fn = t->fn ? t->fn[0].v : nullptr;
    m.ft = t->ft ? t->ft[0].v : nullptr;
    m.nf = t->NF();
    // cyBVH keeps 28-byte nodes (6 floats + word); the ABI wants boxes and words as two arrays
    unsigned maxNode = 0;
    if (t->NF() > 0) {
        std::vector<unsigned> st{1};
        while (!st.empty()) {
            unsigned n = st.back(); st.pop_back();
            if (n > maxNode) maxNode = n;
            if (!t->bvh.IsLeafNode(n)) { st.push_back(t->bvh.GetFirstChildNode(n)); st.push_back(t->bvh.GetSecondChildNode(n)); }
        }
    }
    P.bvh_boxes.emplace_back((maxNode + 1) * 6, 0.f);
    P.bvh_data.emplace_back(maxNode + 1, 0u);
    for (unsigned n = 1; n <= maxNode; n++) {
        memcpy(&P.bvh_boxes.back()[n * 6], t->bvh.GetNodeBounds(n), 6 * sizeof(float));
        P.bvh_data.back()[n] = t->bvh.nodes[n].data;
    }
    m.bvh_boxes = P.bvh_boxes.back().data();
    m.bvh_data = P.bvh_data.back().data();
    m.bvh_nodes = maxNode + 1;
    m.bvh_elements = t->bvh.elements;
    copy3(m.bound_min, t->GetBoundMin());
    copy3(m.bound_max, t->GetBoundMax());
    P.meshes.push_back(m);
    P.mesh_objs.push_back(t);
    return (int)P.meshes.size() - 1;
}

void pack_node(Packed &P, const Node *n, int parent)
{
    rtu_node o;
    memset(&o, 0, sizeof o);
    memcpy(o.tm, n->GetTransform().data, sizeof o.tm);
    memcpy(o.itm, n->GetInverseTransform().data, sizeof o.itm);
    copy3(o.pos, n->GetPosition());
    o.parent = parent;
    o.mesh = -1;
    o.material = -1;
    const Object *obj = n->GetNodeObj();
    o.kind = !obj ? RTU_OBJ_NONE : obj == &theSphere ? RTU_OBJ_SPHERE : obj == &thePlane ? RTU_OBJ_PLANE : RTU_OBJ_MESH;
    if (o.kind == RTU_OBJ_MESH) o.mesh = pack_mesh(P, (const TriObj *)obj);
    for (size_t k = 0; k < materials.size(); k++) if (materials[k] == n->GetMaterial()) o.material = (int)k;
    int me = (int)P.nodes.size();
    P.nodes.push_back(o);
    for (int i = 0; i < n->GetNumChild(); i++) pack_node(P, n->GetChild(i), me); // pre-order = Trace()'s order
}

void pack_scene(Packed &P)
{
    memset(&P.desc, 0, sizeof P.desc);
    pack_node(P, &rootNode, -1);
    for (size_t k = 0; k < materials.size(); k++) {
        const MtlBlinn *b = dynamic_cast<const MtlBlinn *>(materials[k]);
        if (!b) { const MultiMtl *mm = dynamic_cast<const MultiMtl *>(materials[k]); b = mm && !mm->mtls.empty() ? dynamic_cast<const MtlBlinn *>(mm->mtls[0]) : nullptr; }
        rtu_material m;
        memset(&m, 0, sizeof m);
        if (b) {
            m.diffuse = pack_tc(P, b->diffuse); m.specular = pack_tc(P, b->specular); m.reflection = pack_tc(P, b->reflection);
            m.refraction = pack_tc(P, b->refraction); m.emission = pack_tc(P, b->emission);
            m.glossiness = b->glossiness; copyc(m.absorption, b->absorption); m.ior = b->ior;
            m.reflection_glossiness = b->reflectionGlossiness; m.refraction_glossiness = b->refractionGlossiness;
        }
        P.mats.push_back(m);
    }
    for (size_t k = 0; k < lights.size(); k++) {
        rtu_light l;
        memset(&l, 0, sizeof l);
        if (const AmbientLight *a = dynamic_cast<const AmbientLight *>(lights[k])) { l.kind = RTU_LIGHT_AMBIENT; copyc(l.intensity, a->intensity); }
        else if (const DirectLight *d = dynamic_cast<const DirectLight *>(lights[k])) { l.kind = RTU_LIGHT_DIRECT; copyc(l.intensity, d->intensity); copy3(l.v, d->direction); }
        else if (const PointLight *p = dynamic_cast<const PointLight *>(lights[k])) { l.kind = RTU_LIGHT_POINT; copyc(l.intensity, p->intensity); copy3(l.v, p->position); l.size = p->size; }
        P.lts.push_back(l);
    }
    P.desc.background = pack_tc(P, background);
    P.desc.environment = pack_tc(P, environment);
    rtu_camera &c = P.desc.camera;
    copy3(c.pos, camera.pos); copy3(c.dir, camera.dir); copy3(c.up, camera.up);
    c.fov = camera.fov; c.focaldist = camera.focaldist; c.dof = camera.dof;
    c.width = camera.imgWidth; c.height = camera.imgHeight;
    P.desc.nodes = P.nodes.data();         P.desc.n_nodes = (int)P.nodes.size();
    P.desc.meshes = P.meshes.data();       P.desc.n_meshes = (int)P.meshes.size();
    P.desc.materials = P.mats.data();      P.desc.n_materials = (int)P.mats.size();
    P.desc.lights = P.lts.data();          P.desc.n_lights = (int)P.lts.size();
    P.desc.texmaps = P.texmaps.data();     P.desc.n_texmaps = (int)P.texmaps.size();
}

template <class F> F sym(void *lib, const char *name)
{
    void *p = dlsym(lib, name);
    if (!p) { fprintf(stderr, "librtu_b200.so: missing %s\n", name); exit(3); }
    return (F)p;
}

} // namespace

// Replacement of SpawnRenderThreads() (main.cpp:29-64).  The reference hard-codes what Render() estimates: at HEAD it is the
// Monte-Carlo estimator of RenderFunctions.cpp:129-135 (MonteCarlo() with 4 bounces folded into an ambient light, then two
// Shade calls) at maxSampleSize = 1024 samples of the Halton(4,5) pattern, bounceCount 5 (RenderFunctions.cpp:26-31,134).
// Those constants are the defaults here: estimator = RTU_MODE_PATH, spp = 1024, bounces = 5, gi_bounces = 4.  The frame runs
// on the library's worker thread (rtu_render_async, "renderer must run in a separate thread", viewport.cpp:36); the progress
// callback moves renderImage's numRenderedPixels, which the viewport polls (scene.h:585-588, viewport.cpp:390-410), and the
// image in renderImage is refreshed after every slice like the reference's threads fill it pixel by pixel.
struct RtuProgress { long long last; };
static void RtuOnProgress(void *user, int64_t done, int64_t total)
{
    RtuProgress *pr = (RtuProgress *)user;
    (void)total;
    if (done > pr->last) { renderImage.IncrementNumRenderPixel((int)(done - pr->last)); pr->last = done; }
}

int RtuBeginRender(const char *lib_path, int estimator, int spp, int bounces, int gi_bounces, bool reference_pattern, double *device_ms,
                   unsigned long long *rays)
{
    void *lib = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "cannot load %s: %s\n", lib_path, dlerror()); return 3; }
    auto last_error = sym<const char *(*)(void)>(lib, "rtu_last_error");
    auto ctx_create = sym<int (*)(int32_t, void *, rtu_context **)>(lib, "rtu_context_create");
    auto ctx_destroy = sym<void (*)(rtu_context *)>(lib, "rtu_context_destroy");
    auto upload = sym<int (*)(rtu_context *, const rtu_scene_desc *, rtu_scene **)>(lib, "rtu_scene_upload");
    auto destroy = sym<void (*)(rtu_scene *)>(lib, "rtu_scene_destroy");
    auto params_default = sym<void (*)(rtu_params *)>(lib, "rtu_params_default");
    auto render_async = sym<int (*)(rtu_scene *, const rtu_params *, const rtu_image *, rtu_progress_fn, void *, rtu_job **)>(lib, "rtu_render_async");
    auto job_wait = sym<int (*)(rtu_job *)>(lib, "rtu_job_wait");
    auto job_destroy = sym<void (*)(rtu_job *)>(lib, "rtu_job_destroy");
    auto get_stats = sym<int (*)(const rtu_scene *, rtu_stats *)>(lib, "rtu_get_stats");

    Packed P;
    pack_scene(P);
    rtu_context *ctx = nullptr;
    rtu_scene *sc = nullptr;
    int rc = ctx_create(0, nullptr, &ctx);
    if (!rc) rc = upload(ctx, &P.desc, &sc);
    if (rc) { fprintf(stderr, "rtu: %s\n", last_error()); return rc; }
    rtu_params p;
    params_default(&p);
    p.spp = spp;
    p.shade_bounces = bounces;
    p.gi_bounces = gi_bounces;
    p.pattern = reference_pattern ? RTU_PATTERN_REFERENCE : RTU_PATTERN_CENTER;
    p.mode = estimator;
    rtu_image img;
    memset(&img, 0, sizeof img);
    img.rgb8 = &renderImage.GetPixels()[0].r;   // Color24[W*H], row 0 first (scene.h:542,578)
    img.z = renderImage.GetZBuffer();           // float[W*H], BIGFLOAT on miss (scene.h:543,579)
    RtuProgress pr = {0};
    rtu_job *job = nullptr;
    rc = render_async(sc, &p, &img, RtuOnProgress, &pr, &job);   // BeginRender() returns here; the viewport keeps polling
    if (!rc) rc = job_wait(job);                                 // (this headless caller has nothing else to do)
    if (rc) { fprintf(stderr, "rtu_render_async: %s\n", last_error()); return rc; }
    job_destroy(job);
    rtu_stats st;
    get_stats(sc, &st);
    if (device_ms) *device_ms = st.device_ms;
    if (rays) *rays = st.trace_rays + st.shadow_rays;
    destroy(sc);
    ctx_destroy(ctx);
    dlclose(lib);
    return 0;
}
