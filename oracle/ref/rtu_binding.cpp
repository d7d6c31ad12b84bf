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
#include <map>
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


// ---------------------------------------------------------------------------------------------------------------------------
// Operator-level drop-ins (SURVEY 8b "operator surface to preserve").  The reference's own Render() / Trace() / ShadowTrace() /
// Shade() keep running on the host, unmodified; what goes to the device is ONE virtual call at a time:
//   RtuObject    : Object   - IntersectRay(ray, hInfo) of a Sphere / Plane / TriObj = rtu_shadow_trace (the boolean for the
//                             HitInfo's current z, stale-z sphere return included) + rtu_trace (the record) on a scene that
//                             holds just that object
//   RtuMaterial  : Material - Shade(ray, hInfo, lights, bounceCount) = rtu_shade on the whole scene
// One device round trip per call: this is about the interface, not about speed (the frame-level path is RtuBeginRender).
namespace {

struct RtuOps {
    void *lib = nullptr;
    const char *(*last_error)(void) = nullptr;
    int (*trace)(rtu_scene *, const rtu_ray *, int64_t, rtu_hit *) = nullptr;
    int (*shadow_trace)(rtu_scene *, const rtu_ray *, const float *, int64_t, uint8_t *) = nullptr;
    int (*shade)(rtu_scene *, const rtu_ray *, const rtu_hit *, int64_t, int32_t, float *) = nullptr;
    rtu_context *ctx = nullptr;
    rtu_scene *whole = nullptr;
    Packed P;                               // the whole scene (kept alive: the description points into it)
    std::map<const Node *, int> node_index; // pre-order index = rtu_hit::node
};
RtuOps g_ops;

rtu_ray to_rtu(const Ray &r)
{
    rtu_ray o;
    copy3(o.p, r.p);
    copy3(o.dir, r.dir);
    return o;
}

class RtuObject : public Object
{
public:
    const Object *orig = nullptr;
    rtu_scene *one = nullptr;
    bool IntersectRay(const Ray &ray, HitInfo &hInfo, int hitSide = HIT_FRONT) const override
    {
        (void)hitSide;
        const rtu_ray r = to_rtu(ray);
        uint8_t occluded = 0;
        const float z_in = hInfo.z;
        if (g_ops.shadow_trace(one, &r, &z_in, 1, &occluded)) { fprintf(stderr, "rtu_shadow_trace: %s\n", g_ops.last_error()); exit(5); }
        if (!occluded) return false;
        rtu_hit h;
        if (g_ops.trace(one, &r, 1, &h)) { fprintf(stderr, "rtu_trace: %s\n", g_ops.last_error()); exit(5); }
        if (h.node >= 0 && h.z < hInfo.z) { // (a sphere may report a hit without a nearer z: objFunctions.cpp:60-75)
            hInfo.z = h.z;
            hInfo.p = Point3(h.p[0], h.p[1], h.p[2]);
            hInfo.N = Point3(h.N[0], h.N[1], h.N[2]);
            hInfo.uvw = Point3(h.uvw[0], h.uvw[1], h.uvw[2]);
            hInfo.front = h.front != 0;
        }
        return true;
    }
    Box GetBoundBox() const override { return orig->GetBoundBox(); }
};

class RtuMaterial : public Material
{
public:
    Color Shade(const Ray &ray, const HitInfo &hInfo, const LightList &lts, int bounceCount) const override
    {
        (void)lts; // the scene's own light list (what Render() passes at every call site of the Whitted estimator)
        const rtu_ray r = to_rtu(ray);
        rtu_hit h;
        memset(&h, 0, sizeof h);
        h.z = hInfo.z;
        copy3(h.p, hInfo.p); copy3(h.N, hInfo.N); copy3(h.uvw, hInfo.uvw);
        h.node = g_ops.node_index.at(hInfo.node);
        h.face = -1;
        h.front = hInfo.front ? 1 : 0;
        float rgb[3] = {0, 0, 0};
        if (g_ops.shade(g_ops.whole, &r, &h, 1, bounceCount, rgb)) { fprintf(stderr, "rtu_shade: %s\n", g_ops.last_error()); exit(5); }
        return Color(rgb[0], rgb[1], rgb[2]);
    }
};

void index_nodes(const Node *n, int &next)
{
    g_ops.node_index[n] = next++;
    for (int i = 0; i < n->GetNumChild(); i++) index_nodes(n->GetChild(i), next);
}

} // namespace

// what: bit 0 = every Node's object becomes an RtuObject, bit 1 = every Node's material becomes an RtuMaterial
int RtuInstallOperators(const char *lib_path, int what)
{
    RtuOps &G = g_ops;
    G.lib = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
    if (!G.lib) { fprintf(stderr, "cannot load %s: %s\n", lib_path, dlerror()); return 3; }
    G.last_error = sym<const char *(*)(void)>(G.lib, "rtu_last_error");
    G.trace = sym<int (*)(rtu_scene *, const rtu_ray *, int64_t, rtu_hit *)>(G.lib, "rtu_trace");
    G.shadow_trace = sym<int (*)(rtu_scene *, const rtu_ray *, const float *, int64_t, uint8_t *)>(G.lib, "rtu_shadow_trace");
    G.shade = sym<int (*)(rtu_scene *, const rtu_ray *, const rtu_hit *, int64_t, int32_t, float *)>(G.lib, "rtu_shade");
    auto ctx_create = sym<int (*)(int32_t, void *, rtu_context **)>(G.lib, "rtu_context_create");
    auto upload = sym<int (*)(rtu_context *, const rtu_scene_desc *, rtu_scene **)>(G.lib, "rtu_scene_upload");
    pack_scene(G.P); // before any object is replaced: the description is of the reference's own scene
    int next = 0;
    index_nodes(&rootNode, next);
    int rc = ctx_create(0, nullptr, &G.ctx);
    if (!rc) rc = upload(G.ctx, &G.P.desc, &G.whole);
    if (rc) { fprintf(stderr, "rtu: %s\n", G.last_error()); return rc; }
    std::map<const Object *, RtuObject *> proxies;
    std::vector<Node *> stack{&rootNode};
    while (!stack.empty()) {
        Node *n = stack.back();
        stack.pop_back();
        for (int i = 0; i < n->GetNumChild(); i++) stack.push_back(n->GetChild(i));
        if ((what & 1) && n->GetNodeObj()) {
            const Object *obj = n->GetNodeObj();
            RtuObject *&px = proxies[obj];
            if (!px) {
                // a scene of its own: the root (identity, no object) and one identity node holding the object
                static std::vector<std::vector<rtu_node>> keep;
                keep.emplace_back(2);
                rtu_node *nd = keep.back().data();
                for (int k = 0; k < 2; k++) {
                    memset(&nd[k], 0, sizeof(rtu_node));
                    nd[k].tm[0] = nd[k].tm[4] = nd[k].tm[8] = 1.f;
                    nd[k].itm[0] = nd[k].itm[4] = nd[k].itm[8] = 1.f;
                    nd[k].parent = k - 1;
                    nd[k].kind = RTU_OBJ_NONE;
                    nd[k].mesh = -1;
                    nd[k].material = -1;
                }
                rtu_scene_desc d;
                memset(&d, 0, sizeof d);
                d.camera = G.P.desc.camera;
                d.nodes = nd;
                d.n_nodes = 2;
                d.background.texmap = d.environment.texmap = -1;
                nd[1].kind = obj == &theSphere ? RTU_OBJ_SPHERE : obj == &thePlane ? RTU_OBJ_PLANE : RTU_OBJ_MESH;
                if (nd[1].kind == RTU_OBJ_MESH) {
                    int m = -1;
                    for (size_t k = 0; k < G.P.mesh_objs.size(); k++) if (G.P.mesh_objs[k] == obj) m = (int)k;
                    if (m < 0) { fprintf(stderr, "mesh not packed\n"); return 4; }
                    d.meshes = &G.P.meshes[m];
                    d.n_meshes = 1;
                    nd[1].mesh = 0;
                }
                px = new RtuObject;
                px->orig = obj;
                rc = upload(G.ctx, &d, &px->one);
                if (rc) { fprintf(stderr, "rtu_scene_upload: %s\n", G.last_error()); return rc; }
            }
            n->SetNodeObj(px);
        }
        if ((what & 2) && n->GetMaterial()) n->SetMaterial(new RtuMaterial);
    }
    return 0;
}
