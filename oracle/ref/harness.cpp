// TEST INFRASTRUCTURE (oracle/ref): headless driver around the UNMODIFIED reference.
//
// The reference has no headless entry point (main.cpp:74-88 opens a GLUT window) and its
// Render() (RenderFunctions.cpp:55-176) is hard-wired to 1024 spp path tracing, so the
// earlier-project behaviours that BASELINE.json names are reached by calling the
// reference's own functions directly:
//   primary : one Trace() per pixel centre            -> z, node id, face id, p, N, uvw, front
//   whitted : Trace() + Material::Shade(ray,h,lights,5) per sample, sample pattern of
//             RenderFunctions.cpp:77-97               -> linear RGB mean, RGB8 after gamma
//   head    : the reference Render() unchanged (threads joined, not busy-waited)
//   kat     : seeded random rays through Sphere/Plane/Box/BVHBox/TriObj::IntersectRay
//   photonkat: seeded photons through cyPhotonMap::AddPhoton / PrepareForIrradianceEstimation / EstimateIrradiance<100>
//   photon  : GeneratePhotonMap() + PhotonMapping() per pixel centre (RenderFunctions.cpp:341-413)
//   dump    : loader results (node transforms, camera, meshes, BVH, materials, lights)
// Everything numerical is computed by reference code; this file only drives it and writes
// .npy files.  Built by oracle/ref/Makefile into oracle/_ref/ref_harness (git-ignored).
#include "std_first.h"
#define private public
#define protected public
#include "ExternalLibrary/scene.h"
#include "ExternalLibrary/objects.h"
#include "ExternalLibrary/materials.h"
#include "ExternalLibrary/lights.h"
#include "ExternalLibrary/texture.h"
#include "ExternalLibrary/cyPhotonMap.h"
#undef private
#undef protected
extern RenderImage renderImage;
#include "PixelIterator.h"
#include "RenderFunctions.h"

extern Camera camera;
extern Sphere theSphere;
extern Plane thePlane;
extern Node rootNode;
extern MaterialList materials;
extern LightList lights;
extern ObjFileList objList;
extern TexturedColor background;
extern TexturedColor environment;
extern TextureList textureList;
extern thread_local unsigned long long g_traceCalls;
extern thread_local unsigned long long g_shadowCalls;

int LoadScene(const char *filename);
Point3 RefCalculateImageOrigin(float d);
Point3 RefCalculateCurrentPoint(int i, int j, float ox, float oy, Point3 o);
void RefRender(PixelIterator &it);
float BVHBoxIntersection(const Ray &r, Box bvhBox, float t_max);
void RefGeneratePhotonMap();
Color RefPhotonMapping(const Ray &r, const HitInfo &h);
cyPhotonMap *RefPhotonMap();
Color RefMonteCarloPhoton(const HitInfo &h, int x, int y, int n);
int RtuInstallOperators(const char *lib_path, int what);
int RtuBeginRender(const char *lib_path, int estimator, int spp, int bounces, int gi_bounces, bool reference_pattern, double *device_ms,
                   unsigned long long *rays);

// ---------------------------------------------------------------- npy output
static void WriteNpy(const std::string &path, const void *data, const char *descr,
                     size_t itemsize, const std::vector<size_t> &shape)
{
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    std::string sh = "(";
    size_t n = 1;
    for (size_t i = 0; i < shape.size(); i++) { sh += std::to_string(shape[i]) + ","; n *= shape[i]; }
    sh += ")";
    std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + sh + ", }";
    size_t total = 10 + hdr.size() + 1;
    size_t pad = (64 - total % 64) % 64;
    hdr += std::string(pad, ' ');
    hdr += "\n";
    unsigned char magic[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, 0, 0};
    magic[8] = (unsigned char)(hdr.size() & 0xff);
    magic[9] = (unsigned char)(hdr.size() >> 8);
    fwrite(magic, 1, 10, fp);
    fwrite(hdr.data(), 1, hdr.size(), fp);
    if (n) fwrite(data, itemsize, n, fp);
    fclose(fp);
}
static void NpyF(const std::string &p, const std::vector<float> &v, std::vector<size_t> shape) { WriteNpy(p, v.data(), "<f4", 4, shape); }
static void NpyI(const std::string &p, const std::vector<int> &v, std::vector<size_t> shape) { WriteNpy(p, v.data(), "<i4", 4, shape); }
static void NpyU8(const std::string &p, const std::vector<unsigned char> &v, std::vector<size_t> shape) { WriteNpy(p, v.data(), "|u1", 1, shape); }

// ---------------------------------------------------------------- scene indexing
struct NodeRec { Node *node; int parent; };
static std::vector<NodeRec> g_nodes;            // pre-order == the visiting order of Trace()
static std::map<const Node *, int> g_nodeIndex;
static std::vector<const TriObj *> g_meshes;    // unique meshes in first-visit order

static void IndexNodes(Node *n, int parent)
{
    int me = (int)g_nodes.size();
    g_nodes.push_back({n, parent});
    g_nodeIndex[n] = me;
    const Object *o = n->GetNodeObj();
    if (o && o != &theSphere && o != &thePlane) {
        const TriObj *t = (const TriObj *)o;
        if (std::find(g_meshes.begin(), g_meshes.end(), t) == g_meshes.end()) g_meshes.push_back(t);
    }
    for (int i = 0; i < n->GetNumChild(); i++) IndexNodes(n->GetChild(i), me);
}
static int ObjKind(const Object *o) { return !o ? 0 : (o == &theSphere ? 1 : (o == &thePlane ? 2 : 3)); }
static int MeshIndex(const Object *o)
{
    for (size_t i = 0; i < g_meshes.size(); i++) if ((const Object *)g_meshes[i] == o) return (int)i;
    return -1;
}

// Ray in the local coordinates of node idx: chain of ToNodeCoords from the root, as Trace does.
static Ray LocalRay(const Ray &world, int idx)
{
    std::vector<int> chain;
    for (int i = idx; i >= 0; i = g_nodes[i].parent) chain.push_back(i);
    Ray r = world;
    for (int k = (int)chain.size() - 1; k >= 0; k--) r = g_nodes[chain[k]].node->ToNodeCoords(r);
    return r;
}

// Face id of the winning triangle: replay TriObj::IntersectRay's traversal (objFunctions.cpp:
// 333-406) with the reference's own BVHBoxIntersection / IntersectTriangle and remember the
// last face that returned true (each accepted triangle strictly lowers hInfo.z).
static int WinningFace(const TriObj *t, const Ray &ray)
{
    HitInfo h;
    int face = -1;
    if (!t->GetBoundBox().IntersectRay(ray, BIGFLOAT)) return -1;
    std::vector<unsigned int> stack;
    stack.push_back(t->bvh.GetRootNodeID());
    while (!stack.empty()) {
        unsigned int n = stack.back();
        stack.pop_back();
        if (!t->bvh.IsLeafNode(n)) {
            unsigned int c1 = t->bvh.GetFirstChildNode(n), c2 = t->bvh.GetSecondChildNode(n);
            float t1 = BVHBoxIntersection(ray, Box(t->bvh.GetNodeBounds(c1)), BIGFLOAT);
            float t2 = BVHBoxIntersection(ray, Box(t->bvh.GetNodeBounds(c2)), BIGFLOAT);
            if (t1 <= t2) {
                if (t2 != BIGFLOAT) stack.push_back(c2);
                if (t1 != BIGFLOAT) stack.push_back(c1);
            } else if (t1 > t2) {
                if (t1 != BIGFLOAT) stack.push_back(c1);
                if (t2 != BIGFLOAT) stack.push_back(c2);
            }
        } else {
            for (unsigned int i = 0; i < t->bvh.GetNodeElementCount(n); i++) {
                unsigned int f = t->bvh.GetNodeElements(n)[i];
                if (t->IntersectTriangle(ray, h, HIT_FRONT, f)) face = (int)f;
            }
        }
    }
    return face;
}

// ---------------------------------------------------------------- camera rays
static Point3 g_imgOrigin;
// RenderFunctions.cpp:88-97 with the two rand()-driven lens numbers passed in (0,0 when dof==0)
static Ray CameraRay(int x, int y, float offX, float offY, float camOffsetX, float camOffsetY)
{
    Point3 sampledPosition = camera.pos + camera.up * camOffsetY +
                             camera.dir.GetNormalized().Cross(camera.up.GetNormalized()).GetNormalized() * camOffsetX;
    Point3 currentPoint = RefCalculateCurrentPoint(x, y, offX, offY, g_imgOrigin);
    return Ray(sampledPosition, (currentPoint - sampledPosition).GetNormalized());
}

// ---------------------------------------------------------------- options
struct Opts {
    std::string scene, root = ".", mode = "primary", out = "out", pattern = "center", lib = "raytracer-utah_b200/librtu_b200.so", estimator = "whitted", operators;
    int width = 0, height = 0, spp = 1, threads = 1, bounces = 5, n = 100000, seed = 1;
    int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    int s0 = 0, s1 = -1;   // --samples a b: only samples [a,b) of the spp-sample pattern (bounded CPU-baseline runs)
    bool quiet = true;
};

static double Now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class F> static void ParallelRows(int y0, int y1, int threads, F fn)
{
    std::atomic<int> next(y0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t]() {
            for (;;) {
                int y = next.fetch_add(1);
                if (y >= y1) break;
                fn(y, t);
            }
        });
    for (auto &t : th) t.join();
}

// ---------------------------------------------------------------- modes
static void ModePrimary(const Opts &o)
{
    int W = camera.imgWidth, H = camera.imgHeight;
    std::vector<float> z(W * H), p(W * H * 3), N(W * H * 3), uvw(W * H * 3);
    std::vector<int> node(W * H), face(W * H), front(W * H);
    std::vector<unsigned long long> tc(o.threads, 0);
    double t0 = Now();
    ParallelRows(0, H, o.threads, [&](int y, int tid) {
        for (int x = 0; x < W; x++) {
            Ray r = CameraRay(x, y, 0.5f, 0.5f, 0, 0);
            HitInfo h;
            bool hit = Trace(r, &rootNode, h);
            int i = x + W * y;
            z[i] = h.z;
            node[i] = hit ? g_nodeIndex[h.node] : -1;
            front[i] = hit ? (h.front ? 1 : 0) : -1;
            face[i] = -1;
            for (int k = 0; k < 3; k++) { p[i * 3 + k] = hit ? h.p[k] : 0; N[i * 3 + k] = hit ? h.N[k] : 0; uvw[i * 3 + k] = hit ? h.uvw[k] : 0; }
            if (hit && ObjKind(h.node->GetNodeObj()) == 3)
                face[i] = WinningFace((const TriObj *)h.node->GetNodeObj(), LocalRay(r, node[i]));
        }
        tc[tid] = g_traceCalls;
    });
    double dt = Now() - t0;
    size_t uW = W, uH = H;
    NpyF(o.out + "_z.npy", z, {uH, uW});
    NpyI(o.out + "_node.npy", node, {uH, uW});
    NpyI(o.out + "_face.npy", face, {uH, uW});
    NpyI(o.out + "_front.npy", front, {uH, uW});
    NpyF(o.out + "_p.npy", p, {uH, uW, 3});
    NpyF(o.out + "_N.npy", N, {uH, uW, 3});
    NpyF(o.out + "_uvw.npy", uvw, {uH, uW, 3});
    long hits = 0;
    for (int v : node) hits += v >= 0;
    fprintf(stderr, "{\"mode\":\"primary\",\"width\":%d,\"height\":%d,\"hits\":%ld,\"rays\":%ld,\"seconds\":%.6f,\"threads\":%d}\n",
            W, H, hits, (long)W * H, dt, o.threads);
}

static void ModeWhitted(const Opts &o)
{
    int W = camera.imgWidth, H = camera.imgHeight;
    int x0 = o.x0, y0 = o.y0, x1 = o.x1 < 0 ? W : o.x1, y1 = o.y1 < 0 ? H : o.y1;
    int cw = x1 - x0, ch = y1 - y0;
    std::vector<float> rgb((size_t)cw * ch * 3);
    std::vector<unsigned char> rgb8((size_t)cw * ch * 3);
    std::vector<unsigned long long> tc(o.threads, 0), sc(o.threads, 0);
    bool center = o.pattern == "center";
    int spp = o.spp;
    float pixelIncrement = 1.0 / spp;   // RenderFunctions.cpp:71
    srand((unsigned)o.seed);            // the stochastic branches (soft lights, glossy lobes, lens) draw from rand(); Render() seeds by wall clock (:60)
    double t0 = Now();
    ParallelRows(y0, y1, o.threads, [&](int y, int tid) {
        for (int x = x0; x < x1; x++) {
            Color sum(0.0, 0.0, 0.0);
            for (int s = o.s0; s < (o.s1 < 0 ? spp : o.s1); s++) {
                float ox, oy;
                if (center) { ox = 0.5f; oy = 0.5f; }
                else {
                    float cur = s * pixelIncrement;          // :81
                    ox = cur + Halton(s, 4);                   // :84,96
                    oy = cur + Halton(s, 5);                   // :85,96
                }
                float camOffsetX = 0, camOffsetY = 0;
                if (camera.dof > 0) {                          // the lens sample of Render(), :88-91 (two rand() draws)
                    float sampleX = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
                    float sampleTheta = static_cast<float>(rand()) / (static_cast<float>(RAND_MAX / (2 * M_PI)));
                    camOffsetX = sqrt(sampleX * camera.dof * camera.dof) * cos(sampleTheta);
                    camOffsetY = sqrt(sampleX * camera.dof * camera.dof) * sin(sampleTheta);
                }
                Ray r = CameraRay(x, y, ox, oy, camOffsetX, camOffsetY);
                HitInfo h;
                Color c(0.0, 0.0, 0.0);
                if (Trace(r, &rootNode, h)) {
                    const Material *m = h.node->GetMaterial();
                    c = m ? m->Shade(r, h, lights, o.bounces) : Color(1, 1, 1);
                } else {
                    c = background.Sample(Point3((float)x / camera.imgWidth, (float)y / camera.imgHeight, 0));   // :145
                }
                sum += c;
            }
            sum /= (float)spp;                                 // :152
            size_t i = ((size_t)(x - x0) + (size_t)cw * (y - y0)) * 3;
            rgb[i] = sum.r; rgb[i + 1] = sum.g; rgb[i + 2] = sum.b;
            Color g = sum;
            g.r = pow(g.r, 1 / 2.2); g.g = pow(g.g, 1 / 2.2); g.b = pow(g.b, 1 / 2.2);   // :155-157
            Color24 q(g);
            rgb8[i] = q.r; rgb8[i + 1] = q.g; rgb8[i + 2] = q.b;
        }
        tc[tid] = g_traceCalls;
        sc[tid] = g_shadowCalls;
    });
    double dt = Now() - t0;
    unsigned long long T = 0, S = 0;
    for (auto v : tc) T += v;
    for (auto v : sc) S += v;
    if (o.out != "-") {
        NpyF(o.out + "_rgb.npy", rgb, {(size_t)ch, (size_t)cw, 3});
        NpyU8(o.out + "_rgb8.npy", rgb8, {(size_t)ch, (size_t)cw, 3});
    }
    fprintf(stderr, "{\"mode\":\"whitted\",\"width\":%d,\"height\":%d,\"crop\":[%d,%d,%d,%d],\"spp\":%d,\"pattern\":\"%s\",\"trace_rays\":%llu,\"shadow_rays\":%llu,\"rays\":%llu,\"seconds\":%.6f,\"threads\":%d,\"mrays_per_s\":%.4f}\n",
            W, H, x0, y0, x1, y1, spp, o.pattern.c_str(), T, S, T + S, dt, o.threads, (T + S) / dt * 1e-6);
}

static void ModeHead(const Opts &o)
{
    // The reference Render() unchanged: 1024 spp + 4-bounce GI (RenderFunctions.cpp:26-36).
    // main.cpp:50-54 detaches the threads and busy-waits on the ticket counter; joining is the
    // only change.  The PixelIterator off-by-one (A-1) writes one
    // Color24 past the image, so the image is re-allocated with one spare row first.
    int W = camera.imgWidth, H = camera.imgHeight;
    renderImage.Init(W, H + 1);
    renderImage.width = W; renderImage.height = H;
    PixelIterator it;
    double t0 = Now();
    std::vector<std::thread> th;
    for (int t = 0; t < o.threads; t++) th.emplace_back([&]() { RefRender(it); });
    for (auto &t : th) t.join();
    double dt = Now() - t0;
    std::vector<unsigned char> rgb8((size_t)W * H * 3);
    memcpy(rgb8.data(), renderImage.GetPixels(), rgb8.size());
    NpyU8(o.out + "_rgb8.npy", rgb8, {(size_t)H, (size_t)W, 3});
    fprintf(stderr, "{\"mode\":\"head\",\"width\":%d,\"height\":%d,\"seconds\":%.3f,\"threads\":%d}\n", W, H, dt, o.threads);
}

// xorshift64* : only used to draw test rays, not part of any compared arithmetic
static unsigned long long g_rng;
static float Rnd() { g_rng ^= g_rng >> 12; g_rng ^= g_rng << 25; g_rng ^= g_rng >> 27; return (float)((g_rng * 2685821657736338717ULL) >> 40) / 16777216.0f; }
static float RndS(float s) { return (Rnd() * 2 - 1) * s; }

static void DumpHits(const std::string &pre, int n, const std::vector<float> &rays, const std::vector<float> &zin,
                     const std::vector<int> &hit, const std::vector<HitInfo> &h, const std::vector<int> *face)
{
    std::vector<float> z(n), p(n * 3), N(n * 3), uvw(n * 3);
    std::vector<int> front(n);
    for (int i = 0; i < n; i++) {
        z[i] = h[i].z; front[i] = h[i].front;
        for (int k = 0; k < 3; k++) { p[i * 3 + k] = hit[i] ? h[i].p[k] : 0; N[i * 3 + k] = hit[i] ? h[i].N[k] : 0; uvw[i * 3 + k] = hit[i] ? h[i].uvw[k] : 0; }
    }
    size_t un = n;
    NpyF(pre + "_rays.npy", rays, {un, 6});
    NpyF(pre + "_zin.npy", zin, {un});
    NpyI(pre + "_hit.npy", hit, {un});
    NpyF(pre + "_z.npy", z, {un});
    NpyI(pre + "_front.npy", front, {un});
    NpyF(pre + "_p.npy", p, {un, 3});
    NpyF(pre + "_N.npy", N, {un, 3});
    NpyF(pre + "_uvw.npy", uvw, {un, 3});
    if (face) NpyI(pre + "_face.npy", *face, {un});
}

static void ModeKat(const Opts &o)
{
    int n = o.n;
    Node idn;   // identity transform
    g_rng = 0x9E3779B97F4A7C15ULL ^ (unsigned long long)o.seed;
    // ---- sphere / plane: rays aimed near the unit primitive, some with axis-aligned or zero components
    for (int kind = 1; kind <= 2; kind++) {
        std::vector<float> rays(n * 6), zin(n);
        std::vector<int> hit(n);
        std::vector<HitInfo> hs(n);
        for (int i = 0; i < n; i++) {
            Point3 p(RndS(3), RndS(3), RndS(3));
            if (i % 7 == 3) p *= 0.3f;                       // origins inside the sphere
            Point3 tgt(RndS(1.2f), RndS(1.2f), kind == 2 ? 0.f : RndS(1.2f));
            Point3 d = tgt - p;
            if (i % 5 == 0) d.Normalize();
            if (i % 11 == 0) d.x = 0;
            if (i % 13 == 0) d.y = 0;
            if (i % 17 == 0) d.z = 0;
            if (i % 19 == 0) d *= 1 + Rnd() * 30;           // unnormalised, as ToNodeCoords produces
            Ray r(p, d);
            HitInfo h;
            if (i % 3 == 0) h.z = Rnd() * 4;                 // a nearer hit already found / shadow t_max
            zin[i] = h.z;
            // through two identity nodes (rootNode + object node), exactly as Trace() would reach the object
            Ray lr = idn.ToNodeCoords(idn.ToNodeCoords(r));
            bool res = kind == 1 ? theSphere.IntersectRay(lr, h) : thePlane.IntersectRay(lr, h);
            if (res) { idn.FromNodeCoords(h); idn.FromNodeCoords(h); }
            hit[i] = res; hs[i] = h;
            for (int k = 0; k < 3; k++) { rays[i * 6 + k] = p[k]; rays[i * 6 + 3 + k] = d[k]; }
        }
        DumpHits(o.out + (kind == 1 ? "_sphere" : "_plane"), n, rays, zin, hit, hs, nullptr);
    }
    // ---- Box::IntersectRay and BVHBoxIntersection on random boxes
    {
        std::vector<float> rays(n * 6), boxes(n * 6), tmax(n), tb(n);
        std::vector<int> hit(n);
        for (int i = 0; i < n; i++) {
            Point3 a(RndS(2), RndS(2), RndS(2)), b(RndS(2), RndS(2), RndS(2));
            Box bx(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), max(a.x, b.x), max(a.y, b.y), max(a.z, b.z));
            if (i % 23 == 0) bx.pmax.z = bx.pmin.z;          // flat box (the Plane's bound box)
            Point3 p(RndS(4), RndS(4), RndS(4));
            Point3 tgt(RndS(2.5f), RndS(2.5f), RndS(2.5f));
            Point3 d = tgt - p;
            if (i % 11 == 0) d.x = 0;
            if (i % 13 == 0) d.y = 0;
            if (i % 17 == 0) d.z = 0;
            if (i % 29 == 0) p.x = bx.pmin.x;                // origin on a slab plane with d.x==0 -> 0/0
            Ray r(p, d);
            float tm = (i % 4 == 0) ? Rnd() * 6 : BIGFLOAT;
            hit[i] = bx.IntersectRay(r, tm);
            tb[i] = BVHBoxIntersection(r, bx, tm);
            tmax[i] = tm;
            for (int k = 0; k < 3; k++) { rays[i * 6 + k] = p[k]; rays[i * 6 + 3 + k] = d[k]; boxes[i * 6 + k] = bx.pmin[k]; boxes[i * 6 + 3 + k] = bx.pmax[k]; }
        }
        size_t un = n;
        NpyF(o.out + "_box_rays.npy", rays, {un, 6});
        NpyF(o.out + "_box_boxes.npy", boxes, {un, 6});
        NpyF(o.out + "_box_tmax.npy", tmax, {un});
        NpyI(o.out + "_box_hit.npy", hit, {un});
        NpyF(o.out + "_box_tbvh.npy", tb, {un});
    }
    // ---- TriObj::IntersectRay on the first mesh of the loaded scene (rays in mesh-local space)
    if (!g_meshes.empty()) {
        const TriObj *t = g_meshes[0];
        Point3 c = (t->GetBoundMin() + t->GetBoundMax()) * 0.5f;
        Point3 e = (t->GetBoundMax() - t->GetBoundMin()) * 0.5f;
        std::vector<float> rays(n * 6), zin(n);
        std::vector<int> hit(n), face(n);
        std::vector<HitInfo> hs(n);
        for (int i = 0; i < n; i++) {
            Point3 p = c + Point3(RndS(3) * e.x, RndS(3) * e.y, RndS(3) * e.z);
            Point3 tgt = c + Point3(RndS(1) * e.x, RndS(1) * e.y, RndS(1) * e.z);
            Point3 d = tgt - p;
            if (i % 5 == 0) d.Normalize();
            if (i % 31 == 0) d.x = 0;
            if (i % 37 == 0) d.z = 0;
            Ray r(p, d);
            HitInfo h;
            if (i % 3 == 0) h.z = Rnd() * 2;
            zin[i] = h.z;
            Ray lr = idn.ToNodeCoords(idn.ToNodeCoords(r));
            hit[i] = t->IntersectRay(lr, h);
            if (hit[i]) { idn.FromNodeCoords(h); idn.FromNodeCoords(h); }
            hs[i] = h;
            // WinningFace starts from z=BIGFLOAT; the winner is the same whenever the mesh hit at all,
            // because the closest triangle also passes the tighter initial-z gate
            face[i] = hit[i] ? WinningFace(t, lr) : -1;
            for (int k = 0; k < 3; k++) { rays[i * 6 + k] = p[k]; rays[i * 6 + 3 + k] = d[k]; }
        }
        DumpHits(o.out + "_mesh", n, rays, zin, hit, hs, &face);
    }
    fprintf(stderr, "{\"mode\":\"kat\",\"n\":%d}\n", n);
}

static void PushM(std::vector<float> &v, const Matrix3 &m) { for (int i = 0; i < 9; i++) v.push_back(m.data[i]); }
static void PushP(std::vector<float> &v, const Point3 &p) { v.push_back(p.x); v.push_back(p.y); v.push_back(p.z); }
static void PushC(std::vector<float> &v, const Color &c) { v.push_back(c.r); v.push_back(c.g); v.push_back(c.b); }

static void ModeDump(const Opts &o)
{
    size_t nn = g_nodes.size();
    std::vector<float> tm, itm, pos;
    std::vector<int> meta;   // parent, kind, mesh, material index
    for (size_t i = 0; i < nn; i++) {
        Node *n = g_nodes[i].node;
        PushM(tm, n->GetTransform()); PushM(itm, n->GetInverseTransform()); PushP(pos, n->GetPosition());
        int mi = -1;
        for (size_t k = 0; k < materials.size(); k++) if (materials[k] == n->GetMaterial()) mi = (int)k;
        meta.push_back(g_nodes[i].parent); meta.push_back(ObjKind(n->GetNodeObj())); meta.push_back(MeshIndex(n->GetNodeObj())); meta.push_back(mi);
    }
    NpyF(o.out + "_node_tm.npy", tm, {nn, 9});
    NpyF(o.out + "_node_itm.npy", itm, {nn, 9});
    NpyF(o.out + "_node_pos.npy", pos, {nn, 3});
    NpyI(o.out + "_node_meta.npy", meta, {nn, 4});
    std::vector<float> cam;
    PushP(cam, camera.pos); PushP(cam, camera.dir); PushP(cam, camera.up);
    cam.push_back(camera.fov); cam.push_back(camera.focaldist); cam.push_back(camera.dof);
    cam.push_back((float)camera.imgWidth); cam.push_back((float)camera.imgHeight);
    PushP(cam, g_imgOrigin);
    NpyF(o.out + "_camera.npy", cam, {cam.size()});
    for (size_t m = 0; m < g_meshes.size(); m++) {
        const TriObj *t = g_meshes[m];
        std::string pre = o.out + "_mesh" + std::to_string(m);
        size_t nv = t->NV(), nf = t->NF(), nvn = t->NVN(), nvt = t->NVT();
        WriteNpy(pre + "_v.npy", t->v, "<f4", 4, {nv, 3});
        WriteNpy(pre + "_f.npy", t->f, "<u4", 4, {nf, 3});
        WriteNpy(pre + "_vn.npy", t->vn, "<f4", 4, {nvn, 3});
        WriteNpy(pre + "_fn.npy", t->fn, "<u4", 4, {t->fn ? nf : 0, 3});
        WriteNpy(pre + "_vt.npy", t->vt, "<f4", 4, {nvt, 3});
        WriteNpy(pre + "_ft.npy", t->ft, "<u4", 4, {t->ft ? nf : 0, 3});
        // BVH: count nodes by walking from the root
        unsigned int maxNode = 1;
        std::vector<unsigned int> st{1};
        while (!st.empty()) {
            unsigned int n = st.back(); st.pop_back();
            maxNode = max(maxNode, n);
            if (!t->bvh.IsLeafNode(n)) { st.push_back(t->bvh.GetFirstChildNode(n)); st.push_back(t->bvh.GetSecondChildNode(n)); }
        }
        std::vector<float> boxes((maxNode + 1) * 6, 0.f);
        std::vector<unsigned int> data(maxNode + 1, 0);
        for (unsigned int n = 1; n <= maxNode; n++) {
            const float *b = t->bvh.GetNodeBounds(n);
            for (int k = 0; k < 6; k++) boxes[n * 6 + k] = b[k];
            data[n] = t->bvh.nodes[n].data;
        }
        NpyF(pre + "_bvh_boxes.npy", boxes, {(size_t)maxNode + 1, 6});
        WriteNpy(pre + "_bvh_data.npy", data.data(), "<u4", 4, {(size_t)maxNode + 1});
        WriteNpy(pre + "_bvh_elements.npy", t->bvh.elements, "<u4", 4, {nf});
        std::vector<float> bb; PushP(bb, t->GetBoundMin()); PushP(bb, t->GetBoundMax());
        NpyF(pre + "_bound.npy", bb, {6});
    }
    // materials (MtlBlinn only; MultiMtl -> sub-material 0)
    std::vector<float> mt;
    for (size_t k = 0; k < materials.size(); k++) {
        const MtlBlinn *b = dynamic_cast<const MtlBlinn *>(materials[k]);
        if (!b) { const MultiMtl *mm = dynamic_cast<const MultiMtl *>(materials[k]); b = mm && !mm->mtls.empty() ? dynamic_cast<const MtlBlinn *>(mm->mtls[0]) : nullptr; }
        if (!b) { for (int i = 0; i < 22; i++) mt.push_back(0); continue; }
        PushC(mt, b->diffuse.GetColor()); PushC(mt, b->specular.GetColor()); PushC(mt, b->reflection.GetColor());
        PushC(mt, b->refraction.GetColor()); PushC(mt, b->emission.GetColor()); PushC(mt, b->absorption);
        mt.push_back(b->glossiness); mt.push_back(b->ior); mt.push_back(b->reflectionGlossiness); mt.push_back(b->refractionGlossiness);
    }
    NpyF(o.out + "_materials.npy", mt, {materials.size(), 22});
    std::vector<float> lt;
    for (size_t k = 0; k < lights.size(); k++) {
        const Light *l = lights[k];
        float kind = 0; Color I(0, 0, 0); Point3 v(0, 0, 0); float size = 0;
        if (const AmbientLight *a = dynamic_cast<const AmbientLight *>(l)) { kind = 0; I = a->intensity; }
        else if (const DirectLight *d = dynamic_cast<const DirectLight *>(l)) { kind = 1; I = d->intensity; v = d->direction; }
        else if (const PointLight *p = dynamic_cast<const PointLight *>(l)) { kind = 2; I = p->intensity; v = p->position; size = p->size; }
        lt.push_back(kind); PushC(lt, I); PushP(lt, v); lt.push_back(size);
    }
    NpyF(o.out + "_lights.npy", lt, {lights.size(), 8});
    std::vector<float> bg; PushC(bg, background.GetColor()); PushC(bg, environment.GetColor());
    bg.push_back(background.GetTexture() ? 1.f : 0.f); bg.push_back(environment.GetTexture() ? 1.f : 0.f);
    NpyF(o.out + "_bgenv.npy", bg, {bg.size()});
    fprintf(stderr, "{\"mode\":\"dump\",\"nodes\":%zu,\"meshes\":%zu,\"materials\":%zu,\"lights\":%zu}\n", nn, g_meshes.size(), materials.size(), lights.size());
}

// Texture / environment sampling KAT: TexturedColor::Sample and SampleEnvironment on seeded inputs
static void ModeTex(const Opts &o)
{
    int n = o.n;
    g_rng = 0xD1B54A32D192ED03ULL ^ (unsigned long long)o.seed;
    std::vector<float> uvw(n * 3), dirs(n * 3), bgc(n * 3), envc(n * 3);
    std::vector<std::vector<float>> mats(materials.size(), std::vector<float>(n * 12));
    for (int i = 0; i < n; i++) {
        Point3 u(RndS(3), RndS(3), RndS(1)), d(RndS(1), RndS(1), RndS(1));
        d.Normalize();
        Color b = background.Sample(u), e = environment.SampleEnvironment(d);
        for (int k = 0; k < 3; k++) { uvw[i * 3 + k] = u[k]; dirs[i * 3 + k] = d[k]; bgc[i * 3 + k] = b[k]; envc[i * 3 + k] = e[k]; }
        for (size_t m = 0; m < materials.size(); m++) {
            const MtlBlinn *mb = dynamic_cast<const MtlBlinn *>(materials[m]);
            if (!mb) continue;
            Color c4[4] = {mb->diffuse.Sample(u), mb->specular.Sample(u), mb->reflection.Sample(u), mb->refraction.Sample(u)};
            for (int q = 0; q < 4; q++) for (int k = 0; k < 3; k++) mats[m][i * 12 + q * 3 + k] = c4[q][k];
        }
    }
    size_t un = n;
    NpyF(o.out + "_tex_uvw.npy", uvw, {un, 3});
    NpyF(o.out + "_tex_dirs.npy", dirs, {un, 3});
    NpyF(o.out + "_tex_background.npy", bgc, {un, 3});
    NpyF(o.out + "_tex_environment.npy", envc, {un, 3});
    for (size_t m = 0; m < materials.size(); m++) NpyF(o.out + "_tex_mtl" + std::to_string(m) + ".npy", mats[m], {un, 4, 3});
    fprintf(stderr, "{\"mode\":\"tex\",\"n\":%d}\n", n);
}

// The reference binary with its render loop replaced by librtu_b200.so (oracle/ref/rtu_binding.cpp):
// LoadScene() and the PNG writers are the reference's own.
static void ModeGpu(const Opts &o)
{
    double ms = 0;
    unsigned long long rays = 0;
    // --estimator head (default of the binding): what Render() computes at HEAD; whitted: Shade(ray,h,lights,5) alone
    const bool head = o.estimator == "head";
    int rc = RtuBeginRender(o.lib.c_str(), head ? 2 /* RTU_MODE_PATH */ : 1 /* RTU_MODE_WHITTED */, o.spp, o.bounces, 4, head || o.pattern != "center", &ms, &rays);
    if (rc) exit(rc);
    if (renderImage.GetNumRenderedPixels() != camera.imgWidth * camera.imgHeight) { fprintf(stderr, "progress counter ended at %d\n", renderImage.GetNumRenderedPixels()); exit(4); }
    int W = camera.imgWidth, H = camera.imgHeight;
    renderImage.SaveImage((o.out + "_Result.png").c_str());        // main.cpp:59
    renderImage.ComputeZBufferImage();                             // main.cpp:60
    renderImage.SaveZImage((o.out + "_ZBuffer.png").c_str());      // main.cpp:61
    std::vector<unsigned char> rgb8((size_t)W * H * 3), z8((size_t)W * H);
    memcpy(rgb8.data(), renderImage.GetPixels(), rgb8.size());
    memcpy(z8.data(), renderImage.GetZBufferImage(), z8.size());
    NpyU8(o.out + "_rgb8.npy", rgb8, {(size_t)H, (size_t)W, 3});
    NpyU8(o.out + "_z8.npy", z8, {(size_t)H, (size_t)W});
    fprintf(stderr, "{\"mode\":\"gpu\",\"width\":%d,\"height\":%d,\"spp\":%d,\"rays\":%llu,\"device_ms\":%.3f}\n", W, H, o.spp, rays, ms);
}


// ---------------------------------------------------------------- photon map
static_assert(sizeof(cyPhotonMap::Photon) == 24, "photon record is 24 bytes");
static void DumpPhotons(const std::string &path, const cyPhotonMap &m, int first, int count)
{
    std::vector<unsigned char> raw((size_t)count * 24);
    if (count) memcpy(raw.data(), &m.photons[first], raw.size());
    NpyU8(path, raw, {(size_t)count, 24});
}

static void ModePhotonKat(const Opts &o)
{
    // photons on three surfaces (a floor, a wall, a sphere) so that the normal / ellipticity filters matter
    g_rng = 0x9E3779B97F4A7C15ULL ^ (unsigned long long)o.seed;
    const int n = o.n;
    cyPhotonMap m;
    m.Resize(n);
    for (int i = 0; i < n; i++) {
        Point3 pos, nrm;
        int s = i % 3;
        if (s == 0) { pos = Point3(RndS(10), RndS(10), 0.01f * RndS(1)); nrm = Point3(0, 0, 1); }
        else if (s == 1) { pos = Point3(-10 + 0.01f * RndS(1), RndS(10), 10 * Rnd()); nrm = Point3(1, 0, 0); }
        else { Point3 d(RndS(1), RndS(1), RndS(1)); d.Normalize(); pos = Point3(2, 1, 3) + d * 2.5f; nrm = d; }
        Point3 dir(RndS(1), RndS(1), RndS(1));
        dir.Normalize();
        if (Rnd() < 0.8f && dir % nrm > 0) dir = -dir;   // most photons arrive from the front side
        Color pw(Rnd() * 2, Rnd() * 1.5f, Rnd());
        m.AddPhoton(pos, dir, pw);
    }
    m.ScalePhotonPowers(0.37f);
    DumpPhotons(o.out + "_photons_in.npy", m, 1, m.NumPhotons());
    m.PrepareForIrradianceEstimation();
    DumpPhotons(o.out + "_photons_balanced.npy", m, 1, m.NumPhotons());
    const int q = 4096;
    std::vector<float> qp(q * 3), qn(q * 3), irr(q * 3 * 4), dir(q * 3 * 4);
    const float radius[4] = {1.0f, 1.0f, 0.25f, 3.0f}, ell[4] = {0.5f, 1.0f, 0.5f, 0.2f};
    for (int i = 0; i < q; i++) {
        Point3 pos, nrm;
        int s = i % 4;
        if (s == 0) { pos = Point3(RndS(10), RndS(10), 0); nrm = Point3(0, 0, 1); }
        else if (s == 1) { pos = Point3(-10, RndS(10), 10 * Rnd()); nrm = Point3(1, 0, 0); }
        else if (s == 2) { Point3 d(RndS(1), RndS(1), RndS(1)); d.Normalize(); pos = Point3(2, 1, 3) + d * 2.5f; nrm = d; }
        else { pos = Point3(RndS(12), RndS(12), RndS(12)); nrm = Point3(RndS(1), RndS(1), RndS(1)); nrm.Normalize(); }
        for (int k = 0; k < 3; k++) { qp[i * 3 + k] = pos[k]; qn[i * 3 + k] = nrm[k]; }
        for (int v = 0; v < 4; v++) {
            Color c;
            Point3 d;
            m.EstimateIrradiance<100>(c, d, radius[v], pos, &nrm, ell[v]);
            irr[(v * q + i) * 3 + 0] = c.r; irr[(v * q + i) * 3 + 1] = c.g; irr[(v * q + i) * 3 + 2] = c.b;
            for (int k = 0; k < 3; k++) dir[(v * q + i) * 3 + k] = d[k];
        }
    }
    NpyF(o.out + "_qpos.npy", qp, {(size_t)q, 3});
    NpyF(o.out + "_qnormal.npy", qn, {(size_t)q, 3});
    NpyF(o.out + "_irrad.npy", irr, {4, (size_t)q, 3});
    NpyF(o.out + "_dir.npy", dir, {4, (size_t)q, 3});
    fprintf(stderr, "{\"mode\":\"photonkat\",\"photons\":%d,\"queries\":%d,\"radius\":[1.0,1.0,0.25,3.0],\"ellipticity\":[0.5,1.0,0.5,0.2],\"max_photons\":100}\n", n, q);
}

static void ModePhoton(const Opts &o)
{
    int W = camera.imgWidth, H = camera.imgHeight;
    srand((unsigned)o.seed);
    g_traceCalls = 0;
    double t0 = Now();
    fflush(stdout);
    int saved = dup(1);
    FILE *cap = tmpfile();
    dup2(fileno(cap), 1);
    RefGeneratePhotonMap();       // prints "Photon From Light: %i" / "Photon Scale Factor: %f" (RenderFunctions.cpp:386-387)
    fflush(stdout);
    dup2(saved, 1);
    double tGen = Now() - t0;
    unsigned long long emitTraces = g_traceCalls;
    rewind(cap);
    int fromLight = 0;
    float scale = 0;
    char line[256];
    while (fgets(line, sizeof line, cap)) {
        sscanf(line, "Photon From Light: %i", &fromLight);
        sscanf(line, "Photon Scale Factor: %f", &scale);
    }
    fclose(cap);
    cyPhotonMap *m = RefPhotonMap();
    DumpPhotons(o.out + "_photons_balanced.npy", *m, 1, m->NumPhotons());
    std::vector<float> rgb((size_t)W * H * 3), irr((size_t)W * H * 3), dir((size_t)W * H * 3);
    std::vector<int> node((size_t)W * H);
    t0 = Now();
    ParallelRows(0, H, o.threads, [&](int y, int) {
        for (int x = 0; x < W; x++) {
            Ray r = CameraRay(x, y, 0.5f, 0.5f, 0, 0);
            HitInfo h;
            size_t i = x + (size_t)W * y;
            Color c = background.Sample(Point3((float)x / camera.imgWidth, (float)y / camera.imgHeight, 0));
            Color e(0, 0, 0);
            Point3 d(0, 0, 0);
            node[i] = -1;
            if (Trace(r, &rootNode, h)) {
                node[i] = g_nodeIndex[h.node];
                m->EstimateIrradiance<100>(e, d, 1.0f, h.p, &h.N, 0.5f);   // what PhotonMapping computes first (RenderFunctions.cpp:401)
                c = RefPhotonMapping(r, h);
            }
            rgb[i * 3] = c.r; rgb[i * 3 + 1] = c.g; rgb[i * 3 + 2] = c.b;
            irr[i * 3] = e.r; irr[i * 3 + 1] = e.g; irr[i * 3 + 2] = e.b;
            for (int k = 0; k < 3; k++) dir[i * 3 + k] = d[k];
        }
    });
    double tGather = Now() - t0;
    if (o.spp > 1) {
        // the commented-out "Photon Map + MonteCarlo" estimator of Render() (RenderFunctions.cpp:137-139):
        // Shade(ray, h, lights, 5) + MonteCarloPhoton(h, x, y, 1), averaged over --spp evaluations at the pixel centre
        std::vector<float> gi((size_t)W * H * 3);
        ParallelRows(0, H, 1, [&](int y, int) {   // one thread: MonteCarloPhoton draws from rand()
            for (int x = 0; x < W; x++) {
                Ray r = CameraRay(x, y, 0.5f, 0.5f, 0, 0);
                size_t i = x + (size_t)W * y;
                Color sum(0, 0, 0);
                for (int s = 0; s < o.spp; s++) {
                    HitInfo h;
                    if (Trace(r, &rootNode, h)) {
                        sum += h.node->GetMaterial()->Shade(r, h, lights, 5);
                        sum += RefMonteCarloPhoton(h, x, y, 1);
                    } else {
                        sum += background.Sample(Point3((float)x / camera.imgWidth, (float)y / camera.imgHeight, 0));
                    }
                }
                sum /= (float)o.spp;
                gi[i * 3] = sum.r; gi[i * 3 + 1] = sum.g; gi[i * 3 + 2] = sum.b;
            }
        });
        NpyF(o.out + "_gi.npy", gi, {(size_t)H, (size_t)W, 3});
    }
    NpyF(o.out + "_rgb.npy", rgb, {(size_t)H, (size_t)W, 3});
    NpyF(o.out + "_irrad.npy", irr, {(size_t)H, (size_t)W, 3});
    NpyF(o.out + "_dir.npy", dir, {(size_t)H, (size_t)W, 3});
    NpyI(o.out + "_node.npy", node, {(size_t)H, (size_t)W});
    fprintf(stderr, "{\"mode\":\"photon\",\"width\":%d,\"height\":%d,\"photons\":%d,\"from_light\":%d,\"scale\":%.9g,\"emit_traces\":%llu,\"emit_seconds\":%.3f,\"gather_seconds\":%.3f,\"threads\":%d}\n",
            W, H, m->NumPhotons(), fromLight, scale, emitTraces, tGen, tGather, o.threads);
}

int main(int argc, char **argv)
{
    Opts o;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--root") o.root = next();
        else if (a == "--mode") o.mode = next();
        else if (a == "--out") o.out = next();
        else if (a == "--pattern") o.pattern = next();
        else if (a == "--width") o.width = atoi(next());
        else if (a == "--height") o.height = atoi(next());
        else if (a == "--spp") o.spp = atoi(next());
        else if (a == "--threads") o.threads = atoi(next());
        else if (a == "--bounces") o.bounces = atoi(next());
        else if (a == "--n") o.n = atoi(next());
        else if (a == "--seed") o.seed = atoi(next());
        else if (a == "--samples") { o.s0 = atoi(next()); o.s1 = atoi(next()); }
        else if (a == "--crop") { o.x0 = atoi(next()); o.y0 = atoi(next()); o.x1 = atoi(next()); o.y1 = atoi(next()); }
        else if (a == "--verbose") o.quiet = false;
        else if (a == "--lib") o.lib = next();
        else if (a == "--estimator") o.estimator = next();
        else if (a == "--operators") o.operators = next();   // objects | materials | both: the reference's own code on device-backed operators
        else if (a[0] != '-') o.scene = a;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (o.scene.empty()) {
        fprintf(stderr, "usage: ref_harness <scene.xml> --root <asset root> --mode primary|whitted|head|kat|tex|dump [--width W --height H --spp N --pattern center|ref --threads T --out prefix]\n");
        return 2;
    }
    if (o.threads <= 0) { o.threads = std::thread::hardware_concurrency(); if (o.threads == 0) o.threads = 1; }   // main.cpp:39-43
    // absolute paths before chdir
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return 2;
    auto absolutize = [&](const std::string &p) { return (p.empty() || p[0] == '/' || p == "-") ? p : std::string(cwd) + "/" + p; };
    o.scene = absolutize(o.scene);
    o.out = absolutize(o.out);
    o.lib = absolutize(o.lib);
    if (chdir(o.root.c_str()) != 0) { fprintf(stderr, "cannot chdir to %s\n", o.root.c_str()); return 2; }
    int savedStdout = -1;
    if (o.quiet) { fflush(stdout); savedStdout = dup(1); FILE *nul = fopen("/dev/null", "w"); dup2(fileno(nul), 1); }
    int ok = LoadScene(o.scene.c_str());
    if (o.quiet) { fflush(stdout); dup2(savedStdout, 1); }
    if (!ok) { fprintf(stderr, "LoadScene failed for %s\n", o.scene.c_str()); return 1; }
    if (o.width > 0) camera.imgWidth = o.width;
    if (o.height > 0) camera.imgHeight = o.height;
    renderImage.Init(camera.imgWidth, camera.imgHeight);
    IndexNodes(&rootNode, -1);
    g_imgOrigin = RefCalculateImageOrigin(camera.focaldist);   // also sets actualWidth/actualHeight
    if (!o.operators.empty()) {
        // after IndexNodes (the harness's own node table is by Node*, which the replacement keeps)
        int rc = RtuInstallOperators(o.lib.c_str(), o.operators == "objects" ? 1 : o.operators == "materials" ? 2 : 3);
        if (rc) return rc;
        o.threads = 1; // one context, one host thread
    }
    if (o.mode == "primary") ModePrimary(o);
    else if (o.mode == "whitted") ModeWhitted(o);
    else if (o.mode == "head") ModeHead(o);
    else if (o.mode == "kat") ModeKat(o);
    else if (o.mode == "tex") ModeTex(o);
    else if (o.mode == "dump") ModeDump(o);
    else if (o.mode == "gpu") ModeGpu(o);
    else if (o.mode == "photonkat") ModePhotonKat(o);
    else if (o.mode == "photon") ModePhoton(o);
    else { fprintf(stderr, "unknown mode %s\n", o.mode.c_str()); return 2; }
    return 0;
}
