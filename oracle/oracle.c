/*
 * oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar CPU restatement, in plain C, of the reference's render path
 *   Render -> Trace/ShadowTrace -> Object::IntersectRay -> Material::Shade -> Light::Illuminate
 * operating on the flattened scene description of include/rtu.h.  It exists so that the GPU
 * box (which has no /root/reference) can check the CUDA path on arbitrary seeded inputs, and as
 * the "port" CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product (raytracer-utah_b200/) never does.
 *
 * PINNED: tests/test_oracle.py checks every function here against tests/golden/*.npz, which the
 * UNMODIFIED reference produced (oracle/ref harness, tools/make_golden.py): primitive KATs,
 * pixel-centre hit ids / z of 13 scene+size combinations, deterministic Whitted images.
 *
 * Each function cites the reference lines it follows.  Float semantics: compile with
 * -O2 -ffp-contract=off (no FMA contraction), like the reference build of SURVEY.md section 8(c).
 * The stochastic branches (soft shadows, glossy lobes, depth of field) draw from a private
 * xorshift generator: the reference's are seeded by wall clock (RenderFunctions.cpp:60), so only
 * distributions are comparable there.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rtu.h"

#define BIG RTU_BIGFLOAT

typedef struct { float x, y, z; } v3;
typedef struct { v3 p, d; } ray_t;
typedef struct { float r, g, b; } col;

typedef struct {            /* HitInfo (scene.h:150-163) */
    float z;
    v3 p, N, uvw;
    int node, front, face;
} hit_t;

typedef struct {
    const rtu_scene_desc *S;
    int *first_child, *next_sibling;   /* child lists rebuilt from the pre-order parent indices */
    uint64_t trace_rays, shadow_rays, box_tests, tri_tests, node_visits;
    uint64_t rng;
    int shade_bounces;
} ctx_t;

/* ---- cyPoint3f arithmetic (cyPoint.h:292-349) */
static v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 mulf(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static float dot(v3 a, v3 b) { float x = a.x * b.x, y = a.y * b.y, z = a.z * b.z; return x + y + z; }
static v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static float len(v3 a) { return sqrtf(dot(a, a)); }
static v3 unit(v3 a) { float l = len(a); return V(a.x / l, a.y / l, a.z / l); }
static v3 mat(const float *m, v3 p) /* cyMatrix.h:542-547 */
{
    return V(p.x * m[0] + p.y * m[3] + p.z * m[6], p.x * m[1] + p.y * m[4] + p.z * m[7], p.x * m[2] + p.y * m[5] + p.z * m[8]);
}
static col C(float r, float g, float b) { col c = {r, g, b}; return c; }
static col cadd(col a, col b) { return C(a.r + b.r, a.g + b.g, a.b + b.b); }
static col cmul(col a, col b) { return C(a.r * b.r, a.g * b.g, a.b * b.b); }
static col cscale(col a, float s) { return C(a.r * s, a.g * s, a.b * s); }
static int cnonzero(col a) { return a.r != 0 || a.g != 0 || a.b != 0; }

static float smax(float a, float b) { return a < b ? b : a; } /* std::max */
static float smin(float a, float b) { return b < a ? b : a; } /* std::min */

static float urand(ctx_t *c) /* [0,1), stands in for rand()/RAND_MAX */
{
    c->rng ^= c->rng >> 12; c->rng ^= c->rng << 25; c->rng ^= c->rng >> 27;
    return (float)((c->rng * 2685821657736338717ULL) >> 40) / 16777216.0f;
}

/* ---- Box::IntersectRay / BVHBoxIntersection (objFunctions.cpp:143-254, 408-522).
 * The reference spells out four cases; here one pass computes the per-axis slab intervals and the
 * case only selects which axes take part and in which order they enter std::max/std::min. */
static int slab_interval(const ray_t *r, const float *b, float *tEntry, float *tExit)
{
    float t0[3], t1[3];
    const float *p = &r->p.x, *d = &r->d.x;
    int zero = d[0] == 0 ? 0 : (d[1] == 0 ? 1 : (d[2] == 0 ? 2 : 3));
    for (int k = 0; k < 3; k++) {
        if (k == zero) continue;
        t0[k] = (b[k] - p[k]) / d[k];
        t1[k] = (b[k + 3] - p[k]) / d[k];
        if (t0[k] > t1[k]) { float t = t1[k]; t1[k] = t0[k]; t0[k] = t; }
    }
    switch (zero) {
        case 0: *tEntry = smax(t0[2], t0[1]); *tExit = smin(t1[2], t1[1]); break;
        case 1: *tEntry = smax(t0[2], t0[0]); *tExit = smin(t1[2], t1[0]); break;
        case 2: *tEntry = smax(t0[1], t0[0]); *tExit = smin(t1[1], t1[0]); break;
        default: *tEntry = smax(smax(t0[0], t0[1]), t0[2]); *tExit = smin(smin(t1[0], t1[1]), t1[2]); break;
    }
    return zero;
}
static int box_empty(const float *b) { return b[0] > b[3] || b[1] > b[4] || b[2] > b[5]; } /* scene.h:87 */

int oracle_box_intersect(const float *ray6, const float *box6, float t_max)
{
    ray_t r = {{ray6[0], ray6[1], ray6[2]}, {ray6[3], ray6[4], ray6[5]}};
    float a, b;
    if (box_empty(box6)) return 0;
    slab_interval(&r, box6, &a, &b);
    return a <= b && a < t_max;
}
float oracle_bvh_box(const float *ray6, const float *box6, float t_max)
{
    ray_t r = {{ray6[0], ray6[1], ray6[2]}, {ray6[3], ray6[4], ray6[5]}};
    float a, b;
    if (box_empty(box6)) return -t_max;
    slab_interval(&r, box6, &a, &b);
    if (a <= b && a < t_max) return a + 0.01; /* double add, truncated to float (:517) */
    return t_max;
}
static int box_hit(ctx_t *c, const ray_t *r, const float *b, float t_max)
{
    float a, e;
    c->box_tests++;
    if (box_empty(b)) return 0;
    slab_interval(r, b, &a, &e);
    return a <= e && a < t_max;
}
static float bvh_box(ctx_t *c, const ray_t *r, const float *b, float t_max)
{
    float a, e;
    c->box_tests++;
    if (box_empty(b)) return -t_max;
    slab_interval(r, b, &a, &e);
    if (a <= e && a < t_max) return a + 0.01;
    return t_max;
}

/* ---- Sphere::IntersectRay (objFunctions.cpp:15-104) */
static void sphere_fill(const ray_t *r, hit_t *h)
{
    v3 t = add(r->p, mulf(r->d, h->z));
    h->N = h->front ? unit(t) : neg(unit(t));
    h->p = t;
    h->uvw = V(0.5 - atan2f(h->N.x, h->N.y) / (2 * M_PI), 0.5 + asinf(h->N.z) / M_PI, 0);
}
static int sphere_hit(ctx_t *c, const ray_t *r, hit_t *h)
{
    static const float unit_box[6] = {-1, -1, -1, 1, 1, 1};
    if (!box_hit(c, r, unit_box, BIG)) return 0;
    float a = dot(r->d, r->d);
    float b = 2 * dot(r->p, r->d);
    float cc = dot(r->p, r->p) - 1;
    float disc = b * b - 4 * a * cc;
    float m = (-b + sqrtf(disc)) / (2 * a);
    float n = (-b - sqrtf(disc)) / (2 * a);
    if (m == n && m < h->z && m >= 0.001) {
        h->z = m; h->front = 1;
        sphere_fill(r, h);
        return 1;
    }
    /* the two ordered-root branches are mirror images (:45-72 and :73-100): lo is the smaller root */
    float lo, hi;
    if (m < n) { lo = m; hi = n; } else if (n < m) { lo = n; hi = m; } else return 0;
    if (lo < h->z && ((m >= 0.001) | (n >= 0.001))) {
        if (lo <= 0.001 && hi > 0.001 && hi < h->z) { h->z = hi; h->front = 0; }
        else if (lo > 0.001) { h->z = lo; h->front = 1; }
        /* else: z and front keep their previous values but the record is still overwritten (SURVEY A-7) */
        sphere_fill(r, h);
        return 1;
    }
    return 0;
}

/* ---- Plane::IntersectRay (objFunctions.cpp:107-140) */
static int plane_hit(ctx_t *c, const ray_t *r, hit_t *h)
{
    static const float flat_box[6] = {-1, -1, 0, 1, 1, 0};
    if (!box_hit(c, r, flat_box, BIG)) return 0;
    if (r->d.z == 0) return 0;
    float t = (-r->p.z) / r->d.z;
    if (!(t > 0.001 && t < h->z)) return 0;
    v3 q = add(r->p, mulf(r->d, t));
    if (!(q.x > -1 && q.x < 1 && q.y > -1 && q.y < 1)) return 0;
    h->front = r->p.z > 0;
    h->N = V(0, 0, h->front ? 1 : -1);
    h->z = t;
    h->p = V(q.x, q.y, 0);
    h->uvw = V((q.x + 1) / 2, (q.y + 1) / 2, 0);
    return 1;
}

/* ---- TriObj::IntersectTriangle (objFunctions.cpp:257-328) */
static v3 vtx(const float *a, uint32_t i) { return V(a[3 * (size_t)i], a[3 * (size_t)i + 1], a[3 * (size_t)i + 2]); }
static v3 interp(const float *a, const uint32_t *idx, v3 bc) /* cyTriMesh.h:191 */
{
    return add(add(mulf(vtx(a, idx[0]), bc.x), mulf(vtx(a, idx[1]), bc.y)), mulf(vtx(a, idx[2]), bc.z));
}
static float cross2(float ax, float ay, float bx, float by) { return (-ay) * bx + ax * by; } /* cyPoint.h:248 */

static int tri_hit(ctx_t *c, const rtu_mesh *M, const ray_t *r, hit_t *h, uint32_t face)
{
    c->tri_tests++;
    const uint32_t *f = M->f + 3 * (size_t)face;
    v3 A = vtx(M->v, f[0]), B = vtx(M->v, f[1]), Cc = vtx(M->v, f[2]);
    v3 N = unit(cross(sub(B, A), sub(Cc, A)));
    float dn = dot(r->d, N);
    if (!(dn != 0)) return 0;
    float t = dot(sub(A, r->p), N) / dn;
    if (!(t > 0.00001 && t < h->z)) return 0;
    v3 q = add(r->p, mulf(r->d, t));
    float ax = fabsf(N.x), ay = fabsf(N.y), az = fabsf(N.z);
    float mx = smax(smax(ax, ay), az);
    int u, v; /* the two axes kept by the projection */
    if (mx == ax) { u = 1; v = 2; } else if (mx == ay) { u = 0; v = 2; } else if (mx == az) { u = 0; v = 1; } else return 0;
    const float *a = &A.x, *b = &B.x, *cc = &Cc.x, *qq = &q.x;
    float abc = cross2(cc[u] - a[u], cc[v] - a[v], b[u] - a[u], b[v] - a[v]) / 2.0;
    float apc = cross2(cc[u] - a[u], cc[v] - a[v], qq[u] - a[u], qq[v] - a[v]) / 2.0;
    float abp = cross2(qq[u] - a[u], qq[v] - a[v], b[u] - a[u], b[v] - a[v]) / 2.0;
    float b1 = apc / abc, b2 = abp / abc;
    float b3 = 1.0 - b1 - b2; /* double chain, then rounded (:304) */
    if (!(b1 > 0 && b2 > 0 && b3 > 0 && b1 < 1 && b2 < 1 && b3 < 1)) return 0;
    v3 bc = V(b3, b1, b2);
    h->front = dn < 0;
    h->uvw = (M->vt && M->ft) ? interp(M->vt, M->ft + 3 * (size_t)face, bc) : V(0, 0, 0);
    h->N = unit(interp(M->vn, M->fn + 3 * (size_t)face, bc));
    h->z = t;
    h->p = interp(M->v, f, bc);
    h->face = (int)face;
    return 1;
}

/* ---- TriObj::IntersectRay (objFunctions.cpp:333-406): cyBVH node words per cyBVH.h:187-200 */
static int mesh_hit(ctx_t *c, const rtu_mesh *M, const ray_t *r, hit_t *h)
{
    float mb[6] = {M->bound_min[0], M->bound_min[1], M->bound_min[2], M->bound_max[0], M->bound_max[1], M->bound_max[2]};
    if (!box_hit(c, r, mb, BIG)) return 0;
    if (M->nf == 0 || M->bvh_nodes < 2) return 0;
    uint32_t stack[100];
    int top = 0, hit = 0;
    stack[0] = 1;
    while (top >= 0) {
        uint32_t n = stack[top--];
        uint32_t w = M->bvh_data[n];
        if (!(w & 0x80000000u)) {
            uint32_t c1 = w & 0x7fffffffu, c2 = c1 + 1;
            float t1 = bvh_box(c, r, M->bvh_boxes + 6 * (size_t)c1, BIG);
            float t2 = bvh_box(c, r, M->bvh_boxes + 6 * (size_t)c2, BIG);
            if (t1 <= t2) {
                if (t2 != BIG) stack[++top] = c2;
                if (t1 != BIG) stack[++top] = c1;
            } else if (t1 > t2) {
                if (t1 != BIG) stack[++top] = c1;
                if (t2 != BIG) stack[++top] = c2;
            }
        } else {
            uint32_t off = w & 0x0fffffffu, cnt = ((w >> 28) & 7u) + 1u;
            for (uint32_t i = 0; i < cnt; i++) hit |= tri_hit(c, M, r, h, M->bvh_elements[off + i]);
        }
    }
    return hit;
}

static int object_hit(ctx_t *c, const rtu_node *nd, const ray_t *r, hit_t *h)
{
    c->node_visits++;
    switch (nd->kind) {
        case RTU_OBJ_SPHERE: return sphere_hit(c, r, h);
        case RTU_OBJ_PLANE: return plane_hit(c, r, h);
        case RTU_OBJ_MESH: return mesh_hit(c, &c->S->meshes[nd->mesh], r, h);
        default: return 0;
    }
}

/* ---- Node::ToNodeCoords / FromNodeCoords (scene.h:501-512, 235-242) */
static ray_t to_node(const rtu_node *n, const ray_t *r)
{
    v3 pos = V(n->pos[0], n->pos[1], n->pos[2]);
    ray_t o;
    o.p = mat(n->itm, sub(r->p, pos));
    o.d = sub(mat(n->itm, sub(add(r->p, r->d), pos)), o.p);
    return o;
}
static void from_node(const rtu_node *n, hit_t *h)
{
    v3 pos = V(n->pos[0], n->pos[1], n->pos[2]);
    h->p = add(mat(n->tm, h->p), pos);
    const float *m = n->itm;
    v3 t = V(dot(V(m[0], m[1], m[2]), h->N), dot(V(m[3], m[4], m[5]), h->N), dot(V(m[6], m[7], m[8]), h->N));
    h->N = unit(t);
}

/* ---- Trace / ShadowTrace (RenderFunctions.cpp:181-240) */
static int trace_node(ctx_t *c, const ray_t *r, int i, hit_t *h, int any)
{
    const rtu_node *nd = &c->S->nodes[i];
    ray_t lr = to_node(nd, r);
    int hit = 0;
    if (nd->kind != RTU_OBJ_NONE) {
        hit = object_hit(c, nd, &lr, h);
        if (hit) {
            if (any) return 1;
            h->node = i;
            from_node(nd, h);
        }
    }
    for (int ch = c->first_child[i]; ch >= 0; ch = c->next_sibling[ch]) {
        int chit = trace_node(c, &lr, ch, h, any);
        if (chit) {
            if (any) return 1;
            from_node(nd, h);
        }
        hit |= chit;
    }
    return hit;
}
static void hit_init(hit_t *h) /* HitInfo::Init (scene.h:162) */
{
    h->z = BIG; h->node = -1; h->front = 1; h->face = -1;
    h->uvw = V(0.5f, 0.5f, 0.5f); h->p = V(0, 0, 0); h->N = V(0, 0, 0);
}
static int trace(ctx_t *c, const ray_t *r, hit_t *h) { c->trace_rays++; return trace_node(c, r, 0, h, 0); }
static float shadow(ctx_t *c, ray_t r, float t_max) /* GenLight::Shadow (lightFunctions.cpp:27-37) */
{
    hit_t h;
    hit_init(&h);
    h.z = t_max;
    c->shadow_rays++;
    if (trace_node(c, &r, 0, &h, 1) && h.z > 0.0) return 0.0f;
    return 1.0f;
}

/* ---- textures (scene.h:355-365,382,421-431; texture.cpp:95-133) */
static float tile(float v) { float u = v - (int)v; if (u < 0) u += 1; return u; }
static col texmap_sample(const rtu_texmap *T, v3 uvw)
{
    if (T->kind == RTU_TEX_NULL) return C(0, 0, 0);
    v3 u = mat(T->itm, sub(uvw, V(T->pos[0], T->pos[1], T->pos[2])));
    float cu = tile(u.x), cv = tile(u.y);
    if (T->kind == RTU_TEX_CHECKER) {
        int first = (cu <= 0.5f) == (cv <= 0.5f);
        const float *k = first ? T->color1 : T->color2;
        return C(k[0], k[1], k[2]);
    }
    int W = T->width, H = T->height;
    if (W + H == 0 || !T->rgb8) return C(0, 0, 0);
    float x = W * cu, y = H * cv;
    int ix = (int)x, iy = (int)y;
    float fx = x - ix, fy = y - iy;
    if (ix < 0) ix -= (ix / W - 1) * W;
    if (ix >= W) ix -= (ix / W) * W;
    int ixp = ix + 1; if (ixp >= W) ixp -= W;
    if (iy < 0) iy -= (iy / H - 1) * H;
    if (iy >= H) iy -= (iy / H) * H;
    int iyp = iy + 1; if (iyp >= H) iyp -= H;
    const uint8_t *t[4] = {T->rgb8 + 3 * ((size_t)iy * W + ix), T->rgb8 + 3 * ((size_t)iy * W + ixp),
                           T->rgb8 + 3 * ((size_t)iyp * W + ix), T->rgb8 + 3 * ((size_t)iyp * W + ixp)};
    float wt[4] = {(1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy};
    float o[3];
    for (int k = 0; k < 3; k++)
        o[k] = (t[0][k] / 255.0f) * wt[0] + (t[1][k] / 255.0f) * wt[1] + (t[2][k] / 255.0f) * wt[2] + (t[3][k] / 255.0f) * wt[3];
    return C(o[0], o[1], o[2]);
}
static col tc_sample(const rtu_scene_desc *S, const rtu_texcolor *t, v3 uvw)
{
    col c = C(t->color[0], t->color[1], t->color[2]);
    if (t->texmap < 0 || t->texmap >= S->n_texmaps) return c;
    return cmul(c, texmap_sample(&S->texmaps[t->texmap], uvw));
}
static col env_sample(const rtu_scene_desc *S, v3 d)
{
    float z = asinf(-d.z) / (float)M_PI + 0.5f;
    float x = d.x / (fabsf(d.x) + fabsf(d.y));
    float y = d.y / (fabsf(d.x) + fabsf(d.y));
    v3 a = add(mulf(V(0.5f, 0.5f, 0), x), mulf(V(-0.5f, 0.5f, 0), y));
    return tc_sample(S, &S->environment, add(V(0.5f, 0.5f, 0.0f), mulf(a, z)));
}

/* ---- SampleSphere (RenderFunctions.cpp:282-302) */
static v3 sample_ball(ctx_t *c, float radius)
{
    if (!(radius > 0)) return V(0, 0, 0);
    for (;;) {
        v3 o = V(-radius + urand(c) * (radius * 2), -radius + urand(c) * (radius * 2), -radius + urand(c) * (radius * 2));
        if (!(len(o) > radius)) return o;
    }
}

/* ---- Light::Illuminate (lightFunctions.cpp:39-84, lights.h:32,48) and Direction */
static col illuminate(ctx_t *c, const rtu_light *L, v3 p)
{
    col I = C(L->intensity[0], L->intensity[1], L->intensity[2]);
    v3 pos = V(L->v[0], L->v[1], L->v[2]);
    if (L->kind == RTU_LIGHT_AMBIENT) return I;
    if (L->kind == RTU_LIGHT_DIRECT) {
        ray_t r = {p, neg(pos)};
        return cscale(I, shadow(c, r, BIG));
    }
    v3 target = pos;
    if (L->size > 0) {
        float sr = urand(c) * L->size, th = urand(c) * (float)(2 * M_PI);
        float ox = sr * cosf(th), oy = sr * sinf(th);
        v3 n = unit(sub(pos, p));
        v3 v1 = unit(cross(n, V(0, 0, 1)));
        v3 v2 = unit(cross(v1, n));
        target = add(add(pos, mulf(v1, ox)), mulf(v2, oy));
    }
    ray_t r = {p, unit(sub(target, p))};
    float s = shadow(c, r, len(sub(p, target)));
    v3 w = sub(pos, p);
    return cscale(cscale(I, s), 1 / dot(w, w));
}
static v3 light_direction(const rtu_light *L, v3 p)
{
    v3 v = V(L->v[0], L->v[1], L->v[2]);
    if (L->kind == RTU_LIGHT_DIRECT) return v;
    if (L->kind == RTU_LIGHT_POINT) return unit(sub(p, v));
    return V(0, 0, 0);
}

/* ---- MtlBlinn::Shade (mtlFunctions.cpp:120-298) */
/* amb != NULL: the light list is one AmbientLight of that intensity (the list MonteCarlo() builds,
 * RenderFunctions.cpp:587-590); amb == NULL: the scene's lights */
static col shade(ctx_t *c, const ray_t *ray, const hit_t *h, int bounce, const col *amb);

#define shade_hit_of(c, r, h, bounce) shade(c, r, h, bounce, amb)

static v3 mirror(v3 d, v3 n) { return unit(sub(d, mulf(n, 2 * dot(d, n)))); }

static col shade(ctx_t *c, const ray_t *ray, const hit_t *h, int bounce, const col *amb)
{
    const rtu_scene_desc *S = c->S;
    int mi = S->nodes[h->node].material;
    if (mi < 0) return C(1, 1, 1); /* the reference would dereference NULL here */
    const rtu_material *M = &S->materials[mi];
    col out = C(0, 0, 0);
    if (h->front && amb) {
        out = cadd(out, cmul(tc_sample(S, &M->diffuse, h->uvw), *amb)); /* :131-133 with the single ambient light */
    } else if (h->front) {
        for (int i = 0; i < S->n_lights; i++) {
            const rtu_light *L = &S->lights[i];
            if (L->kind == RTU_LIGHT_AMBIENT) {
                out = cadd(out, cmul(tc_sample(S, &M->diffuse, h->uvw), illuminate(c, L, h->p)));
                continue;
            }
            v3 view = unit(sub(V(S->camera.pos[0], S->camera.pos[1], S->camera.pos[2]), h->p)); /* camera.pos, not the ray (:137) */
            v3 ld = unit(neg(light_direction(L, h->p)));
            v3 hv = unit(add(view, ld));
            float ndl = dot(h->N, ld), ndh = dot(h->N, hv);
            if (ndl < 0.0) ndl = 0.0;
            if (ndh < 0.0) ndh = 0.0;
            col il = illuminate(c, L, h->p);
            col brdf = cadd(tc_sample(S, &M->diffuse, h->uvw), cscale(tc_sample(S, &M->specular, h->uvw), powf(ndh, M->glossiness)));
            out = cadd(out, cmul(cscale(il, ndl), brdf));
        }
    }
    if (bounce <= 0) return out;
    col Kt = tc_sample(S, &M->refraction, h->uvw);
    if (cnonzero(Kt)) {
        v3 so = add(h->p, h->N);
        v3 sn = unit(sub(add(so, sample_ball(c, M->refraction_glossiness)), h->p));
        float cos1 = dot(sn, neg(ray->d));
        float sin1 = sqrt(1 - pow(cos1, 2));
        if (sin1 > 1) sin1 = 1.0;
        if (sin1 < -1) sin1 = -1.0;
        if (cos1 > 1) cos1 = 1.0;
        if (cos1 < -1) cos1 = -1.0;
        float n1 = M->ior, n2 = 1.0;
        if (h->front) { n1 = 1.0; n2 = M->ior; }
        float sin2 = (n1 / n2) * sin1;
        float cos2 = sqrtf(1 - sin2 * sin2);
        if (cos2 > 1) cos2 = 1.0;
        v3 sv = unit(cross(sn, unit(cross(sn, neg(ray->d)))));
        if (sin2 > 1) { /* total internal reflection; absorption from an un-traced HitInfo (z = BIGFLOAT, SURVEY A-11) */
            ray_t rr = {h->p, mirror(ray->d, sn)};
            hit_t rh;
            hit_init(&rh);
            col ab = C(expf((-rh.z) * M->absorption[0]), expf((-rh.z) * M->absorption[1]), expf((-rh.z) * M->absorption[2]));
            if (trace(c, &rr, &rh)) out = cadd(out, cmul(ab, shade_hit_of(c, &rr, &rh, bounce - 1)));
        } else {
            v3 sn2 = unit(sub(add(so, sample_ball(c, M->refraction_glossiness)), h->p)); /* second sample shadows the first (:225-227) */
            v3 rd = unit(add(mulf(neg(sn2), cos2), mulf(sv, sin2)));
            ray_t rr = {h->p, rd};
            hit_t rh;
            hit_init(&rh);
            if (trace(c, &rr, &rh)) {
                float R0 = pow((n1 - n2) / (n1 + n2), 2);
                float Fr = R0 + (1.0 - R0) * pow((1.0 - cos1), 5);
                ray_t fr = {h->p, mirror(ray->d, sn2)};
                hit_t fh;
                hit_init(&fh);
                col fres;
                if (trace(c, &fr, &fh)) fres = cmul(Kt, shade_hit_of(c, &fr, &fh, bounce - 1));
                else fres = env_sample(S, fr.d);
                col rres = shade_hit_of(c, &rr, &rh, bounce - 1);
                col ab = C(1, 1, 1);
                if (!rh.front) ab = C(expf((-rh.z) * M->absorption[0]), expf((-rh.z) * M->absorption[1]), expf((-rh.z) * M->absorption[2]));
                out = cadd(out, cadd(cscale(cmul(cmul(ab, Kt), rres), 1.0 - Fr), cscale(fres, Fr)));
            } else {
                out = cadd(out, env_sample(S, rd));
            }
        }
    }
    col Kr = tc_sample(S, &M->reflection, h->uvw);
    if (cnonzero(Kr)) {
        v3 so = add(h->p, h->N);
        v3 sn = unit(sub(add(so, sample_ball(c, M->reflection_glossiness)), h->p));
        ray_t rr = {h->p, mirror(ray->d, sn)};
        hit_t rh;
        hit_init(&rh);
        if (trace(c, &rr, &rh)) out = cadd(out, cmul(Kr, shade_hit_of(c, &rr, &rh, bounce - 1)));
        else out = cadd(out, cmul(env_sample(S, rr.d), C(M->reflection.color[0], M->reflection.color[1], M->reflection.color[2])));
    }
    return out;
}

/* ---- SampleHemiSphereCosine (RenderFunctions.cpp:320-337), radius 1; the tangent is built from
 * normal x (s,s,s) with s the first random number (SURVEY A-14) */
static v3 sample_hemi_cos(ctx_t *c, v3 n)
{
    float sx = urand(c);
    float phi = urand(c) * (float)(2 * M_PI);
    float theta = 0.5 * acos(1 - 2 * sx);
    v3 v1 = unit(cross(n, V(sx, sx, sx)));
    v3 v2 = unit(cross(v1, n));
    return add(add(mulf(n, cosf(theta)), mulf(v1, sinf(theta) * cosf(phi))), mulf(v2, sinf(theta) * sinf(phi)));
}

/* ---- MonteCarlo (RenderFunctions.cpp:454-591): intensity of the AmbientLight it appends */
static col monte_carlo(ctx_t *c, const hit_t *h, int bounces)
{
    if (bounces <= 0) return C(0.1f, 0.1f, 0.1f);                                /* :584 */
    ray_t r = {h->p, unit(sample_hemi_cos(c, h->N))};                             /* :561-562 */
    hit_t h2;
    hit_init(&h2);
    if (!trace(c, &r, &h2)) return env_sample(c->S, r.d);                         /* :575 */
    col amb2 = monte_carlo(c, &h2, bounces - 1);                                  /* :568 */
    return cadd(shade(c, &r, &h2, 5, &amb2), shade(c, &r, &h2, 5, NULL));         /* :569-570; /= monteCarloSampleSize (1) */
}

/* ---- camera (RenderFunctions.cpp:243-269, 88-97) */
typedef struct { v3 pos, origin, u, v, lx, ly; float dof; } cam_t;
static void make_camera(const rtu_camera *c, int W, int H, cam_t *o)
{
    v3 pos = V(c->pos[0], c->pos[1], c->pos[2]), dir = V(c->dir[0], c->dir[1], c->dir[2]), up = V(c->up[0], c->up[1], c->up[2]);
    float d = c->focaldist;
    float ah = tan((c->fov / 2) * M_PI / 180.0) * 2 * d;
    float aw = ((float)W / (float)H) * ah;
    v3 right = unit(cross(unit(dir), unit(up)));
    v3 top = add(add(pos, mulf(unit(dir), d)), mulf(unit(up), ah / 2));
    o->pos = pos;
    o->origin = sub(top, mulf(right, aw / 2));
    o->u = mulf(right, aw / (float)W);
    o->v = mulf(mulf(unit(up), -1), ah / (float)H);
    o->lx = right; o->ly = up; o->dof = c->dof;
}
static ray_t camera_ray(ctx_t *c, const cam_t *cm, int x, int y, float ox, float oy)
{
    v3 lens = cm->pos;
    if (cm->dof > 0) {
        float sx = urand(c), th = urand(c) * (float)(2 * M_PI);
        float cx = sqrtf(sx * cm->dof * cm->dof) * cosf(th), cy = sqrtf(sx * cm->dof * cm->dof) * sinf(th);
        lens = add(add(cm->pos, mulf(cm->ly, cy)), mulf(cm->lx, cx));
    }
    v3 pt = add(add(cm->origin, mulf(cm->u, x + ox)), mulf(cm->v, y + oy));
    ray_t r = {lens, unit(sub(pt, lens))};
    return r;
}
static float halton(int index, int base) /* scene.h:130-139 */
{
    float r = 0, f = 1.0f / (float)base;
    for (int i = index; i > 0; i /= base) { r += f * (i % base); f /= (float)base; }
    return r;
}

/* ---- context */
static int ctx_init(ctx_t *c, const rtu_scene_desc *S, uint64_t seed)
{
    memset(c, 0, sizeof *c);
    c->S = S;
    c->rng = 0x9E3779B97F4A7C15ULL ^ (seed * 0xD1B54A32D192ED03ULL + 1);
    int n = S->n_nodes;
    c->first_child = (int *)malloc(sizeof(int) * (size_t)n * 2);
    if (!c->first_child) return 0;
    c->next_sibling = c->first_child + n;
    int *last = (int *)malloc(sizeof(int) * (size_t)n);
    if (!last) { free(c->first_child); return 0; }
    for (int i = 0; i < n; i++) { c->first_child[i] = -1; c->next_sibling[i] = -1; last[i] = -1; }
    for (int i = 1; i < n; i++) {
        int p = S->nodes[i].parent;
        if (last[p] < 0) c->first_child[p] = i; else c->next_sibling[last[p]] = i;
        last[p] = i;
    }
    free(last);
    return 1;
}
static void ctx_free(ctx_t *c) { free(c->first_child); }

/* ---- public entry points */
typedef struct { uint64_t trace_rays, shadow_rays, box_tests, tri_tests, node_visits; double seconds; } oracle_stats;

static void fill_hit(rtu_hit *o, const hit_t *h, int hit)
{
    o->z = h->z; o->node = hit ? h->node : -1; o->front = h->front; o->face = hit ? h->face : -1;
    o->p[0] = h->p.x; o->p[1] = h->p.y; o->p[2] = h->p.z;
    o->N[0] = h->N.x; o->N[1] = h->N.y; o->N[2] = h->N.z;
    o->uvw[0] = h->uvw.x; o->uvw[1] = h->uvw.y; o->uvw[2] = h->uvw.z;
}

int oracle_trace(const rtu_scene_desc *S, const rtu_ray *rays, int64_t n, rtu_hit *hits)
{
    ctx_t c;
    if (!ctx_init(&c, S, 0)) return 1;
    for (int64_t i = 0; i < n; i++) {
        ray_t r = {{rays[i].p[0], rays[i].p[1], rays[i].p[2]}, {rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]}};
        hit_t h;
        hit_init(&h);
        int hit = trace(&c, &r, &h);
        if (hit && S->nodes[h.node].kind != RTU_OBJ_MESH) h.face = -1;
        fill_hit(&hits[i], &h, hit);
    }
    ctx_free(&c);
    return 0;
}

int oracle_shadow_trace(const rtu_scene_desc *S, const rtu_ray *rays, const float *t_max, int64_t n, uint8_t *occ)
{
    ctx_t c;
    if (!ctx_init(&c, S, 0)) return 1;
    for (int64_t i = 0; i < n; i++) {
        ray_t r = {{rays[i].p[0], rays[i].p[1], rays[i].p[2]}, {rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]}};
        occ[i] = shadow(&c, r, t_max[i]) == 0.0f;
    }
    ctx_free(&c);
    return 0;
}

int oracle_shade(const rtu_scene_desc *S, const rtu_ray *rays, const rtu_hit *hits, int64_t n, int bounces, float *rgb)
{
    ctx_t c;
    if (!ctx_init(&c, S, 0)) return 1;
    for (int64_t i = 0; i < n; i++) {
        rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = 0;
        if (hits[i].node < 0) continue;
        ray_t r = {{rays[i].p[0], rays[i].p[1], rays[i].p[2]}, {rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]}};
        hit_t h;
        h.z = hits[i].z; h.node = hits[i].node; h.front = hits[i].front; h.face = hits[i].face;
        h.p = V(hits[i].p[0], hits[i].p[1], hits[i].p[2]);
        h.N = V(hits[i].N[0], hits[i].N[1], hits[i].N[2]);
        h.uvw = V(hits[i].uvw[0], hits[i].uvw[1], hits[i].uvw[2]);
        col o = shade(&c, &r, &h, bounces, NULL);
        rgb[3 * i] = o.r; rgb[3 * i + 1] = o.g; rgb[3 * i + 2] = o.b;
    }
    ctx_free(&c);
    return 0;
}

typedef struct {
    const rtu_scene_desc *S;
    const rtu_params *P;
    rtu_image *out;
    int W, H, y0, y1, x0, x1;
    cam_t cam;
    volatile int *next_row;
    oracle_stats st;
    uint64_t seed;
} job_t;

static col photon_mapping(ctx_t *c, const ray_t *r, const hit_t *h);
static col monte_carlo_photon(ctx_t *c, const hit_t *h, int x, int y, int W, int H, int bounces);

static void *render_rows(void *arg)
{
    job_t *J = (job_t *)arg;
    ctx_t c;
    if (!ctx_init(&c, J->S, J->seed)) return NULL;
    const rtu_params *P = J->P;
    int W = J->W;
    int s0 = 0, s1 = P->spp;
    if (P->sample_begin || P->sample_end) { s0 = P->sample_begin; s1 = P->sample_end; }
    float inc = 1.0 / P->spp; /* RenderFunctions.cpp:71 */
    for (;;) {
        int y = __sync_fetch_and_add(J->next_row, 1);
        if (y >= J->y1) break;
        for (int x = J->x0; x < J->x1; x++) {
            size_t i = (size_t)x + (size_t)W * y;
            if (P->mode == RTU_MODE_PRIMARY || J->out->z || J->out->node_id || J->out->face_id) {
                ray_t r = camera_ray(&c, &J->cam, x, y, 0.5f, 0.5f);
                hit_t h;
                hit_init(&h);
                int hit = trace(&c, &r, &h);
                if (J->out->z) J->out->z[i] = h.z;
                if (J->out->node_id) J->out->node_id[i] = hit ? h.node : -1;
                if (J->out->face_id) J->out->face_id[i] = (hit && J->S->nodes[h.node].kind == RTU_OBJ_MESH) ? h.face : -1;
                if (P->mode == RTU_MODE_PRIMARY) continue;
            }
            col sum = C(0, 0, 0);
            for (int s = s0; s < s1; s++) {
                float ox = 0.5f, oy = 0.5f;
                if (P->pattern == RTU_PATTERN_REFERENCE) { float cur = s * inc; ox = cur + halton(s, 4); oy = cur + halton(s, 5); }
                ray_t r = camera_ray(&c, &J->cam, x, y, ox, oy);
                hit_t h;
                hit_init(&h);
                col v;
                if (trace(&c, &r, &h)) {
                    if (P->mode == RTU_MODE_PHOTON) {                             /* RenderFunctions.cpp:141-142 */
                        v = photon_mapping(&c, &r, &h);
                    } else if (P->mode == RTU_MODE_PHOTON_GATHER) {               /* RenderFunctions.cpp:137-139 */
                        v = cadd(shade(&c, &r, &h, P->shade_bounces, NULL), monte_carlo_photon(&c, &h, x, y, W, J->H, P->gi_bounces));
                    } else if (P->mode == RTU_MODE_PATH) {                        /* RenderFunctions.cpp:129-135 */
                        col amb = monte_carlo(&c, &h, P->gi_bounces);
                        v = cadd(shade(&c, &r, &h, P->shade_bounces, &amb), shade(&c, &r, &h, P->shade_bounces, NULL));
                    } else {
                        v = shade(&c, &r, &h, P->shade_bounces, NULL);
                    }
                }
                else v = tc_sample(J->S, &J->S->background, V((float)x / W, (float)y / J->H, 0)); /* :145 */
                sum = cadd(sum, v);
            }
            float n = (float)P->spp;
            sum = C(sum.r / n, sum.g / n, sum.b / n); /* :152 */
            if (J->out->rgb) { J->out->rgb[3 * i] = sum.r; J->out->rgb[3 * i + 1] = sum.g; J->out->rgb[3 * i + 2] = sum.b; }
            if (J->out->rgb8) {
                float g[3] = {(float)pow(sum.r, 1 / 2.2), (float)pow(sum.g, 1 / 2.2), (float)pow(sum.b, 1 / 2.2)}; /* :155-157 */
                for (int k = 0; k < 3; k++) {
                    int q = (int)(g[k] * 255); /* cyColor.h:245-246 */
                    J->out->rgb8[3 * i + k] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
                }
            }
        }
    }
    J->st.trace_rays = c.trace_rays; J->st.shadow_rays = c.shadow_rays; J->st.box_tests = c.box_tests;
    J->st.tri_tests = c.tri_tests; J->st.node_visits = c.node_visits;
    ctx_free(&c);
    return NULL;
}

/* crop = {x0,y0,x1,y1} or NULL; threads >= 1.  Pixels outside the crop are left untouched. */
int oracle_render(const rtu_scene_desc *S, const rtu_params *P, rtu_image *out, const int *crop, int threads, oracle_stats *st)
{
    int W = P->width > 0 ? P->width : S->camera.width, H = P->height > 0 ? P->height : S->camera.height;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    job_t jobs[256];
    pthread_t th[256];
    volatile int next = crop ? crop[1] : 0;
    cam_t cam;
    make_camera(&S->camera, W, H, &cam);
    for (int t = 0; t < threads; t++) {
        job_t *J = &jobs[t];
        memset(J, 0, sizeof *J);
        J->S = S; J->P = P; J->out = out; J->W = W; J->H = H; J->cam = cam; J->next_row = &next;
        J->x0 = crop ? crop[0] : 0; J->x1 = crop ? crop[2] : W; J->y0 = crop ? crop[1] : 0; J->y1 = crop ? crop[3] : H;
        J->seed = P->seed * 1315423911u + (uint64_t)t;
        if (P->row_begin || P->row_end) { if (J->y0 < P->row_begin) J->y0 = P->row_begin; if (J->y1 > P->row_end) J->y1 = P->row_end; }
    }
    if (P->row_begin || P->row_end) { if (next < P->row_begin) next = P->row_begin; }
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, render_rows, &jobs[t]);
    if (st) memset(st, 0, sizeof *st);
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        if (st) {
            st->trace_rays += jobs[t].st.trace_rays; st->shadow_rays += jobs[t].st.shadow_rays; st->box_tests += jobs[t].st.box_tests;
            st->tri_tests += jobs[t].st.tri_tests; st->node_visits += jobs[t].st.node_visits;
        }
    }
    /* RenderImage::ComputeZBufferImage (scene.h:590-612) */
    if (out->z && out->z8 && !crop) {
        size_t n = (size_t)W * H;
        float zmin = BIG, zmax = 0;
        for (size_t i = 0; i < n; i++) {
            if (out->z[i] == BIG) continue;
            if (zmin > out->z[i]) zmin = out->z[i];
            if (zmax < out->z[i]) zmax = out->z[i];
        }
        for (size_t i = 0; i < n; i++) {
            if (out->z[i] == BIG) { out->z8[i] = 0; continue; }
            float f = (zmax - out->z[i]) / (zmax - zmin);
            int q = (int)(f * 255);
            out->z8[i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
        }
    }
    return 0;
}

/* TexturedColor::Sample / SampleEnvironment on caller inputs (texture KATs) */
int oracle_sample_texcolor(const rtu_scene_desc *S, const rtu_texcolor *t, const float *uvw, int64_t n, float *rgb)
{
    for (int64_t i = 0; i < n; i++) {
        col c = tc_sample(S, t, V(uvw[3 * i], uvw[3 * i + 1], uvw[3 * i + 2]));
        rgb[3 * i] = c.r; rgb[3 * i + 1] = c.g; rgb[3 * i + 2] = c.b;
    }
    return 0;
}
int oracle_sample_environment(const rtu_scene_desc *S, const float *dir, int64_t n, float *rgb)
{
    for (int64_t i = 0; i < n; i++) {
        col c = env_sample(S, V(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
        rgb[3 * i] = c.r; rgb[3 * i + 1] = c.g; rgb[3 * i + 2] = c.b;
    }
    return 0;
}

/* ================================================================== photon map
 * cyPhotonMap.h restated on rtu_photon arrays.  Pinned by tests/golden/kat_photonmap.npz, which the reference's own
 * cyPhotonMap produced: the balanced array is reproduced byte for byte and the estimates bit for bit. */

/* Photon::GetDirection (cyPhotonMap.h:158-181), including the `dirY-dirY` slip at :168 that leaves z = sqrt(1 - x^2) */
static v3 photon_direction(const rtu_photon *p)
{
    v3 d;
    d.x = (float)p->dir_x / (float)0x7FFF;
    d.y = (float)p->dir_y / (float)0x7FFF;
    int xy2 = p->dir_x * p->dir_x + p->dir_y - p->dir_y;
    if (xy2 > 0x3FFF0001) xy2 = 0x3FFF0001;
    int z2 = 0x3FFF0001 - xy2;
    int z = 0, place = 0x40000000, rem = z2;
    while (place > rem) place >>= 2;
    while (place) {
        if (rem >= z + place) {
            rem = rem - z - place;
            z = z + (place << 1);
        }
        z >>= 1;
        place >>= 2;
    }
    d.z = (float)z / (float)0x7FFF;
    if (p->plane_dirz & 0x8) d.z = -d.z;
    return d;
}

typedef struct {
    rtu_photon *work; /* 1-based scratch copy that gets partitioned in place */
    rtu_photon *out;  /* 1-based heap-ordered result */
} balance_t;

/* PhotonMap::BalanceSegment (cyPhotonMap.h:230-290) */
static void balance_segment(balance_t *b, v3 lo, v3 hi, int index, int start, int end)
{
    int median = 1;
    while (4 * median <= end - start + 1) median += median;
    if (3 * median <= end - start + 1) {
        median += median;
        median += start - 1;
    } else {
        median = end - median + 1;
    }
    int axis = 2;
    v3 ext = sub(hi, lo);
    if (ext.x > ext.y) {
        if (ext.x > ext.z) axis = 0;
    } else if (ext.y > ext.z) {
        axis = 1;
    }
    rtu_photon *ph = b->work;
    int left = start, right = end;
    while (right > left) {
        const float v = ph[right].position[axis];
        int i = left - 1, j = right;
        while (ph[++i].position[axis] < v) {}
        while (ph[--j].position[axis] > v && j > left) {}
        while (i < j) {
            rtu_photon t = ph[i]; ph[i] = ph[j]; ph[j] = t;
            while (ph[++i].position[axis] < v) {}
            while (ph[--j].position[axis] > v && j > left) {}
        }
        rtu_photon t = ph[i]; ph[i] = ph[right]; ph[right] = t;
        if (i >= median) right = i - 1;
        if (i <= median) left = i + 1;
    }
    b->out[index] = ph[median];
    b->out[index].plane_dirz = (uint8_t)((b->out[index].plane_dirz & 0x8) | (axis & 0x3)); /* SetPlane (:64) */
    float split = b->out[index].position[axis];
    if (median > start) {
        if (start < median - 1) {
            v3 h2 = hi;
            if (axis == 0) h2.x = split; else if (axis == 1) h2.y = split; else h2.z = split;
            balance_segment(b, lo, h2, 2 * index, start, median - 1);
        } else {
            b->out[2 * index] = ph[start];
        }
    }
    if (median < end) {
        if (median + 1 < end) {
            v3 l2 = lo;
            if (axis == 0) l2.x = split; else if (axis == 1) l2.y = split; else l2.z = split;
            balance_segment(b, l2, hi, 2 * index + 1, median + 1, end);
        } else {
            b->out[2 * index + 1] = ph[end];
        }
    }
}

/* PhotonMap::PrepareForIrradianceEstimation (cyPhotonMap.h:207-228).  in: n photons; out: n+1 records (out[0] is
 * the unused, zeroed slot 0 of the reference's vector).  The bounding box starts from that zero slot (:212-213). */
int oracle_balance_photons(const rtu_photon *in, uint32_t n, rtu_photon *out)
{
    memset(out, 0, sizeof(rtu_photon) * ((size_t)n + 1));
    if (n == 0) return 0;
    balance_t b;
    b.work = (rtu_photon *)malloc(sizeof(rtu_photon) * ((size_t)n + 1));
    if (!b.work) return 1;
    memset(&b.work[0], 0, sizeof(rtu_photon));
    memcpy(&b.work[1], in, sizeof(rtu_photon) * (size_t)n);
    b.out = out;
    v3 lo = V(0, 0, 0), hi = V(0, 0, 0);
    for (uint32_t i = 1; i <= n; i++) {
        const float *p = b.work[i].position;
        if (lo.x > p[0]) lo.x = p[0];
        if (hi.x < p[0]) hi.x = p[0];
        if (lo.y > p[1]) lo.y = p[1];
        if (hi.y < p[1]) hi.y = p[1];
        if (lo.z > p[2]) lo.z = p[2];
        if (hi.z < p[2]) hi.z = p[2];
    }
    balance_segment(&b, lo, hi, 1, 1, (int)n);
    free(b.work);
    return 0;
}

#define ORACLE_MAX_PHOTONS 100 /* photonSampleSize (RenderFunctions.cpp:33) */
typedef struct {
    v3 pos, normal;
    int has_normal;
    float norm_scale;
    int found;
    float dist2[ORACLE_MAX_PHOTONS + 1];
    rtu_photon photon[ORACLE_MAX_PHOTONS + 1];
} nearest_t;

/* PhotonMap::LocatePhotons (cyPhotonMap.h:350-424) */
static void locate_photons(const rtu_photon *map, int half_stored, nearest_t *np, int index)
{
    const rtu_photon *p = &map[index];
    int axis = p->plane_dirz & 0x3;
    if (index < half_stored) {
        float qa = axis == 0 ? np->pos.x : (axis == 1 ? np->pos.y : np->pos.z);
        float dist = qa - p->position[axis];
        if (dist > 0) {
            locate_photons(map, half_stored, np, 2 * index + 1);
            if (dist * dist < np->dist2[0]) locate_photons(map, half_stored, np, 2 * index);
        } else {
            locate_photons(map, half_stored, np, 2 * index);
            if (dist * dist < np->dist2[0]) locate_photons(map, half_stored, np, 2 * index + 1);
        }
    }
    v3 dif = sub(V(p->position[0], p->position[1], p->position[2]), np->pos);
    float dist2 = dot(dif, dif);
    if (dist2 < np->dist2[0]) {
        if (np->has_normal) {
            v3 dir = photon_direction(p);
            if (dot(dir, np->normal) >= 0) return;
            if (np->norm_scale > 0) {
                float perp = dot(dif, np->normal);
                dif = add(dif, mulf(np->normal, perp * np->norm_scale));
                dist2 = dot(dif, dif);
                if (dist2 >= np->dist2[0]) return;
            }
        }
        if (np->found < ORACLE_MAX_PHOTONS) {
            np->found++;
            np->dist2[np->found] = dist2;
            np->photon[np->found] = *p;
            if (np->found == ORACLE_MAX_PHOTONS) { /* heapify */
                int half_found = np->found >> 1;
                for (int k = half_found; k >= 1; k--) {
                    int parent = k;
                    rtu_photon tp = np->photon[k];
                    float td2 = np->dist2[k];
                    while (parent <= half_found) {
                        int j = parent + parent;
                        if (j < np->found && np->dist2[j] < np->dist2[j + 1]) j++;
                        if (td2 >= np->dist2[j]) break;
                        np->dist2[parent] = np->dist2[j];
                        np->photon[parent] = np->photon[j];
                        parent = j;
                    }
                    np->photon[parent] = tp;
                    np->dist2[parent] = td2;
                }
            }
        } else {
            int parent = 1, j = 2;
            while (j <= np->found) {
                if (j < np->found && np->dist2[j] < np->dist2[j + 1]) j++;
                if (dist2 > np->dist2[j]) break;
                np->dist2[parent] = np->dist2[j];
                np->photon[parent] = np->photon[j];
                parent = j;
                j <<= 1;
            }
            np->photon[parent] = *p;
            np->dist2[parent] = dist2;
            np->dist2[0] = np->dist2[1];
        }
    }
}

/* PhotonMap::EstimateIrradiance<100>, FILTER_TYPE_CONSTANT (cyPhotonMap.h:276-323).
 * map: n+1 balanced records (oracle_balance_photons), pos/normal: nq x 3. */
int oracle_estimate_irradiance(const rtu_photon *map, uint32_t n, const float *pos, const float *normal, int64_t nq, float radius,
                               float ellipticity, float *irrad, float *direction, int32_t *found)
{
    int half_stored = (int)n / 2 - 1;
    for (int64_t q = 0; q < nq; q++) {
        col e = C(0, 0, 0);
        v3 d = V(0, 0, 0);
        nearest_t np;
        np.pos = V(pos[q * 3], pos[q * 3 + 1], pos[q * 3 + 2]);
        np.has_normal = normal != NULL;
        if (normal) np.normal = V(normal[q * 3], normal[q * 3 + 1], normal[q * 3 + 2]);
        np.norm_scale = ellipticity == 1 ? 0 : 1 / ellipticity - 1;
        np.found = 0;
        np.dist2[0] = radius * radius;
        if (n > 0) locate_photons(map, half_stored, &np, 1);
        for (int i = 1; i <= np.found; i++) {
            const rtu_photon *p = &np.photon[i];
            col pw = cscale(C(p->color[0] / 255.0f, p->color[1] / 255.0f, p->color[2] / 255.0f), p->power); /* GetPower (:57) */
            float filter = 1;
            e = cadd(e, cscale(pw, filter));
            d = add(d, mulf(photon_direction(p), filter * p->power));
        }
        if (np.found > 0) {
            float area = (float)M_PI * np.dist2[0];
            if (area > 0) {
                const float inv = 1.0f / area;
                e = cscale(e, inv);
            }
            d = unit(d); /* Normalize(): *this /= Length() */
        }
        irrad[q * 3] = e.r; irrad[q * 3 + 1] = e.g; irrad[q * 3 + 2] = e.b;
        direction[q * 3] = d.x; direction[q * 3 + 1] = d.y; direction[q * 3 + 2] = d.z;
        if (found) found[q] = np.found;
    }
    return 0;
}


/* The map RTU_MODE_PHOTON renders with: balanced records (n+1) and the estimate's radius / ellipticity. */
static const rtu_photon *g_map = NULL;
static uint32_t g_map_n = 0;
static float g_map_radius = 1.0f, g_map_ellipticity = 0.5f;

int oracle_set_photon_map(const rtu_photon *balanced, uint32_t n, float radius, float ellipticity)
{
    g_map = balanced; g_map_n = n; g_map_radius = radius; g_map_ellipticity = ellipticity;
    return 0;
}

/* PhotonMapping (RenderFunctions.cpp:394-413): Shade(r, hInfo, {PhotonLight(irradiance, direction)}, 0).
 * PhotonLight (lights.h:61-74): Illuminate = intensity (no shadow ray), Direction = direction.GetNormalized(). */
static col photon_mapping(ctx_t *c, const ray_t *r, const hit_t *h)
{
    const rtu_scene_desc *S = c->S;
    float pos[3] = {h->p.x, h->p.y, h->p.z}, nrm[3] = {h->N.x, h->N.y, h->N.z}, e[3], d[3];
    oracle_estimate_irradiance(g_map, g_map_n, pos, nrm, 1, g_map_radius, g_map_ellipticity, e, d, NULL);
    v3 dir = unit(V(d[0], d[1], d[2])); /* SetDirection normalises again; 0/0 = NaN when no photon was found */
    int mi = S->nodes[h->node].material;
    if (mi < 0) return C(1, 1, 1);
    const rtu_material *M = &S->materials[mi];
    col out = C(0, 0, 0);
    if (h->front) {
        v3 view = unit(sub(V(S->camera.pos[0], S->camera.pos[1], S->camera.pos[2]), h->p));
        v3 ld = unit(neg(dir));
        v3 hv = unit(add(view, ld));
        float ndl = dot(h->N, ld), ndh = dot(h->N, hv);
        if (ndl < 0.0) ndl = 0.0;
        if (ndh < 0.0) ndh = 0.0;
        col il = C(e[0], e[1], e[2]);
        col brdf = cadd(tc_sample(S, &M->diffuse, h->uvw), cscale(tc_sample(S, &M->specular, h->uvw), powf(ndh, M->glossiness)));
        out = cadd(out, cmul(cscale(il, ndl), brdf));
    }
    (void)r;
    return out;
}


/* MonteCarloPhoton(hInfo, x, y, 1) (RenderFunctions.cpp:416-451).  Every bounce samples the hemisphere of the SAME
 * first hit, and the HitInfo of the sample rays is initialised once and never reset, so a later sample only "hits"
 * what is nearer than the previous sample's hit (SURVEY A-16 for the same pattern in the photon bounces). */
static col monte_carlo_photon(ctx_t *c, const hit_t *h, int x, int y, int W, int H, int bounces)
{
    col sum = C(0, 0, 0);
    hit_t hs;
    hit_init(&hs);
    int actual = 0;
    for (int b = 0; b < bounces; b++) {
        ray_t r = {h->p, unit(sample_hemi_cos(c, h->N))};                          /* :428-429 */
        actual++;
        if (trace(c, &r, &hs)) {
            sum = cadd(sum, photon_mapping(c, &r, &hs));                           /* :436-437 */
        } else {
            sum = cadd(sum, tc_sample(c->S, &c->S->background, V((float)x / W, (float)y / H, 0))); /* :440 */
            break;
        }
    }
    float n = (float)actual;
    return C(sum.r / n, sum.g / n, sum.b / n);                                     /* :445; /= monteCarloSampleSize (1) */
}
