/* CPU restatement of the ALGORITHM of the device photon estimate (raytracer-utah_b200/csrc/photon_kernels.cu:
 * k_knn_init / k_knn_merge / k_knn_candidates / k_knn_replay), line for line where it matters: the heap-free candidate walk
 * with its histogram radius, subtree position / direction boxes and stackless traversal, then the replay against the real heap.
 * TEST INFRASTRUCTURE: tests/test_host.py compares it with the oracle's LocatePhotons recursion bit for bit, so that the
 * argument "the candidate list contains the reference's inserts in order" is checked without a GPU.
 * float arithmetic: compile with -ffp-contract=off. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float position[3]; float power; uint8_t color[3]; uint8_t plane_dirz; int16_t dir_x, dir_y; } photon_t;
typedef struct { float lo[3], hi[3]; signed char dlo[3], dhi[3]; } box_t;

#define CAP 1024
#define K 100

static unsigned fbits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static float bfl(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static float dot3(float ax, float ay, float az, float bx, float by, float bz) { return (ax * bx + ay * by) + az * bz; }

/* Photon::GetDirection (cyPhotonMap.h:158-181) with its dirY-dirY slip */
static void direction_of(const photon_t *p, float *d)
{
    int ix = p->dir_x, iy = p->dir_y;
    d[0] = (float)ix / 32767.0f;
    d[1] = (float)iy / 32767.0f;
    int xy2 = ix * ix + iy - iy;
    if (xy2 > 0x3FFF0001) xy2 = 0x3FFF0001;
    int v = 0x3FFF0001 - xy2, z = 0;
    for (int b = 1 << 15; b; b >>= 1) { int t = z | b; if ((long long)t * t <= v) z = t; }
    d[2] = (float)z / 32767.0f;
    if (p->plane_dirz & 8) d[2] = -d[2];
}

typedef struct { float d2[K + 2]; int idx[K + 2]; int found; float r2; } heap_t;
static void heap_insert(heap_t *h, float dist2, int index) /* cyPhotonMap.h:384-418 */
{
    if (h->found < K) {
        h->found++;
        h->d2[h->found] = dist2; h->idx[h->found] = index;
        if (h->found == K)
            for (int k = K / 2; k >= 1; k--) {
                int parent = k; float td = h->d2[k]; int ti = h->idx[k];
                while (parent <= K / 2) {
                    int j = parent + parent;
                    if (j < K && h->d2[j] < h->d2[j + 1]) j++;
                    if (td >= h->d2[j]) break;
                    h->d2[parent] = h->d2[j]; h->idx[parent] = h->idx[j]; parent = j;
                }
                h->d2[parent] = td; h->idx[parent] = ti;
            }
    } else {
        int parent = 1, j = 2;
        while (j <= K) {
            if (j < K && h->d2[j] < h->d2[j + 1]) j++;
            if (dist2 > h->d2[j]) break;
            h->d2[parent] = h->d2[j]; h->idx[parent] = h->idx[j]; parent = j; j <<= 1;
        }
        h->d2[parent] = dist2; h->idx[parent] = index;
        h->r2 = h->d2[1];
    }
}

/* k_knn_init + k_knn_merge */
static box_t *build_boxes(const photon_t *map, int n, int half)
{
    box_t *bx = (box_t *)malloc(sizeof(box_t) * ((size_t)n + 1));
    for (int i = n; i >= 1; i--) {
        float d[3];
        direction_of(&map[i], d);
        for (int k = 0; k < 3; k++) {
            bx[i].lo[k] = bx[i].hi[k] = map[i].position[k];
            bx[i].dlo[k] = (signed char)(int)floorf(d[k] * 127.0f);
            bx[i].dhi[k] = (signed char)(int)ceilf(d[k] * 127.0f);
        }
        if (i < half)
            for (int c = 2 * i; c <= 2 * i + 1; c++)
                for (int k = 0; k < 3; k++) {
                    if (bx[c].lo[k] < bx[i].lo[k]) bx[i].lo[k] = bx[c].lo[k];
                    if (bx[c].hi[k] > bx[i].hi[k]) bx[i].hi[k] = bx[c].hi[k];
                    if (bx[c].dlo[k] < bx[i].dlo[k]) bx[i].dlo[k] = bx[c].dlo[k];
                    if (bx[c].dhi[k] > bx[i].dhi[k]) bx[i].dhi[k] = bx[c].dhi[k];
                }
    }
    return bx;
}

/* k_knn_candidates for one query; returns the list length, bit 31 = list overflow */
static unsigned candidates(const photon_t *map, const box_t *bx, int n, int half, const float *q, const float *nrm, int has_n,
                           float radius, float norm_scale, unsigned *ld, unsigned *li, long *steps)
{
    unsigned char hist[32];
    memset(hist, 0, 32);
    const float r2_0 = radius * radius;
    const int key_top = (int)(fbits(r2_0) >> 21);
    const int can_shrink = key_top >= 32 && key_top < (0x7f800000 >> 21);
    const float kcoef = norm_scale > 0.f ? (2.f * norm_scale + norm_scale * norm_scale) * 0.9999f : 0.f;
    unsigned cur = 1, up = 0, cnt = 0, below = 0, flag = 0;
    int B = 31;
    float bound = r2_0;
    if (n <= 0) return 0;
    for (;;) {
        (*steps)++;
        const photon_t *p = &map[cur];
        int process = 1;
        if (up == 0u && (int)cur < half) {
            const float ax = bx[cur].lo[0] - q[0], ay = bx[cur].lo[1] - q[1], az = bx[cur].lo[2] - q[2];
            const float hx = bx[cur].hi[0] - q[0], hy = bx[cur].hi[1] - q[1], hz = bx[cur].hi[2] - q[2];
            const float mx = fmaxf(fmaxf(ax, -hx), 0.f), my = fmaxf(fmaxf(ay, -hy), 0.f), mz = fmaxf(fmaxf(az, -hz), 0.f);
            float lb = dot3(mx, my, mz, mx, my, mz);
            int cull = 0;
            if (has_n) {
                const float t0 = nrm[0] * ax, t1 = nrm[0] * hx, t2 = nrm[1] * ay, t3 = nrm[1] * hy, t4 = nrm[2] * az, t5 = nrm[2] * hz;
                const float plo = (fminf(t0, t1) + fminf(t2, t3)) + fminf(t4, t5), phi = (fmaxf(t0, t1) + fmaxf(t2, t3)) + fmaxf(t4, t5);
                const float mag = (fmaxf(fabsf(t0), fabsf(t1)) + fmaxf(fabsf(t2), fabsf(t3))) + fmaxf(fabsf(t4), fabsf(t5));
                float pm = fmaxf(fmaxf(plo, -phi), 0.f) - 4e-7f * mag;
                pm = fmaxf(pm, 0.f);
                lb = lb + kcoef * (pm * pm);
                const float u0 = nrm[0] * (float)bx[cur].dlo[0], u1 = nrm[0] * (float)bx[cur].dhi[0];
                const float u2 = nrm[1] * (float)bx[cur].dlo[1], u3 = nrm[1] * (float)bx[cur].dhi[1];
                const float u4 = nrm[2] * (float)bx[cur].dlo[2], u5 = nrm[2] * (float)bx[cur].dhi[2];
                cull = (fminf(u0, u1) + fminf(u2, u3)) + fminf(u4, u5) > 0.02f;
            }
            cull = cull || lb > bound * 1.00001f;
            if (cull) { up = cur; cur >>= 1; process = 0; if (cur == 0u) break; continue; }
        }
        if (up != 0u || (int)cur < half) {
            const int axis = p->plane_dirz & 3;
            const float dist = q[axis] - p->position[axis];
            const unsigned near_child = dist > 0 ? 2u * cur + 1u : 2u * cur;
            if (up == 0u) { cur = near_child; process = 0; }
            else if (up == near_child && dist * dist < bound) { cur = near_child ^ 1u; up = 0u; process = 0; }
        }
        if (!process) continue;
        float fx = p->position[0] - q[0], fy = p->position[1] - q[1], fz = p->position[2] - q[2];
        float d2 = dot3(fx, fy, fz, fx, fy, fz);
        int take = d2 < bound;
        unsigned mark = 0;
        if (take && has_n) {
            float d[3];
            direction_of(p, d);
            if (dot3(d[0], d[1], d[2], nrm[0], nrm[1], nrm[2]) >= 0.f) take = 0;
            else if (norm_scale > 0.f) {
                const float perp = dot3(fx, fy, fz, nrm[0], nrm[1], nrm[2]);
                const float s = perp * norm_scale;
                fx = fx + nrm[0] * s; fy = fy + nrm[1] * s; fz = fz + nrm[2] * s;
                const float d2e = dot3(fx, fy, fz, fx, fy, fz);
                if (d2e < d2) mark = 0x80000000u;
                d2 = d2e;
                if (d2 >= bound) take = 0;
            }
        }
        if (take) {
            if (cnt < CAP) { ld[cnt] = fbits(d2); li[cnt] = cur | mark; } else flag = 1;
            int c = 31 - (key_top - (int)(fbits(d2) >> 21));
            if (c < 0) c = 0;
            if (cnt == 100u) {
                int t = 31;
                for (; t > 0; t--) if (hist[t]) break;
                hist[t]--;
                if (t < B) below--;
            }
            cnt++;
            if (c < B) { hist[c]++; below++; } else if (hist[c] < 255) hist[c]++;
            if (can_shrink && cnt > 100u)
                while (below >= 100u) { B--; below -= hist[B]; bound = bfl((unsigned)(key_top - 31 + B + 1) << 21); }
        }
        up = cur; cur >>= 1;
        if (cur == 0u) break;
    }
    return cnt | (flag << 31);
}

/* the whole estimate for nq queries; returns the number of queries whose list overflowed (they are left untouched) */
int knn_two_phase_estimate(const photon_t *map, int n, const float *pos, const float *normal, long nq, float radius, float ellipticity,
                           float *irrad, float *direction, int *found, double *mean_steps, double *mean_list)
{
    const int half = n / 2 - 1;
    const float norm_scale = ellipticity == 1.f ? 0.f : 1.f / ellipticity - 1.f;
    box_t *bx = n > 0 ? build_boxes(map, n, half) : NULL;
    unsigned *ld = (unsigned *)malloc(4 * CAP), *li = (unsigned *)malloc(4 * CAP);
    long steps = 0, lists = 0;
    int overflowed = 0;
    for (long i = 0; i < nq; i++) {
        const float *q = pos + 3 * i, *nr = normal ? normal + 3 * i : NULL;
        unsigned L = candidates(map, bx, n, half, q, nr, normal != NULL, radius, norm_scale, ld, li, &steps);
        if (L >> 31) { overflowed++; continue; }
        lists += L;
        heap_t h;
        h.found = 0; h.r2 = radius * radius;
        unsigned head = L < K ? L : K;
        for (unsigned k = 0; k < head; k++) heap_insert(&h, bfl(ld[k]), (int)(li[k] & 0x7fffffffu));
        for (unsigned k = K; k < L; k++) {
            const float d = bfl(ld[k]);
            int in = d < h.r2;
            if (in && (li[k] >> 31)) {
                const photon_t *p = &map[li[k] & 0x7fffffffu];
                const float fx = p->position[0] - q[0], fy = p->position[1] - q[1], fz = p->position[2] - q[2];
                in = dot3(fx, fy, fz, fx, fy, fz) < h.r2;
            }
            if (in) heap_insert(&h, d, (int)(li[k] & 0x7fffffffu));
        }
        float e[3] = {0, 0, 0}, o[3] = {0, 0, 0};
        for (int k = 1; k <= h.found; k++) {
            const photon_t *p = &map[h.idx[k]];
            float d[3];
            direction_of(p, d);
            for (int c = 0; c < 3; c++) e[c] = e[c] + ((float)p->color[c] / 255.0f) * p->power;
            for (int c = 0; c < 3; c++) o[c] = o[c] + d[c] * p->power;
        }
        if (h.found > 0) {
            const float area = 3.14159274101257324f * h.r2;
            if (area > 0.f) { const float inv = 1.0f / area; for (int c = 0; c < 3; c++) e[c] = e[c] * inv; }
            const float len = sqrtf(dot3(o[0], o[1], o[2], o[0], o[1], o[2]));
            for (int c = 0; c < 3; c++) o[c] = o[c] / len;
        }
        for (int c = 0; c < 3; c++) { irrad[3 * i + c] = e[c]; direction[3 * i + c] = o[c]; }
        if (found) found[i] = h.found;
    }
    if (mean_steps) *mean_steps = nq ? (double)steps / nq : 0;
    if (mean_list) *mean_list = nq ? (double)lists / nq : 0;
    free(bx); free(ld); free(li);
    return overflowed;
}
