/*
 * oracle/raytrace.c -- CPU ray/mesh intersector.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (dreammat_b200/) never does.
 *
 * Restates the closest-hit semantics the reference obtains from the un-vendored
 * third-party `_raytracing` extension (ashawkey/raytracing @ git HEAD, unpinned,
 * requirements.txt:25), as seen from its call sites:
 *   threestudio/models/renderers/raytracing_renderer.py:31,61  trace(o,d)->(pos,face_normal,depth)
 *   threestudio/models/renderers/raytracing_renderer.py:318-324 depth >= 10 => miss
 * Published algorithm of that dependency (instant-ngp triangle BVH): closest hit over a
 * binary BVH with t initialised to MAX_DIST = 10; per-triangle test is the
 * determinant form (n = e1 x e2, q = (o-a) x d, u = -q.e2/(d.n), v = q.e1/(d.n),
 * t = -n.(o-a)/(d.n); reject u<0,u>1,v<0,u+v>1,t<0), no back-face culling.
 *
 * Two independent paths are provided so the CPU tests can cross-check them:
 *   rt_trace_brute  : O(F) per ray
 *   rt_trace_bvh    : median-split BVH, ordered traversal
 * Build: gcc -O2 -fopenmp -shared -fPIC (see oracle/build.py).  -ffp-contract=off keeps
 * the arithmetic a plain IEEE fp32 sequence.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RT_MAX_DIST 10.0f

typedef struct { float x, y, z; } v3;
static inline v3 sub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline v3 cross(v3 a, v3 b) {
    v3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

/* returns t (or a huge value on miss); writes barycentrics */
static inline float tri_hit(v3 o, v3 d, v3 a, v3 b, v3 c, float *uo, float *vo) {
    v3 e1 = sub(b, a), e2 = sub(c, a), r = sub(o, a);
    v3 n = cross(e1, e2);
    v3 q = cross(r, d);
    float inv = 1.0f / dot(d, n);
    float u = inv * -dot(q, e2);
    float v = inv * dot(q, e1);
    float t = inv * -dot(n, r);
    if (!(u >= 0.0f) || u > 1.0f || !(v >= 0.0f) || (u + v) > 1.0f || !(t >= 0.0f)) return 3.0e38f;
    *uo = u; *vo = v;
    return t;
}

/* ------------------------------------------------------------------ brute force */
void rt_trace_brute(const float *verts, const int32_t *tris, int64_t n_tris,
                    const float *ro, const float *rd, int64_t n_rays,
                    float *out_t, int32_t *out_tri, float *out_uv) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_rays; ++i) {
        v3 o = {ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]};
        v3 d = {rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]};
        float best = RT_MAX_DIST, bu = 0, bv = 0; int32_t bi = -1;
        for (int64_t f = 0; f < n_tris; ++f) {
            const float *pa = verts + 3 * tris[3 * f], *pb = verts + 3 * tris[3 * f + 1], *pc = verts + 3 * tris[3 * f + 2];
            v3 a = {pa[0], pa[1], pa[2]}, b = {pb[0], pb[1], pb[2]}, c = {pc[0], pc[1], pc[2]};
            float u, v, t = tri_hit(o, d, a, b, c, &u, &v);
            if (t < best) { best = t; bi = (int32_t)f; bu = u; bv = v; }
        }
        out_t[i] = best; out_tri[i] = bi;
        if (out_uv) { out_uv[2 * i] = bu; out_uv[2 * i + 1] = bv; }
    }
}

/* ------------------------------------------------------------------ BVH */
typedef struct {
    float lo[3], hi[3];
    int32_t left, right;  /* children (internal) */
    int32_t first, count; /* leaf range in perm (count>0 => leaf) */
} Node;

typedef struct {
    Node *nodes; int32_t n_nodes;
    int32_t *perm;
    const float *verts; const int32_t *tris; int64_t n_tris;
    float *cent;
} Bvh;

static void tri_bounds(const Bvh *b, int32_t f, float lo[3], float hi[3]) {
    for (int k = 0; k < 3; ++k) { lo[k] = 3e38f; hi[k] = -3e38f; }
    for (int j = 0; j < 3; ++j) {
        const float *p = b->verts + 3 * b->tris[3 * f + j];
        for (int k = 0; k < 3; ++k) { if (p[k] < lo[k]) lo[k] = p[k]; if (p[k] > hi[k]) hi[k] = p[k]; }
    }
}

static int g_axis; static const float *g_cent;
static int cmp_axis(const void *a, const void *b) {
    float ca = g_cent[3 * (*(const int32_t *)a) + g_axis], cb = g_cent[3 * (*(const int32_t *)b) + g_axis];
    return (ca > cb) - (ca < cb);
}

static int32_t build_rec(Bvh *b, int32_t first, int32_t count) {
    int32_t id = b->n_nodes++;
    Node *n = &b->nodes[id];
    for (int k = 0; k < 3; ++k) { n->lo[k] = 3e38f; n->hi[k] = -3e38f; }
    float clo[3] = {3e38f, 3e38f, 3e38f}, chi[3] = {-3e38f, -3e38f, -3e38f};
    for (int32_t i = first; i < first + count; ++i) {
        float lo[3], hi[3]; tri_bounds(b, b->perm[i], lo, hi);
        for (int k = 0; k < 3; ++k) {
            if (lo[k] < n->lo[k]) n->lo[k] = lo[k];
            if (hi[k] > n->hi[k]) n->hi[k] = hi[k];
            float c = b->cent[3 * b->perm[i] + k];
            if (c < clo[k]) clo[k] = c; if (c > chi[k]) chi[k] = c;
        }
    }
    if (count <= 4) { n->first = first; n->count = count; n->left = n->right = -1; return id; }
    int axis = 0; float ext = chi[0] - clo[0];
    for (int k = 1; k < 3; ++k) if (chi[k] - clo[k] > ext) { ext = chi[k] - clo[k]; axis = k; }
    g_axis = axis; g_cent = b->cent;
    qsort(b->perm + first, (size_t)count, sizeof(int32_t), cmp_axis);
    int32_t half = count / 2;
    n->count = 0; n->first = 0;
    int32_t l = build_rec(b, first, half);
    int32_t r = build_rec(b, first + half, count - half);
    b->nodes[id].left = l; b->nodes[id].right = r;
    return id;
}

void *rt_bvh_build(const float *verts, const int32_t *tris, int64_t n_tris) {
    Bvh *b = (Bvh *)calloc(1, sizeof(Bvh));
    b->verts = verts; b->tris = tris; b->n_tris = n_tris;
    b->nodes = (Node *)malloc(sizeof(Node) * (size_t)(2 * n_tris + 1));
    b->perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_tris);
    b->cent = (float *)malloc(sizeof(float) * 3 * (size_t)n_tris);
    for (int64_t f = 0; f < n_tris; ++f) {
        b->perm[f] = (int32_t)f;
        for (int k = 0; k < 3; ++k)
            b->cent[3 * f + k] = (verts[3 * tris[3 * f] + k] + verts[3 * tris[3 * f + 1] + k] + verts[3 * tris[3 * f + 2] + k]) / 3.0f;
    }
    build_rec(b, 0, (int32_t)n_tris);
    return b;
}

void rt_bvh_free(void *h) {
    Bvh *b = (Bvh *)h; if (!b) return;
    free(b->nodes); free(b->perm); free(b->cent); free(b);
}

static inline int box_hit(const Node *n, v3 o, v3 inv, float tmax, float *tnear) {
    float t0 = 0.0f, t1 = tmax;
    const float oo[3] = {o.x, o.y, o.z}, ii[3] = {inv.x, inv.y, inv.z};
    for (int k = 0; k < 3; ++k) {
        /* slightly widened slab so a hit exactly on a box face is never culled */
        float a = (n->lo[k] - 1e-6f - oo[k]) * ii[k], b = (n->hi[k] + 1e-6f - oo[k]) * ii[k];
        float lo = a < b ? a : b, hi = a < b ? b : a;
        if (lo != lo || hi != hi) continue; /* 0 * inf: ray parallel inside slab plane */
        if (lo > t0) t0 = lo; if (hi < t1) t1 = hi;
    }
    *tnear = t0;
    return t0 <= t1;
}

void rt_trace_bvh(const void *h, const float *ro, const float *rd, int64_t n_rays,
                  float *out_t, int32_t *out_tri, float *out_uv) {
    const Bvh *b = (const Bvh *)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n_rays; ++i) {
        v3 o = {ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]};
        v3 d = {rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]};
        v3 inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
        float best = RT_MAX_DIST, bu = 0, bv = 0; int32_t bi = -1;
        int32_t stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node *n = &b->nodes[stack[--sp]];
            float tn;
            if (!box_hit(n, o, inv, best, &tn)) continue;
            if (n->count > 0) {
                for (int32_t k = n->first; k < n->first + n->count; ++k) {
                    int32_t f = b->perm[k];
                    const float *pa = b->verts + 3 * b->tris[3 * f], *pb = b->verts + 3 * b->tris[3 * f + 1], *pc = b->verts + 3 * b->tris[3 * f + 2];
                    v3 a = {pa[0], pa[1], pa[2]}, bb = {pb[0], pb[1], pb[2]}, c = {pc[0], pc[1], pc[2]};
                    float u, v, t = tri_hit(o, d, a, bb, c, &u, &v);
                    /* ties resolved to the lowest face index, like the brute-force scan */
                    if (t < best || (t == best && bi >= 0 && f < bi)) { best = t; bi = f; bu = u; bv = v; }
                }
            } else {
                stack[sp++] = n->left; stack[sp++] = n->right;
            }
        }
        out_t[i] = best; out_tri[i] = bi;
        if (out_uv) { out_uv[2 * i] = bu; out_uv[2 * i + 1] = bv; }
    }
}
