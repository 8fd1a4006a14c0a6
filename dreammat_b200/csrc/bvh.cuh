// bvh.cuh -- device-side BVH2 traversal (closest-hit and any-hit) over the layout built in bvh.cu.
//
// Replaces the `_raytracing` extension the reference calls at
// models/renderers/raytracing_renderer.py:31,61 (semantics: closest hit with t initialised to
// 10; `depth >= 10` is a miss, :322).  Triangle test = the determinant form the published
// instant-ngp / ashawkey-raytracing BVH uses (no culling, u,v in [0,1], u+v<=1, t>=0).
//
// Node = 4 x float4 (64 B): both child boxes live in the parent, so one node fetch (two 32 B
// sectors) decides both children.  Leaves hold <= 4 triangles stored as 3 x float4 (a|id, b, c).
#pragma once
#include "common.cuh"

#define DM_RT_MAX_DIST 10.0f
#define DM_BVH_STACK 48

struct dm_bvh {
    float4* nodes;   // [n_nodes*4]
    float4* tris;    // [n_tris*3]   leaf order
    int32_t n_nodes;
    int32_t n_tris;
    int32_t root;    // child code of the root (>=0 node, <0 leaf)
};

struct BvhView {
    const float4* __restrict__ nodes;
    const float4* __restrict__ tris;
    int32_t root;
};

__device__ __forceinline__ float tri_hit_dev(f3 o, f3 d, f3 a, f3 b, f3 c, float& u, float& v) {
    f3 e1 = b - a, e2 = c - a, r = o - a;
    f3 n = cross3(e1, e2);
    f3 q = cross3(r, d);
    float inv = 1.0f / dot3(d, n);
    u = inv * -dot3(q, e2);
    v = inv * dot3(q, e1);
    float t = inv * -dot3(n, r);
    if (!(u >= 0.0f) || u > 1.0f || !(v >= 0.0f) || (u + v) > 1.0f || !(t >= 0.0f)) return 3.0e38f;
    return t;
}

__device__ __forceinline__ bool slab(float lox, float loy, float loz, float hix, float hiy, float hiz, f3 o, f3 inv,
                                     float tmax, float& tn) {
    float ax = (lox - 1e-6f - o.x) * inv.x, bx = (hix + 1e-6f - o.x) * inv.x;
    float ay = (loy - 1e-6f - o.y) * inv.y, by = (hiy + 1e-6f - o.y) * inv.y;
    float az = (loz - 1e-6f - o.z) * inv.z, bz = (hiz + 1e-6f - o.z) * inv.z;
    // fminf/fmaxf drop NaNs (0*inf when the ray lies in a slab plane)
    float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    tn = t0;
    return t0 <= t1;
}

// ANY=true : returns as soon as some triangle is hit with t < tmax (occlusion rays)
// ANY=false: closest hit; ties go to the lowest original triangle id (matches the oracle's scan order)
template <bool ANY>
__device__ __forceinline__ bool bvh_trace(const BvhView& bv, f3 o, f3 d, float& best_t, int& best_id, float& bu,
                                          float& bvv) {
    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[DM_BVH_STACK];
    int sp = 0;
    int cur = bv.root;
    best_t = DM_RT_MAX_DIST;
    best_id = -1;
    bu = bvv = 0.0f;
    while (true) {
        if (cur >= 0) {
            const float4* n = bv.nodes + (int64_t)cur * 4;
            float4 n0 = __ldg(n), n1 = __ldg(n + 1), n2 = __ldg(n + 2), n3 = __ldg(n + 3);
            float tl, tr;
            bool hl = slab(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, o, inv, best_t, tl);
            bool hr = slab(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, o, inv, best_t, tr);
            int cl = __float_as_int(n3.x), cr = __float_as_int(n3.y);
            if (hl && hr) {
                // near child first
                bool lfirst = tl <= tr;
                int nearc = lfirst ? cl : cr, farc = lfirst ? cr : cl;
                if (sp < DM_BVH_STACK) stack[sp++] = farc;
                cur = nearc;
                continue;
            } else if (hl) {
                cur = cl;
                continue;
            } else if (hr) {
                cur = cr;
                continue;
            }
        } else {
            int code = ~cur;
            int first = code >> 2, cnt = (code & 3) + 1;
            for (int k = 0; k < cnt; ++k) {
                const float4* tp = bv.tris + (int64_t)(first + k) * 3;
                float4 A = __ldg(tp), B = __ldg(tp + 1), C = __ldg(tp + 2);
                float u, v;
                float t = tri_hit_dev(o, d, mk3(A.x, A.y, A.z), mk3(B.x, B.y, B.z), mk3(C.x, C.y, C.z), u, v);
                int id = __float_as_int(A.w);
                if (ANY) {
                    if (t < best_t) { best_t = t; best_id = id; return true; }
                } else {
                    if (t < best_t || (t == best_t && best_id >= 0 && id < best_id)) {
                        best_t = t; best_id = id; bu = u; bvv = v;
                    }
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return best_id >= 0;
}
