// Per-sample mesh queries: exact closest triangle, +x ray parity (inside test), nearest vertex, per-view sample
// visibility.  Replaces cal_vis_sdf_batch (src/lib/dataset/mesh_util.py:498-524: kaolin point_to_mesh_distance +
// check_sign + barycentric blend) and pytorch3d knn_points(K=1) (src/networks.py:28), which the reference issues
// 2 + 3 times per pass by brute force (N x F).  Here: one BVH traversal each, with conservative pruning so that
// the result is identical to the brute-force first-minimum of oracle/geom_oracle.c (ties -> lowest index).
#pragma once
#include <cstdio>
#include "common.cuh"
#include "rays.cuh"

#define BVH_STACK 48
#ifdef GEOM_COUNT            // developer build: traversal statistics (printed by k_geom_print after every query launch)
__device__ unsigned long long d_geom_cnt[8];   // samples, tri nodes popped, past box, past slab, tri slab tests, exact tests, vtx nodes, parity tris
#define GC(i, n) cnt[i] += (n)
#else
#define GC(i, n) ((void)0)
#endif
#ifndef GEOM_RAY_LANES
#define GEOM_RAY_LANES 8       // rays per warp (sweep: 32 -> 30.0 ms, 16 -> 26.4, 8 -> 25.5, 4 -> 26.4; 0 = old mapping, 32 depths of one ray: 35.0)
#endif
#ifndef GEOM_NODE_LB
#define GEOM_NODE_LB 1       // per-node slab bound (plane + thickness + bounding circle) after the box test
#endif
#ifndef GEOM_TRI_LB
#define GEOM_TRI_LB 1        // per-triangle lower bound (plane distance + bounding circle) before the exact distance
#endif

// ab = b - a and ac = c - a are passed in (precomputed per frame with the same rounding as xsub)
__device__ __forceinline__ float point_tri_dist2(const float* p, const float* a, const float* b, const float* c,
                                                 float ab0, float ab1, float ab2, float ac0, float ac1, float ac2) {
    const float ap0 = xsub(p[0], a[0]), ap1 = xsub(p[1], a[1]), ap2 = xsub(p[2], a[2]);
    float q0, q1, q2;
    const float d1 = xdot3(ab0, ab1, ab2, ap0, ap1, ap2), d2 = xdot3(ac0, ac1, ac2, ap0, ap1, ap2);
    bool done = false;
    if (d1 <= 0.0f && d2 <= 0.0f) { q0 = a[0]; q1 = a[1]; q2 = a[2]; done = true; }
    float d3 = 0, d4 = 0, d5 = 0, d6 = 0, vc = 0, vb = 0;
    if (!done) {
        const float bp0 = xsub(p[0], b[0]), bp1 = xsub(p[1], b[1]), bp2 = xsub(p[2], b[2]);
        d3 = xdot3(ab0, ab1, ab2, bp0, bp1, bp2);
        d4 = xdot3(ac0, ac1, ac2, bp0, bp1, bp2);
        if (d3 >= 0.0f && d4 <= d3) { q0 = b[0]; q1 = b[1]; q2 = b[2]; done = true; }
    }
    if (!done) {
        vc = xsub(xmul(d1, d4), xmul(d3, d2));
        if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
            const float v = xdiv(d1, xsub(d1, d3));
            q0 = xadd(a[0], xmul(v, ab0)); q1 = xadd(a[1], xmul(v, ab1)); q2 = xadd(a[2], xmul(v, ab2));
            done = true;
        }
    }
    if (!done) {
        const float cp0 = xsub(p[0], c[0]), cp1 = xsub(p[1], c[1]), cp2 = xsub(p[2], c[2]);
        d5 = xdot3(ab0, ab1, ab2, cp0, cp1, cp2);
        d6 = xdot3(ac0, ac1, ac2, cp0, cp1, cp2);
        if (d6 >= 0.0f && d5 <= d6) { q0 = c[0]; q1 = c[1]; q2 = c[2]; done = true; }
    }
    if (!done) {
        vb = xsub(xmul(d5, d2), xmul(d1, d6));
        if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
            const float w = xdiv(d2, xsub(d2, d6));
            q0 = xadd(a[0], xmul(w, ac0)); q1 = xadd(a[1], xmul(w, ac1)); q2 = xadd(a[2], xmul(w, ac2));
            done = true;
        }
    }
    if (!done) {
        const float va = xsub(xmul(d3, d6), xmul(d5, d4));
        const float e43 = xsub(d4, d3), e56 = xsub(d5, d6);
        if (va <= 0.0f && e43 >= 0.0f && e56 >= 0.0f) {
            const float w = xdiv(e43, xadd(e43, e56));
            q0 = xadd(b[0], xmul(w, xsub(c[0], b[0])));
            q1 = xadd(b[1], xmul(w, xsub(c[1], b[1])));
            q2 = xadd(b[2], xmul(w, xsub(c[2], b[2])));
        } else {
            const float denom = xdiv(1.0f, xadd(xadd(va, vb), vc));
            const float v = xmul(vb, denom), w = xmul(vc, denom);
            q0 = xadd(xadd(a[0], xmul(ab0, v)), xmul(ac0, w));
            q1 = xadd(xadd(a[1], xmul(ab1, v)), xmul(ac1, w));
            q2 = xadd(xadd(a[2], xmul(ab2, v)), xmul(ac2, w));
        }
    }
    const float e0 = xsub(p[0], q0), e1 = xsub(p[1], q1), e2 = xsub(p[2], q2);
    return xdot3(e0, e1, e2, e0, e1, e2);
}

// squared distance from p to an AABB (lower bound of the distance to anything inside); plain arithmetic is fine,
// pruning is made conservative by the caller
__device__ __forceinline__ float aabb_dist2(const float4& mn, const float4& mx, const float* p) {
    const float dx = fmaxf(fmaxf(mn.x - p[0], 0.0f), p[0] - mx.x);
    const float dy = fmaxf(fmaxf(mn.y - p[1], 0.0f), p[1] - mx.y);
    const float dz = fmaxf(fmaxf(mn.z - p[2], 0.0f), p[2] - mx.z);
    return dx * dx + dy * dy + dz * dz;
}

// bound_d: an upper bound of the result (squared distance to any mesh vertex, slightly inflated by the caller) or +inf.
// It only prunes: the first-minimum face is still found and its distance is the one point_tri_dist2 computes.
__device__ __forceinline__ void closest_face(const FrameDev& fr, const float* p, float bound_d, float& best_d, int& best_f, unsigned* cnt) {
    (void)cnt;
    int stack[BVH_STACK];
    int sp = 0;
    stack[sp++] = 0;
    best_d = bound_d;
    best_f = 0x7fffffff;
    while (sp > 0) {
        const int ni = stack[--sp];
        const float4 mn = fr.tri_nodes[2 * ni], mx = fr.tri_nodes[2 * ni + 1];
        GC(1, 1);
        if (aabb_dist2(mn, mx, p) > best_d * 1.00001f + 1e-12f) continue;
        GC(2, 1);
#if GEOM_NODE_LB
        {   // slab bound of the node (bvh::triangle_node_bounds): for a smooth patch far from p it is much tighter than the
            // box, whose slack lets every node within ~sqrt(2 D size) of the foot point through
            const float4 l0 = fr.tri_node_lb[2 * ni], l1 = fr.tri_node_lb[2 * ni + 1];
            const float ex = p[0] - l1.x, ey = p[1] - l1.y, ez = p[2] - l1.z;
            const float h = l0.x * ex + l0.y * ey + l0.z * ez;
            const float e2 = ex * ex + ey * ey + ez * ez;
            const float u = fmaxf(fabsf(h) - l0.w, 0.0f), s = fmaxf(sqrtf(fmaxf(e2 - h * h, 0.0f)) - l1.w, 0.0f);
            if (u * u + s * s > best_d * 1.0001f + 1e-10f) continue;
        }
#endif
        GC(3, 1);
        const int a = __float_as_int(mn.w), b = __float_as_int(mx.w);
        if (a < 0) {
            const int first = ~a;
            for (int i = 0; i < b; ++i) {
                const float4* rec = fr.tri_rec + TRI_REC_F4 * (first + i);
                GC(4, 1);
#if GEOM_TRI_LB
                {   // Lower bound of the distance to the triangle: it lies in its plane (unit normal n, through c) inside
                    // the circle (c, r), so dist^2 >= h^2 + max(0, sqrt(|p - c|^2 - h^2) - r)^2 with h = n . (p - c).
                    // An AABB bound lets every triangle within ~sqrt(2 D e) of the foot point through (D = distance,
                    // e = edge length: ~60 exact tests per far sample); this one passes the few whose circle covers it.
                    // Only a prune (r is inflated on the host, slack 1e-4 relative): the result stays the exact first minimum.
                    const float4 l0 = rec[4], l1 = rec[5];
                    const float ex = p[0] - l1.x, ey = p[1] - l1.y, ez = p[2] - l1.z;
                    const float h = l0.x * ex + l0.y * ey + l0.z * ez;
                    const float e2 = ex * ex + ey * ey + ez * ez;
                    const float s = fmaxf(sqrtf(fmaxf(e2 - h * h, 0.0f)) - l0.w, 0.0f);
                    if (h * h + s * s > best_d * 1.0001f + 1e-10f) continue;
                }
#endif
                const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2];
                const float4 r3 = rec[3];
                GC(5, 1);
                const int f = __float_as_int(r0.w);
                const float va[3] = {r0.x, r0.y, r0.z}, vb[3] = {r1.x, r1.y, r1.z}, vc[3] = {r2.x, r2.y, r2.z};
                const float d = point_tri_dist2(p, va, vb, vc, r1.w, r2.w, r3.x, r3.y, r3.z, r3.w);
                if (d < best_d || (d == best_d && f < best_f)) { best_d = d; best_f = f; }
            }
        } else {
            // near child last (popped first); a child that cannot beat the current best is not pushed at all
            const float da = aabb_dist2(fr.tri_nodes[2 * a], fr.tri_nodes[2 * a + 1], p);
            const float db = aabb_dist2(fr.tri_nodes[2 * b], fr.tri_nodes[2 * b + 1], p);
            const float thr = best_d * 1.00001f + 1e-12f;
            if (da < db) { if (db <= thr) stack[sp++] = b; if (da <= thr) stack[sp++] = a; }
            else { if (da <= thr) stack[sp++] = a; if (db <= thr) stack[sp++] = b; }
        }
    }
}

// parity of +x ray crossings (Moller-Trumbore, dir = (1,0,0)); same per-triangle arithmetic as vo_check_sign
__device__ __forceinline__ bool inside_parity(const FrameDev& fr, const float* p) {
    int stack[BVH_STACK];
    int sp = 0;
    stack[sp++] = 0;
    int cnt = 0;
    const float m = 1e-4f;      // metres; fp32 rounding of the hit test is ~1e-8 at this scale
    while (sp > 0) {
        const int ni = stack[--sp];
        const float4 mn = fr.tri_nodes[2 * ni], mx = fr.tri_nodes[2 * ni + 1];
        if (p[1] < mn.y - m || p[1] > mx.y + m || p[2] < mn.z - m || p[2] > mx.z + m || p[0] > mx.x + m) continue;
        const int a = __float_as_int(mn.w), b = __float_as_int(mx.w);
        if (a < 0) {
            const int first = ~a;
            for (int i = 0; i < b; ++i) {
                const float4* rec = fr.tri_rec + TRI_REC_F4 * (first + i);
                const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3];
                const float v0[3] = {r0.x, r0.y, r0.z};
                const float e10 = r1.w, e11 = r2.w, e12 = r3.x;        // v1 - v0
                const float e20 = r3.y, e21 = r3.z, e22 = r3.w;        // v2 - v0
                const float aa = xsub(xmul(e12, e21), xmul(e11, e22));
                if (fabsf(aa) < 1e-20f) continue;
                const float inv = xdiv(1.0f, aa);
                const float s0 = xsub(p[0], v0[0]), s1 = xsub(p[1], v0[1]), s2 = xsub(p[2], v0[2]);
                const float u = xmul(inv, xsub(xmul(s2, e21), xmul(s1, e22)));
                if (!(u >= 0.0f)) continue;
                const float qx = xsub(xmul(s1, e12), xmul(s2, e11));
                const float qy = xsub(xmul(s2, e10), xmul(s0, e12));
                const float qz = xsub(xmul(s0, e11), xmul(s1, e10));
                const float v = xmul(inv, qx);
                if (!(v >= 0.0f) || !(xadd(u, v) <= 1.0f)) continue;
                const float t = xmul(inv, xdot3(e20, e21, e22, qx, qy, qz));
                if (t > 0.0f) cnt++;
            }
        } else {
            stack[sp++] = a;
            stack[sp++] = b;
        }
    }
    return (cnt & 1) != 0;
}

__device__ __forceinline__ int nearest_vertex(const FrameDev& fr, const float* p, float& best_d_out) {
    int stack[BVH_STACK];
    int sp = 0;
    stack[sp++] = 0;
    float best_d = __int_as_float(0x7f800000);
    int best_i = 0x7fffffff;
    while (sp > 0) {
        const int ni = stack[--sp];
        const float4 mn = fr.vtx_nodes[2 * ni], mx = fr.vtx_nodes[2 * ni + 1];
        if (aabb_dist2(mn, mx, p) > best_d * 1.00001f + 1e-12f) continue;
        const int a = __float_as_int(mn.w), b = __float_as_int(mx.w);
        if (a < 0) {
            const int first = ~a;
            for (int i = 0; i < b; ++i) {
                const float4 vr = fr.vtx_rec[first + i];
                const int j = __float_as_int(vr.w);
                const float dx = xsub(p[0], vr.x), dy = xsub(p[1], vr.y), dz = xsub(p[2], vr.z);
                const float d = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
                if (d < best_d || (d == best_d && j < best_i)) { best_d = d; best_i = j; }
            }
        } else {
            const float da = aabb_dist2(fr.vtx_nodes[2 * a], fr.vtx_nodes[2 * a + 1], p);
            const float db = aabb_dist2(fr.vtx_nodes[2 * b], fr.vtx_nodes[2 * b + 1], p);
            if (da < db) { stack[sp++] = b; stack[sp++] = a; }
            else { stack[sp++] = a; stack[sp++] = b; }
        }
    }
    best_d_out = best_d;
    return best_i;
}

// barycentric_coordinates_of_projection (mesh_util.py:321-356) -> b0,b1,b2 on face f
__device__ __forceinline__ void bary_of_projection(const FrameDev& fr, const float* p, int f, float* bw) {
    const int* fv = fr.faces + 3 * f;
    const float* v0 = fr.verts + 3 * fv[0];
    const float* v1 = fr.verts + 3 * fv[1];
    const float* v2 = fr.verts + 3 * fv[2];
    const float u0 = xsub(v1[0], v0[0]), u1 = xsub(v1[1], v0[1]), u2 = xsub(v1[2], v0[2]);
    const float w0 = xsub(v2[0], v0[0]), w1 = xsub(v2[1], v0[1]), w2 = xsub(v2[2], v0[2]);       // "v" in the reference
    const float n0 = xsub(xmul(u1, w2), xmul(u2, w1)), n1 = xsub(xmul(u2, w0), xmul(u0, w2)), n2 = xsub(xmul(u0, w1), xmul(u1, w0));
    float s = xadd(xadd(xmul(n0, n0), xmul(n1, n1)), xmul(n2, n2));
    if (s == 0.0f) s = 1e-6f;
    const float inv = xdiv(1.0f, s);
    const float q0 = xsub(p[0], v0[0]), q1 = xsub(p[1], v0[1]), q2 = xsub(p[2], v0[2]);          // "w" in the reference
    // cross(u, q)
    const float c0 = xsub(xmul(u1, q2), xmul(u2, q1)), c1 = xsub(xmul(u2, q0), xmul(u0, q2)), c2 = xsub(xmul(u0, q1), xmul(u1, q0));
    const float b2 = xmul(xadd(xadd(xmul(c0, n0), xmul(c1, n1)), xmul(c2, n2)), inv);
    // cross(q, v)
    const float g0 = xsub(xmul(q1, w2), xmul(q2, w1)), g1 = xsub(xmul(q2, w0), xmul(q0, w2)), g2 = xsub(xmul(q0, w1), xmul(q1, w0));
    const float b1 = xmul(xadd(xadd(xmul(g0, n0), xmul(g1, n1)), xmul(g2, n2)), inv);
    bw[0] = xsub(xsub(1.0f, b1), b2);
    bw[1] = b1;
    bw[2] = b2;
}

// One thread per sample.  Outputs may be NULL.
// pts_in != NULL: explicit query points (VANeRF.query called directly) instead of cam_pos + dir * z.
#ifndef GEOM_MINB
#define GEOM_MINB 12       // 40 registers: 48 warps per SM hide the dependent node loads (sweep: 8 -> 36.7 ms, 12 -> 36.0, 16 -> 43.3)
#endif
__global__ void __launch_bounds__(128, GEOM_MINB) k_geom_query(FrameDev fr, TargetDev tar, const float* __restrict__ rays, const float* __restrict__ z,
                             const float* __restrict__ pts_in, int R, int S, float* __restrict__ pts, float* __restrict__ sdf,
                             int* __restrict__ face, int* __restrict__ nn_vert, unsigned char* __restrict__ qvis) {
    const long long N = (long long)R * S;
    long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#if GEOM_RAY_LANES
    // Thread -> sample mapping for ray batches: the 32 lanes of a warp take the SAME depth index of 32 consecutive rays
    // (neighbouring pixels: a few millimetres apart in space), consecutive warps the next depth index.  With 32 consecutive
    // depths of one ray per warp the lanes were up to 11 cm apart and every lane walked its own BVH path; now the warp
    // walks (nearly) one path.  Results do not depend on the mapping; outputs stay indexed by n = ray * S + depth.
    if (!pts_in) {
        // a warp = GEOM_RAY_LANES consecutive rays x (32 / GEOM_RAY_LANES) consecutive depth indices
        constexpr int RL = GEOM_RAY_LANES, DL = 32 / RL;
        const long long gw = n >> 5;
        const int lane = (int)(n & 31);
        const int sb_per_ray_block = (S + DL - 1) / DL;
        const int rb = (int)(gw / sb_per_ray_block), sb = (int)(gw - (long long)rb * sb_per_ray_block);
        const int rr = rb * RL + (lane % RL), si = sb * DL + lane / RL;
        if (rr >= R || si >= S) return;
        n = (long long)rr * S + si;
    }
#endif
    if (n >= N) return;
    const int r = (int)(n / S);
    float p[3];
    if (pts_in) { p[0] = pts_in[3 * n]; p[1] = pts_in[3 * n + 1]; p[2] = pts_in[3 * n + 2]; }
    else sample_point(rays + (size_t)r * VANERF_RAY_STRIDE, tar.cam_pos, z[n], p);
    if (pts) { pts[3 * n] = p[0]; pts[3 * n + 1] = p[1]; pts[3 * n + 2] = p[2]; }
    // nearest vertex first: its squared distance bounds the closest-face distance from above (the vertex belongs to a
    // face), which prunes most of the triangle traversal.  Inflated by 1e-4 relative so that rounding differences
    // between the two distance formulas (~1e-6) cannot exclude the true closest face.
#ifndef GEOM_ABLATE
#define GEOM_ABLATE 0        // developer timing experiments (wrong results): 1 no nearest vertex, 2 no closest face, 4 no parity
#endif
    unsigned cnt[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    (void)cnt;
    float dnn2 = 1e-4f;
    int nnv = 0;
    if (!(GEOM_ABLATE & 1)) nnv = nearest_vertex(fr, p, dnn2);
    float d2 = 1e-4f; int f = 0;
    if (!(GEOM_ABLATE & 2)) {
        closest_face(fr, p, dnn2 * 1.0001f + 1e-12f, d2, f, cnt);
        if (f == 0x7fffffff) closest_face(fr, p, __int_as_float(0x7f800000), d2, f, cnt);     // never expected; keeps the result exact
    }
    const bool in = (GEOM_ABLATE & 4) ? false : inside_parity(fr, p);
    // pts_sdf = sqrt(d2 + 1e-6) * (-2 * (inside - 0.5))   (mesh_util.py:510-512)
    const float dist = xsqrt(xadd(d2, 1e-6f));
    const float sign = xmul(-2.0f, xsub(in ? 1.0f : 0.0f, 0.5f));
    if (sdf) sdf[n] = xmul(dist, sign);
    if (face) face[n] = f;
    if (nn_vert) nn_vert[n] = nnv;
    if (qvis) {
        float bw[3];
        bary_of_projection(fr, p, f, bw);
        const int* fv = fr.faces + 3 * f;
        for (int v = 0; v < fr.V; ++v) {
            const float* vis = fr.vis + (size_t)v * fr.n_verts;
            const float blend = xadd(xadd(xmul(vis[fv[0]], bw[0]), xmul(vis[fv[1]], bw[1])), xmul(vis[fv[2]], bw[2]));
            qvis[(size_t)v * N + n] = blend >= 1e-1f ? 1 : 0;
        }
    }
#ifdef GEOM_COUNT
    for (int i = 0; i < 8; ++i) if (cnt[i]) atomicAdd(&d_geom_cnt[i], (unsigned long long)cnt[i]);
#endif
}
#ifdef GEOM_COUNT
__global__ void k_geom_print() {
    const double n = (double)d_geom_cnt[0];
    printf("GEOM_COUNT samples %llu per sample: tri nodes popped %.1f past box %.1f past slab %.1f tri slab tests %.1f exact tests %.1f\n",
           d_geom_cnt[0], d_geom_cnt[1] / n, d_geom_cnt[2] / n, d_geom_cnt[3] / n, d_geom_cnt[4] / n, d_geom_cnt[5] / n);
}
#endif
