// Device-side construction of the per-frame acceleration structures: the triangle BVH and the vertex BVH
// the mesh queries of geom.cuh traverse, the per-primitive leaf records and the per-node slab bounds.  The reference has no
// counterpart (kaolin / pytorch3d answer cal_vis_sdf_batch and knn_points by brute force, mesh_util.py:498-524,
// networks.py:28); any valid tree gives the same query results because the searches are exact (first minimum, ties ->
// lowest index), so the builder is free to choose its own splits.
//
// Tree = median split on the longest axis of the centroid bounds, top down, one level per iteration, one CTA per tree
// (3108 triangles / 1558 vertices: the whole build is a few dozen microseconds and needs no host round trip).  The split of
// a node is the exact median of its contiguous range in the order (centroid coordinate, primitive index) - one shared-memory
// bitonic sort of all primitives per level, range start as the leading key field - so the node ranges are a deterministic
// function of the mesh.
// Node = 2 x float4 {min.xyz, as_float(a)}, {max.xyz, as_float(b)}: inner a / b = children, leaf a = ~first primitive
// (negative), b = count; primitives are stored in leaf order.
#pragma once
#include "common.cuh"

#define BVH_MAX_NODES 2048
#define BVH_BUILD_THREADS 1024
#define BVH_MAX_PRIMS 4096              // primitives per tree (sort keys are staged in shared memory)
#define BVH_SLAB_MAX_PRIMS 256          // nodes with more triangles get no slab bound (their slabs are too thick to prune anything)

struct BvhBuildArgs {
    const float* verts;        // (Nv,3)
    const int* faces;          // (n,3) for the triangle tree, NULL for the vertex tree (primitive i = vertex i)
    int n, leaf;
    float4* nodes;             // out: 2 float4 per node
    int* prims;                // out: primitive ids in leaf order
    int2* node_range;          // out: (first, count) of every node
    int* n_nodes;              // out
    // scratch, context-owned: box 6n floats, cen 3n floats, key n floats (unused since the keys moved to shared memory), prim_b n, node_a n, node_b n ints,
    // nb 12 * BVH_MAX_NODES ints, axis BVH_MAX_NODES ints, active 2 * BVH_MAX_NODES ints
    float* box; float* cen; float* key;
    int* prim_b; int* node_a; int* node_b; int* nb; int* axis; int* active;
};

// order-preserving map float -> int (for atomicMin / atomicMax on floats)
__device__ __forceinline__ int bvh_f2o(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float bvh_o2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

__global__ void __launch_bounds__(BVH_BUILD_THREADS) k_bvh_build(BvhBuildArgs A0, BvhBuildArgs A1) {
    const BvhBuildArgs& A = blockIdx.x ? A1 : A0;
    const int tid = threadIdx.x, nt = blockDim.x, n = A.n;
    __shared__ int s_nodes, s_active, s_next;
    __shared__ int s_scan[32];
    __shared__ unsigned long long s_sort[BVH_MAX_PRIMS];     // per level: (range start, ordered centroid coordinate, primitive id)
    int np2 = 1;
    while (np2 < n) np2 <<= 1;                               // bitonic network size
    // ---- primitive boxes and centroids
    for (int i = tid; i < n; i += nt) {
        float mn[3], mx[3];
        if (A.faces) {
            const float* a = A.verts + 3 * A.faces[3 * i];
            const float* b = A.verts + 3 * A.faces[3 * i + 1];
            const float* c = A.verts + 3 * A.faces[3 * i + 2];
            for (int k = 0; k < 3; ++k) { mn[k] = fminf(a[k], fminf(b[k], c[k])); mx[k] = fmaxf(a[k], fmaxf(b[k], c[k])); }
        } else {
            for (int k = 0; k < 3; ++k) mn[k] = mx[k] = A.verts[3 * i + k];
        }
        for (int k = 0; k < 3; ++k) {
            A.box[6 * i + k] = mn[k]; A.box[6 * i + 3 + k] = mx[k];
            A.cen[3 * i + k] = xmul(0.5f, xadd(mn[k], mx[k]));
        }
        A.prims[i] = i;
        A.node_a[i] = 0;
    }
    if (tid == 0) {
        s_nodes = 1; s_active = 1; s_next = 0;
        A.active[0] = 0;
        A.node_range[0] = make_int2(0, n);
    }
    __syncthreads();
    int* prims = A.prims;
    int* nodeof = A.node_a;
    int* act = A.active; int* act_o = A.active + BVH_MAX_NODES;
    for (int level = 0; level < 64; ++level) {
        const int na = s_active;
        if (na == 0 || na > nt) break;                       // a level never has more than BVH_BUILD_THREADS nodes (leaf >= 8, n <= 4096)
        // ---- bounds of the active nodes: primitive boxes and centroids, min / max through ordered-int atomics
        for (int k = tid; k < na; k += nt) {
            int* b = A.nb + 12 * act[k];
            for (int c = 0; c < 3; ++c) { b[c] = 0x7fffffff; b[3 + c] = (int)0x80000000; b[6 + c] = 0x7fffffff; b[9 + c] = (int)0x80000000; }
        }
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += nt) {           // uniform trip count: the warp votes / shuffles below need every lane
            const int i = i0 + tid;
            const int nd = i < n ? nodeof[i] : -1;     // -1: past the end, or the position belongs to a finished leaf
            int lo[6], hi[6];                          // ordered-int box min / centroid min, box max / centroid max
#pragma unroll
            for (int c = 0; c < 6; ++c) { lo[c] = 0x7fffffff; hi[c] = (int)0x80000000; }
            if (nd >= 0) {
                const int p = prims[i];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    lo[c] = bvh_f2o(A.box[6 * p + c]); hi[c] = bvh_f2o(A.box[6 * p + 3 + c]);
                    lo[3 + c] = hi[3 + c] = bvh_f2o(A.cen[3 * p + c]);
                }
            }
            // node ranges are contiguous, so the lanes of a warp form runs of equal node id: segmented warp reduction (a lane
            // absorbs the lane o above it while that lane is in the same run), then only the first lane of every run issues the
            // 12 atomics - one set per (warp, node) instead of one per primitive (the root alone would otherwise serialise
            // 37 000 atomics on 12 addresses, the deep levels 37 000 scattered ones each)
            const int lane = tid & 31;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ndo = __shfl_down_sync(0xffffffffu, nd, o);
                const bool take = (lane + o < 32) && ndo == nd;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const int l2 = __shfl_down_sync(0xffffffffu, lo[c], o), h2 = __shfl_down_sync(0xffffffffu, hi[c], o);
                    if (take) { lo[c] = min(lo[c], l2); hi[c] = max(hi[c], h2); }
                }
            }
            const int ndp = __shfl_up_sync(0xffffffffu, nd, 1);
            if (nd < 0 || (lane > 0 && ndp == nd)) continue;      // not the head of a run
            int* b = A.nb + 12 * nd;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                atomicMin(&b[c], lo[c]); atomicMax(&b[3 + c], hi[c]);
                atomicMin(&b[6 + c], lo[3 + c]); atomicMax(&b[9 + c], hi[3 + c]);
            }
        }
        __syncthreads();
        // ---- node boxes, leaf / split decision
        for (int k = tid; k < na; k += nt) {
            const int nd = act[k];
            const int* b = A.nb + 12 * nd;
            const int2 rg = A.node_range[nd];
            float4 mn = make_float4(bvh_o2f(b[0]), bvh_o2f(b[1]), bvh_o2f(b[2]), 0.f);
            float4 mx = make_float4(bvh_o2f(b[3]), bvh_o2f(b[4]), bvh_o2f(b[5]), 0.f);
            int ax = -1;
            if (rg.y <= A.leaf) {
                mn.w = __int_as_float(~rg.x);
                mx.w = __int_as_float(rg.y);
            } else {
                const float e0 = bvh_o2f(b[9]) - bvh_o2f(b[6]), e1 = bvh_o2f(b[10]) - bvh_o2f(b[7]), e2 = bvh_o2f(b[11]) - bvh_o2f(b[8]);
                ax = 0;
                float best = e0;
                if (e1 > best) { ax = 1; best = e1; }
                if (e2 > best) ax = 2;
            }
            A.axis[nd] = ax;
            A.nodes[2 * nd] = mn;
            A.nodes[2 * nd + 1] = mx;
        }
        __syncthreads();
        // ---- children of the splitting nodes, numbered in the order of the active list (deterministic): block-wide exclusive
        // scan of the split flags (one active node per thread)
        {
            const int nd = tid < na ? act[tid] : -1;
            const int flag = (nd >= 0 && A.axis[nd] >= 0) ? 1 : 0;
            int incl = flag;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((tid & 31) >= o) incl += t;
            }
            if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
            __syncthreads();
            if (tid < 32) {
                const int wtot = tid < (nt >> 5) ? s_scan[tid] : 0;
                int winc = wtot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, winc, o);
                    if (tid >= o) winc += t;
                }
                s_scan[tid] = winc - wtot;                     // exclusive prefix of the warp totals
                if (tid == 31) s_next = 2 * winc;              // children created on this level
            }
            __syncthreads();
            if (flag) {
                const int rank = s_scan[tid >> 5] + incl - 1;  // splitting nodes before this one
                const int2 rg = A.node_range[nd];
                const int mid = rg.y / 2, l = s_nodes + 2 * rank, r = l + 1;
                float4 mn = A.nodes[2 * nd], mx = A.nodes[2 * nd + 1];
                mn.w = __int_as_float(l); mx.w = __int_as_float(r);
                A.nodes[2 * nd] = mn; A.nodes[2 * nd + 1] = mx;
                A.node_range[l] = make_int2(rg.x, mid);
                A.node_range[r] = make_int2(rg.x + mid, rg.y - mid);
                act_o[2 * rank] = l; act_o[2 * rank + 1] = r;
            }
        }
        // ---- median split = sort of every splitting node's range by (centroid coordinate on its axis, primitive id): one
        // bitonic sort of the whole array per level with the range start as the most significant key field, so that ranges
        // (and the positions of finished leaves, keyed by their own index) stay where they are
        for (int i = tid; i < np2; i += nt) {
            unsigned long long key = ~0ull;
            if (i < n) {
                const int nd = nodeof[i];
                const int p = prims[i];
                const int ax = nd < 0 ? -1 : A.axis[nd];
                if (ax < 0) key = ((unsigned long long)i << 44) | (unsigned long long)p;
                else {
                    const unsigned ukey = (unsigned)bvh_f2o(A.cen[3 * p + ax]) ^ 0x80000000u;
                    key = ((unsigned long long)A.node_range[nd].x << 44) | ((unsigned long long)ukey << 12) | (unsigned long long)p;
                }
            }
            s_sort[i] = key;
        }
        __syncthreads();
        for (int k = 2; k <= np2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (np2 >> 1); t += nt) {
                    const int i = ((t / j) * 2 * j) + (t % j), q = i + j;
                    const unsigned long long x = s_sort[i], y = s_sort[q];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { s_sort[i] = y; s_sort[q] = x; }
                }
                __syncthreads();
            }
        // ---- new primitive order; the lower half of a split range belongs to the left child
        for (int i = tid; i < n; i += nt) {
            const int nd = nodeof[i];
            prims[i] = (int)(s_sort[i] & 0xfffull);
            if (nd < 0) continue;
            if (A.axis[nd] < 0) { nodeof[i] = -1; continue; }
            const int2 rg = A.node_range[nd];
            nodeof[i] = (i - rg.x) < rg.y / 2 ? __float_as_int(A.nodes[2 * nd].w) : __float_as_int(A.nodes[2 * nd + 1].w);
        }
        __syncthreads();
        { int* t = act; act = act_o; act_o = t; }
        if (tid == 0) { s_nodes += s_next; s_active = s_next; s_next = 0; }
        __syncthreads();
    }
    if (tid == 0) *A.n_nodes = s_nodes;
}

// Leaf-order triangle records (FrameDev::tri_rec): a | face id, b | ab.x, c | ab.y, ab.z | ac, unit normal | circle radius,
// centroid.  ab = b - a and ac = c - a are rounded once, exactly like the kernels' xsub; the lower-bound part (normal,
// centroid, radius) is computed in double and inflated so that it can only loosen the bound.
__global__ void k_tri_records(const float* __restrict__ verts, const int* __restrict__ faces, const int* __restrict__ prims, int n,
                              float4* __restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int fi = prims[i];
    const float* a = verts + 3 * faces[3 * fi];
    const float* b = verts + 3 * faces[3 * fi + 1];
    const float* c = verts + 3 * faces[3 * fi + 2];
    float ab[3], ac[3];
    for (int k = 0; k < 3; ++k) { ab[k] = xsub(b[k], a[k]); ac[k] = xsub(c[k], a[k]); }
    const double nx = (double)ab[1] * ac[2] - (double)ab[2] * ac[1], ny = (double)ab[2] * ac[0] - (double)ab[0] * ac[2],
                 nz = (double)ab[0] * ac[1] - (double)ab[1] * ac[0];
    const double nl = sqrt(nx * nx + ny * ny + nz * nz);
    float cen[3];
    for (int k = 0; k < 3; ++k) cen[k] = (float)(((double)a[k] + (double)b[k] + (double)c[k]) / 3.0);
    double rad = 0.0;
    const float* vs[3] = {a, b, c};
    for (int q = 0; q < 3; ++q) {
        double d2 = 0.0;
        for (int k = 0; k < 3; ++k) { const double d = (double)xsub(vs[q][k], cen[k]); d2 += d * d; }
        const double r = sqrt(d2);
        rad = r > rad ? r : rad;
    }
    float4* r = rec + (size_t)TRI_REC_F4 * i;
    r[0] = make_float4(a[0], a[1], a[2], __int_as_float(fi));
    r[1] = make_float4(b[0], b[1], b[2], ab[0]);
    r[2] = make_float4(c[0], c[1], c[2], ab[1]);
    r[3] = make_float4(ab[2], ac[0], ac[1], ac[2]);
    const bool ok = nl > 1e-20;
    r[4] = make_float4(ok ? (float)(nx / nl) : 0.f, ok ? (float)(ny / nl) : 0.f, ok ? (float)(nz / nl) : 0.f, (float)(rad * 1.0001 + 1e-7));
    r[5] = make_float4(cen[0], cen[1], cen[2], 0.f);
}

__global__ void k_vtx_records(const float* __restrict__ verts, const int* __restrict__ prims, int n, float4* __restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int vi = prims[i];
    rec[i] = make_float4(verts[3 * vi], verts[3 * vi + 1], verts[3 * vi + 2], __int_as_float(vi));
}

__device__ __forceinline__ double bvh_warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double bvh_warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}

// Per-node slab bound for the closest-triangle search (geom.cuh): every point of the node's triangles lies within
// |n . (x - c)| <= t of the plane (unit n, through c) and, projected into it, within r of c.  2 float4 per node:
// n.xyz, t | c.xyz, r (t and r inflated: rounding can only loosen the bound).  The direction is whichever of
// {area-weighted normal of the node's triangles, x, y, z} gives the thinnest slab.  One warp per node.
__global__ void k_tri_node_bounds(const float* __restrict__ verts, const int* __restrict__ faces, const int* __restrict__ prims,
                                  const int2* __restrict__ node_range, const int* __restrict__ n_nodes, float4* __restrict__ lb) {
    const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= *n_nodes) return;
    const int2 rg = node_range[w];
    if (rg.y > BVH_SLAB_MAX_PRIMS) {               // upper levels: no slab bound (h = 0 < t, radial distance < r: never prunes)
        if (lane == 0) { lb[2 * w] = make_float4(0.f, 0.f, 0.f, 3.0e38f); lb[2 * w + 1] = make_float4(0.f, 0.f, 0.f, 3.0e38f); }
        return;
    }
    double an[3] = {0, 0, 0}, cs[3] = {0, 0, 0};
    for (int k = rg.x + lane; k < rg.x + rg.y; k += 32) {
        const int f = prims[k];
        const float* v0 = verts + 3 * faces[3 * f];
        const float* v1 = verts + 3 * faces[3 * f + 1];
        const float* v2 = verts + 3 * faces[3 * f + 2];
        const double u[3] = {(double)v1[0] - v0[0], (double)v1[1] - v0[1], (double)v1[2] - v0[2]};
        const double q[3] = {(double)v2[0] - v0[0], (double)v2[1] - v0[1], (double)v2[2] - v0[2]};
        an[0] += u[1] * q[2] - u[2] * q[1]; an[1] += u[2] * q[0] - u[0] * q[2]; an[2] += u[0] * q[1] - u[1] * q[0];
        for (int c = 0; c < 3; ++c) cs[c] += (double)v0[c] + (double)v1[c] + (double)v2[c];
    }
    float cf[3];
    for (int c = 0; c < 3; ++c) {
        an[c] = bvh_warp_sum(an[c]);
        cf[c] = (float)(bvh_warp_sum(cs[c]) / (3.0 * rg.y));
    }
    double best_t = 1e300, best_r = 0.0;
    float best_n[3] = {0.f, 0.f, 0.f};
    for (int q = 0; q < 4; ++q) {
        const double cd[3] = {q == 0 ? an[0] : (q == 1 ? 1.0 : 0.0), q == 0 ? an[1] : (q == 2 ? 1.0 : 0.0), q == 0 ? an[2] : (q == 3 ? 1.0 : 0.0)};
        const double nl = sqrt(cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2]);
        if (nl < 1e-30) continue;                  // warp-uniform
        float nf[3];
        for (int c = 0; c < 3; ++c) nf[c] = (float)(cd[c] / nl);
        const double nfl = sqrt((double)nf[0] * nf[0] + (double)nf[1] * nf[1] + (double)nf[2] * nf[2]);    // ~1: the stored vector
        double tt = 0.0, rr = 0.0;
        for (int k = rg.x + lane; k < rg.x + rg.y; k += 32) {
            const int f = prims[k];
            for (int j = 0; j < 3; ++j) {
                const float* v = verts + 3 * faces[3 * f + j];
                const double e[3] = {(double)v[0] - cf[0], (double)v[1] - cf[1], (double)v[2] - cf[2]};
                const double h = (nf[0] * e[0] + nf[1] * e[1] + nf[2] * e[2]) / nfl;
                const double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
                const double ah = h < 0 ? -h : h, s2 = e2 - h * h;
                const double s = sqrt(s2 > 0.0 ? s2 : 0.0);
                tt = ah > tt ? ah : tt;
                rr = s > rr ? s : rr;
            }
        }
        tt = bvh_warp_max(tt);
        rr = bvh_warp_max(rr);
        if (tt < best_t) { best_t = tt; best_r = rr; for (int c = 0; c < 3; ++c) best_n[c] = nf[c]; }
    }
    if (lane == 0) {
        lb[2 * w] = make_float4(best_n[0], best_n[1], best_n[2], (float)(best_t * 1.001 + 1e-6));
        lb[2 * w + 1] = make_float4(cf[0], cf[1], cf[2], (float)(best_r * 1.001 + 1e-6));
    }
}
