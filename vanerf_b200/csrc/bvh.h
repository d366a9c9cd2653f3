// Host-side BVH builder (median split on the longest centroid axis).  Per frame: 3108 triangles / 1558 vertices.
// Node = 2 x float4: {min.xyz, as_float(a)}, {max.xyz, as_float(b)};  inner: a = left child, b = right child;
// leaf: a = ~first_prim (negative), b = prim count.  Prims are indices into the original arrays, in leaf order.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace bvh {

struct Box { float mn[3], mx[3]; };

struct Tree {
    std::vector<float> nodes;    // 8 floats per node
    std::vector<int> prims;
    int n_nodes() const { return (int)(nodes.size() / 8); }
};

inline float as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }

inline void build(const std::vector<Box>& boxes, int leaf_size, Tree& out) {
    const int n = (int)boxes.size();
    out.nodes.clear();
    out.prims.resize(n);
    for (int i = 0; i < n; ++i) out.prims[i] = i;
    std::vector<float> cen(3 * (size_t)n);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) cen[3 * i + c] = 0.5f * (boxes[i].mn[c] + boxes[i].mx[c]);
    struct Job { int first, count, node; };
    std::vector<Job> stack;
    out.nodes.resize(8);
    stack.push_back({0, n, 0});
    while (!stack.empty()) {
        Job j = stack.back();
        stack.pop_back();
        Box b;
        float cmn[3], cmx[3];
        for (int c = 0; c < 3; ++c) { b.mn[c] = cmn[c] = 1e30f; b.mx[c] = cmx[c] = -1e30f; }
        for (int i = j.first; i < j.first + j.count; ++i) {
            const int p = out.prims[i];
            for (int c = 0; c < 3; ++c) {
                b.mn[c] = std::min(b.mn[c], boxes[p].mn[c]);
                b.mx[c] = std::max(b.mx[c], boxes[p].mx[c]);
                cmn[c] = std::min(cmn[c], cen[3 * p + c]);
                cmx[c] = std::max(cmx[c], cen[3 * p + c]);
            }
        }
        float* nd = &out.nodes[8 * (size_t)j.node];
        nd[0] = b.mn[0]; nd[1] = b.mn[1]; nd[2] = b.mn[2];
        nd[4] = b.mx[0]; nd[5] = b.mx[1]; nd[6] = b.mx[2];
        if (j.count <= leaf_size) {
            nd[3] = as_float(~j.first);
            nd[7] = as_float(j.count);
            continue;
        }
        int ax = 0;
        if (cmx[1] - cmn[1] > cmx[ax] - cmn[ax]) ax = 1;
        if (cmx[2] - cmn[2] > cmx[ax] - cmn[ax]) ax = 2;
        const int mid = j.count / 2;
        std::nth_element(out.prims.begin() + j.first, out.prims.begin() + j.first + mid,
                         out.prims.begin() + j.first + j.count,
                         [&](int a, int c) { return cen[3 * a + ax] < cen[3 * c + ax] || (cen[3 * a + ax] == cen[3 * c + ax] && a < c); });
        const int l = (int)(out.nodes.size() / 8), r = l + 1;
        out.nodes.resize(out.nodes.size() + 16);
        nd = &out.nodes[8 * (size_t)j.node];     // resize may move
        nd[3] = as_float(l);
        nd[7] = as_float(r);
        stack.push_back({j.first, mid, l});
        stack.push_back({j.first + mid, j.count - mid, r});
    }
}

inline void build_triangles(const float* verts, const int32_t* faces, int n_faces, Tree& out) {
    std::vector<Box> boxes(n_faces);
    for (int f = 0; f < n_faces; ++f) {
        for (int c = 0; c < 3; ++c) {
            float a = verts[3 * faces[3 * f] + c], b = verts[3 * faces[3 * f + 1] + c], d = verts[3 * faces[3 * f + 2] + c];
            boxes[f].mn[c] = std::min(a, std::min(b, d));
            boxes[f].mx[c] = std::max(a, std::max(b, d));
        }
    }
    // leaf sizes from sweeps on B200 (2/4/8/16 triangles with the per-triangle lower bound x 4/8/16 vertices); VANERF_TRI_LEAF: developer override
    const char* e = std::getenv("VANERF_TRI_LEAF");
    build(boxes, e ? std::max(1, std::min(32, std::atoi(e))) : 8, out);
}

// Per-node slab bound for the closest-triangle search (geom.cuh): every point of the node's triangles lies within
// |n . (x - c)| <= t of the plane (unit n, through c) and, projected into it, within r of c.  8 floats per node:
// n.xyz, t | c.xyz, r (t and r inflated so that rounding can only loosen the bound).  The direction is whichever of
// {area-weighted normal of the node's triangles, x, y, z} gives the thinnest slab.
inline void triangle_node_bounds(const Tree& t, const float* verts, const int32_t* faces, std::vector<float>& lb) {
    const int nn = t.n_nodes();
    lb.assign((size_t)nn * 8, 0.0f);
    std::vector<int> first(nn), count(nn);
    auto as_int = [](float f) { int v; std::memcpy(&v, &f, 4); return v; };
    for (int i = nn - 1; i >= 0; --i) {                      // children have larger indices than their parent
        const int a = as_int(t.nodes[8 * (size_t)i + 3]), b = as_int(t.nodes[8 * (size_t)i + 7]);
        if (a < 0) { first[i] = ~a; count[i] = b; }
        else { first[i] = std::min(first[a], first[b]); count[i] = count[a] + count[b]; }
    }
    std::vector<double> pts;
    for (int i = 0; i < nn; ++i) {
        pts.clear();
        double an[3] = {0, 0, 0}, c[3] = {0, 0, 0};
        for (int k = first[i]; k < first[i] + count[i]; ++k) {
            const int f = t.prims[k];
            const float* v0 = verts + 3 * faces[3 * f];
            const float* v1 = verts + 3 * faces[3 * f + 1];
            const float* v2 = verts + 3 * faces[3 * f + 2];
            const double u[3] = {(double)v1[0] - v0[0], (double)v1[1] - v0[1], (double)v1[2] - v0[2]};
            const double w[3] = {(double)v2[0] - v0[0], (double)v2[1] - v0[1], (double)v2[2] - v0[2]};
            an[0] += u[1] * w[2] - u[2] * w[1]; an[1] += u[2] * w[0] - u[0] * w[2]; an[2] += u[0] * w[1] - u[1] * w[0];
            for (const float* v : {v0, v1, v2}) { pts.push_back(v[0]); pts.push_back(v[1]); pts.push_back(v[2]); c[0] += v[0]; c[1] += v[1]; c[2] += v[2]; }
        }
        const size_t np = pts.size() / 3;
        float cf[3];
        for (int k = 0; k < 3; ++k) cf[k] = (float)(c[k] / (double)np);
        double cand[4][3] = {{an[0], an[1], an[2]}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        double best_t = 1e300, best_r = 0.0;
        float best_n[3] = {0, 0, 0};
        for (int q = 0; q < 4; ++q) {
            const double nl = std::sqrt(cand[q][0] * cand[q][0] + cand[q][1] * cand[q][1] + cand[q][2] * cand[q][2]);
            if (nl < 1e-30) continue;
            float nf[3];
            for (int k = 0; k < 3; ++k) nf[k] = (float)(cand[q][k] / nl);
            const double nfl = std::sqrt((double)nf[0] * nf[0] + (double)nf[1] * nf[1] + (double)nf[2] * nf[2]);   // ~1: the stored vector
            double tt = 0.0, rr = 0.0;
            for (size_t j = 0; j < np; ++j) {
                const double e[3] = {pts[3 * j] - cf[0], pts[3 * j + 1] - cf[1], pts[3 * j + 2] - cf[2]};
                const double h = (nf[0] * e[0] + nf[1] * e[1] + nf[2] * e[2]) / nfl;
                const double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
                tt = std::max(tt, std::fabs(h));
                rr = std::max(rr, std::sqrt(std::max(e2 - h * h, 0.0)));
            }
            if (tt < best_t) { best_t = tt; best_r = rr; for (int k = 0; k < 3; ++k) best_n[k] = nf[k]; }
        }
        float* o = &lb[8 * (size_t)i];
        o[0] = best_n[0]; o[1] = best_n[1]; o[2] = best_n[2]; o[3] = (float)(best_t * 1.001 + 1e-6);
        o[4] = cf[0]; o[5] = cf[1]; o[6] = cf[2]; o[7] = (float)(best_r * 1.001 + 1e-6);
    }
}

inline void build_points(const float* pts, int n, Tree& out) {
    std::vector<Box> boxes(n);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) boxes[i].mn[c] = boxes[i].mx[c] = pts[3 * i + c];
    build(boxes, 16, out);
}

}  // namespace bvh
