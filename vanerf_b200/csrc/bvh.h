// Host-side BVH builder (median split on the longest centroid axis).  Per frame: 3108 triangles / 1558 vertices.
// Node = 2 x float4: {min.xyz, as_float(a)}, {max.xyz, as_float(b)};  inner: a = left child, b = right child;
// leaf: a = ~first_prim (negative), b = prim count.  Prims are indices into the original arrays, in leaf order.
#pragma once
#include <algorithm>
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

inline void build_points(const float* pts, int n, Tree& out) {
    std::vector<Box> boxes(n);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) boxes[i].mn[c] = boxes[i].mx[c] = pts[3 * i + c];
    build(boxes, 16, out);
}

}  // namespace bvh
