/* ORACLE — test infrastructure only.  Nothing under vanerf_b200/ may import, link or call this.
 *
 * CPU restatement (plain C, fp32, no FMA contraction: build with -ffp-contract=off) of the third-party
 * geometry queries the VANeRF render path calls through kaolin 0.15.0 / pytorch3d 0.7.5
 * (un-vendored wheels pinned in reference requirements.txt:81,158; call sites
 * src/lib/dataset/mesh_util.py:498-524, :284-318, src/networks.py:27-33):
 *
 *   vo_point_mesh_distance  kaolin.metrics.trianglemesh.point_to_mesh_distance (mesh_util.py:509)
 *   vo_check_sign           kaolin.ops.mesh.check_sign                         (mesh_util.py:511)
 *   vo_knn1                 pytorch3d.ops.knn_points(K=1)                      (networks.py:28)
 *   vo_rasterize            pytorch3d.renderer.mesh.rasterize_meshes, naive path, faces_per_pixel=1,
 *                           blur_radius=0, perspective_correct, cull_backfaces (mesh_util.py:303,
 *                           src/lib/common/render_utils.py:169-177)
 *
 * PARITY UNPINNED at this boundary: neither library is installed here, no network, and the reference
 * has no test that pins their arithmetic.  The semantics below restate the published algorithms
 * (SURVEY.md Appendix A.3): exact point-triangle squared distance (Ericson, RTCD 5.1.5) with
 * first-minimum face index; ray-parity containment along +x (Moller-Trumbore); squared-distance
 * nearest vertex with first-minimum index; pytorch3d's naive z-buffer rasteriser conventions.
 * Brute force on purpose: the CUDA path uses acceleration structures and must reproduce these
 * results exactly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline void sub3(const float* a, const float* b, float* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }

/* squared distance from p to triangle (a,b,c); closest-point regions after Ericson */
static float point_tri_dist2(const float* p, const float* a, const float* b, const float* c) {
    float ab[3], ac[3], ap[3], bp[3], cp[3], q[3], d[3];
    sub3(b, a, ab); sub3(c, a, ac); sub3(p, a, ap);
    float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) { q[0] = a[0]; q[1] = a[1]; q[2] = a[2]; goto done; }
    sub3(p, b, bp);
    float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) { q[0] = b[0]; q[1] = b[1]; q[2] = b[2]; goto done; }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        float v = d1 / (d1 - d3);
        for (int i = 0; i < 3; ++i) q[i] = a[i] + v * ab[i];
        goto done;
    }
    sub3(p, c, cp);
    float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) { q[0] = c[0]; q[1] = c[1]; q[2] = c[2]; goto done; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        float w = d2 / (d2 - d6);
        for (int i = 0; i < 3; ++i) q[i] = a[i] + w * ac[i];
        goto done;
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int i = 0; i < 3; ++i) q[i] = b[i] + w * (c[i] - b[i]);
        goto done;
    }
    {
        float denom = 1.0f / ((va + vb) + vc);
        float v = vb * denom, w = vc * denom;
        for (int i = 0; i < 3; ++i) q[i] = (a[i] + ab[i] * v) + ac[i] * w;
    }
done:
    sub3(p, q, d);
    return dot3(d, d);
}

void vo_point_mesh_distance(const float* pts, int64_t N, const float* verts, const int64_t* faces, int64_t F,
                            float* out_d2, int64_t* out_idx) {
    for (int64_t n = 0; n < N; ++n) {
        const float* p = pts + 3 * n;
        float best = INFINITY; int64_t bi = 0;
        for (int64_t f = 0; f < F; ++f) {
            const float* a = verts + 3 * faces[3 * f + 0];
            const float* b = verts + 3 * faces[3 * f + 1];
            const float* c = verts + 3 * faces[3 * f + 2];
            float d = point_tri_dist2(p, a, b, c);
            if (d < best) { best = d; bi = f; }        /* strict: first minimal face wins */
        }
        out_d2[n] = best; out_idx[n] = bi;
    }
}

/* +x ray parity; a point inside both (overlapping) hands counts as outside (XOR) */
void vo_check_sign(const float* pts, int64_t N, const float* verts, const int64_t* faces, int64_t F, uint8_t* inside) {
    for (int64_t n = 0; n < N; ++n) {
        const float* p = pts + 3 * n;
        int cnt = 0;
        for (int64_t f = 0; f < F; ++f) {
            const float* v0 = verts + 3 * faces[3 * f + 0];
            const float* v1 = verts + 3 * faces[3 * f + 1];
            const float* v2 = verts + 3 * faces[3 * f + 2];
            float e1[3], e2[3], s[3];
            sub3(v1, v0, e1); sub3(v2, v0, e2);
            float a = e1[2] * e2[1] - e1[1] * e2[2];          /* e1 . (dir x e2), dir = +x */
            if (fabsf(a) < 1e-20f) continue;
            float inv = 1.0f / a;
            sub3(p, v0, s);
            float u = inv * (s[2] * e2[1] - s[1] * e2[2]);
            if (!(u >= 0.0f)) continue;
            float qx = s[1] * e1[2] - s[2] * e1[1];
            float qy = s[2] * e1[0] - s[0] * e1[2];
            float qz = s[0] * e1[1] - s[1] * e1[0];
            float v = inv * qx;                               /* dir . q */
            if (!(v >= 0.0f) || !(u + v <= 1.0f)) continue;
            float t = inv * ((e2[0] * qx + e2[1] * qy) + e2[2] * qz);
            if (t > 0.0f) cnt++;
        }
        inside[n] = (uint8_t)(cnt & 1);
    }
}

void vo_knn1(const float* pts, int64_t N, const float* verts, int64_t Nv, int64_t* out_idx) {
    for (int64_t n = 0; n < N; ++n) {
        const float* p = pts + 3 * n;
        float best = INFINITY; int64_t bi = 0;
        for (int64_t j = 0; j < Nv; ++j) {
            float dx = p[0] - verts[3 * j], dy = p[1] - verts[3 * j + 1], dz = p[2] - verts[3 * j + 2];
            float d = (dx * dx + dy * dy) + dz * dz;
            if (d < best) { best = d; bi = j; }
        }
        out_idx[n] = bi;
    }
}

static inline float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

/* pytorch3d naive rasteriser restated: square image S, NDC +X left / +Y up, pixel (row i, col j) centre at
 * (1-(2j+1)/S, 1-(2i+1)/S); K=1 nearest pz, ties -> lowest face index; -1 = empty. */
void vo_rasterize(const float* xyz, const int64_t* faces, int64_t F, int S, int row0, int row1, int64_t* pix_to_face) {
    const float kEps = 1e-8f;
    for (int i = row0; i < row1; ++i) {
        for (int j = 0; j < S; ++j) {
            int yi = S - 1 - i, xi = S - 1 - j;
            float pxf = -1.0f + (2.0f * (float)xi + 1.0f) / (float)S;
            float pyf = -1.0f + (2.0f * (float)yi + 1.0f) / (float)S;
            float bestz = INFINITY; int64_t bf = -1;
            for (int64_t f = 0; f < F; ++f) {
                const float* v0 = xyz + 3 * faces[3 * f + 0];
                const float* v1 = xyz + 3 * faces[3 * f + 1];
                const float* v2 = xyz + 3 * faces[3 * f + 2];
                float zmax = fmaxf(v0[2], fmaxf(v1[2], v2[2]));
                float xmin = fminf(v0[0], fminf(v1[0], v2[0])), xmax = fmaxf(v0[0], fmaxf(v1[0], v2[0]));
                float ymin = fminf(v0[1], fminf(v1[1], v2[1])), ymax = fmaxf(v0[1], fmaxf(v1[1], v2[1]));
                int outside = (pxf > xmax) || (pxf < xmin) || (pyf > ymax) || (pyf < ymin);
                float area = edge_fn(v0[0], v0[1], v1[0], v1[1], v2[0], v2[1]);
                int back = area < 0.0f;
                int zero_area = (area <= kEps && area >= -kEps);
                if (zmax < kEps || back || outside || zero_area) continue;
                float barea = edge_fn(v2[0], v2[1], v0[0], v0[1], v1[0], v1[1]) + kEps;
                float w0 = edge_fn(pxf, pyf, v1[0], v1[1], v2[0], v2[1]) / barea;
                float w1 = edge_fn(pxf, pyf, v2[0], v2[1], v0[0], v0[1]) / barea;
                float w2 = edge_fn(pxf, pyf, v0[0], v0[1], v1[0], v1[1]) / barea;
                float t0 = w0 * v1[2] * v2[2], t1 = v0[2] * w1 * v2[2], t2 = v0[2] * v1[2] * w2;
                float den = fmaxf((t0 + t1) + t2, kEps);
                float b0 = t0 / den, b1 = t1 / den, b2 = t2 / den;
                float pz = (b0 * v0[2] + b1 * v1[2]) + b2 * v2[2];
                if (pz < 0.0f) continue;
                if (!(b0 > 0.0f && b1 > 0.0f && b2 > 0.0f)) continue;   /* blur_radius = 0: inside only */
                if (pz < bestz) { bestz = pz; bf = f; }
            }
            pix_to_face[(int64_t)i * S + j] = bf;
        }
    }
}
