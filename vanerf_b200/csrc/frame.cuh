// Per-frame setup kernels: NHWC repack of the source-view maps, vertex projection, per-view vertex visibility
// (z-buffer raster restating pytorch3d's naive rasteriser as used by get_visibility,
// src/lib/dataset/mesh_util.py:284-318, settings src/lib/common/render_utils.py:169-177) and the
// visibility-premultiplied vertex feature tables (src/networks.py:83,96,270-279 + KNN_vis :27-33).
// The reference recomputes all of this in every one of its 32 passes per image; here it runs once per frame.
#pragma once
#include "common.cuh"

#define RASTER_S 256

// (V,C,h,w) -> (V,h,w,C)
__global__ void k_repack_nhwc(const float* __restrict__ src, float* __restrict__ dst, int V, int C, int h, int w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)V * C * h * w;
    if (i >= total) return;
    const int c = (int)(i % C);
    const long long pix = i / C;
    const int x = (int)(pix % w);
    const int y = (int)((pix / w) % h);
    const int v = (int)(pix / ((long long)w * h));
    dst[i] = src[(((long long)v * C + c) * h + y) * w + x];
}

// img (V,3,H,W) + fg (V,H,W) u8 -> (V,H,W,4)
__global__ void k_repack_imgm(const float* __restrict__ img, const unsigned char* __restrict__ fg,
                              float4* __restrict__ dst, int V, int H, int W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long hw = (long long)H * W;
    if (i >= V * hw) return;
    const int v = (int)(i / hw);
    const long long p = i % hw;
    const float* b = img + (long long)v * 3 * hw;
    dst[i] = make_float4(b[p], b[hw + p], b[2 * hw + p], fg[i] ? 1.0f : 0.0f);
}

// A.2: per (view, vertex): raster-space xyz = (cat(xy/(W-1|H-1), (z-znear)/(zfar-znear)) + 1)/2, and the [-1,1]
// image coordinates used for vertex feature sampling (src/model.py:845-853).
__global__ void k_project_verts(FrameDev fr, float* __restrict__ xyz_ndc, float* __restrict__ xy11) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fr.V * fr.n_verts) return;
    const int v = i / fr.n_verts, j = i % fr.n_verts;
    const float* p = fr.verts + 3 * j;
    const float* M = fr.KRT[v];
    const float hx = xaffine(M, 0, p[0], p[1], p[2]);
    const float hy = xaffine(M, 1, p[0], p[1], p[2]);
    const float hz = xaffine(M, 2, p[0], p[1], p[2]);
    const float ze = xadd(hz, 1e-8f);
    const float vx = xdiv(hx, ze), vy = xdiv(hy, ze);
    const float wm = (float)fr.W - 1.0f, hm = (float)fr.H - 1.0f;
    const float x01 = xdiv(vx, wm), y01 = xdiv(vy, hm);
    const float z01 = xdiv(xsub(hz, fr.znear), fr.z_range);
    xyz_ndc[3 * i + 0] = xdiv(xadd(x01, 1.0f), 2.0f);
    xyz_ndc[3 * i + 1] = xdiv(xadd(y01, 1.0f), 2.0f);
    xyz_ndc[3 * i + 2] = xdiv(xadd(z01, 1.0f), 2.0f);
    xy11[2 * i + 0] = xsub(xmul(2.0f, x01), 1.0f);
    xy11[2 * i + 1] = xsub(xmul(2.0f, y01), 1.0f);
}

__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return xsub(xmul(xsub(px, ax), xsub(by, ay)), xmul(xsub(py, ay), xsub(bx, ax)));
}

// One thread per (view, face): scan the pixels of its bounding box, atomicMin (depth bits << 32 | face) per pixel.
// Per-pixel arithmetic is that of oracle/geom_oracle.c: vo_rasterize.
__global__ void k_raster_faces(FrameDev fr, const float* __restrict__ xyz_ndc, unsigned long long* __restrict__ zbuf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fr.V * fr.n_faces) return;
    const int v = i / fr.n_faces, f = i % fr.n_faces;
    const int S = RASTER_S;
    const float kEps = 1e-8f;
    const int* fv = fr.faces + 3 * f;
    const float* base = xyz_ndc + (size_t)v * fr.n_verts * 3;
    const float* v0 = base + 3 * fv[0];
    const float* v1 = base + 3 * fv[1];
    const float* v2 = base + 3 * fv[2];
    const float zmax = fmaxf(v0[2], fmaxf(v1[2], v2[2]));
    const float xmin = fminf(v0[0], fminf(v1[0], v2[0])), xmax = fmaxf(v0[0], fmaxf(v1[0], v2[0]));
    const float ymin = fminf(v0[1], fminf(v1[1], v2[1])), ymax = fmaxf(v0[1], fmaxf(v1[1], v2[1]));
    const float area = edge_fn(v0[0], v0[1], v1[0], v1[1], v2[0], v2[1]);
    if (zmax < kEps || area < 0.0f || (area <= kEps && area >= -kEps)) return;
    if (!(xmax >= -1.0f && xmin <= 1.0f && ymax >= -1.0f && ymin <= 1.0f)) return;     // also drops NaN boxes
    // pixel centre: pf = -1 + (2*k + 1)/S with k = S-1-col (or row).  Conservative k range, exact test inside.
    int kx0 = (int)floorf((fmaxf(xmin, -1.0f) + 1.0f) * 0.5f * S - 0.5f) - 1, kx1 = (int)ceilf((fminf(xmax, 1.0f) + 1.0f) * 0.5f * S - 0.5f) + 1;
    int ky0 = (int)floorf((fmaxf(ymin, -1.0f) + 1.0f) * 0.5f * S - 0.5f) - 1, ky1 = (int)ceilf((fminf(ymax, 1.0f) + 1.0f) * 0.5f * S - 0.5f) + 1;
    kx0 = max(kx0, 0); ky0 = max(ky0, 0); kx1 = min(kx1, S - 1); ky1 = min(ky1, S - 1);
    const float barea = xadd(edge_fn(v2[0], v2[1], v0[0], v0[1], v1[0], v1[1]), kEps);
    unsigned long long* zb = zbuf + (size_t)v * S * S;
    for (int yi = ky0; yi <= ky1; ++yi) {
        const float pyf = xadd(-1.0f, xdiv(xadd(xmul(2.0f, (float)yi), 1.0f), (float)S));
        if (pyf > ymax || pyf < ymin) continue;
        for (int xi = kx0; xi <= kx1; ++xi) {
            const float pxf = xadd(-1.0f, xdiv(xadd(xmul(2.0f, (float)xi), 1.0f), (float)S));
            if (pxf > xmax || pxf < xmin) continue;
            const float w0 = xdiv(edge_fn(pxf, pyf, v1[0], v1[1], v2[0], v2[1]), barea);
            const float w1 = xdiv(edge_fn(pxf, pyf, v2[0], v2[1], v0[0], v0[1]), barea);
            const float w2 = xdiv(edge_fn(pxf, pyf, v0[0], v0[1], v1[0], v1[1]), barea);
            const float t0 = xmul(xmul(w0, v1[2]), v2[2]), t1 = xmul(xmul(v0[2], w1), v2[2]), t2 = xmul(xmul(v0[2], v1[2]), w2);
            const float den = fmaxf(xadd(xadd(t0, t1), t2), kEps);
            const float b0 = xdiv(t0, den), b1 = xdiv(t1, den), b2 = xdiv(t2, den);
            float pz = xadd(xadd(xmul(b0, v0[2]), xmul(b1, v1[2])), xmul(b2, v2[2]));
            if (pz < 0.0f) continue;
            if (!(b0 > 0.0f && b1 > 0.0f && b2 > 0.0f)) continue;
            if (pz == 0.0f) pz = 0.0f;                 // canonical +0
            const int row = S - 1 - yi, col = S - 1 - xi;
            const unsigned long long key = ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned)f;
            atomicMin(&zb[row * S + col], key);
        }
    }
}

// visible faces -> visible vertices; an empty pixel contributes index -1 = the LAST face
__global__ void k_resolve_vis(FrameDev fr, const unsigned long long* __restrict__ zbuf, float* __restrict__ vis) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int S2 = RASTER_S * RASTER_S;
    if (i >= fr.V * S2) return;
    const int v = i / S2;
    const unsigned long long key = zbuf[i];
    const int f = (key == 0xffffffffffffffffull) ? fr.n_faces - 1 : (int)(key & 0xffffffffu);
    const int* fv = fr.faces + 3 * f;
    float* o = vis + (size_t)v * fr.n_verts;
    o[fv[0]] = 1.0f; o[fv[1]] = 1.0f; o[fv[2]] = 1.0f;
}

__device__ __forceinline__ float tap_nhwc(const float* __restrict__ map, int C, int c, const Bilin& b) {
    const float v00 = map[(size_t)b.i00 * C + c];
    const float v01 = b.i01 >= 0 ? map[(size_t)b.i01 * C + c] : 0.0f;
    const float v10 = b.i10 >= 0 ? map[(size_t)b.i10 * C + c] : 0.0f;
    const float v11 = b.i11 >= 0 ? map[(size_t)b.i11 * C + c] : 0.0f;
    return bilin_mix(b, v00, v01, v10, v11);
}

// One thread per (view, vertex, channel) over the 64 + 8 + 32 table channels.
__global__ void k_vertex_tables(FrameDev fr, const float* __restrict__ xy11, const float* __restrict__ gfeat,
                                float* __restrict__ T64, float* __restrict__ T8, float* __restrict__ Ttex) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int CH = 64 + 8 + 32;
    if (i >= (long long)fr.V * fr.n_verts * CH) return;
    const int c = (int)(i % CH);
    const int vj = (int)(i / CH);
    const int v = vj / fr.n_verts;
    const float x = xy11[2 * vj], y = xy11[2 * vj + 1];
    const float vis = fr.vis[vj];
    if (c < 64) {
        const Bilin b = bilin_setup(x, y, fr.g0w, fr.g0h);
        T64[(size_t)vj * 64 + c] = tap_nhwc(fr.geo0 + (size_t)v * fr.g0h * fr.g0w * 64, 64, c, b) * vis;
    } else if (c < 72) {
        const Bilin b = bilin_setup(x, y, fr.g1w, fr.g1h);
        T8[(size_t)vj * 8 + (c - 64)] = tap_nhwc(fr.geo1 + (size_t)v * fr.g1h * fr.g1w * 8, 8, c - 64, b) * vis;
    } else {
        const int k = c - 72;
        float val = 0.0f;
        if (k < 3) {
            const Bilin b = bilin_setup(x, y, fr.W, fr.H);
            val = tap_nhwc(fr.imgm + (size_t)v * fr.H * fr.W * 4, 4, k, b);
        } else if (k < 11) {
            const Bilin b = bilin_setup(x, y, fr.tw, fr.th);
            val = tap_nhwc(fr.tex + (size_t)v * fr.th * fr.tw * 8, 8, k - 3, b);
        } else if (k < 29) {
            val = gfeat[(size_t)vj * 18 + (k - 11)];
        }
        Ttex[(size_t)vj * 32 + k] = val * vis;
    }
}

// keypoints in each source camera frame (src/spatial.py:84); one thread per (view, keypoint) — host fills FrameDev
// from the returned buffer, so this is plain host code in api.cu (42 x V points).
