// Shared definitions of the vanerf_b200 kernels.
#pragma once

#include <stdint.h>
#include <stddef.h>

#ifdef VANERF_HOST_EMUL
#include "host_emul.h"
#define DYN_SMEM(type, name) EMUL_DYN_SMEM(type, name)
#define HD inline
#else
#include <cuda_runtime.h>
#define DYN_SMEM(type, name) extern __shared__ __align__(128) unsigned char name##_raw_[]; \
    type* name = reinterpret_cast<type*>(name##_raw_)
#define HD __host__ __device__ __forceinline__
#define VANERF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif

#include "../../include/vanerf_b200.h"

#define MAXV VANERF_MAX_VIEWS
#define NKPT VANERF_N_KPT
#define TRI_REC_F4 6             // float4 per triangle record (geom.cuh)
#define NUM_V_HAND 779          // src/networks.py:25  (twin vertex = (id + 779) mod 1558)

// ---------------------------------------------------------------------------------------------------------------
// Exact fp32 arithmetic: one correctly rounded operation per call, never contracted into FMA by the compiler.
// The bit-exact part of the path (ray setup, sample depths/positions, projections feeding masks, geometry
// queries) is written with these only, in the operand order of oracle/oracle_torch.py.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// (a0*b0 + a1*b1) + a2*b2
__device__ __forceinline__ float xdot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return xadd(xadd(xmul(a0, b0), xmul(a1, b1)), xmul(a2, b2));
}
// torch-CPU 2-norm of a 3-vector: sqrt(fma(z,z, fma(y,y, x*x)))
__device__ __forceinline__ float xnorm3(float x, float y, float z) {
    return xsqrt(xfma(z, z, xfma(y, y, xmul(x, x))));
}
// ((p0*m0 + p1*m1) + p2*m2) + m3   for row j of a row-major 4x4
__device__ __forceinline__ float xaffine(const float* M, int j, float p0, float p1, float p2) {
    return xadd(xdot3(p0, p1, p2, M[4 * j + 0], M[4 * j + 1], M[4 * j + 2]), M[4 * j + 3]);
}

struct Bilin {            // grid_sample(bilinear, border, align_corners=True) tap set, src/utils.py:136-151
    int i00, i01, i10, i11;        // linear pixel indices (y*W + x) of nw, ne, sw, se; -1 = out of range (weight-0 tap)
    float nw, ne, sw, se;
};
__device__ __forceinline__ Bilin bilin_setup(float x, float y, int Wf, int Hf) {
    float ix = xmul(xdiv(xadd(x, 1.0f), 2.0f), (float)(Wf - 1));
    float iy = xmul(xdiv(xadd(y, 1.0f), 2.0f), (float)(Hf - 1));
    ix = fminf((float)(Wf - 1), fmaxf(ix, 0.0f));
    iy = fminf((float)(Hf - 1), fmaxf(iy, 0.0f));
    float x0 = floorf(ix), y0 = floorf(iy);
    float w = xsub(ix, x0), e = xsub(1.0f, w), n = xsub(iy, y0), s = xsub(1.0f, n);
    Bilin b;
    b.nw = xmul(s, e); b.ne = xmul(s, w); b.sw = xmul(n, e); b.se = xmul(n, w);
    int x0i = (int)x0, y0i = (int)y0, x1i = x0i + 1, y1i = y0i + 1;
    bool xo = x1i < Wf, yo = y1i < Hf;
    b.i00 = y0i * Wf + x0i;
    b.i01 = xo ? y0i * Wf + x1i : -1;
    b.i10 = yo ? y1i * Wf + x0i : -1;
    b.i11 = (xo && yo) ? y1i * Wf + x1i : -1;
    return b;
}
// torch-CPU accumulation order: fma(se,SE, fma(sw,SW, fma(ne,NE, nw*NW)))
__device__ __forceinline__ float bilin_mix(const Bilin& b, float vnw, float vne, float vsw, float vse) {
    return xfma(vse, b.se, xfma(vsw, b.sw, xfma(vne, b.ne, xmul(vnw, b.nw))));
}

// ---------------------------------------------------------------------------------------------------------------
// Device-side parameter blocks
// ---------------------------------------------------------------------------------------------------------------
struct FrameDev {                 // filled by vanerf_frame_setup, passed by value to kernels
    int V, H, W, n_verts, n_faces;
    float znear, zfar, z_range;
    float KRT[MAXV][16];
    float extrin[MAXV][16];
    float src_pos[MAXV][3];
    const float* kpt_cam;          // (V,42,3) keypoints in each source camera frame (src/spatial.py:84)
    // maps, NHWC fp32
    const float* geo0; int g0h, g0w;       // (V,h,w,64)
    const float* geo1; int g1h, g1w;       // (V,h,w,8)
    const float* tex;  int th, tw;         // (V,h,w,8)
    const float* imgm;                     // (V,H,W,4) = r,g,b,fg
    // vertex tables, premultiplied by vert_vis
    const float* T64;                      // (V,Nv,64)
    const float* T8;                       // (V,Nv,8)
    const float* Ttex;                     // (V,Nv,32): img3, tex8, gf18, 0,0,0
    const float* vis;                      // (V,Nv)
    // mesh + acceleration structures
    const float* verts;                    // (Nv,3)
    const int* faces;                      // (F,3)
    const float4* tri_nodes; const int* tri_prims;   // BVH over triangles (prims = face ids in leaf order)
    const float4* tri_node_lb;                       // per node: slab bound n.xyz, t | c.xyz, r (bvh::triangle_node_bounds)
    const float4* vtx_nodes; const int* vtx_prims;   // BVH over vertices
    // per-primitive records in leaf order, so that a leaf visit is one dependent load instead of prims -> faces -> verts:
    // triangle = 4 x float4 {a.xyz, as_float(face id)}, {b.xyz, ab.x}, {c.xyz, ab.y}, {ab.z, ac.xyz} with ab = b - a,
    // ac = c - a rounded exactly as the kernels' xsub does; vertex = {xyz, as_float(vertex id)}
    const float4* tri_rec;      // TRI_REC_F4 float4 per triangle, leaf order: a|id, b|ab.x, c|ab.y, ab.z|ac, n|r, centre
    const float4* vtx_rec;
};

struct TargetDev {
    float inv_K[9], R[9], cam_pos[3], znear, zfar, bmin[3], bmax[3];
};

// Row record written by the gather kernel, one per (sample, view), fp32.  Offsets in floats.
#define REC_PX64 0       // pixel-aligned geo0 (64)
#define REC_A64 64       // T64[nn] * vis[nn]
#define REC_B64 128      // T64[twin] * vis[twin]
#define REC_PX8 192      // pixel-aligned geo1 (8)
#define REC_A8 200
#define REC_B8 208
#define REC_QIMG 216     // img r,g,b + fg-mask tap
#define REC_QTEX 220     // pixel-aligned tex (8)
#define REC_ATEX 228     // Ttex[nn]   (img3, tex8, gf18, 0,0,0) * vis
#define REC_BTEX 260     // Ttex[twin]
#define REC_SDF 292
#define REC_QVIS 293
#define REC_VN 294
#define REC_VT 295
#define REC_CAM 296      // camera-space xyz (3) + pad
#define REC_RD 300       // ray_diff (4)
#define REC_PW 304       // pix_weight
#define REC_MASK 305     // out_mask
#define REC_STRIDE 308   // floats (77 float4 units)

// One packed linear layer in device memory: wt is the TRANSPOSED weight (K rows of Npad floats, zero padded),
// b has Npad floats (zeros when the layer has no bias).
struct LayerDev {
    const float* wt;
    const float* b;
    int K, N, Npad, pad_;
};
enum LayerId {
    L_GEO_AT0 = 0, L_GEO_AT1, L_GEO_F0, L_GEO_F1, L_GEO8_AT0, L_GEO8_AT1, L_GEO8_F0, L_GEO8_F1,
    L_MLP0, L_MLP1, L_MLP2, L_MLP3, L_POST0, L_POST1, L_POST2, L_COMPRESS,
    L_TEX_AT0, L_TEX_AT1, L_TEX_F0, L_TEX_F1,
    L_RAY0, L_RAY1, L_BASE0, L_BASE1, L_VIS1_0, L_VIS1_1, L_VIS2_0, L_VIS2_1, L_OUT0, L_OUT1, L_OUT2,
    L_COUNT
};
struct NetDev {
    LayerDev layer[L_COUNT];
    float ani_al_abs, beta;
};

#define CUDA_TRY(ctx, expr)                                                   \
    do {                                                                      \
        cudaError_t e_ = (expr);                                              \
        if (e_ != cudaSuccess) return ctx_fail((ctx), e_, #expr, __LINE__);   \
    } while (0)
