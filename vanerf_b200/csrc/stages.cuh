// Stage-level primitives behind the reference's per-stage callables (SURVEY.md §8(b) "kept surface"): feat_sample
// (src/utils.py:136-151), the K=1 nearest-vertex search of KNN_vis (src/networks.py:27-33), one dense layer with the
// activations the path uses (Conv1d(k=1) / Linear: src/networks.py:47-71,224-235, src/utils.py:670-685,
// src/model.py:1578-1591) and SpatialEncoder's rel_z_decay encoding (src/spatial.py:109-117).  vanerf_b200/stages.py
// composes them into GeoVisFusion / TexVisFusion / MLPUNetFusion / IBRRenderingHead / SpatialEncoder with the
// reference's signatures.  The render path itself never calls these: it runs the fused kernels (gather*.cuh, mlp_*.cuh),
// which compute the same stages without materialising their outputs.
#pragma once
#include "common.cuh"

// feat (B,C,H,W) NCHW, uv (B,N,2) in [-1,1] -> out (B,N,C); grid_sample(bilinear, border, align_corners=True)
__global__ void k_feat_sample(const float* __restrict__ feat, int B, int C, int H, int W, const float* __restrict__ uv, int N,
                              float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N * C) return;
    const int c = (int)(i % C);
    const long long bn = i / C;
    const int b = (int)(bn / N);
    const Bilin t = bilin_setup(uv[2 * bn], uv[2 * bn + 1], W, H);
    const float* f = feat + ((size_t)b * C + c) * H * W;
    out[i] = bilin_mix(t, f[t.i00], t.i01 >= 0 ? f[t.i01] : 0.0f, t.i10 >= 0 ? f[t.i10] : 0.0f, t.i11 >= 0 ? f[t.i11] : 0.0f);
}

// nearest of Nv points for each of N queries: squared distance ((dx*dx + dy*dy) + dz*dz), ties -> lowest index
__global__ void k_knn1(const float* __restrict__ q, int N, const float* __restrict__ vert, int Nv, int* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
    float best = __int_as_float(0x7f800000);
    int bi = 0;
    for (int j = 0; j < Nv; ++j) {
        const float dx = xsub(x, vert[3 * j]), dy = xsub(y, vert[3 * j + 1]), dz = xsub(z, vert[3 * j + 2]);
        const float d = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
        if (d < best) { best = d; bi = j; }
    }
    idx[i] = bi;
}

enum StageAct { SA_NONE = 0, SA_RELU, SA_SOFTPLUS, SA_SIGMOID, SA_ELU };
__device__ __forceinline__ float stage_act(float v, int act) {
    switch (act) {
        case SA_RELU: return fmaxf(v, 0.0f);
        case SA_SOFTPLUS: return 100.0f * v > 20.0f ? v : log1pf(expf(100.0f * v)) / 100.0f;      // Softplus(beta=100, threshold=20)
        case SA_SIGMOID: return 1.0f / (1.0f + expf(-v));
        case SA_ELU: return v > 0.0f ? v : expm1f(v);
        default: return v;
    }
}
// y (M,N) = act(x (M,K) @ w (N,K)^T + b); one thread per output element, k ascending
__global__ void k_dense(const float* __restrict__ x, int M, int K, const float* __restrict__ w, const float* __restrict__ b, int N, int act,
                        float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    const int n = (int)(i % N);
    const long long m = i / N;
    const float* xr = x + m * K;
    const float* wr = w + (size_t)n * K;
    float acc = 0.0f;
    for (int k = 0; k < K; ++k) acc = fmaf(xr[k], wr[k], acc);
    if (b) acc += b[n];
    y[i] = stage_act(acc, act);
}

// SpatialEncoder rel_z_decay: cxyz (BV,N,3) camera-space samples, kxyz (BV,Kp,3) camera-space keypoints ->
// out (BV,N,(1 + 2L) Kp), row-major (function, keypoint): [dz | sin(pi dz) | cos(pi dz) | sin(2 pi dz) | ...] * w
__global__ void k_rel_z_decay(const float* __restrict__ cxyz, const float* __restrict__ kxyz, int BV, int N, int Kp, int L, float scale,
                              float sigma, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)BV * N * Kp) return;
    const int k = (int)(i % Kp);
    const long long bn = i / Kp;
    const int bv = (int)(bn / N);
    const float* c = cxyz + 3 * bn;
    const float* kp = kxyz + ((size_t)bv * Kp + k) * 3;
    const float dx = c[0] - kp[0], dy = c[1] - kp[1], dzr = c[2] - kp[2];
    const float wgt = expf(-(dx * dx + dy * dy + dzr * dzr) / (2.0f * (sigma * sigma)));
    const float dz = scale * dzr;
    float* o = out + bn * (size_t)((1 + 2 * L) * Kp);
    o[k] = dz * wgt;
    float f = 3.14159274101257324f;                       // np.float32(np.pi), doubled per level (pe_vector)
    for (int l = 0; l < L; ++l) {
        const float a = dz * f;
        o[(size_t)(1 + 2 * l) * Kp + k] = sinf(a) * wgt;
        o[(size_t)(2 + 2 * l) * Kp + k] = cosf(a) * wgt;
        f *= 2.0f;
    }
}
