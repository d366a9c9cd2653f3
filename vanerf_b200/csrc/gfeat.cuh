// Per-frame TexVisFusion global vertex feature as kernels (reference src/networks.py:246-279):
//   gf_img = fconv4(img)      Conv2d(3,21,3) - LayerNorm([H,W]) - ReLU - Conv2d(21,42,3) - LayerNorm([H,W]) - ReLU - AdaptiveAvgPool2d(3)
//   gf_tex = fconv3(feat_tex) Conv2d(8,21,3) - ... same on the (H/4, W/4) map
//   g = cat([gf_img (42,9), gf_tex (42,9)], -1) -> (42,18)
//   fconv_gt(g)               Conv1d(42,779,3) - LayerNorm(18) - ReLU - Conv1d(779,1558,3) - LayerNorm(18) - ReLU -> (1558,18)
// once per frame and source view.  The reference (and round 1 of this repo) leaves these to cuDNN: ~2.2 ms of small library
// launches per frame, two thirds of the per-frame setup that every rank of a multi-GPU render repeats.  Here: direct fp32
// convolutions (FFMA; 9.7 GFLOP per frame at V = 3) with the LayerNorm statistics reduced in the producing kernel, the
// normalisation + ReLU applied while the next convolution stages its input tile, and the 3x3 adaptive average pool fused with
// the second normalisation.  All convolutions are bias-free (src/networks.py:238-262), LayerNorm eps = 1e-6.
#pragma once
#include "common.cuh"

#define GF_TILE 16                      // output pixels per block edge
#define GF_CCHUNK 7                     // input channels staged per round (3, 8 -> one / two rounds; 21 -> three)
#define GF_MID 21
#define GF_OUT 42
#define GF_POOL_CHUNKS 8                // row chunks per (view, channel) plane in k_gf_norm_pool

// y (V,COUT,H,W) = conv3x3(pad 1) of x (V,CIN,H,W).  NORM: x is first normalised per (view, channel) with the LayerNorm
// statistics in `st_in` (sum, sum of squares over H*W, double) and the affine maps lnw / lnb (H,W), then ReLU; padding is zero
// AFTER that (the reference pads the activated tensor).  Per (view, output channel) sum and sum of squares of y: every block
// writes its partial sums to `st_part` (view, block, 2 COUT) and k_gf_reduce adds them up in a fixed order, so the statistics
// - and with them every bit of the frame's output - do not depend on the order in which blocks finish (no atomics).
template <int CIN, int COUT, bool NORM>
__global__ void __launch_bounds__(GF_TILE * GF_TILE)
k_gf_conv3x3(const float* __restrict__ x, const float* __restrict__ w, int H, int W, const double* __restrict__ st_in,
             const float* __restrict__ lnw, const float* __restrict__ lnb, float* __restrict__ y, double* __restrict__ st_part) {
    __shared__ float tile[GF_CCHUNK][GF_TILE + 2][GF_TILE + 2];
    __shared__ __align__(16) float wsm[GF_CCHUNK][9][((COUT + 3) / 4) * 4];
    __shared__ float red[2 * COUT][GF_TILE * GF_TILE / 32];
    const int v = blockIdx.z, tx = threadIdx.x % GF_TILE, ty = threadIdx.x / GF_TILE;
    const int x0 = blockIdx.x * GF_TILE, y0 = blockIdx.y * GF_TILE;
    const int px = x0 + tx, py = y0 + ty;
    const size_t HW = (size_t)H * W;
    __shared__ float s_mu[CIN], s_rstd[CIN];
    if (NORM && threadIdx.x < CIN) {
        const double s = st_in[2 * ((size_t)v * CIN + threadIdx.x)], q = st_in[2 * ((size_t)v * CIN + threadIdx.x) + 1];
        const double mean = s / (double)HW, var = q / (double)HW - mean * mean;
        s_mu[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + 1e-6));
    }
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.0f;
    for (int c0 = 0; c0 < CIN; c0 += GF_CCHUNK) {
        const int nc = min(GF_CCHUNK, CIN - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * (GF_TILE + 2) * (GF_TILE + 2); i += blockDim.x) {
            const int c = i / ((GF_TILE + 2) * (GF_TILE + 2)), r = i % ((GF_TILE + 2) * (GF_TILE + 2));
            const int yy = y0 + r / (GF_TILE + 2) - 1, xx = x0 + r % (GF_TILE + 2) - 1;
            float val = 0.0f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                val = x[((size_t)v * CIN + c0 + c) * HW + (size_t)yy * W + xx];
                if (NORM) val = fmaxf((val - s_mu[c0 + c]) * s_rstd[c0 + c] * lnw[(size_t)yy * W + xx] + lnb[(size_t)yy * W + xx], 0.0f);
            }
            tile[c][r / (GF_TILE + 2)][r % (GF_TILE + 2)] = val;
        }
        for (int i = threadIdx.x; i < nc * 9 * COUT; i += blockDim.x) {
            const int co = i % COUT, t = (i / COUT) % 9, c = i / (COUT * 9);
            wsm[c][t][co] = w[((size_t)co * CIN + c0 + c) * 9 + t];
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float xv = tile[c][ty + t / 3][tx + t % 3];
#pragma unroll
                for (int co = 0; co < COUT; ++co) acc[co] = fmaf(xv, wsm[c][t][co], acc[co]);
            }
        }
    }
    const bool live = px < W && py < H;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        const float a = live ? acc[co] : 0.0f;
        if (live) y[((size_t)v * COUT + co) * HW + (size_t)py * W + px] = a;
        float s = a, q = a * a;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (lane == 0) { red[2 * co][wid] = s; red[2 * co + 1][wid] = q; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * COUT) {
        double t = 0.0;
        for (int k = 0; k < GF_TILE * GF_TILE / 32; ++k) t += (double)red[threadIdx.x][k];
        const size_t nb = (size_t)gridDim.x * gridDim.y, b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        st_part[((size_t)v * nb + b) * (2 * COUT) + threadIdx.x] = t;       // [2 co] = sum, [2 co + 1] = sum of squares
    }
}

// out[g][i] = sum_k part[g][k][i] (g = blockIdx.y): the deterministic second stage of the block-partial reductions.  One warp
// per output: lane l adds parts l, l + 32, ... in ascending order, then a fixed shuffle tree combines the lanes, so the result
// depends on nothing but the values.
template <typename T>
__global__ void k_gf_reduce(const T* __restrict__ part, int n_parts, int n, T* __restrict__ out) {
    const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const T* p = part + (size_t)blockIdx.y * n_parts * n;
    T t = 0;
    for (int k = lane; k < n_parts; k += 32) t += p[(size_t)k * n + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) out[(size_t)blockIdx.y * n + i] = t;
}

// pooled_part (chunk,V,C,9) = this row chunk's share of AdaptiveAvgPool2d(3)(ReLU(LayerNorm[H,W](y))): one block per (channel,
// view, row chunk); k_gf_reduce adds the chunks.  Region i covers rows floor(i H / 3) .. ceil((i + 1) H / 3) (torch's
// definition: neighbouring regions overlap when H % 3 != 0).
__global__ void __launch_bounds__(256) k_gf_norm_pool(const float* __restrict__ y, int C, int H, int W, const double* __restrict__ st,
                                                      const float* __restrict__ lnw, const float* __restrict__ lnb, float* __restrict__ pooled) {
    const int c = blockIdx.x, v = blockIdx.y;
    const size_t HW = (size_t)H * W;
    const double s = st[2 * ((size_t)v * C + c)], q = st[2 * ((size_t)v * C + c) + 1];
    const double mean = s / (double)HW, var = q / (double)HW - mean * mean;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + 1e-6));
    const float* yp = y + ((size_t)v * C + c) * HW;
    int r0[3], r1[3], c0[3], c1[3];
    for (int i = 0; i < 3; ++i) {
        r0[i] = (i * H) / 3; r1[i] = ((i + 1) * H + 2) / 3;
        c0[i] = (i * W) / 3; c1[i] = ((i + 1) * W + 2) / 3;
    }
    float acc[9];
    for (int k = 0; k < 9; ++k) acc[k] = 0.0f;
    const int rows = (H + gridDim.z - 1) / gridDim.z;
    const size_t i_begin = (size_t)blockIdx.z * rows * W, i_end = min(HW, (size_t)(blockIdx.z + 1) * rows * W);
    for (size_t i = i_begin + threadIdx.x; i < i_end; i += blockDim.x) {
        const int yy = (int)(i / W), xx = (int)(i % W);
        const float a = fmaxf((yp[i] - mu) * rstd * lnw[i] + lnb[i], 0.0f);
#pragma unroll
        for (int ri = 0; ri < 3; ++ri)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
                if (yy >= r0[ri] && yy < r1[ri] && xx >= c0[ci] && xx < c1[ci]) acc[3 * ri + ci] += a;
    }
    __shared__ float red[9][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < 9; ++k) {
        float t = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[k][wid] = t;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        float t = 0.0f;
        for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
        const int ri = threadIdx.x / 3, ci = threadIdx.x % 3;
        pooled[(((size_t)blockIdx.z * gridDim.y + v) * C + c) * 9 + threadIdx.x] = t / (float)((r1[ri] - r0[ri]) * (c1[ci] - c0[ci]));
    }
}

// fconv_gt: out (V,COUT,18) = ReLU(LayerNorm(18)(Conv1d(CIN,COUT,3,pad 1)(g))), g (V,CIN,18) = cat([g_img (V,CIN,9) | g_tex (V,CIN,9)])
// for the first layer (g_tex != NULL) or a plain (V,CIN,18) tensor.  One warp per (view, output channel): lanes stride the input
// channels (coalesced weight rows), 18 partial sums per lane, warp reduction.
__global__ void __launch_bounds__(128) k_gf_conv1d_ln(const float* __restrict__ g_a, const float* __restrict__ g_b, const float* __restrict__ w, int V,
                                                      int CIN, int COUT, const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                      float* __restrict__ out) {
    const int gw = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (gw >= V * COUT) return;
    const int co = gw % COUT, v = gw / COUT;
    float acc[18];
#pragma unroll
    for (int l = 0; l < 18; ++l) acc[l] = 0.0f;
    for (int ci = lane; ci < CIN; ci += 32) {
        float row[20];
        row[0] = 0.0f; row[19] = 0.0f;
        if (g_b) {
#pragma unroll
            for (int l = 0; l < 9; ++l) { row[1 + l] = g_a[((size_t)v * CIN + ci) * 9 + l]; row[10 + l] = g_b[((size_t)v * CIN + ci) * 9 + l]; }
        } else {
#pragma unroll
            for (int l = 0; l < 18; ++l) row[1 + l] = g_a[((size_t)v * CIN + ci) * 18 + l];
        }
        const float* wr = w + ((size_t)co * CIN + ci) * 3;
        const float w0 = wr[0], w1 = wr[1], w2 = wr[2];
#pragma unroll
        for (int l = 0; l < 18; ++l) acc[l] = fmaf(w2, row[l + 2], fmaf(w1, row[l + 1], fmaf(w0, row[l], acc[l])));
    }
    float mean = 0.0f;
#pragma unroll
    for (int l = 0; l < 18; ++l) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[l] += __shfl_xor_sync(0xffffffffu, acc[l], o);
        mean += acc[l];
    }
    mean /= 18.0f;
    float var = 0.0f;
#pragma unroll
    for (int l = 0; l < 18; ++l) var += (acc[l] - mean) * (acc[l] - mean);
    const float rstd = 1.0f / sqrtf(var / 18.0f + 1e-6f);
    if (lane < 18) {
        float o = 0.0f;
#pragma unroll
        for (int l = 0; l < 18; ++l) if (l == lane) o = fmaxf((acc[l] - mean) * rstd * lnw[l] + lnb[l], 0.0f);
        out[((size_t)v * COUT + co) * 18 + lane] = o;
    }
}
