// fp32 reference-accuracy shading kernel: SpatialEncoder + GeoVisFusion + MLPUNetFusion + TexVisFusion +
// IBRRenderingHead + eval_func for a tile of 64 samples x V views per CTA, everything between the gathered
// records and the (N,5) rgba rows staying in shared memory.
//   SpatialEncoder.forward (rel_z_decay)   src/spatial.py:59-117
//   GeoVisFusion.forward                   src/networks.py:75-106
//   MLPUNetFusion.forward                  src/utils.py:633-649 (MLPUNet :822-852, PoolModule :744-779, pool_ops :854-880)
//   ibr_compress_gfeat + TexVisFusion      src/model.py:921, src/networks.py:281-293
//   IBRRenderingHead.forward               src/model.py:1600-1636
//   eval_func                              src/model.py:1140-1160
// This is the "fp32 path" of the north star (tolerance 1e-3): plain FFMA register-tile GEMMs (4 rows x 8 columns per
// thread, activations k-major in shared memory, weights streamed from L2 in 16-row chunks).  The tensor-core path
// (mlp_tc.cu) shares the record format and the epilogue math.
#pragma once
#include "common.cuh"

#define MLP_THREADS 256
#define TS 64               // samples per tile
#define KC 16               // weight rows staged per chunk
#define WS_FLOATS (KC * 128)
static_assert(KC * 128 / 4 <= 2 * MLP_THREADS, "gemm_acc stages a weight chunk with two float4 per thread");

enum Act { ACT_NONE = 0, ACT_RELU, ACT_SOFTPLUS, ACT_SIGMOID, ACT_ELU };

__device__ __forceinline__ float act_apply(float x, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(x, 0.0f);
        case ACT_SOFTPLUS: { const float bx = 100.0f * x; return bx > 20.0f ? x : log1pf(expf(bx)) / 100.0f; }   // Softplus(beta=100, threshold=20)
        case ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
        case ACT_ELU: return x > 0.0f ? x : expm1f(x);
        default: return x;
    }
}

// acc[4][8] += A[K][64]^T (rows 4tx..4tx+3) * Wt[K][Npad] (cols 8ty..8ty+7).  All threads must call.
__device__ __forceinline__ void gemm_acc(float (&acc)[4][8], const float* A, int K, const float* __restrict__ Wt,
                                         int Npad, float* WS) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const bool active = 8 * ty < Npad;
    // Weight chunks (KC rows of Npad <= 128 floats = at most two float4 per thread) go global -> registers -> shared: the
    // loads of chunk i + 1 are issued before the FMAs of chunk i, so their L2 latency hides behind the arithmetic instead
    // of sitting between two block barriers in front of every chunk.
    float4 pre[2];
    auto fetch = [&](int k0) {
        const int n4 = (min(KC, K - k0) * Npad) >> 2;
        const float4* src = reinterpret_cast<const float4*>(Wt + (size_t)k0 * Npad);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = tid + j * MLP_THREADS;
            pre[j] = i < n4 ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += KC) {
        const int kc = min(KC, K - k0);
        const int n4 = (kc * Npad) >> 2;
        __syncthreads();                                   // readers of the previous chunk are done
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = tid + j * MLP_THREADS;
            if (i < n4) reinterpret_cast<float4*>(WS)[i] = pre[j];
        }
        __syncthreads();
        if (k0 + KC < K) fetch(k0 + KC);
        if (active) {
            for (int k = 0; k < kc; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(A + (size_t)(k0 + k) * TS + 4 * tx);
                const float4 w0 = *reinterpret_cast<const float4*>(WS + k * Npad + 8 * ty);
                const float4 w1 = *reinterpret_cast<const float4*>(WS + k * Npad + 8 * ty + 4);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
            }
        }
    }
}

__device__ __forceinline__ void acc_zero(float (&acc)[4][8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
}

// OUT[col][row] = act(acc + b[col]) for col < N
__device__ __forceinline__ void acc_store(const float (&acc)[4][8], const LayerDev& L, int act, float* OUT) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = 8 * ty + j;
        if (col < L.N) {
            const float b = L.b[col];
            float4 o;
            o.x = act_apply(acc[0][j] + b, act);
            o.y = act_apply(acc[1][j] + b, act);
            o.z = act_apply(acc[2][j] + b, act);
            o.w = act_apply(acc[3][j] + b, act);
            *reinterpret_cast<float4*>(OUT + (size_t)col * TS + 4 * tx) = o;
        }
    }
}

__device__ __forceinline__ void dense(const float* A, const LayerDev& L, int act, float* OUT, float* WS) {
    float acc[4][8];
    acc_zero(acc);
    gemm_acc(acc, A, L.K, L.wt, L.Npad, WS);
    acc_store(acc, L, act, OUT);
}

// per-view scalar rows in SC: [V][16][TS]
#define SC_SDF 0
#define SC_QVIS 1
#define SC_VN 2
#define SC_VT 3
#define SC_CAM 4
#define SC_RD 8
#define SC_PW 12
#define SC_MASK 13

__host__ __device__ inline size_t mlp_simt_smem_floats(int V) {
    return 12544 + 8192 + 5376 + 2688 + 2816 + WS_FLOATS + 1536 + (size_t)V * (4096 + 1024);
}

// rec: records of this chunk ((n_chunk*V) rows); outputs indexed by the global sample index sample0 + i.
__global__ void __launch_bounds__(MLP_THREADS, 1)
k_mlp_simt(const NetDev* __restrict__ netp, const float* __restrict__ kpt_cam, int V, const float* __restrict__ rec,
           long long sample0, int n_chunk, float* __restrict__ rgba, float* __restrict__ raw_out,
           float* __restrict__ dbg_latent) {
    DYN_SMEM(float, smem);
    const NetDev& net = *netp;
    float* X = smem;
    float* Y = X + 12544;
    float* PEB = Y + 8192;
    float* PECH = PEB + 5376;
    float* G8 = PECH + 2688;            // x28 [28][TS] | t8 [8][TS] | out8 [8][TS]
    float* WS = G8 + 2816;
    float* SM = WS + WS_FLOATS;         // at [16][TS] | gate [8][TS]
    float* H3 = SM + 1536;              // [V][64][TS]  (later: F [V][40][TS])
    float* SC = H3 + (size_t)V * 4096;  // [V][16][TS]
    float* AT = SM;
    float* GATE = SM + 16 * TS;

    const int tid = threadIdx.x;
    const int row = tid & 63, q = tid >> 6;
    const int n_tiles = (n_chunk + TS - 1) / TS;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int i0 = tile * TS;
        // =============================== per-view geometry branch ===============================
        for (int v = 0; v < V; ++v) {
            __syncthreads();
            {   // ---- load the geo part of the record rows (transposing): thread = (row lr, unit phase kq)
                const int lr = tid >> 2, kq = tid & 3;
                const int isamp = min(i0 + lr, n_chunk - 1);
                const float4* r4 = reinterpret_cast<const float4*>(rec + ((size_t)isamp * V + v) * REC_STRIDE);
                float* sc = SC + (size_t)v * 16 * TS;
                for (int u = kq; u < 54; u += 4) {
                    const float4 val = r4[u];
                    float* dst = (u < 48) ? (X + (size_t)(4 * u) * TS + lr) : (G8 + (size_t)(4 * (u - 48)) * TS + lr);
                    dst[0] = val.x; dst[TS] = val.y; dst[2 * TS] = val.z; dst[3 * TS] = val.w;
                }
                if (kq == 0) {
                    const float4 s = r4[REC_SDF / 4];
                    X[192 * TS + lr] = s.x; X[193 * TS + lr] = s.y; X[194 * TS + lr] = s.z; X[195 * TS + lr] = s.w;
                    G8[24 * TS + lr] = s.x; G8[25 * TS + lr] = s.y; G8[26 * TS + lr] = s.z; G8[27 * TS + lr] = s.w;
                    sc[SC_SDF * TS + lr] = s.x; sc[SC_QVIS * TS + lr] = s.y; sc[SC_VN * TS + lr] = s.z; sc[SC_VT * TS + lr] = s.w;
                } else if (kq == 1) {
                    const float4 c = r4[REC_CAM / 4];
                    sc[(SC_CAM + 0) * TS + lr] = c.x; sc[(SC_CAM + 1) * TS + lr] = c.y; sc[(SC_CAM + 2) * TS + lr] = c.z;
                } else if (kq == 2) {
                    const float4 d = r4[REC_RD / 4];
                    sc[(SC_RD + 0) * TS + lr] = d.x; sc[(SC_RD + 1) * TS + lr] = d.y; sc[(SC_RD + 2) * TS + lr] = d.z; sc[(SC_RD + 3) * TS + lr] = d.w;
                } else {
                    const float4 w = r4[REC_PW / 4];
                    sc[SC_PW * TS + lr] = w.x; sc[SC_MASK * TS + lr] = w.y;
                }
            }
            __syncthreads();
            {   // ---- keypoint-relative depth offsets and Gaussian weights (src/spatial.py:109-113)
                const float* sc = SC + (size_t)v * 16 * TS;
                const float cx = sc[(SC_CAM + 0) * TS + row], cy = sc[(SC_CAM + 1) * TS + row], cz = sc[(SC_CAM + 2) * TS + row];
                for (int kp = q; kp < NKPT; kp += 4) {
                    const float* kc = kpt_cam + ((size_t)v * NKPT + kp) * 3;
                    const float dx = cx - kc[0], dy = cy - kc[1], dz = cz - kc[2];
                    PEB[kp * TS + row] = dz;
                    PEB[(NKPT + kp) * TS + row] = expf(-(dx * dx + dy * dy + dz * dz) / 0.02f);
                }
            }
            // ---- GeoVisFusion, 64-channel scale (src/networks.py:83-94)
            dense(X, net.layer[L_GEO_AT0], ACT_RELU, AT, WS);
            dense(AT, net.layer[L_GEO_AT1], ACT_SIGMOID, GATE, WS);
            __syncthreads();
            for (int k = q; k < 192; k += 4) X[k * TS + row] *= GATE[(k >> 6) * TS + row];
            dense(X, net.layer[L_GEO_F0], ACT_RELU, Y, WS);
            dense(Y, net.layer[L_GEO_F1], ACT_NONE, Y + 64 * TS, WS);           // out64
            // ---- GeoVisFusion, 8-channel scale (src/networks.py:96-104)
            dense(G8, net.layer[L_GEO8_AT0], ACT_RELU, AT, WS);
            dense(AT, net.layer[L_GEO8_AT1], ACT_SIGMOID, GATE, WS);
            __syncthreads();
            for (int k = q; k < 24; k += 4) G8[k * TS + row] *= GATE[(k >> 3) * TS + row];
            dense(G8, net.layer[L_GEO8_F0], ACT_RELU, G8 + 28 * TS, WS);
            dense(G8 + 28 * TS, net.layer[L_GEO8_F1], ACT_NONE, G8 + 36 * TS, WS);   // out8
            // ---- MLPUNet layer 0: [PE 294 | out64] -> 128, PE generated 42 columns at a time (src/spatial.py:20-35)
            {
                float acc[4][8];
                acc_zero(acc);
                const LayerDev& L0 = net.layer[L_MLP0];
                for (int f = 0; f < 7; ++f) {
                    __syncthreads();
                    for (int kp = q; kp < NKPT; kp += 4) {
                        const float dz = PEB[kp * TS + row], w = PEB[(NKPT + kp) * TS + row];
                        float val;
                        if (f == 0) val = dz;
                        else {
                            const float freq = (f <= 2) ? 3.14159274f : ((f <= 4) ? 6.28318548f : 12.5663710f);   // float32(pi * 2^l)
                            const float ang = dz * freq;
                            val = (f & 1) ? sinf(ang) : cosf(ang);
                        }
                        PECH[kp * TS + row] = val * w;
                    }
                    gemm_acc(acc, PECH, NKPT, L0.wt + (size_t)f * NKPT * L0.Npad, L0.Npad, WS);
                }
                gemm_acc(acc, Y + 64 * TS, 64, L0.wt + (size_t)294 * L0.Npad, L0.Npad, WS);
                acc_store(acc, L0, ACT_SOFTPLUS, X);                                // h0 [128]
            }
            dense(X, net.layer[L_MLP1], ACT_SOFTPLUS, Y, WS);                       // h1 [128]
            {
                float acc[4][8];
                acc_zero(acc);
                const LayerDev& L2 = net.layer[L_MLP2];
                gemm_acc(acc, Y, 128, L2.wt, L2.Npad, WS);
                gemm_acc(acc, G8 + 36 * TS, 8, L2.wt + (size_t)128 * L2.Npad, L2.Npad, WS);
                acc_store(acc, L2, ACT_SOFTPLUS, X);                                // h2 [120]
            }
            dense(X, net.layer[L_MLP3], ACT_NONE, H3 + (size_t)v * 64 * TS, WS);    // h3 [64]
        }
        // =============================== view pooling + density head ===============================
        __syncthreads();
        for (int c = q; c < 64; c += 4) {                                           // pool_ops mean/var (src/utils.py:854-880)
            float mean = 0.f;
            for (int v = 0; v < V; ++v) mean += SC[((size_t)v * 16 + SC_PW) * TS + row] * H3[((size_t)v * 64 + c) * TS + row];
            float var = 0.f;
            for (int v = 0; v < V; ++v) {
                const float d = H3[((size_t)v * 64 + c) * TS + row] - mean;
                var += SC[((size_t)v * 16 + SC_PW) * TS + row] * d * d;
            }
            X[c * TS + row] = mean;
            X[(64 + c) * TS + row] = var;
        }
        __syncthreads();
        if (dbg_latent) {
            const int isamp = i0 + row;
            if (isamp < n_chunk)
                for (int c = q; c < 128; c += 4) dbg_latent[(size_t)(sample0 + isamp) * 128 + c] = X[c * TS + row];
        }
        float* LAT24 = PEB;                    // [24][TS]
        float* OO = PEB + 24 * TS;             // [8][TS]: o0, o1
        dense(X, net.layer[L_COMPRESS], ACT_NONE, LAT24, WS);
        dense(X, net.layer[L_POST0], ACT_SOFTPLUS, Y, WS);
        dense(Y, net.layer[L_POST1], ACT_SOFTPLUS, Y + 64 * TS, WS);
        dense(Y + 64 * TS, net.layer[L_POST2], ACT_NONE, OO, WS);
        // =============================== texture branch per view (TexVisFusion) ===============================
        float* F = H3;                         // [V][40][TS]   (H3 is dead: pooled)
        float* SRC = PEB + 32 * TS;            // [V][3][TS] source colours (rgb_feat[..., :3] before the ray encoder)
        float* SV = SRC + 3 * MAXV * TS;       // [V][TS] blending logits
        float* WT = SV + MAXV * TS;            // [V][TS] anisotropic weights
        for (int v = 0; v < V; ++v) {
            __syncthreads();
            {
                const int lr = tid >> 2, kq = tid & 3;
                const int isamp = min(i0 + lr, n_chunk - 1);
                const float* rr = rec + ((size_t)isamp * V + v) * REC_STRIDE;
                // y96 = [img3,tex8 | a11 | b11 | a18 | b18 | lat24 | qvis,vn,vt]  (src/networks.py:284-286)
                for (int k = kq; k < 96; k += 4) {
                    float val;
                    if (k < 3) val = rr[REC_QIMG + k];
                    else if (k < 11) val = rr[REC_QTEX + (k - 3)];
                    else if (k < 22) val = rr[REC_ATEX + (k - 11)];
                    else if (k < 33) val = rr[REC_BTEX + (k - 22)];
                    else if (k < 51) val = rr[REC_ATEX + 11 + (k - 33)];
                    else if (k < 69) val = rr[REC_BTEX + 11 + (k - 51)];
                    else if (k < 93) val = LAT24[(k - 69) * TS + lr];
                    else val = rr[REC_QVIS + (k - 93)];
                    X[k * TS + lr] = val;
                }
            }
            dense(X, net.layer[L_TEX_AT0], ACT_RELU, Y, WS);
            dense(Y, net.layer[L_TEX_AT1], ACT_SIGMOID, GATE, WS);
            __syncthreads();
            for (int k = q; k < 93; k += 4) {
                const int g = k < 11 ? 0 : k < 22 ? 1 : k < 33 ? 2 : k < 51 ? 3 : k < 69 ? 4 : 5;
                X[k * TS + row] *= GATE[g * TS + row];
            }
            dense(X, net.layer[L_TEX_F0], ACT_RELU, Y, WS);
            dense(Y, net.layer[L_TEX_F1], ACT_NONE, F + (size_t)v * 40 * TS, WS);
            __syncthreads();
            if (q < 3) SRC[((size_t)v * 3 + q) * TS + row] = F[((size_t)v * 40 + q) * TS + row];
            // ray encoder 4 -> 16 -> 40 (ELU), added to the features (src/model.py:1612-1618)
            dense(SC + ((size_t)v * 16 + SC_RD) * TS, net.layer[L_RAY0], ACT_ELU, Y, WS);
            dense(Y, net.layer[L_RAY1], ACT_ELU, Y + 16 * TS, WS);
            __syncthreads();
            for (int c = q; c < 40; c += 4) F[((size_t)v * 40 + c) * TS + row] += Y[(16 + c) * TS + row];
        }
        // =============================== IBRRenderingHead over the view axis ===============================
        __syncthreads();
        if (q == 0) {                           // anisotropic weights (src/model.py:1620-1623)
            float e[MAXV], emin = 3.0e38f, sum = 0.f;
            for (int v = 0; v < V; ++v) {
                e[v] = expf(net.ani_al_abs * (SC[((size_t)v * 16 + SC_RD + 3) * TS + row] - 1.0f));
                emin = fminf(emin, e[v]);
            }
            for (int v = 0; v < V; ++v) { e[v] = (e[v] - emin) * SC[((size_t)v * 16 + SC_MASK) * TS + row]; sum += e[v]; }
            for (int v = 0; v < V; ++v) WT[v * TS + row] = e[v] / (sum + 1e-8f);
        }
        __syncthreads();
        for (int c = q; c < 40; c += 4) {       // fused_mean_variance (src/utils.py:153-157) -> X[0:40], X[40:80]
            float mean = 0.f;
            for (int v = 0; v < V; ++v) mean += F[((size_t)v * 40 + c) * TS + row] * WT[v * TS + row];
            float var = 0.f;
            for (int v = 0; v < V; ++v) {
                const float d = F[((size_t)v * 40 + c) * TS + row] - mean;
                var += WT[v * TS + row] * d * d;
            }
            X[c * TS + row] = mean;
            X[(40 + c) * TS + row] = var;
        }
        for (int v = 0; v < V; ++v) {
            __syncthreads();
            const float* sc = SC + (size_t)v * 16 * TS;
            for (int c = q; c < 40; c += 4) X[(80 + c) * TS + row] = F[((size_t)v * 40 + c) * TS + row];
            dense(X, net.layer[L_BASE0], ACT_ELU, Y, WS);                      // 120 -> 64
            float* XV = X + 120 * TS;                                           // x  [32]
            float* T32 = X + 160 * TS;                                          // scratch [32]
            float* PV = Y + 64 * TS;                                            // [40] (33 used)
            dense(Y, net.layer[L_BASE1], ACT_ELU, XV, WS);                      // 64 -> 32
            __syncthreads();
            for (int c = q; c < 32; c += 4) T32[c * TS + row] = XV[c * TS + row] * WT[v * TS + row];
            dense(T32, net.layer[L_VIS1_0], ACT_ELU, Y, WS);                    // 32 -> 32
            dense(Y, net.layer[L_VIS1_1], ACT_ELU, PV, WS);                     // 32 -> 33
            __syncthreads();
            for (int c = q; c < 32; c += 4) {
                const float xn = XV[c * TS + row] + PV[c * TS + row];
                XV[c * TS + row] = xn;
                T32[c * TS + row] = xn * act_apply(PV[32 * TS + row], ACT_SIGMOID) * sc[SC_MASK * TS + row];
            }
            dense(T32, net.layer[L_VIS2_0], ACT_ELU, Y, WS);                    // 32 -> 32
            dense(Y, net.layer[L_VIS2_1], ACT_SIGMOID, Y + 32 * TS, WS);        // 32 -> 1
            __syncthreads();
            // out_layer input [x 32 | vis 1 | ray_diff 4] is contiguous at XV: X[120..156]
            if (q == 0) XV[32 * TS + row] = Y[32 * TS + row] * sc[SC_MASK * TS + row];
            else XV[(32 + q) * TS + row] = sc[(SC_RD + q - 1) * TS + row];
            if (q == 1) XV[36 * TS + row] = sc[(SC_RD + 3) * TS + row];
            dense(XV, net.layer[L_OUT0], ACT_ELU, Y, WS);                       // 37 -> 16
            dense(Y, net.layer[L_OUT1], ACT_ELU, Y + 16 * TS, WS);              // 16 -> 8
            dense(Y + 16 * TS, net.layer[L_OUT2], ACT_NONE, Y + 24 * TS, WS);   // 8 -> 1
            __syncthreads();
            if (q == 0) SV[v * TS + row] = (sc[SC_MASK * TS + row] == 0.0f) ? -1e4f : Y[24 * TS + row];
        }
        __syncthreads();
        if (q == 0 && i0 + row < n_chunk) {     // softmax blend of the source colours + eval_func (model.py:1634-1635, 1140-1160)
            float smax = -3.0e38f;
            for (int v = 0; v < V; ++v) smax = fmaxf(smax, SV[v * TS + row]);
            float den = 0.f, rgb[3] = {0.f, 0.f, 0.f};
            for (int v = 0; v < V; ++v) {
                const float e = expf(SV[v * TS + row] - smax);
                den += e;
                for (int c = 0; c < 3; ++c) rgb[c] += SRC[((size_t)v * 3 + c) * TS + row] * e;
            }
            const float inv = 1.0f / den;
            const float o0 = OO[row], o1 = OO[TS + row];
            const float m = SC[SC_MASK * TS + row];          // out_mask is identical for all views; valid = sum > 0
            const size_t n = (size_t)(sample0 + i0 + row);
            if (raw_out) {
                float* o = raw_out + n * 5;
                o[0] = o0; o[1] = o1; o[2] = rgb[0] * inv; o[3] = rgb[1] * inv; o[4] = rgb[2] * inv;
            }
            if (rgba) {
                float* o = rgba + n * 5;
                o[0] = m * fmaxf(o1, 0.0f);
                o[1] = m * o0 + (1.0f - m) * 0.001f;
                o[2] = rgb[0] * inv; o[3] = rgb[1] * inv; o[4] = rgb[2] * inv;
            }
        }
    }
}
