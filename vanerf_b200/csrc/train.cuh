// Kernels of the training branch (BASELINE.json configs[4]; reference: the `net.training` paths of src/model.py:748-957,
// :1103-1422 and their autograd).  The training step keeps the path's own kernels for everything that is not a dense
// layer: ray / depth sampling and the mesh queries (rays.cuh, geom.cuh: no gradient flows through them, the reference wraps
// them in no_grad), the projection / mask / boundary-weight stage below, the bilinear gathers with their scatter-add
// backward, and alpha compositing with its hand-written backward.  The dense layers of the unfused training graph are plain
// library GEMMs (torch.nn.functional.linear -> cuBLAS) under torch autograd (vanerf_b200/train.py).
#pragma once
#include "common.cuh"
#include "gather.cuh"
#include "composite.cuh"

// Per sample and source view, everything VANeRF.query derives from the sample position alone (src/model.py:780-821,
// :936-946, src/spatial.py:71-72), with the same device functions as the fused gather kernel (exact-op masks):
//   xy (V,N,2) normalised image coordinates; mask (N) = all-views AND of in-frustum and foreground tests; pw_raw (V,N) =
//   product of the three boundary sigmoids (before the mask and the normalisation over views); cam (V,N,3) camera-space
//   position; rd (V,N,4) ray difference [unit(view - s), s . view].
__global__ void k_project_samples(FrameDev fr, TargetDev tar, const float* __restrict__ rays, const float* __restrict__ z, int S, long long N,
                                  float* __restrict__ xy, unsigned char* __restrict__ mask, float* __restrict__ pw_raw,
                                  float* __restrict__ cam, float* __restrict__ rd) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* ray = rays + (size_t)(n / S) * VANERF_RAY_STRIDE;
    float p[3];
    sample_point(ray, tar.cam_pos, z[n], p);
    bool m = true;
    for (int v = 0; v < fr.V; ++v) {
        ViewProj pr = project_sample(fr, v, p);
        const Bilin b = bilin_setup(pr.x, pr.y, fr.W, fr.H);
        const float* mp = fr.imgm + (size_t)v * fr.H * fr.W * 4 + 3;
        const float fgv = bilin_mix(b, mp[(size_t)b.i00 * 4], b.i01 >= 0 ? mp[(size_t)b.i01 * 4] : 0.f,
                                    b.i10 >= 0 ? mp[(size_t)b.i10 * 4] : 0.f, b.i11 >= 0 ? mp[(size_t)b.i11 * 4] : 0.f);
        m = m && pr.in && (fgv > 0.1f);
        float w = 1.0f;
        const float q3[3] = {pr.x, pr.y, pr.zn};
        for (int c = 0; c < 3; ++c) {
            const float q = 0.5f * q3[c] + 0.5f;
            const float d = fminf(q, 1.0f - q);
            w *= sigmoidf_(5.0f * (d / 0.1f - 1.0f));
        }
        const size_t i = (size_t)v * N + n;
        xy[2 * i] = pr.x; xy[2 * i + 1] = pr.y;
        pw_raw[i] = w;
        const float* E = fr.extrin[v];
        cam[3 * i] = xaffine(E, 0, p[0], p[1], p[2]); cam[3 * i + 1] = xaffine(E, 1, p[0], p[1], p[2]); cam[3 * i + 2] = xaffine(E, 2, p[0], p[1], p[2]);
        float s0 = p[0] - fr.src_pos[v][0], s1 = p[1] - fr.src_pos[v][1], s2 = p[2] - fr.src_pos[v][2];
        const float inv = 1.0f / fmaxf(sqrtf(s0 * s0 + s1 * s1 + s2 * s2), 1e-12f);
        s0 *= inv; s1 *= inv; s2 *= inv;
        const float e0 = ray[0] - s0, e1 = ray[1] - s1, e2 = ray[2] - s2;
        const float ninv = 1.0f / fmaxf(sqrtf(e0 * e0 + e1 * e1 + e2 * e2), 1e-6f);
        rd[4 * i] = e0 * ninv; rd[4 * i + 1] = e1 * ninv; rd[4 * i + 2] = e2 * ninv; rd[4 * i + 3] = s0 * ray[0] + s1 * ray[1] + s2 * ray[2];
    }
    mask[n] = m ? 1 : 0;
}

// Backward of feat_sample (grid_sample bilinear / border / align_corners=True) with respect to the map: scatter-add of
// d_out (B,N,C) into d_feat (B,C,H,W) with the four tap weights.  d_feat must be zeroed by the caller.
__global__ void k_feat_sample_bwd(const float* __restrict__ d_out, int B, int C, int H, int W, const float* __restrict__ uv, int N,
                                  float* __restrict__ d_feat) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N * C) return;
    const int c = (int)(i % C);
    const long long bn = i / C;
    const int b = (int)(bn / N);
    const Bilin t = bilin_setup(uv[2 * bn], uv[2 * bn + 1], W, H);
    float* f = d_feat + ((size_t)b * C + c) * H * W;
    const float g = d_out[i];
    atomicAdd(f + t.i00, g * t.nw);
    if (t.i01 >= 0) atomicAdd(f + t.i01, g * t.ne);
    if (t.i10 >= 0) atomicAdd(f + t.i10, g * t.sw);
    if (t.i11 >= 0) atomicAdd(f + t.i11, g * t.se);
}

// Backward of VANeRF.rgba2out (src/model.py:1465-1494), warp per ray.  With tau_s = sigma_s dist_s, T_s = exp(-sum_{k<s} tau_k)
// and w_s = (1 - exp(-tau_s)) T_s:  dL/dtau_s = g_s T_{s+1} - sum_{j>s} g_j w_j  (g_s = dL/dw_s), no division by (1 - c).
// Inputs: gradients of color (R,3), alpha (R), depth (R), sdf_out (R) (any may be NULL).  Outputs: d_rgba (R,S,5) and
// d_beta (one float, atomically accumulated; zeroed by the caller).
__global__ void __launch_bounds__(COMP_WARPS * 32)
k_composite_bwd(const float* __restrict__ rgba, const float* __restrict__ z, const float* __restrict__ mesh_sdf, int R, int S, float beta,
                const float* __restrict__ g_color, const float* __restrict__ g_alpha, const float* __restrict__ g_depth,
                const float* __restrict__ g_sdf, float* __restrict__ d_rgba, float* __restrict__ d_beta) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * COMP_WARPS;
    const int per = (S + 31) / 32;
    float dbeta_acc = 0.0f;
    for (int r0 = 0; r0 < R; r0 += n_warps) {
        const int r = r0 + warp0;
        const bool live = r < R;
        const int rr = live ? r : R - 1;
        const float* zr = z + (size_t)rr * S;
        const float* ar = rgba + (size_t)rr * S * 5;
        const float* sr = mesh_sdf + (size_t)rr * S;
        float c[COMP_MAX_PER_LANE], sig[COMP_MAX_PER_LANE], sg[COMP_MAX_PER_LANE], dst[COMP_MAX_PER_LANE], local = 1.0f;
        const int s0 = lane * per;
#pragma unroll
        for (int i = 0; i < COMP_MAX_PER_LANE; ++i) {
            const int s = s0 + i;
            c[i] = 0.0f; sig[i] = 0.0f; sg[i] = 0.0f; dst[i] = 0.0f;
            if (i < per && s < S) {
                const float a = ar[(size_t)s * 5] + sr[s];
                sg[i] = 1.0f / (1.0f + expf(a / beta));                  // sigmoid(-a / beta)
                sig[i] = sg[i] / beta;
                dst[i] = (s + 1 < S) ? (zr[s + 1] - zr[s]) : 1e10f;
                c[i] = 1.0f - expf(-sig[i] * dst[i]);
                local *= (1.0f - c[i]);
            }
        }
        float incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl *= t;
        }
        float T = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) T = 1.0f;
        // forward sums needed by the depth / sdf quotients
        float w[COMP_MAX_PER_LANE], Tn[COMP_MAX_PER_LANE], acc_a = 0.f, acc_z = 0.f, acc_s = 0.f;
#pragma unroll
        for (int i = 0; i < COMP_MAX_PER_LANE; ++i) {
            const int s = s0 + i;
            w[i] = 0.0f; Tn[i] = 0.0f;
            if (i < per && s < S) {
                w[i] = c[i] * T;
                T *= (1.0f - c[i]);
                Tn[i] = T;                                               // T_{s+1}
                acc_a += w[i]; acc_z += zr[s] * w[i]; acc_s += ar[(size_t)s * 5 + 1] * w[i];
            }
        }
        acc_a = warp_sum(acc_a); acc_z = warp_sum(acc_z); acc_s = warp_sum(acc_s);
        const float gc0 = g_color ? g_color[3 * (size_t)rr] : 0.f, gc1 = g_color ? g_color[3 * (size_t)rr + 1] : 0.f, gc2 = g_color ? g_color[3 * (size_t)rr + 2] : 0.f;
        const float ga = g_alpha ? g_alpha[rr] : 0.f, gd = g_depth ? g_depth[rr] : 0.f, gs = g_sdf ? g_sdf[rr] : 0.f;
        const float den = acc_a + 1e-8f;
        const float gz_w = gd / den, gs_w = gs / den;                    // d depth / d Z, d sdf / d Ssum
        const float ga_tot = ga - gd * acc_z / (den * den) - gs * acc_s / (den * den);
        // g_s = dL/dw_s and the suffix sums of g_j w_j
        float g[COMP_MAX_PER_LANE], lsum = 0.f;
#pragma unroll
        for (int i = 0; i < COMP_MAX_PER_LANE; ++i) {
            const int s = s0 + i;
            g[i] = 0.f;
            if (i < per && s < S) {
                const float* px = ar + (size_t)s * 5;
                g[i] = gc0 * px[2] + gc1 * px[3] + gc2 * px[4] + ga_tot + gz_w * zr[s] + gs_w * px[1];
                lsum += g[i] * w[i];
            }
        }
        float after = 0.f;                                               // sum over the lanes after this one
        {
            float inc = lsum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_down_sync(0xffffffffu, inc, o);
                if (lane + o < 32) inc += t;
            }
            after = inc - lsum;
        }
        float suffix = after;                                            // sum_{j > s} g_j w_j, walking this lane's samples backwards
#pragma unroll
        for (int i = COMP_MAX_PER_LANE - 1; i >= 0; --i) {
            const int s = s0 + i;
            if (i < per && s < S) {
                const float dtau = g[i] * Tn[i] - suffix;
                suffix += g[i] * w[i];
                const float dsig = dtau * dst[i];
                const float a = ar[(size_t)s * 5] + sr[s];
                const float dsg = sg[i] * (1.0f - sg[i]);                // derivative of sigmoid at u = -a / beta
                if (live) {
                    float* o = d_rgba + ((size_t)r * S + s) * 5;
                    o[0] = dsig * (-dsg / (beta * beta));
                    o[1] = gs_w * w[i];
                    o[2] = gc0 * w[i]; o[3] = gc1 * w[i]; o[4] = gc2 * w[i];
                    dbeta_acc += dsig * ((dsg * a / beta - sg[i]) / (beta * beta));
                }
            }
        }
    }
    dbeta_acc = warp_sum(dbeta_acc);
    if (lane == 0 && d_beta && dbeta_acc != 0.0f) atomicAdd(d_beta, dbeta_acc);
}
