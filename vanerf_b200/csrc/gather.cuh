// Fused projection / mask / pix_weight / bilinear feature gather kernel (north-star kernel 1).
// Replaces, per sample and source view: the projection and masks of VANeRF.query (src/model.py:780-821), the five
// feat_sample calls (src/model.py:800,826,829,906,919 -> src/utils.py:136-151), the three KNN_vis gathers
// (src/networks.py:27-33 <- :84,97,281), the camera-space transform of SpatialEncoder (src/spatial.py:71-72) and
// the ray-difference encoding of query_color (src/model.py:936-946).
//
// Mapping: 16 lanes (a half warp) own one sample; lane l owns float4 "units" l, l+16, ... of the 77-unit record of
// each (sample, view) row, so every tap of a 64-channel NHWC map is one fully coalesced 256-byte read by the half
// warp and every record row is written with coalesced 16-byte stores.  Per-view camera matrices and keypoints live
// in the kernel parameter block (constant bank).  Masks follow the exact-op contract (bit-exact vs the oracle).
#pragma once
#include "common.cuh"
#include "rays.cuh"

#define GATHER_THREADS 256

struct ViewProj {
    float x, y, zn;      // normalised image coordinates in [-1,1] and normalised depth
    bool in, fg;
};

__device__ __forceinline__ ViewProj project_sample(const FrameDev& fr, int v, const float* p) {
    const float* M = fr.KRT[v];
    const float hx = xaffine(M, 0, p[0], p[1], p[2]);
    const float hy = xaffine(M, 1, p[0], p[1], p[2]);
    const float hz = xaffine(M, 2, p[0], p[1], p[2]);
    ViewProj o;
    o.x = xsub(xmul(2.0f, xdiv(xdiv(hx, hz), (float)fr.W - 1.0f)), 1.0f);
    o.y = xsub(xmul(2.0f, xdiv(xdiv(hy, hz), (float)fr.H - 1.0f)), 1.0f);
    o.zn = xsub(xdiv(xmul(2.0f, xsub(hz, fr.znear)), fr.z_range), 1.0f);
    const float lo = -1.01f, hi = 1.01f;          // (float)(-1.0 - 1e-2), (float)(1.0 + 1e-2)
    o.in = (o.x >= lo) && (o.x <= hi) && (o.y >= lo) && (o.y <= hi) && (o.zn >= -1.0f);
    return o;
}

__device__ __forceinline__ float4 tap4(const float* __restrict__ map, int C, int c4, const Bilin& b) {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 v00 = *reinterpret_cast<const float4*>(map + (size_t)b.i00 * C + c4);
    const float4 v01 = b.i01 >= 0 ? *reinterpret_cast<const float4*>(map + (size_t)b.i01 * C + c4) : z4;
    const float4 v10 = b.i10 >= 0 ? *reinterpret_cast<const float4*>(map + (size_t)b.i10 * C + c4) : z4;
    const float4 v11 = b.i11 >= 0 ? *reinterpret_cast<const float4*>(map + (size_t)b.i11 * C + c4) : z4;
    return make_float4(bilin_mix(b, v00.x, v01.x, v10.x, v11.x), bilin_mix(b, v00.y, v01.y, v10.y, v11.y),
                       bilin_mix(b, v00.z, v01.z, v10.z, v11.z), bilin_mix(b, v00.w, v01.w, v10.w, v11.w));
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// rec: (n_samples_in_chunk * V, REC_STRIDE).  sample0 = first global sample index of this chunk.
// valid_out: optional (N) 0/1 = VANeRF.query's `valid`.
__global__ void __launch_bounds__(GATHER_THREADS)
k_gather(FrameDev fr, TargetDev tar, const float* __restrict__ rays, const float* __restrict__ z,
         const float* __restrict__ pts_in, const float* __restrict__ view_in, int S, long long sample0, int n_chunk, long long N_total, const float* __restrict__ sdf, const int* __restrict__ nn_vert,
         const unsigned char* __restrict__ qvis, float* __restrict__ rec, unsigned char* __restrict__ valid_out) {
    const int lane = threadIdx.x & 15;
    const int group = (blockIdx.x * GATHER_THREADS + threadIdx.x) >> 4;
    const int n_groups = (gridDim.x * GATHER_THREADS) >> 4;
    const int V = fr.V;
    for (int i = group; i < n_chunk; i += n_groups) {
        const long long n = sample0 + i;
        float p[3];
        const float* ray;
        if (pts_in) {            // explicit points + view directions (VANeRF.query called directly)
            p[0] = pts_in[3 * n]; p[1] = pts_in[3 * n + 1]; p[2] = pts_in[3 * n + 2];
            ray = view_in + 3 * n;
        } else {
            ray = rays + (size_t)(n / S) * VANERF_RAY_STRIDE;
            sample_point(ray, tar.cam_pos, z[n], p);
        }
        // ---- pass 1: projections, masks, smooth boundary weights (all views; lane-redundant, registers only)
        ViewProj pr[MAXV];
        float pw[MAXV];
        bool m = true;
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            if (v < V) {
                pr[v] = project_sample(fr, v, p);
                const Bilin b = bilin_setup(pr[v].x, pr[v].y, fr.W, fr.H);
                const float* mp = fr.imgm + (size_t)v * fr.H * fr.W * 4 + 3;
                const float fgv = bilin_mix(b, mp[(size_t)b.i00 * 4], b.i01 >= 0 ? mp[(size_t)b.i01 * 4] : 0.f,
                                            b.i10 >= 0 ? mp[(size_t)b.i10 * 4] : 0.f, b.i11 >= 0 ? mp[(size_t)b.i11 * 4] : 0.f);
                pr[v].fg = fgv > 0.1f;
                m = m && pr[v].in && pr[v].fg;
            }
        }
        const float mf = m ? 1.0f : 0.0f;
        float pw_sum = 0.f;
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            if (v < V) {
                float w = 1.0f;
                const float q3[3] = {pr[v].x, pr[v].y, pr[v].zn};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float q = 0.5f * q3[c] + 0.5f;
                    const float d = fminf(q, 1.0f - q);
                    w *= sigmoidf_(5.0f * (d / 0.1f - 1.0f));
                }
                pw[v] = w * mf;
                pw_sum += pw[v];
            }
        }
        if (valid_out && lane == 0) valid_out[n] = m ? 1 : 0;
        const int nn = nn_vert[n];
        const int tw = (nn + NUM_V_HAND) % (2 * NUM_V_HAND);
        const float sdfv = sdf[n];
        // ---- pass 2: gather, one record row per view
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            if (v >= V) break;
            float4* row = reinterpret_cast<float4*>(rec + ((size_t)i * V + v) * REC_STRIDE);
            const float x = pr[v].x, y = pr[v].y;
            const size_t vb = (size_t)v * fr.n_verts;
            for (int u = lane; u < REC_STRIDE / 4; u += 16) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u < 16) {
                    const Bilin b = bilin_setup(x, y, fr.g0w, fr.g0h);
                    o = tap4(fr.geo0 + (size_t)v * fr.g0h * fr.g0w * 64, 64, 4 * u, b);
                } else if (u < 32) {
                    o = *reinterpret_cast<const float4*>(fr.T64 + (vb + nn) * 64 + 4 * (u - 16));
                } else if (u < 48) {
                    o = *reinterpret_cast<const float4*>(fr.T64 + (vb + tw) * 64 + 4 * (u - 32));
                } else if (u < 50) {
                    const Bilin b = bilin_setup(x, y, fr.g1w, fr.g1h);
                    o = tap4(fr.geo1 + (size_t)v * fr.g1h * fr.g1w * 8, 8, 4 * (u - 48), b);
                } else if (u < 52) {
                    o = *reinterpret_cast<const float4*>(fr.T8 + (vb + nn) * 8 + 4 * (u - 50));
                } else if (u < 54) {
                    o = *reinterpret_cast<const float4*>(fr.T8 + (vb + tw) * 8 + 4 * (u - 52));
                } else if (u == 54) {
                    const Bilin b = bilin_setup(x, y, fr.W, fr.H);
                    o = tap4(fr.imgm + (size_t)v * fr.H * fr.W * 4, 4, 0, b);
                } else if (u < 57) {
                    const Bilin b = bilin_setup(x, y, fr.tw, fr.th);
                    o = tap4(fr.tex + (size_t)v * fr.th * fr.tw * 8, 8, 4 * (u - 55), b);
                } else if (u < 65) {
                    o = *reinterpret_cast<const float4*>(fr.Ttex + (vb + nn) * 32 + 4 * (u - 57));
                } else if (u < 73) {
                    o = *reinterpret_cast<const float4*>(fr.Ttex + (vb + tw) * 32 + 4 * (u - 65));
                } else if (u == 73) {
                    o = make_float4(sdfv, qvis[(size_t)v * N_total + n] ? 1.0f : 0.0f, fr.vis[vb + nn], fr.vis[vb + tw]);
                } else if (u == 74) {
                    const float* E = fr.extrin[v];
                    o = make_float4(xaffine(E, 0, p[0], p[1], p[2]), xaffine(E, 1, p[0], p[1], p[2]),
                                    xaffine(E, 2, p[0], p[1], p[2]), 0.f);
                } else if (u == 75) {
                    // ray difference (model.py:936-946): s = normalize(p - c_src); rd = [(view - s)/max(|.|,1e-6), s.view]
                    float s0 = p[0] - fr.src_pos[v][0], s1 = p[1] - fr.src_pos[v][1], s2 = p[2] - fr.src_pos[v][2];
                    const float inv = 1.0f / fmaxf(sqrtf(s0 * s0 + s1 * s1 + s2 * s2), 1e-12f);
                    s0 *= inv; s1 *= inv; s2 *= inv;
                    const float e0 = ray[0] - s0, e1 = ray[1] - s1, e2 = ray[2] - s2;
                    const float ninv = 1.0f / fmaxf(sqrtf(e0 * e0 + e1 * e1 + e2 * e2), 1e-6f);
                    o = make_float4(e0 * ninv, e1 * ninv, e2 * ninv, s0 * ray[0] + s1 * ray[1] + s2 * ray[2]);
                } else {   // u == 76
                    o = make_float4(pw[v] / (pw_sum + 1e-6f), mf, 0.f, 0.f);
                }
                row[u] = o;
            }
        }
    }
}
