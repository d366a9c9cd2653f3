// Gather kernel of the tensor-core path (north-star kernel 1, bf16 variant).  Same per-sample work as k_gather
// (projection + masks + pix_weight of VANeRF.query src/model.py:780-821, the feat_sample taps src/utils.py:136-151,
// the KNN_vis rows src/networks.py:27-33, camera-space position src/spatial.py:71-72, ray difference
// src/model.py:936-946) but
//   * reads bf16 NHWC maps and bf16 vertex tables (built once per frame): one bilinear tap of the 64-channel map is
//     one 128-byte line, one lane moves one 16-byte chunk (8 channels);
//   * writes, per (tile of 128 samples, view), five 16 KB operand images in exactly the K-major 128-byte-swizzle
//     layout the tcgen05 MMAs of k_mlp_tc consume (the MLP kernel brings them in with one bulk copy per image),
//     plus a 64-byte fp32 side record per (sample, view).
// Masks / `valid` keep the exact-op contract (bit-exact vs the oracle); features are rounded to bf16 once.
#pragma once
#include <cuda_bf16.h>
#include <cstdio>
#include "common.cuh"
#include "gather.cuh"
#include "tc_prims.cuh"

#define TC_ROWS 128
#define TC_SLOT 16384
#define TC_REC_IMAGES 5                 // R0 px64 | R1 a64 | R2 b64 | R3 misc | R4 tex
#define TC_AUX_BYTES 64                 // cam xyz, rd0 | rd1..3, pw | mask,0,0,0 | 8 bf16 extras

// ---- per-frame bf16 copies ------------------------------------------------------------------------------------
__global__ void k_f32_to_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}
// Ttex (V*Nv, 32) fp32 [img3 | tex8 | gf18 | 0,0,0] -> (V*Nv, 32) bf16 [tex8 | gf0-7 | gf8-15 | gf16, gf17, img r,g,b, 0,0,0]
__global__ void k_ttex_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int n_rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * 32) return;
    const int r = i >> 5, c = i & 31;
    int s;
    if (c < 8) s = 3 + c;
    else if (c < 26) s = 11 + (c - 8);
    else if (c < 29) s = c - 26;
    else s = 31;
    dst[i] = __float2bfloat16_rn(src[r * 32 + s]);
}

struct FrameTc {                       // bf16 companions of FrameDev
    const __nv_bfloat16* geo0;         // (V,h,w,64)
    const __nv_bfloat16* geo1;         // (V,h,w,8)
    const __nv_bfloat16* tex;          // (V,h,w,8)
    const __nv_bfloat16* T64;          // (V,Nv,64)
    const __nv_bfloat16* T8;           // (V,Nv,8)
    const __nv_bfloat16* Ttex;         // (V,Nv,32) chunk order, see k_ttex_bf16
};

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(tc::pack_bf16(f[0], f[1]), tc::pack_bf16(f[2], f[3]), tc::pack_bf16(f[4], f[5]), tc::pack_bf16(f[6], f[7]));
}
// bilinear tap of 8 consecutive bf16 channels (one 16-byte chunk per corner); fp32 mix, same tap order as bilin_mix
__device__ __forceinline__ uint4 tap8_bf16(const __nv_bfloat16* __restrict__ map, int C, int c8, const Bilin& b) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    const uint4 q00 = *reinterpret_cast<const uint4*>(map + (size_t)b.i00 * C + c8);
    const uint4 q01 = b.i01 >= 0 ? *reinterpret_cast<const uint4*>(map + (size_t)b.i01 * C + c8) : z;
    const uint4 q10 = b.i10 >= 0 ? *reinterpret_cast<const uint4*>(map + (size_t)b.i10 * C + c8) : z;
    const uint4 q11 = b.i11 >= 0 ? *reinterpret_cast<const uint4*>(map + (size_t)b.i11 * C + c8) : z;
    float a[8], c[8], d[8], e[8], o[8];
    unpack8(q00, a); unpack8(q01, c); unpack8(q10, d); unpack8(q11, e);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(e[i], b.se, fmaf(d[i], b.sw, fmaf(c[i], b.ne, a[i] * b.nw)));
    return pack8(o);
}

#ifndef GTC_THREADS
#define GTC_THREADS 256
#endif
#ifndef GTC_MINB
#define GTC_MINB 4                      // resident CTAs per SM the register allocation aims for
#endif

// rec: (n_tiles, V, 5, 16 KB) operand images; aux: (n_tiles*128, V, 64 B).  Rows past n_chunk in the last tile are
// written as copies of the last sample (finite values; their outputs are never stored).
__global__ void __launch_bounds__(GTC_THREADS, GTC_MINB)
k_gather_tc(FrameDev fr, FrameTc ft, TargetDev tar, const float* __restrict__ rays, const float* __restrict__ z,
            const float* __restrict__ pts_in, const float* __restrict__ view_in, int S, long long sample0, int n_chunk,
            long long N_total, const float* __restrict__ sdf, const int* __restrict__ nn_vert,
            const unsigned char* __restrict__ qvis, unsigned char* __restrict__ rec, unsigned char* __restrict__ aux,
            unsigned char* __restrict__ valid_out) {
    // A warp handles 8 samples per iteration.  Pass 1 (per (sample, view): projection, masks, boundary weight, image and
    // foreground taps, the tap set of the 64-channel map; exact-op sequences) runs on lane = 4 * sample + view, i.e. on 24
    // of 32 lanes at V = 3 (it used to run on 3 lanes of each 8-lane sample group: 12 of 32).  Pass 2 (the wide copies,
    // 8 lanes x 16 bytes per sample) takes the 8 samples in two rounds of four and reads pass 1 by shuffle.
    const int lane32 = threadIdx.x & 31;
    const int lane = lane32 & 7;
    const int warp_g = (blockIdx.x * GTC_THREADS + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * GTC_THREADS) >> 5;
    const int V = fr.V;
    const int n_rows = ((n_chunk + TC_ROWS - 1) / TC_ROWS) * TC_ROWS;          // multiple of 8
    for (int i8 = warp_g * 8; i8 < n_rows; i8 += n_warps * 8) {
        // ---- pass 1
        float mx = 0.f, my = 0.f, mw = 0.f, mcr = 0.f, mcg = 0.f, mcb = 0.f;
        int mok = 1;
        Bilin mb;
        mb.i00 = mb.i01 = mb.i10 = mb.i11 = 0; mb.nw = mb.ne = mb.sw = mb.se = 0.f;
        {
            const int i1 = i8 + (lane32 >> 2), v = lane32 & 3;
            const long long n1 = sample0 + (i1 < n_chunk ? i1 : n_chunk - 1);
            if (v < V) {
                float p[3];
                if (pts_in) { p[0] = pts_in[3 * n1]; p[1] = pts_in[3 * n1 + 1]; p[2] = pts_in[3 * n1 + 2]; }
                else sample_point(rays + (size_t)(n1 / S) * VANERF_RAY_STRIDE, tar.cam_pos, z[n1], p);
                const ViewProj q = project_sample(fr, v, p);
                const Bilin b = bilin_setup(q.x, q.y, fr.W, fr.H);
                // the image tap (r, g, b) reads the same four texels as the foreground-mask tap: do both here
                const float4* ip = reinterpret_cast<const float4*>(fr.imgm) + (size_t)v * fr.H * fr.W;
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef GTC_DEBUG
                if (b.i00 < 0 || b.i00 >= fr.H * fr.W || b.i01 >= fr.H * fr.W || b.i10 >= fr.H * fr.W || b.i11 >= fr.H * fr.W || n1 < 0 || n1 >= N_total)
                    printf("GTC_DEBUG pass1 blk %d thr %d i1 %d n1 %lld v %d idx %d %d %d %d\n", blockIdx.x, threadIdx.x, i1, n1, v, b.i00, b.i01, b.i10, b.i11);
#endif
                const float4 t00 = ip[b.i00], t01 = b.i01 >= 0 ? ip[b.i01] : z4, t10 = b.i10 >= 0 ? ip[b.i10] : z4, t11 = b.i11 >= 0 ? ip[b.i11] : z4;
                const float fgv = bilin_mix(b, t00.w, t01.w, t10.w, t11.w);
                mcr = bilin_mix(b, t00.x, t01.x, t10.x, t11.x);
                mcg = bilin_mix(b, t00.y, t01.y, t10.y, t11.y);
                mcb = bilin_mix(b, t00.z, t01.z, t10.z, t11.z);
                mok = (q.in && fgv > 0.1f) ? 1 : 0;
                float w = 1.0f;
                const float q3[3] = {q.x, q.y, q.zn};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float qq = 0.5f * q3[c] + 0.5f;
                    const float d = fminf(qq, 1.0f - qq);
                    w *= sigmoidf_(5.0f * (d / 0.1f - 1.0f));
                }
                mx = q.x; my = q.y; mw = w;
                mb = bilin_setup(q.x, q.y, fr.g0w, fr.g0h);
            }
        }
        // ---- pass 2, four samples per round
#pragma unroll 1
        for (int rnd = 0; rnd < 2; ++rnd) {
        const int s8 = 4 * rnd + (lane32 >> 3);                 // sample of this 8-lane group within the warp's eight
        const int src0 = 4 * s8;                                // pass-1 lane of (sample, view 0)
        const int i = i8 + s8;
        const bool real = i < n_chunk;
        const long long n = sample0 + (real ? i : n_chunk - 1);
        float p[3];
        const float* ray;
        if (pts_in) {
            p[0] = pts_in[3 * n]; p[1] = pts_in[3 * n + 1]; p[2] = pts_in[3 * n + 2];
            ray = view_in + 3 * n;
        } else {
            ray = rays + (size_t)(n / S) * VANERF_RAY_STRIDE;
            sample_point(ray, tar.cam_pos, z[n], p);
        }
        float pw[MAXV];
        bool m = true;
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            if (v < V) {
                pw[v] = __shfl_sync(0xffffffffu, mw, src0 + v);
                const int okv = __shfl_sync(0xffffffffu, mok, src0 + v);       // unconditionally: `m && shfl(...)` would let
                m = m && (okv != 0);                                            // lanes with m == false skip a warp-wide shuffle
            }
        }
        const float mf = m ? 1.0f : 0.0f;
        float pw_sum = 0.f;
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            if (v < V) {
                pw[v] = pw[v] * mf;
                pw_sum += pw[v];
            }
        }
        if (valid_out && lane == 0 && real) valid_out[n] = m ? 1 : 0;
        const int nn = nn_vert[n];
        const int tw = (nn + NUM_V_HAND) % (2 * NUM_V_HAND);
        const float sdfv = sdf[n];
        const int tile = i / TC_ROWS, row = i % TC_ROWS;
        const uint32_t off = tc::slot_chunk_off(row, lane);
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            if (v >= V) break;
            unsigned char* img = rec + ((size_t)tile * V + v) * (TC_REC_IMAGES * TC_SLOT);
            const float x = __shfl_sync(0xffffffffu, mx, src0 + v), y = __shfl_sync(0xffffffffu, my, src0 + v);
            const float cr = __shfl_sync(0xffffffffu, mcr, src0 + v), cg = __shfl_sync(0xffffffffu, mcg, src0 + v), cb = __shfl_sync(0xffffffffu, mcb, src0 + v);
            Bilin b64;
            b64.i00 = __shfl_sync(0xffffffffu, mb.i00, src0 + v); b64.i01 = __shfl_sync(0xffffffffu, mb.i01, src0 + v);
            b64.i10 = __shfl_sync(0xffffffffu, mb.i10, src0 + v); b64.i11 = __shfl_sync(0xffffffffu, mb.i11, src0 + v);
            b64.nw = __shfl_sync(0xffffffffu, mb.nw, src0 + v); b64.ne = __shfl_sync(0xffffffffu, mb.ne, src0 + v);
            b64.sw = __shfl_sync(0xffffffffu, mb.sw, src0 + v); b64.se = __shfl_sync(0xffffffffu, mb.se, src0 + v);
            const size_t vb = (size_t)v * fr.n_verts;
            const float qv = qvis[(size_t)v * N_total + n] ? 1.0f : 0.0f, vn = fr.vis[vb + nn], vt = fr.vis[vb + tw];
#ifdef GTC_DEBUG
            {
                const int lim = fr.g0h * fr.g0w;
                if (b64.i00 < 0 || b64.i00 >= lim || b64.i01 >= lim || b64.i10 >= lim || b64.i11 >= lim || nn < 0 || nn >= fr.n_verts ||
                    tw < 0 || tw >= fr.n_verts || n < 0 || n >= N_total || tile < 0) {
                    printf("GTC_DEBUG blk %d thr %d i %d n %lld v %d src0 %d idx %d %d %d %d lim %d nn %d tw %d\n", blockIdx.x, threadIdx.x, i, n, v, src0,
                           b64.i00, b64.i01, b64.i10, b64.i11, lim, nn, tw);
                    continue;
                }
            }
#endif
            // R0: pixel-aligned geo0, R1 / R2: nearest / twin vertex rows (already multiplied by visibility)
            *reinterpret_cast<uint4*>(img + off) = tap8_bf16(ft.geo0 + (size_t)v * fr.g0h * fr.g0w * 64, 64, 8 * lane, b64);
            *reinterpret_cast<uint4*>(img + TC_SLOT + off) = *reinterpret_cast<const uint4*>(ft.T64 + (vb + nn) * 64 + 8 * lane);
            *reinterpret_cast<uint4*>(img + 2 * TC_SLOT + off) = *reinterpret_cast<const uint4*>(ft.T64 + (vb + tw) * 64 + 8 * lane);
            // the two 8-channel bilinear taps (geo1 on lane 2, tex on lane 0) run as ONE code path with per-lane map
            // parameters: a warp executes every divergent branch body, so two separate branches would cost twice
            uint4 tap8c = make_uint4(0, 0, 0, 0);
            if (lane == 0 || lane == 2) {
                const bool g = lane == 2;
                const int mw = g ? fr.g1w : fr.tw, mh = g ? fr.g1h : fr.th;
                const __nv_bfloat16* mp = g ? ft.geo1 + (size_t)v * fr.g1h * fr.g1w * 8 : ft.tex + (size_t)v * fr.th * fr.tw * 8;
                const Bilin b = bilin_setup(x, y, mw, mh);
                tap8c = tap8_bf16(mp, 8, 0, b);
            }
            // R3 (misc): c0 [sdf,qvis,vn,vt,0..] | c1 0 | c2 px8 | c3 a8 | c4 b8 | c5 [sdf,qvis,vn,vt,0..] | c6,c7 0
            {
                uint4 o = make_uint4(0, 0, 0, 0);
                if (lane == 0 || lane == 5) {
                    o.x = tc::pack_bf16(sdfv, qv);
                    o.y = tc::pack_bf16(vn, vt);
                } else if (lane == 2) {
                    o = tap8c;
                } else if (lane == 3) {
                    o = *reinterpret_cast<const uint4*>(ft.T8 + (vb + nn) * 8);
                } else if (lane == 4) {
                    o = *reinterpret_cast<const uint4*>(ft.T8 + (vb + tw) * 8);
                }
                *reinterpret_cast<uint4*>(img + 3 * TC_SLOT + off) = o;
            }
            // R4 (tex): c0 qtex8 | c1 atex8 | c2 btex8 | c3,c4 agf0-15 | c5,c6 bgf0-15 | c7 [agf16,agf17,bgf16,bgf17,q r,g,b,a r]
            {
                uint4 o;
                const __nv_bfloat16* ta = ft.Ttex + (vb + nn) * 32;
                const __nv_bfloat16* tb = ft.Ttex + (vb + tw) * 32;
                if (lane == 0) o = tap8c;
                else if (lane < 7) {       // lanes 1..6: one row chunk each, a per-lane pointer instead of six branches
                    const __nv_bfloat16* rowp = (lane == 1 || lane == 3 || lane == 4) ? ta : tb;
                    const int coff = lane < 3 ? 0 : ((lane == 3 || lane == 5) ? 8 : 16);
                    o = *reinterpret_cast<const uint4*>(rowp + coff);
                } else {
                    const uint4 qa = *reinterpret_cast<const uint4*>(ta + 24);      // gf16, gf17, r, g, b, 0,0,0
                    const uint4 qb = *reinterpret_cast<const uint4*>(tb + 24);
                    o.x = qa.x;                                                      // agf16, agf17
                    o.y = qb.x;                                                      // bgf16, bgf17
                    o.z = tc::pack_bf16(cr, cg);                                     // q r, g (tapped in pass 1)
                    o.w = tc::pack_bf16(cb, __uint_as_float(qa.y << 16));            // q b, a r
                }
                *reinterpret_cast<uint4*>(img + 4 * TC_SLOT + off) = o;
            }
            // side record
            if (lane < 4) {
                uint4 o;
                if (lane < 2) {        // lanes 0 and 1 share the direction normalisations (one code path)
                    float s0 = p[0] - fr.src_pos[v][0], s1 = p[1] - fr.src_pos[v][1], s2 = p[2] - fr.src_pos[v][2];
                    const float inv = 1.0f / fmaxf(sqrtf(s0 * s0 + s1 * s1 + s2 * s2), 1e-12f);
                    s0 *= inv; s1 *= inv; s2 *= inv;
                    const float e0 = ray[0] - s0, e1 = ray[1] - s1, e2 = ray[2] - s2;
                    const float ninv = 1.0f / fmaxf(sqrtf(e0 * e0 + e1 * e1 + e2 * e2), 1e-6f);
                    if (lane == 0) {
                        const float* E = fr.extrin[v];
                        o = make_uint4(__float_as_uint(xaffine(E, 0, p[0], p[1], p[2])), __float_as_uint(xaffine(E, 1, p[0], p[1], p[2])),
                                       __float_as_uint(xaffine(E, 2, p[0], p[1], p[2])), __float_as_uint(e0 * ninv));
                    } else {
                        o = make_uint4(__float_as_uint(e1 * ninv), __float_as_uint(e2 * ninv),
                                       __float_as_uint(s0 * ray[0] + s1 * ray[1] + s2 * ray[2]), __float_as_uint(pw[v] / (pw_sum + 1e-6f)));
                    }
                } else if (lane == 2) {
                    o = make_uint4(__float_as_uint(mf), 0, 0, 0);
                } else {
                    const uint4 qa = *reinterpret_cast<const uint4*>(ft.Ttex + (vb + nn) * 32 + 24);
                    const uint4 qb = *reinterpret_cast<const uint4*>(ft.Ttex + (vb + tw) * 32 + 24);
                    // [a g, a b | b r, b g | b b, qvis | vn, vt]
                    o.x = (qa.y >> 16) | (qa.z << 16);
                    o.y = qb.y;
                    o.z = (qb.z & 0xffffu) | (tc::pack_bf16(0.f, qv) & 0xffff0000u);
                    o.w = tc::pack_bf16(vn, vt);
                }
                *reinterpret_cast<uint4*>(aux + ((size_t)i * V + v) * TC_AUX_BYTES + 16 * lane) = o;
            }
        }
        }   // round
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Split-precision ("fp32") path: the fp32 gather records of k_gather (one 308-float row per (sample, view), gather.cuh /
// common.cuh REC_*) -> the same five operand images as k_gather_tc, each as a bf16 hi image and a bf16 lo image
// (value = hi + lo to 16 mantissa bits), plus a 96-byte side record whose eight "extras" stay fp32.
// rec_img: (n_tiles, V, 10, 16 KB) = images R0..R4 hi, then R0..R4 lo; aux: (n_tiles * 128, V, 96 B).
// One thread per (row, view, image, 16-byte chunk); rows past n_chunk replicate the last sample (never stored).
// record offset of element j of chunk c of image img (-1 = zero); mirrors the operand layout of k_gather_tc
__device__ __forceinline__ int rec_split_offset(int img, int c, int j) {
    switch (img) {
        case 0: return REC_PX64 + 8 * c + j;
        case 1: return REC_A64 + 8 * c + j;
        case 2: return REC_B64 + 8 * c + j;
        case 3:
            if (c == 0 || c == 5) return j < 4 ? REC_SDF + j : -1;               // sdf, qvis, vn, vt
            if (c == 2) return REC_PX8 + j;
            if (c == 3) return REC_A8 + j;
            if (c == 4) return REC_B8 + j;
            return -1;
        default:
            if (c == 0) return REC_QTEX + j;
            if (c == 1) return REC_ATEX + 3 + j;
            if (c == 2) return REC_BTEX + 3 + j;
            if (c == 3 || c == 4) return REC_ATEX + 11 + 8 * (c - 3) + j;
            if (c == 5 || c == 6) return REC_BTEX + 11 + 8 * (c - 5) + j;
            switch (j) {                                                          // c == 7
                case 0: return REC_ATEX + 27;
                case 1: return REC_ATEX + 28;
                case 2: return REC_BTEX + 27;
                case 3: return REC_BTEX + 28;
                case 4: return REC_QIMG;
                case 5: return REC_QIMG + 1;
                case 6: return REC_QIMG + 2;
                default: return REC_ATEX;
            }
    }
}
__global__ void __launch_bounds__(256) k_rec_split(const float* __restrict__ rec, int V, int n_chunk, unsigned char* __restrict__ rec_img,
                                                   unsigned char* __restrict__ aux) {
    __shared__ short tab[TC_REC_IMAGES * 8 * 8];                                // the column map, built once per block: no divergent switches per element
    for (int k = threadIdx.x; k < TC_REC_IMAGES * 8 * 8; k += blockDim.x) tab[k] = (short)rec_split_offset(k >> 6, (k >> 3) & 7, k & 7);
    __syncthreads();
    const int n_rows = ((n_chunk + TC_ROWS - 1) / TC_ROWS) * TC_ROWS;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_row = TC_REC_IMAGES * 8 + 6;                                  // 40 image chunks + 6 float4 of the side record
    if (t >= (long long)n_rows * V * per_row) return;
    const int u = (int)(t % per_row);
    const long long iv = t / per_row;
    const int v = (int)(iv % V), i = (int)(iv / V);
    const float* r = rec + ((size_t)(i < n_chunk ? i : n_chunk - 1) * V + v) * REC_STRIDE;
    if (u < TC_REC_IMAGES * 8) {
        const int img = u >> 3, c = u & 7;
        float f[8], fh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int o = tab[8 * u + j]; f[j] = o >= 0 ? r[o] : 0.0f; }
        const uint4 hi = pack8(f);
        unpack8(hi, fh);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] -= fh[j];
        const int tile = i / TC_ROWS, row = i % TC_ROWS;
        unsigned char* base = rec_img + ((size_t)tile * V + v) * (2 * TC_REC_IMAGES * TC_SLOT) + tc::slot_chunk_off(row, c);
        *reinterpret_cast<uint4*>(base + (size_t)img * TC_SLOT) = hi;
        *reinterpret_cast<uint4*>(base + (size_t)(TC_REC_IMAGES + img) * TC_SLOT) = pack8(f);
    } else {
        const int q = u - TC_REC_IMAGES * 8;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q == 0) o = make_float4(r[REC_CAM], r[REC_CAM + 1], r[REC_CAM + 2], r[REC_RD]);
        else if (q == 1) o = make_float4(r[REC_RD + 1], r[REC_RD + 2], r[REC_RD + 3], r[REC_PW]);
        else if (q == 2) o = make_float4(r[REC_MASK], 0.f, 0.f, 0.f);
        else if (q == 3) o = make_float4(r[REC_ATEX + 1], r[REC_ATEX + 2], r[REC_BTEX], r[REC_BTEX + 1]);
        else if (q == 4) o = make_float4(r[REC_BTEX + 2], r[REC_QVIS], r[REC_VN], r[REC_VT]);
        *reinterpret_cast<float4*>(aux + ((size_t)i * V + v) * 96 + 16 * q) = o;
    }
}
