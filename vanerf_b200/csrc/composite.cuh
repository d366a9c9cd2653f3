// Warp-per-ray alpha compositing and hierarchical (importance) resampling (north-star kernel 3).
//   VANeRF.rgba2out + sdf_activation   src/model.py:1465-1494, :879-882
//   VANeRF.importance_sample + sort    src/model.py:1425-1462, :1301-1307
// Transmittance is an exclusive prefix product over the samples of a ray: each lane multiplies its own run of
// consecutive samples, then a 5-step warp-shuffle scan combines the lanes.  The pdf normaliser and the cdf of the
// resampler are summed left to right (the oracle's defined order), so fine depths are bit-exact given the same
// contrib; the merge of coarse and fine depths is a rank sort (values identical to torch.sort).
#pragma once
#include "common.cuh"

#define COMP_WARPS 4
#define COMP_MAX_PER_LANE 8      // S <= 256

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(COMP_WARPS * 32)
k_composite(const float* __restrict__ rgba, const float* __restrict__ z, const float* __restrict__ mesh_sdf, int R, int S,
            float beta, float* __restrict__ color, float* __restrict__ depth, float* __restrict__ alpha,
            float* __restrict__ sdf_out, float* __restrict__ contrib_out) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * COMP_WARPS;
    const int per = (S + 31) / 32;
    // every warp runs the same number of iterations so that the shuffles stay convergent
    for (int r0 = 0; r0 < R; r0 += n_warps) {
        const int r = r0 + warp0;
        const bool live = r < R;
        const int rr = live ? r : R - 1;
        const float* zr = z + (size_t)rr * S;
        const float* ar = rgba + (size_t)rr * S * 5;
        const float* sr = mesh_sdf + (size_t)rr * S;
        float c[COMP_MAX_PER_LANE], local = 1.0f;
        const int s0 = lane * per;
#pragma unroll
        for (int i = 0; i < COMP_MAX_PER_LANE; ++i) {
            const int s = s0 + i;
            c[i] = 0.0f;
            if (i < per && s < S) {
                const float a = ar[(size_t)s * 5] + sr[s];
                const float sigma = (1.0f / (1.0f + expf(a / beta))) / beta;        // sigmoid(-a/beta)/beta
                const float dist = (s + 1 < S) ? (zr[s + 1] - zr[s]) : 1e10f;
                c[i] = 1.0f - expf(-sigma * dist);
                local *= (1.0f - c[i]);
            }
        }
        // exclusive prefix product across lanes
        float incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl *= t;
        }
        float T = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) T = 1.0f;
        float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_a = 0.f, acc_z = 0.f, acc_s = 0.f;
#pragma unroll
        for (int i = 0; i < COMP_MAX_PER_LANE; ++i) {
            const int s = s0 + i;
            if (i < per && s < S) {
                const float w = c[i] * T;
                T *= (1.0f - c[i]);
                const float* px = ar + (size_t)s * 5;
                acc_r += px[2] * w; acc_g += px[3] * w; acc_b += px[4] * w;
                acc_a += w; acc_z += zr[s] * w; acc_s += px[1] * w;
                if (contrib_out && live) contrib_out[(size_t)r * S + s] = w;
            }
        }
        acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b);
        acc_a = warp_sum(acc_a); acc_z = warp_sum(acc_z); acc_s = warp_sum(acc_s);
        if (lane == 0 && live) {
            if (color) { color[3 * (size_t)r] = acc_r; color[3 * (size_t)r + 1] = acc_g; color[3 * (size_t)r + 2] = acc_b; }
            if (alpha) alpha[r] = acc_a;
            if (depth) depth[r] = acc_z / (acc_a + 1e-8f);
            if (sdf_out) sdf_out[r] = acc_s / (acc_a + 1e-8f);
        }
    }
}

// One warp per ray.  Dynamic shared memory per warp: cdf[S-1] | zmid[S-1] | vals[S+nf]  (floats).
__global__ void __launch_bounds__(COMP_WARPS * 32)
k_importance(const float* __restrict__ contrib, const float* __restrict__ z, const float* __restrict__ zmid_in, int R, int S,
             const float* __restrict__ u, int nf, int u_per_ray, float* __restrict__ z_fine_only, float* __restrict__ z_out,
             unsigned char* __restrict__ src_map) {      // optional (R, S + nf): merged slot -> index into [z | z_fine]
    DYN_SMEM(float, sm);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int per_warp = 2 * (S - 1) + (S + nf);
    float* cdf = sm + (size_t)wib * per_warp;
    float* zmid = cdf + (S - 1);
    float* vals = zmid + (S - 1);
    const int warp0 = blockIdx.x * COMP_WARPS + wib;
    const int n_warps = gridDim.x * COMP_WARPS;
    const int nb = S - 2;                 // pdf bins = contrib[1:-1]
    for (int r0 = 0; r0 < R; r0 += n_warps) {
        const int r = r0 + warp0;
        const bool live = r < R;
        const int rr = live ? r : R - 1;
        // reference calling convention (zmid_in != NULL): contrib holds the S-2 inner bins, zmid_in the S-1 midpoints
        const float* cr = zmid_in ? contrib + (size_t)rr * (S - 2) - 1 : contrib + (size_t)rr * S;
        const float* zr = zmid_in ? nullptr : z + (size_t)rr * S;
        __syncwarp();
        if (zmid_in) {
            for (int i = lane; i < S - 1; i += 32) zmid[i] = zmid_in[(size_t)rr * (S - 1) + i];
        } else {
            for (int i = lane; i < S - 1; i += 32) zmid[i] = xmul(0.5f, xadd(zr[i + 1], zr[i]));
            for (int i = lane; i < S; i += 32) vals[i] = zr[i];
        }
        // left-to-right sums (every lane computes the same values; 2*(S-2) dependent adds)
        float tot = 0.0f;
        for (int i = 0; i < nb; ++i) tot = xadd(tot, xadd(cr[1 + i], 1e-5f));
        if (lane == 0) {
            float run = 0.0f;
            cdf[0] = 0.0f;
            for (int i = 0; i < nb; ++i) {
                run = xadd(run, xdiv(xadd(cr[1 + i], 1e-5f), tot));
                cdf[i + 1] = run;
            }
        }
        __syncwarp();
        for (int j = lane; j < nf; j += 32) {
            const float uj = u_per_ray ? u[(size_t)rr * nf + j] : u[j];
            // searchsorted(right=True): number of cdf entries <= u   (cdf is non-decreasing, S-1 entries)
            int lo_b = 0, hi_b = S - 1;
            while (lo_b < hi_b) {
                const int mid = (lo_b + hi_b) >> 1;
                if (cdf[mid] <= uj) lo_b = mid + 1; else hi_b = mid;
            }
            const int idx = lo_b;
            const int lo = max(idx - 1, 0), hi = min(idx, S - 2);
            const float cl = cdf[lo], ch = cdf[hi];
            float den = xsub(ch, cl);
            if (den < 1e-5f) den = 1.0f;
            const float zf = xadd(zmid[lo], xmul(xdiv(xsub(uj, cl), den), xsub(zmid[hi], zmid[lo])));
            vals[S + j] = zf;
            if (z_fine_only && live) z_fine_only[(size_t)r * nf + j] = zf;
        }
        __syncwarp();
        if (!z_out) continue;
        const int n = S + nf;
        for (int i = lane; i < n; i += 32) {
            const float x = vals[i];
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const float y = vals[j];
                rank += (y < x || (y == x && j < i)) ? 1 : 0;
            }
            if (live) {
                z_out[(size_t)r * n + rank] = x;
                if (src_map) src_map[(size_t)r * n + rank] = (unsigned char)i;
            }
        }
    }
}

// Merged fine-pass inputs of rgba2out from the two places they were evaluated (coarse reuse, see vanerf_render_rays):
// slot k of ray r comes from coarse sample src < S (rgba_c, sdf_c) or from new fine sample src - S (rgba_f, sdf_f).
// Geometry of the merged fine set assembled from the coarse pass and the new depths (vanerf_render_rays): merged slot k of
// ray r takes sdf / nearest vertex / per-view sample visibility from index src_map[k] of [coarse | new].  The merged set
// contains the coarse depths bit for bit (sort of cat[z, z_fine], src/model.py:1301-1307) and the mesh queries depend on
// the sample position only, so this is the array cal_vis_sdf_batch / knn_points would return for the 128 depths.
__global__ void k_merge_geom(const unsigned char* __restrict__ src_map, const float* __restrict__ sdf_c, const int* __restrict__ nn_c,
                             const unsigned char* __restrict__ qv_c, const float* __restrict__ sdf_f, const int* __restrict__ nn_f,
                             const unsigned char* __restrict__ qv_f, int R, int S, int nf, int V, float* __restrict__ sdf_m,
                             int* __restrict__ nn_m, unsigned char* __restrict__ qv_m) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = S + nf;
    if (k >= (long long)R * n) return;
    const int r = (int)(k / n), src = src_map[k];
    const bool c = src < S;
    const size_t j = c ? (size_t)r * S + src : (size_t)r * nf + (src - S);
    const size_t Nc = (size_t)R * S, Nf = (size_t)R * nf, Nm = (size_t)R * n;
    sdf_m[k] = c ? sdf_c[j] : sdf_f[j];
    nn_m[k] = c ? nn_c[j] : nn_f[j];
    for (int v = 0; v < V; ++v) qv_m[(size_t)v * Nm + k] = c ? qv_c[(size_t)v * Nc + j] : qv_f[(size_t)v * Nf + j];
}

__global__ void k_merge_reuse(const unsigned char* __restrict__ src_map, const float* __restrict__ rgba_c, const float* __restrict__ sdf_c,
                              const float* __restrict__ rgba_f, const float* __restrict__ sdf_f, int R, int S, int nf,
                              float* __restrict__ rgba_m, float* __restrict__ sdf_m) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = S + nf;
    if (k >= (long long)R * n) return;
    const int r = (int)(k / n), src = src_map[k];
    const bool c = src < S;
    const size_t j = c ? (size_t)r * S + src : (size_t)r * nf + (src - S);
    const float* q = (c ? rgba_c : rgba_f) + 5 * j;
    float* o = rgba_m + 5 * (size_t)k;
    o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3]; o[4] = q[4];
    sdf_m[k] = (c ? sdf_c : sdf_f)[j];
}

// planes: color (stride,3) | depth (stride) | alpha (stride) | sdf (stride)  ->  rows (n,8) = r,g,b,depth,alpha,sdf,0,0
__global__ void k_pack_out(const float* __restrict__ pl, int stride, int n, float* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float4* o = reinterpret_cast<float4*>(out + (size_t)r * 8);
    o[0] = make_float4(pl[3 * (size_t)r], pl[3 * (size_t)r + 1], pl[3 * (size_t)r + 2], pl[3 * (size_t)stride + r]);
    o[1] = make_float4(pl[4 * (size_t)stride + r], pl[5 * (size_t)stride + r], 0.f, 0.f);
}
