// Ray generation, frustum distances, box clip and coarse sample depths.
// Replaces src/model.py:1190-1238 and VANeRF.ray_bbox_intersection (src/model.py:1497-1570).
// Bit-exact against oracle/oracle_torch.py: make_rays / sample_z (exact-op wrappers only, same operand order).
#pragma once
#include "common.cuh"

// rays: (R, 8) = dir.xyz, near, far, hit, box_near, box_far
// ztab: (S) shared by all rays, or (R,S) when t_per_ray != 0 (training: stratified jitter, src/model.py:1226-1230)
__global__ void k_sample_rays(TargetDev tar, const int* __restrict__ pix_xy, int R, const float* __restrict__ ztab, int t_per_ray,
                              int S, float* __restrict__ rays, float* __restrict__ z) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float x = (float)pix_xy[2 * r], y = (float)pix_xy[2 * r + 1];
    const float* iK = tar.inv_K;
    float d[3], dn[3], df[3];
    const float xn = xmul(tar.znear, x), yn = xmul(tar.znear, y), on = xmul(tar.znear, 1.0f);
    const float xf = xmul(tar.zfar, x), yf = xmul(tar.zfar, y), of = xmul(tar.zfar, 1.0f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        d[j] = xdot3(x, y, 1.0f, iK[j], iK[3 + j], iK[6 + j]);
        dn[j] = xdot3(xn, yn, on, iK[j], iK[3 + j], iK[6 + j]);
        df[j] = xdot3(xf, yf, of, iK[j], iK[3 + j], iK[6 + j]);
    }
    const float znear_r = xnorm3(dn[0], dn[1], dn[2]);
    const float zfar_r = xnorm3(df[0], df[1], df[2]);
    float w[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) w[j] = xdot3(d[0], d[1], d[2], tar.R[j], tar.R[3 + j], tar.R[6 + j]);
    const float nrm = fmaxf(xnorm3(w[0], w[1], w[2]), 1e-12f);
    float dir[3], dd[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        dir[j] = xdiv(w[j], nrm);
        dd[j] = fabsf(dir[j]) < 1e-5f ? 1e-5f : dir[j];          // sign is lost on purpose (model.py:1523)
    }
    // six plane hits, order [min_x,min_y,min_z,max_x,max_y,max_z]
    const float eps = 1e-6f;
    int n_in = 0;
    float p_first[3] = {0, 0, 0}, p_second[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float bound = (k < 3) ? tar.bmin[k] : tar.bmax[k - 3];
        const float t = xdiv(xsub(bound, tar.cam_pos[k % 3]), dd[k % 3]);
        float p[3];
        bool in = true;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p[c] = xadd(xmul(t, dd[c]), tar.cam_pos[c]);
            in = in && (p[c] >= xsub(tar.bmin[c], eps)) && (p[c] <= xadd(tar.bmax[c], eps));
        }
        if (in) {
            if (n_in == 0) { p_first[0] = p[0]; p_first[1] = p[1]; p_first[2] = p[2]; }
            else if (n_in == 1) { p_second[0] = p[0]; p_second[1] = p[1]; p_second[2] = p[2]; }
            n_in++;
        }
    }
    const bool hit = (n_in == 2);
    float bnear = 1.0f, bfar = 1.0f;
    if (hit) {
        const float nr = xnorm3(dd[0], dd[1], dd[2]);
        const float d0 = xdiv(xnorm3(xsub(p_first[0], tar.cam_pos[0]), xsub(p_first[1], tar.cam_pos[1]),
                                     xsub(p_first[2], tar.cam_pos[2])), nr);
        const float d1 = xdiv(xnorm3(xsub(p_second[0], tar.cam_pos[0]), xsub(p_second[1], tar.cam_pos[1]),
                                     xsub(p_second[2], tar.cam_pos[2])), nr);
        bnear = fminf(d0, d1);
        bfar = fmaxf(d0, d1);
    }
    // m*a + (1-m)*b blends (model.py:1217-1220)
    const float m1 = (hit && bnear > znear_r) ? 1.0f : 0.0f;
    const float zn = xadd(xmul(m1, bnear), xmul(xsub(1.0f, m1), znear_r));
    const float m2 = (hit && bfar < zfar_r) ? 1.0f : 0.0f;
    const float zf = xadd(xmul(m2, bfar), xmul(xsub(1.0f, m2), zfar_r));
    float* o = rays + (size_t)r * VANERF_RAY_STRIDE;
    o[0] = dir[0]; o[1] = dir[1]; o[2] = dir[2]; o[3] = zn; o[4] = zf; o[5] = hit ? 1.0f : 0.0f;
    o[6] = bnear; o[7] = bfar;
    const float span = xsub(zf, zn);
    const float* tt = t_per_ray ? ztab + (size_t)r * S : ztab;
    for (int s = 0; s < S; ++s) z[(size_t)r * S + s] = xadd(zn, xmul(span, tt[s]));
}

// eval_pts = cam_pos + dir * z  (mul, then add; model.py:1234)
__device__ __forceinline__ void sample_point(const float* ray, const float* cam_pos, float zz, float* p) {
    p[0] = xadd(cam_pos[0], xmul(ray[0], zz));
    p[1] = xadd(cam_pos[1], xmul(ray[1], zz));
    p[2] = xadd(cam_pos[2], xmul(ray[2], zz));
}
