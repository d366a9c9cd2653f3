// libvanerf_b200.so: context, weight packing, per-frame setup and kernel launchers behind the C ABI of
// include/vanerf_b200.h.  Single translation unit: the kernels live in the .cuh files included below.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "bvh_build.cuh"
#include "rays.cuh"
#include "geom.cuh"
#include "frame.cuh"
#include "gather.cuh"
#include "mlp_simt.cuh"
#include "composite.cuh"
#include "stages.cuh"
#include "train.cuh"
#ifndef VANERF_HOST_EMUL
#include "mlp_tc.cuh"
#include "gfeat.cuh"
#endif

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

enum KernelClass { KCL_SETUP = 0, KCL_RAYS, KCL_GEOM, KCL_GATHER, KCL_MLP, KCL_COMPOSITE, KCL_IMPORTANCE, KCL_COUNT };

#ifndef VANERF_HOST_EMUL
struct TimedEv { int kc; cudaEvent_t a, b; };
#endif

struct vanerf_ctx {
    int device = 0, sm_count = 148;
    int64_t launches = 0;
    char err[512] = {0};
    // optional per-kernel-class CUDA-event timing (bench.py roofline numbers)
    bool timing = false;
    double t_ms[KCL_COUNT] = {0};
    int64_t t_cnt[KCL_COUNT] = {0};
#ifndef VANERF_HOST_EMUL
    std::vector<TimedEv> evs;
#endif
    // weights
    DevBuf wblob, netdev;
#ifndef VANERF_HOST_EMUL
    // tensor-core path: step tables + bf16 weight images, bf16 maps / vertex tables, operand images, error flag
    TcTables h_tc;                     // host copy of the biases / small fp32 layers (weights) + camera-space keypoints (frame)
    TcProg h_prog;                     // static MMA program (layer shapes only); uploaded to __constant__ c_prog
    bool tc_tab_dirty = true;
    int tc_waves = 16;
    int tc_grid = 1 << 30;            // developer override (VANERF_TC_GRID): cap on the CTAs of k_mlp_tc<false> (scaling experiments)
    int split_waves = 8;               // tiles per CTA and launch of the split-precision path
    unsigned tcw_lo_off = 0;           // byte offset of the lo weight images inside tcw
    bool fp32_simt = false;            // VANERF_FP32_SIMT=1: fp32 path on the FFMA kernel (k_mlp_simt) instead of the split-precision tensor-core kernel
    DevBuf tcw, tctab, geo0b, geo1b, texb, T64b, T8b, Ttexb, tc_rec, tc_aux, gf_scratch;
    FrameTc ft;
    int* tc_err_host = nullptr;        // mapped pinned int written by the kernels (bounded waits that gave up)
    int* tc_err_dev = nullptr;
#endif
    NetDev h_net;
    bool have_weights = false;
    // frame
    FrameDev fr;
    bool have_frame = false;
    DevBuf geo0, geo1, tex, imgm, T64, T8, Ttex, vis, tri_nodes, tri_prims, vtx_nodes, vtx_prims,
        xyz_ndc, xy11, zbuf, tri_rec, vtx_rec, tri_node_lb, upload, bvh_scratch[2];
    // pinned staging of the per-frame host inputs (vertices, faces, camera-space keypoints): two buffers + "upload done" events
    void* stage[2] = {nullptr, nullptr};
    size_t stage_cap[2] = {0, 0};
    cudaEvent_t stage_ev[2];
    bool stage_ev_ok[2] = {false, false};
    int stage_i = 0;
    // scratch
    DevBuf rec, s_rays, s_z, s_z2, s_sdf, s_nn, s_qvis, s_rgba, s_contrib, s_valid, s_tab;
    DevBuf s_zf, s_srcmap, s_sdf_f, s_rgba_f, s_sdf_m, s_rgba_m;      // coarse reuse (vanerf_set_reuse_coarse)
    DevBuf s_nn_f, s_qvis_f, s_nn_m, s_qvis_m;                        // geometry reuse (vanerf_set_reuse_geometry)
    bool reuse_coarse = false;
    bool reuse_geometry = true;
    unsigned char* imp_src_map = nullptr;   // set around the importance launch of vanerf_render_rays (coarse reuse)
};

static int ctx_fail(vanerf_ctx* ctx, cudaError_t e, const char* what, int line) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "CUDA error %d (%s) at %s [vanerf_b200.cu:%d]", (int)e, cudaGetErrorString(e), what, line);
    return VANERF_ERR_CUDA;
}
static int ctx_state(vanerf_ctx* ctx) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "call order violated: load weights, then frame_setup, then render");
    return VANERF_ERR_STATE;
}
static int ctx_unsupported(vanerf_ctx* ctx, const char* msg) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "unsupported: %s", msg);
    return VANERF_ERR_UNSUPPORTED;
}
static int ctx_invalid(vanerf_ctx* ctx, const char* msg) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "invalid argument: %s", msg);
    return VANERF_ERR_INVALID;
}

static int ensure(vanerf_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CUDA_TRY(ctx, cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CUDA_TRY(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}
#define ENSURE(ctx, buf, bytes) do { int rc_ = ensure((ctx), (buf), (bytes)); if (rc_) return rc_; } while (0)
#define CHECK_LAUNCH(ctx) do { (ctx)->launches++; CUDA_TRY((ctx), cudaGetLastError()); } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Brackets the launches of one kernel class with CUDA events on the launching stream when timing is enabled.
struct TimedScope {
#ifndef VANERF_HOST_EMUL
    vanerf_ctx* c; cudaStream_t s; TimedEv e; bool on;
    TimedScope(vanerf_ctx* ctx, int kc, cudaStream_t st) : c(ctx), s(st), on(ctx->timing) {
        if (on) { e.kc = kc; cudaEventCreate(&e.a); cudaEventCreate(&e.b); cudaEventRecord(e.a, s); }
    }
    ~TimedScope() { if (on) { cudaEventRecord(e.b, s); c->evs.push_back(e); } }
#else
    TimedScope(vanerf_ctx*, int, cudaStream_t) {}
#endif
};

// Every entry point that touches the device runs with the context's device current and restores the caller's.
struct DeviceGuard {
    int prev = -1; bool sw = false;
    explicit DeviceGuard(const vanerf_ctx* c) {
        if (c && cudaGetDevice(&prev) == cudaSuccess && prev != c->device) sw = cudaSetDevice(c->device) == cudaSuccess;
    }
    ~DeviceGuard() { if (sw) cudaSetDevice(prev); }
};

extern "C" {

int vanerf_ctx_create(vanerf_ctx** out, int device) {
    if (!out) return VANERF_ERR_INVALID;
    vanerf_ctx* c = new vanerf_ctx();
    c->device = device;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};      // the caller's current device is kept
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return VANERF_ERR_CUDA; }
    int sm = 0;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { delete c; return VANERF_ERR_CUDA; }
    c->sm_count = sm;
    memset(&c->fr, 0, sizeof(c->fr));
    memset(&c->h_net, 0, sizeof(c->h_net));
    // opt-in shared-memory sizes are per device: set here, on this context's device (a second context on another GPU of
    // the same process sets them again for its own device)
    if (cudaFuncSetAttribute(k_mlp_simt, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(mlp_simt_smem_floats(MAXV) * sizeof(float)) + 1024) != cudaSuccess) { delete c; return VANERF_ERR_CUDA; }
#ifndef VANERF_HOST_EMUL
    if (cudaHostAlloc((void**)&c->tc_err_host, 8 * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&c->tc_err_dev, c->tc_err_host, 0) != cudaSuccess) { delete c; return VANERF_ERR_CUDA; }
    memset(c->tc_err_host, 0, 8 * sizeof(int));
    memset(&c->ft, 0, sizeof(c->ft));
    memset(&c->h_tc, 0, sizeof(c->h_tc));
    memset(&c->h_prog, 0, sizeof(c->h_prog));
    if (cudaFuncSetAttribute(k_mlp_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_mlp_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_tc_mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * TC_SLOT + 1024) != cudaSuccess) {
        cudaFreeHost(c->tc_err_host); delete c; return VANERF_ERR_CUDA;
    }
    if (const char* w = getenv("VANERF_TC_WAVES")) c->tc_waves = std::max(1, std::min(16, atoi(w)));      // developer override
    if (const char* w = getenv("VANERF_TC_GRID")) c->tc_grid = std::max(1, atoi(w));                         // developer override
    if (const char* w = getenv("VANERF_FP32_SIMT")) c->fp32_simt = atoi(w) != 0;                            // developer override
    if (const char* w = getenv("VANERF_REUSE_GEOM")) c->reuse_geometry = atoi(w) != 0;                      // developer override
#endif
    *out = c;
    return VANERF_OK;
}

void vanerf_ctx_destroy(vanerf_ctx* c) {
    if (!c) return;
    DeviceGuard dg_(c);
#ifndef VANERF_HOST_EMUL
    DevBuf* tcb[] = {&c->tcw, &c->tctab, &c->geo0b, &c->geo1b, &c->texb, &c->T64b, &c->T8b, &c->Ttexb, &c->tc_rec, &c->tc_aux, &c->gf_scratch};
    for (DevBuf* b : tcb) if (b->p) cudaFree(b->p);
    if (c->tc_err_host) cudaFreeHost(c->tc_err_host);
#endif
    for (int i = 0; i < 2; ++i) {
        if (c->stage_ev_ok[i]) { cudaEventSynchronize(c->stage_ev[i]); cudaEventDestroy(c->stage_ev[i]); }
        if (c->stage[i]) cudaFreeHost(c->stage[i]);
    }
    DevBuf* all[] = {&c->wblob, &c->netdev, &c->geo0, &c->geo1, &c->tex, &c->imgm, &c->T64, &c->T8, &c->Ttex, &c->vis,
                     &c->tri_nodes, &c->tri_prims, &c->vtx_nodes, &c->vtx_prims, &c->upload, &c->bvh_scratch[0], &c->bvh_scratch[1],
                     &c->xyz_ndc, &c->xy11, &c->zbuf, &c->tri_rec, &c->vtx_rec, &c->tri_node_lb, &c->rec, &c->s_rays, &c->s_z, &c->s_z2, &c->s_sdf, &c->s_nn,
                     &c->s_qvis, &c->s_rgba, &c->s_contrib, &c->s_valid, &c->s_tab,
                     &c->s_zf, &c->s_srcmap, &c->s_sdf_f, &c->s_rgba_f, &c->s_sdf_m, &c->s_rgba_m,
                     &c->s_nn_f, &c->s_qvis_f, &c->s_nn_m, &c->s_qvis_m};
    for (DevBuf* b : all) if (b->p) cudaFree(b->p);
    delete c;
}

const char* vanerf_status_str(int s) {
    switch (s) {
        case VANERF_OK: return "ok";
        case VANERF_ERR_INVALID: return "invalid argument";
        case VANERF_ERR_CUDA: return "CUDA runtime error";
        case VANERF_ERR_STATE: return "call order violated (load weights, then frame_setup, then render)";
        case VANERF_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}
const char* vanerf_last_error(const vanerf_ctx* c) { return c ? c->err : "no context"; }
int vanerf_sm_count(const vanerf_ctx* c) { return c ? c->sm_count : 0; }
int64_t vanerf_launch_count(const vanerf_ctx* c) { return c ? c->launches : 0; }

// ------------------------------------------------------------------------------------------------ weights
int vanerf_load_weights(vanerf_ctx* ctx, const vanerf_weights* w, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !w) return VANERF_ERR_INVALID;
    const vanerf_linear* src[L_COUNT] = {
        &w->geo_at[0], &w->geo_at[1], &w->geo_f[0], &w->geo_f[1], &w->geo8_at[0], &w->geo8_at[1], &w->geo8_f[0], &w->geo8_f[1],
        &w->mlp[0], &w->mlp[1], &w->mlp[2], &w->mlp[3], &w->post[0], &w->post[1], &w->post[2], &w->compress,
        &w->tex_at[0], &w->tex_at[1], &w->tex_f[0], &w->tex_f[1],
        &w->ray[0], &w->ray[1], &w->base[0], &w->base[1], &w->vis1[0], &w->vis1[1], &w->vis2[0], &w->vis2[1],
        &w->outl[0], &w->outl[1], &w->outl[2]};
    static const int expect[L_COUNT][2] = {   // (out, in) of configs/vanerf.json
        {10, 196}, {3, 10}, {64, 196}, {64, 64}, {10, 28}, {3, 10}, {8, 28}, {8, 8},
        {128, 358}, {128, 128}, {120, 136}, {64, 120}, {64, 128}, {64, 64}, {2, 64}, {24, 128},
        {96, 96}, {6, 96}, {96, 96}, {40, 96},
        {16, 4}, {40, 16}, {64, 120}, {32, 64}, {32, 32}, {33, 32}, {32, 32}, {1, 32}, {16, 37}, {8, 16}, {1, 8}};
    size_t total = 0;
    std::vector<size_t> off_w(L_COUNT), off_b(L_COUNT);
    for (int i = 0; i < L_COUNT; ++i) {
        if (!src[i]->w) return ctx_invalid(ctx, "NULL weight matrix");
        if (src[i]->out_dim != expect[i][0] || src[i]->in_dim != expect[i][1]) {
            snprintf(ctx->err, sizeof(ctx->err), "layer %d: expected (%d,%d) got (%d,%d)", i, expect[i][0], expect[i][1], src[i]->out_dim, src[i]->in_dim);
            return VANERF_ERR_UNSUPPORTED;
        }
        const int npad = (src[i]->out_dim + 7) & ~7;
        off_w[i] = total; total += (size_t)src[i]->in_dim * npad;
        off_b[i] = total; total += npad;
        total = (total + 3) & ~(size_t)3;
    }
    std::vector<float> blob(total, 0.0f);
    ENSURE(ctx, ctx->wblob, total * sizeof(float));
    ENSURE(ctx, ctx->netdev, sizeof(NetDev));
    for (int i = 0; i < L_COUNT; ++i) {
        const int N = src[i]->out_dim, K = src[i]->in_dim, npad = (N + 7) & ~7;
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) blob[off_w[i] + (size_t)k * npad + n] = src[i]->w[(size_t)n * K + k];
        if (src[i]->b) for (int n = 0; n < N; ++n) blob[off_b[i] + n] = src[i]->b[n];
        LayerDev& L = ctx->h_net.layer[i];
        L.wt = (const float*)ctx->wblob.p + off_w[i];
        L.b = (const float*)ctx->wblob.p + off_b[i];
        L.K = K; L.N = N; L.Npad = npad; L.pad_ = 0;
    }
    ctx->h_net.ani_al_abs = fabsf(w->ani_al);
    ctx->h_net.beta = fmaxf(w->sigmoid_beta, 2e-3f);           // sdf_activation clamp (src/model.py:880)
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->wblob.p, blob.data(), total * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->netdev.p, &ctx->h_net, sizeof(NetDev), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CUDA_TRY(ctx, cudaStreamSynchronize((cudaStream_t)stream));   // host staging buffers go out of scope
#ifndef VANERF_HOST_EMUL
    {   // tensor-core path: step tables + swizzled bf16 weight images
        std::vector<uint16_t> img, img_lo;
        float kpt_keep[TC_MAXV * NKPT * 4];
        memcpy(kpt_keep, ctx->h_tc.kpt4, sizeof(kpt_keep));
        tc_build(src, w->ani_al, ctx->h_tc, ctx->h_prog, img, &img_lo);
        memcpy(ctx->h_tc.kpt4, kpt_keep, sizeof(kpt_keep));
        ctx->tc_tab_dirty = true;
        // bf16 weight images [hi | lo]: the bf16 path streams the hi half, the split-precision path both
        ctx->tcw_lo_off = (unsigned)(img.size() * 2);
        ENSURE(ctx, ctx->tcw, img.size() * 4);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tcw.p, img.data(), img.size() * 2, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        CUDA_TRY(ctx, cudaMemcpyAsync((unsigned char*)ctx->tcw.p + img.size() * 2, img_lo.data(), img_lo.size() * 2, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        CUDA_TRY(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    }
#endif
    ctx->have_weights = true;
    return VANERF_OK;
}

// ------------------------------------------------------------------------------------------------ frame
int vanerf_frame_setup(vanerf_ctx* ctx, const vanerf_frame* f, float* vert_vis_out, void* stream_) {
    DeviceGuard dg_(ctx);
    if (!ctx || !f) return VANERF_ERR_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int V = f->n_views, H = f->height, W = f->width, Nv = f->n_verts, F = f->n_faces;
    if (V < 1 || V > MAXV) return ctx_invalid(ctx, "n_views out of range (1..4)");
    if (Nv != 2 * NUM_V_HAND) return ctx_invalid(ctx, "mesh must have 1558 vertices (two sealed MANO hands)");
    if (!f->KRT || !f->extrin || !f->src_cam_pos || !f->kpt3d || !f->verts || !f->faces || !f->img || !f->fg_mask ||
        !f->feat_geo0 || !f->feat_geo1 || !f->feat_tex || !f->vert_gfeat)
        return ctx_invalid(ctx, "NULL pointer in vanerf_frame");
    FrameDev& fr = ctx->fr;
    fr.V = V; fr.H = H; fr.W = W; fr.n_verts = Nv; fr.n_faces = F;
    fr.znear = f->znear; fr.zfar = f->zfar; fr.z_range = f->z_range;
    memcpy(fr.KRT, f->KRT, sizeof(float) * 16 * V);
    memcpy(fr.extrin, f->extrin, sizeof(float) * 16 * V);
    memcpy(fr.src_pos, f->src_cam_pos, sizeof(float) * 3 * V);
    fr.g0h = f->g0_h; fr.g0w = f->g0_w; fr.g1h = f->g1_h; fr.g1w = f->g1_w; fr.th = f->t_h; fr.tw = f->t_w;

    // keypoints in each source camera frame: kc = kpt3d @ R^T + t (src/spatial.py:84), exact-op order on the host
    std::vector<float> kc((size_t)V * NKPT * 3);
    for (int v = 0; v < V; ++v)
        for (int k = 0; k < NKPT; ++k)
            for (int j = 0; j < 3; ++j) {
                const float* M = f->extrin + 16 * v;
                const float* p = f->kpt3d + 3 * k;
                volatile float a = p[0] * M[4 * j], b = p[1] * M[4 * j + 1], c = p[2] * M[4 * j + 2];
                volatile float s = a + b;
                volatile float t = s + c;
                kc[((size_t)v * NKPT + k) * 3 + j] = t + M[4 * j + 3];
            }
    // ---- host inputs -> device through context-owned pinned staging (two buffers, so that this call never has to wait
    // for anything but the upload issued two frames ago); nothing below synchronises the stream
    const size_t o_verts = 0, o_faces = o_verts + (((size_t)Nv * 12 + 255) & ~(size_t)255), o_kc = o_faces + (((size_t)F * 12 + 255) & ~(size_t)255);
    const size_t up_bytes = o_kc + ((kc.size() * 4 + 255) & ~(size_t)255);
    {
        const int sb = ctx->stage_i ^= 1;
        if (ctx->stage_cap[sb] < up_bytes) {
            if (ctx->stage[sb]) { CUDA_TRY(ctx, cudaEventSynchronize(ctx->stage_ev[sb])); CUDA_TRY(ctx, cudaFreeHost(ctx->stage[sb])); }
            ctx->stage[sb] = nullptr; ctx->stage_cap[sb] = 0;
            CUDA_TRY(ctx, cudaHostAlloc(&ctx->stage[sb], up_bytes, cudaHostAllocDefault));
            ctx->stage_cap[sb] = up_bytes;
            if (!ctx->stage_ev_ok[sb]) { CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[sb], cudaEventDisableTiming)); ctx->stage_ev_ok[sb] = true; }
        } else {
            CUDA_TRY(ctx, cudaEventSynchronize(ctx->stage_ev[sb]));      // upload of two frames ago (long complete)
        }
        unsigned char* h = (unsigned char*)ctx->stage[sb];
        memcpy(h + o_verts, f->verts, (size_t)Nv * 12);
        memcpy(h + o_faces, f->faces, (size_t)F * 12);
        memcpy(h + o_kc, kc.data(), kc.size() * 4);
        ENSURE(ctx, ctx->upload, up_bytes);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->upload.p, h, up_bytes, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->stage_ev[sb], stream));
    }
    const float* d_verts = (const float*)((unsigned char*)ctx->upload.p + o_verts);
    const int* d_faces = (const int*)((unsigned char*)ctx->upload.p + o_faces);

    const size_t g0n = (size_t)V * 64 * fr.g0h * fr.g0w, g1n = (size_t)V * 8 * fr.g1h * fr.g1w, txn = (size_t)V * 8 * fr.th * fr.tw;
    ENSURE(ctx, ctx->geo0, g0n * 4); ENSURE(ctx, ctx->geo1, g1n * 4); ENSURE(ctx, ctx->tex, txn * 4);
    ENSURE(ctx, ctx->imgm, (size_t)V * H * W * 16);
    ENSURE(ctx, ctx->T64, (size_t)V * Nv * 64 * 4); ENSURE(ctx, ctx->T8, (size_t)V * Nv * 8 * 4); ENSURE(ctx, ctx->Ttex, (size_t)V * Nv * 32 * 4);
    ENSURE(ctx, ctx->vis, (size_t)V * Nv * 4);
    ENSURE(ctx, ctx->tri_nodes, (size_t)BVH_MAX_NODES * 32); ENSURE(ctx, ctx->tri_prims, (size_t)F * 4);
    ENSURE(ctx, ctx->vtx_nodes, (size_t)BVH_MAX_NODES * 32); ENSURE(ctx, ctx->vtx_prims, (size_t)Nv * 4);
    ENSURE(ctx, ctx->tri_rec, (size_t)F * TRI_REC_F4 * 16); ENSURE(ctx, ctx->vtx_rec, (size_t)Nv * 16);
    ENSURE(ctx, ctx->tri_node_lb, (size_t)BVH_MAX_NODES * 32);
    ENSURE(ctx, ctx->xyz_ndc, (size_t)V * Nv * 12); ENSURE(ctx, ctx->xy11, (size_t)V * Nv * 8);
    ENSURE(ctx, ctx->zbuf, (size_t)V * RASTER_S * RASTER_S * 8);
    if (F > BVH_MAX_PRIMS || Nv > BVH_MAX_PRIMS) return ctx_invalid(ctx, "mesh too large for the per-frame BVH builder (4096 primitives per tree)");

    fr.geo0 = (const float*)ctx->geo0.p; fr.geo1 = (const float*)ctx->geo1.p; fr.tex = (const float*)ctx->tex.p;
    fr.imgm = (const float*)ctx->imgm.p;
    fr.T64 = (const float*)ctx->T64.p; fr.T8 = (const float*)ctx->T8.p; fr.Ttex = (const float*)ctx->Ttex.p;
    fr.vis = (const float*)ctx->vis.p;
    fr.verts = d_verts; fr.faces = d_faces;
    fr.tri_nodes = (const float4*)ctx->tri_nodes.p; fr.tri_prims = (const int*)ctx->tri_prims.p;
    fr.tri_node_lb = (const float4*)ctx->tri_node_lb.p;
    fr.vtx_nodes = (const float4*)ctx->vtx_nodes.p; fr.vtx_prims = (const int*)ctx->vtx_prims.p;
    fr.kpt_cam = (const float*)((unsigned char*)ctx->upload.p + o_kc);
    fr.tri_rec = (const float4*)ctx->tri_rec.p; fr.vtx_rec = (const float4*)ctx->vtx_rec.p;

    const int T = 256;
    TimedScope ts(ctx, KCL_SETUP, stream);
    {   // acceleration structures, built on the device (bvh_build.cuh): triangle tree (leaves of 8) and vertex tree (leaves of 16)
        // in one two-CTA launch, then leaf records and per-node slab bounds
        const int np[2] = {F, Nv};
        BvhBuildArgs ba[2];
        for (int t = 0; t < 2; ++t) {
            const size_t n = (size_t)np[t];
            const size_t bytes = n * (6 + 3 + 1 + 3) * 4 + (size_t)BVH_MAX_NODES * (12 + 1 + 2) * 4 + (size_t)BVH_MAX_NODES * 8 + 256;
            ENSURE(ctx, ctx->bvh_scratch[t], bytes);
            unsigned char* q = (unsigned char*)ctx->bvh_scratch[t].p;
            BvhBuildArgs& a = ba[t];
            a.verts = d_verts; a.faces = t == 0 ? d_faces : nullptr; a.n = np[t];
            a.nodes = (float4*)(t == 0 ? ctx->tri_nodes.p : ctx->vtx_nodes.p);
            a.prims = (int*)(t == 0 ? ctx->tri_prims.p : ctx->vtx_prims.p);
            a.node_range = (int2*)q; q += (size_t)BVH_MAX_NODES * 8;
            a.box = (float*)q; q += n * 24; a.cen = (float*)q; q += n * 12; a.key = (float*)q; q += n * 4;
            a.prim_b = (int*)q; q += n * 4; a.node_a = (int*)q; q += n * 4; a.node_b = (int*)q; q += n * 4;
            a.nb = (int*)q; q += (size_t)BVH_MAX_NODES * 48; a.axis = (int*)q; q += (size_t)BVH_MAX_NODES * 4;
            a.active = (int*)q; q += (size_t)BVH_MAX_NODES * 8;
            a.n_nodes = (int*)q;
        }
        // leaf sizes from sweeps on B200 (2/4/8/16 triangles with the per-triangle lower bound x 4/8/16 vertices)
        ba[0].leaf = 8; ba[1].leaf = 16;
        VANERF_LAUNCH(k_bvh_build, 2, BVH_BUILD_THREADS, 0, stream, ba[0], ba[1]); CHECK_LAUNCH(ctx);
        VANERF_LAUNCH(k_tri_records, cdiv(F, 128), 128, 0, stream, d_verts, d_faces, (const int*)ctx->tri_prims.p, F, (float4*)ctx->tri_rec.p); CHECK_LAUNCH(ctx);
        VANERF_LAUNCH(k_vtx_records, cdiv(Nv, 128), 128, 0, stream, d_verts, (const int*)ctx->vtx_prims.p, Nv, (float4*)ctx->vtx_rec.p); CHECK_LAUNCH(ctx);
        VANERF_LAUNCH(k_tri_node_bounds, cdiv(BVH_MAX_NODES * 32, 128), 128, 0, stream, d_verts, d_faces, (const int*)ctx->tri_prims.p,
                      (const int2*)ba[0].node_range, (const int*)ba[0].n_nodes, (float4*)ctx->tri_node_lb.p); CHECK_LAUNCH(ctx);
    }
    VANERF_LAUNCH(k_repack_nhwc, cdiv(g0n, T), T, 0, stream, f->feat_geo0, (float*)ctx->geo0.p, V, 64, fr.g0h, fr.g0w); CHECK_LAUNCH(ctx);
    VANERF_LAUNCH(k_repack_nhwc, cdiv(g1n, T), T, 0, stream, f->feat_geo1, (float*)ctx->geo1.p, V, 8, fr.g1h, fr.g1w); CHECK_LAUNCH(ctx);
    VANERF_LAUNCH(k_repack_nhwc, cdiv(txn, T), T, 0, stream, f->feat_tex, (float*)ctx->tex.p, V, 8, fr.th, fr.tw); CHECK_LAUNCH(ctx);
    VANERF_LAUNCH(k_repack_imgm, cdiv((long long)V * H * W, T), T, 0, stream, f->img, f->fg_mask, (float4*)ctx->imgm.p, V, H, W); CHECK_LAUNCH(ctx);
    VANERF_LAUNCH(k_project_verts, cdiv(V * Nv, T), T, 0, stream, fr, (float*)ctx->xyz_ndc.p, (float*)ctx->xy11.p); CHECK_LAUNCH(ctx);
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->zbuf.p, 0xff, (size_t)V * RASTER_S * RASTER_S * 8, stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->vis.p, 0, (size_t)V * Nv * 4, stream));
    VANERF_LAUNCH(k_raster_faces, cdiv(V * F, 128), 128, 0, stream, fr, (const float*)ctx->xyz_ndc.p, (unsigned long long*)ctx->zbuf.p); CHECK_LAUNCH(ctx);
    VANERF_LAUNCH(k_resolve_vis, cdiv(V * RASTER_S * RASTER_S, T), T, 0, stream, fr, (const unsigned long long*)ctx->zbuf.p, (float*)ctx->vis.p); CHECK_LAUNCH(ctx);
    VANERF_LAUNCH(k_vertex_tables, cdiv((long long)V * Nv * 104, T), T, 0, stream, fr, (const float*)ctx->xy11.p, f->vert_gfeat,
                  (float*)ctx->T64.p, (float*)ctx->T8.p, (float*)ctx->Ttex.p); CHECK_LAUNCH(ctx);
#ifndef VANERF_HOST_EMUL
    {   // bf16 companions for the tensor-core path; camera-space keypoints go to the constant tables
        memset(ctx->h_tc.kpt4, 0, sizeof(ctx->h_tc.kpt4));
        for (size_t i = 0; i < std::min<size_t>(kc.size() / 3, (size_t)TC_MAXV * NKPT); ++i)
            for (int j = 0; j < 3; ++j) ctx->h_tc.kpt4[4 * i + j] = kc[3 * i + j];
        ctx->tc_tab_dirty = true;
        const size_t t64n = (size_t)V * Nv * 64, t8n = (size_t)V * Nv * 8, ttn = (size_t)V * Nv * 32;
        ENSURE(ctx, ctx->geo0b, g0n * 2); ENSURE(ctx, ctx->geo1b, g1n * 2); ENSURE(ctx, ctx->texb, txn * 2);
        ENSURE(ctx, ctx->T64b, t64n * 2); ENSURE(ctx, ctx->T8b, t8n * 2); ENSURE(ctx, ctx->Ttexb, ttn * 2);
        VANERF_LAUNCH(k_f32_to_bf16, cdiv(g0n, T), T, 0, stream, (const float*)ctx->geo0.p, (__nv_bfloat16*)ctx->geo0b.p, (long long)g0n); CHECK_LAUNCH(ctx);
        VANERF_LAUNCH(k_f32_to_bf16, cdiv(g1n, T), T, 0, stream, (const float*)ctx->geo1.p, (__nv_bfloat16*)ctx->geo1b.p, (long long)g1n); CHECK_LAUNCH(ctx);
        VANERF_LAUNCH(k_f32_to_bf16, cdiv(txn, T), T, 0, stream, (const float*)ctx->tex.p, (__nv_bfloat16*)ctx->texb.p, (long long)txn); CHECK_LAUNCH(ctx);
        VANERF_LAUNCH(k_f32_to_bf16, cdiv(t64n, T), T, 0, stream, (const float*)ctx->T64.p, (__nv_bfloat16*)ctx->T64b.p, (long long)t64n); CHECK_LAUNCH(ctx);
        VANERF_LAUNCH(k_f32_to_bf16, cdiv(t8n, T), T, 0, stream, (const float*)ctx->T8.p, (__nv_bfloat16*)ctx->T8b.p, (long long)t8n); CHECK_LAUNCH(ctx);
        VANERF_LAUNCH(k_ttex_bf16, cdiv(ttn, T), T, 0, stream, (const float*)ctx->Ttex.p, (__nv_bfloat16*)ctx->Ttexb.p, V * Nv); CHECK_LAUNCH(ctx);
        ctx->ft.geo0 = (const __nv_bfloat16*)ctx->geo0b.p; ctx->ft.geo1 = (const __nv_bfloat16*)ctx->geo1b.p;
        ctx->ft.tex = (const __nv_bfloat16*)ctx->texb.p; ctx->ft.T64 = (const __nv_bfloat16*)ctx->T64b.p;
        ctx->ft.T8 = (const __nv_bfloat16*)ctx->T8b.p; ctx->ft.Ttex = (const __nv_bfloat16*)ctx->Ttexb.p;
    }
#endif
    if (vert_vis_out)
        CUDA_TRY(ctx, cudaMemcpyAsync(vert_vis_out, ctx->vis.p, (size_t)V * Nv * 4, cudaMemcpyDeviceToDevice, stream));
    ctx->have_frame = true;
    return VANERF_OK;
}

// ------------------------------------------------------------------------------------------------ per-ray stages
static TargetDev make_target(const vanerf_target* t) {
    TargetDev d;
    memcpy(d.inv_K, t->inv_K, sizeof(d.inv_K));
    memcpy(d.R, t->R, sizeof(d.R));
    memcpy(d.cam_pos, t->cam_pos, sizeof(d.cam_pos));
    d.znear = t->znear; d.zfar = t->zfar;
    for (int c = 0; c < 3; ++c) {          // bounds + boffset (-0.01, +0.01), fp32 adds (src/model.py:1518)
        volatile float lo = t->bounds[c] + (-0.01f), hi = t->bounds[3 + c] + 0.01f;
        d.bmin[c] = lo; d.bmax[c] = hi;
    }
    return d;
}

int vanerf_sample_rays(vanerf_ctx* ctx, const vanerf_target* tar, const int32_t* pix_xy, int32_t R, const float* ztab,
                       int32_t S, float* rays, float* z, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !tar || !pix_xy || !ztab || !rays || !z || R <= 0 || S <= 0) return ctx_invalid(ctx, "vanerf_sample_rays");
    TimedScope ts(ctx, KCL_RAYS, (cudaStream_t)stream);
    VANERF_LAUNCH(k_sample_rays, cdiv(R, 128), 128, 0, stream, make_target(tar), pix_xy, R, ztab, 0, S, rays, z);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}

int vanerf_sample_rays_t(vanerf_ctx* ctx, const vanerf_target* tar, const int32_t* pix_xy, int32_t R, const float* ttab,
                         int32_t S, int32_t t_per_ray, float* rays, float* z, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !tar || !pix_xy || !ttab || !rays || !z || R <= 0 || S <= 0) return ctx_invalid(ctx, "vanerf_sample_rays_t");
    TimedScope ts(ctx, KCL_RAYS, (cudaStream_t)stream);
    VANERF_LAUNCH(k_sample_rays, cdiv(R, 128), 128, 0, stream, make_target(tar), pix_xy, R, ttab, t_per_ray ? 1 : 0, S, rays, z);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}

int vanerf_geom_query(vanerf_ctx* ctx, const vanerf_target* tar, const float* rays, const float* z, int32_t R, int32_t S,
                      float* pts, float* sdf, int32_t* face, int32_t* nn_vert, uint8_t* qvis, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !tar || !rays || !z || R <= 0 || S <= 0) return ctx_invalid(ctx, "vanerf_geom_query");
    if (!ctx->have_frame) return ctx_state(ctx);
    TimedScope ts(ctx, KCL_GEOM, (cudaStream_t)stream);
    // ray batches: one warp per (block of 32 rays, depth index), see GEOM_RAY_LANES in geom.cuh
#if GEOM_RAY_LANES
    const long long n_thr = (long long)cdiv(R, GEOM_RAY_LANES) * cdiv(S, 32 / GEOM_RAY_LANES) * 32;
#else
    const long long n_thr = (long long)R * S;
#endif
    VANERF_LAUNCH(k_geom_query, cdiv(n_thr, 128), 128, 0, stream, ctx->fr, make_target(tar), rays, z, (const float*)nullptr, R, S, pts,
                  sdf, face, nn_vert, qvis);
    CHECK_LAUNCH(ctx);
#ifdef GEOM_COUNT
    VANERF_LAUNCH(k_geom_print, 1, 1, 0, stream);
#endif
    return VANERF_OK;
}

#ifndef VANERF_HOST_EMUL
// Reports (and clears) the abort record of the tensor-core kernels: a bounded mbarrier wait that gave up made its launch
// drain with garbage results instead of hanging the GPU.  The record is written through mapped host memory.
static bool tc_take_error(vanerf_ctx* ctx) {
    volatile int* e = ctx->tc_err_host;
    if (!e[0]) return false;
    snprintf(ctx->err, sizeof(ctx->err), "tensor-core kernel: a bounded wait gave up (code %d; pending issue %d acc %d producer %d rec %d pe %d); "
             "the results of that call are invalid", e[0], e[2], e[3], e[4], e[5], e[6]);
    for (int i = 0; i < 8; ++i) e[i] = 0;
    return true;
}
// Tensor-core shading: gather (bf16 operand images) + fused MLP per chunk of 2 x SM-count tiles of 128 samples, so that
// one chunk's images (<= 75 MB at V = 3) stay L2 resident between the two kernels.
static int tc_shade(vanerf_ctx* ctx, const TargetDev& td, const float* rays, const float* z, int S, long long N, const float* sdf,
                    const int* nn, const unsigned char* qvis, float* rgba, unsigned char* valid, float* raw_out,
                    float* dbg_latent, cudaStream_t stream, const float* pts_in, const float* view_in) {
    const int V = ctx->fr.V;
    if (V > TC_MAXV) {
        snprintf(ctx->err, sizeof(ctx->err), "bf16 tensor-core path supports up to %d source views (got %d)", TC_MAXV, V);
        return VANERF_ERR_UNSUPPORTED;
    }
    if (tc_take_error(ctx)) return VANERF_ERR_CUDA;        // an earlier launch gave up: reported once, then the path is usable again
    // tiles per launch: `tc_waves` tile pairs per CTA (default 16: ~1.2 GB of operand images per launch at V = 3; one pair per
    // CTA keeps the images L2 resident but costs sixteen times the launches, and the gather runs 22 % faster on the larger grid)
    const int max_tiles = TC_TILES * ctx->sm_count * ctx->tc_waves;
    const long long chunk = (long long)max_tiles * TC_ROWS;
    ENSURE(ctx, ctx->tc_rec, (size_t)max_tiles * V * TC_REC_IMAGES * TC_SLOT);
    ENSURE(ctx, ctx->tc_aux, (size_t)max_tiles * TC_ROWS * V * TC_AUX_BYTES);
    if (!tc_program_matches(ctx->h_prog)) {           // the kernels execute the compile-time copy (kProg)
        snprintf(ctx->err, sizeof(ctx->err), "tensor-core path: packing script and compiled MMA program disagree");
        return VANERF_ERR_STATE;
    }
#if !TC_TAB_PARAM
    if (ctx->tc_tab_dirty) {            // stream-ordered: earlier launches on this stream have read the old tables
        ENSURE(ctx, ctx->tctab, sizeof(TcTables));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tctab.p, &ctx->h_tc, sizeof(TcTables), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(stream));          // h_tc may change again before an async copy would read it
        ctx->tc_tab_dirty = false;
    }
#endif
    for (long long s0 = 0; s0 < N; s0 += chunk) {
        const int nc = (int)((N - s0) < chunk ? (N - s0) : chunk);
        const int n_tiles = cdiv(nc, TC_ROWS);
        {
            TimedScope ts(ctx, KCL_GATHER, stream);
            const int gblocks = min(cdiv((long long)n_tiles * TC_ROWS, GTC_THREADS / 4), ctx->sm_count * 8);      // 8 samples per warp and iteration
            VANERF_LAUNCH(k_gather_tc, gblocks, GTC_THREADS, 0, stream, ctx->fr, ctx->ft, td, rays, z, pts_in, view_in, S, s0, nc, N,
                          sdf, nn, qvis, (unsigned char*)ctx->tc_rec.p, (unsigned char*)ctx->tc_aux.p, valid);
            CHECK_LAUNCH(ctx);
        }
        TimedScope ts(ctx, KCL_MLP, stream);
        TcArgs a;
        a.tab = (const TcTables*)ctx->tctab.p; a.wblob = (const unsigned char*)ctx->tcw.p;
        a.rec = (const unsigned char*)ctx->tc_rec.p; a.aux = (const unsigned char*)ctx->tc_aux.p;
        a.V = V; a.n_chunk = nc; a.sample0 = s0; a.wblob_lo_off = ctx->tcw_lo_off;
        a.rgba = rgba; a.raw_out = raw_out; a.dbg_latent = dbg_latent; a.err = ctx->tc_err_dev;
#if TC_TAB_PARAM
        // the tables (biases, small fp32 layers, camera-space keypoints of the frame) travel by value in the kernel
        // parameter: constant-bank reads with warp-uniform addresses, no shared memory, no context-global state
        VANERF_LAUNCH(k_mlp_tc<false>, min(min(cdiv(n_tiles, TC_TILES), ctx->sm_count), ctx->tc_grid), TC_THREADS, TC_SMEM_BYTES, stream, a, ctx->h_tc);
#else
        VANERF_LAUNCH(k_mlp_tc<false>, min(cdiv(n_tiles, TC_TILES), ctx->sm_count), TC_THREADS, TC_SMEM_BYTES, stream, a);
#endif
        CHECK_LAUNCH(ctx);
    }
    return VANERF_OK;
}
#endif

#if !defined(VANERF_HOST_EMUL) && TC_TAB_PARAM
// Split-precision ("fp32") tensor-core shading: fp32 gather records (k_gather) -> bf16 hi / lo operand images (k_rec_split) ->
// k_mlp_tc<true> (three MMAs per K step, one tile per CTA), in chunks of sm_count x split_waves tiles.
static int tc_shade_split(vanerf_ctx* ctx, const TargetDev& td, const float* rays, const float* z, int S, long long N, const float* sdf,
                          const int* nn, const unsigned char* qvis, float* rgba, unsigned char* valid, float* raw_out, float* dbg_latent,
                          cudaStream_t stream, const float* pts_in, const float* view_in) {
    const int V = ctx->fr.V;
    if (tc_take_error(ctx)) return VANERF_ERR_CUDA;
    if (!tc_program_matches(ctx->h_prog)) {
        snprintf(ctx->err, sizeof(ctx->err), "tensor-core path: packing script and compiled MMA program disagree");
        return VANERF_ERR_STATE;
    }
    const int max_tiles = ctx->sm_count * ctx->split_waves;
    const long long chunk = (long long)max_tiles * TC_ROWS;
    ENSURE(ctx, ctx->rec, (size_t)chunk * V * REC_STRIDE * 4);
    ENSURE(ctx, ctx->tc_rec, (size_t)max_tiles * V * 2 * TC_REC_IMAGES * TC_SLOT);
    ENSURE(ctx, ctx->tc_aux, (size_t)max_tiles * TC_ROWS * V * TC_AUX_BYTES_SPLIT);
    for (long long s0 = 0; s0 < N; s0 += chunk) {
        const int nc = (int)((N - s0) < chunk ? (N - s0) : chunk);
        const int n_tiles = cdiv(nc, TC_ROWS);
        {
            TimedScope ts(ctx, KCL_GATHER, stream);
            const int gblocks = min(cdiv(nc, GATHER_THREADS / 16), ctx->sm_count * 8);
            VANERF_LAUNCH(k_gather, gblocks, GATHER_THREADS, 0, stream, ctx->fr, td, rays, z, pts_in, view_in, S, s0, nc, N, sdf, nn, qvis,
                          (float*)ctx->rec.p, valid);
            CHECK_LAUNCH(ctx);
            const long long n_thr = (long long)n_tiles * TC_ROWS * V * (TC_REC_IMAGES * 8 + 6);
            VANERF_LAUNCH(k_rec_split, cdiv(n_thr, 256), 256, 0, stream, (const float*)ctx->rec.p, V, nc, (unsigned char*)ctx->tc_rec.p,
                          (unsigned char*)ctx->tc_aux.p);
            CHECK_LAUNCH(ctx);
        }
        TimedScope ts(ctx, KCL_MLP, stream);
        TcArgs a;
        a.tab = nullptr; a.wblob = (const unsigned char*)ctx->tcw.p; a.wblob_lo_off = ctx->tcw_lo_off;
        a.rec = (const unsigned char*)ctx->tc_rec.p; a.aux = (const unsigned char*)ctx->tc_aux.p;
        a.V = V; a.n_chunk = nc; a.sample0 = s0;
        a.rgba = rgba; a.raw_out = raw_out; a.dbg_latent = dbg_latent; a.err = ctx->tc_err_dev;
        VANERF_LAUNCH(k_mlp_tc<true>, min(n_tiles, ctx->sm_count), TC_THREADS, TC_SMEM_BYTES, stream, a, ctx->h_tc);
        CHECK_LAUNCH(ctx);
    }
    return VANERF_OK;
}
#endif

#define SHADE_CHUNK 65536      // samples per gather/MLP round; records: chunk * V * 1232 B

// pts_in/view_in != NULL: explicit points and view directions ((N,3) each, R*S == N) instead of rays + depths
static int shade_impl(vanerf_ctx* ctx, int precision, const TargetDev& td, const float* rays, const float* z, int R, int S,
                      const float* sdf, const int* nn, const unsigned char* qvis, float* rgba, unsigned char* valid,
                      float* raw_out, float* dbg_latent, cudaStream_t stream, const float* pts_in = nullptr,
                      const float* view_in = nullptr) {
    const long long N = (long long)R * S;
    const int V = ctx->fr.V;
#ifndef VANERF_HOST_EMUL
    if (precision == VANERF_BF16)
        return tc_shade(ctx, td, rays, z, S, N, sdf, nn, qvis, rgba, valid, raw_out, dbg_latent, stream, pts_in, view_in);
#if TC_TAB_PARAM
    // fp32 path: split-precision tensor-core kernel (up to 3 source views); the FFMA kernel below serves V = 4 and the
    // VANERF_FP32_SIMT=1 developer switch
    if (!ctx->fp32_simt && V <= TC_MAXV)
        return tc_shade_split(ctx, td, rays, z, S, N, sdf, nn, qvis, rgba, valid, raw_out, dbg_latent, stream, pts_in, view_in);
#endif
#endif
    const int chunk = (int)(N < SHADE_CHUNK ? N : SHADE_CHUNK);
    ENSURE(ctx, ctx->rec, (size_t)chunk * V * REC_STRIDE * 4);
    const size_t smem = mlp_simt_smem_floats(V) * sizeof(float);
    for (long long s0 = 0; s0 < N; s0 += chunk) {
        const int nc = (int)((N - s0) < chunk ? (N - s0) : chunk);
        const int gblocks = min(cdiv(nc, GATHER_THREADS / 16), ctx->sm_count * 8);
        {
            TimedScope ts(ctx, KCL_GATHER, stream);
            VANERF_LAUNCH(k_gather, gblocks, GATHER_THREADS, 0, stream, ctx->fr, td, rays, z, pts_in, view_in, S, s0, nc, N, sdf,
                          nn, qvis, (float*)ctx->rec.p, valid);
            CHECK_LAUNCH(ctx);
        }
        TimedScope ts(ctx, KCL_MLP, stream);
        const int mblocks = min(cdiv(nc, TS), ctx->sm_count);
        VANERF_LAUNCH(k_mlp_simt, mblocks, MLP_THREADS, smem, stream, (const NetDev*)ctx->netdev.p, ctx->fr.kpt_cam, V,
                      (const float*)ctx->rec.p, s0, nc, rgba, raw_out, dbg_latent);
        CHECK_LAUNCH(ctx);
    }
    return VANERF_OK;
}

int vanerf_shade(vanerf_ctx* ctx, int precision, const vanerf_target* tar, const float* rays, const float* z, int32_t R,
                 int32_t S, const float* sdf, const int32_t* nn_vert, const uint8_t* qvis, float* rgba, uint8_t* valid,
                 float* raw_out, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !tar || !rays || !z || !sdf || !nn_vert || !qvis || R <= 0 || S <= 0) return ctx_invalid(ctx, "vanerf_shade");
    if (!ctx->have_frame || !ctx->have_weights) return ctx_state(ctx);
    if (precision != VANERF_FP32 && precision != VANERF_BF16) return ctx_invalid(ctx, "precision");
    return shade_impl(ctx, precision, make_target(tar), rays, z, R, S, sdf, nn_vert, qvis, rgba, valid, raw_out, nullptr,
                      (cudaStream_t)stream);
}

// test hook: like vanerf_shade (fp32) but also returns the pooled 128-wide latent of MLPUNetFusion (N,128)
int vanerf_shade_debug(vanerf_ctx* ctx, const vanerf_target* tar, const float* rays, const float* z, int32_t R, int32_t S,
                       const float* sdf, const int32_t* nn_vert, const uint8_t* qvis, float* rgba, uint8_t* valid,
                       float* raw_out, float* latent, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !tar) return VANERF_ERR_INVALID;
    if (!ctx->have_frame || !ctx->have_weights) return ctx_state(ctx);
    return shade_impl(ctx, VANERF_FP32, make_target(tar), rays, z, R, S, sdf, nn_vert, qvis, rgba, valid, raw_out, latent,
                      (cudaStream_t)stream);
}

// test hooks of the tensor-core path ------------------------------------------------------------------------------
// vanerf_shade_debug for the bf16 path (latent = pooled [mean | var] before bf16 rounding)
int vanerf_shade_debug_bf16(vanerf_ctx* ctx, const vanerf_target* tar, const float* rays, const float* z, int32_t R, int32_t S,
                            const float* sdf, const int32_t* nn_vert, const uint8_t* qvis, float* rgba, uint8_t* valid,
                            float* raw_out, float* latent, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !tar) return VANERF_ERR_INVALID;
    if (!ctx->have_frame || !ctx->have_weights) return ctx_state(ctx);
    return shade_impl(ctx, VANERF_BF16, make_target(tar), rays, z, R, S, sdf, nn_vert, qvis, rgba, valid, raw_out, latent,
                      (cudaStream_t)stream);
}
// Cycle trace of CTA 0 / thread 0 of the next k_mlp_tc launches: buf dev (capacity, 2) int64 pairs (tag, clock64), NULL = off.
// Returns the number of pairs recorded so far (after a stream synchronise) when buf == NULL.
int vanerf_tc_profile(vanerf_ctx* ctx, long long* buf, int32_t capacity) {
    DeviceGuard dg_(ctx);
#if !defined(VANERF_HOST_EMUL) && defined(VANERF_TC_TRACE)
    if (!ctx) return VANERF_ERR_INVALID;
    int n = 0, zero = 0;
    if (!buf) {
        if (cudaMemcpyFromSymbol(&n, d_tc_prof_n, sizeof(int)) != cudaSuccess) return VANERF_ERR_CUDA;
        capacity = 0;
    }
    if (cudaMemcpyToSymbol(d_tc_prof, &buf, sizeof(buf)) != cudaSuccess || cudaMemcpyToSymbol(d_tc_prof_cap, &capacity, sizeof(int)) != cudaSuccess ||
        cudaMemcpyToSymbol(d_tc_prof_n, &zero, sizeof(int)) != cudaSuccess) return VANERF_ERR_CUDA;
    return n;
#else
    (void)ctx; (void)buf; (void)capacity;          // trace support is compiled in with -DVANERF_TC_TRACE only
    return 0;
#endif
}
// nonzero = a bounded wait inside a tensor-core kernel gave up (valid after the stream has been synchronised)
int vanerf_tc_error(vanerf_ctx* ctx) {
#ifndef VANERF_HOST_EMUL
    return ctx ? *ctx->tc_err_host : 0;
#else
    (void)ctx;
    return 0;
#endif
}
// Completion check of the bf16 path: synchronises `stream`, then returns VANERF_ERR_CUDA (message in vanerf_last_error) and
// clears the record if any tensor-core launch since the last check gave up; VANERF_OK otherwise.
int vanerf_tc_check(vanerf_ctx* ctx, void* stream) {
    if (!ctx) return VANERF_ERR_INVALID;
    DeviceGuard dg_(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize((cudaStream_t)stream));
#ifndef VANERF_HOST_EMUL
    if (tc_take_error(ctx)) return VANERF_ERR_CUDA;
#endif
    return VANERF_OK;
}
// D (128, Npad) = bf16(A (128, K) dev) x bf16(W (N, K) host)^T through one tcgen05 step; K % 16 == 0, K <= 256, N <= 128
int vanerf_tc_selftest(vanerf_ctx* ctx, const float* A_dev, const float* W_host, int32_t K, int32_t N, float* D_dev, void* stream_) {
    DeviceGuard dg_(ctx);
#ifndef VANERF_HOST_EMUL
    if (!ctx || !A_dev || !W_host || !D_dev || K <= 0 || K > 256 || (K & 15) || N <= 0 || N > 128) return ctx_invalid(ctx, "vanerf_tc_selftest");
    cudaStream_t stream = (cudaStream_t)stream_;
    std::vector<TcProg> Pv(1);         // one-step program of this test (heap: the structs are several KB)
    std::vector<TcTables> Tv(1);       // biases etc. are not used by the self test
    memset(&Tv[0], 0, sizeof(TcTables));
    TcProg& P = Pv[0];
    std::vector<uint16_t> img;
    tc_build_single(W_host, N, K, P, img);
    struct Tmp { DevBuf b; ~Tmp() { if (b.p) cudaFree(b.p); } } blob_, tab_;     // freed on every return path
    DevBuf& blob = blob_.b; DevBuf& tab = tab_.b;
    ENSURE(ctx, blob, img.size() * 2);
    ENSURE(ctx, tab, sizeof(TcTables));
    CUDA_TRY(ctx, cudaMemcpyAsync(tab.p, &Tv[0], sizeof(TcTables), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(blob.p, img.data(), img.size() * 2, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(ctx, cudaMemcpyToSymbolAsync(c_prog, &P, sizeof(TcProg), 0, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(stream));          // the host copies above go out of scope with this call
    VANERF_LAUNCH(k_tc_selftest, 1, TC_THREADS, TC_SMEM_BYTES, stream, (const TcTables*)tab.p, (const unsigned char*)blob.p, A_dev, K,
                  (N + 15) & ~15, D_dev, ctx->tc_err_dev);
    CHECK_LAUNCH(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(stream));
    if (tc_take_error(ctx)) return VANERF_ERR_CUDA;
    return VANERF_OK;
#else
    (void)A_dev; (void)W_host; (void)K; (void)N; (void)D_dev; (void)stream_;
    return ctx_invalid(ctx, "tensor-core path needs the CUDA build");
#endif
}

// 0 when the weight-packing script (run here with zero weights, no GPU needed) yields exactly the MMA program the
// tensor-core kernels were compiled with.
int vanerf_tc_program_check(void) {
#ifndef VANERF_HOST_EMUL
    static const int dims[L_COUNT][2] = {
        {10, 196}, {3, 10}, {64, 196}, {64, 64}, {10, 28}, {3, 10}, {8, 28}, {8, 8},
        {128, 358}, {128, 128}, {120, 136}, {64, 120}, {64, 128}, {64, 64}, {2, 64}, {24, 128},
        {96, 96}, {6, 96}, {96, 96}, {40, 96},
        {16, 4}, {40, 16}, {64, 120}, {32, 64}, {32, 32}, {33, 32}, {32, 32}, {1, 32}, {16, 37}, {8, 16}, {1, 8}};
    std::vector<std::vector<float>> w(L_COUNT);
    std::vector<vanerf_linear> lin(L_COUNT);
    const vanerf_linear* src[L_COUNT];
    for (int i = 0; i < L_COUNT; ++i) {
        w[i].assign((size_t)dims[i][0] * dims[i][1], 0.0f);
        lin[i].w = w[i].data(); lin[i].b = nullptr; lin[i].out_dim = dims[i][0]; lin[i].in_dim = dims[i][1];
        src[i] = &lin[i];
    }
    static TcTables T;
    static TcProg P;
    std::vector<uint16_t> img;
    tc_build(src, 0.0f, T, P, img);
    return tc_program_matches(P) ? 0 : 1;
#else
    return 0;
#endif
}

// Developer measurement: pacing of small tcgen05.mma (see k_tc_mma_probe).  out_host: (2 warps, 2) cycles of CTA 0
// [issue, issue + completion]; n_ctas CTAs run the same stream concurrently.
int vanerf_tc_mma_probe(vanerf_ctx* ctx, int32_t n, int32_t reps, int32_t n_acc, int32_t mode, int32_t n_ctas, long long* out_host) {
    DeviceGuard dg_(ctx);
#ifndef VANERF_HOST_EMUL
    if (!ctx || !out_host || n < 16 || n > 256 || (n & 15) || reps <= 0 || n_acc <= 0 || n_acc * n > 256 || n_ctas <= 0) return ctx_invalid(ctx, "vanerf_tc_mma_probe");
    long long* d = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d, (size_t)n_ctas * 4 * sizeof(long long)));
    struct Free { long long* p; ~Free() { cudaFree(p); } } free_d{d};
    CUDA_TRY(ctx, cudaMemset(d, 0, (size_t)n_ctas * 4 * sizeof(long long)));
    k_tc_mma_probe<<<n_ctas, 128, 6 * TC_SLOT + 1024>>>(n, reps, n_acc, mode, d);
    CHECK_LAUNCH(ctx);
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    CUDA_TRY(ctx, cudaMemcpy(out_host, d, 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    return VANERF_OK;
#else
    (void)n; (void)reps; (void)n_acc; (void)mode; (void)n_ctas; (void)out_host;
    return ctx_invalid(ctx, "tensor-core path needs the CUDA build");
#endif
}

// VANeRF.query called directly on arbitrary points (src/model.py:748-877): geometry + gather + networks.
int vanerf_query_points(vanerf_ctx* ctx, int precision, const vanerf_target* tar, const float* pts, const float* view,
                        int32_t N, const float* sdf_in, const uint8_t* qvis_in, float* raw_out, uint8_t* valid, float* rgba,
                        void* stream_) {
    DeviceGuard dg_(ctx);
    if (!ctx || !tar || !pts || !view || N <= 0) return ctx_invalid(ctx, "vanerf_query_points");
    if (!ctx->have_frame || !ctx->have_weights) return ctx_state(ctx);
    if (precision != VANERF_FP32 && precision != VANERF_BF16) return ctx_invalid(ctx, "precision");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int V = ctx->fr.V;
    ENSURE(ctx, ctx->s_sdf, (size_t)N * 4);
    ENSURE(ctx, ctx->s_nn, (size_t)N * 4);
    ENSURE(ctx, ctx->s_qvis, (size_t)N * V);
    const TargetDev td = make_target(tar);
    {
        TimedScope ts(ctx, KCL_GEOM, stream);
        VANERF_LAUNCH(k_geom_query, cdiv(N, 128), 128, 0, stream, ctx->fr, td, (const float*)nullptr, (const float*)nullptr, pts, N, 1,
                      (float*)nullptr, (float*)ctx->s_sdf.p, (int*)nullptr, (int*)ctx->s_nn.p, (unsigned char*)ctx->s_qvis.p);
        CHECK_LAUNCH(ctx);
    }
    const float* sdf = sdf_in ? sdf_in : (const float*)ctx->s_sdf.p;
    const unsigned char* qv = qvis_in ? qvis_in : (const unsigned char*)ctx->s_qvis.p;
    return shade_impl(ctx, precision, td, nullptr, nullptr, N, 1, sdf, (const int*)ctx->s_nn.p, qv, rgba, valid, raw_out, nullptr,
                      stream, pts, view);
}

int vanerf_timing_enable(vanerf_ctx* ctx, int on) {
    if (!ctx) return VANERF_ERR_INVALID;
    ctx->timing = on != 0;
    return VANERF_OK;
}

// Synchronises the recorded events and accumulates; ms_out / count_out have 7 entries:
// setup, rays, geom, gather, mlp, composite, importance.
int vanerf_timing_read(vanerf_ctx* ctx, double* ms_out, int64_t* count_out, int reset) {
    DeviceGuard dg_(ctx);
    if (!ctx || !ms_out || !count_out) return VANERF_ERR_INVALID;
#ifndef VANERF_HOST_EMUL
    cudaError_t first = cudaSuccess;
    for (auto& e : ctx->evs) {
        float ms = 0.f;
        cudaError_t rc = cudaEventSynchronize(e.b);
        if (rc == cudaSuccess) rc = cudaEventElapsedTime(&ms, e.a, e.b);
        if (rc == cudaSuccess) { ctx->t_ms[e.kc] += ms; ctx->t_cnt[e.kc] += 1; }
        else if (first == cudaSuccess) first = rc;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    ctx->evs.clear();
    CUDA_TRY(ctx, first);
#endif
    for (int i = 0; i < KCL_COUNT; ++i) { ms_out[i] = ctx->t_ms[i]; count_out[i] = ctx->t_cnt[i]; }
    if (reset) for (int i = 0; i < KCL_COUNT; ++i) { ctx->t_ms[i] = 0; ctx->t_cnt[i] = 0; }
    return VANERF_OK;
}

int vanerf_composite(vanerf_ctx* ctx, const float* rgba, const float* z, const float* mesh_sdf, int32_t R, int32_t S,
                     float* color, float* depth, float* alpha, float* sdf_out, float* contrib, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !rgba || !z || !mesh_sdf || R <= 0 || S <= 0) return ctx_invalid(ctx, "vanerf_composite");
    if (S > 32 * COMP_MAX_PER_LANE) return ctx_unsupported(ctx, "vanerf_composite: too many samples per ray");
    if (!ctx->have_weights) return ctx_state(ctx);
    const int blocks = min(cdiv(R, COMP_WARPS), ctx->sm_count * 16);
    TimedScope ts(ctx, KCL_COMPOSITE, (cudaStream_t)stream);
    VANERF_LAUNCH(k_composite, blocks, COMP_WARPS * 32, 0, stream, rgba, z, mesh_sdf, R, S, ctx->h_net.beta, color, depth, alpha,
                  sdf_out, contrib);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}

int vanerf_importance(vanerf_ctx* ctx, const float* contrib, const float* z, int32_t R, int32_t S, const float* u,
                      int32_t nf, int32_t u_per_ray, float* z_fine_only, float* z_out, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !contrib || !z || !u || !z_out || R <= 0 || S < 3 || nf <= 0) return ctx_invalid(ctx, "vanerf_importance");
    const float* zmid_in = nullptr;
    const size_t smem = (size_t)COMP_WARPS * (2 * (S - 1) + S + nf) * sizeof(float);
    if (smem > 48 * 1024) return ctx_unsupported(ctx, "importance sampling: depths per ray exceed the shared-memory budget");
    const int blocks = min(cdiv(R, COMP_WARPS), ctx->sm_count * 16);
    TimedScope ts(ctx, KCL_IMPORTANCE, (cudaStream_t)stream);
    VANERF_LAUNCH(k_importance, blocks, COMP_WARPS * 32, smem, stream, contrib, z, zmid_in, R, S, u, nf, u_per_ray, z_fine_only, z_out,
                  ctx->imp_src_map);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}

int vanerf_set_reuse_coarse(vanerf_ctx* ctx, int on) {
    if (!ctx) return VANERF_ERR_INVALID;
    ctx->reuse_coarse = on != 0;
    return VANERF_OK;
}

int vanerf_set_reuse_geometry(vanerf_ctx* ctx, int on) {
    if (!ctx) return VANERF_ERR_INVALID;
    ctx->reuse_geometry = on != 0;
    return VANERF_OK;
}

// Reference calling convention of VANeRF.importance_sample (src/model.py:1425-1462): contrib_inner (R, D-2),
// z_mid (R, D-1) -> z_fine (R, n_fine); no merge.
int vanerf_importance_mid(vanerf_ctx* ctx, const float* contrib_inner, const float* z_mid, int32_t R, int32_t D, const float* u,
                          int32_t nf, int32_t u_per_ray, float* z_fine, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !contrib_inner || !z_mid || !u || !z_fine || R <= 0 || D < 3 || nf <= 0) return ctx_invalid(ctx, "vanerf_importance_mid");
    const size_t smem = (size_t)COMP_WARPS * (2 * (D - 1) + D + nf) * sizeof(float);
    if (smem > 48 * 1024) return ctx_unsupported(ctx, "importance sampling: depths per ray exceed the shared-memory budget");
    const int blocks = min(cdiv(R, COMP_WARPS), ctx->sm_count * 16);
    TimedScope ts(ctx, KCL_IMPORTANCE, (cudaStream_t)stream);
    VANERF_LAUNCH(k_importance, blocks, COMP_WARPS * 32, smem, stream, contrib_inner, (const float*)nullptr, z_mid, R, D, u, nf,
                  u_per_ray, z_fine, (float*)nullptr, (unsigned char*)nullptr);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}

// ------------------------------------------------------------------------------------------------ per-frame global feature
// TexVisFusion global vertex feature (src/networks.py:246-279) as kernels: gfeat.cuh.  All pointers are device pointers.
int vanerf_global_vertex_feature(vanerf_ctx* ctx, const vanerf_gfeat_weights* w, const float* img, const float* tex, int32_t V, int32_t H,
                                 int32_t W, int32_t th, int32_t tw, float* out, void* stream_) {
    DeviceGuard dg_(ctx);
#ifndef VANERF_HOST_EMUL
    if (!ctx || !w || !img || !tex || !out || V <= 0 || V > MAXV || H <= 0 || W <= 0 || th <= 0 || tw <= 0) return ctx_invalid(ctx, "vanerf_global_vertex_feature");
    const vanerf_conv_stack* st3[3] = {&w->img, &w->tex, &w->gt};
    for (const vanerf_conv_stack* s : st3)
        if (!s->conv0 || !s->ln1_w || !s->ln1_b || !s->conv3 || !s->ln4_w || !s->ln4_b) return ctx_invalid(ctx, "vanerf_global_vertex_feature: NULL weight");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t HW = (size_t)H * W, hw = (size_t)th * tw;
    // scratch: stats (4 x V x 42 x 2 doubles) | stat partials (V x blocks x 84 doubles) | pooled img, tex (V,42,9 each) | pooled
    // partials (chunks x 2 x V,42,9) | mid (V,779,18) | y1 img | y2 img | y1 tex | y2 tex
    const dim3 gi(cdiv(W, GF_TILE), cdiv(H, GF_TILE), V), gt(cdiv(tw, GF_TILE), cdiv(th, GF_TILE), V);
    const int nbi = gi.x * gi.y, nbt = gt.x * gt.y;
    const size_t n_stat = (size_t)4 * V * GF_OUT * 2;
    size_t off = n_stat * 8;
    const size_t o_part = off; off += (size_t)V * nbi * 2 * GF_OUT * 8;
    const size_t o_pi = off; off += (size_t)V * GF_OUT * 9 * 4;
    const size_t o_pt = off; off += (size_t)V * GF_OUT * 9 * 4;
    const size_t o_pp = off; off += (size_t)GF_POOL_CHUNKS * V * GF_OUT * 9 * 4;
    const size_t o_mid = off; off += (size_t)V * NUM_V_HAND * 18 * 4;
    off = (off + 255) & ~(size_t)255;
    const size_t o_y1i = off; off += (size_t)V * GF_MID * HW * 4;
    const size_t o_y2i = off; off += (size_t)V * GF_OUT * HW * 4;
    const size_t o_y1t = off; off += (size_t)V * GF_MID * hw * 4;
    const size_t o_y2t = off; off += (size_t)V * GF_OUT * hw * 4;
    ENSURE(ctx, ctx->gf_scratch, off);
    unsigned char* q = (unsigned char*)ctx->gf_scratch.p;
    double* st = (double*)q;
    double *st1i = st, *st2i = st + (size_t)V * GF_OUT * 2, *st1t = st + (size_t)2 * V * GF_OUT * 2, *st2t = st + (size_t)3 * V * GF_OUT * 2;
    double* part = (double*)(q + o_part);
    float *pi = (float*)(q + o_pi), *pt = (float*)(q + o_pt), *pp = (float*)(q + o_pp), *mid = (float*)(q + o_mid);
    float *y1i = (float*)(q + o_y1i), *y2i = (float*)(q + o_y2i), *y1t = (float*)(q + o_y1t), *y2t = (float*)(q + o_y2t);
    TimedScope ts(ctx, KCL_SETUP, stream);
    const int T = GF_TILE * GF_TILE;
    // per-view statistics: the (V, blocks, 2 C) partials are reduced per view (part rows of view v are contiguous)
    auto reduce_stats = [&](int nb, int C, double* dst) -> int {
        k_gf_reduce<double><<<dim3(cdiv(2 * C * 32, 128), V), 128, 0, stream>>>(part, nb, 2 * C, dst);
        CHECK_LAUNCH(ctx);
        return VANERF_OK;
    };
    const int np = V * GF_OUT * 9;
    int rc;
    k_gf_conv3x3<3, GF_MID, false><<<gi, T, 0, stream>>>(img, w->img.conv0, H, W, nullptr, nullptr, nullptr, y1i, part); CHECK_LAUNCH(ctx);
    if ((rc = reduce_stats(nbi, GF_MID, st1i))) return rc;
    k_gf_conv3x3<GF_MID, GF_OUT, true><<<gi, T, 0, stream>>>(y1i, w->img.conv3, H, W, st1i, w->img.ln1_w, w->img.ln1_b, y2i, part); CHECK_LAUNCH(ctx);
    if ((rc = reduce_stats(nbi, GF_OUT, st2i))) return rc;
    k_gf_norm_pool<<<dim3(GF_OUT, V, GF_POOL_CHUNKS), 256, 0, stream>>>(y2i, GF_OUT, H, W, st2i, w->img.ln4_w, w->img.ln4_b, pp); CHECK_LAUNCH(ctx);
    k_gf_reduce<float><<<dim3(cdiv(np * 32, 128), 1), 128, 0, stream>>>(pp, GF_POOL_CHUNKS, np, pi); CHECK_LAUNCH(ctx);
    k_gf_conv3x3<8, GF_MID, false><<<gt, T, 0, stream>>>(tex, w->tex.conv0, th, tw, nullptr, nullptr, nullptr, y1t, part); CHECK_LAUNCH(ctx);
    if ((rc = reduce_stats(nbt, GF_MID, st1t))) return rc;
    k_gf_conv3x3<GF_MID, GF_OUT, true><<<gt, T, 0, stream>>>(y1t, w->tex.conv3, th, tw, st1t, w->tex.ln1_w, w->tex.ln1_b, y2t, part); CHECK_LAUNCH(ctx);
    if ((rc = reduce_stats(nbt, GF_OUT, st2t))) return rc;
    k_gf_norm_pool<<<dim3(GF_OUT, V, GF_POOL_CHUNKS), 256, 0, stream>>>(y2t, GF_OUT, th, tw, st2t, w->tex.ln4_w, w->tex.ln4_b, pp); CHECK_LAUNCH(ctx);
    k_gf_reduce<float><<<dim3(cdiv(np * 32, 128), 1), 128, 0, stream>>>(pp, GF_POOL_CHUNKS, np, pt); CHECK_LAUNCH(ctx);
    k_gf_conv1d_ln<<<cdiv((long long)V * NUM_V_HAND * 32, 128), 128, 0, stream>>>(pi, pt, w->gt.conv0, V, GF_OUT, NUM_V_HAND, w->gt.ln1_w, w->gt.ln1_b, mid); CHECK_LAUNCH(ctx);
    k_gf_conv1d_ln<<<cdiv((long long)V * 2 * NUM_V_HAND * 32, 128), 128, 0, stream>>>(mid, nullptr, w->gt.conv3, V, NUM_V_HAND, 2 * NUM_V_HAND, w->gt.ln4_w, w->gt.ln4_b, out); CHECK_LAUNCH(ctx);
    return VANERF_OK;
#else
    (void)w; (void)img; (void)tex; (void)V; (void)H; (void)W; (void)th; (void)tw; (void)out; (void)stream_;
    return ctx_unsupported(ctx, "vanerf_global_vertex_feature needs the CUDA build");
#endif
}

// ------------------------------------------------------------------------------------------------ stage primitives
int vanerf_feat_sample(vanerf_ctx* ctx, const float* feat, int32_t B, int32_t C, int32_t H, int32_t W, const float* uv, int32_t N,
                       float* out, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !feat || !uv || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0) return ctx_invalid(ctx, "vanerf_feat_sample");
    VANERF_LAUNCH(k_feat_sample, cdiv((long long)B * N * C, 256), 256, 0, stream, feat, B, C, H, W, uv, N, out);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}
int vanerf_knn1(vanerf_ctx* ctx, const float* query, int32_t N, const float* vert, int32_t Nv, int32_t* idx, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !query || !vert || !idx || N <= 0 || Nv <= 0) return ctx_invalid(ctx, "vanerf_knn1");
    VANERF_LAUNCH(k_knn1, cdiv(N, 128), 128, 0, stream, query, N, vert, Nv, idx);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}
int vanerf_dense(vanerf_ctx* ctx, const float* x, int32_t M, int32_t K, const float* w, const float* b, int32_t N, int32_t act,
                 float* y, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !x || !w || !y || M <= 0 || K <= 0 || N <= 0 || act < 0 || act > SA_ELU) return ctx_invalid(ctx, "vanerf_dense");
    VANERF_LAUNCH(k_dense, cdiv((long long)M * N, 256), 256, 0, stream, x, M, K, w, b, N, act, y);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}
int vanerf_rel_z_decay(vanerf_ctx* ctx, const float* cxyz, const float* kxyz, int32_t BV, int32_t N, int32_t n_kpt, int32_t levels,
                       float scale, float sigma, float* out, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !cxyz || !kxyz || !out || BV <= 0 || N <= 0 || n_kpt <= 0 || levels < 0) return ctx_invalid(ctx, "vanerf_rel_z_decay");
    VANERF_LAUNCH(k_rel_z_decay, cdiv((long long)BV * N * n_kpt, 256), 256, 0, stream, cxyz, kxyz, BV, N, n_kpt, levels, scale, sigma, out);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}

// ------------------------------------------------------------------------------------------------ training branch
int vanerf_project_samples(vanerf_ctx* ctx, const vanerf_target* tar, const float* rays, const float* z, int32_t R, int32_t S,
                           float* xy, uint8_t* mask, float* pw_raw, float* cam, float* ray_diff, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !tar || !rays || !z || !xy || !mask || !pw_raw || !cam || !ray_diff || R <= 0 || S <= 0) return ctx_invalid(ctx, "vanerf_project_samples");
    if (!ctx->have_frame) return ctx_state(ctx);
    const long long N = (long long)R * S;
    VANERF_LAUNCH(k_project_samples, cdiv(N, 128), 128, 0, stream, ctx->fr, make_target(tar), rays, z, S, N, xy, mask, pw_raw, cam, ray_diff);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}
int vanerf_feat_sample_bwd(vanerf_ctx* ctx, const float* d_out, int32_t B, int32_t C, int32_t H, int32_t W, const float* uv, int32_t N,
                           float* d_feat, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !d_out || !uv || !d_feat || B <= 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0) return ctx_invalid(ctx, "vanerf_feat_sample_bwd");
    VANERF_LAUNCH(k_feat_sample_bwd, cdiv((long long)B * N * C, 256), 256, 0, stream, d_out, B, C, H, W, uv, N, d_feat);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}
int vanerf_composite_bwd(vanerf_ctx* ctx, const float* rgba, const float* z, const float* mesh_sdf, int32_t R, int32_t S, float beta,
                         const float* g_color, const float* g_alpha, const float* g_depth, const float* g_sdf, float* d_rgba, float* d_beta,
                         void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !rgba || !z || !mesh_sdf || !d_rgba || R <= 0 || S <= 0 || !(beta > 0.0f)) return ctx_invalid(ctx, "vanerf_composite_bwd");
    if (S > 32 * COMP_MAX_PER_LANE) return ctx_unsupported(ctx, "vanerf_composite_bwd: too many samples per ray");
    const int blocks = min(cdiv(R, COMP_WARPS), ctx->sm_count * 16);
    VANERF_LAUNCH(k_composite_bwd, blocks, COMP_WARPS * 32, 0, stream, rgba, z, mesh_sdf, R, S, beta, g_color, g_alpha, g_depth, g_sdf, d_rgba, d_beta);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}
// rgba2out with an explicit beta (the training graph owns sigmoid_beta as a parameter)
int vanerf_composite_beta(vanerf_ctx* ctx, const float* rgba, const float* z, const float* mesh_sdf, int32_t R, int32_t S, float beta,
                          float* color, float* depth, float* alpha, float* sdf_out, float* contrib, void* stream) {
    DeviceGuard dg_(ctx);
    if (!ctx || !rgba || !z || !mesh_sdf || R <= 0 || S <= 0 || !(beta > 0.0f)) return ctx_invalid(ctx, "vanerf_composite_beta");
    if (S > 32 * COMP_MAX_PER_LANE) return ctx_unsupported(ctx, "vanerf_composite_beta: too many samples per ray");
    const int blocks = min(cdiv(R, COMP_WARPS), ctx->sm_count * 16);
    TimedScope ts(ctx, KCL_COMPOSITE, (cudaStream_t)stream);
    VANERF_LAUNCH(k_composite, blocks, COMP_WARPS * 32, 0, stream, rgba, z, mesh_sdf, R, S, beta, color, depth, alpha, sdf_out, contrib);
    CHECK_LAUNCH(ctx);
    return VANERF_OK;
}

size_t vanerf_scratch_bytes(const vanerf_ctx* ctx, int32_t R, int32_t S) {
    const int V = ctx && ctx->have_frame ? ctx->fr.V : MAXV;
    const size_t N = (size_t)R * S;
    return N * (4 + 4 + 4 + V + 20 + 4 + 1) + (size_t)R * 32 + (size_t)SHADE_CHUNK * V * REC_STRIDE * 4;
}

// rays per internal chunk of vanerf_render_rays: 65 536 rays x 128 depths = 8.4 M samples (~0.4 GB of per-sample scratch); with 16 tile
// pairs per shading launch a 334x512 view is ~140 kernel launches
#define RENDER_RAY_CHUNK 65536

int vanerf_render_rays(vanerf_ctx* ctx, int precision, const vanerf_target* tar, const int32_t* pix_xy, int32_t R,
                       int32_t Sc, int32_t Sf, int32_t fine, const float* ztab, const float* utab, float* out_coarse,
                       float* out_fine, void* stream_) {
    DeviceGuard dg_(ctx);
    if (!ctx || !tar || !pix_xy || !ztab || !out_coarse || R <= 0 || Sc < 3) return ctx_invalid(ctx, "vanerf_render_rays");
    if (fine && (!utab || !out_fine || Sf <= 0)) return ctx_invalid(ctx, "vanerf_render_rays: fine pass needs utab/out_fine");
    if (!ctx->have_frame || !ctx->have_weights) return ctx_state(ctx);
    if (precision != VANERF_FP32 && precision != VANERF_BF16) return ctx_invalid(ctx, "precision");
    cudaStream_t stream = (cudaStream_t)stream_;
    const TargetDev td = make_target(tar);
    const int V = ctx->fr.V;
    const int S2 = fine ? Sc + Sf : Sc;
    const int RC = R < RENDER_RAY_CHUNK ? R : RENDER_RAY_CHUNK;
    const size_t Nmax = (size_t)RC * S2;
    ENSURE(ctx, ctx->s_rays, (size_t)RC * VANERF_RAY_STRIDE * 4);
    ENSURE(ctx, ctx->s_z, (size_t)RC * Sc * 4);
    ENSURE(ctx, ctx->s_z2, Nmax * 4);
    ENSURE(ctx, ctx->s_sdf, Nmax * 4);
    ENSURE(ctx, ctx->s_nn, Nmax * 4);
    ENSURE(ctx, ctx->s_qvis, Nmax * V);
    ENSURE(ctx, ctx->s_rgba, Nmax * 20);
    ENSURE(ctx, ctx->s_contrib, (size_t)RC * Sc * 4);
    float* rays = (float*)ctx->s_rays.p; float* z = (float*)ctx->s_z.p; float* z2 = (float*)ctx->s_z2.p;
    float* sdf = (float*)ctx->s_sdf.p; int* nn = (int*)ctx->s_nn.p; unsigned char* qv = (unsigned char*)ctx->s_qvis.p;
    float* rgba = (float*)ctx->s_rgba.p; float* contrib = (float*)ctx->s_contrib.p;
    for (int r0 = 0; r0 < R; r0 += RC) {
        const int rc = (R - r0) < RC ? (R - r0) : RC;
        int st;
        float* oc = out_coarse + (size_t)r0 * 8;
        if ((st = vanerf_sample_rays(ctx, tar, pix_xy + 2 * (size_t)r0, rc, ztab, Sc, rays, z, stream_))) return st;
        if ((st = vanerf_geom_query(ctx, tar, rays, z, rc, Sc, nullptr, sdf, nullptr, nn, qv, stream_))) return st;
        if ((st = shade_impl(ctx, precision, td, rays, z, rc, Sc, sdf, nn, qv, rgba, nullptr, nullptr, nullptr, stream))) return st;
        // composite writes planes (color | depth | alpha | sdf); k_pack_out interleaves them into (R,8) rows
        ENSURE(ctx, ctx->s_tab, (size_t)RC * 6 * 4);
        float* pl = (float*)ctx->s_tab.p;      // color (RC,3) | depth | alpha | sdf
        if ((st = vanerf_composite(ctx, rgba, z, sdf, rc, Sc, pl, pl + 3 * (size_t)RC, pl + 4 * (size_t)RC, pl + 5 * (size_t)RC, contrib, stream_))) return st;
        VANERF_LAUNCH(k_pack_out, cdiv(rc, 256), 256, 0, stream, pl, RC, rc, oc); CHECK_LAUNCH(ctx);
        if (fine) {
            float* of = out_fine + (size_t)r0 * 8;
            if (ctx->reuse_coarse) {
                // The merged fine set contains the Sc coarse depths bit for bit, and a sample's outputs depend on nothing
                // but its point and ray: the reference re-evaluates those Sc samples (src/model.py:1301-1349), here only
                // the Sf new depths go through geometry / gather / networks and the merged arrays are assembled from the
                // two evaluations (k_merge_reuse).  Same bits out, a third fewer evaluations per ray.
                ENSURE(ctx, ctx->s_zf, (size_t)RC * Sf * 4); ENSURE(ctx, ctx->s_srcmap, Nmax);
                ENSURE(ctx, ctx->s_sdf_f, (size_t)RC * Sf * 4); ENSURE(ctx, ctx->s_rgba_f, (size_t)RC * Sf * 20);
                ENSURE(ctx, ctx->s_sdf_m, Nmax * 4); ENSURE(ctx, ctx->s_rgba_m, Nmax * 20);
                float* zf = (float*)ctx->s_zf.p; float* sdf_f = (float*)ctx->s_sdf_f.p; float* rgba_f = (float*)ctx->s_rgba_f.p;
                float* sdf_m = (float*)ctx->s_sdf_m.p; float* rgba_m = (float*)ctx->s_rgba_m.p;
                ctx->imp_src_map = (unsigned char*)ctx->s_srcmap.p;
                st = vanerf_importance(ctx, contrib, z, rc, Sc, utab, Sf, 0, zf, z2, stream_);
                ctx->imp_src_map = nullptr;
                if (st) return st;
                if ((st = vanerf_geom_query(ctx, tar, rays, zf, rc, Sf, nullptr, sdf_f, nullptr, nn, qv, stream_))) return st;
                if ((st = shade_impl(ctx, precision, td, rays, zf, rc, Sf, sdf_f, nn, qv, rgba_f, nullptr, nullptr, nullptr, stream))) return st;
                VANERF_LAUNCH(k_merge_reuse, cdiv((long long)rc * S2, 256), 256, 0, stream, (const unsigned char*)ctx->s_srcmap.p, rgba, sdf,
                              rgba_f, sdf_f, rc, Sc, Sf, rgba_m, sdf_m);
                CHECK_LAUNCH(ctx);
                if ((st = vanerf_composite(ctx, rgba_m, z2, sdf_m, rc, S2, pl, pl + 3 * (size_t)RC, pl + 4 * (size_t)RC, pl + 5 * (size_t)RC, nullptr, stream_))) return st;
            } else if (ctx->reuse_geometry) {
                // Default: the networks evaluate all S2 merged samples like the reference (src/model.py:1328-1349), but the
                // mesh queries (cal_vis_sdf_batch / knn_points: functions of the sample position only) run for the Sf new
                // depths only; the Sc coarse depths sit in the merged set bit for bit and keep their coarse-pass results.
                ENSURE(ctx, ctx->s_zf, (size_t)RC * Sf * 4); ENSURE(ctx, ctx->s_srcmap, Nmax);
                ENSURE(ctx, ctx->s_sdf_f, (size_t)RC * Sf * 4); ENSURE(ctx, ctx->s_nn_f, (size_t)RC * Sf * 4);
                ENSURE(ctx, ctx->s_qvis_f, (size_t)RC * Sf * V);
                ENSURE(ctx, ctx->s_sdf_m, Nmax * 4); ENSURE(ctx, ctx->s_nn_m, Nmax * 4); ENSURE(ctx, ctx->s_qvis_m, Nmax * V);
                float* zf = (float*)ctx->s_zf.p; float* sdf_f = (float*)ctx->s_sdf_f.p; int* nn_f = (int*)ctx->s_nn_f.p;
                unsigned char* qv_f = (unsigned char*)ctx->s_qvis_f.p;
                float* sdf_m = (float*)ctx->s_sdf_m.p; int* nn_m = (int*)ctx->s_nn_m.p; unsigned char* qv_m = (unsigned char*)ctx->s_qvis_m.p;
                ctx->imp_src_map = (unsigned char*)ctx->s_srcmap.p;
                st = vanerf_importance(ctx, contrib, z, rc, Sc, utab, Sf, 0, zf, z2, stream_);
                ctx->imp_src_map = nullptr;
                if (st) return st;
                if ((st = vanerf_geom_query(ctx, tar, rays, zf, rc, Sf, nullptr, sdf_f, nullptr, nn_f, qv_f, stream_))) return st;
                {
                    TimedScope ts(ctx, KCL_GEOM, stream);
                    VANERF_LAUNCH(k_merge_geom, cdiv((long long)rc * S2, 256), 256, 0, stream, (const unsigned char*)ctx->s_srcmap.p, sdf, nn, qv,
                                  sdf_f, nn_f, qv_f, rc, Sc, Sf, V, sdf_m, nn_m, qv_m);
                    CHECK_LAUNCH(ctx);
                }
                if ((st = shade_impl(ctx, precision, td, rays, z2, rc, S2, sdf_m, nn_m, qv_m, rgba, nullptr, nullptr, nullptr, stream))) return st;
                if ((st = vanerf_composite(ctx, rgba, z2, sdf_m, rc, S2, pl, pl + 3 * (size_t)RC, pl + 4 * (size_t)RC, pl + 5 * (size_t)RC, nullptr, stream_))) return st;
            } else {
            if ((st = vanerf_importance(ctx, contrib, z, rc, Sc, utab, Sf, 0, nullptr, z2, stream_))) return st;
            if ((st = vanerf_geom_query(ctx, tar, rays, z2, rc, S2, nullptr, sdf, nullptr, nn, qv, stream_))) return st;
            if ((st = shade_impl(ctx, precision, td, rays, z2, rc, S2, sdf, nn, qv, rgba, nullptr, nullptr, nullptr, stream))) return st;
            if ((st = vanerf_composite(ctx, rgba, z2, sdf, rc, S2, pl, pl + 3 * (size_t)RC, pl + 4 * (size_t)RC, pl + 5 * (size_t)RC, nullptr, stream_))) return st;
            }
            VANERF_LAUNCH(k_pack_out, cdiv(rc, 256), 256, 0, stream, pl, RC, rc, of); CHECK_LAUNCH(ctx);
        }
    }
    return VANERF_OK;
}

}  // extern "C"
