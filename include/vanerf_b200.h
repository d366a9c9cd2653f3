/* libvanerf_b200.so — C ABI of the B200-native VANeRF novel-view render path.
 *
 * The reference (XuanHuang0/VANeRF) has no FFI / operator interface: its hot path is Python methods over torch.
 * Each entry point below replaces the body of the reference function(s) cited next to it; the Python class
 * `vanerf_b200.model.VANeRF` keeps the reference's call surface (src/model.py:748,1027,1103,1425,1465,1497) and
 * binds these symbols with ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; no torch types cross the boundary.
 *   - "dev" pointers are device pointers owned by the caller (e.g. the PyTorch caching allocator), "host" pointers
 *     are host memory read synchronously during the call.
 *   - every call enqueues its work on `stream` (a cudaStream_t passed as void*) and returns without synchronising the
 *     stream, except vanerf_load_weights (packs from host memory once per model), vanerf_timing_read, vanerf_tc_check,
 *     vanerf_tc_selftest and vanerf_tc_mma_probe, which say so.  vanerf_frame_setup stages its host inputs in pinned
 *     buffers owned by the context (it waits at most for the upload it issued two frames earlier).  A context grows its
 *     scratch with cudaMalloc on first use / on larger sizes.
 *   - the context owns only packed weights, per-frame acceleration structures, staging and scratch; it holds no
 *     process-global mutable state, and every entry point runs on the context's device and restores the caller's.
 *   - return value: 0 on success, a negative vanerf_status otherwise; no exceptions cross the ABI.
 *   - one vanerf_ctx per device per process; calls on a context are ordered by the stream.
 *   - sample index is the fastest dimension of every (ray, sample) array, as in the reference
 *     (eval_pts.view(B,-1,3), src/model.py:1234-1235).
 */
#ifndef VANERF_B200_H
#define VANERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VANERF_MAX_VIEWS 4       /* source views of the fp32 path */
#define VANERF_MAX_VIEWS_BF16 3  /* source views of the bf16 tensor-core path (TMEM column budget of the batched rendering head);
                                    vanerf_shade / vanerf_render_rays / vanerf_query_points with VANERF_BF16 and more views
                                    return VANERF_ERR_UNSUPPORTED, nothing is truncated silently */
#define VANERF_N_KPT 42          /* configs/vanerf.json: sp_args.n_kpt */
#define VANERF_N_VERT 1558       /* 2 x (778 MANO + 1 seal vertex), src/networks.py:25 */
#define VANERF_RAY_STRIDE 8      /* floats per ray record: dir.xyz, near, far, hit, pad, pad */

typedef enum {
    VANERF_OK = 0,
    VANERF_ERR_INVALID = -1,     /* bad argument (NULL pointer, size out of range) */
    VANERF_ERR_CUDA = -2,        /* CUDA runtime error (see vanerf_last_error) */
    VANERF_ERR_STATE = -3,       /* call order violated (weights / frame not loaded) */
    VANERF_ERR_UNSUPPORTED = -4  /* configuration outside the compiled limits */
} vanerf_status;

typedef enum { VANERF_FP32 = 0, VANERF_BF16 = 1 } vanerf_precision;

typedef struct vanerf_ctx vanerf_ctx;

/* One (out,in) row-major fp32 matrix with optional bias, host memory.  Weight-norm is already folded
 * (W = g * v / ||v||, src/utils.py:670-685); Conv1d(k=1) weights have their trailing axis dropped. */
typedef struct {
    const float* w;
    const float* b;              /* NULL = no bias */
    int32_t out_dim, in_dim;
} vanerf_linear;

/* Hot-path parameters, host pointers. */
typedef struct {
    vanerf_linear geo_at[2], geo_f[2];      /* GeoVisFusion.fconv_at / fconv_ated   (src/networks.py:47-58)  */
    vanerf_linear geo8_at[2], geo8_f[2];    /* GeoVisFusion.fconv_at1 / fconv_ated1 (src/networks.py:60-71)  */
    vanerf_linear mlp[4];                   /* MLPUNet layers1 358-128-128-(136)120-64 (src/utils.py:822-852) */
    vanerf_linear post[3];                  /* MLP layers2 128-64-64-2              (src/utils.py:687-719)   */
    vanerf_linear compress;                 /* ibr_compress_gfeat 128-24            (src/model.py:633,921)   */
    vanerf_linear tex_at[2], tex_f[2];      /* TexVisFusion.fconv_at / fconv        (src/networks.py:224-235) */
    vanerf_linear ray[2], base[2], vis1[2], vis2[2], outl[3];   /* IBRRenderingHead (src/model.py:1578-1591) */
    float ani_al;                           /* IBRRenderingHead.ani_al */
    float sigmoid_beta;                     /* VANeRF.sigmoid_beta (clamped to >= 2e-3 by the library, model.py:880) */
} vanerf_weights;

/* Per-frame inputs: source cameras, source-view maps, two-hand mesh (reference layouts, fp32, NCHW). */
typedef struct {
    int32_t n_views, height, width;         /* source views V and source image size */
    float znear, zfar;                      /* cam_in["znear"/"zfar"] */
    float z_range;                          /* (float)(zfar - znear) evaluated in double, like the reference's python floats */
    const float* KRT;                       /* host (V,4,4)  cam_in["KRT"]                     src/model.py:312   */
    const float* extrin;                    /* host (V,4,4)  sp_data["extrin"]                 src/model.py:306   */
    const float* src_cam_pos;               /* host (V,3)    inverse(KRT)[:, :3, 3]            src/model.py:937   */
    const float* kpt3d;                     /* host (42,3)   sp_data["kpt3d"]                                     */
    const float* verts;                     /* host (n_verts,3) targets["vert_world"]                             */
    const int32_t* faces;                   /* host (n_faces,3) targets["face_world"]                             */
    int32_t n_verts, n_faces;
    const float* img;                       /* dev (V,3,H,W)                                                     */
    const uint8_t* fg_mask;                 /* dev (V,H,W) 0/1   src_foreground_mask                              */
    const float* feat_geo0; int32_t g0_h, g0_w;   /* dev (V,64,h,w) geo encoder level 0  src/networks.py:76      */
    const float* feat_geo1; int32_t g1_h, g1_w;   /* dev (V,8,h,w)  geo encoder level 1  src/networks.py:77      */
    const float* feat_tex;  int32_t t_h, t_w;     /* dev (V,8,h,w)  tex encoder          src/networks.py:78      */
    const float* vert_gfeat;                /* dev (V,n_verts,18) TexVisFusion global feature (fconv_gt output,
                                               src/networks.py:273-279), produced per frame by torch */
} vanerf_frame;

/* Target camera of one render (per-frame 3x3 matrices come from the same torch calls as the reference). */
typedef struct {
    float inv_K[9];      /* inverse(K[:3,:3]).T, row-major: d_cam = [x,y,1] @ inv_K      src/model.py:1208 */
    float R[9];          /* RT[:3,:3] row-major: d_world = d_cam @ R                     src/model.py:1212 */
    float cam_pos[3];    /* -(t^T R)                                                     src/model.py:1213 */
    float znear, zfar;   /* frustum distances before the box clip                        src/model.py:1137 */
    float bounds[6];     /* config["bounds"] (2,3): min xyz, max xyz (offset +-0.01 applied inside) :1216 */
} vanerf_target;

/* Per-frame TexVisFusion convolution stacks (src/networks.py:238-262), DEVICE pointers, fp32, reference layouts:
 * conv0 / conv3 = the two bias-free convolutions, ln1 / ln4 = LayerNorm affine maps.
 *   img (fconv4): conv0 (21,3,3,3),  ln1 (H,W),   conv3 (42,21,3,3),   ln4 (H,W)
 *   tex (fconv3): conv0 (21,8,3,3),  ln1 (h,w),   conv3 (42,21,3,3),   ln4 (h,w)
 *   gt (fconv_gt): conv0 (779,42,3), ln1 (18),    conv3 (1558,779,3),  ln4 (18) */
typedef struct { const float *conv0, *ln1_w, *ln1_b, *conv3, *ln4_w, *ln4_b; } vanerf_conv_stack;
typedef struct { vanerf_conv_stack img, tex, gt; } vanerf_gfeat_weights;

int  vanerf_ctx_create(vanerf_ctx** out, int device);
void vanerf_ctx_destroy(vanerf_ctx* ctx);
const char* vanerf_status_str(int status);
const char* vanerf_last_error(const vanerf_ctx* ctx);
int  vanerf_sm_count(const vanerf_ctx* ctx);

/* Packs (transposes, pads, bf16-converts) the parameters into the context.  Replaces module construction +
 * load_state_dict for the hot path (src/model.py:134-138, :605-667). */
int vanerf_load_weights(vanerf_ctx* ctx, const vanerf_weights* w, void* stream);

/* Per-frame setup: NHWC repack of the maps, vertex projection + per-view vertex visibility raster
 * (src/model.py:1244-1255, mesh_util.py:284-318,484-489), visibility-premultiplied vertex feature tables
 * (src/networks.py:83,96,270-279), camera-space keypoints (src/spatial.py:74-84), triangle / vertex BVHs.
 * vert_vis_out: optional dev (V,n_verts) fp32 copy of the visibility table. */
int vanerf_frame_setup(vanerf_ctx* ctx, const vanerf_frame* f, float* vert_vis_out, void* stream);

/* TexVisFusion global vertex feature of one frame (src/networks.py:273-279: fconv4(img), fconv3(feat_tex) -> (42,18) ->
 * fconv_gt) as kernels: img dev (V,3,H,W), tex dev (V,8,th,tw) -> out dev (V,1558,18), the `vert_gfeat` input of
 * vanerf_frame_setup.  LayerNorm shapes follow the map sizes. */
int vanerf_global_vertex_feature(vanerf_ctx* ctx, const vanerf_gfeat_weights* w, const float* img, const float* tex, int32_t n_views,
                                 int32_t height, int32_t width, int32_t tex_h, int32_t tex_w, float* out, void* stream);

/* Ray generation, box clip, coarse depths (src/model.py:1190-1238, :1497-1570).
 * pix_xy dev (R,2) int32 target pixels; ztab dev (S) = linspace(0,1,S); rays dev (R,8) out; z dev (R,S) out. */
int vanerf_sample_rays(vanerf_ctx* ctx, const vanerf_target* tar, const int32_t* pix_xy, int32_t n_rays,
                       const float* ztab, int32_t n_samples, float* rays, float* z, void* stream);

/* Same with explicit interpolation parameters: ttab dev (S), or (R,S) when t_per_ray != 0.  The training branch draws
 * t = z_lower + rand * (z_upper - z_lower) per ray and sample (stratified jitter, src/model.py:1226-1230) with torch and
 * passes the table in; z = near + (far - near) * t like the uniform branch (:1230,:1232). */
int vanerf_sample_rays_t(vanerf_ctx* ctx, const vanerf_target* tar, const int32_t* pix_xy, int32_t n_rays,
                         const float* ttab, int32_t n_samples, int32_t t_per_ray, float* rays, float* z, void* stream);

/* Signed distance to the mesh, closest face, nearest vertex, per-view sample visibility
 * (mesh_util.py:498-524 = kaolin point_to_mesh_distance + check_sign; networks.py:28 = pytorch3d knn_points).
 * Outputs dev: pts (N,3) sample positions, sdf (N), face (N), nn_vert (N), qvis (V,N).  Any output may be NULL
 * except sdf/nn_vert/qvis when followed by vanerf_shade. */
int vanerf_geom_query(vanerf_ctx* ctx, const vanerf_target* tar, const float* rays, const float* z,
                      int32_t n_rays, int32_t n_samples, float* pts, float* sdf, int32_t* face,
                      int32_t* nn_vert, uint8_t* qvis, void* stream);

/* VANeRF.query + eval_func for N = n_rays*n_samples points (src/model.py:748-957, :1140-1160):
 * projection, masks, pix_weight, feature gather, SpatialEncoder, GeoVisFusion, MLPUNetFusion, TexVisFusion,
 * IBRRenderingHead.  rgba dev (N,5) = [valid*relu(o1), valid*o0 + (1-valid)*0.001, r, g, b]; valid dev (N) 0/1.
 * raw_out: optional dev (N,5) = VANeRF.query's own output [o0,o1,r,g,b]. */
int vanerf_shade(vanerf_ctx* ctx, int precision, const vanerf_target* tar, const float* rays, const float* z,
                 int32_t n_rays, int32_t n_samples, const float* sdf, const int32_t* nn_vert, const uint8_t* qvis,
                 float* rgba, uint8_t* valid, float* raw_out, void* stream);

/* VANeRF.rgba2out (src/model.py:1465-1494), warp per ray.  mesh_sdf dev (R,S) is the signed mesh distance.
 * out dev: color (R,3), depth (R), alpha (R), sdf_out (R), contrib (R,S); any may be NULL. */
int vanerf_composite(vanerf_ctx* ctx, const float* rgba, const float* z, const float* mesh_sdf, int32_t n_rays,
                     int32_t n_samples, float* color, float* depth, float* alpha, float* sdf_out, float* contrib,
                     void* stream);

/* VANeRF.importance_sample + merge sort (src/model.py:1301-1307, :1425-1462).  u dev (n_fine) shared by all rays
 * (linspace(0,1,n_fine) for uniform=True) or (R,n_fine) when u_per_ray != 0.  z_out dev (R, S + n_fine) sorted. */
int vanerf_importance(vanerf_ctx* ctx, const float* contrib, const float* z, int32_t n_rays, int32_t n_samples,
                      const float* u, int32_t n_fine, int32_t u_per_ray, float* z_fine_only, float* z_out,
                      void* stream);

/* Same sampler with the reference's own argument convention: contrib_inner dev (R, D-2) = contrib[..., 1:-1],
 * z_mid dev (R, D-1); z_fine dev (R, n_fine) out, no merge. */
int vanerf_importance_mid(vanerf_ctx* ctx, const float* contrib_inner, const float* z_mid, int32_t n_rays, int32_t n_depths,
                          const float* u, int32_t n_fine, int32_t u_per_ray, float* z_fine, void* stream);

/* One call for VANeRF.batch_render_pifu_nerf's ray batch (src/model.py:1103-1422, inference branch):
 * rays -> coarse pass -> composite -> importance -> fine pass -> composite, chunked internally.
 * ztab dev (n_coarse) = linspace(0,1,n_coarse), utab dev (n_fine) = linspace(0,1,n_fine) (uniform=True).
 * out_coarse / out_fine dev (R,8): r,g,b, depth, alpha, sdf, 0, 0 (out_fine may be NULL when fine == 0). */
int vanerf_render_rays(vanerf_ctx* ctx, int precision, const vanerf_target* tar, const int32_t* pix_xy,
                       int32_t n_rays, int32_t n_coarse, int32_t n_fine, int32_t fine, const float* ztab,
                       const float* utab, float* out_coarse, float* out_fine, void* stream);

/* Coarse reuse in vanerf_render_rays (default off = the reference's evaluation count).  The merged fine set of
 * src/model.py:1301-1307 contains the n_coarse coarse depths bit for bit and a sample's outputs depend only on its
 * point and ray, so the reference evaluates those samples twice (:1280-1294 and :1340-1349).  With reuse on, only the
 * n_fine new depths go through geometry / gather / networks in the fine pass and the merged arrays are assembled from
 * the two evaluations: identical output bits, n_coarse + n_fine instead of 2 n_coarse + n_fine evaluations per ray. */
int vanerf_set_reuse_coarse(vanerf_ctx* ctx, int on);

/* Geometry reuse in the fine pass of vanerf_render_rays (default ON).  The mesh queries of the reference
 * (cal_vis_sdf_batch mesh_util.py:498-524, knn_points networks.py:28) are functions of the sample position alone and the
 * merged fine set holds the n_coarse coarse depths bit for bit, so the fine pass queries the mesh for the n_fine new
 * depths only and takes the coarse pass's sdf / nearest vertex / sample visibility for the rest.  The NETWORKS still
 * evaluate all n_coarse + n_fine merged samples, as the reference does (src/model.py:1328-1349): same evaluation count,
 * identical output bits (tests: *_geometry_reuse_is_bit_identical).  0 = query the mesh again for every merged sample. */
int vanerf_set_reuse_geometry(vanerf_ctx* ctx, int on);

/* Scratch the context needs for vanerf_render_rays / vanerf_shade at the given sizes (bytes). */
size_t vanerf_scratch_bytes(const vanerf_ctx* ctx, int32_t n_rays, int32_t n_samples);

/* Test hook: vanerf_shade (fp32) that also returns MLPUNetFusion's pooled latent (N,128). */
int vanerf_shade_debug(vanerf_ctx* ctx, const vanerf_target* tar, const float* rays, const float* z, int32_t n_rays,
                       int32_t n_samples, const float* sdf, const int32_t* nn_vert, const uint8_t* qvis, float* rgba,
                       uint8_t* valid, float* raw_out, float* latent, void* stream);

/* VANeRF.query on explicit points (src/model.py:748-877): pts, view dev (N,3).  sdf_in / qvis_in: optional
 * caller-provided query_sdf (N) / query_vis (V,N) (the reference passes them in); NULL = computed from the mesh.
 * raw_out dev (N,5) = [o0,o1,r,g,b], valid dev (N), rgba dev (N,5) (eval_func layout); outputs may be NULL. */
int vanerf_query_points(vanerf_ctx* ctx, int precision, const vanerf_target* tar, const float* pts, const float* view,
                        int32_t n_points, const float* sdf_in, const uint8_t* qvis_in, float* raw_out, uint8_t* valid,
                        float* rgba, void* stream);

/* Stage-level primitives: what the reference's per-stage callables are made of.  vanerf_b200/stages.py composes them into
 * feat_sample, KNN_vis, SpatialEncoder, GeoVisFusion, MLPUNetFusion, TexVisFusion and IBRRenderingHead with the reference's
 * signatures; the render path itself runs the fused kernels and never materialises these outputs.
 *   vanerf_feat_sample  feat_sample (src/utils.py:136-151): feat dev (B,C,H,W), uv dev (B,N,2) -> out dev (B,N,C)
 *   vanerf_knn1         pytorch3d knn_points(K=1) inside KNN_vis (src/networks.py:28): query dev (N,3), vert dev (Nv,3) -> idx (N)
 *   vanerf_dense        one Conv1d(k=1) / Linear layer: y (M,N) = act(x (M,K) @ w (N,K)^T + b), act: 0 none, 1 ReLU,
 *                       2 Softplus(beta=100, threshold=20), 3 sigmoid, 4 ELU; b may be NULL
 *   vanerf_rel_z_decay  SpatialEncoder "rel_z_decay" (src/spatial.py:109-117): cxyz dev (BV,N,3) camera-space samples, kxyz dev
 *                       (BV,n_kpt,3) camera-space keypoints -> out dev (BV,N,(1 + 2 levels) n_kpt) */
int vanerf_feat_sample(vanerf_ctx* ctx, const float* feat, int32_t B, int32_t C, int32_t H, int32_t W, const float* uv, int32_t N,
                       float* out, void* stream);
int vanerf_knn1(vanerf_ctx* ctx, const float* query, int32_t N, const float* vert, int32_t Nv, int32_t* idx, void* stream);
int vanerf_dense(vanerf_ctx* ctx, const float* x, int32_t M, int32_t K, const float* w, const float* b, int32_t N, int32_t act,
                 float* y, void* stream);
int vanerf_rel_z_decay(vanerf_ctx* ctx, const float* cxyz, const float* kxyz, int32_t BV, int32_t N, int32_t n_kpt, int32_t levels,
                       float scale, float sigma, float* out, void* stream);

/* Training branch (BASELINE.json configs[4]; reference: the net.training paths of src/model.py:748-957, :1103-1422 and their
 * autograd).  Sampling and mesh queries reuse vanerf_sample_rays_t / vanerf_importance (per-ray u) / vanerf_geom_query (no
 * gradient flows through them).  The unfused training graph of vanerf_b200/train.py adds:
 *   vanerf_project_samples  per sample and view, what VANeRF.query derives from the position alone (src/model.py:780-821,
 *                           :936-946, src/spatial.py:71-72): xy dev (V,N,2), mask dev (N) = all-views AND of the in-frustum and
 *                           foreground tests (before view dropout), pw_raw dev (V,N) boundary weight before mask / normalisation,
 *                           cam dev (V,N,3) camera-space position, ray_diff dev (V,N,4)
 *   vanerf_feat_sample_bwd  backward of feat_sample w.r.t. the map: d_out dev (B,N,C) scatter-added into d_feat dev (B,C,H,W)
 *                           (zeroed by the caller)
 *   vanerf_composite_beta   vanerf_composite with an explicit beta (sigmoid_beta is a parameter of the training graph)
 *   vanerf_composite_bwd    backward of rgba2out: g_color (R,3), g_alpha, g_depth, g_sdf (R) (any NULL) -> d_rgba dev (R,S,5) and
 *                           d_beta dev (1) (accumulated; zeroed by the caller)
 * Dense layers of the training graph are library GEMMs (cuBLAS through torch.nn.functional.linear) under torch autograd. */
int vanerf_project_samples(vanerf_ctx* ctx, const vanerf_target* tar, const float* rays, const float* z, int32_t n_rays, int32_t n_samples,
                           float* xy, uint8_t* mask, float* pw_raw, float* cam, float* ray_diff, void* stream);
int vanerf_feat_sample_bwd(vanerf_ctx* ctx, const float* d_out, int32_t B, int32_t C, int32_t H, int32_t W, const float* uv, int32_t N,
                           float* d_feat, void* stream);
int vanerf_composite_beta(vanerf_ctx* ctx, const float* rgba, const float* z, const float* mesh_sdf, int32_t n_rays, int32_t n_samples, float beta,
                          float* color, float* depth, float* alpha, float* sdf_out, float* contrib, void* stream);
int vanerf_composite_bwd(vanerf_ctx* ctx, const float* rgba, const float* z, const float* mesh_sdf, int32_t n_rays, int32_t n_samples, float beta,
                         const float* g_color, const float* g_alpha, const float* g_depth, const float* g_sdf, float* d_rgba, float* d_beta,
                         void* stream);

/* Per-kernel-class device timing with CUDA events on the launching stream (used by bench.py for the roofline
 * numbers).  vanerf_timing_read synchronises the recorded events; ms_out / count_out have 7 entries:
 * frame setup, rays, geometry, gather, mlp, composite, importance. */
int vanerf_timing_enable(vanerf_ctx* ctx, int on);
int vanerf_timing_read(vanerf_ctx* ctx, double* ms_out, int64_t* count_out, int reset);

/* Test hooks of the bf16 tensor-core path (tcgen05 / TMEM; precision == VANERF_BF16 in the calls above).
 * vanerf_shade_debug_bf16: vanerf_shade_debug on the tensor-core kernel (latent before bf16 rounding).
 * vanerf_tc_error: nonzero when a bounded barrier wait inside a tensor-core kernel gave up (read after a stream
 *   synchronise; such a launch produced garbage instead of hanging).
 * vanerf_tc_selftest: D dev (128, (N+15)&~15) = bf16(A dev (128,K)) x bf16(W host (N,K))^T through one tcgen05 step
 *   with the same operand layout, weight ring and TMEM read-back as the shading kernel; K % 16 == 0, K <= 256, N <= 128. */
int vanerf_shade_debug_bf16(vanerf_ctx* ctx, const vanerf_target* tar, const float* rays, const float* z, int32_t n_rays,
                            int32_t n_samples, const float* sdf, const int32_t* nn_vert, const uint8_t* qvis, float* rgba,
                            uint8_t* valid, float* raw_out, float* latent, void* stream);
int vanerf_tc_error(vanerf_ctx* ctx);
/* Completion check of the bf16 path (a SYNCHRONISING call): waits for `stream`, then returns VANERF_ERR_CUDA and clears the
 * record if a tensor-core launch since the last check gave up on a bounded wait (its results are invalid), VANERF_OK
 * otherwise.  Without it the condition is reported by the next bf16 call on the context. */
int vanerf_tc_check(vanerf_ctx* ctx, void* stream);
/* Developer aid: cycle trace (tag, clock64) pairs of CTA 0 / thread 0 of the following tensor-core launches into
 * buf dev (capacity, 2) int64; buf == NULL switches the trace off and returns the number of pairs recorded. */
int vanerf_tc_profile(vanerf_ctx* ctx, long long* buf, int32_t capacity);
int vanerf_tc_selftest(vanerf_ctx* ctx, const float* A_dev, const float* W_host, int32_t K, int32_t N, float* D_dev, void* stream);
/* 0 when the weight-packing script and the compile-time MMA program of the tensor-core kernels agree (CPU only). */
int vanerf_tc_program_check(void);
/* Developer measurement: cycles to issue / to complete `reps` tcgen05.mma of shape 128 x n x 16 round-robin over n_acc
 * accumulators (mode bit 0: two issuing warps per CTA); out_host (2, 2) int64 = per warp [issue, total] of CTA 0. */
int vanerf_tc_mma_probe(vanerf_ctx* ctx, int32_t n, int32_t reps, int32_t n_acc, int32_t mode, int32_t n_ctas, long long* out_host);

/* Number of kernels launched by this context since creation (for bench.py's gpu_launches). */
int64_t vanerf_launch_count(const vanerf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VANERF_B200_H */
