/*
 * anerf_b200 -- C ABI of the B200-native A-NeRF ray-marching hot path.
 *
 * Everything here takes plain pointers and sizes; device pointers are CUDA device addresses on the
 * current device, `stream` is a cudaStream_t passed as void*.  No torch types cross this boundary.
 * Every function returns 0 on success or a negative anerf_status; anerf_last_error() gives the
 * message for the calling thread.  Nothing in this library has a CPU fallback.
 *
 * Reference interfaces each entry point stands in for (LemonATsu/A-NeRF, paths relative to the
 * reference root):
 *   anerf_plan_create / anerf_pack_net  <- NeRF.__init__ + RayCaster.load_state_dict
 *                                          (core/networks/nerf.py:12-88, core/raycasters.py:768-788)
 *   anerf_render_fwd                    <- RayCaster.render_rays  (core/raycasters.py:361-474) incl.
 *                                          get_near_far_in_cylinder (core/utils/ray_utils.py:292-344),
 *                                          sample_from_lineseg (:204-251), isample_from_lineseg (:255-289),
 *                                          encode_inputs (core/raycasters.py:476-555), run_network (:557-577),
 *                                          NeRF.forward + raw2outputs (core/networks/nerf.py:133-205)
 *   anerf_render_fwd_host               <- the same call as core/trainer.py:64-79 batchify_rays makes it,
 *                                          with host buffers (H2D/D2H inside)
 *   anerf_render_frame                  <- get_rays (core/utils/ray_utils.py:6-28) + the per-ray expansion of the pose in
 *                                          run_nerf.render_path (run_nerf.py:77-98) + render_rays, per frame
 *   anerf_density_points                <- RayCaster.render_pts_density / render_mesh_density
 *                                          (core/raycasters.py:579-648)
 *   anerf_render_bwd                    <- what loss.backward() does to the graph of render_rays in training
 *                                          (core/trainer.py:187-203 optimize -> autograd through
 *                                          core/networks/nerf.py:94-205, core/cutoff_embedder.py:111-174,
 *                                          core/encoders.py:8-37, core/networks/embedding.py:17-34):
 *                                          gradients of the network weights, the framecodes and the per-ray
 *                                          bone transforms (pose refinement, core/pose_opt.py:435)
 */
#ifndef ANERF_B200_H_
#define ANERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ANERF_OK = 0,
  ANERF_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
  ANERF_ERR_CUDA = -2,         /* a CUDA runtime call failed */
  ANERF_ERR_DEVICE = -3,       /* the kernel reported a protocol error (see anerf_last_error) */
  ANERF_ERR_NOMEM = -4
} anerf_status;

/* Shape of one density/radiance MLP and of its encodings (create_raycaster flags,
 * core/raycasters.py:17-104).  multires is fixed at 7 (all shipped configs); multires_views is 4 or 0
 * (configs/surreal/surreal_single.txt:32: raw bone-local view directions, no sin/cos features). */
typedef struct {
  int32_t n_joints;        /* 1..24 */
  int32_t depth;           /* netdepth D (pts_linears), 2..8 */
  int32_t width;           /* netwidth W: 64, 128 or 256 */
  int32_t skip;            /* reference skips=[4]: layer skip+1 takes cat[encoding, h]; -1 if >= depth-1 */
  int32_t framecode_ch;    /* 0, or 16 when opt_framecode */
  int32_t n_framecodes;    /* rows of the framecode table */
  int32_t operand_format;  /* 0 = fp16 hi/lo split (what RayCaster uses by default), 1 = bf16 hi/lo split */
  int32_t view_freqs;      /* multires_views: 4 or 0 */
} anerf_net_config;

/* Pointers to one network's fp32 parameters on the device, reference state_dict layout
 * ([out, in] row-major weights). */
typedef struct {
  const float* pts_w[8];   /* pts_linears.i.weight */
  const float* pts_b[8];
  const float* alpha_w;    /* [1, W] */
  const float* alpha_b;    /* [1] */
  const float* feature_w;  /* [W, W] */
  const float* feature_b;
  const float* views_w;    /* [W/2, W + 3*(1 + 2*view_freqs)*J (+ framecode_ch)] */
  const float* views_b;
  const float* rgb_w;      /* [3, W/2] */
  const float* rgb_b;
  const float* framecodes; /* [n_framecodes, framecode_ch] or NULL */
} anerf_net_params;

typedef struct anerf_plan anerf_plan;   /* opaque: layer program + K maps for one anerf_net_config */

int anerf_plan_create(const anerf_net_config* cfg, anerf_plan** out);
void anerf_plan_destroy(anerf_plan* plan);
/* Bytes of the packed (tensor-core operand layout) image of one network. */
size_t anerf_packed_bytes(const anerf_plan* plan);
/* fp32 parameters -> packed image (device to device, asynchronous on `stream`).  Call after every
 * parameter update (load_state_dict, optimizer step). */
int anerf_pack_net(const anerf_plan* plan, const anerf_net_params* params, void* packed, void* stream);

/* Per-call options of render_rays (render_kwargs of core/raycasters.py:156-178). */
typedef struct {
  int32_t n_rays;
  int32_t n_samples;         /* coarse samples per ray (N_samples) */
  int32_t n_importance;      /* extra fine samples (N_importance), 0 = coarse pass only */
  int32_t lindisp;
  int32_t softplus;          /* density_type: 0 relu, 1 softplus */
  int32_t eval_mean_framecode; /* eval with all cams < 0: use the mean code (embedding.py:21-22) */
  float density_scale;       /* B */
  float softplus_shift;
  float tau_pts, tau_views;  /* CutoffEmbedder.tau of embed_fn / embeddirs_fn */
  float cutoff_pts[24];      /* CutoffEmbedder.cutoff_dist per joint */
  float cutoff_views[24];
  int32_t single_net;        /* --single_net: one network for both passes (pass the same packed image twice) and the
                                blurred importance pdf of core/utils/ray_utils.py:271-277 */
  int32_t reserved;
} anerf_render_opts;

/* Device inputs.  rays: [N,8] = origin(3), direction(3), near, far (the first 8 columns of the
 * reference's ray_batch; viewdirs are unused by the path, core/raycasters.py:413-417).
 * skts: [N,J,4,4] world->bone transforms per ray.  cyls: [N,5].  cams: [N] float camera indices or NULL.
 * Optional random draws (training): t_rand [N,Sc], u_rand [N,Si], noise0 [N,Sc], noise1 [N,Sc+Si]
 * (already multiplied by raw_noise_std); NULL = deterministic sampling / no noise. */
typedef struct {
  const float* rays;
  const float* skts;
  const float* cyls;
  const float* cams;
  const float* t_rand;
  const float* u_rand;
  const float* noise0;
  const float* noise1;
  const int32_t* pose_idx;  /* optional [N]: ray -> pose.  When set, `skts` is [P,J,4,4] (one transform set per POSE, read
                               through the index) and anerf_render_bwd's g_skts is [P,J,4,4], summed over each pose's rays */
  int32_t n_poses;          /* P (rows of `skts`) when pose_idx is set: indices are clamped to [0, P) on the device, so a bad
                               index can never read or write outside the buffers; 0 = unchecked */
  int32_t reserved;
} anerf_render_inputs;

/* Device outputs ([N,...], fp32).  The *0 entries and z_all may be NULL; with n_importance == 0 the
 * coarse results go to rgb_map/disp_map/acc_map/alpha. */
typedef struct {
  float* rgb_map;   /* [N,3] */
  float* disp_map;  /* [N] */
  float* acc_map;   /* [N] */
  float* alpha;     /* [N, Sc+Si] */
  float* rgb0;      /* [N,3] */
  float* disp0;
  float* acc0;
  float* alpha0;    /* [N,Sc] */
  float* z_all;     /* optional tap: sorted depths of the fine pass [N,Sc+Si] */
  float* raw;       /* optional tap: network outputs of the last pass [N,S,4] */
} anerf_render_outputs;

size_t anerf_render_workspace_bytes(int32_t n_rays);

/* One chunk of rays through the whole path on the device.  packed_fine may equal packed_coarse
 * (opts->single_net) and is ignored when n_importance == 0. */
int anerf_render_fwd(const anerf_plan* plan, const void* packed_coarse, const void* packed_fine,
                     const anerf_render_opts* opts, const anerf_render_inputs* in,
                     const anerf_render_outputs* out, void* workspace, size_t workspace_bytes, void* stream);

/* One frame's camera and pose for anerf_render_frame: what run_nerf.render_path hands to render() per frame
 * (run_nerf.py:77-98) before get_rays (core/utils/ray_utils.py:6-28) and the per-ray replication of the pose
 * (run_nerf.py:84-90) -- here the rays are generated inside the kernels and the pose is read once. */
typedef struct {
  float c2w[12];           /* rows 0..2 of the camera-to-world matrix, row-major [3][4] */
  float focal_x, focal_y;
  float center_x, center_y;/* principal point; get_rays' default is (W/2, H/2) */
  float near, far;         /* ray bounds before the cylinder intersection (the reference passes 0 and 1) */
  int32_t width, height;
  int32_t pixel0;          /* ray r of the call is pixel pixel0 + r (row-major j*W + i) when `pixels` is NULL */
  const int32_t* pixels;   /* device, optional: [n_rays] flat pixel indices (e.g. the pixels inside the skeleton's box) */
  const float* skts;       /* device: [J,4,4], ONE pose for the frame */
  const float* cyl;        /* device: [5] bounding cylinder of the frame */
  float cam;               /* the frame's camera index (framecodes); ignored otherwise */
  int32_t reserved;
} anerf_frame_inputs;

/* A chunk of n_rays pixels of one frame: like anerf_render_fwd on the rays get_rays would produce for those pixels with
 * the pose replicated per ray, without materialising either (no random draws: rendering is deterministic). */
int anerf_render_frame(const anerf_plan* plan, const void* packed_coarse, const void* packed_fine,
                       const anerf_render_opts* opts, const anerf_frame_inputs* frame, const anerf_render_outputs* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Same, with HOST buffers for inputs and outputs (pinned or pageable); copies in and out on
 * `stream` around the kernels and synchronises the stream before returning. */
int anerf_render_fwd_host(const anerf_plan* plan, const void* packed_coarse, const void* packed_fine,
                          const anerf_render_opts* opts, const anerf_render_inputs* host_in,
                          const anerf_render_outputs* host_out, void* stream);

/* What core/trainer.py:64-79 batchify_rays does with a whole frame of HOST buffers: opts->n_rays rays (any number) are
 * processed in chunks of `chunk` rays (each chunk exactly as one anerf_render_fwd call, so the near/far repair stays a
 * chunk-wide mean), with the copies of chunk c+1 (host -> device) and of chunk c-1 (device -> host) overlapping the
 * kernels of chunk c: three internal streams and two device arenas that are kept across calls.  Pinned host memory
 * gives real overlap; pageable memory works but serialises.  Outputs that are NULL are neither computed into host memory
 * nor copied (render_path reads rgb/disp/acc only).  Synchronises before returning. */
int anerf_render_fwd_host_chunked(const anerf_plan* plan, const void* packed_coarse, const void* packed_fine,
                                  const anerf_render_opts* opts, int32_t chunk, const anerf_render_inputs* host_in,
                                  const anerf_render_outputs* host_out);

/* ---- training: backward of one chunk ------------------------------------------------------------------ */

/* Gradient buffers of one network, same shapes as anerf_net_params, fp32 on the device.  Gradients are ADDED to
 * the buffers (zero-fill them for a fresh gradient); a NULL entry skips that parameter (frozen layer). */
typedef struct {
  float* pts_w[8];
  float* pts_b[8];
  float* alpha_w;
  float* alpha_b;
  float* feature_w;
  float* feature_b;
  float* views_w;
  float* views_b;
  float* rgb_w;
  float* rgb_b;
  float* framecodes;
} anerf_net_grads;

/* dL/d(outputs of anerf_render_fwd), device, same shapes as anerf_render_outputs; NULL = zero. */
typedef struct {
  const float* rgb_map;
  const float* disp_map;
  const float* acc_map;
  const float* alpha;
  const float* rgb0;
  const float* disp0;
  const float* acc0;
  const float* alpha0;
} anerf_render_grads;

size_t anerf_render_bwd_workspace_bytes(const anerf_plan* plan, int32_t n_rays, int32_t n_samples, int32_t n_importance);

/* Backward of anerf_render_fwd for the same opts / inputs.  `nearfar` [N,2] is the forward call's workspace
 * (the repaired near/far of every ray), `z_all` [N,Sc+Si] its z_all tap (ignored when n_importance == 0); the
 * sample positions carry no gradient (the reference detaches them, core/utils/ray_utils.py:285).  The pass
 * recomputes the activations of both networks layer by layer in fp32 from the fp32 parameters (`coarse`,
 * `fine`: the same pointers that were packed), so no activations have to be kept from the forward call.
 * g_skts [N,J,4,4] (added to; NULL = the pose needs no gradient). */
int anerf_render_bwd(const anerf_plan* plan, const anerf_net_params* coarse, const anerf_net_params* fine,
                     const anerf_render_opts* opts, const anerf_render_inputs* in, const float* nearfar,
                     const float* z_all, const anerf_render_grads* grad_out, const anerf_net_grads* g_coarse,
                     const anerf_net_grads* g_fine, float* g_skts, void* workspace, size_t workspace_bytes,
                     void* stream);

/* The same, one network pass at a time: pass_mask 1 = the coarse network's pass, 2 = the fine network's, 3 = both.  Lets a
 * trainer exchange the coarse network's gradients (NCCL, another stream) while the fine pass still runs. */
int anerf_render_bwd_pass(const anerf_plan* plan, const anerf_net_params* coarse, const anerf_net_params* fine,
                          const anerf_render_opts* opts, const anerf_render_inputs* in, const float* nearfar,
                          const float* z_all, const anerf_render_grads* grad_out, const anerf_net_grads* g_coarse,
                          const anerf_net_grads* g_fine, float* g_skts, void* workspace, size_t workspace_bytes,
                          int32_t pass_mask, void* stream);

/* Training step without the recomputation: the forward of a training step run as the layer-wise GEMM chain of the
 * backward pass (same fp32 parameters, same tensor-core engine), KEEPING the activations of both network passes in a
 * caller-owned `state` buffer, and a backward that starts from that buffer -- three GEMM passes per step instead of the
 * four of anerf_render_fwd + anerf_render_bwd (the reference: torch autograd keeps every activation as well,
 * core/networks/nerf.py:94-148).  Outputs and their meaning are those of anerf_render_fwd (core/raycasters.py:711-724;
 * `out->raw` is not provided); the per-ray stages (compositing, importance sampling, sorted merge) are the fused kernel's
 * own code.
 * anerf_train_state_bytes: size of `state` for a batch, or 0 when the batch is too large to keep resident (more than
 * 524,288 samples in one network pass) -- use anerf_render_fwd + anerf_render_bwd then.
 * anerf_render_fwd_train: `nearfar_out` [N,2] / out->z_all [N,Sc+Si] (optional) receive copies of what the state holds,
 * so that a caller can still fall back to anerf_render_bwd.  The state stays valid for anerf_render_bwd_saved until the next
 * anerf_render_fwd_train on the same buffer; both calls must see the same ANERF_TRAIN_GEMM setting. */
size_t anerf_train_state_bytes(const anerf_plan* plan, int32_t n_rays, int32_t n_samples, int32_t n_importance);
int anerf_render_fwd_train(const anerf_plan* plan, const anerf_net_params* coarse, const anerf_net_params* fine,
                           const anerf_render_opts* opts, const anerf_render_inputs* in, const anerf_render_outputs* out,
                           float* nearfar_out, void* state, size_t state_bytes, void* stream);
int anerf_render_bwd_saved(const anerf_plan* plan, const anerf_net_params* coarse, const anerf_net_params* fine,
                           const anerf_render_opts* opts, const anerf_render_inputs* in, const anerf_render_grads* grad_out,
                           const anerf_net_grads* g_coarse, const anerf_net_grads* g_fine, float* g_skts, void* state,
                           size_t state_bytes, int32_t pass_mask, void* stream);

/* Raw (pre-activation) density of `n_points` world points under ONE pose: pts [P,3], skts [J,4,4],
 * sigma [P].  Uses tau_pts / cutoff_pts of `opts` (other fields ignored). */
int anerf_density_points(const anerf_plan* plan, const void* packed, const anerf_render_opts* opts,
                         const float* pts, const float* skts, int64_t n_points, float* sigma, void* stream);

/* The same for voxels [first, first + count) of the flattened (res+1)^3 grid that RayCaster.render_mesh_density builds
 * (core/raycasters.py:579-595: np.meshgrid(t, t, t) with t = np.linspace(-radius, radius, res+1), plus kps[0,0]), with
 * the points generated inside the kernel: nothing but the densities touches HBM.  origin: DEVICE pointer to the root
 * joint position (3 floats); sigma [count] in flat grid order. */
int anerf_density_grid(const anerf_plan* plan, const void* packed, const anerf_render_opts* opts, const float* origin,
                       double radius, int32_t res, int64_t first, int64_t count, const float* skts, float* sigma,
                       void* stream);

/* Build-time self test of the tensor-core building blocks on one CTA pair: D[256,N] = A[256,K] * B[N,K]^T with
 * the split-precision operand path (A, B fp32 on the device, D [2,256,N] fp32: two passes; K multiple of 128;
 * N in {64,128,256}).
 * format: 1 bf16, 0 fp16. */
int anerf_selftest_gemm(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t format,
                        void* stream);

/* Self test of the training path's tensor-core GEMM (tc_gemm.cuh) with explicit strides:
 *   C(m,n) (op)= sum_k A(m,k) B(n,k),  A(m,k) = A[m*a_ms + k*a_ks], B(n,k) = B[n*b_ns + k*b_ks], C(m,n) = C[m*c_ms + n*c_ns]
 * epilogue: + bias[n], ReLU, zero where mask[m*mask_ms + n] <= 0; mode 0 store, 1 add to C, 2 atomic add;
 * slice_chunks > 0 splits K into slices of slice_chunks*32 (use with mode 2).  bf16 hi/lo split, 3 MMAs per product. */
int anerf_selftest_tc_gemm(const float* A, int64_t a_ms, int64_t a_ks, int32_t M, int32_t K, const float* B, int64_t b_ns,
                           int64_t b_ks, int32_t N, float* C, int64_t c_ms, int64_t c_ns, const float* bias,
                           const float* mask, int64_t mask_ms, int32_t relu, int32_t mode, int32_t slice_chunks,
                           void* stream);

/* ---- pose refinement: the kinematic chain ------------------------------------------ */

/* PoseOptLayer.calculate_kinematic (core/pose_opt.py:372-445, with unrolled_kinematic_chain :482-521 and torch.inverse
 * :435) for n_poses poses: rots [P,J,3,3] per-joint rotations (what rot6d_to_rotmat / axisang_to_rot produce), rest_pose
 * [n_rest,J,3] with n_rest in {1, P}, pelvis [P,3]  ->  l2ws [P,J,4,4] (pelvis-shifted, as the reference returns them),
 * skts [P,J,4,4] = l2ws^-1 (closed-form rigid inverse), kps [P,J,3] (may be NULL).  parents: HOST array [J], the
 * skeleton's joint_trees (parents[root_id] == root_id; every other parent precedes its child).  All tensors on the device. */
int anerf_pose_chain_fwd(int32_t n_poses, int32_t n_joints, const int32_t* parents, int32_t root_id, const float* rots,
                         const float* rest_pose, int32_t n_rest, const float* pelvis, float* l2ws, float* skts, float* kps,
                         void* stream);
size_t anerf_pose_chain_bwd_scratch_bytes(int32_t n_poses, int32_t n_joints);
/* Its backward: cotangents of skts / l2ws / kps (each [P,...] or NULL) -> g_rots [P,J,3,3], g_pelvis [P,3] (overwritten).
 * g_skts is what anerf_render_bwd accumulates per POSE when it is given `pose_idx` (the segment sum over the rays of
 * each pose happens there, with atomics), so the [N,J,4,4] per-ray gradient of the reference's graph never exists. */
int anerf_pose_chain_bwd(int32_t n_poses, int32_t n_joints, const int32_t* parents, int32_t root_id, const float* rots,
                         const float* rest_pose, int32_t n_rest, const float* pelvis, const float* l2ws, const float* skts,
                         const float* g_skts, const float* g_l2ws, const float* g_kps, float* g_rots, float* g_pelvis,
                         void* scratch, size_t scratch_bytes, void* stream);

/* ---- mesh extraction ------------------------------------------------------------------ */

/* Marching cubes on a density volume resident on the device (reference: mcubes.marching_cubes(sigma, threshold) on a
 * host copy, run_render.py:983-986).  volume: n0 x n1 x n2 fp32 voxels with element strides s0, s1, s2 (so the
 * transposed view RayCaster.render_mesh_density returns needs no copy).  Pass 1 writes the number of triangles of each
 * of the (n0-1)(n1-1)(n2-1) cells (cell index = (i (n1-1) + j)(n2-1) + k); the caller scans them; pass 2 writes, for
 * triangle t and corner v, the vertex position in index coordinates (verts [T,3,3]) and the id of the volume edge it
 * lies on (keys [T,3]; equal ids = the same vertex, for welding).  Inside = value > iso; normals point to lower values. */
int anerf_mc_count(const float* volume, int32_t n0, int32_t n1, int32_t n2, int64_t s0, int64_t s1, int64_t s2, float iso,
                   int32_t* counts, void* stream);
int anerf_mc_emit(const float* volume, int32_t n0, int32_t n1, int32_t n2, int64_t s0, int64_t s1, int64_t s2, float iso,
                  const int64_t* offsets, float* verts, int64_t* keys, void* stream);

/* ---- training-ray sampler ----------------------------------------------------------- */

/* The dataset tensors of BaseH5Dataset (core/dataset.py:108-160 init_meta / the .h5 layout), resident on the device. */
typedef struct {
  const uint8_t* masks;    /* [F, H*W] sampling masks: a pixel may be drawn when > 0 ('sampling_masks') */
  const uint8_t* imgs;     /* [F, H*W, 3] ('imgs') */
  const uint8_t* fgs;      /* [F, H*W] foreground masks ('masks') or NULL */
  const uint8_t* bgs;      /* [B, H*W, 3] backgrounds or NULL */
  const int32_t* bg_idx;   /* [F] background of each image or NULL */
  const float* c2ws;       /* [F, 3, 4] */
  const float* focals;     /* [F, 2] (fx, fy) */
  const float* centers;    /* [F, 2] principal points or NULL (image centre) */
  int32_t height, width;
  int32_t n_frames;        /* F: entries of `frames` outside [0, F) make their image report -1 valid pixels and produce no rows */
  int32_t fg_is_255;       /* foreground masks stored as 0/255 instead of 0/1 */
  int32_t mask_img;        /* target = img * fg + (1 - fg) * bg (the reference's mask_img) */
} anerf_sampler_inputs;

typedef struct {
  float* rays;             /* [N, 8] origin, direction, near = 0, far = 1   (N = n_images * rays_per_image) */
  float* target;           /* [N, 3] colours in [0, 1] */
  float* fg;               /* [N] or NULL */
  float* bg;               /* [N, 3] or NULL */
  int32_t* pixel_idx;      /* [N] flat pixel index (row-major), increasing within an image */
  int32_t* frame_of_ray;   /* [N] image index of every ray (-> camera index / pose index) */
} anerf_sampler_outputs;

/* BaseH5Dataset.__getitem__ for a batch of images (core/dataset.py:57-105: sample_pixels :277-322 with patch_size 1 and
 * no box sampling, get_rays :346-362, get_img_data :258-275) on the device: for each of the n_images images listed in
 * `frames` (DEVICE int32 array), rays_per_image DISTINCT pixels drawn uniformly from its sampling mask, in increasing
 * pixel order.  n_valid (DEVICE [n_images]) receives each image's number of valid pixels; images with fewer than
 * rays_per_image valid pixels produce no rows (the caller checks).  The random stream is a counter-based hash of
 * (seed, image, pixel), not numpy's generator. */
int anerf_sample_rays(const anerf_sampler_inputs* in, const int32_t* frames, int32_t n_images, int32_t rays_per_image,
                      uint64_t seed, const anerf_sampler_outputs* out, int32_t* n_valid, void* stream);

/* ---- optimizer step ------------------------------------------------------------------- */

/* Photometric loss of the reference's trainer and its gradient seed in one launch (Trainer._compute_nerf_loss,
 * core/trainer.py:352-381 with img2l1 / img2mse :9-41): pred = rgb + (1 - acc) bg when use_background (bg [N,3] or, when
 * NULL, bg_const), loss = weight * mean(|pred - target|) (mse = 0) or weight * mean((pred - target)^2) (mse = 1).
 * Writes d loss / d rgb [N,3] and d loss / d acc [N]; ADDS the un-normalised sum of loss terms and the sum of squared
 * errors (for the PSNR) to sums[0..1]. */
int anerf_loss_seed(const float* rgb, const float* acc, const float* target, const float* bg, float bg_const,
                    int32_t use_background, int32_t mse, int32_t n_rays, float weight, float* g_rgb, float* g_acc,
                    float* sums, void* stream);

/* torch.optim.Adam.step (amsgrad off) for n_tensors fp32 tensors in ONE launch (reference: Trainer.optimize,
 * core/trainer.py:451-483; optimizer built at core/raycasters.py:116).  params / grads / exp_avg / exp_avg_sq: HOST
 * arrays of device pointers, sizes[i] elements each; `step` = the update count including this one; grads are first
 * multiplied by grad_scale (1/world after a summing all-reduce).  Hyper-parameters are doubles (Python floats): 1 - beta is
 * formed in double like torch does before it reaches the fp32 kernel. */
int anerf_adam_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                    float* const* exp_avg_sq, const int64_t* sizes, int64_t step, double lr, double beta1, double beta2, double eps,
                    double weight_decay, double grad_scale, void* stream);

/* Debug aid: device buffer of 3 x 1024 int64; while set, launches record a clock64 timeline of CTA 0
 * (stream 0 MMA thread, 1/2 worker groups; entries = tag << 48 | clock).  NULL switches it off. */
void anerf_debug_set_trace(long long* device_buffer);

const char* anerf_last_error(void);
int anerf_version(void);
/* The kernels never spin forever: a protocol error is recorded in a pinned status word and the kernel traps.  The
 * asynchronous entry points cannot see that; call this after synchronising the stream (0 = clean, ANERF_ERR_DEVICE
 * + anerf_last_error() otherwise).  [the reference drops into pdb on NaNs, core/utils/ray_utils.py:247] */
int anerf_check_status(void);

#ifdef __cplusplus
}
#endif
#endif  /* ANERF_B200_H_ */
