/*
 * lidar4d_b200.h - C-ABI of the B200 (sm_100a) LiDAR4D hot-path library
 * (liblidar4d_b200.so).  Plain C: pointers, sizes, POD structs; no torch / ATen
 * types cross this boundary.
 *
 * What it replaces (paths relative to the reference checkout):
 *   - the tiny-cuda-nn plugin surface used by the hot path
 *       tcnn.Encoding(HashGrid)   model/hash_field.py:47-57, :107-117; model/flow_field.py:67-77
 *       tcnn.Encoding(Frequency)  model/lidar4d.py:68-74
 *       tcnn.Network(FullyFusedMLP) model/lidar4d.py:83-117
 *   - the PyTorch op graph around it
 *       LiDAR_Renderer.run        model/renderer.py:44-140   (sampling, compositing)
 *       LiDAR4D.density/attribute model/lidar4d.py:139-223
 *       LiDAR4D.flow              model/lidar4d.py:124-137
 *       Planes4D / F.grid_sample  model/planes_field.py:56-141
 *       HashGrid4D / HashGridT    model/hash_field.py:65-88, :141-172
 *       FlowField                 model/flow_field.py:102-130
 *   - error convention: unlike the reference's chamfer plugin (utils/chamfer3D/chamfer3D.cu:144-150,
 *     printf + ignored return), every entry point returns 0 on success and a negative L4D_E* code
 *     on failure; l4d_last_error() gives the message; nothing throws or aborts across the ABI.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless named host_*; the caller owns every buffer
 *     (inputs, outputs, workspaces, gradients); the library allocates nothing on the hot path.
 *   - all work is enqueued on the caller's stream (`stream` = cudaStream_t cast to void*); no
 *     device synchronisation inside; re-entrant (no global state except the thread-local error string).
 *   - host-side decisions (frame index, first/last-frame branches, time-slice indices, Lagrange
 *     bases) arrive pre-computed in L4DFrame so that no device->host sync is needed
 *     (the reference syncs at lidar4d.py:143, hash_field.py:82, lidar4d.py:201).
 *   - level geometry (scale/resolution/entries) arrives as data in L4DGrid so that hash
 *     indices are bit-exact with whatever computed that geometry.
 */
#ifndef LIDAR4D_B200_H_
#define LIDAR4D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L4D_ABI_VERSION 1
#define L4D_MAX_LEVELS 16
#define L4D_MAX_PLANE_SCALES 4
#define L4D_MAX_TIME_SLICES 16
#define L4D_HIDDEN 64

enum {
  L4D_OK = 0,
  L4D_EINVAL = -1,   /* bad argument / unsupported configuration */
  L4D_ECUDA = -2,    /* a CUDA runtime call or launch failed      */
  L4D_ESIZE = -3     /* caller-provided buffer too small          */
};

/* One multi-resolution hash grid (one tcnn HashGrid encoding). */
typedef struct L4DGrid {
  uint32_t n_dims;        /* 2 or 3 */
  uint32_t n_levels;      /* <= L4D_MAX_LEVELS */
  uint32_t n_features;    /* 4 (hash_static / hash_dynamic) or 8 (flow) */
  uint32_t reserved;
  float    scale[L4D_MAX_LEVELS];
  uint32_t resolution[L4D_MAX_LEVELS];
  uint32_t entries[L4D_MAX_LEVELS];       /* hashmap_size of the level */
  uint32_t offset[L4D_MAX_LEVELS + 1];    /* first entry of the level; offset[n_levels] = total entries */
} L4DGrid;

/* Model hyper-parameters (LiDAR4D.__init__, model/lidar4d.py:23-45). */
typedef struct L4DConfig {
  L4DGrid  hash_static;          /* 3D, F=4 */
  L4DGrid  hash_dynamic[3];      /* 2D xy / xz / yz, F=4; one table per time slice */
  L4DGrid  flow;                 /* 3D, F=8 */
  uint32_t n_plane_scales;       /* 4 */
  uint32_t plane_res[L4D_MAX_PLANE_SCALES];  /* spatial resolution per scale (32,64,128,256) */
  uint32_t time_resolution;      /* 8: time planes' H and number of hash time slices */
  uint32_t num_frames;
  uint32_t active_sensor;        /* renderer.py:101 */
  uint32_t view_degree;          /* 12 */
  uint32_t sigma_in_dim;         /* 120 */
  uint32_t sigma_in_pad;         /* 128 */
  uint32_t attr_in_dim;          /* 87 */
  uint32_t attr_in_pad;          /* 96 */
  float    bound;
  float    near_lidar;
  float    far_lidar;
  float    density_scale;
  uint32_t mlp_fp16;             /* 1: MLP weights are consumed as fp16-rounded working copies (as tcnn's half
                                    params) and the dense kernels run on tcgen05 tensor cores; 0: fp32 FMA */
  uint32_t reserved;
} L4DConfig;

/* One (.,tau) query of the time-sliced grids (model/hash_field.py:76-88). */
typedef struct L4DTimeQuery {
  float    tau;
  uint32_t slice_lo, slice_hi;
  float    w_lo, w_hi;           /* (idx2-idx), (idx-idx1) */
  uint32_t single;               /* idx1==idx2: feature = G_lo only */
  float    basis[4];             /* cubic Lagrange basis at tau (hash_field.py:65-74) */
} L4DTimeQuery;

/* Per-frame constants decided on the host (model/lidar4d.py:143,157-173). */
typedef struct L4DFrame {
  float        time;
  uint32_t     frame_idx;
  uint32_t     has_fwd, has_bwd;
  L4DTimeQuery cur, fwd, bwd;
  float        flow_basis[4];    /* Lagrange basis at `time` (flow_field.py:102-111,122) */
} L4DFrame;

/* fp32 master parameters in the reference's state_dict layouts. */
typedef struct L4DMasterParams {
  const float* hash_static;                                     /* hash_encoder.hash_static.params */
  const float* hash_dynamic[3][L4D_MAX_TIME_SLICES];            /* hash_encoder.hash_dynamic.P.hash_t.S.params */
  const float* flow_grid;                                       /* flow_net.grid_enc.params */
  const float* planes[L4D_MAX_PLANE_SCALES][6];                 /* planes_encoder.planes.S.C  [1,8,H,W] */
  const float* sigma_net;                                       /* [64*in_pad + 16*64] */
  const float* intensity_net;                                   /* [64*96 + 64*64 + 16*64] */
  const float* raydrop_net;
  const float* flow_mlp[3];                                     /* flow_net.mlp.{0,2,4}.weight */
} L4DMasterParams;

/* fp32 gradient buffers, same layouts as L4DMasterParams (accumulated into). */
typedef struct L4DMasterGrads {
  float* hash_static;
  float* hash_dynamic[3][L4D_MAX_TIME_SLICES];
  float* flow_grid;
  float* planes[L4D_MAX_PLANE_SCALES][6];
  float* sigma_net;
  float* intensity_net;
  float* raydrop_net;
  float* flow_mlp[3];
} L4DMasterGrads;

/* Ray batch + sampling arguments of LiDAR_Renderer.run (model/renderer.py:44-89). */
typedef struct L4DRays {
  const float* rays_o;      /* [n_rays,3] */
  const float* rays_d;      /* [n_rays,3] */
  uint32_t     n_rays;
  uint32_t     n_steps;     /* num_steps (768) */
  uint32_t     perturb;     /* z jitter on/off */
  uint32_t     reserved;
  uint64_t     seed;        /* counter-based jitter stream: u = f(seed, ray_offset+ray, sample) */
  uint64_t     ray_offset;  /* global index of ray 0 (rank-independent streams under ray sharding) */
} L4DRays;

int         l4d_abi_version(void);
const char* l4d_last_error(void);

/* --- parameter staging: fp32 masters -> kernel working set (fp16 tables, channels-last planes,
 *     MLP weights in both orientations).  Call after every optimiser step. ------------------- */
size_t l4d_staged_bytes(const L4DConfig* cfg);
int    l4d_stage_params(const L4DConfig* cfg, const L4DMasterParams* master,
                        void* staged, size_t staged_bytes, void* stream);

/* Same, selectively: L4D_STAGE_TABLES = the three hash-table groups (fp16 casts / slice-pair packing),
 * L4D_STAGE_SMALL = planes (channels-last), MLP working copies and tensor-core operand copies - ONE launch for all of
 * them.  `staged` must have been zero-filled once when it was allocated (operand padding is never rewritten).
 * After l4d_adam_step with a `master` table only L4D_STAGE_SMALL is left to do. */
#define L4D_STAGE_TABLES 1u
#define L4D_STAGE_SMALL  2u
int    l4d_stage_params_ex(const L4DConfig* cfg, const L4DMasterParams* master,
                           void* staged, size_t staged_bytes, uint32_t what, void* stream);

/* --- optimiser step (replaces torch.optim.Adam of main_lidar4d.py:298-300 + the re-staging pass) -----
 * Adam (no weight decay, no amsgrad; torch's fused-kernel operation order) over FLAT fp32 arenas: params, grads and the
 * two moment buffers hold every tensor of the hot path at the same offsets (each tensor starts on a multiple of
 * L4D_ADAM_CHUNK floats).  `groups` = sorted, disjoint [begin,end) float ranges with their learning rate (the
 * reference's param groups, lidar4d.py:226-237); ranges not covered are left alone.  grads are multiplied by inv_scale
 * (1/loss-scale) on the fly; zero_grad=1 clears them in the same pass.  If `master` is given (pointers INTO `params`),
 * a group that is exactly one hash table also emits that table's fp16 working copy into `staged` (static / flow: cast;
 * dynamic slice t: the lo half of pair t and the hi half of pair t-1), i.e. l4d_stage_params_ex(L4D_STAGE_TABLES) for free. */
#define L4D_ADAM_CHUNK 1024
#define L4D_ADAM_MAX_SEGMENTS 64
typedef struct L4DAdamGroup {
  uint64_t begin, end;     /* floats, relative to the arena start; begin % L4D_ADAM_CHUNK == 0, end % 4 == 0 */
  float    lr;
  uint32_t reserved;
} L4DAdamGroup;
int    l4d_adam_step(const L4DConfig* cfg, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                     uint64_t n_floats, const L4DAdamGroup* groups, uint32_t n_groups,
                     float beta1, float beta2, float eps, uint32_t step, float inv_scale, uint32_t zero_grad,
                     const L4DMasterParams* master_or_null, void* staged_or_null, size_t staged_bytes, void* stream);

/* --- LiDAR_Renderer.run forward (renderer.py:44-140 + lidar4d.py:139-223), one fused kernel.
 *     depth[n], image[n,2] (ch0 raydrop, ch1 intensity), wsum[n]; weights/z_vals [n,S] optional.
 *     `saved` (l4d_saved_bytes) enables a later backward; NULL = inference through the single kernel.
 *     With `saved` the split pipeline runs: the buffer is also its exchange workspace and receives the
 *     launch's contracted tables (dynamic hash x time constants of `frame`, time-plane rows) - the
 *     backward must be called with the same saved / frame / rays / staged as its forward. -------- */
size_t l4d_saved_bytes(const L4DConfig* cfg, uint32_t n_rays, uint32_t n_steps);
int    l4d_render_forward(const L4DConfig* cfg, const void* staged, const L4DFrame* frame,
                          const L4DRays* rays, float* depth, float* image, float* wsum,
                          float* weights_or_null, float* zvals_or_null,
                          void* saved_or_null, size_t saved_bytes, void* stream);

/* --- backward of the above: upstream grads g_depth[n], g_image[n,2], optional g_wsum[n],
 *     g_weights[n,S].  Hash-table gradients are accumulated straight into `grads`; plane and
 *     MLP gradients into `grad_work` (zero it ONCE after allocating it: every kernel that consumes a
 *     region - l4d_unstage_grads, the fold kernels of the slice- / basis- / time-row-independent
 *     accumulators it also holds - clears it again) and folded into `grads` by l4d_unstage_grads.
 *     One grad_work per stream that runs backwards concurrently. ------------------------------ */
size_t l4d_grad_work_bytes(const L4DConfig* cfg);
int    l4d_render_backward(const L4DConfig* cfg, const void* staged, const L4DFrame* frame,
                           const L4DRays* rays, const void* saved, size_t saved_bytes,
                           const float* g_depth, const float* g_image,
                           const float* g_wsum_or_null, const float* g_weights_or_null,
                           const L4DMasterGrads* grads, void* grad_work, size_t grad_work_bytes,
                           void* stream);
/* Same; additionally records the CUDA event `ev_hash_done_or_null` (a cudaEvent_t) on `stream` at the point from which the
 * hash_static / hash_dynamic gradient buffers are final (after the static scatter and the dynamic fold, before the flow
 * backward): a data-parallel caller starts all-reducing that bucket - 60 % of the gradient bytes - on a side stream while the
 * rest of the backward still runs (lidar4d_b200/parallel.py). */
int    l4d_render_backward_ex(const L4DConfig* cfg, const void* staged, const L4DFrame* frame,
                              const L4DRays* rays, const void* saved, size_t saved_bytes,
                              const float* g_depth, const float* g_image,
                              const float* g_wsum_or_null, const float* g_weights_or_null,
                              const L4DMasterGrads* grads, void* grad_work, size_t grad_work_bytes,
                              void* ev_hash_done_or_null, void* stream);
int    l4d_unstage_grads(const L4DConfig* cfg, const void* grad_work, size_t grad_work_bytes,
                         const L4DMasterGrads* grads, void* stream);

/* --- LiDAR4D.flow (lidar4d.py:124-137): x[q,3] in [-bound,bound] -> flow[q,6].
 *     flow_saved (q*16 floats) enables the backward. ------------------------------------------ */
int    l4d_flow_forward(const L4DConfig* cfg, const void* staged, const L4DFrame* frame,
                        const float* x, uint32_t n_points, float* flow,
                        float* flow_saved_or_null, void* stream);
int    l4d_flow_backward(const L4DConfig* cfg, const void* staged, const L4DFrame* frame,
                         const float* x, uint32_t n_points, const float* flow_saved,
                         const float* g_flow, const L4DMasterGrads* grads,
                         void* grad_work, size_t grad_work_bytes, void* stream);

/* --- parity / debug entry points (not on the hot path) -------------------------------------- */
/* grid_id: 0 static, 1..3 dynamic xy/xz/yz, 4 flow.  x[n,D] -> idx[n,2^D] (uint32, modulo the
 * level's entries), w[n,2^D]. */
int    l4d_hash_indices(const L4DConfig* cfg, uint32_t grid_id, uint32_t level,
                        const float* x, uint32_t n, uint32_t* idx, float* w, void* stream);
/* LiDAR4D.density on explicit points x[n,3] in [-bound,bound]: sigma[n], geo[n,15],
 * optional features[n,sigma_in_dim] and flow[n,6]. */
int    l4d_density_forward(const L4DConfig* cfg, const void* staged, const L4DFrame* frame,
                           const float* x, uint32_t n, float* sigma, float* geo,
                           float* features_or_null, float* flow_or_null, void* stream);

/* LiDAR4D.attribute (lidar4d.py:191-223) on explicit points: unit directions d[n,3], geo_feat[n,15],
 * optional mask[n] (uint8, 0 = row stays zero) -> out[n,2] = (raydrop, intensity) after the sigmoid. */
int    l4d_attribute_forward(const L4DConfig* cfg, const void* staged, const float* d, const float* geo,
                             const unsigned char* mask_or_null, uint32_t n, float* out, void* stream);

/* --- SURVEY 8(f) "next" row 1: the chamfer / nearest-neighbour op of utils/chamfer3D --------------------------
 * Replaces chamfer_cuda.cpp:17-32 `forward(xyz1,xyz2,dist1,dist2,idx1,idx2)` / `backward(...)` (kernels
 * chamfer3D.cu:11-133,154-174).  xyz1[b,n,3], xyz2[b,m,3] fp32 contiguous; dist1[b,n] = squared distance to the
 * nearest point of xyz2 and idx1[b,n] its index (the smallest index on ties), dist2/idx2 likewise for xyz2.
 * `work` = l4d_chamfer_work_bytes(b,n,m) scratch bytes.  backward ACCUMULATES into g_xyz1[b,n,3], g_xyz2[b,m,3]
 * (zero them first, as dist_chamfer_3D.py:66-70 does). */
size_t l4d_chamfer_work_bytes(uint32_t b, uint32_t n, uint32_t m);
int    l4d_chamfer_forward(const float* xyz1, const float* xyz2, uint32_t b, uint32_t n, uint32_t m,
                           float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                           void* work, size_t work_bytes, void* stream);
int    l4d_chamfer_backward(const float* xyz1, const float* xyz2, uint32_t b, uint32_t n, uint32_t m,
                            const float* g_dist1, const float* g_dist2, const int32_t* idx1, const int32_t* idx2,
                            float* g_xyz1, float* g_xyz2, void* stream);

/* --- SURVEY 8(f) "next" row 2: the thin layers either side of the render kernels, one launch each ------------------------
 * l4d_lidar_rays replaces data/base_dataset.py:15-102 `get_lidar_rays` (+ the GT gather of kitti360_dataset.py:170-178):
 * pose = DEVICE [4][4] row-major cam2world; inds = n flat pixel ids row*W+col (NULL = all H*W pixels in order, n = H*W);
 * writes rays_o[n,3], rays_d[n,3] and, if gt != NULL, gt[n,C] = image[inds,C] (image = DEVICE [H*W,C]).
 * l4d_lidar_loss replaces model/runner.py:179-213 (default criteria: L1 depth, MSE intensity / raydrop, label smoothing):
 * ACCUMULATES the summed loss into *loss (zero it first) and writes d loss / d depth [n] and d loss / d image [n,2]. */
int    l4d_lidar_rays(const float* pose, float fov_up, float fov, uint32_t H, uint32_t W, const long long* inds_or_null,
                      uint32_t n, const float* image_or_null, uint32_t C, float* rays_o, float* rays_d, float* gt_or_null,
                      void* stream);
int    l4d_lidar_loss(const float* depth, const float* image2, const float* gt3, uint32_t n, float alpha_d, float alpha_r,
                      float alpha_i, float smooth, float* loss, float* g_depth, float* g_image2, void* stream);

/* --- profiling aid: while started, CUDA events are recorded on the launch stream around every kernel of
 *     l4d_render_forward / l4d_render_backward.  l4d_profile_stop returns the number of (kernel name, ms)
 *     pairs written (static strings), or a negative error code.  Not thread-safe. --------------------------- */
/* kernels this library has launched in this process so far (bench.py "gpu_launches"; counted at the launch sites). */
unsigned long long l4d_launch_count(void);
int    l4d_profile_start(void);
int    l4d_profile_stop(const char** names, float* ms, int cap);

/* tcgen05/TMEM self-test of the MLP engine's building blocks: C[128,N] = A[128,K] * B[N,K]^T,
 * fp16 operands, fp32 accumulate (N%16==0 in [16,256], K%16==0 in [16,512]). */
int    l4d_tc_selftest(const void* A_half, const void* B_half, float* C, uint32_t N, uint32_t K, void* stream);
/* same with operands pre-arranged in the engine's tile format, M in {64,128}, K-major or MN-major operands
 * (the weight-gradient GEMMs X^T*delta read sample-major tiles as MN-major operands). */
int    l4d_tc_selftest2(const void* A_tile, const void* B_tile, float* C, uint32_t M, uint32_t N, uint32_t K,
                        uint32_t a_mn_major, uint32_t b_mn_major, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIDAR4D_B200_H_ */
