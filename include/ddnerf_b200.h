/*
 * ddnerf_b200 -- C ABI of the B200-native DDNeRF / mip-NeRF per-ray hot path.
 *
 * The reference (dadonda89/DDNeRF) is pure Python/PyTorch and has no FFI; each entry point
 * below replaces one of its Python functions (cited as file:line of the reference) and is what
 * a Python binding (ctypes) of that function calls.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (nothing is allocated or freed
 *     here); outputs are pre-sized by the caller; fp32 row-major unless stated;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no host sync;
 *   - return value 0 = ok, non-zero = error, message via ddnerf_last_error() (thread-local);
 *   - N = rays, S = samples (intervals) of the current pass, fence-posts t[N, S+1];
 *   - N == 0 (an empty chunk or shard) is valid for every per-ray entry point: the per-ray pointers may then be NULL,
 *     no kernel is launched, scalar outputs (losses, regs) are written as for an empty sum.
 */
#ifndef DDNERF_B200_H_
#define DDNERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDNERF_ABI_VERSION 1

/* ---- library ------------------------------------------------------------------------- */
int         ddnerf_version(void);
const char* ddnerf_last_error(void);
/* 1 when the running device is compute capability 10.x (tcgen05 path usable). */
int         ddnerf_device_is_sm100(void);
/* Number of kernels this library has launched in the calling process (monotonic). */
int64_t     ddnerf_launch_count(void);

/* ---- K3: samplers ---------------------------------------------------------------------- */
/* sample_first_cycle, models/samplers.py:30-62.  near/far: per-ray scalars with element
 * strides (rays[N,12] columns 7/8 -> stride 12).  t_rand [N,S+1] or NULL (perturb off). */
int ddnerf_sample_first_cycle(const float* near, const float* far, int64_t ray_stride,
                              const float* t_rand, float* t_out, int64_t N, int S,
                              int lindisp, void* stream);

/* sample_pdf, models/samplers.py:64-121 (mip-NeRF inverse-CDF resampling).
 * bins [N,S+1], weights [N,S], rand [N,n] uniform draws or NULL (det=True), out [N,n];
 * idx_out [N,n] int32 or NULL receives the interval index j of every sample. */
int ddnerf_sample_pdf(const float* bins, const float* weights, const float* rand, float* out,
                      int32_t* idx_out, int64_t N, int S, int n, int pdf_padding, void* stream);

/* sample_pdf_with_mu_sigma, models/samplers.py:124-215 (DDNeRF Gaussian-in-cell resampling,
 * endpoints pinned to near_cfg/far_cfg, ascending sort). */
int ddnerf_sample_pdf_mu_sigma(const float* bins, const float* weights, const float* mus,
                               const float* sigmas, const float* part_inside,
                               const float* left_tail, const float* rand, float* out,
                               int32_t* idx_out, int64_t N, int S, int n, int pdf_padding,
                               float near_cfg, float far_cfg, void* stream);

/* The same resampler fed by ddnerf_composite_dd_forward: `sigmas` are the UNSMOOTHED sigmas; the kernel multiplies by
 * gaussian_smooth_factor (`smooth`, or *smooth_dev when non-NULL: a replayed CUDA graph reads the annealed factor from
 * device memory) and evaluates smoothed_left_tail / smoothed_part_inside (models/models.py:268-273) per cell. */
int ddnerf_sample_pdf_mu_sigma_fused(const float* bins, const float* weights, const float* mus,
                                     const float* sigmas, float smooth, const float* smooth_dev,
                                     const float* rand, float* out, int32_t* idx_out, int64_t N, int S,
                                     int n, int pdf_padding, float near_cfg, float far_cfg, void* stream);

/* The interval search alone (samplers.py:106-116): idx = #{cdf <= u} - 1, on caller-provided
 * CDFs.  Used by the bit-exactness test ("identical CDFs -> identical indices"). */
int ddnerf_find_interval(const float* cdf, const float* u, int32_t* idx_out, int64_t N,
                         int S, int n, void* stream);

/* ---- a2: ray packing --------------------------------------------------------------------- */
/* get_rays_batches, models/models.py:144-158: rays [N,12] = (origin 3, direction 3, radius, near, far,
 * direction / ||direction||_2) from ray_origins [N,3], ray_directions [N,3], ray_radii [N] and the
 * dataset's near / far planes.  One launch instead of the reference's norm / div / ones_like / cat. */
int ddnerf_pack_rays(const float* ray_origins, const float* ray_directions, const float* ray_radii,
                     float near, float far, int64_t N, float* rays, void* stream);

/* ---- K2: encoding ------------------------------------------------------------------------ */
/* cast_rays + integrated_pos_enc + positional_encoding of run_network,
 * models/models.py:117-133, general_utils/math_utils.py:7-166, nerf_helpers.py:127-171.
 * rays [N,12] = (o3,d3,radius,near,far,viewdir3).  Writes the 96 IPE features of sample row
 * r = ray*S+i to enc_out[r*ld_enc + 0..95] and the 27 view-direction features to
 * dir_out[r*ld_dir + 0..26].  ray_shape: 0 cone, 1 cylinder. */
int ddnerf_encode(const float* rays, const float* t_vals, float* enc_out, int64_t ld_enc,
                  float* dir_out, int64_t ld_dir, int64_t N, int S, int ray_shape, void* stream);

/* ---- K1: the NeRF MLP (models/base_architectures.py:3-126) ------------------------------- */
/* Parameter table: 13 (weight,bias) pairs in state_dict order
 *   0..7 layers_xyz.{0..7}, 8 fc_feat, 9 fc_alpha, 10 layers_dir.0, 11 fc_rgb, 12 fc_mu_sigma
 * (entry 12 NULL for MipNeRFModel).  Weights are nn.Linear layout [out,in], fp32. */
#define DDNERF_MLP_NPARAMS 13
typedef struct {
    const float* w[DDNERF_MLP_NPARAMS];
    const float* b[DDNERF_MLP_NPARAMS];
} ddnerf_mlp_params;
typedef struct {
    float* w[DDNERF_MLP_NPARAMS];
    float* b[DDNERF_MLP_NPARAMS];
} ddnerf_mlp_grads;

/* Bytes of activation workspace the fp32 path needs for `rows` sample rows. */
int64_t ddnerf_mlp_f32_workspace_bytes(int64_t rows);

/* fp32 forward over rows = N*S samples, features produced in-kernel from rays/t (K2 fused as
 * producer).  out [rows, C] with C = 4 (rgb,density) or 6 (+raw_mu,raw_sigma).  `workspace`
 * keeps every layer's activations for the backward pass. */
int ddnerf_mlp_f32_forward(const ddnerf_mlp_params* p, const float* rays, const float* t_vals,
                           int64_t N, int S, int ray_shape, int out_channels, float* out,
                           void* workspace, void* stream);
/* fp32 forward from a caller-provided feature matrix x [rows,123] (MipNeRFModel.forward). */
int ddnerf_mlp_f32_forward_x(const ddnerf_mlp_params* p, const float* x, int64_t rows,
                             int out_channels, float* out, void* workspace, void* stream);
/* fp32 backward: grad_out [rows,C] -> ACCUMULATES into g (caller zeroes it); dx [rows,123]
 * or NULL.  `workspace` is the one the matching forward filled. */
int ddnerf_mlp_f32_backward(const ddnerf_mlp_params* p, const ddnerf_mlp_grads* g,
                            const float* grad_out, int64_t rows, int out_channels, float* dx,
                            void* workspace, void* stream);

/* ---- K1, bf16 throughput mode: fused tcgen05 chain (csrc/mlp_tc.cu) ----------------------------
 * Weights are re-packed (after every optimizer step) into bf16 stage images in the order the
 * kernel streams them; biases into an fp32 table.  Rows are processed in 256-row work items; the
 * encoder writes the bf16 operand image of every item (IPE xyz blocks + view-direction block). */
int64_t ddnerf_mlp_tc_wimg_bytes(void);
int64_t ddnerf_mlp_tc_bias_floats(void);
int64_t ddnerf_mlp_tc_items(int64_t rows);
int64_t ddnerf_mlp_tc_enc_bytes(int64_t rows);
int64_t ddnerf_mlp_tc_act_save_bytes(int64_t rows);
int64_t ddnerf_mlp_tc_mask_save_bytes(int64_t rows);
int ddnerf_mlp_tc_pack(const ddnerf_mlp_params* p, int out_channels, void* wimg, float* bias_pack,
                       void* stream);
/* cast_rays + integrated_pos_enc + positional_encoding (models/models.py:117-133) -> bf16 images */
int ddnerf_mlp_tc_encode(const float* rays, const float* t_vals, int64_t N, int S, int ray_shape,
                         void* enc_img, void* stream);
/* MipNeRFModel / DepthMipNeRFModel forward (base_architectures.py:40-61 / 103-126) over `rows`
 * samples; out [rows, C] fp32.  act_save / mask_save: NULL for inference, else the areas sized by
 * the two *_save_bytes functions (kept for the backward kernels). */
int ddnerf_mlp_tc_forward(const void* wimg, const float* bias_pack, const void* enc_img,
                          int64_t rows, int out_channels, float* out, void* act_save,
                          void* mask_save, void* stream);

/* Backward of the bf16 MLP (autograd of base_architectures.py:40-61 / 103-126), two kernels:
 * (1) the dX chain: grad_out [rows,C] -> dZ tile images of every layer (dz_save, same size and
 *     format as act_save), using the ReLU masks the training forward stored;
 * (2) the weight/bias gradients dW_l = dZ_l^T . A_{l-1}, db_l = colsum(dZ_l), ACCUMULATED with fp32
 *     atomics into `grads` (caller zeroes the buffers; [out,in] nn.Linear layout, any 4-byte
 *     alignment).  No gradient is produced for the encoded inputs (nothing upstream is trainable).
 * max_ctas > 0 caps the number of CTAs (= SMs) a call occupies, so that the HBM-read-bound dW of one pass
 * and the tensor/HBM-write-bound dX chain of the next can share the GPU from two streams; 0 = all SMs. */
int ddnerf_mlp_tc_backward_dx(const void* wimg, const float* bias_pack, const float* grad_out,
                              int64_t rows, int out_channels, const void* mask_save,
                              void* dz_save, int max_ctas, void* stream);
int ddnerf_mlp_tc_backward_dw(const void* act_save, const void* dz_save, const void* enc_img,
                              const float* grad_out, const ddnerf_mlp_grads* grads, int64_t rows,
                              int out_channels, int max_ctas, void* stream);

/* Host-only consistency hooks (no device work): the static ring/op programs of the chain kernels
 * (0 = consistent) and the split of backward_dw over `sms` SMs: the (layer-op, tile) line cut into at most `sms`
 * equal-cost pieces, a piece that crosses a layer-op boundary being two (rarely three) work items run back to back by
 * one CTA; written as (op, first tile, end tile) uint32 triples; returns the number of work items (<= sms + 12). */
/* The forward with the encoder inside the kernel (models/models.py:117-142 as ONE launch): two extra warps of the chain
 * kernel compute cone -> Gaussian, IPE and the view-direction encoding of the next 256-row work item while the GEMMs of
 * the current one run.  Give enc_img (ddnerf_mlp_tc_enc_bytes(N*S) bytes, written here and read by
 * ddnerf_mlp_tc_backward_dw) for the training forward, or enc_scratch (ddnerf_mlp_tc_enc_scratch_bytes() bytes, a
 * per-SM double buffer that stays in L2) for inference -- exactly one of the two. */
int64_t ddnerf_mlp_tc_enc_scratch_bytes(void);
int ddnerf_mlp_tc_forward_rays(const void* wimg, const float* bias_pack, const float* rays,
                               const float* t_vals, int64_t N, int S, int ray_shape, int out_channels,
                               float* out, void* enc_img, void* enc_scratch, void* act_save,
                               void* mask_save, void* stream);
int ddnerf_mlp_tc_program_check(void);
/* Kernel variant of the forward and dX chains (base_architectures.py:89-126, same arithmetic, bit-identical results):
 * 1 = CTA pairs (cluster of 2, tcgen05 cta_group::2 with the weight chunks split across the pair; default),
 * 0 = one CTA per 256-row work item, -1 = re-read the DDNERF_TC_PAIR environment variable.  Returns the previous setting. */
int ddnerf_mlp_tc_set_pair_mode(int mode);
/* Diagnostic hook of the dW kernel: a device buffer of >= 4 * 480 uint64 receives, per work item of the next
 * launches, {layer-op, tiles, cycles until its last MMA completed, cycles of its flush}; NULL switches it off. */
int ddnerf_mlp_tc_dw_set_profile_buffer(void* dev_u64);

/* Diagnostic hook: a device buffer of >= 8 * n_SMs uint64 into which the chain kernels write per-CTA cycle
 * counters (issuer total / waiting on epilogues / waiting on ring stages, epilogues waiting on MMAs / busy /
 * count); NULL (default) switches the instrumentation off. */
int ddnerf_mlp_tc_set_profile_buffer(void* dev_u64);
int ddnerf_mlp_tc_dw_plan(int64_t rows, int sms, uint32_t* triples, int max_items);

/* Descriptor self-test of the tcgen05 path (test infrastructure of the bf16 MLP): one CTA computes
 * D[128,N] = A.B^T from two operand tile images given in their shared-memory byte layout, with the
 * shared-memory descriptors (start address 0), instruction descriptor and per-k16-step address
 * stepping {a_step, a_steps_per_block, a_block_pitch, b_step, b_steps_per_block, b_block_pitch}
 * supplied by the caller. */
int ddnerf_tc_gemm_selftest(const void* a_img, int64_t a_bytes, const void* b_img, int64_t b_bytes,
                            float* d_out, int N, int nk16, uint64_t a_desc, uint64_t b_desc,
                            uint32_t idesc, const uint32_t* stepping, void* stream);

/* Diagnostic: cycles one CTA (pair = 0) or a CTA pair (pair = 1, tcgen05 cta_group::2) needs for n_mma
 * back-to-back M = 128 (256) x N x 16 bf16 MMAs on resident operands, committing to an mbarrier every
 * `commit_every` MMAs (0 = never).  cycles_out: one uint64 per CTA of a 148-CTA launch. */
int ddnerf_tc_mma_rate(int pair, int N, int n_mma, int commit_every, void* cycles_out, void* stream);

/* ---- K4: alpha compositing (general_utils/volume_rendering_utils.py:6-84) ---------------- */
/* raw [N,S,raw_stride] (channels 0..3 = r,g,b,density), t [N,S+1], rd = ray directions with
 * row stride rd_stride, noise [N,S] unit normal or NULL (density += noise*noise_std),
 * mus [N,S] or NULL.  Outputs: rgb_map [N,3], disp [N], acc [N], weights [N,S], depth [N],
 * cdisp [N] (only when mus), rgb [N,S,3] or NULL. */
int ddnerf_composite_forward(const float* raw, int raw_stride, const float* t, const float* rd,
                             int64_t rd_stride, const float* noise, float noise_std,
                             const float* mus, int white_background, int blender,
                             float* rgb_map, float* disp, float* acc, float* weights,
                             float* depth, float* cdisp, float* rgb, int64_t N, int S,
                             void* stream);
/* Analytic backward.  Cotangents g_* may each be NULL (= zero).  Writes g_raw [N,S,4] and
 * g_mus [N,S] (when mus and g_mus non-NULL). */
int ddnerf_composite_backward(const float* raw, int raw_stride, const float* t, const float* rd,
                              int64_t rd_stride, const float* noise, float noise_std,
                              const float* mus, int white_background, int blender,
                              const float* g_rgb_map, const float* g_disp, const float* g_acc,
                              const float* g_weights, const float* g_depth, const float* g_cdisp,
                              float* g_raw, float* g_mus, int64_t N, int S, void* stream);

/* ---- f1: ray generation (general_utils/nerf_helpers.py:67-125 get_ray_bundle; with ndc != 0 followed by
 * data_utils/dataset_helpers.py:3-42 ndc_mipnerf_rays(H, W, focal, ro, rd, near = ndc_near)) ------------- */
/* c2w_host: HOST pointer to the first 12 floats of the row-major 4x4 camera-to-world matrix (read during the
 * call).  Pixel rows [row_lo, row_hi) of the H x W frame are written: ray_origins / ray_directions
 * [rows, W, 3], radii [rows, W, 1] (directions un-normalised, as the reference returns them). */
int ddnerf_ray_bundle(int H, int W, double focal, const float* c2w_host, int ndc, float ndc_near,
                      int row_lo, int row_hi, float* ray_origins, float* ray_directions, float* radii,
                      void* stream);

/* Same with the pose in DEVICE memory (12 floats), so that a captured CUDA graph can be replayed for a new pose. */
int ddnerf_ray_bundle_dev(int H, int W, double focal, const float* c2w_dev, int ndc, float ndc_near,
                          int row_lo, int row_hi, float* ray_origins, float* ray_directions, float* radii,
                          void* stream);

/* ---- f4: frame post-processing (validation_utils/visualization.py:11-27 cast_to_disparity_image /
 * cast_to_image; render_video.py:96-101 video frame) ---------------------------------------------------- */
/* minmax[0..1] <- min, max of disp[0..n) (NaNs skipped).  workspace: >= 16 bytes, zeroed ONCE by the caller; the
 * call leaves it zeroed again.  With a frame split across ranks the caller reduces minmax (MIN, MAX) between the
 * two calls. */
int ddnerf_frame_minmax(const float* disp, int64_t n, void* workspace, float* minmax, void* stream);
/* rgb [rows*W,3], disp [rows*W] -> any of: rgb8 [rows,W,3] = trunc(clamp(255 rgb, 0, 255)); disp8 [rows,W] =
 * trunc(255 clamp((disp - min)/(max - min), 0, 1)); video_bgr [rows, 2W, 3] = [rgb8 | disp8 x3] in BGR order. */
int ddnerf_frame_pack_u8(const float* rgb, const float* disp, const float* minmax, uint8_t* rgb8, uint8_t* disp8,
                         uint8_t* video_bgr, int rows, int W, void* stream);

/* ---- f3: device-resident training ray store (data_utils/dataset.py:8-59 TrainDataset) ----------------- */
/* rows [n,12] <- {origin 3, direction 3, radius 1, target rgb 3, 0, 0} per ray (48-byte rows, 16-byte aligned). */
int ddnerf_raystore_pack(const float* ray_origins, const float* ray_directions, const float* radii,
                         const float* target_rgb, int64_t n, float* rows, void* stream);
/* Batch i takes row base + idx[i] (dataset.py:52-53: origins[idxs], directions[idxs], radii[idxs], target[idxs]).
 * idx: int64 on the device.  An index outside [0, total_rows) sets *bad_index_flag (may be null) and reads row 0. */
int ddnerf_raystore_gather(const float* rows, int64_t total_rows, const int64_t* idx, int64_t n, int64_t base,
                           float* ray_origins, float* ray_directions, float* radii, float* target_rgb,
                           int* bad_index_flag, void* stream);

/* ---- K4 + DDNeRF depth-distribution head (the coarse pass of DDNerfModel.predict, models/models.py:242-273) ----------
 * The compositor above with the glue around it folded in.  raw6 [N,S,6] = (r,g,b,density,raw_mu,raw_sigma), read in
 * place.  Writes the compositor outputs (depth = the corrected depth, volume_rendering_utils.py:76-83) plus
 *   mus = sigmoid(raw_mu) [N,S], sigmas = sigmoid(raw_sigma) + 0.001 [N,S]                       (models.py:245-246)
 *   regs[4] = {mus_loss, sig_loss, mus_reg, sig_reg} = {sum raw_mu^2 / N, sum raw_sigma^2 / N, dist_reg_coef x each}
 *                                                                                             (models.py:248-252)
 * The tails of models.py:254-258 / 268-273 are computed by their consumers (ddnerf_sample_pdf_mu_sigma_fused,
 * ddnerf_dp_loss_* with NULL tails).  scratch: ddnerf_composite_dd_scratch_floats(N) floats, contents irrelevant. */
int64_t ddnerf_composite_dd_scratch_floats(int64_t N);
int ddnerf_composite_dd_forward(const float* raw6, const float* t, const float* rd, int64_t rd_stride,
                                const float* noise, float noise_std, int white_background, int blender,
                                float dist_reg_coef, float* rgb_map, float* disp, float* acc,
                                float* weights, float* depth, float* cdisp, float* mus, float* sigmas,
                                float* regs, float* scratch, int64_t N, int S, void* stream);
/* Cotangents (each may be NULL) of the maps, the weights, mus, sigmas and regs[4] -> g_raw6 [N,S,6], the cotangent of the
 * network output (sigmoid' of the mu/sigma head and the regulariser gradients 2 raw / N included). */
int ddnerf_composite_dd_backward(const float* raw6, const float* t, const float* rd, int64_t rd_stride,
                                 const float* noise, float noise_std, int white_background, int blender,
                                 float dist_reg_coef, const float* g_rgb_map, const float* g_disp,
                                 const float* g_acc, const float* g_weights, const float* g_depth,
                                 const float* g_cdisp, const float* g_mus, const float* g_sigmas,
                                 const float* g_regs, float* g_raw6, int64_t N, int S, void* stream);

/* ---- K5: depth-distribution loss (models/dd_utils.py:6-78) ------------------------------- */
/* scratch: >= 4 + 2*N floats of caller-owned workspace: [0] = sum of the per-ray KL values, [1] = number of
 * rays kept by the blender row filter (dd_utils.py:12-28), [4 .. 4+N) per-ray KL, [4+N .. 4+2N) per-ray kept
 * flag; the mean is taken over the per-ray values in a fixed order (bit-reproducible).
 * loss_out: 1 float = kl_div(..., 'mean'). */
/* lt0 / pin0 may both be NULL: left_tail = Phi((0 - mu) / sigma) and part_inside = Phi((1 - mu) / sigma) - left_tail
 * (models/models.py:254-258) are then evaluated per coarse cell inside the kernels (constants of the loss, as the
 * reference detaches them). */
int ddnerf_dp_loss_forward(const float* t1, const float* t0, const float* w1, const float* w0,
                           const float* mus0, const float* sigmas0, const float* lt0,
                           const float* pin0, int blender, float* loss_out, float* scratch,
                           int64_t N, int S0, int S1, void* stream);
/* g_loss: device pointer to the scalar cotangent.  Writes g_w0, g_mus0, g_sigmas0 [N,S0].
 * scratch is the buffer the forward filled (holds the relevant-ray count). */
int ddnerf_dp_loss_backward(const float* t1, const float* t0, const float* w1, const float* w0,
                            const float* mus0, const float* sigmas0, const float* lt0,
                            const float* pin0, int blender, const float* g_loss,
                            const float* scratch, float* g_w0, float* g_mus0, float* g_sigmas0,
                            int64_t N, int S0, int S1, void* stream);
/* The loss term as DDNerfModel.predict forms it (models/models.py:287-289):
 *   loss_out[0] = kl_div * scale + regs[2] + regs[3]
 * with scale = the number of fine cells (t_vals.shape[1] - 1) and regs = {mus_loss, sig_loss, mus_reg, sig_reg} as
 * ddnerf_composite_dd_forward wrote them.  Same kernels as ddnerf_dp_loss_forward (the finishing kernel adds the terms):
 * replaces the reference's mul + two adds, and in the backward the slice / mul / add chain autograd builds for them. */
int ddnerf_dp_loss_total_forward(const float* t1, const float* t0, const float* w1, const float* w0,
                                 const float* mus0, const float* sigmas0, const float* lt0,
                                 const float* pin0, int blender, float scale, const float* regs,
                                 float* loss_out, float* scratch, int64_t N, int S0, int S1, void* stream);
/* g_loss: device pointer to the cotangent of loss_out[0].  Writes g_w0, g_mus0, g_sigmas0 [N,S0] (cotangent times scale
 * through the KL term) and g_regs[4] = {0, 0, g, g}. */
int ddnerf_dp_loss_total_backward(const float* t1, const float* t0, const float* w1, const float* w0,
                                  const float* mus0, const float* sigmas0, const float* lt0,
                                  const float* pin0, int blender, float scale, const float* g_loss,
                                  const float* scratch, float* g_w0, float* g_mus0, float* g_sigmas0,
                                  float* g_regs, int64_t N, int S0, int S1, void* stream);

/* ---- training-step tail on the flat parameter bucket (train_model.py:156-177) ------------ */
/* loss = sum_j coef_j * mse(rgb_j, target); writes g_rgb_j = coef_j*2*(rgb_j-target)/(3N).
 * mse_out [3] = {mse(rgb_0), mse(rgb_1), coef_0 mse_0 + coef_1 mse_1}.  rgb1/g_rgb1 may be NULL. */
int ddnerf_mse_loss(const float* rgb0, const float* rgb1, const float* target, float coef0,
                    float coef1, float* g_rgb0, float* g_rgb1, float* mse_out, int64_t N,
                    void* stream);
/* torch.optim.Adam (no weight decay, no amsgrad) on a flat fp32 bucket; grad_scale multiplies
 * the gradient first (1/world_size after an all-reduce sum). */
int ddnerf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                     int64_t n, float lr, double beta1, double beta2, double eps, int step,
                     float grad_scale, void* stream);

/* The same update with its scalars in device memory: hyper[10] = {lr, beta1, beta2, eps, 1 - beta1^step,
 * sqrt(1 - beta2^step), grad_scale, (unused here), 1 - beta1, 1 - beta2} (the complements rounded from double, as
 * torch.optim.Adam passes them).  Lets a CUDA graph of the whole training step be replayed while the
 * learning rate and bias corrections change every step. */
int ddnerf_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                         int64_t n, const float* hyper, void* stream);

/* The per-iteration host scalars of the driver loop, train_model.py:135-150 (annealed gaussian_smooth_factor,
 * learning_rate_decay of general_utils/nerf_helpers.py:211-245) and Adam's bias corrections, computed on the device
 * from a device-resident counter so that a replayed CUDA graph of the step reads nothing the host mutates.
 * state[2] (int64, device) = {iteration i, Adam steps taken}: read, then both advanced by one.
 * hyper[10] (device) receives {lr(i), beta1, beta2, eps, 1-beta1^t, sqrt(1-beta2^t), grad_scale, smooth(i), 1-beta1,
 * 1-beta2}, t = steps + 1.
 * sched[13] (HOST doubles, read at call time) = {lr_init, lr_final, max_steps, lr_delay_steps, lr_delay_mult, beta1,
 * beta2, eps, grad_scale, smooth0, dsmooth, final_smooth, finnish_smooth}. */
int ddnerf_train_schedule(int64_t* state, float* hyper, const double* sched, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDNERF_B200_H_ */
