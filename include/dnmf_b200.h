/*
 * dnmf_b200 -- C ABI of the B200-native dNMF fit hot path.
 *
 * The reference (mathdiane/dNMF) is pure Python and has no FFI; the boundary it exposes is the
 * class surface of Demix/dNMF.py (ExponentialFP / DeformableNMF).  Each entry point below names
 * the reference code it replaces (file:line relative to the reference tree).  The Python host
 * layer in dnmf_b200/model.py binds these with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).
 *
 * Conventions
 *   - every call returns 0 on success, non-zero on failure; dnmf_last_error() gives the text.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - "_dev" pointers are device pointers, "_host" pointers are host pointers.
 *   - layouts are the reference's own: beta[10][3][T], C[K][T] (T innermost), frames
 *     [n][X][Y][Z] (Z innermost, what the reference's DataLoader yields, Demix/dNMF.py:214).
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 *   - frame ids (frame_ids_dev / frame_ids_host) index this context's slab: 0 <= id < T.  Device ids are checked
 *     on the device before any kernel indexes with them (the kernels read a copy clamped into the slab, so an id
 *     out of range cannot touch memory outside it); asynchronous calls report such an id at the NEXT entry
 *     point or at dnmf_check_status, synchronous ones (dnmf_mu_stats, dnmf_motion_step_host, the frame-parallel
 *     dnmf_motion_epoch) immediately.  An id listed twice in one gradient batch gets the sum of its
 *     occurrences' gradients, like the reference's autograd (index_put with accumulate); dnmf_mu_stats and
 *     dnmf_motion_step_host reject duplicates.
 */
#ifndef DNMF_B200_H
#define DNMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dnmf_ctx dnmf_ctx;

#define DNMF_ABI_VERSION 7

int dnmf_abi_version(void);
const char* dnmf_last_error(void);

/* Context for one GPU's slab of T frames of an X*Y*Z volume with K neurons.
 * Replaces the constructor state of ExponentialFP.__init__ (Demix/dNMF.py:19-43): voxel grid,
 * basis and the dense Gaussian volume A are never materialised. */
int dnmf_create(dnmf_ctx** out, int X, int Y, int Z, int K, int T, int device);
void dnmf_destroy(dnmf_ctx* ctx);

/* Kernel 1a: per-axis truncated Gaussian tables and integer node ranges from positions/sigma
 * (Demix/dNMF.py:29-40).  cutoff <= 0 disables truncation.  pos_host[K][3], sigma_host[K]. */
int dnmf_set_footprints(dnmf_ctx* ctx, const float* pos_host, const float* sigma_host, float cutoff,
                        void* stream);
int dnmf_get_ranges(dnmf_ctx* ctx, int32_t* ranges_host /* [K][3][2] */);
int dnmf_get_table(dnmf_ctx* ctx, int axis, float* table_host /* [K][s_axis+3][2] */);

/* Launch geometry of the fused kernel: warps per CTA along x and y (warp footprint is 8x4 voxels),
 * tile depth tz (0 = whole Z), staged-slot capacity (0 = automatic), y-adjacent sub-tiles carried by each
 * warp (1, or 2 for the 1x1, 2x1 and 2x2 warp layouts: the two sub-tiles then run as one packed FP32x2 stream).
 * warps_z = 2 or 4 (1x1 layout with two sub-tiles only) puts that many warps on the one 8 x 8 tile, each marching its
 * part of the z range: the short lists of the small tile with enough warps per SM for dense configurations.
 * Without this call the library picks a layout from the list lengths.  One CTA walks up to 8 consecutive
 * frames of the batch through its tile (environment override for tuning: DNMF_FPC=<1..32>). */
int dnmf_set_tiling(dnmf_ctx* ctx, int warps_x, int warps_y, int tz, int slot_capacity, int subtiles_y,
                    int warps_z);
int dnmf_get_tiling(dnmf_ctx* ctx, int32_t* out /* [12]: tx,ty,tz,ntx,nty,ntz,warps_x,warps_y,cap,subtiles_y,
                                                     exact_fast_division_verified,warps_z */);

/* Resident video slab [T][X][Y][Z] on the device (ingest of SimulatedVideoDataset.video,
 * Demix/dNMF.py:203,214-215; negative values are clamped to 0 like __getitem__ does). */
int dnmf_upload_frames(dnmf_ctx* ctx, const float* frames_host, int t0, int n, int clamp_negative,
                       void* stream);
int dnmf_video_devptr(dnmf_ctx* ctx, float** out_dev);
/* Zero-copy alternative for frames that are already on the device: the context reads the caller's slab
 * frames_dev[T][X][Y][Z] in place (the caller keeps it alive and owns it; NULL detaches).  clamp_negative != 0
 * clamps it IN PLACE, which is what the reference's dataset does to its own video (Demix/dNMF.py:215). */
int dnmf_attach_frames(dnmf_ctx* ctx, float* frames_dev, int clamp_negative, void* stream);

/* Kernel 1b (stand-alone form): deterministic neuron-to-tile binning for B frames; the fused
 * kernel runs the same device code in its prologue.  Outputs are device arrays:
 * counts[B*nt], offsets[B*nt+1] (int64), windows[B*nt][3][2], ids[ids_capacity]; the total list
 * length is returned through total_host (ids beyond ids_capacity are not written). */
int dnmf_bin_tiles(dnmf_ctx* ctx, const float* beta_dev, const int32_t* frame_ids_dev, int B,
                   int32_t* counts_dev, int64_t* offsets_dev, int32_t* windows_dev, int32_t* ids_dev,
                   int64_t ids_capacity, int64_t* total_host, void* stream);

/* Kernel 2: fused forward + loss + analytic beta gradient of one minibatch
 * (ExponentialFP.forward + F.mse_loss + backward, Demix/dNMF.py:54-58,187-190).
 * frames_dev == NULL reads the resident slab by frame id, else frames_dev is [B][X][Y][Z].
 * grad_dev[10][3][T] receives the columns of the batch frames (scaled by 2/(B_global*N));
 * sse_dev[B] (double) receives each frame's sum of squared residuals. */
int dnmf_loss_grad(dnmf_ctx* ctx, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                   int B_global, const float* beta_dev, const float* C_dev, float* grad_dev,
                   double* sse_dev, void* stream);

/* Affine fits (rows 4..9 of beta frozen, the `affine` argument of the step calls below): with affine != 0
 * dnmf_loss_grad no longer produces the gradient rows 4..9 (they are returned as zero, which is what the Adam step
 * of an affine fit makes of them) and frames whose quadratic coefficients are all zero take main loops without the
 * z^2 terms.  The step calls do the same for the duration of a call made with affine != 0.  The reference has no
 * such switch (Demix/dNMF.py:53-58 always carries the full basis); 0 (the default) is its behaviour. */
int dnmf_set_affine(dnmf_ctx* ctx, int affine);

/* Kernel 3a: dense Adam over all 30*T entries (torch.optim.Adam step at Demix/dNMF.py:191,
 * demo.py:42).  `step` is the 1-based step count; affine != 0 freezes rows 4..9.  grad_dev is
 * consumed and reset to zero.  loss_dev (optional) receives sum(sse[0..B)) / (B_global*N). */
int dnmf_adam_step(dnmf_ctx* ctx, float* beta_dev, float* grad_dev, float* m_dev, float* v_dev,
                   double lr, double beta1, double beta2, double eps, int64_t step, int affine,
                   const double* sse_dev, int B, int B_global, double* loss_dev, void* stream);

/* One iteration of DeformableNMF.update_motion's inner loop (Demix/dNMF.py:186-191):
 * dnmf_loss_grad into the context's gradient buffer followed by dnmf_adam_step. */
int dnmf_motion_step(dnmf_ctx* ctx, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                     int B_global, float* beta_dev, float* m_dev, float* v_dev, const float* C_dev,
                     double lr, double beta1, double beta2, double eps, int64_t step, int affine,
                     double* loss_dev, void* stream);

/* All minibatches of one epoch of update_motion over the RESIDENT video (Demix/dNMF.py:185-191 with the
 * loop over the DataLoader inside the library): batch i = frame_ids_dev[batch_offsets_host[i] ..
 * batch_offsets_host[i+1]), its global batch size = its length * global_batch_scale (ranks of a sharded fit),
 * Adam step number first_step + i, loss written to loss_dev[i] (may be NULL).  Bit-identical results to
 * nbatches calls of dnmf_motion_step.  When no frame occurs twice in the epoch (a DataLoader pass) the
 * minibatches touch disjoint columns of beta and of the Adam moments -- the model has no parameter shared
 * between frames -- so the library runs them as ONE fused launch over all frames of the epoch; each column
 * replays the zero-gradient Adam steps of the other minibatches before and after its own gradient step, in
 * the same fp32 operations as the step-by-step schedule.  Epochs with repeated frames run batch by batch. */
int dnmf_motion_epoch(dnmf_ctx* ctx, const int32_t* frame_ids_dev, const int32_t* batch_offsets_host,
                      int nbatches, int global_batch_scale, float* beta_dev, float* m_dev, float* v_dev,
                      const float* C_dev, double lr, double beta1, double beta2, double eps,
                      int64_t first_step, int affine, double* loss_dev, void* stream);
/* sequential: 1 = dnmf_motion_epoch always runs batch by batch, 0 = automatic, negative = leave unchanged.
 * last_parallel_out (may be NULL): 1 when the most recent dnmf_motion_epoch ran frame-parallel. */
int dnmf_epoch_mode(dnmf_ctx* ctx, int sequential, int* last_parallel_out);

/* Same, with HOST buffers: copies frames_host[B][X][Y][Z] and the ids to the device, runs the
 * step and copies the loss back (synchronises the stream).  This is the end-to-end call. */
int dnmf_motion_step_host(dnmf_ctx* ctx, const float* frames_host, const int32_t* frame_ids_host,
                          int B, int B_global, float* beta_dev, float* m_dev, float* v_dev,
                          const float* C_dev, double lr, double beta1, double beta2, double eps,
                          int64_t step, int affine, double* loss_host, void* stream);

/* ExponentialFP.forward outputs (Demix/dNMF.py:54-62): A_tC[B][X][Y][Z]; optionally the dense
 * A_t[B][K][X][Y][Z] and the normalised grid[X][Y][Z][3][B] (small problems only). */
int dnmf_forward(dnmf_ctx* ctx, const int32_t* frame_ids_dev, int B, const float* beta_dev,
                 const float* C_dev, float* AtC_dev, float* At_dev, float* grid_dev, void* stream);

/* Kernel 3b: sufficient statistics of the trace update for B frames
 * (G_t = A_t^T A_t, b_t = A_t^T Y_t, Demix/dNMF.py:141-142; hoisted out of the sweep loop, which
 * is bit-identical in the reference) accumulated in fp64 into the context. */
int dnmf_mu_stats(dnmf_ctx* ctx, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                  const float* beta_dev, void* stream);
int dnmf_get_mu_stats(dnmf_ctx* ctx, int t, double* G_host /* [K][K] */, double* b_host /* [K] */);
/* dnmf_mu_stats has two device paths: the fused kernel's tiles with register-blocked accumulators (short
 * neuron lists, the default when the tiling allows it) and a shared-memory panel kernel (any list length,
 * also the automatic redo when a list outgrows the staged capacity).  The sweeps read G either dense or
 * compacted to the static neighbour lists (neurons whose truncated supports overlap; the default when the
 * lists are shorter than K/2); without temporal coupling (gamma None or 0) dnmf_mu_sweeps runs all sweeps of
 * a frame in one CTA.  A third statistics kernel takes the dense configurations (lists of up to 127 neurons per
 * 8 x 8 x Z tile): the panel Gram on the tensor cores (tcgen05.mma.kind::tf32, 3xTF32 split, accumulators in TMEM);
 * automatic order: fused tiles -> tensor-core panel -> SIMT panel.  Whatever the first stage, the per-frame sum
 * over tiles is a fixed-order fp64 reduction (bitwise reproducible, G_t exactly symmetric).
 * flags: bit 0 = skip the fused tiles (SIMT panel unless bit 3), bit 1 = always dense sweeps, bit 2 = one launch
 * per sweep, bit 3 = tensor-core panel first, 0 = automatic, negative = leave unchanged.  last_path_out (may be
 * NULL): bit 0 = the most recent dnmf_mu_stats ran on the fused tiles, bit 1 = the most recent dnmf_mu_begin set up
 * sparse sweeps, bit 2 = the most recent dnmf_mu_stats ran on the tensor-core panel kernel. */
int dnmf_mu_path(dnmf_ctx* ctx, int flags, int* last_path_out);

/* Multiplicative sweeps C <- C (b + g nbr) / (G C + 2 g C + 1e-32) over all T frames in fp64
 * (Demix/dNMF.py:143-148,172-173); use_gamma == 0 reproduces gamma=None.
 * dnmf_mu_sweeps = begin + iters * sweep + end.  The split form lets a multi-GPU caller exchange
 * the boundary trace columns between sweeps: halo_prev/next_dev (double[K]) replace the
 * edge-replicated neighbours of the first/last local frame when non-NULL. */
int dnmf_mu_begin(dnmf_ctx* ctx, const float* C_dev, void* stream);
int dnmf_mu_sweep(dnmf_ctx* ctx, double gamma, int use_gamma, const double* halo_prev_dev,
                  const double* halo_next_dev, void* stream);
int dnmf_mu_boundary(dnmf_ctx* ctx, double* first_dev, double* last_dev, void* stream);
int dnmf_mu_end(dnmf_ctx* ctx, float* C_dev, void* stream);
int dnmf_mu_sweeps(dnmf_ctx* ctx, float* C_dev, double gamma, int use_gamma, int iters, void* stream);

/* Nearest-neighbour registered video Y_i of ExponentialFP.spatial_pushforward / image_iwarp
 * (Demix/dNMF.py:81-83,90-103) for B frames: out[B][X][Y][Z]. */
int dnmf_iwarp(dnmf_ctx* ctx, const float* frames_dev, const int32_t* frame_ids_dev, int B,
               const float* beta_dev, float* out_dev, void* stream);

/* EXTENSION (no counterpart in the reference, off unless called): gradients of the same loss with respect
 * to the shared parameters -- positions pos[K][3], widths sigma[K] and a scalar background b added to the
 * model -- for callers that learn them next to the deformation.  dnmf_ext_enable allocates the derivative
 * tables (call dnmf_set_footprints afterwards).  dnmf_ext_loss_grad does what dnmf_loss_grad does (with b
 * added to Yhat) and also writes gpos_dev[K][3], gsig_dev[K], gbg_dev[1] (fp64, scaled by 2/(B_global*N),
 * summed over the B frames); with frames sharded over GPUs these three arrays are the only gradients that
 * need an all-reduce. */
int dnmf_ext_enable(dnmf_ctx* ctx);
int dnmf_ext_loss_grad(dnmf_ctx* ctx, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                       int B_global, const float* beta_dev, const float* C_dev, float background,
                       float* grad_beta_dev, double* sse_dev, double* gpos_dev, double* gsig_dev,
                       double* gbg_dev, void* stream);

/* The same iteration without leaving the device (what DeformableNMF.update_motion runs after
 * enable_shared_learning): dnmf_ext_step_begin = the two gradient kernels with the context's own background, the
 * beta gradient kept in the context and ONE packed buffer for the caller's all-reduce,
 *   packed[0..3K) = dL/dpos, [3K..4K) = dL/dsigma, [4K] = dL/db, [4K+1] = sum of the batch's SSE    (doubles);
 * dnmf_ext_step_end (after the all-reduce) = Adam on beta (dnmf_adam_step), Adam with the same formula on
 * pos / sigma / b from the packed buffer (state kept in the context; sigma is clamped to >= sigma_min), then ranges,
 * tables and the static candidate lists rebuilt on the device at the current launch geometry.  No host
 * synchronisation in either call; loss_dev (optional) = packed[4K+1] / (B_global * N).
 * dnmf_ext_set_params sets the background (and optionally clears the Adam state of the shared parameters),
 * dnmf_ext_get_params copies pos[K][3], sigma[K], b into the caller's device buffers. */
int dnmf_ext_step_begin(dnmf_ctx* ctx, const float* frames_dev, const int32_t* frame_ids_dev, int B, int B_global,
                        const float* beta_dev, const float* C_dev, double* packed_dev, void* stream);
int dnmf_ext_step_end(dnmf_ctx* ctx, float* beta_dev, float* m_dev, float* v_dev, double lr, double beta1,
                      double beta2, double eps, int64_t step, int affine, const double* packed_dev, int B_global,
                      double lr_pos, double lr_sigma, double lr_background, float sigma_min, double* loss_dev,
                      void* stream);
int dnmf_ext_set_params(dnmf_ctx* ctx, float background, int reset_adam_state, void* stream);
int dnmf_ext_get_params(dnmf_ctx* ctx, float* pos_dev_out, float* sigma_dev_out, float* background_dev_out,
                        void* stream);

/* FFMA microbenchmark: best-of-`repeats` dense FP32 throughput of the device in TFLOP/s (FMA = 2);
 * the roofline denominator for the FP32-bound fused kernel (MEASURED_PEAKS.json has no FP32 entry). */
int dnmf_measure_fp32_peak(int device, int repeats, double* tflops_out);

/* Build variant of the library: bit 0 = checked build (-DDNMF_CHECKED: device-side assertions on every unclamped
 * index of the kernels; `python -m dnmf_b200.build --checked`).  No reference counterpart (test infrastructure). */
int dnmf_build_info(void);
/* Launches a kernel whose only statement is a failing device assertion and synchronises: returns non-zero in a
 * checked build (the context is unusable afterwards: call it from a throw-away process), 0 in the normal build.
 * Proves that the checked build's assertions are live.  No reference counterpart (test infrastructure). */
int dnmf_debug_trip_assert(int device);

/* ---- callers and data formats either side of the fit path ---------------------------- */

/* Synthetic generator, clean frames (WUtils/Simulator.py:66-77,197-203: every cell is a Gaussian with
 * cov = shape_std * I rescaled to peak 1, weighted by its trace): out[nT][X][Y][Z] =
 * sum_k traces[k][t] * exp(-|p - pos[k][:][t]|^2 / (2 * shape_std)) for t = t0 .. t0+nT-1.
 * pos_dev[K][3][T], traces_dev[K][T] (float32, device).  Z <= 64.  No context needed. */
int dnmf_render_cells(const float* pos_dev, const float* traces_dev, int K, int T, int t0, int nT, int X, int Y,
                      int Z, float shape_std, float* out_dev, void* stream);

/* DeformableNMF.update_spatial (Demix/dNMF.py:151-160), fp64 like the reference's numpy einsum, voxels P =
 * the flattened leading axes: A[P][K], C[K][T], Y_i[P][T] -> out[P][K] = A * (Y_i C^T) / (A (C C^T) + gamma D + 1e-32).
 * use_D: 0 = no penalty (D=None), 1 = D_dev[P][K] given, 2 = D computed on the fly from pos_dev[K][3] on the
 * X*Y*Z = P grid, D = 1 - exp(-0.01 |p - pos_k|) (Demix/dNMF.py:133-135).  scratch_KK_dev: K*K doubles. */
int dnmf_update_spatial(const double* A_dev, const double* C_dev, const double* Yi_dev, const double* D_dev,
                        const float* pos_dev, int gX, int gY, int gZ, double gamma, int use_D, int64_t P, int K,
                        int T, double* scratch_KK_dev, double* out_dev, void* stream);

/* The static DeformableNMF.update_temporal on DENSE arrays (Demix/dNMF.py:139-149), fp64: A_t[P][K][T] (the
 * reference's [X,Y,Z,K,T] array), C[K][T], Y[P][T] -> out[K][T].  scratch_dev: (K*K + K) * T doubles. */
int dnmf_update_temporal_dense(const double* At_dev, const double* C_dev, const double* Y_dev, double gamma,
                               int use_gamma, int64_t P, int K, int T, double* scratch_dev, double* out_dev,
                               void* stream);

/* Max-projection along z of the deformed footprints (demo.py:50-52, A_t.max(2)) for B frames, straight from the
 * per-axis tables: out[B][K][X][Y]; the dense A_t[B][K][X][Y][Z] is never formed.  Z <= 64. */
int dnmf_forward_maxz(dnmf_ctx* ctx, const int32_t* frame_ids_dev, int B, const float* beta_dev, float* out_dev,
                      void* stream);
/* The same for frames (Y.max(2), Y_i.max(2)): frames_dev[columns][Z] -> out_dev[columns]. */
int dnmf_frames_maxz(const float* frames_dev, int64_t columns, int Z, float* out_dev, void* stream);

/* Waits for `stream` and reports an error that an earlier asynchronous call found on the device (see the
 * conventions above); returns 0 when there is none.  The error is cleared by reporting it. */
int dnmf_check_status(dnmf_ctx* ctx, void* stream);

/* Counters for bench accounting: number of fused-kernel launches etc. since creation. */
int dnmf_get_counters(dnmf_ctx* ctx, int64_t* out /* [8] */);

#ifdef __cplusplus
}
#endif
#endif /* DNMF_B200_H */
