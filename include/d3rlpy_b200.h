/*
 * d3rlpy_b200 — C ABI of the B200-native offline-RL update path.
 *
 * Plain pointers and sizes only: every pointer is a DEVICE pointer unless its name
 * says `pinned`/`host`; `stream` is a cudaStream_t passed as void*.  All calls are
 * asynchronous on `stream`, allocate nothing, and return 0 (D3B_OK) or a negative
 * code with text available from d3b_last_error().  No exceptions, no aborts, no
 * CPU fallback.  Row-major fp32 everywhere; `ld*` are leading dimensions in
 * elements; `stride_*` are per-ensemble-member strides in elements (0 = operand
 * shared by all members).
 *
 * The reference (thanhkaist/d3rlpy 1.1.0) has no FFI for this path — its boundary
 * is a Python class API — so each entry point cites the reference code it
 * replaces (paths relative to the reference root).
 */
#ifndef D3RLPY_B200_H
#define D3RLPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3B_ABI_VERSION 1
#define D3B_OK 0
#define D3B_ERR_ARG (-1)     /* bad shape / null pointer / unsupported configuration */
#define D3B_ERR_CUDA (-2)    /* CUDA runtime error (text in d3b_last_error) */

/* ---- library ------------------------------------------------------------------ */
const char* d3b_last_error(void);
int d3b_abi_version(void);
int64_t d3b_launch_count(void); /* kernels + memset/copy nodes issued by this library so far */
int d3b_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);
int d3b_set_pdl(int enabled); /* programmatic dependent launch between the update's kernels (default on) */

/* ---- K1/K1b: replay-buffer gather ---------------------------------------------
 * Replaces TransitionMiniBatch.__cinit__/_assign_to_batch/_assign_observation/
 * _assign_action (d3rlpy/dataset.pyx:1139-1342) and _stack_frames
 * (d3rlpy/dataset.pyx:1051-1096); optional fused StandardScaler.transform
 * (d3rlpy/preprocessing/scalers.py:350-354) when scaler_mean/std are non-NULL.
 * `meta` is int32[T][4] = {step, episode_start_step, episode_last_step, flags}
 * per transition (flags bit 0 = terminal, bit 1 = next_observation is the all-zero dummy: both set on the last
 * transition of a terminal episode, dataset.pyx:86-96; the online ReplayBuffer sets bit 0 alone on the transition
 * into the terminal state, online/buffers.py:283-300); `indices` int64[batch] are transition indices. */
int d3b_gather_vector(const float* obs, int obs_dim, const void* actions, int act_dim, int discrete,
                      const float* rewards, const void* meta, const int64_t* indices, int batch, int n_steps,
                      float gamma, float* out_obs, void* out_act, float* out_rew, float* out_next, float* out_term,
                      float* out_nsteps, const float* scaler_mean, const float* scaler_std, float scaler_eps,
                      void* stream);
int d3b_gather_frames(const uint8_t* frames, int frame_bytes, const void* meta, const int64_t* indices, int batch,
                      int n_frames, int n_steps, uint8_t* out_obs, uint8_t* out_next, void* stream);

/* StandardScaler.transform (d3rlpy/preprocessing/scalers.py:350-354) for a host-staged batch: x = (x - mean) / (std + eps).
 * MinMaxScaler.transform (scalers.py:209-218) is the same call with mean = min, std = max - min (float32), eps = 0. */
int d3b_standardize(float* x, const float* mean, const float* std, float eps, int rows, int dim, void* stream);

/* MinMaxActionScaler.transform / reverse_transform (d3rlpy/preprocessing/action_scalers.py:185-206), in place on
 * a[rows][dim]:  scale   a = ((a - min) / (max - min)) * 2 - 1      (TorchMiniBatch, torch_utility.py:182-183)
 *                unscale a = ((max - min) * ((a + 1) / 2)) + min    (predict_best_action / sample_action,
 *                                                                    algos/torch/base.py:60-62,77-79) */
int d3b_scale_actions(float* a, const float* minimum, const float* maximum, int rows, int dim, void* stream);
int d3b_unscale_actions(float* a, const float* minimum, const float* maximum, int rows, int dim, void* stream);

/* RewardScaler.transform (d3rlpy/preprocessing/reward_scalers.py:125-126,176-177,261-264,356-359,470-472) in place on
 * r[n]:  r = (mul * (clamp(r, lo, hi) - sub)) / div  -- Multiply (mul), Clip (lo, hi, mul), MinMax (sub = min,
 * div = max - min), Standard (sub = mean, div = std + eps), ReturnBased (div = return_max - return_min); unused
 * constants are -inf / +inf / 0 / 1 / 1, which leave the value bit-identical. */
int d3b_scale_rewards(float* r, int n, float lo, float hi, float sub, float mul, float div, void* stream);

/* ---- K2/K3: batched-ensemble dense layers (fp32 mode) ---------------------------
 * forward:  y[e] = act(x[e] w[e]^T + b[e])      replaces nn.Linear + ReLU in
 *           _VectorEncoder._fc_encode (d3rlpy/models/torch/encoders.py:265-275) and the
 *           Python loop over members (q_functions/ensemble_q_function.py:144-146,168-170).
 * backward_data:   dx[e] = (dy[e] w[e]) * [relu_src[e] > 0]      (autograd of the above)
 * backward_weight: dw[e] += dy[e]^T x[e]; dbias[e] += colsum(dy[e])  (accumulates with RED) */
int d3b_linear_forward(const float* x, int64_t ldx, int64_t stride_x, const float* w, int64_t ldw, int64_t stride_w,
                       const float* bias, int64_t stride_b, float* y, int64_t ldy, int64_t stride_y, int rows,
                       int out_features, int in_features, int members, int relu, void* stream);
int d3b_linear_backward_data(const float* dy, int64_t lddy, int64_t stride_dy, const float* w, int64_t ldw,
                             int64_t stride_w, float* dx, int64_t lddx, int64_t stride_dx, const float* relu_src,
                             int64_t ld_src, int64_t stride_src, int rows, int out_features, int in_features,
                             int members, void* stream);
int d3b_linear_backward_weight(const float* dy, int64_t lddy, int64_t stride_dy, const float* x, int64_t ldx,
                               int64_t stride_x, float* dw, int64_t lddw, int64_t stride_dw, float* dbias,
                               int64_t stride_db, int rows, int out_features, int in_features, int members,
                               void* stream);

/* fp32-mode engine of the three calls above: 1 (default) = error-compensated TF32 ("3xTF32") on the tcgen05 tensor
 * cores (csrc/gemm_tf32x3.cu: every fp32 operand is split x = hi + lo in registers, three kind::tf32 MMAs per step,
 * fp32 accumulate; fp32-grade accuracy, so the 1e-5 parity of fp32 mode holds), 0 = SIMT FFMA GEMMs.  Layers with
 * fewer than 64 rows always run on the SIMT kernels.  Environment override at first use: D3B_FP32_ENGINE=simt. */
int d3b_set_fp32_engine(int engine);
int d3b_get_fp32_engine(void);
/* The 3xTF32 GEMM itself: c[e] (m x n) (+)= sum_r a(m, r) b(n, r), fp32 row-major operands of any leading dimension.
 * a_rc: a stored [m][r] (else [r][m]); b_rc: b stored [n][r] (else [r][n]) — forward (1,1), data gradient (1,0),
 * weight gradient (0,0).  Epilogue: + bias[n], ReLU, keep where mask[m][n] > 0, store or RED.ADD (`atomic`, needed
 * for splits > 1 over r); colsum[e][m] += sum_r a(m, r) (bias gradient; weight-gradient form only). */
int d3b_tc32_gemm(const float* a, int64_t lda, int64_t stride_a, int a_rc, const float* b, int64_t ldb,
                  int64_t stride_b, int b_rc, float* c, int64_t ldc, int64_t stride_c, int m, int n, int r,
                  int members, int splits, const float* bias, int64_t stride_bias, int relu, const float* mask,
                  int64_t ld_mask, int64_t stride_mask, float* colsum, int64_t stride_colsum, int atomic,
                  void* stream);

/* Narrow heads (out_features <= 32): Q head, mu|logstd, VAE heads, discrete Q head.
 * Replace `_fc/_mu/_logstd` (q_functions/mean_q_function.py:21,69; policies.py:55,92,153-158;
 * imitators.py:45-54). act_tanh fuses DeterministicPolicy's tanh (policies.py:57-59). */
int d3b_head_forward(const float* x, int64_t ldx, int64_t stride_x, const float* w, int64_t ldw, int64_t stride_w,
                     const float* bias, int64_t stride_b, float* y, int64_t ldy, int64_t stride_y, int rows,
                     int out_features, int in_features, int members, int act_tanh, void* stream);
int d3b_head_backward_data(const float* dy, int64_t lddy, int64_t stride_dy, const float* w, int64_t ldw,
                           int64_t stride_w, float* dx, int64_t lddx, int64_t stride_dx, const float* relu_src,
                           int64_t ld_src, int64_t stride_src, int rows, int out_features, int in_features,
                           int members, void* stream);
int d3b_head_backward_weight(const float* dy, int64_t lddy, int64_t stride_dy, const float* x, int64_t ldx,
                             int64_t stride_x, float* dw, int64_t lddw, int64_t stride_dw, float* dbias,
                             int64_t stride_db, int rows, int out_features, int in_features, int members,
                             void* stream);

/* ---- K2/K3 (bf16 mode): tcgen05 tensor-core dense layers ----------------------------
 * umma_gemm: c[e] (m x n) = a[e] (m x k) * b[e] (n x k)^T with bf16 K-major operands staged by TMA and
 *   fp32 accumulation in TMEM; stride_* == 0 shares the operand across members.  Forward, dgrad and
 *   wgrad of nn.Linear (encoders.py:265-275, every ensemble member of ensemble_q_function.py:144-170 in
 *   one launch) are all expressed in this form over K-major shadows.  Optional fused epilogue: +bias,
 *   ReLU, ReLU-mask (keep where mask>0), bf16 output, transposed bf16 output, fp32 store / RED.ADD
 *   (split-K over `splits` when accumulating weight gradients).
 * shadow_weights / to_bf16: fp32 -> bf16 (+ transposed) shadows of parameters and first-layer inputs.
 * head_*_bf16 / colsum_bf16: narrow heads and bias gradients over bf16 activations. */
int d3b_umma_gemm(const void* a, int64_t lda, int64_t stride_a, const void* b, int64_t ldb, int64_t stride_b, int m,
                  int n, int k, int members, int splits, const float* bias, int64_t stride_bias, int relu,
                  const void* mask, int64_t ld_mask, int64_t stride_mask, void* out_bf16, int64_t ldo,
                  int64_t stride_o, void* out_t_bf16, int64_t ldt, int64_t stride_t, float* out_f32, int64_t ldf,
                  int64_t stride_f, int atomic, void* stream);
/* profiling hook: 0 switches off the cluster split-K latency configuration umma_gemm picks for launches far below
 * one wave (128 x 64 tiles, the reduction split over a thread-block cluster, partial tiles summed through distributed
 * shared memory) */
int d3b_umma_set_cluster(int enabled);
/* weight-gradient form c[e] (m x n) (+)= a[e]^T b[e], a [k][m] and b [k][n] row-major bf16 (MN-major UMMA
 * operands loaded by TMA straight from the saved dZ_l / H_{l-1}; no transposed copies) */
int d3b_umma_gemm_tn(const void* a, int64_t lda, int64_t stride_a, const void* b, int64_t ldb, int64_t stride_b, int m,
                     int n, int k, int members, int splits, float* out_f32, int64_t ldf, int64_t stride_f, int atomic,
                     void* stream);
/* mlp_forward_bf16: the whole ReLU trunk + narrow head of one network for all members in ONE persistent launch
 * (activations chained through shared memory / TMEM; H_l TMA-stored only when acts_host[l] != NULL).
 * Host arrays: dims_host = {K_0, N_0..N_{L-1}}; w_host/bias_host/acts_host[l] = member-0 pointers of layer l
 * (bf16 K-major weight shadows, fp32 biases, bf16 [members][rows][ld_act] outputs).  Layer widths must be
 * multiples of 16 and <= 256, n_layers <= 4, n_head <= 32 (0 = trunk only); save_rows > 0 limits the stored
 * activations to the first save_rows rows (the rest of the rows are forward-only).  Replaces encoders.py:265-339 +
 * ensemble_q_function.py:141-175 + the `_fc/_mu/_logstd` heads in one call. */
int d3b_mlp_forward_bf16(const void* x, int64_t ldx, int64_t stride_x, int rows, int members, int n_layers,
                         const int* dims_host, const void* const* w_host, const int64_t* ldw_host, int64_t stride_w,
                         const float* const* bias_host, int64_t stride_bias, void* const* acts_host,
                         const int64_t* ld_act_host, const int64_t* stride_act_host, const float* head_w,
                         const float* head_b, int64_t stride_head, int n_head, int head_tanh, float* head_out,
                         int save_rows, void* stream);
/* mlp_backward_bf16: autograd of mlp_forward_bf16 w.r.t. activations in ONE persistent launch: dZ_{L-1} from
 * d_head and the head weights, then dZ_{l-1} = (dZ_l W_l) * [H_{l-1} > 0] with W_l fed as MN-major tiles, optional
 * dX for input columns [dx_col0, dx_col0+dx_cols).  dbias_host != NULL: also store every dZ_l (operands of the
 * weight-gradient GEMMs d3b_umma_gemm_tn) and RED-add bias / head gradients into the gradient arena; with
 * d_head_bf16 != NULL (n_head > 1) a bf16 [members][rows][16] copy of d_head is written instead of reducing the
 * head's dW in-kernel, and the caller adds {d_head_bf16, H_{L-1}} as one more weight-gradient GEMM problem. */
int d3b_mlp_backward_bf16(int rows, int members, int n_layers, const int* dims_host, const void* const* w_host,
                          const int64_t* ldw_host, int64_t stride_w, const void* const* acts_host,
                          const int64_t* ld_act_host, const int64_t* stride_act_host, void* const* dz_host,
                          const int64_t* ld_dz_host, const int64_t* stride_dz_host, const float* d_head,
                          const float* head_w, int64_t stride_head, int n_head, float* const* dbias_host,
                          float* d_head_w, float* d_head_b, int64_t stride_grad, void* d_head_bf16, float* dx,
                          int64_t lddx, int64_t stride_dx, int dx_col0, int dx_cols, void* stream);
/* all weight-gradient GEMMs of one network in one launch (host arrays of length n_problems <= 5; same k rows and
 * member count; RED.ADD into the gradient arena, split-K chosen so that ~2 CTAs per SM are in flight) */
int d3b_umma_gemm_tn_batched(int n_problems, const void* const* a_host, const int64_t* lda_host,
                             const int64_t* stride_a_host, const void* const* b_host, const int64_t* ldb_host,
                             const int64_t* stride_b_host, const int* m_host, const int* n_host, int k, int members,
                             float* const* out_host, const int64_t* ldf_host, int64_t stride_f, void* stream);
int d3b_umma_set_debug(void* device_buffer);
int d3b_tc32_set_ops(int mask);  /* which layer GEMMs use the 3xTF32 engine: bit 0 forward, 1 data gradient, 2 weight gradient (default 7) */
int d3b_tc32_set_variant(int variant);  /* profiling hook: knock out loads (1) / smem stores (2) / MMAs (4) */
int d3b_tc32_set_debug(void* device_buffer);  /* profiling hook: 64 clock64 phase stamps per CTA of d3b_tc32_gemm */
int d3b_mlp_set_debug(void* device_buffer);  /* profiling hook: 16 clock64 phase stamps per CTA of mlp_forward_bf16 */ /* profiling hook: 8 clock64 phase stamps per CTA */
int d3b_shadow_weights(const float* src, int64_t src_member_stride, void* dst_bf16, int64_t dst_member_stride,
                       const int64_t* table_host, int n_entries, int members, void* stream);
int d3b_to_bf16(const float* src, int64_t lds, int rows, int cols, void* dst, int64_t ldd, void* dst_t, int64_t ldt,
                void* stream);
int d3b_head_forward_bf16(const void* x, int64_t ldx, int64_t stride_x, const float* w, int64_t ldw,
                          int64_t stride_w, const float* bias, int64_t stride_b, float* y, int64_t ldy,
                          int64_t stride_y, int rows, int out_features, int in_features, int members, int act_tanh,
                          void* stream);
int d3b_head_backward_data_bf16(const float* dy, int64_t lddy, int64_t stride_dy, const float* w, int64_t ldw,
                                int64_t stride_w, void* dx, int64_t lddx, int64_t stride_dx, void* dx_t, int64_t ldt,
                                int64_t stride_dxt, const void* relu_src, int64_t ld_src, int64_t stride_src,
                                int rows, int out_features, int in_features, int members, void* stream);
int d3b_head_backward_weight_bf16(const float* dy, int64_t lddy, int64_t stride_dy, const void* x, int64_t ldx,
                                  int64_t stride_x, float* dw, int64_t lddw, int64_t stride_dw, float* dbias,
                                  int64_t stride_db, int rows, int out_features, int in_features, int members,
                                  void* stream);
int d3b_colsum_bf16(const void* dz, int64_t ld, int64_t stride_z, float* dbias, int64_t stride_db, int rows,
                    int cols, int members, void* stream);

/* The callable Q-function API.  ensemble_reduce: _reduce_ensemble (q_functions/ensemble_q_function.py:9-24) over the
 * member axis of q[members][n] -> out[n]; mode 0 = min, 1 = max, 2 = mean, 3 = mix (lam * min + (1 - lam) * max).
 * td_error: EnsembleQFunction.compute_error (ensemble_q_function.py:81-106) — out[0] = sum_members mean_b loss(q_e[b] -
 * (r[b] + gamma * target[b] * (1 - terminal[b]))), squared error (mean_q_function.py:74-87) or Huber with beta 1
 * (utility.py:27-32); gamma_rows (one gamma per row, e.g. gamma ** n_steps) overrides the scalar when non-NULL. */
int d3b_ensemble_reduce(const float* q, int64_t stride_member, int n, int members, int mode, float lam, float* out,
                        void* stream);
int d3b_td_error(const float* q, int64_t stride_member, const float* rewards, const float* target,
                 const float* terminals, const float* gamma_rows, float gamma, int n, int members, int huber,
                 float* out, void* stream);

/* ---- advantage-weighted actor steps over a non-squashed Gaussian policy (AWAC, CRR; csrc/awr.cu) -----------------
 * dist = Normal(tanh(mu), exp(logstd)) (policies.py:160-181, distributions.py:33-88); logstd is the sigmoid-squashed
 * parameter logstd_param[act_dim] when non-NULL (AWAC, policies.py:248-253), else the clamped head columns
 * [act_dim, 2 act_dim) of `head` (CRR).
 * gauss_policy_rows: x[b*n + k] = [obs_b | clamp(loc_b + scale * eps[k][b], -1, 1)], eps laid out (n, batch, act_dim).
 * awr_weights: advantage = reduce_members(q_data)[b] - reduce_n(reduce_members(q_samples)[b*n + k]);
 *   member_reduce 0 = min (AWAC, awac_impl.py:126-148), 1 = mean (CRR, crr_impl.py:104-141); value_reduce 0 = mean,
 *   1 = max over the n samples; weight_mode 0 = softmax over the batch of adv / temperature, times batch
 *   (awac_impl.py:150-154), 1 = clamp(exp(adv / temperature), 0, max_weight), 2 = [adv > 0] (crr_impl.py:94-102).
 * gauss_wll_loss: loss = -scale * sum_b weights[b] * log pi(a_b | s_b) (awac_impl.py:103-116 with scale 1,
 *   crr_impl.py:82-92 with scale 1 / batch), d loss / d head (mu columns; logstd columns for the head form) and the
 *   accumulated gradient of the logstd parameter; metric_mean_std (optional) = mean exp(logstd parameter). */
int d3b_gauss_mean_std(const float* logstd_param, float min_logstd, float max_logstd, int act_dim, float* out,
                       void* stream);  /* mean_j exp(squashed logstd parameter): AWAC's `mean_std` metric (awac_impl.py:97-99) */
int d3b_gauss_policy_rows(const float* head, int64_t ld_head, const float* logstd_param, float min_logstd,
                          float max_logstd, const float* eps, const float* obs, int64_t ld_obs, float* x, int64_t ldx,
                          int batch, int n, int obs_dim, int act_dim, void* stream);
int d3b_awr_weights(const float* q_data, int64_t stride_q_data, const float* q_samples, int64_t stride_q_samples,
                    int members, int batch, int n, int member_reduce, int value_reduce, int weight_mode,
                    float temperature, float max_weight, float* weights, void* stream);
int d3b_gauss_wll_loss(const float* head, int64_t ld_head, const float* logstd_param, const float* actions,
                       int64_t ld_act, const float* weights, float min_logstd, float max_logstd, float scale,
                       float* d_head, int64_t ld_dhead, float* dlogstd_param, float* metric_loss,
                       float* metric_mean_std, int batch, int act_dim, void* stream);

/* PLAS glue (plas_impl.py:138-168): scaled_concat_rows x[b] = [obs_b | scale * z_b] (the latent action 2 * pi(s) next
 * to the observation, the decoder's input); tanh_backward out = scale * dy * (1 - y^2) for y = tanh(pre). */
int d3b_scaled_concat_rows(const float* obs, int64_t ldo, const float* z, int64_t ldz, float scale, float* x, int64_t ldx,
                           int batch, int obs_dim, int z_dim, void* stream);
int d3b_tanh_backward(const float* dy, int64_t lddy, const float* y, int64_t ldy, float scale, float* out, int64_t ldo,
                      int rows, int cols, void* stream);

/* ---- BEAR (csrc/bear.cu; d3rlpy/algos/torch/bear_impl.py) ----------------------------------------------------------
 * bear_latent_rows: x[k*batch + b] = [obs_b | clamp(latent[k*batch + b], +-clip)], the decoder input of
 *   ConditionalVAE.sample_n_without_squash (imitators.py:94-118; rows in sample-major order).
 * bear_mmd: per observation the MMD (bear_impl.py:233-281) between the n raw policy samples mu + exp(clamp(logstd)) *
 *   eps[k][b] (head = mu | logstd) and the n raw decoder outputs behavior_raw[k*batch + b]; Laplacian or Gaussian kernel
 *   (bear_impl.py:27-38); sum_out[0] += sum_b (sqrt(mmd_b + 1e-6) - threshold); when d_head is non-NULL the gradient of
 *   exp(log_alpha) * inv_batch * sum_b mmd_b w.r.t. the policy head is ADDED to d_head.
 * bear_alpha_step: update_alpha (bear_impl.py:215-231) from that sum: loss, Adam on log_alpha ({p,g,m,v} at float
 *   offsets 0,4,8,12), clamp to [-5, 10], metrics.  bear_actor_metric: [SAC actor loss] + exp(log_alpha) * mean(mmd - thr).
 * bear_target: compute_target (bear_impl.py:283-303) from q[members][batch*n] and logp[batch*n]. */
int d3b_bear_latent_rows(const float* obs, int64_t ldo, const float* latent, float clip, float* x, int64_t ldx, int batch,
                         int n, int obs_dim, int latent_dim, void* stream);
int d3b_bear_mmd(const float* head, int64_t ld_head, const float* eps, const float* behavior_raw, int64_t ld_behavior,
                 int gaussian_kernel, float sigma, float min_logstd, float max_logstd, const float* log_alpha,
                 float threshold, float inv_batch, float* d_head, int64_t ld_dhead, float* sum_out, int batch, int n,
                 int act_dim, void* stream);
int d3b_bear_alpha_step(const float* mmd_sum, float* alpha_scalar, const int* step, double lr, float inv_batch,
                        float* metric_loss, float* metric_alpha, void* stream);
int d3b_bear_actor_metric(const float* sac_loss, const float* mmd_sum, const float* log_alpha, float inv_batch,
                          float* metric, void* stream);
int d3b_bear_target(const float* q, int64_t stride_q, const float* logp, const float* log_temp, float lam, float* q_tpn,
                    int batch, int n, int members, void* stream);

/* ---- K4-K7: row assembly, sampling, losses ---------------------------------------
 * concat_rows: x[b*n+k] = [obs[b] | f(act[b*n+k])]  — torch.cat([x, action]) of
 *   VectorEncoderWithAction.forward (encoders.py:328-339) plus the repeat/transpose/reshape of
 *   cql_impl.py:153-161,176-180, bcq_impl.py:163-168; f = TD3 target smoothing
 *   (td3_impl.py:66-73) when `noise` != NULL, or clamp(+-act_clip) (bcq_impl.py:141,181). */
int d3b_concat_rows(const float* obs, int64_t ldo, const float* act, int64_t lda, const float* noise, float sigma,
                    float noise_clip, float act_clip, float* x, int64_t ldx, int batch, int n_repeat, int obs_dim,
                    int act_dim, void* stream);
int d3b_concat_rows_bf16(const float* obs, int64_t ldo, const float* act, int64_t lda, const float* noise, float sigma,
                         float noise_clip, float act_clip, void* x_bf16, int64_t ldx, int batch, int n_repeat,
                         int obs_dim, int act_dim, void* stream); /* same rows as bf16 GEMM operands */
/* policy_sample_rows: SquashedNormalPolicy sample_with_log_prob / sample_n_with_log_prob /
 *   best_action (policies.py:167-249, distributions.py:91-143).  head = [mu | raw logstd],
 *   eps laid out [n][batch][act] as Normal.rsample((n,)). */
int d3b_policy_sample_rows(const float* head, int64_t ld_head, const float* eps, const float* obs, int64_t ldo,
                           float* x, int64_t ldx, float* act_out, float* logp, int batch, int n_samples, int obs_dim,
                           int act_dim, float min_logstd, float max_logstd, int deterministic, void* stream);
/* critic_loss: TD error summed over members (ddpg_impl.py:154-165, ensemble_q_function.py:81-106,
 *   mean_q_function.py:74-87) + CQL conservative term with its softmax gradient
 *   (cql_impl.py:196-223); q rows = [data B | pi(s_t) B*N | pi(s_t+1) B*N | random B*N].
 *   sums[0..2] += {sum (q-y)^2, sum logsumexp, sum q_data}. */
int d3b_critic_loss(const float* q, int64_t stride_q, const float* q_targ, int64_t stride_qt, int targ_members,
                    const float* q_tpn, const float* rewards, const float* terminals, const float* n_steps,
                    float gamma, const float* logp_t, const float* logp_tp1, int n_action_samples, int act_dim,
                    const float* log_alpha, float conservative_weight, float* dq, int64_t stride_dq, float* sums,
                    float* y_out, int batch, int members, float inv_batch, int td_enabled, void* stream);
/* cql_finalize: scalar tail of compute_critic_loss (mode 0) / update_alpha (mode 1)
 *   (cql_impl.py:110-141,217-223). */
int d3b_cql_finalize(const float* sums, const float* log_alpha, float inv_batch, int members,
                     float conservative_weight, float alpha_threshold, int mode, int conservative, float* metric,
                     float* grad_log_alpha, void* stream);
int d3b_scalar_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, const int* step, double lr,
                    double beta1, double beta2, double eps, float* out_exp, void* stream);
/* SACImpl.compute_actor_loss / update_temp (sac_impl.py:114-146) */
int d3b_sac_actor_loss(const float* q, int64_t stride_q, const float* logp, const float* log_temp, float* dq,
                       int64_t stride_dq, float* loss_sum, int batch, int members, float inv_batch, void* stream);
int d3b_sac_actor_backward(const float* head, int64_t ld_head, const float* eps, const float* dx_action,
                           int64_t lddx, int64_t stride_dx, int members, const float* log_temp, float* dhead,
                           int64_t ld_dhead, int batch, int act_dim, float min_logstd, float max_logstd,
                           float inv_batch, void* stream);
int d3b_sac_temp_loss(const float* logp, const float* log_temp, int batch, int act_dim, float inv_batch,
                      float* metric, float* grad, int accumulate, void* stream);
/* SACImpl.compute_target, soft backup (sac_impl.py:148-162; CQL soft_q_backup=True, cql_impl.py:225-231) */
int d3b_sac_soft_backup(const float* q_targ, int64_t stride_q, int members, const float* logp, const float* log_temp,
                        float* q_tpn, int batch, void* stream);
/* ---- fused glue kernels of the CQL/SAC update (single-GPU bf16 path; csrc/cql_fused.cu) -------------------
 * begin_step: counters[i] += 1 for the bits of mask, loss partial sums zeroed.
 * cql_rows: all critic input rows of one update (critic-step and alpha-step importance-sampling groups
 *   [data | pi(s_t) | pi(s_t+1) | random], target row, actor row) as bf16 GEMM operands + every tanh-Gaussian
 *   log-prob (cql_impl.py:143-204, policies.py:167-249, distributions.py:91-143).  ptrs_host / rows_host: see .cu.
 * sac_temp_step: update_temp loss + gradient + Adam (sac_impl.py:123-146); scalar = {p,g,m,v} at float stride 4.
 * cql_loss_step: conservative + TD loss and gradient seed, then the scalar tail by the last block: mode 0 critic
 *   metric (cql_impl.py:110-117), mode 1 alpha loss + Adam on log_alpha (cql_impl.py:119-141).
 * sac_actor_step: SACImpl.compute_actor_loss (sac_impl.py:114-121) + gradient seed + metric.
 * done_counter (cql_loss_step, sac_actor_step): zero-initialised workspace of 32-bit words: word 0 = self-resetting block
 *   counter, words [4, 4 + 3 * ceil(batch * members / 8)) (cql_loss_step) or [4, 4 + ceil(batch / 256)) (sac_actor_step) =
 *   per-block partial sums; the last block adds them in a fixed order, so `sums` and the metrics are bit-reproducible. */
int d3b_begin_step(int* counters, int n, unsigned mask, float* slots, int n_slots, void* stream);
/* begin_step + noise_fill (epoch = counters[draw_index] after its bump) + optional fp32 -> bf16 conversion of the
 * policy input rows, in ONE launch; done_counter: one zero-initialised 32-bit word (self-resetting).  The counters are
 * bumped by the last block to finish, so the launch needs no grid barrier. */
int d3b_update_prologue(int* counters, int n_counters, unsigned mask, int draw_index, float* slots, int n_slots,
                        float* noise, int64_t n_normal, int64_t n_uniform, uint64_t seed, const float* src,
                        int64_t lds, int rows, int cols, void* dst_bf16, int64_t ldd, void* done_counter,
                        void* stream);
int d3b_cql_rows(const float* head, const float* obs, const float* next_obs, const float* act, int batch,
                 int n_action_samples, int obs_dim, int act_dim, float min_logstd, float max_logstd, void* x_bf16,
                 int64_t ldx, int n_groups, const void* const* ptrs_host, const int64_t* rows_host, void* stream);
/* cql_rows with fp32 output rows (fp32 mode) */
int d3b_cql_rows_f32(const float* head, const float* obs, const float* next_obs, const float* act, int batch,
                     int n_action_samples, int obs_dim, int act_dim, float min_logstd, float max_logstd, float* x,
                     int64_t ldx, int n_groups, const void* const* ptrs_host, const int64_t* rows_host, void* stream);
int d3b_sac_temp_step(const float* logp, float* scalar, const int* step, int batch, int act_dim, float inv_batch,
                      double lr, float* metric_loss, float* metric_exp, void* stream);
int d3b_cql_loss_step(const float* q, int64_t stride_q, const float* q_targ, int64_t stride_qt, int targ_members,
                      const float* q_tpn, const float* rewards, const float* terminals, const float* n_steps,
                      float gamma, const float* logp_t, const float* logp_tp1, int n_action_samples, int act_dim,
                      float* scalar_alpha, float conservative_weight, float alpha_threshold, float* dq,
                      int64_t stride_dq, float* sums, void* done_counter, int batch, int members, float inv_batch,
                      int mode, const int* step_alpha, double lr_alpha, float* metric, float* metric_exp,
                      void* stream);
int d3b_sac_actor_step(const float* q, int64_t stride_q, const float* logp, const float* log_temp, float* dq,
                       int64_t stride_dq, float* loss_sum, void* done_counter, float* metric, int batch, int members,
                       float inv_batch, void* stream);

/* TD3PlusBCImpl.compute_actor_loss (td3_plus_bc_impl.py:64-70) in three phases so the
 * batch-global lambda can be all-reduced between stats and seed when the batch is sharded. */
int d3b_td3bc_actor_stats(const float* q0, const float* a, int64_t lda, const float* a_data, int64_t ldd,
                          float* sums, int batch, int act_dim, void* stream);
int d3b_td3bc_actor_seed(const float* sums, float alpha, float inv_batch, int act_dim, float* dq, int64_t stride_dq,
                         int batch, int members, float* metric, void* stream);
int d3b_td3bc_actor_backward(const float* a, int64_t lda, const float* a_data, int64_t ldd, const float* dx_action,
                             int64_t lddx, float* dz, int64_t lddz, int batch, int act_dim, float inv_batch,
                             void* stream);

/* ---- K8: BCQ (imitators.py:63-86, bcq_impl.py:115-226, policies.py:94-97, q_functions/__init__.py:8-63) --
 * vae_sample_rows: z = mu + exp(clamp(logstd)) eps, decoder rows [obs | z], KL(N(mu,sd)||N(0,1)) sum.
 * vae_recon / vae_backward / vae_finalize: MSE + beta*KL loss of ConditionalVAE.compute_error and its gradient.
 * residual_rows / residual_backward: a = clamp(sampled + scale*tanh(z), -1, 1) of DeterministicResidualPolicy.
 * bcq_target_reduce: max over sampled actions of (1-lam) max_e Q + lam min_e Q.
 * neg_mean_seed: actor loss -mean(Q_0) and its seed gradient. */
int d3b_vae_sample_rows(const float* head, int64_t ld_head, const float* eps, const float* obs, int64_t ldo, float* x,
                        int64_t ldx, float* kl_sum, int batch, int obs_dim, int latent, float min_logstd,
                        float max_logstd, void* stream);
int d3b_vae_recon(const float* y, const float* actions, int64_t lda, float* dpre, float* sq_sum, int batch,
                  int act_dim, float inv_batch, void* stream);
int d3b_vae_backward(const float* head, int64_t ld_head, const float* eps, const float* dz, int64_t lddz,
                     float* dhead, int64_t ld_dhead, int batch, int latent, float min_logstd, float max_logstd,
                     float beta, float inv_batch, void* stream);
int d3b_vae_finalize(const float* sums, int act_dim, int latent, float beta, float inv_batch, float* metric,
                     void* stream);
int d3b_residual_rows(const float* z, int64_t ldz, const float* sampled, int64_t lds, const float* obs, int64_t ldo,
                      float* x, int64_t ldx, float scale, int rows, int n_repeat, int obs_dim, int act_dim,
                      void* stream);
int d3b_residual_backward(const float* z, int64_t ldz, const float* sampled, int64_t lds, const float* da,
                          int64_t ldda, float* dz, int64_t lddz, float scale, int batch, int act_dim, void* stream);
int d3b_bcq_target_reduce(const float* q, int64_t stride_q, float* q_tpn, int batch, int n_actions, int members,
                          float lam, void* stream);
int d3b_neg_mean_seed(const float* q0, float* dq, float* loss_sum, int batch, float inv_batch, void* stream);
/* ---- K9 (loss side): DQN/DoubleDQN target and DiscreteCQL loss (dqn_impl.py:97-171, cql_impl.py:279-302,
 * mean_q_function.py:26-42, utility.py:27-32) */
int d3b_dqn_target(const float* q_online, int64_t stride_qo, const float* q_targ, int64_t stride_qt, float* q_tpn,
                   int batch, int n_actions, int members, void* stream);
int d3b_dcql_loss(const float* q, int64_t stride_q, const float* q_tpn, const float* actions, const float* rewards,
                  const float* terminals, const float* n_steps, float gamma, float alpha, float* dq,
                  int64_t stride_dq, float* sums, int batch, int n_actions, int members, float inv_batch,
                  int conservative, void* stream);
int d3b_dcql_finalize(const float* sums, float inv_batch, float alpha, int conservative, float* metric, void* stream);
/* Quantile-regression Q head (QRQFunctionFactory) of the same algorithms.  theta = the head's outputs
 * [members][batch][n_actions][n_quantiles] (DiscreteQRQFunction._compute_quantiles, qr_q_function.py:38-42).
 * qr_target: greedy action of mean_e mean_i theta_select (online net: DoubleDQNImpl.compute_target, dqn_impl.py:162-171;
 *   target net: DQNImpl.compute_target, :133-141), then the quantiles [batch][n_quantiles] of the member whose mean is
 *   smallest (pick_quantile_value_by_action utility.py:17-24 + _reduce_quantile_ensemble "min",
 *   ensemble_q_function.py:47-52).
 * qr_loss: quantile Huber loss with fixed mid-point taus (qr_q_function.py:15-19,50-78; utility.py:35-61) summed over
 *   members (ensemble_q_function.py:81-106) [+ the DiscreteCQL term on the quantile means, cql_impl.py:290-302] and the
 *   gradient w.r.t. every theta; sums[0] += sum_e sum_b L, sums[1] += sum_b (logsumexp - data), accumulated in a fixed
 *   order through the workspace partials[2 * batch] (bit-reproducible); dcql_finalize turns the sums into the metric.
 * qr_values: values[e][b][a] = mean_i theta (DiscreteQRQFunction.forward, qr_q_function.py:44-48). */
int d3b_qr_target(const float* theta_select, int64_t stride_select, const float* theta_targ, int64_t stride_targ,
                  float* q_tpn, int batch, int n_actions, int n_quantiles, int members, void* stream);
int d3b_qr_loss(const float* theta, int64_t stride_theta, const float* q_tpn, const float* actions,
                const float* rewards, const float* terminals, const float* n_steps, float gamma, float alpha,
                float* dtheta, int64_t stride_dtheta, float* partials, float* sums, int batch, int n_actions,
                int n_quantiles, int members, float inv_batch, int conservative, void* stream);
int d3b_qr_values(const float* theta, int64_t stride_theta, float* values, int64_t stride_values, int batch,
                  int n_actions, int n_quantiles, int members, void* stream);
/* backward of qr_values: dtheta[v][i] = dvalues[v] / n_quantiles (the actor losses differentiate through the mean of a
 * ContinuousQRQFunction's quantiles, qr_q_function.py:118-122).  The continuous QR critics reuse qr_target / qr_loss /
 * qr_values with n_actions = 1 (ContinuousQRQFunction, qr_q_function.py:91-165). */
int d3b_qr_values_backward(const float* dvalues, float* dtheta, int64_t n_values, int n_quantiles, void* stream);

/* ---- IQL (sibling algorithm on the same building blocks; d3rlpy/algos/torch/iql_impl.py:109-141).
 * iql_value_loss: expectile regression mean_b |expectile - 1[d < 0]| d^2, d = min_e Q'_e(s,a) - V(s); writes dL/dV and
 *   the metric.
 * iql_actor_loss: -mean_b w_b log N(a_b; tanh(mu_b), exp(logstd)) with w_b = min(exp(weight_temp (min_e Q'_e - V)),
 *   max_weight) and logstd = min + sigmoid(param) (max - min) (policies.py:168-181,248-253); writes dL/dmu (pre-tanh),
 *   ACCUMULATES dL/dparam into dlogstd[act_dim], writes the metric.  One block each, fixed summation order. */
int d3b_iql_value_loss(const float* q_targ, int64_t stride_q, int members, const float* v, float expectile,
                       float inv_batch, float* dv, float* metric, int batch, void* stream);
int d3b_iql_actor_loss(const float* mu, int64_t ld_mu, const float* logstd_param, const float* actions, int64_t ld_act,
                       const float* q_targ, int64_t stride_q, int members, const float* v, float weight_temp,
                       float max_weight, float min_logstd, float max_logstd, float inv_batch, float* dmu,
                       int64_t ld_dmu, float* dlogstd, float* metric, int batch, int act_dim, void* stream);

/* ---- K9 (encoder side): Nature-DQN convolutions as patch gather + the dense-layer GEMMs above
 * (PixelEncoder.forward, d3rlpy/models/torch/encoders.py:81-162; nn.Conv2d at :100).
 * im2col: patches[e][(b,oh,ow)][(ic,kh,kw)] = x[e][b][ic][oh*s+kh][ow*s+kw] / divisor with generic element
 *   strides (sb,sc,sh,sw), so the same call reads the NCHW uint8 minibatch (divisor = 255 fuses PixelScaler,
 *   scalers.py:109-110, and the float cast of torch_utility.py:146-149) and the NHWC fp32 outputs of the
 *   previous layer; the flatten + fc of encoders.py:150-162 is the ksize == height case.
 * col2im: gradient of im2col onto the NHWC input, times the ReLU mask [y_prev > 0] of its producer (bf16 mask
 *   => bf16 gradient output, the tensor-core path).  im2col x_is_u8: 0 fp32, 1 uint8, 2 bf16 input. */
int d3b_im2col(const void* x, int x_is_u8, int64_t stride_x, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
               void* out, int out_is_bf16, int64_t ldo, int64_t stride_o, int images, int channels, int height,
               int width, int ksize, int stride, float divisor, int members, void* stream);
int d3b_col2im(const float* dpatch, int64_t ldp, int64_t stride_p, const void* y_prev, int y_is_bf16, int64_t ldy,
               int64_t stride_y, void* dx, int64_t lddx, int64_t stride_dx, int images, int channels, int height,
               int width, int ksize, int stride, int members, void* stream);

/* ---- K10: optimizer / target sync ----------------------------------------------
 * adam_step: torch.optim.Adam.step as built by AdamFactory (d3rlpy/models/optimizers.py:106-138;
 *   call sites ddpg_impl.py:150,181, sac_impl.py:141, cql_impl.py:137, bcq_impl.py:159,
 *   dqn_impl.py:109) over one flat arena; `step` is a device int holding t (already
 *   incremented by d3b_tick); `target` != NULL fuses soft_sync of the new params.
 * soft_sync / hard_sync: d3rlpy/torch_utility.py:27-41. */
int d3b_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* target, int64_t n,
                  const int* step, double lr, double beta1, double beta2, double eps, float tau, int zero_grad,
                  void* stream);
/* adam_step with torch.optim.Adam's weight_decay: grad += weight_decay * param before the moments (AWAC's actor
 * optimizer, d3rlpy/algos/awac.py:105). */
int d3b_adam_step_wd(float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* target, int64_t n,
                     const int* step, double lr, double beta1, double beta2, double eps, float weight_decay, float tau,
                     int zero_grad, void* stream);
/* adam_step that also rewrites the bf16 K-major weight shadows (params and, when synced, target) in the same
 * pass; table_host: n_segments x {param_off, rows, cols, shadow_off, ld} int64 per trunk weight matrix. */
int d3b_adam_step_shadow(float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* target, int64_t n,
                         const int* step, double lr, double beta1, double beta2, double eps, float tau,
                         void* shadow_params, void* shadow_target, const int64_t* table_host, int n_segments,
                         int64_t member_size, int64_t shadow_member, void* stream);
int d3b_soft_sync(float* target, const float* params, int64_t n, float tau, void* stream);
int d3b_hard_sync(float* target, const float* params, int64_t n, void* stream);
int d3b_tick(int* counters, int n, unsigned mask, void* stream);
/* Philox4x32-10 fill: first n_normal floats ~ N(0,1), next n_uniform ~ U(-1,1); replaces
 * torch.randn / uniform_ / Normal.rsample draws (td3_impl.py:67, cql_impl.py:186,
 * distributions.py:105,118, bcq_impl.py:136,178, imitators.py:85). */
int d3b_noise_fill(float* out, int64_t n_normal, int64_t n_uniform, uint64_t seed, const int* draw_counter,
                   void* stream);

/* ---- K11: data-parallel exchange (new: the reference is single-device) -------------
 * One sum all-reduce per optimizer step over the flat gradient arena (+ the loss partial sums), on the
 * update stream, capturable into the update's CUDA graph.  NCCL is bound with dlopen at run time;
 * `id` is a 128-byte ncclUniqueId created on rank 0 and distributed by the host (torch.distributed). */
int d3b_comm_load(const char* libnccl_path);
int d3b_comm_unique_id(void* id_out_128);
int d3b_comm_init(const void* id_128, int world_size, int rank, void** comm_out);
int d3b_allreduce_sum(void* comm, float* buf, int64_t n, void* stream);
int d3b_comm_destroy(void* comm);

/* ---- K10+K11 fused over NVLink peer memory (CUDA IPC): the optimizer kernel reads every rank's gradient arena
 * itself (fixed rank order => bit-identical sums on all ranks) and applies Adam + Polyak + shadow refresh — no
 * NCCL call on the update path.  flags: per-rank int32 block, {ready, done} epoch pairs; epoch: device counter of
 * the update.  peer_export/import map a device allocation into the other ranks of the box. */
/* profiling hook: device buffer of 4 * 64 * 4 int64 receiving, per traced peer kernel (0 wait-and-zero, 1 scalar
 * steps, 2 / 3 critic / actor Adam exchange) and update, {globaltimer at entry, cycles spent waiting for peers, cycles
 * of block 0, epoch}; NULL switches it off */
int d3b_peer_set_trace(void* device_buffer);
int d3b_peer_export(const void* ptr, void* handle_out_64, int64_t* offset_out);
int d3b_peer_import(const void* handle_64, int64_t offset, void** ptr_out);
int d3b_peer_wait_zero(const void* const* flags_host, int world, int rank, int done_index, const int* epoch,
                       float* grads, int64_t n, int done_index2, float* grads2, int64_t n2, void* stream);
/* grads2 (optional): a second arena (its own done flag) waited for and zeroed by the same launch. */
/* Data-parallel CQL (cql_impl.py:119-141, sac_impl.py:128-146): all-reduce of vec = {-, sum logsumexp, sum data value,
 * temperature-loss sum} over the exchange block, then the temperature step (loss metric, Adam, exp metric; skipped when
 * temp_scalar is NULL) and the alpha step (loss, gradient, Adam, exp metric) in one launch.  *_scalar: {p,g,m,v} blocks
 * at float offsets 0,4,8,12. */
int d3b_dp_scalar_steps(float* vec, const void* const* xchg_host, const void* const* flags_host, int world, int rank,
                        int channel, const int* epoch, float* temp_scalar, const int* step_temp, double lr_temp,
                        float* metric_temp_loss, float* metric_temp, float* alpha_scalar, const int* step_alpha,
                        double lr_alpha, float inv_members_batch, float conservative_weight, float alpha_threshold,
                        float* metric_alpha_loss, float* metric_alpha, void* stream);
int d3b_peer_allreduce_small(float* vec, int n, const void* const* xchg_host, const void* const* flags_host, int world,
                             int rank, int channel, const int* epoch, void* stream);
int d3b_adam_step_peer(float* params, float* exp_avg, float* exp_avg_sq, float* target, int64_t n, const int* step,
                       double lr, double beta1, double beta2, double eps, float tau, void* shadow_params,
                       void* shadow_target, const int64_t* table_host, int n_segments, int64_t member_size,
                       int64_t shadow_member, const void* const* grads_host, const void* const* flags_host, int world,
                       int rank, int flag_index, const int* epoch, void* block_counter, float* small_vec, int small_n,
                       const void* const* xchg_host, int small_channel, const void* const* gred_host,
                       void* block_counter2, void* stream);
/* gred_host (optional): every rank's reduced-gradient buffer (n floats each, peer-mapped) — selects the TWO-SHOT
 * exchange: each rank sums its 1/W slice over all ranks and pushes it to everybody (reduce-scatter + all-gather inside
 * the kernel, flag_index + 2 = "reduced" flag, block_counter2 = a second zero-initialised uint32), then every rank
 * runs Adam from its local copy: 2 (W-1)/W arena sizes over NVLink per rank instead of W-1. */
/* small_vec (<= 16 floats, optional): loss partial sums that ride along with the gradient exchange (in: this
 * rank's values, out: all-rank sums) through exchange channel small_channel. */

/* ---- plumbing: staged copies, CUDA-graph capture of a whole update ------------------ */
int d3b_memset_zero(void* ptr, int64_t bytes, void* stream);
int d3b_copy_h2d(void* dst, const void* src_pinned, int64_t bytes, void* stream);
int d3b_copy_d2h(void* dst_pinned, const void* src, int64_t bytes, void* stream);
int d3b_copy_d2d(void* dst, const void* src, int64_t bytes, void* stream);
/* The same small copies as a KERNEL (either direction; the host side must be pinned, i.e. device-addressable under
 * unified addressing): inside the update graph a kernel node starts sooner than a copy-engine node.  Replaces the
 * `torch.tensor(array, device=...)` upload of _convert_to_torch (d3rlpy/torch_utility.py:146-149) and the
 * `loss.cpu().detach().numpy()` read-back (e.g. d3rlpy/algos/torch/ddpg_impl.py:152) for small minibatches. */
int d3b_copy_mapped(void* dst, const void* src, int64_t bytes, void* stream);
int d3b_stream_sync(void* stream);
int d3b_spin(int64_t ns, void* stream); /* measurement helper: busy-wait kernel (keeps the stream ahead of the host) */
/* fork/join a side stream (graph branches under capture): independent steps of one update run concurrently */
int d3b_stream_fork(void* main_stream, void* side_stream);
int d3b_stream_join(void* main_stream, void* side_stream);
int d3b_graph_begin(void* stream);
int d3b_graph_end(void* stream, void** graph_exec, int* n_nodes);
int d3b_graph_launch(void* graph_exec, void* stream);
int d3b_graph_destroy(void* graph_exec);

#ifdef __cplusplus
}
#endif
#endif /* D3RLPY_B200_H */
