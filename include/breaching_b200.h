/* breaching_b200 -- C ABI of the sm_100a gradient-inversion engine.
 *
 * Drop-in boundary.  The reference has no native code and no FFI of its own; the
 * interface a maintainer would bind is the body of
 *   OptimizationBasedAttacker._run_trial      breaching/attacks/optimization_based_attack.py:90-143
 *   closure of _compute_objective             breaching/attacks/optimization_based_attack.py:145-189
 *   GradientLoss.forward / _grad_fn_single_step  breaching/attacks/auxiliaries/objectives.py:26-46
 *   the *_sim / _euclidean list reductions    breaching/attacks/auxiliaries/objectives.py:91-95,135-141,160-164,185-196
 *   TotalVariation / Norm / DeepInversion / Feature regularizers   breaching/attacks/auxiliaries/regularizers.py
 *   optimizer_lookup (Adam/AdamW/SGD + LR)    breaching/attacks/auxiliaries/common.py:5-40
 *   _score_trial                              breaching/attacks/optimization_based_attack.py:191-204
 * Each entry point below names the reference lines it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: plain C, no torch types.  All functions return 0 on success or a negative bre_status;
 * bre_last_error() gives the message of the last failure on the calling thread.  Pointers marked
 * "device or host" are copied with cudaMemcpyDefault (UVA), so either works; "device" pointers must
 * be resident on the engine's GPU.  One engine per GPU, not re-entrant.
 */
#ifndef BREACHING_B200_H
#define BREACHING_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bre_engine bre_engine;

enum bre_status {
  BRE_OK = 0,
  BRE_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  BRE_ERR_CUDA = -2,        /* a CUDA runtime call failed */
  BRE_ERR_STATE = -3,       /* call order violated (e.g. run before load) */
  BRE_ERR_UNSUPPORTED = -4  /* feature not implemented by the engine (never silently falls back) */
};

/* ---- layer program (produced by breaching_b200.compiler from the nn.Module) ---------------------- */
enum bre_op_kind { BRE_OP_CONV = 1, BRE_OP_BNACT = 2, BRE_OP_MAXPOOL = 3, BRE_OP_AVGPOOL = 4, BRE_OP_LINEAR = 5,
                   /* token-sequence programs (transformer / TAG path): tensors are [rows = batch * seq_len, C, 1, 1] */
                   BRE_OP_POSADD = 6,      /* out = candidate + positional embedding `w` of position (row mod seq_len) */
                   BRE_OP_LAYERNORM = 7,   /* gamma, beta, eps */
                   BRE_OP_ATTENTION = 8 }; /* tin = fused (q | k | v) projection, R = heads, S = seq_len, no mask */
enum bre_param_perm { BRE_PERM_NONE = 0, BRE_PERM_OIHW_TO_OHWI = 1, BRE_PERM_LINEAR_CHW_TO_HWC = 2 };

typedef struct bre_tensor_desc { int32_t N, C, H, W; } bre_tensor_desc; /* tensor 0 = candidate (NCHW) */

typedef struct bre_param_desc {
  int64_t numel;
  int32_t perm;          /* bre_param_perm: engine-internal layout of this tensor */
  int32_t d0, d1, d2;    /* OIHW->OHWI: d0=O, d1=I, d2=H*W ; LINEAR: d0=out, d1=C, d2=H*W ; NONE: d0 = elements to reserve if > numel
                            (zero tail, e.g. the rows of a vocabulary padded to the GEMM tile width; see option "logits_valid") */
} bre_param_desc;

typedef struct bre_op_desc {
  int32_t kind;
  int32_t tin, tout, res;         /* tensor ids (res = -1 if none) */
  int32_t R, S, stride, pad;      /* conv / maxpool geometry */
  int32_t w, b;                   /* parameter indices in model.parameters() order, -1 = none */
  int32_t has_bn, relu;           /* BNACT: out = relu?( bn?(in) + res? ) */
  int32_t gamma, beta;            /* parameter indices of BN weight / bias */
  int32_t bn_buffer;              /* index into the running-stat arrays passed to bre_engine_load_model */
  float eps;
  int32_t acc_in, acc_res;        /* reverse sweeps: accumulate into (1) or overwrite (0) the input delta */
  int32_t bn_train;               /* BNACT: BN uses the batch statistics of its input (model in train mode without buffers,
                                     base_attack.py:192-197) instead of running statistics */
} bre_op_desc;

/* ---- attack configuration (cfg_attack of the reference, flattened) ------------------------------- */
enum bre_objective {              /* objectives.py:496-506 objective_lookup */
  BRE_OBJ_EUCLIDEAN = 0, BRE_OBJ_COSINE = 1, BRE_OBJ_L1 = 2, BRE_OBJ_TAG_EUCLIDEAN = 3,
  BRE_OBJ_ANGULAR = 4, BRE_OBJ_FAST_COSINE = 5, BRE_OBJ_MASKED_COSINE = 6
};
enum bre_optimizer { BRE_OPT_ADAM = 0, BRE_OPT_ADAMW = 1, BRE_OPT_SGD = 2 }; /* common.py:6-17 */
enum bre_sign { BRE_SIGN_NONE = 0, BRE_SIGN_HARD = 1, BRE_SIGN_SOFT = 2 };  /* optimization_based_attack.py:175-184 */

typedef struct bre_attack_cfg {
  int32_t objective;              /* bre_objective */
  float obj_scale, task_regularization, tag_scale, mask_value, angular_fudge;
  int32_t optimizer;              /* bre_optimizer */
  float beta1, beta2, adam_eps, weight_decay, momentum;
  int32_t nesterov;
  int32_t signed_mode;            /* bre_sign */
  int32_t boxed;                  /* optimization_based_attack.py:117-118 */
  int32_t max_iterations;         /* cfg.optim.max_iterations (soft-sign schedule + LR table length) */
  float langevin_noise;           /* :167-170 */
  float grad_clip;                /* :171-174 ; < 0 = disabled */
  uint64_t noise_seed;            /* Philox seed for the Langevin noise */
  /* regularizers.py -- a scale of 0 disables the term */
  float tv_scale, tv_inner_exp, tv_outer_exp, tv_eps; int32_t tv_double_opponents;
  float norm_scale, norm_p;
  float di_scale, di_first_bn_multiplier;
  float feat_scale;
  int32_t orthogonality;          /* regularizers.py:156-181 (the reference ignores its `scale`; != 0 enables the term) */
  int32_t objective_excludes_task;/* Pearlmutter* objectives (objectives.py:279-365, 468-493): task_regularization enters the
                                     candidate gradient (:362) but not the reported objective value */
} bre_attack_cfg;

/* ---- engine life cycle -------------------------------------------------------------------------- */
/* Build an engine for one model replica on `device`.  Replaces the per-trial set-up of _run_trial
 * (optimization_based_attack.py:93-107). */
int bre_engine_create(const bre_tensor_desc* tensors, int32_t n_tensors,
                      const bre_op_desc* ops, int32_t n_ops,
                      const bre_param_desc* params, int32_t n_params,
                      int32_t logits_tensor, const bre_attack_cfg* cfg, int32_t device, bre_engine** out);
void bre_engine_destroy(bre_engine* e);

/* Model state of the attacked network as the server payload holds it (base_attack.py:169-212):
 * `params[i]` = pointer to parameter i (model.parameters() order, torch-contiguous, fp32),
 * `bn_mean[j]` / `bn_var[j]` = running statistics of BN layer j (order of bre_op_desc.bn_buffer).
 * Pointers: device or host. */
int bre_engine_load_model(bre_engine* e, const float* const* params, int32_t n_params,
                          const float* const* bn_mean, const float* const* bn_var, int32_t n_bn);

/* The user's shared gradient (shared_data[i]["gradients"], users.py:176-186), same order/layout as params,
 * per-tensor TAG weights (objectives.py:115-124; NULL = all ones), labels (int64, [N]),
 * normalisation box (base_attack.py:53-57): mean/std per input channel (NULL = 0/1).
 * Pointers: device or host. */
int bre_engine_load_targets(bre_engine* e, const float* const* grads, int32_t n_params,
                            const float* tensor_weights, const int64_t* labels, int32_t n_labels,
                            const float* mean, const float* std, int32_t n_channels);

/* FedAvg / multi-step local updates (objectives.py:48-72, users.py:336-413): the user ran `steps` SGD steps of size `lr`,
 * step k on the candidate slice [k*B mod total_images, ... + B) (B = batch of the layer program) with labels
 * labels[k*B .. (k+1)*B); the matched quantity becomes W_K - W_0.  Re-sizes the candidate state to `total_images`.
 * Call after bre_engine_load_model / load_targets and before bre_engine_begin_trial.  labels: device or host. */
int bre_engine_set_local_steps(bre_engine* e, int32_t total_images, int32_t steps, float lr, const int64_t* labels);

/* Joint data / label optimisation (OptimizationJointAttacker, optimization_with_label_attack.py:145-189): the closure hands
 * `labels.softmax(dim=-1)` to the loss as class probabilities (:154).  `probabilities` [N, classes] fp32 (device or host)
 * replaces the index labels in the task loss until cleared with NULL.  After bre_engine_objective_and_gradient,
 * bre_engine_label_gradient writes d(objective)/d(probabilities) [N, classes] (the caller chains it through its softmax,
 * as autograd does for the reference at :162).  bre_engine_set_labels replaces the index labels (scoring with
 * `labels.argmax`, :67-70). */
int bre_engine_load_soft_labels(bre_engine* e, const float* probabilities, int64_t numel);
int bre_engine_label_gradient(bre_engine* e, float* grad_out);
int bre_engine_set_labels(bre_engine* e, const int64_t* labels, int32_t n_labels);

/* Measured features for the `features` regulariser (regularizers.py:31-43): [N, F] fp32, device or host. */
int bre_engine_load_feature_targets(bre_engine* e, const float* measured, int64_t numel);

/* Start a trial: candidate [N,C,H,W] fp32 (device or host), LR table of length n_lr (host; entry `it` is
 * the step size used by optimiser step `it`, common.py:19-38).  Resets Adam state, best-so-far, history. */
int bre_engine_begin_trial(bre_engine* e, const float* candidate, const float* lr_table, int32_t n_lr);

/* Enqueue `n_iters` iterations of optimization_based_attack.py:110-138 (closure + step + projection +
 * best-so-far + history) on the engine's stream; returns without waiting. */
int bre_engine_run(bre_engine* e, int32_t n_iters);
int bre_engine_sync(bre_engine* e);
/* Same as bre_engine_run + bre_engine_sync, bracketed by CUDA events on the engine's stream: *ms_out = device time. */
int bre_engine_run_timed(bre_engine* e, int32_t n_iters, float* ms_out);

/* After sync: iterations recorded in the history (== len(stats["Trial_k_Val"])), whether a non-finite
 * objective stopped the trial (:131-133), minimal objective so far and last task loss. */
int bre_engine_status(bre_engine* e, int32_t* iters_recorded, int32_t* stopped, double* min_objective,
                      double* last_task_loss);
int bre_engine_read_history(bre_engine* e, float* out_host, int32_t n);
/* Copy best / current candidate ([N,C,H,W] fp32) to `out` (device or host). */
int bre_engine_get_best(bre_engine* e, float* out);
int bre_engine_get_candidate(bre_engine* e, float* out);

/* _score_trial (optimization_based_attack.py:191-204): objective `scoring` (bre_objective: euclidean or
 * cosine, scale 1) of `candidate`; non-finite is reported as +inf. */
int bre_engine_score(bre_engine* e, const float* candidate, int32_t scoring, double* out_score);

/* ---- evaluation without an optimiser step (used by tests and by label/feature tooling) ----------- */
/* Runs the four sweeps + regularisers for `candidate` and returns the total objective and
 * d objective / d candidate (unprocessed, i.e. before noise / clip / sign) into grad_out (device or host). */
int bre_engine_objective_and_gradient(bre_engine* e, const float* candidate, double* objective, float* grad_out);
/* Objective terms of the last evaluation: match, task loss, tv, norm, deep inversion, features. */
int bre_engine_last_terms(bre_engine* e, double* terms6);
/* Debug access (tests): parameter-gradient list G / direction v in torch layout for parameter `index`;
 * which: 0 = G, 1 = v, 2 = W, 3 = g.  out: host, numel floats. */
int bre_engine_debug_param(bre_engine* e, int32_t which, int32_t index, float* out_host);
/* which: 0 = activation, 1 = delta (sweep B), 2 = tangent, 3 = tangent delta; NCHW fp32 to host. */
int bre_engine_debug_tensor(bre_engine* e, int32_t which, int32_t tensor, float* out_host);
/* Number of kernel launches per iteration (for gpu_launches in bench.py) and whether graphs are used. */
int bre_engine_launches_per_iteration(bre_engine* e, int32_t* out);
int bre_engine_set_option(bre_engine* e, const char* name, int64_t value);

/* Joint data + label optimisation on the device (OptimizationJointAttacker._run_trial, optimization_with_label_attack.py:89-143;
 * closure :145-189): like bre_engine_begin_trial, plus the label-logit leaf [N, classes] (rows = batch * seq_len for token
 * models).  Every bre_engine_run iteration then evaluates softmax(labels) as the soft targets of the task loss, the objective,
 * both gradients (candidate and label logits, post-processed separately), steps both leaves with the configured optimiser,
 * projects the candidate only, and keeps the best-so-far pair.  Pointers: device or host. */
int bre_engine_begin_joint_trial(bre_engine* e, const float* candidate, const float* label_logits, int64_t n_label_elems,
                                 const float* lr_table, int32_t n_lr);
/* Label logits of the joint trial: best != 0 -> the best-so-far copy, else the current iterate. */
int bre_engine_get_joint_labels(bre_engine* e, int32_t best, float* out);

/* Candidate augmentations of the closure (optimization_based_attack.py:149-153, auxiliaries/augmentations.py): the model and the
 * priors see view(candidate); `differentiable` != 0 pulls the gradient back through the transposed view, 0 replaces the candidate
 * by its view every iteration (the reference assigns candidate.data).  Pipeline: up to 4 permutation steps in config order
 * (kinds[s]: 1 = discrete_shift with params[s] = lim, 2 = flip with params[s] = p), then the optional continuous_shift
 * (bilinear grid sample, shift in pixels, "circular" wrap as in the reference), then the composite colour affine per (image,
 * channel): out = in * cj_scale + cj_shift (colorjitter; NULL = none; device or host [N * C]).  Random draws: Philox(seed,
 * iteration).  n_steps = 0, cs_enabled = 0 and cj_scale = NULL switch augmentations off. */
int bre_engine_set_augmentations(bre_engine* e, int32_t n_steps, const int32_t* kinds, const float* params, int32_t cs_enabled,
                                 float cs_shift, int32_t cs_circular, const float* cj_scale, const float* cj_shift,
                                 int32_t differentiable, uint64_t seed);
/* The draws of the last evaluation (for parity tests): roll offsets / flip flags per step, continuous-shift uniforms per image. */
int bre_engine_last_augmentation(bre_engine* e, int32_t* o1, int32_t* o2, float* sx, float* sy);
/* Stand-alone view (transpose = 0) or pull-back (transpose = 1: x is the gradient w.r.t. the view and is clobbered when the
 * continuous shift is on; scratch: same size) with explicit draws o1 / o2 per step and sx / sy per image (NULL = no continuous shift). */
int bre_augment_view(const float* x, float* out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t n_steps, const int32_t* kinds,
                     const int32_t* o1, const int32_t* o2, float cs_shift, int32_t cs_circular, const float* sx, const float* sy,
                     const float* cj_scale, const float* cj_shift, int32_t transpose, float* scratch, void* stream);

/* ---- the steps either side of the hot path ----------------------------------------- */
/* User-side update production (cases/users.py:148-169 `_compute_batch_gradient`): one forward + backward of the loaded model
 * on `data` (candidate layout, device or host) with index `labels` -> gradient of the mean task loss w.r.t. every parameter,
 * written to grads_out[i] (model.parameters() order, torch layout, device or host); loss_out (host, may be NULL) = task loss. */
int bre_engine_param_gradients(bre_engine* e, const float* data, const int64_t* labels, int32_t n_labels,
                               float* const* grads_out, int32_t n_params, double* loss_out);
/* Train-mode BN (user without public buffers, users.py:140-143): batch mean and *biased* variance that layer `bn_index` saw in
 * the last forward, from which the user's shipped buffers follow (momentum None: running_mean = mean, running_var = unbiased). */
int bre_engine_bn_batch_stats(bre_engine* e, int32_t bn_index, float* mean_out, float* var_out);
/* Model forward only (analysis/analysis.py:66-69 feature comparison): logits [N, classes] (device or host). */
int bre_engine_forward(bre_engine* e, const float* data, float* logits_out);
/* Per-example mean squared error between the de-normalised ([x * std + mean], per channel; NULL = identity), optionally
 * [0,1]-clamped reconstruction and ground truth (analysis/analysis.py:228-242); PSNR per example = 10 log10(1 / mse)
 * (analysis/metrics.py:108-130) follows on the host.  rec, ref: device fp32 [N, C, HW]; mean / std: host [C]; mse: host [N]. */
int bre_image_mse(const float* rec, const float* ref, int32_t N, int32_t C, int32_t HW, const float* mean, const float* std,
                  int32_t clamp01, double* mse_host, void* stream);

/* Bilinear resize of an NCHW fp32 batch on the device, F.interpolate(mode="bilinear", align_corners=False) semantics: the
 * stage-to-stage up-sampling of MultiScaleOptimizationAttacker (multiscale_optimization_attack.py:45-69). */
int bre_resize_bilinear(const float* src, float* dst, int32_t N, int32_t C, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, void* stream);

/* ---- stand-alone kernels (each is also a stage of the engine; exposed for parity tests + rooflines) */
/* Multi-tensor gradient-matching reduction (objectives.py:91-95,135-141,160-164,185-196).
 * G, g: device fp32 [n]; chunk_weights: device fp32 [ceil(n/1024)] or NULL.  sums5 (host, double):
 * <G,g>, |G|^2, |g|^2, sum (G-g)^2, sum w |G-g|; NULL = launch only, no read-back (for timing the bare kernel).  */
int bre_match_reduce(const float* G, const float* g, const float* chunk_weights, int64_t n, float mask_value,
                     double* sums5_host, void* stream);
/* TotalVariation value + gradient (regularizers.py:130-147) for x [N,3,H,W] device fp32;
 * grad (device) is overwritten when accumulate == 0. */
int bre_total_variation(const float* x, float* grad, int32_t N, int32_t H, int32_t W, float scale, float inner_exp,
                        float outer_exp, float eps, int32_t double_opponents, int32_t accumulate,
                        double* value_host, void* stream);
/* Token-sequence ops of the transformer / TAG path (language_models.py:150-205 as attacked in embedding space; SURVEY
 * section 8 rows a15 / a16), stand-alone, one call per sweep (0 forward, 1 backward, 2 tangent-forward, 3 tangent-backward;
 * rules in oracle/transformer_interp.py).  fp32 device pointers, [rows, C] row-major, rows = batch * seq_len.
 *   layernorm: x = the op's input; in1..in3 per sweep: (1) dy | (2) x' | (3) dy', dy, x'; stats [rows, 2] is written by
 *              sweep 0 and read by the others; sweep 1 also writes g_gamma / g_beta when non-NULL.
 *   attention: qkv [rows, 3 d] = (q | k | v) projections, heads x dh = d, no mask; in1..in3 per sweep: (1) dO [rows, d] |
 *              (2) (qkv)' | (3) dO', dO, (qkv)'; P / Pd [B, heads, T, T] are written by sweeps 0 / 2 and read later.
 * The engine's sweeps do not dispatch to these kernels yet (the transformer model family is the next row to be built). */
int bre_token_layernorm(int32_t sweep, const float* x, const float* in1, const float* in2, const float* in3, const float* gamma,
                        const float* beta, const float* v_gamma, const float* v_beta, float eps, int32_t rows, int32_t C, float* stats,
                        float* out, float* g_gamma, float* g_beta, void* stream);
int bre_token_attention(int32_t sweep, const float* qkv, const float* in1, const float* in2, const float* in3, int32_t B, int32_t T,
                        int32_t heads, int32_t dh, float* P, float* Pd, float* out, void* stream);

/* Token recovery of the text attacks (replaces `_postprocess_text_data._max_similarity`, breaching/attacks/base_attack.py:126-133):
 * tokens[n] = argmax_v <r_n - mean, e_v - mean> / |r_n - mean|^2 / |e_v - mean|^2 over the V rows of emb [*, d] (rows picked through
 * `subset` [V] when non-NULL; ids are then positions in that list).  rec [rows, d], tokens [rows] int64; device pointers. */
int bre_token_match(const float* rec, const float* emb, const int64_t* subset, int32_t rows, int32_t d, int32_t V, int64_t* tokens,
                    void* stream);

/* Implicit-GEMM convolution family, NHWC activations / OHWI weights, fp32:
 * mode 0 fprop  : out[N,Ho,Wo,Co]  = conv(in[N,H,W,Ci], w[Co,R,S,Ci]) (+ conv(in2, w2) when in2 != NULL)
 * mode 1 dgrad  : din[N,H,W,Ci]    = conv^T(dout[N,Ho,Wo,Co], w) (+ conv^T(dout2, w2))
 * mode 2 wgrad  : dw[Co,R,S,Ci]    = sum_pixels dout (x) in
 * backend 0 = SIMT fp32, 1 = tcgen05 TF32 (where available, else BRE_ERR_UNSUPPORTED), 2 = the engine's dispatch
 * (tcgen05 where the shape is covered, SIMT otherwise). */
int bre_conv_gemm(int32_t mode, int32_t backend, const float* a, const float* w, const float* a2, const float* w2,
                  float* out, int32_t N, int32_t H, int32_t W, int32_t Ci, int32_t Co, int32_t R, int32_t S,
                  int32_t stride, int32_t pad, void* stream);

const char* bre_last_error(void);
const char* bre_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BREACHING_B200_H */
