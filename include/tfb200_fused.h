/*
 * tfb200_fused.h -- C ABI of the fused transformer-layer kernels in libmsda_b200.so (sm_100a).
 *
 * These have NO counterpart in the reference's native code: the reference runs them as chains of PyTorch ops
 *   x = norm(x + dropout(branch))        src/trackformer/models/deformable_transformer.py:284-285, 291-292,
 *                                        360-361, 370-371, 377-378
 * They sit on the hot path and are exposed through the same Python module as extra
 * functions; the reference-facing surface stays include/msda_b200.h.
 *
 * Conventions as in msda_b200.h: caller owns all (device) buffers, work is enqueued on `stream`, 0 = success,
 * negative = argument error, positive = cudaError_t.
 */
#ifndef TFB200_FUSED_H_
#define TFB200_FUSED_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFB200_E_NULLPTR (-1)
#define TFB200_E_SHAPE   (-2)   /* C must be a multiple of 4, <= 512 (LayerNorm backward) / 1024 (forward, colsum) */
#define TFB200_LN_MAX_CTAS 592  /* persistent grid: 148 SMs x 4 CTAs; also the row count of `partial_ws` */

/* number of CTAs (= rows of the [ctas][2][C] fp32 workspace the backward needs) for a problem of `rows` rows */
int tfb200_ln_partial_ctas(int64_t rows);

/* s = x + branch * keep_mask / keep_prob   (keep_mask == NULL: s = x + branch)
 * y = LayerNorm(s) * gamma + beta          mean/rstd [rows] and s_out [rows][C] are optional outputs (NULL to skip) */
int tfb200_add_dropout_layernorm_fwd_f32(const float* x, const float* branch, const uint8_t* keep_mask,
                                         const float* gamma, const float* beta, float* s_out, float* y,
                                         float* mean, float* rstd, int64_t rows, int C, float keep_prob, float eps,
                                         void* stream);

/* dx = dL/dx (= dL/ds), dbranch = dL/dbranch (NULL to skip), dgamma/dbeta [C]; partial_ws: [ctas][2][C] fp32 */
int tfb200_add_dropout_layernorm_bwd_f32(const float* dy, const float* s, const uint8_t* keep_mask,
                                         const float* gamma, const float* mean, const float* rstd, float* dx,
                                         float* dbranch, float* dgamma, float* dbeta, float* partial_ws,
                                         int64_t rows, int C, float keep_prob, void* stream);

/* Same pair with mask-free dropout: keep decisions are a counter-based hash of (*seed_dev, element index) against
 * keep_prob (the scheme of tfb200_relu_dropout_*), so no mask tensor is drawn, stored or read; forward and backward
 * must be given the same seed.                                                                                        */
int tfb200_add_dropout_layernorm_seeded_fwd_f32(const float* x, const float* branch, const int64_t* seed_dev,
                                                const float* gamma, const float* beta, float* s_out, float* y,
                                                float* mean, float* rstd, int64_t rows, int C, float keep_prob,
                                                float eps, void* stream);
int tfb200_add_dropout_layernorm_seeded_bwd_f32(const float* dy, const float* s, const int64_t* seed_dev,
                                                const float* gamma, const float* mean, const float* rstd, float* dx,
                                                float* dbranch, float* dgamma, float* dbeta, float* partial_ws,
                                                int64_t rows, int C, float keep_prob, void* stream);

/* out[c] = sum_r x[r][c]  (bias gradient of a Linear over a long token axis); C % 4 == 0, C <= 1024;
 * partial_ws: [tfb200_ln_partial_ctas(rows)][C] fp32.  Deterministic (fixed summation order).                      */
int tfb200_colsum_f32(const float* x, float* out, float* partial_ws, int64_t rows, int C, void* stream);

/* h = dropout(relu(a)) in one pass (FFN hidden activation, deformable_transformer.py:283,359).  Inverted dropout with a
 * counter-based keep decision hashed from (*seed_dev, element index): no mask tensor.  backward: grad_a from grad_h and
 * h alone (h > 0 <=> a > 0 and kept).  n % 4 == 0.  training == 0: plain ReLU.                                        */
int tfb200_relu_dropout_fwd_f32(const float* a, float* h, const int64_t* seed_dev, int64_t n, float keep_prob,
                                int training, void* stream);
int tfb200_relu_dropout_bwd_f32(const float* grad_h, const float* h, float* grad_a, int64_t n, float keep_prob,
                                int training, void* stream);

/* proj [rows][3*M*L*P] = per query [offsets (M,L,P,2) | attention logits (M,L*P)]  ->  sampling locations
 * loc [rows][M][L][P][2] and softmax attention weights attn [rows][M][L][P] in one pass (ms_deform_attn.py:69-82).
 * ref: [rows][L][ref_dim] reference points (ref_dim 2) or boxes (4); shapes_f32: [L][2] level sizes AS STORED (H, W)
 * in fp32 (only read for ref_dim == 2).  L*P in {4, 8, 16, 32}, (M*L*P) % 32 == 0.  backward: grad_proj from grad_loc,
 * grad_attn and the saved attn (no gradient for ref: callers use it only when ref does not require one).            */
int tfb200_sampling_prep_fwd_f32(const float* proj, const float* ref, const float* shapes_f32, float* loc, float* attn,
                                 int64_t rows, int M, int L, int P, int ref_dim, void* stream);
int tfb200_sampling_prep_bwd_f32(const float* grad_loc, const float* grad_attn, const float* attn, const float* ref,
                                 const float* shapes_f32, float* grad_proj, int64_t rows, int M, int L, int P, int ref_dim,
                                 void* stream);

/* Device-side Hungarian matching (replaces the host scipy.optimize.linear_sum_assignment calls of
 * src/trackformer/models/matcher.py:104,127).  cost: [K][B][Q][T] fp32 (K decoder layers, B images, Q queries, T = all
 * ground-truth boxes of the batch, image b owning columns offsets[b] .. offsets[b+1]-1; offsets_dev: int32 [B+1] on the
 * device).  Every (k, b) problem assigns each of image b's boxes to a distinct query (boxes per image <= max_targets <= Q).
 * src/tgt: int64 [K][T]; for each problem the pairs are written to columns offsets[b].. in ascending query order:
 * src = query index, tgt = global box index.  *status_dev (int32, caller-zeroed) becomes 1 (more boxes than queries)
 * or 2 (infeasible: a box whose costs are all +inf).  Double-precision duals, like scipy.                           */
int tfb200_lsa_f32(const float* cost, const int* offsets_dev, int64_t* src, int64_t* tgt, int K, int B, int Q, int T,
                   int max_targets, int* status_dev, void* stream);

/* Detection post-processing for the online tracker (deformable_detr.py:286-334 + the per-track read-backs of
 * tracker.py:306-330): logits [N][Q][C], boxes [N][Q][4] (cx, cy, w, h normalised), sizes_hw [N][2] int64 (h, w) on the
 * DEVICE  ->  packed [N][Q][6] = {sigmoid score of the best class, its index, x0, y0, x1, y1 in pixels (unclipped)} and,
 * optionally, labels [N][Q] int64.  Ties between classes resolve to the smallest index (torch.max).                 */
int tfb200_detect_postprocess_f32(const float* logits, const float* boxes, const int64_t* sizes_hw, float* packed,
                                  int64_t* labels, int N, int Q, int C, void* stream);

/* clip_grad_norm_ + AdamW (decoupled weight decay) over one contiguous fp32 range of n parameters in one pass
 * (engine.py:147-151 + torch.optim.AdamW of train.py:118).  grad_norm_dev: device scalar holding the global gradient
 * 2-norm, or NULL for no clipping; the clip coefficient min(1, max_norm / (norm + 1e-6)) is applied on the fly, the
 * gradient buffer is not modified.  step >= 1 is the update count (bias correction).  Pointers 16-byte aligned.      */
int tfb200_flat_adamw_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                          const float* grad_norm_dev, float max_norm, float lr, float beta1, float beta2, float eps,
                          float weight_decay, int64_t step, void* stream);

/* Frozen batch-norm + optional residual add + optional ReLU over channels-last activations ([pixels][C], C % 4 == 0):
 *   y = act(x * scale[c] + shift[c] (+ residual))          backbone.py:46-55 + torchvision Bottleneck.forward
 * backward: g = relu ? dy * (y > 0) : dy;  dx = g * scale[c] (NULL to skip);  dresidual = g (NULL when there was no
 * residual).  scale / shift are the folded per-channel terms (weight * rsqrt(var + eps), bias - mean * scale).        */
int tfb200_frozen_bn_act_fwd_f32(const float* x, const float* residual, const float* scale, const float* shift,
                                 float* y, int64_t pixels, int C, int relu, void* stream);
int tfb200_frozen_bn_act_bwd_f32(const float* dy, const float* y, const float* scale, float* dx, float* dresidual,
                                 int64_t pixels, int C, int relu, void* stream);

/* y[M][N] = act(x[M][K] . w[N][K]^T + bias[N])  -- the nn.Linear of the long-token projections / FFN
 * (ops/modules/ms_deform_attn.py:64,69,70,88; models/deformable_transformer.py:282-286) as ONE hand-written
 * tcgen05 kernel: TMA-staged 128x32 fp32 tiles, TF32 tensor-core products (tcgen05.mma kind::tf32), fp32
 * accumulation in TMEM, bias (+ optional ReLU) in the epilogue, TMA store.  All row-major, contiguous, 16-byte aligned.
 * Requires N % 128 == 0 and K % 32 == 0 (tfb200_tf32_linear_supported); M is arbitrary.  bias may be NULL.
 * Returns -5 (unsupported) outside that domain: the caller keeps the library GEMM for those.                          */
int tfb200_tf32_linear_supported(int64_t M, int N, int K);
int tfb200_tf32_linear_f32(const float* x, const float* w, const float* bias, float* y, int64_t M, int N, int K,
                           int relu, void* stream);
/* The layer's two backward products on the same pipeline:
 *   dx[M][K]  = dy[M][N] . w[N][K]            (N % 32 == 0, K % 128 == 0)
 *   dw[N][K]  = dy[M][N]^T . x[M][K]          (N, K % 128 == 0; the token axis M is split over all SMs, the partial
 *                                              products are added into dw by TMA reductions; dw is zero-filled here) */
int tfb200_tf32_linear_dgrad_f32(const float* dy, const float* w, float* dx, int64_t M, int N, int K, void* stream);
int tfb200_tf32_linear_wgrad_f32(const float* dy, const float* x, float* dw, int64_t M, int N, int K, void* stream);

/* out[r][c] = sigmoid(delta[r][c] + inverse_sigmoid(ref[r][c]))  for c < ref_dim (2 or 4), sigmoid(delta[r][c]) beyond;
 * inverse_sigmoid(x) = log(clamp(clamp(x,0,1), eps) / clamp(1 - clamp(x,0,1), eps))  (util/misc.py:515-519) -- the box
 * refinement of deformable_transformer.py:412-422 / deformable_detr.py:229-248 in one launch.  delta/out [rows][4],
 * ref [rows][ref_dim].  backward: grad_delta [rows][4], grad_ref [rows][ref_dim] (NULL to skip), from grad_out and out. */
int tfb200_refine_boxes_fwd_f32(const float* delta, const float* ref, float* out, int64_t rows, int ref_dim, float eps,
                                void* stream);
int tfb200_refine_boxes_bwd_f32(const float* grad_out, const float* out, const float* ref, float* grad_delta,
                                float* grad_ref, int64_t rows, int ref_dim, float eps, void* stream);

/* SetCriterion / HungarianMatcher on the device (csrc/set_loss.cu; reference: models/matcher.py:60-100,
 * models/detr.py:213-328, util/misc.py:540-571, util/box_ops.py:24-61).  R = K*B*Q rows (layer, image, query).
 *   match_cost : cost[R][T] = w_bbox * L1 + w_class * focal class cost - w_giou * GIoU against the T boxes of the batch
 *   set_loss_fwd: src/tgt [K][T] are the matched (query, global box) pairs, columns offsets[b]..offsets[b+1]-1 belong to
 *                 image b; writes out5k[5][K] = loss_ce, loss_bbox, loss_giou, cardinality_error, class_error and the
 *                 UNIT gradients (unit_logits [R][C], unit_l1 / unit_giou [R][4]) that set_loss_bwd scales by the incoming
 *                 gradients g_*[K] / num_boxes.  n_gt [B] = boxes per image (float), num_boxes = device scalar.
 *                 row_loss [R], row_flags [R], pair_l1 / pair_giou [K][T] are scratch.                                  */
int tfb200_match_cost_f32(const float* logits, const float* boxes, const int64_t* tgt_ids, const float* tgt_boxes,
                          float* cost, int64_t R, int C, int T, float w_class, float w_bbox, float w_giou, float alpha,
                          float gamma, void* stream);
int tfb200_set_loss_fwd_f32(const float* logits, const float* boxes, const int64_t* src, const int64_t* tgt,
                            const int64_t* tgt_ids, const float* tgt_boxes, const int* offsets, const float* n_gt,
                            const float* num_boxes, float* unit_logits, float* unit_l1, float* unit_giou, float* row_loss,
                            int* row_flags, float* pair_l1, float* pair_giou, float* out5k, int K, int B, int Q, int C,
                            int T, float alpha, float gamma, void* stream);
int tfb200_set_loss_bwd_f32(const float* unit_logits, const float* unit_l1, const float* unit_giou, const float* g_ce,
                            const float* g_l1, const float* g_giou, const float* num_boxes, float* grad_logits,
                            float* grad_boxes, int K, int B, int Q, int C, void* stream);

/* Stem of the ResNet trunk in one pass (csrc/frozen_bn_act.cu): y = maxpool 3x3 / stride 2 / pad 1 of
 * relu(x * scale[c] + shift[c]); x [N][H][W][C] channels-last, y [N][(H+1)/2][(W+1)/2][C]; forward only (frozen stem). */
int tfb200_frozen_bn_relu_maxpool_f32(const float* x, const float* scale, const float* shift, float* y, int N, int H, int W,
                                      int C, void* stream);

/* Dense self-attention over a few hundred positions, 32 channels per head (csrc/small_attn.cu; replaces the core of the
 * decoder's nn.MultiheadAttention, reference models/deformable_transformer.py:342,366-368).  q / k / v / out (and dout /
 * dq / dk / dv) are [L][B][H][32] VIEWS: channel stride 1, head stride 32, sequence / batch strides (in elements) in
 *   strides8  = {q_l, q_b, k_l, k_b, v_l, v_b, out_l, out_b}
 *   strides16 = strides8 + {dout_l, dout_b, dq_l, dq_b, dk_l, dk_b, dv_l, dv_b}
 * out = dropout(softmax(scale * q k^T + mask)) v; key_pad [B][L] (1 = ignore that key) or NULL; seed_dev NULL = no
 * dropout, else keep_prob in (0, 1] with the mask-free hash RNG; lse [B][H][L] (saved for the backward);
 * delta_ws [B][H][L] scratch.                                                                                       */
int tfb200_small_attn_fwd_f32(const float* q, const float* k, const float* v, const uint8_t* key_pad,
                              const int64_t* seed_dev, float* out, float* lse, int B, int H, int L,
                              const int64_t* strides8, float scale, float keep_prob, void* stream);
int tfb200_small_attn_bwd_f32(const float* q, const float* k, const float* v, const uint8_t* key_pad,
                              const int64_t* seed_dev, const float* out, const float* lse, const float* dout,
                              float* dq, float* dk, float* dv, float* delta_ws, int B, int H, int L,
                              const int64_t* strides16, float scale, float keep_prob, void* stream);

/* One frame of the online tracker's bookkeeping on the device (csrc/track_step.cu; reference models/tracker.py:266-548:
 * score thresholds 337-373 / 425-436, both NMS passes 388-406 / 494-515, public-detection gating 122-164, ReID 166-264,
 * results 533-545, reid_sim_only 547-548).  One CTA takes every decision of Tracker.step in the reference's order.
 *
 * State is a struct of arrays in track order: entries [0, n_active) are the active tracks, [n_active, n_active +
 * n_inactive) the inactive ones, both in the reference's list order.  header = {n_active, n_inactive, track_num,
 * num_reids, n_query_next, error, n_results, 0}.  `in` is CONSUMED (it is the kernel's work table, rows beyond the
 * tracks are used for this frame's object queries: capacity >= n_active + n_inactive + nq, <= 2048); `out` receives
 * the new state, `q_boxes` / `q_embeds` the track queries of the next frame (active tracks, then the inactive ones that
 * survive the patience / positive-area test, boxes as cx, cy, w, h divided by the image size), and `result` the header
 * followed by n_results rows of 8 words {id, obj_ind, score, x0, y0, x1, y1, 0} (ints and fp32 bit patterns).
 * error: 1 = n_query does not match the state, 2 = capacity.                                                         */
typedef struct TfbTrackState {
  int32_t* header;             /* [8] */
  int32_t* ids;                /* [capacity] */
  float* pos;                  /* [capacity][4]  x0, y0, x1, y1 in pixels */
  float* anchor;               /* [capacity][4]  position at the start of the step (Track.last_pos[-1]) */
  float* score;                /* [capacity] */
  int32_t* obj_ind;            /* [capacity] */
  int32_t* count_inactive;     /* [capacity] */
  int32_t* count_termination;  /* [capacity] */
  float* bank;                 /* [capacity][hidden] last output embedding of every track */
} TfbTrackState;

typedef struct TfbTrackStepArgs {
  const float* rows;           /* [n_query + nq][6] packed detections of this frame (tfb200_detect_postprocess_f32) */
  const float* hs_embeds;      /* [n_query + nq][hidden] decoder output embeddings */
  const float* public_dets;    /* [n_public][4] pixel xyxy, or NULL */
  TfbTrackState in, out;
  float* q_boxes;              /* [capacity][4] */
  float* q_embeds;             /* [capacity][hidden] */
  int32_t* result;             /* [8 + 8 * capacity] */
  int32_t* iscratch;           /* [16 * capacity] */
  float* fscratch;             /* [8 * capacity + capacity * max(nq, n_public, 1)] */
  double* dscratch;            /* [4 * capacity] */
  double inactive_patience, reid_sim_threshold;
  float detection_obj_score_thresh, track_obj_score_thresh, reid_score_thresh, detection_nms_thresh, track_nms_thresh;
  int32_t capacity, hidden, nq, n_query, n_public, img_h, img_w;
  int32_t overflow_boxes, public_mode /* 0 off, 1 center_distance, 2 min_iou_0_5 */, reid_greedy_matching,
      reid_sim_only, steps_termination, detection_nms_on, track_nms_on;
} TfbTrackStepArgs;

int tfb200_track_step_f32(const TfbTrackStepArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TFB200_FUSED_H_ */
