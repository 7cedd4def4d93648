/*
 * avdf.h — C-ABI of libavdf_sm100.so: the B200 (sm_100a) kernels behind the temporal-localization
 * inference path of audio-visual/Audio_Visual_Deepfake_Detection.
 *
 * Conventions
 *   - every entry point returns AVDF_OK (0) or a negative AVDF_ERR_* code and never throws;
 *     avdf_last_error() returns the message of the calling thread's last failure;
 *   - all pointers are DEVICE pointers owned by the caller unless a comment says "host";
 *     nothing is allocated behind the caller's back (scratch comes in through workspace
 *     pointers whose size is queried with the matching *_workspace_bytes function);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream;
 *   - activations are token-major ("channel-last"): [batch, rows_per_video, channels], channels
 *     contiguous; weights of a conv with k taps are [c_out, k * c_in] with the tap index outermost
 *     inside a row (W[co][j * c_in + ci] = reference weight[co][ci][j]).
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   avdf_nms_hard / avdf_nms_soft   pybind11 module nms_1d_cpu {nms, softnms},
 *                                   libs/utils/csrc/nms_cpu.cpp:19-58, 67-160, 172-182
 *   avdf_postprocess                libs/modeling/av_fd_no_recon.py:760-876 (inference_single_video,
 *                                   postprocessing) + libs/utils/nms.py:8-190 (NMSop, SoftNMSop,
 *                                   seg_voting, batched_nms)
 *   avdf_interp_concat              libs/datasets/deepfake_video_audio.py:513-547 (F.interpolate x3 + cat)
 *   avdf_host_pack                  HOST: DataLoader collate + pin of the raw stream arrays,
 *                                   libs/datasets/deepfake_video_audio.py:547-558 (dataset item) and the
 *                                   per-video `.to(device)` of libs/modeling/av_fd_no_recon.py:476-477
 *   avdf_pack_feats                 preprocessing, libs/modeling/av_fd_no_recon.py:431-479 (pad + batch layout)
 *   avdf_conv_gemm                  MaskedConv1D (libs/modeling/blocks.py:13-63) as used by the
 *                                   embedding (backbones.py:437-445), the 1x1 projections and MLP of the
 *                                   transformer blocks (blocks.py:1169-1171,1223,1291-1297), FPN laterals
 *                                   (necks.py:69-73), head towers (av_fd_no_recon.py:75-89,144-159) and
 *                                   DownBlock convs (blocks.py:1495-1516), with LayerNorm (blocks.py:70-112),
 *                                   activation, positional encoding and residual/AffineDropPath fused
 *   avdf_ln_dwconv_ln               LN -> depthwise MaskedConv1D(k3, stride) -> LN of LocalMaskedMHCA /
 *                                   LocalMaskedMMHCA (blocks.py:1159-1165, 726-741) incl. the nearest
 *                                   up/down-sampling of backbones.py:487,490 and MaxPool1d skip (blocks.py:1277-1281)
 *   avdf_attention                  banded / global softmax attention (blocks.py:977-1224, 274-313)
 *   avdf_ln_rows                    LayerNorm before the MLP (blocks.py:1311)
 *   avdf_instnorm_lrelu             InstanceNorm1d + LeakyReLU of DownBlock (blocks.py:1508-1515)
 *   avdf_fpn_fuse                   FPN1D top-down sum + depthwise conv + LN (necks.py:75-93)
 *   avdf_head_final                 last conv of the cls / reg heads + Scale + ReLU (av_fd_no_recon.py:82-89,152-159)
 *   avdf_vcls_exp12 / _exp13        video-level classifier tails (blocks.py:1608-1626, 1682-1700)
 */
#ifndef AVDF_H_
#define AVDF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AVDF_API __attribute__((visibility("default")))
#else
#define AVDF_API
#endif

/* 3: avdf_conv_gemm_args grew (dot_w / dot_n / dot_out, tap_rows); avdf_head_combine, avdf_logmel, avdf_byola_* added */
#define AVDF_ABI_VERSION 3
#define AVDF_MAX_LEVELS 8
#define AVDF_MAX_SEGS 1024

enum { AVDF_OK = 0, AVDF_ERR_INVALID = -1, AVDF_ERR_CUDA = -2, AVDF_ERR_UNSUPPORTED = -3 };
enum { AVDF_DTYPE_F32 = 0, AVDF_DTYPE_BF16 = 1, AVDF_DTYPE_F16 = 2 };
enum { AVDF_ACT_NONE = 0, AVDF_ACT_RELU = 1, AVDF_ACT_GELU = 2 };

/* ---- runtime ---- */
AVDF_API int avdf_abi_version(void);
AVDF_API const char* avdf_last_error(void);
/* fills SM count and compute capability of the current device */
AVDF_API int avdf_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---- K1: per-stream linear resize to t_out + channel concat (video | byola | emo) ----
 * Streams are packed row-major [sum_b T_s(b), C_s] fp32 with per-stream prefix offsets [batch+1]
 * (rows). A stream with C_s == 0 is absent. out: [batch, t_out, c_video+c_byola+c_emo]. */
AVDF_API int avdf_interp_concat(const float* video, const float* byola, const float* emo,
                       const int32_t* video_off, const int32_t* byola_off, const int32_t* emo_off,
                       int32_t batch, int32_t c_video, int32_t c_byola, int32_t c_emo, int32_t t_out,
                       void* out, int32_t out_dtype, void* stream);

/* The same with the raw streams stored as bf16 (in_dtype = AVDF_DTYPE_BF16): the opt-in 16-bit feature-shard format of the
 * ingestion path (half the host -> device bytes per video; the interpolation arithmetic stays fp32). */
AVDF_API int avdf_interp_concat_in(const void* video, const void* byola, const void* emo, int32_t in_dtype,
                          const int32_t* video_off, const int32_t* byola_off, const int32_t* emo_off,
                          int32_t batch, int32_t c_video, int32_t c_byola, int32_t c_emo, int32_t t_out,
                          void* out, int32_t out_dtype, void* stream);

/* ---- preprocessing (av_fd_no_recon.py:431-479): one video's feats [channels, t] fp32 (dataset item layout)
 * -> token-major rows out[t_padded, channels], zero-padded from t to t_padded. */
AVDF_API int avdf_pack_feats(const float* feats_ct, int32_t channels, int32_t t, int32_t t_padded, void* out,
                    int32_t out_dtype, void* stream);

/* ---- NMS: same contracts as nms_1d_cpu.nms / nms_1d_cpu.softnms, on device memory ----
 * segs [n,2], scores [n]; out_idx [n] int64 (kept input indices, descending score / pick order);
 * out_count [1]. max_num <= 0: run to completion (the reference's behaviour); max_num > 0: stop after
 * that many picks (the wrappers consume no more, nms.py:29-30,56-63). dets [n,3] is written in place
 * for the picks, like the reference. */
AVDF_API size_t avdf_nms_workspace_bytes(int32_t n);
AVDF_API int avdf_nms_hard(const float* segs, const float* scores, int32_t n, float iou_threshold, int32_t max_num,
                  int64_t* out_idx, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);
AVDF_API int avdf_nms_soft(const float* segs, const float* scores, int32_t n, float* dets, float iou_threshold,
                  float sigma, float min_score, int32_t method, int32_t max_num, int64_t* out_idx,
                  int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* ---- decode + batched_nms + seconds conversion, one CTA per video ---- */
typedef struct avdf_postprocess_args {
  int32_t batch;
  /* decode inputs (logits == NULL: candidates are provided in cand_* instead) */
  const float* logits;           /* [batch, P]      P = sum(level_len), level-major per video */
  const float* offsets;          /* [batch, P, 2]   */
  const uint8_t* mask;           /* [batch, P]      */
  int32_t n_levels;
  int32_t level_len[AVDF_MAX_LEVELS];
  float level_stride[AVDF_MAX_LEVELS];
  float pre_nms_thresh; int32_t pre_nms_topk; float duration_thresh;
  /* candidates: written by decode (and read by voting), or given */
  float* cand_segs;              /* [batch, cand_cap, 2] */
  float* cand_scores;            /* [batch, cand_cap]    */
  int32_t* cand_count;           /* [batch]              */
  int32_t cand_cap;
  /* batched_nms (class-agnostic path, nms.py:160-190) */
  float iou_threshold, min_score, sigma, voting_thresh;
  int32_t max_seg_num, use_soft_nms, soft_method;   /* use_soft_nms: 0 hard, 1 soft, -1 no NMS (test_cfg nms_method
                                                     * 'none', av_fd_no_recon.py:847: the decoded candidates go straight
                                                     * to the seconds conversion; out_* then hold cand_cap entries per video) */
  /* seconds conversion (all NULL: stay on the feature grid) */
  const float* vid_feat_stride;  /* [batch] */
  const float* vid_half_nframes; /* [batch] 0.5 * feat_num_frames */
  const float* vid_fps;          /* [batch] */
  const float* vid_duration;     /* [batch] */
  /* outputs, sorted by descending score */
  float* out_segs;               /* [batch, max_seg_num, 2] */
  float* out_scores;             /* [batch, max_seg_num]    */
  int32_t* out_count;            /* [batch]                 */
  void* workspace; size_t workspace_bytes;
  /* optional result records for the multi-GPU gather (the reference's per-video JSON record,
   * libs/utils/train_utils.py:577-595): every video appends ONE fixed-size fp32 row
   *   [video index, count, video_cls, scores[max_seg_num], segs[max_seg_num][2]]
   * to rec_ring at row atomicAdd(rec_counter, 1) % rec_cap. All NULL / 0: no records. */
  float* rec_ring;               /* [rec_cap, 3 + 3 * max_seg_num] */
  uint32_t* rec_counter;         /* [1] device counter, owned by the caller */
  int32_t rec_cap;
  const int32_t* vid_index;      /* [batch] global video index of every row of the batch */
  const float* vid_cls;          /* [batch] video-level logit */
  /* optional: index (into the video's candidate list) of every returned segment, in output order - the label of a
   * pick in class-agnostic NMS over several classes is cls_idxs[index] (libs/utils/nms.py:159-180) */
  int32_t* out_index;            /* [batch, max_seg_num] */
} avdf_postprocess_args;          /* host struct */
AVDF_API size_t avdf_postprocess_workspace_bytes(int32_t batch, int32_t cand_cap);
AVDF_API int avdf_postprocess(const avdf_postprocess_args* args, void* stream);

/* ---- conv-as-GEMM with fused epilogue ----
 * out[b, seg_o_row + t, n] = epi( sum_{j<taps} sum_{c<c_in} W[n, j*c_in + c] * A[b, seg_a_row + stride*t + j - taps/2, c] )
 * (rows outside [0, stride * seg_t_out) of the level read as zero), for every segment `seg`. Segments are levels of a
 * pyramid (shared weights) or independent problems stacked along the rows with their own weight block (seg_w_row):
 * e.g. the q, k, v projections of one attention block in a single launch.
 * epi(v): v += bias[n]; v *= mask[row]; v = LN_n(v) * ln_w + ln_b; v = act(v); v += pe[t, n] * mask[row];
 *         v = residual[row, n] * mask[row] + gamma[n] * v            (each step only if its pointer is set;
 *         ln_after_residual moves the LayerNorm behind the residual step, see the field below)
 * dtype F32 -> fp32 CUDA-core path (parity mode); BF16 / F16 -> TMA + tcgen05 tensor-core path (A and W in that
 * 16-bit format, fp32 accumulate in TMEM). out_h (optional) receives a 16-bit copy in out_h_dtype (BF16 | F16). */
#define AVDF_MAX_TAPS 9
typedef struct avdf_conv_gemm_args {
  int32_t batch, n_out, c_in, taps, stride, n_seg;
  int32_t seg_t_out[AVDF_MAX_LEVELS];
  int32_t seg_a_row[AVDF_MAX_LEVELS];
  int32_t seg_o_row[AVDF_MAX_LEVELS];
  int32_t seg_w_row[AVDF_MAX_LEVELS];  /* row offset into w / bias / ln_* / gamma for this segment (0: shared weights);
                                        * w then holds n_w_rows >= seg_w_row + n_out rows (0 means n_out) */
  int32_t n_w_rows;
  int64_t a_rows_per_video, o_rows_per_video;
  const void* a;                 /* [batch, a_rows_per_video, c_in] */
  const void* w;                 /* [n_out, taps * c_in] */
  int32_t dtype;
  const float* bias; const uint8_t* row_mask; const float* ln_w; const float* ln_b; int32_t act;
  const float* pe; const float* residual; const float* gamma;
  float* out_f32; void* out_h; int32_t out_h_dtype;   /* [batch, o_rows_per_video, n_out] */
  void* workspace; size_t workspace_bytes;
  /* ln_after_residual != 0 (16-bit path, n_out = 256, residual + ln_w + out_f32 + out_h all set): the LayerNorm is applied
   * AFTER the residual step instead of before it - out_f32 receives y = residual * mask + gamma * ((acc + bias) * mask)
   * (the block's residual stream, blocks.py:1309-1310 / 868-869) and out_h receives LN(y) * ln_w + ln_b, the operand of
   * the block's MLP (blocks.py:1311): the attention projection and LN2 in one launch. act must be NONE, pe NULL. */
  int32_t ln_after_residual;
  /* tap_mode 0: taps centred on the output position (offsets j - taps/2, MaskedConv1D); 1: forward taps (offsets
   * 0 .. taps-1, stride 1 only): the row pair (x[t], x[t+1]) a stride-2 ConvTranspose1d(k=3, padding=1, output_padding=1)
   * reads for its outputs 2t and 2t+1 (blocks.py:1443-1491) - see libs/modeling/engine.py `up_block` */
  int32_t tap_mode;
  /* Row dot products instead of (or next to) the output tensor (16-bit path, n_out = 256, wide configuration):
   * dot_out[row, j] = sum_n dot_w[j, n] * v[row, n] for j < dot_n <= 6, v = the epilogue's result. With dot_out set,
   * out_f32 and out_h may both be NULL: the last layer of the cls / reg towers then emits only the per-tap partial sums of
   * the heads' final k3 convolution (av_fd_no_recon.py:82-89, 152-159), which avdf_head_combine turns into logits / offsets. */
  const float* dot_w;            /* [dot_n, n_out] fp32 or NULL */
  int32_t dot_n;
  float* dot_out;                /* [batch, o_rows_per_video, dot_n] fp32 or NULL */
  /* Explicit tap table (stride 1, tap_mode 0, taps <= AVDF_MAX_TAPS): tap j reads input row t + tap_rows[j] instead of
   * t + j - taps/2. With the rows of a (time, mel) grid flattened and padded (csrc/byola.cu "grid layout") the nine offsets
   * {-1,0,1} * row_pitch + {-1,0,1} make the launch a 3x3 Conv2d (audio_feature/content_audio/byol_a/models.py:59, 64). */
  const int32_t* tap_rows;       /* host [taps] or NULL */
} avdf_conv_gemm_args;            /* host struct */
AVDF_API size_t avdf_conv_gemm_workspace_bytes(const avdf_conv_gemm_args* args);
AVDF_API int avdf_conv_gemm(const avdf_conv_gemm_args* args, void* stream);

/* ---- LN -> depthwise conv k3 (stride 1|2) * mask -> LN, for up to 3 streams sharing one source ----
 * src [batch, t_src, C] fp32. Virtual input position p in [0, t_virt) reads source row
 * (shift >= 0 ? p >> shift : p << -shift)  (nearest up / down-sampling of backbones.py:487,490).
 * out_s[b, t', :] = LN_out_s( (sum_j dw_s[:, j] * u_s[stride*t' + j - 1]) * mask_out[b, t'] ),
 * u_s[p] = LN(src[map(p)]) * ln_in_w_s + ln_in_b_s inside [0, t_virt), 0 outside.
 * skip_out (optional, stride 2): MaxPool1d(3, 2, 1) of the raw source rows. */
typedef struct avdf_ln_dwconv_ln_args {
  int32_t batch, channels, t_src, t_virt, shift, stride, n_streams;
  const float* src;
  const uint8_t* mask_out;       /* [batch, t_virt / stride] */
  const float* ln_in_w[3]; const float* ln_in_b[3];
  const float* dw_w[3];          /* [C, 3] */
  const float* ln_out_w[3]; const float* ln_out_b[3];
  void* out[3];                  /* [batch, out_rows_per_video, C], row t' of video b at b * out_rows_per_video + t' */
  int32_t out_dtype;
  int32_t out_rows_per_video;    /* 0: t_virt / stride (dense). Larger: several streams interleaved per video */
  float* skip_out;               /* [batch, t_virt / stride, C] fp32 or NULL */
  int32_t tile_rows;             /* output rows per warp tile: 0 = chosen from the problem size, or 2 / 4 / 8 */
} avdf_ln_dwconv_ln_args;
AVDF_API int avdf_ln_dwconv_ln(const avdf_ln_dwconv_ln_args* args, void* stream);

/* ---- fused transformer MLP (blocks.py:1236-1243 `self.mlp`, applied at blocks.py:1315-1316) ----
 * out = residual * mask + gamma * ((GELU(x W1^T + b1) W2^T + b2) * mask), exact-erf GELU, fp32 accumulate; the
 * [rows, hidden] activations stay on chip (16-bit, like the unfused path's intermediate tensor). channels = 256,
 * hidden = 1024, dtype = AVDF_DTYPE_F16 | AVDF_DTYPE_BF16 (x, w1, w2). Equivalent to two avdf_conv_gemm calls. */
typedef struct avdf_mlp_fused_args {
  int32_t rows, channels, hidden, dtype;
  const void* x;                 /* [rows, channels] */
  const void* w1; const float* b1;   /* [hidden, channels], [hidden] */
  const void* w2; const float* b2;   /* [channels, hidden], [channels] */
  const uint8_t* row_mask;       /* [rows] or NULL */
  const float* residual;         /* [rows, channels] fp32 */
  const float* gamma;            /* [channels] or NULL (AffineDropPath scale, blocks.py:1316) */
  float* out;                    /* [rows, channels] fp32 */
  void* out_h;                   /* optional 16-bit copy of out (same dtype as x) or NULL */
  int32_t out_h_t, out_h_pitch, out_h_row0;   /* 0: the copy is dense [rows, channels]; else rows = batch * out_h_t and row
                                               * b * out_h_t + t goes to row b * out_h_pitch + out_h_row0 + t (one level of
                                               * a [batch, P, channels] pyramid buffer, necks.py:62-93's input) */
  /* "block tail" (att != NULL): the attention output projection and LN2 of the block run in the same launch
   * (blocks.py:1223, 1309-1316 / 779, 868-872):
   *   y = skip * mask + gamma_attn * ((att w_o^T + b_o) * mask);  x = LN(y) * ln2_w + ln2_b (16-bit, never leaves the SM);
   *   out = (y + GELU(x w1^T + b1) w2^T + b2) * mask
   * y stays in the tensor-memory accumulator and the second GEMM accumulates on top of it, so the MLP's AffineDropPath
   * scale must be FOLDED by the caller: pass w2 = diag(scale) w2, b2 = scale * b2 and gamma = NULL. x and residual are
   * ignored; y (optional, may be NULL) receives a copy of the residual stream between the two halves of the block.
   * Equivalent to avdf_conv_gemm(ln_after_residual) followed by the plain avdf_mlp_fused. */
  const void* att;               /* [rows, channels] 16-bit (dtype) or NULL */
  const void* w_o; const float* b_o;   /* [channels, channels], [channels] */
  const float* gamma_attn;       /* [channels] or NULL */
  const float* ln2_w; const float* ln2_b;   /* [channels] */
  const float* skip;             /* [rows, channels] fp32: the block's input (or its max-pooled copy) */
  float* y;                      /* [rows, channels] fp32 or NULL */
} avdf_mlp_fused_args;
AVDF_API int avdf_mlp_fused(const avdf_mlp_fused_args* args, void* stream);

/* ---- multi-head attention over q,k,v [batch, t, C]; window > 1: band |i-j| <= window/2 with additive
 * -1e4 on masked keys and zeroed masked query rows (blocks.py:1152-1224); window <= 1: global with
 * -inf on masked keys (blocks.py:274-313). Row i of video b of q/k/v is at (b * qkv_rows_per_video + i) * C from the
 * respective pointer (qkv_rows_per_video 0 = t; 3t when q,k,v of a video are stacked in one buffer). out [batch, t, C]. */
AVDF_API int avdf_attention(const void* q, const void* k, const void* v, const uint8_t* kv_mask, void* out,
                   int32_t in_dtype, int32_t out_dtype, int32_t batch, int32_t t, int32_t qkv_rows_per_video,
                   int32_t channels, int32_t n_head, int32_t window, void* stream);

/* ---- LayerNorm over channels of fp32 rows -> fp32 or bf16 ---- */
AVDF_API int avdf_ln_rows(const float* x, const float* w, const float* b, void* out, int32_t out_dtype, int64_t rows,
                 int32_t channels, void* stream);

/* ---- InstanceNorm1d over T (per video, channel; eps 1e-5, biased) + LeakyReLU(slope) ---- */
AVDF_API int avdf_instnorm_lrelu(const float* x, void* out, int32_t out_dtype, int32_t batch, int32_t t, int32_t channels,
                        float slope, void* stream);

/* ---- FPN top-down fuse: L_l[t] = sum_{j>=l} lat_j[t >> (j-l)]; F_l = LN_l(dwconv3_l(L_l) * mask) ----
 * lat / out are pyramids [batch, P, C] with levels concatenated per video. dw_w [n_levels, C, 3],
 * ln_w / ln_b [n_levels, C]. */
AVDF_API int avdf_fpn_fuse(const float* lat, const uint8_t* mask, const float* dw_w, const float* ln_w, const float* ln_b,
                  void* out, int32_t out_dtype, int32_t batch, int32_t channels, int32_t n_levels,
                  const int32_t* level_len /* host */, void* stream);

/* ---- last conv (k3) of both heads: logits [batch,P] = cls_w . x_cls + cls_b (* mask);
 * offsets [batch,P,2] = relu((reg_w . x_reg + reg_b) * mask * scale_l). Towers are pyramids [batch,P,C]. */
AVDF_API int avdf_head_final(const void* cls_feat, const void* reg_feat, int32_t dtype, const uint8_t* mask,
                    const float* cls_w /* [1,3*C] */, const float* cls_b, const float* reg_w /* [2,3*C] */,
                    const float* reg_b, const float* level_scale /* host [n_levels] */, float* logits,
                    float* offsets, int32_t batch, int32_t channels, int32_t n_levels,
                    const int32_t* level_len /* host */, void* stream);

/* The same from per-tap partial sums (avdf_conv_gemm dot_out of the last tower layers): cls_dots [batch, P, 3] holds, for
 * every pyramid row, the dot products of its tower output with the three taps of cls_head.cls_head.conv.weight; reg_dots
 * [batch, P, 6] = (output o, tap j) at o * 3 + j. logit[t] = dots[t-1][0] + dots[t][1] + dots[t+1][2] + bias inside the level. */
AVDF_API int avdf_head_combine(const float* cls_dots, const float* reg_dots, const uint8_t* mask, const float* cls_b,
                      const float* reg_b, const float* level_scale /* host [n_levels] */, float* logits, float* offsets,
                      int32_t batch, int32_t n_levels, const int32_t* level_len /* host */, void* stream);

/* ---- video-level classifier tails ---- */
/* exp12 (blocks.py:1608-1626): z [batch, t <= 32, C] (after the last DownBlock) -> logit [batch]. The two dense
 * weights are passed TRANSPOSED ([in, out]: conv0_wt[k][c] = conv0.weight[c][k], lin1_wt[j][c] = conv1.weight[c][j]). */
AVDF_API int avdf_vcls_exp12(const void* z, int32_t dtype, const float* conv0_wt /* [C,C] */, const float* lin1_wt /* [2C,C] */,
                    const float* ln_w, const float* ln_b, const float* lin2_w /* [C] */, const float* lin2_b,
                    float* out, int32_t batch, int32_t t, int32_t channels, void* stream);
/* exp13 (blocks.py:1682-1700): z [batch, t, C] -> logit [batch] */
AVDF_API int avdf_vcls_exp13(const void* z, int32_t dtype, const float* conv0_w /* [C,C] */, const float* seg_w /* [C] */,
                    const float* seg_b, const float* cls_w /* [2] */, const float* cls_b, float* out,
                    int32_t batch, int32_t t, int32_t channels, void* stream);

/* ---- upstream BYOL-A feature extractor (audio_feature/content_audio) ----
 * A batch of clips is packed along time. frames(clip) = 1 + samples / 160 (centre = True); all index arrays are DEVICE
 * int arrays prepared by the host side (libs/features/byola.py `BatchPlan`). */
/* replaces torchaudio MelSpectrogram(16 kHz, n_fft = win = 1024, hop 160, 64 mels, 60-7800 Hz) + log(x + eps) +
 * PrecomputedNorm (extract_audio_feature_one.py:34-42, 66; byol_a/augmentations.py:218-219):
 * wav = the clips back to back, clip c = samples [clip_sample_off[c], clip_sample_off[c+1]) (each > 512 samples),
 * its frames are rows [clip_frame_off[c], clip_frame_off[c+1]) of lms [total_frames, 64] (time-major). One CTA transforms
 * frames 2j and 2j+1 of a clip together: clip_pair_off[c] = sum over earlier clips of ceil(frames / 2), total_pairs its end.
 * window [1024] (periodic Hann), twiddle [1024][2] = (cos, -sin)(2 pi j / 1024); the mel triangles in sparse form: filter m
 * weighs the power bins mel_lo[m] .. mel_lo[m] + mel_cnt[m] - 1 with mel_w[m * mel_stride + 0 ..] (all DEVICE arrays;
 * mel_lo[m] + mel_cnt[m] <= 513). */
AVDF_API int avdf_logmel(const float* wav, const int64_t* clip_sample_off, const int32_t* clip_frame_off,
                const int32_t* clip_pair_off, int32_t n_clips, int32_t total_frames, int32_t total_pairs, const float* window,
                const float* twiddle, const int32_t* mel_lo, const int32_t* mel_cnt, const float* mel_w, int32_t mel_stride,
                float mean, float std, float* lms, void* stream);
/* features.0-3 of AudioNTT2020Task6 (byol_a/models.py:54-57): Conv2d(1, 64, 3, padding 1) + BatchNorm2d (eval; folded by
 * the caller into w [64, 9] (mel tap major) and b [64]) + ReLU + MaxPool2d(2) -> level-1 grid layout
 * out [n_steps * 34, 64] (csrc/byola.cu): step s holds pooled time step t_of_step[s] of clip clip_of_step[s] (-1: an
 * all-zero separator step). mask_out (optional) [n_steps * 34] receives 1 for rows holding data, 0 for padding rows. */
AVDF_API int avdf_byola_conv1_pool(const float* lms, const int32_t* clip_frame_off, const float* w, const float* b,
                const int32_t* clip_of_step, const int32_t* t_of_step, int32_t n_steps, void* out, int32_t out_dtype,
                uint8_t* mask_out, void* stream);
/* MaxPool2d(2) (models.py:62, 67) from the grid layout with mel_in rows per step (clip c starts at step clip_step_in[c])
 * into the next grid layout (pad_out 1: [n_steps_out * (mel_in / 2 + 2), 64]) or into dense rows for the fc layers
 * (pad_out 0: [n_steps_out, (mel_in / 2) * 64], feature index = mel * 64 + channel as models.py:80-82 flattens it). */
AVDF_API int avdf_byola_pool(const void* in, int32_t dtype, int32_t mel_in, const int32_t* clip_step_in,
                const int32_t* clip_of_step, const int32_t* t_of_step, int32_t n_steps_out, int32_t pad_out, void* out,
                uint8_t* mask_out, void* stream);

/* ---- HOST: gather n byte spans (src[i] -> dst[i], nbytes[i] bytes; all HOST pointers, any alignment; dst is normally
 * a slice of a pinned staging buffer) with up to n_threads threads of a pool that lives inside the library
 * (non-temporal stores). Synchronous; concurrent calls are serialised. No CUDA call is made. */
AVDF_API int avdf_host_pack(const void* const* src, void* const* dst, const size_t* nbytes, int32_t n, int32_t n_threads);

/* HOST: 1 when every span [src[i], src[i] + nbytes[i]) is page-locked host memory (cudaHostAlloc / cudaHostRegister), else 0. */
AVDF_API int avdf_host_all_pinned(const void* const* src, const size_t* nbytes, int32_t n);
/* n asynchronous copies src[i] (pinned HOST) -> dst[i] (DEVICE) on `stream`: ingestion without a staging copy. The sources
 * must stay alive and unchanged until the stream has passed the copies. */
AVDF_API int avdf_h2d_gather(const void* const* src, void* const* dst, const size_t* nbytes, int32_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVDF_H_ */
