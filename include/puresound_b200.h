/*
 * puresound_b200 — C ABI of the B200 (sm_100a) separator-forward engine.
 *
 * Drop-in boundary for the hot path of mcw519/PureSound.
 * The reference has no native code: every entry point below replaces one or
 * more implicit ATen library calls made from the reference's nn.Module.forward()
 * bodies; the reference file:line each one replaces is cited per function
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub
 * a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless stated;
 *   - activations are "frames-major": [batch, rows(frames), channels], channel
 *     index fastest (the reference's [N, C, T] transposed; ps_transpose converts);
 *   - no allocation and no host synchronisation: callers pass outputs/scratch
 *     and a cudaStream_t (as void*).  Re-entrant across host threads, streams
 *     and devices: the only process-wide state is a set of std::atomic caches
 *     of idempotent facts (SM count per device, "dynamic shared-memory limit
 *     raised for kernel X on device D", A/B switches read from the environment),
 *     each stored only after the call it stands for succeeded;
 *   - return value: 0 on success, negative ps_status otherwise
 *     (ps_error_string() gives text; the Python host maps them to the
 *     reference's exception types).
 */
#ifndef PURESOUND_B200_H
#define PURESOUND_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_ABI_VERSION 1
#if defined(__GNUC__)
#define PS_API __attribute__((visibility("default")))
#else
#define PS_API
#endif

typedef enum {
  PS_OK = 0,
  PS_ERR_INVALID_ARG = -1, /* NULL pointer, negative size, inconsistent strides   */
  PS_ERR_UNSUPPORTED = -2, /* legal in the reference but outside this engine      */
  PS_ERR_CUDA = -3,        /* launch failed; see ps_last_cuda_error()             */
  PS_ERR_NO_DEVICE = -4    /* not an sm_100 device                                */
} ps_status;

/* activation codes (prologue and epilogue) */
enum { PS_ACT_NONE = 0, PS_ACT_PRELU = 1, PS_ACT_RELU = 2, PS_ACT_TANH = 3, PS_ACT_SIGMOID = 4 };
/* GEMM / depthwise prologue modes: what is applied to the input on load */
enum {
  PS_PRO_NONE = 0,
  PS_PRO_AFFINE = 1,  /* act(x*scale[b,k] + shift[b,k]): gLN/gGN/bN1d folded to scale/shift */
  PS_PRO_ROWNORM = 2, /* act((x-mean[b,r])*rstd[b,r]*gamma[k] + beta[k]): cLN             */
  PS_PRO_MASK = 3     /* x * act(x2[b,r,k]): real mask apply (base_nn.py:146-159)          */
};
/* GEMM back ends */
enum { PS_GEMM_AUTO = 0, PS_GEMM_SIMT = 1, PS_GEMM_TCGEN05 = 2 };

PS_API const char* ps_error_string(int status);
PS_API const char* ps_last_cuda_error(void); /* thread-local text of the last CUDA failure */
PS_API int ps_version(void);
/* 1 if the current device is compute capability 10.x, else 0 (negative on error) */
PS_API int ps_device_ok(void);
/* sizeof() of the descriptor structs, so FFI hosts can verify their mirror: 0 ps_gemm_t, 1 ps_dwconv_t, 2 ps_lstm_t, 3 ps_stream_dw_t, 4 ps_gated_t, 5 ps_stream_hop_block_t, 6 ps_stream_hop_t */
PS_API int64_t ps_struct_size(int which);

/* --------------------------------------------------------------------------
 * Dense contraction  Y[b,r,m] = epi( sum_k pro(X[b,r,k]) * W[m,k] )
 *
 * Replaces every groups=1 kernel-size-1 nn.Conv1d / nn.Linear on the path
 *   conv_tasnet.py:44-46,65  lobe/cnn.py:75-76  dprnn.py:89-91,101-103,107-109
 *   lobe/trivial.py:137-142  lobe/pooling.py:71-86  egs/tse/model.py:133
 * the LSTM input projections (W_ih x + b, dprnn.py:67-103 via nn.LSTM), and — with
 * x_row_stride = hop < K — the framed filterbank / STFT analysis convolutions
 *   lobe/encoder.py:50-56,370-373   and the synthesis GEMMs of :62-68,:429-430.
 * The prologue carries the preceding norm + PReLU (lobe/norm.py:20-50, nn.PReLU)
 * or the mask product; the epilogue carries bias, the per-item embedding bias
 * (conv_tasnet.py:80-83 folded: W[:,C:]*e), an activation, the residual add
 * (conv_tasnet.py:88) and the Welford partials of the output for the next gLN.
 * -------------------------------------------------------------------------- */
typedef struct {
  int64_t batch, rows, M, K;
  const float* X; int64_t x_batch_stride, x_row_stride; /* floats; row stride may be < K */
  const float* W; int64_t w_row_stride;                 /* [M,K], K contiguous           */
  float* Y; int64_t y_batch_stride, y_row_stride;
  /* prologue */
  int32_t pro_mode, pro_act;
  const float* pro_a;   /* AFFINE: scale [batch or 1, K]; ROWNORM: gamma [K]             */
  const float* pro_b;   /* AFFINE: shift                ; ROWNORM: beta  [K]             */
  int64_t pro_batch_stride; /* AFFINE: K, or 0 when scale/shift are batch independent    */
  const float* pro_rowstats; /* ROWNORM: [batch, rows, 2] = (mean, rstd)                 */
  const float* pro_slope;    /* device scalar: PReLU slope (nn.PReLU() has one)          */
  const float* X2;           /* MASK: second operand, same strides as X                  */
  /* epilogue */
  const float* bias;         /* [M] or NULL                                               */
  const float* bias_batch;   /* [batch, M] or NULL                                        */
  int32_t epi_act; int32_t backend; /* PS_ACT_*, PS_GEMM_*                                 */
  const float* epi_slope;    /* device scalar for PS_ACT_PRELU in the epilogue            */
  const float* residual; int64_t res_batch_stride, res_row_stride;
  float* stats_partials;     /* NULL or [batch, ps_gemm_stats_slots, 3] (count,mean,M2)   */
  /* tcgen05 path only: weights pre-packed by ps_gemm_pack_weights (else NULL)            */
  const void* W_packed;
  /* optional fused gLN/gGN finalize (needs stats_partials): the folded affine ps_stats_finalize would produce from the
   * partials is written to fin_scale / fin_shift [batch, M] by the CTA that completes an item's last tile (no extra
   * launch; same fixed merge order).  fin_counter: [batch] uint32, zero before the first use, left zero afterwards. */
  const float* fin_gamma; const float* fin_beta; float fin_eps;
  float* fin_scale; float* fin_shift; uint32_t* fin_counter;
  /* optional row LayerNorm in the epilogue (ln_eps > 0): Y = residual + LN_M(X W^T + bias) * ln_gamma + ln_beta, the
   * Linear -> nn.LayerNorm -> + of the DPRNN blocks (dprnn.py:161-163,173-175).  Fused into the tcgen05 kernel when
   * M == 128, otherwise the library runs ps_rownorm after the GEMM (Y must not alias residual in that case). */
  const float* ln_gamma; const float* ln_beta; float ln_eps;
} ps_gemm_t;

PS_API int ps_gemm(const ps_gemm_t* d, void* stream);
/* which kernel ps_gemm would run for this descriptor on the current device (no launch): 0 exact-fp32 CUDA cores,
 * 1 single-CTA tcgen05, 2 CTA-pair tcgen05 (128-frame tiles), 3 wide CTA-pair tcgen05 (256-frame tiles), 4 few-channel
 * tcgen05 (M <= 128: frames on the MMA's M side, weights resident in shared memory); negative ps_status on error.  For tests, profiles and bench labels. */
PS_API int ps_gemm_path(const ps_gemm_t* d);
/* number of (count,mean,M2) slots per batch item ps_gemm writes for this shape */
PS_API int64_t ps_gemm_stats_slots(int64_t rows, int64_t M);
/* bytes of the packed-weight image for the tcgen05 path (0 if shape not eligible) */
PS_API int64_t ps_gemm_packed_bytes(int64_t M, int64_t K);
PS_API int ps_gemm_pack_weights(const float* W, int64_t w_row_stride, int64_t M, int64_t K, void* packed, void* stream);

/* --------------------------------------------------------------------------
 * gLN / gGN statistics (lobe/norm.py:20-34, :96): merge the per-CTA Welford
 * partials of one tensor, in a fixed order, into per-item (mean, biased var) and
 * emit the folded affine  scale[b,c] = gamma[c]*rstd_b,  shift[b,c] = beta[c] -
 * mean_b*rstd_b*gamma[c]  consumed by PS_PRO_AFFINE.  meanvar [batch,2] optional.
 * -------------------------------------------------------------------------- */
PS_API int ps_stats_finalize(const float* partials, int64_t batch, int64_t slots, const float* gamma, const float* beta,
                      float eps, int64_t C, float* scale, float* shift, float* meanvar, void* stream);
/* Welford partials (count, mean, M2) of a strided region, `slots` per item: element (n, m, r, c) at
 * x + n*batch_stride + m*mid_stride + r*row_stride + c.  partials [batch, slots, 3] feed ps_stats_finalize.  gLN of the
 * U-Net layers (lobe/norm.py:20-34 on [N,C,F,T], unet.py:121,147), whose conv outputs are strided views. */
PS_API int ps_stats_region(const float* x, int64_t batch, int64_t mid, int64_t rows, int64_t C, int64_t batch_stride,
                           int64_t mid_stride, int64_t row_stride, int64_t slots, float* partials, void* stream);

/* bN1d in eval mode (lobe/norm.py:94): scale = w/sqrt(rv+eps), shift = b - rm*scale; [C] each */
PS_API int ps_bn_fold(const float* weight, const float* bias, const float* running_mean, const float* running_var,
               float eps, int64_t C, float* scale, float* shift, void* stream);
/* cLN / nn.LayerNorm statistics: out[row] = (mean, 1/sqrt(biased var + eps)) */
PS_API int ps_rowstats(const float* x, int64_t rows, int64_t C, int64_t row_stride, float eps, float* out, void* stream);

/* --------------------------------------------------------------------------
 * Dilated depthwise Conv1d with the preceding norm+PReLU in its prologue
 * (lobe/cnn.py:62-74; padding rule :58-60; the causal right-trim of :100-101 is
 * realised by causal taps).  x,y: [batch, T, C]; w: [C, P]; taps read
 * x[t + (p - (P-1)/2)*d] (non-causal, P odd) or x[t - (P-1-p)*d] (causal); taps
 * outside [0,T) contribute 0 after the prologue.
 * -------------------------------------------------------------------------- */
typedef struct {
  int64_t batch, T, C; int32_t P, dilation, causal;
  const float* x; float* y;
  const float* w; const float* bias; /* [C,P], [C] */
  int32_t pro_mode, pro_act;         /* PS_PRO_NONE/AFFINE/ROWNORM */
  const float* pro_a; const float* pro_b; int64_t pro_batch_stride;
  const float* pro_rowstats; const float* pro_slope;
  float* stats_partials;             /* NULL or [batch, ps_dwconv_stats_slots, 3] */
  int64_t stats_slots;               /* set by the library: row pitch of stats_partials (callers leave 0) */
  /* optional fused gLN/gGN finalize, as in ps_gemm_t (fin_scale / fin_shift are [batch, C]) */
  const float* fin_gamma; const float* fin_beta; float fin_eps;
  float* fin_scale; float* fin_shift; uint32_t* fin_counter;
} ps_dwconv_t;
PS_API int ps_dwconv(const ps_dwconv_t* d, void* stream);
PS_API int64_t ps_dwconv_stats_slots(int64_t T, int64_t C);

/* --------------------------------------------------------------------------
 * Row-wise normalisation  y[r,:] = (res ? res[r,:] : 0) + act((x-mean_r)*rstd_r*w + b)
 * nn.LayerNorm + residual of the DPRNN blocks (dprnn.py:161-163,173-175), the
 * FiLM input norm (lobe/trivial.py:158-159), and cLN+PReLU for streaming.
 * -------------------------------------------------------------------------- */
PS_API int ps_rownorm(const float* x, const float* res, float* y, int64_t rows, int64_t C, const float* w, const float* b,
               float eps, int32_t act, const float* slope, void* stream);

/* Learned-filterbank / iSTFT overlap-add (lobe/encoder.py:62-68,85-94; lobe/stft.py:103-115;
 * encoder.py:449-454) fused with the output constraint (base_nn.py:414-424):
 * y[b,j] = sum_t frames[b,t,j-t*hop];  if wsum: y /= wsum[j] where wsum[j] > 1e-10;
 * constraint 0 none, 1 clamp(-1,1), 2 sigmoid.  frames [batch,T,win]; y [batch,(T-1)*hop+win]. */
PS_API int ps_ola(const float* frames, int64_t batch, int64_t T, int64_t win, int64_t hop, const float* wsum,
           int32_t constraint, float* y, void* stream);

/* Mask activation + apply (base_nn.py:81-95,41-79,97-112,146-159).  complex=0: y=f*act(m);
 * complex=1 on [.., 2F] channel halves: (a+ib)(c+id) with (c,d)=act(m).  n = batch*rows. */
PS_API int ps_mask_apply(const float* feats, const float* mask, float* y, int64_t n_rows, int64_t C, int32_t act,
                  int32_t is_complex, void* stream);

/* Magnitude (lobe/trivial.py:35-58): y[r,f] = sqrt(re^2 + im^2 + 1e-8) on channel halves,
 * optional DC drop and log1p.  x [n_rows, 2F]; y [n_rows, F - drop_first]. */
PS_API int ps_magnitude(const float* x, float* y, int64_t n_rows, int64_t F, int32_t drop_first, int32_t log1p, void* stream);

/* SpecAugment of the mel speaker front-end (lobe/trivial.py:307-335 -> torchaudio mask_along_axis): in place,
 * x[b,t,c] = value where bounds[0] <= c < bounds[1] or bounds[2] <= t < bounds[3]; one band per axis for the whole
 * batch.  bounds = 4 int32 in DEVICE memory (the host draws them per call; a captured graph replays with new bands). */
PS_API int ps_band_fill(float* x, int64_t batch, int64_t T, int64_t C, const int32_t* bounds, float value, void* stream);

/* AttentiveStatisticsPooling tail (lobe/pooling.py:108-126): softmax over frames of
 * logits[b,t,c], weighted mean and sqrt(clamp(sum w (x-mean)^2, 1e-12)).  out [batch, 2C]. */
PS_API int ps_asp_pool(const float* x, const float* logits, int64_t batch, int64_t T, int64_t C, float* out, void* stream);

/* F.normalize(p=2, dim=1, eps=1e-12) (conv_tasnet.py:348-349, dprnn.py:128-129). x,y [rows, E] */
PS_API int ps_l2normalize(const float* x, float* y, int64_t rows, int64_t E, void* stream);

/* DPRNN segmentation (lobe/trivial.py:178-241, dprnn.py:133-145,180-189).
 * overlap=1: seg[b,q,k,:] = x[b, q*K/2 + k - K/2, :] (0 outside [0,T)); merge averages the
 * two covers.  overlap=0: zero-padded reshape / crop.  x [batch,T,C]; seg [batch,S,K,C]. */
PS_API int ps_segment(const float* x, float* seg, int64_t batch, int64_t T, int64_t C, int64_t K, int64_t S, int32_t overlap, void* stream);
PS_API int ps_merge(const float* seg, float* y, int64_t batch, int64_t T, int64_t C, int64_t K, int64_t S, int32_t overlap, void* stream);

/* --------------------------------------------------------------------------
 * LSTM recurrence (nn.LSTM num_layers=1, gate order i,f,g,o; dprnn.py:67-103,160,170).
 * gx [positions, D*4H] holds W_ih x + b_ih + b_hh for every position (ps_gemm);
 * w_hh_t [D, H, 4H] is W_hh transposed per direction.  Sequence q (< n_seq), step t
 * (< L) lives at position  (q / inner)*outer_stride + (q % inner)*inner_stride +
 * t*step_stride, which expresses both the intra-chunk ([N*S] x K) and the
 * inter-chunk ([N*K] x S) passes over one [N,S,K,*] tensor without a permute.
 * h0/c0/hn/cn: [D, n_seq, H] or NULL.  out [positions, D*H]; direction 1 runs t reversed.
 * -------------------------------------------------------------------------- */
typedef struct {
  int64_t n_seq, L, H; int32_t D;
  int64_t inner, outer_stride, inner_stride, step_stride;
  const float* gx; const float* w_hh_t;
  const float* h0; const float* c0;
  float* out; float* hn; float* cn;
  /* image built by ps_lstm_pack_weights from w_hh_t, or NULL (exact-fp32 CUDA-core kernel reading w_hh_t).  H <= 128 with
   * H % 32 == 0: the tensor-core kernel's resident image; every other H <= 256: the CUDA-core kernel's gate-minor fp32 image
   * [D][k][unit][4 gates] (one 16-byte weight vector per k and thread, streamed through a cp.async ring; same results). */
  const void* w_packed;
  /* 0: gx rows are [D][4 gates][H] (nn.LSTM order); 1: [D][H][4 gates] (rows of W_ih permuted by the caller so the four
   * gates of a unit are one 16-byte load); needs w_packed */
  int32_t gx_interleaved;
} ps_lstm_t;
PS_API int ps_lstm(const ps_lstm_t* d, void* stream);
/* bytes of the packed recurrent-weight image for this H (0 if H > 256 or D is not 1 / 2) */
PS_API int64_t ps_lstm_packed_bytes(int64_t H, int32_t D);
/* w_hh_t [D, H, 4H] -> per direction: W_hh as bf16 hi (shared-memory tile image) | bf16 lo (row-major, goes to TMEM) for
 * the tensor-core sizes; the gate-minor fp32 image for the others */
PS_API int ps_lstm_pack_weights(const float* w_hh_t, int64_t H, int32_t D, void* packed, void* stream);

/* FiLM combine (lobe/trivial.py:163-165): y = sb[:, :C] * xn + sb[:, C:]; sb [rows, 2C] */
PS_API int ps_film_combine(const float* sb, const float* xn, float* y, int64_t rows, int64_t C, void* stream);

/* Gated product of GatedTCN (conv_tasnet.py:141-205):  y = a' * sigmoid(b'),  a' = act_a(pro_a(a)), b' = act_b(pro_b(b)),
 * with pro_* as in ps_gemm_t (NONE / AFFINE / ROWNORM; rowstats are [batch, rows, 2]) so the two branch norms + PReLUs
 * are applied on load.  b == NULL: y = a' (strided copy-transform, used to write the FiLM-modulated right-branch input
 * into the zero-padded conv buffer, conv_tasnet.py:197-200).  All strides in floats (float4 path when C, the strides
 * and the pointers are 16-byte friendly, scalar otherwise). */
typedef struct {
  int64_t batch, rows, C;
  const float* a; int64_t a_batch_stride, a_row_stride;
  const float* b; int64_t b_batch_stride, b_row_stride;
  float* y; int64_t y_batch_stride, y_row_stride;
  /* optional middle level (0 or 1: none): element (n, m, r, c) at batch*n + mid*m + row*r + c; the per-item AFFINE
   * parameters are indexed by n, ROWNORM statistics by ((n*mid + m)*rows + r).  Used by the U-Net shell to copy
   * [N, T, F, C] tensors into the time-shifted, frequency-padded tap buffers of its 2-D convs (unet.py:100-175). */
  int64_t mid, a_mid_stride, b_mid_stride, y_mid_stride;
  int32_t a_mode, a_act; const float* a_pa; const float* a_pb; int64_t a_pro_batch_stride; const float* a_rowstats; const float* a_slope;
  int32_t b_mode, b_act; const float* b_pa; const float* b_pb; int64_t b_pro_batch_stride; const float* b_rowstats; const float* b_slope;
} ps_gated_t;
PS_API int ps_gated(const ps_gated_t* d, void* stream);

/* Multi-head self-attention core (nn.MultiheadAttention inside lobe/attention.py:37-113, used by DPARNblock2D,
 * dparn.py:12-108): qkv [batch, L, 3E] = fused in-projection (q | k | v), out [batch, L, E];
 * out[b,t,head] = sum_j softmax_j(q_t . k_j / sqrt(E/heads)) v_j, j <= t only when causal.  Head dim in {4,8,16,32,64}. */
PS_API int ps_attention(const float* qkv, float* out, int64_t batch, int64_t L, int64_t E, int32_t heads, int32_t causal,
                        void* stream);

/* SDR / SI-SNR scoring (loss/sdr.py:104-185, 263-299): out[r] = 10 log10(|t|^2 / (|e|^2 + tau |t|^2 + eps) + eps) in dB with
 * t = alpha s2 (alpha = <s1,s2>/(<s2,s2>+eps) when scaled, else 1), e = s1 - t (s1 - s2 when scale_dependent), both signals
 * mean-removed first when zero_mean; tau = 10^(-sdr_max/10) or 0.  s1, s2: rows of L samples `stride` floats apart. */
PS_API int ps_sdr(const float* s1, const float* s2, int64_t rows, int64_t L, int64_t stride1, int64_t stride2, int32_t scaled,
                  int32_t scale_dependent, int32_t zero_mean, float tau, float eps, float* out, void* stream);

/* [batch, R, C] -> [batch, C, R] (boundary conversion to/from the reference's [N,C,T]) */
PS_API int ps_transpose(const float* x, float* y, int64_t batch, int64_t R, int64_t C, void* stream);

/* --------------------------------------------------------------------------
 * Streaming (new; API pattern of puresound/streaming/skim_inference.py:142-218,
 * oracle = offline causal forward).  One call advances every stream by one frame.
 * step is a device int64 counter (so a captured CUDA graph can be replayed).
 * -------------------------------------------------------------------------- */
/* causal dilated depthwise step with device ring buffers:
 *   v = act(norm1(u[s,:])) -> ring[s, step % RL, :];  y[s,:] = act(norm2(bias + sum_p w[:,p] * ring[s,(step-(P-1-p)*d) % RL,:]))
 * norm kind 0: cLN (gamma/beta, eps) computed over the row; 1: per-channel affine (bN1d folded). RL=(P-1)*d+1. */
typedef struct {
  int64_t streams, C; int32_t P, dilation;
  const float* u; float* y; float* ring; const int64_t* step;
  const float* w; const float* bias;
  int32_t norm_kind; float eps;
  const float* n1_a; const float* n1_b; const float* slope1;
  const float* n2_a; const float* n2_b; const float* slope2;
} ps_stream_dw_t;
PS_API int ps_stream_dwconv_step(const ps_stream_dw_t* d, void* stream);
/* frame assembly: frame[s,:] = [hist[s,:] | chunk[s,:]], hist <- last (win-hop) samples */
PS_API int ps_stream_push(const float* chunk, float* hist, float* frame, int64_t streams, int64_t win, int64_t hop, void* stream);
/* overlap-add emit: acc[s,:] += frame[s,:]; out[s,:hop] = constrain(acc[s,:hop]); acc <- shift left by hop */
PS_API int ps_stream_ola(const float* frame, float* acc, float* out, int64_t streams, int64_t win, int64_t hop, int32_t constraint, void* stream);
PS_API int ps_stream_advance(int64_t* step, void* stream);

/* --------------------------------------------------------------------------
 * One hop of every concurrent stream of a causal Conv-TasNet as ONE persistent cooperative kernel (round 2): frame
 * assembly, encoder, every TCN block (1x1 conv, per-frame norm + PReLU, dilation-history ring + causal depthwise taps,
 * 1x1 conv, 1x1 conv + residual), mask + decoder, overlap-add emit - phases of one grid (one CTA per SM) separated by
 * grid barriers, exact fp32.  Same state and arithmetic as the ps_stream_* chain above + ps_gemm (API pattern:
 * StreamingSkiM.step_frame, streaming/skim_inference.py:176-218; arithmetic: conv_tasnet.py:11-90, lobe/cnn.py:58-79,
 * lobe/norm.py:37-50, lobe/encoder.py:50-94).  norm_kind 0: cLN (n*_a / n*_b = gamma / beta), 1: eval-BatchNorm folded to
 * a per-channel affine (n*_a / n*_b = scale / shift).  All pointers are device pointers; `blocks` is a device array.
 * -------------------------------------------------------------------------- */
typedef struct {
  const float* w_in; int64_t w_in_ld;   /* [H, >= C]: the first C columns of in_conv (speaker columns are folded) */
  const float* ebias;                   /* [streams, H] per-stream speaker bias W_in[:, C:] e, or NULL           */
  const float* n1_a; const float* n1_b; const float* slope1;
  const float* dw_w; const float* dw_b; /* depthwise taps [H, P], bias [H] or NULL                                */
  const float* n2_a; const float* n2_b; const float* slope2;
  const float* w_pw; const float* b_pw; /* [H, H], [H] or NULL                                                    */
  const float* n3_a; const float* n3_b; const float* slope3;
  const float* w_out; const float* b_out; /* [C, H], [C] or NULL                                                  */
  float* ring;                          /* [streams, (P-1)*dilation + 1, H] dilation history                     */
  int32_t P, dilation;
  /* optional bf16 split of the three 1x1-conv weights for the tensor-core path (many streams): [M][K] hi then [M][K] lo
   * bf16 (hi = bf16(w), lo = bf16(w - hi)); NULL = exact-fp32 FFMA path */
  const uint16_t* w_in_p; const uint16_t* w_pw_p; const uint16_t* w_out_p;
} ps_stream_hop_block_t;

typedef struct {
  int64_t streams;
  int32_t C, H, win, hop, n_blocks, norm_kind, enc_relu, mask_act, constraint; float eps;
  const float* w_enc;                   /* [C, win]                                                               */
  const float* w_dec_t;                 /* [win, C] (the ConvTranspose1d weight transposed)                       */
  const ps_stream_hop_block_t* blocks;  /* device array [n_blocks]                                                */
  const float* chunk;                   /* [streams, hop] new samples                                             */
  float* hist; float* frame; float* frame_out; float* acc; float* out; /* [S,win-hop] [S,win] [S,win] [S,win] [S,hop] */
  int64_t* step;                        /* frames processed so far (advanced by the kernel)                       */
  float* feats; float* x; float* u1; float* u2; float* u3; /* scratch [S,C] [S,C] [S,H] [S,H] [S,H]                */
  uint32_t* barrier;                    /* 2 words, zero before the first launch, owned by the kernel afterwards  */
  const uint16_t* w_enc_p; const uint16_t* w_dec_p; /* optional bf16 splits of w_enc / w_dec_t (see the block struct)   */
} ps_stream_hop_t;
PS_API int ps_stream_hop(const ps_stream_hop_t* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PURESOUND_B200_H */
