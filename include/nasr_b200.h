/*
 * nasr_b200.h — C ABI of libnasr_b200.so, the B200 (sm_100a) engine for the
 * TCN / GCN eval-mode forward of francescopapaleo/neural-audio-spring-reverb.
 *
 * The reference has no FFI of its own (pure Python / PyTorch); the drop-in
 * boundary is its Python surface.  Every entry point below names the
 * reference interface it replaces (paths relative to the reference root,
 * src/nasr = src/neural_audio_spring_reverb):
 *
 *   nasr_engine_create     <- TCN.__init__ / GCN.__init__ + load_state_dict
 *                             (src/nasr/networks/tcn.py:94-148, gcn.py:80-138,
 *                              networks/model_utils.py:160-163)
 *   nasr_set_cond          <- FiLM.forward's adaptor Linear + eval BatchNorm
 *                             (src/nasr/networks/custom_layers.py:32-42)
 *   nasr_forward           <- TCN.forward / GCN.forward
 *                             (src/nasr/networks/tcn.py:150-155, gcn.py:140-147)
 *   nasr_forward_host      <- the timed region of make_inference: host tensor
 *                             -> device -> model(input, c) -> host
 *                             (src/nasr/inference.py:36-61,77)
 *   nasr_stream_reset /
 *   nasr_forward_chunk     <- PaddingCached / Conv1dCached streaming state
 *                             (src/nasr/wrapper.py:14-57)
 *   nasr_block_forward     <- TCNBlock.forward / GCNBlock.forward on its own
 *                             (src/nasr/networks/tcn.py:73-86, gcn.py:53-61)
 *   nasr_postprocess       <- the post-processing of make_inference: peak normalise,
 *                             torchaudio highpass_biquad (lfilter, clamp), peak normalise
 *                             (src/nasr/inference.py:70-78)
 *   nasr_eval_metrics      <- the MAE / ESR / DC metrics of evaluate_model (src/nasr/eval.py:38-40,118-121)
 *   nasr_rt60              <- measure_rt60's Schroeder integration (src/nasr/tools/rt60.py:49-70)
 *   nasr_convolve_full     <- the direct convolution of measure_model_ir (src/nasr/tools/ir_model.py:138-140)
 *
 * Conventions
 *   - All tensors are fp32, contiguous, reference layout: x [B, in_ch, T],
 *     y [B, out_ch, T], cond [B, cond_dim].
 *   - Pointers named *_dev are device pointers on the engine's device; *_host
 *     are host pointers (pinned memory makes the copies asynchronous).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     All device work is enqueued on it; nothing synchronises internally
 *     except nasr_forward_host, which returns after y_host is complete.
 *   - Every function returns NASR_OK (0) or a negative nasr_status; the text
 *     of the last error on a handle is available from nasr_last_error().
 *     No C++ exception crosses this boundary.
 *   - A handle is bound to one device and is not thread-safe.
 *   - There is no CPU fallback: without a CUDA device nasr_engine_create
 *     fails with NASR_ERR_CUDA.
 */
#ifndef NASR_B200_H
#define NASR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NASR_API __attribute__((visibility("default")))
#else
#define NASR_API
#endif

#define NASR_MAX_BLOCKS 64

typedef enum nasr_status {
  NASR_OK = 0,
  NASR_ERR_INVALID = -1,     /* bad argument / unsupported shape (Python: ValueError)   */
  NASR_ERR_CUDA = -2,        /* CUDA runtime / driver failure     (Python: RuntimeError) */
  NASR_ERR_STATE = -3,       /* call order: cond not set, stream not reset, ...          */
  NASR_ERR_NOMEM = -4        /* device or host allocation failed                         */
} nasr_status;

enum { NASR_ARCH_TCN = 0, NASR_ARCH_GCN = 1 };

/* precision / kernel selection */
enum {
  NASR_PATH_AUTO = 0,        /* tensor-core (tcgen05 split-16) blocks where eligible, else fp32 FFMA */
  NASR_PATH_FP32 = 1,        /* fp32 FFMA kernels only                                              */
  NASR_PATH_TC_GATHER = 2    /* as AUTO, but only the tap-gather tcgen05 kernel (no accumulator ring) */
};

/*
 * Model descriptor = the constructor arguments of the reference classes.
 *   TCN(n_channels, n_layers, dilation_growth, in_ch, out_ch, kernel_size, cond_dim)
 *   GCN(in_ch, out_ch, n_blocks, n_channels, dilation_growth, kernel_size, cond_dim)
 * dilations[i] = dilation_growth ** i is spelled out so that WaveNet-style
 * schedules (src/nasr/networks/wavenet.py:159-169) fit the same descriptor.
 */
typedef struct nasr_model_desc {
  int32_t arch;                       /* NASR_ARCH_TCN | NASR_ARCH_GCN                       */
  int32_t n_blocks;                   /* <= NASR_MAX_BLOCKS                                  */
  int32_t in_ch, out_ch;
  int32_t n_channels;                 /* C; GCN conv width is 2C                             */
  int32_t kernel_size;
  int32_t cond_dim;
  int32_t has_film;                   /* TCN: cond_dim > 0 (tcn.py:63-64); GCN: 1 (gcn.py:45) */
  int32_t final_tanh;                 /* GCN: 1 (gcn.py:145-146); TCN: 0                     */
  int32_t path;                       /* NASR_PATH_*                                         */
  float   bn_eps;                     /* 1e-5, nn.BatchNorm1d default                        */
  int32_t dilations[NASR_MAX_BLOCKS];
} nasr_model_desc;

typedef struct nasr_engine nasr_engine;

/*
 * Number of floats nasr_engine_create expects in `weights`, and their order.
 * Per block i (state_dict keys of the reference module):
 *   blocks.i.conv.conv.weight [W, Cin, k]   (W = C for TCN, 2C for GCN)
 *   blocks.i.conv.conv.bias   [W]
 *   if has_film: blocks.i.film.adaptor.weight [2W, cond_dim], .adaptor.bias [2W],
 *                blocks.i.film.bn.weight [W], .bn.bias [W],
 *                .bn.running_mean [W], .bn.running_var [W]
 *   if TCN:      blocks.i.act.weight [1]
 *   blocks.i.res.weight [C, Cin]
 * then out_net.weight [out_ch, C].
 */
NASR_API size_t nasr_weight_count(const nasr_model_desc* desc);

/* weights: host pointer, nasr_weight_count(desc) floats. device: CUDA ordinal. */
NASR_API int nasr_engine_create(const nasr_model_desc* desc, const float* weights_host,
                       size_t n_weights, int device, nasr_engine** out);
NASR_API void nasr_engine_destroy(nasr_engine* e);

/* Human-readable text of the last failure on this handle (or of the last
 * failed nasr_engine_create when e == NULL). Never NULL. */
NASR_API const char* nasr_last_error(const nasr_engine* e);

/* Fold conv bias + eval BatchNorm + FiLM(cond) into per-(clip, channel)
 * scale/shift for every block. cond_dev: [B, cond_dim] (may be NULL when
 * cond_dim == 0). Must precede nasr_forward / nasr_forward_chunk with the same B. */
NASR_API int nasr_set_cond(nasr_engine* e, const float* cond_dev, int B, void* stream);

/* One-shot forward with zero history (reference TCN/GCN.forward). */
NASR_API int nasr_forward(nasr_engine* e, const float* x_dev, float* y_dev, int B, int64_t T,
                 void* stream);

/* nasr_forward that is correct for ANY input range, like the reference forward (tcn.py:150-155 has no range limit):
 * runs the call, waits for `stream`, and if an inter-block activation left the fp16 range of the tensor-core path
 * (see nasr_saturated) runs it again on the fp32 kernels before returning. *redone (may be NULL) = 1 if it did.
 * This is what the Python forward uses for device tensors; nasr_forward stays fully asynchronous. */
NASR_API int nasr_forward_checked(nasr_engine* e, const float* x_dev, float* y_dev, int B, int64_t T,
                         void* stream, int* redone);
/* how many calls on this handle were redone on the fp32 kernels so far (nasr_forward_host, nasr_forward_checked) */
NASR_API int64_t nasr_sat_fallbacks(const nasr_engine* e);

/* nasr_forward with a CUDA-event pair around every block launch on `stream`;
 * waits for completion and writes the n_blocks device durations (ms) to block_ms.
 * Used by bench.py for the live per-kernel roofline numbers. */
NASR_API int nasr_forward_profiled(nasr_engine* e, const float* x_dev, float* y_dev, int B, int64_t T,
                          void* stream, float* block_ms);

/* The tcgen05 path keeps inter-block activations as fp16 hi + fp16 lo pairs; a value beyond
 * +-65504 is clamped and flagged. nasr_forward_host checks the flag itself and transparently
 * redoes the call on the fp32 kernels. After nasr_forward / nasr_forward_chunk (asynchronous)
 * call this to find out: waits for `stream`, returns 1 if the last forward saturated (then
 * re-create the engine with path = NASR_PATH_FP32), 0 if not, < 0 on error. */
NASR_API int nasr_saturated(nasr_engine* e, void* stream);

/* Host-buffer forward: H2D(x, cond) -> set_cond -> forward -> D2H(y), all on
 * `stream`, then waits for completion. cond_host may be NULL when cond_dim == 0. */
NASR_API int nasr_forward_host(nasr_engine* e, const float* x_host, const float* cond_host,
                      float* y_host, int B, int64_t T, void* stream);

/* Streaming: per-block history of the last (k-1)*d input rows is carried
 * across calls (wrapper.py:14-30). reset zero-fills it for B clips. */
NASR_API int nasr_stream_reset(nasr_engine* e, int B, void* stream);
NASR_API int nasr_forward_chunk(nasr_engine* e, const float* x_dev, float* y_dev, int B,
                       int64_t T_chunk, void* stream);

/* A single block on its own: x [B, Cin, T] -> y [B, C, T] (zero history). */
NASR_API int nasr_block_forward(nasr_engine* e, int block, const float* x_dev, float* y_dev,
                       int B, int64_t T, void* stream);

/*
 * Post-processing of make_inference on the device (inference.py:70-78), independent of any engine handle:
 *   out = y / max|y|  ->  lfilter(out, a_coeffs, b_coeffs) per row (zero initial state, clamped to [-1, 1] when
 *   clamp != 0, as torchaudio.functional.lfilter does by default)  ->  out / max|out|.
 * y_dev, out_dev: [rows, T] fp32 on the current device (may alias); b_coeffs / a_coeffs: 3 host floats each, the
 * values torchaudio.functional.highpass_biquad passes to lfilter (un-normalised, a_coeffs[0] != 0).
 * workspace_dev: nasr_postprocess_workspace_bytes(rows, T) bytes of device memory. Asynchronous on `stream`.
 * The recursion runs as a chunked scan in fp64 (the reference's sequential fp32 recursion carries ~1e-3 of
 * rounding noise for the 20 Hz filter).
 */
NASR_API size_t nasr_postprocess_workspace_bytes(int rows, int64_t T);
NASR_API int nasr_postprocess(const float* y_dev, float* out_dev, int rows, int64_t T, const float* b_coeffs,
                     const float* a_coeffs, int clamp, void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * Evaluation metrics of evaluate_model on the device (eval.py:38-40,118-121), one pass over (pred, target):
 *   out_dev[0] = torch.nn.L1Loss()(pred, target)            mean |pred - target| over all elements
 *   out_dev[1] = auraloss.time.ESRLoss()(pred, target)      mean over rows of sum (t - p)^2 / (sum t^2 + 1e-8)
 *   out_dev[2] = auraloss.time.DCLoss()(pred, target)       mean over rows of mean(t - p)^2 / (mean t^2 + 1e-8)
 * pred_dev / target_dev: [rows, T] fp32 (rows = batch x channels); out_dev: 3 doubles on the device.
 * (auraloss 0.4.0 is a third-party dependency that is absent here: its published formulas are restated, parity
 * unpinned. The mel-scaled multi-resolution STFT loss of eval.py:41-49 is not built.)
 */
NASR_API size_t nasr_eval_metrics_workspace_bytes(int rows);
NASR_API int nasr_eval_metrics(const float* pred_dev, const float* target_dev, int rows, int64_t T, double* out_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * RT60 of an impulse response by Schroeder integration (tools/rt60.py:49-70): energy[i] = sum_{j >= i} h[j]^2,
 * truncated before its last nonzero sample, in dB relative to energy[0]; t_5 / t_decay = first samples below -5 dB /
 * -decay_db; rt60 = (60 / decay_db) * (t_decay - t_5), 0 when a crossing does not exist (the reference's except branch).
 * out_dev: 4 doubles {rt60 seconds, i_5db, i_decay, i_nz} (-1 = not found). The reference sums in fp32; here fp64.
 */
NASR_API size_t nasr_rt60_workspace_bytes(int64_t n);
NASR_API int nasr_rt60(const float* h_dev, int64_t n, double sample_rate, double decay_db, double* out_dev,
              void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * Full linear convolution out[i] = sum_j a[j] * b[i - j], i in [0, n + m - 1), fp64 multiply-adds: the
 * scipy.signal.convolve(..., method="direct") of measure_model_ir (tools/ir_model.py:138-140) as a hand-written
 * direct (O(n m)) kernel. a_dev [n], b_dev [m], out_dev [n + m - 1]: doubles on the current device.
 */
NASR_API int nasr_convolve_full(const double* a_dev, int64_t n, const double* b_dev, int64_t m, double* out_dev,
                       void* stream);

/* Bytes of device workspace the engine holds / would hold for (B, T). */
NASR_API size_t nasr_workspace_bytes(const nasr_engine* e, int B, int64_t T);

/* Receptive field in samples (tcn.py:157-164, gcn.py:149-160). */
NASR_API int64_t nasr_receptive_field(const nasr_engine* e);

/* Introspection used by bench.py / tests. */
NASR_API int64_t nasr_launch_count(const nasr_engine* e);       /* kernels launched by this handle so far   */
NASR_API int nasr_block_path(const nasr_engine* e, int block);  /* 0 = fp32 FFMA, 1 = tcgen05 tap-gather, 2 = tcgen05 accumulator-ring kernel */
NASR_API const char* nasr_version(void);

/* dev only: per-CTA timeline stamps of the last ring-kernel launch made with NASR_RB_DBG=8 (tools/ring_timeline.py) */
NASR_API int nasr_debug_ring_stamps(unsigned long long* host, int max_ctas);
NASR_API int nasr_debug_ring_steps(unsigned long long* host);   /* same launch: per-step stamps of CTA 0, host[4][64] */
NASR_API int nasr_debug_toep_stamps(unsigned long long* host, int n);   /* NASR_TOEP_DBG=8, tools/toep_timeline.py */
/* dev / tests: host-side launch plan of the accumulator-ring kernel (no device needed); out16 = {mode, G, L, n, S, NP,
 * spans_per_strip, total_spans, grid, stages, NS, NW, tmem_cols, smem_bytes, n_grp, rext}; 0 on success */
NASR_API int nasr_debug_ring_plan(int arch, int k, int d, int B, int64_t T, int64_t in_row0, int sm_count, int64_t* out16);

#ifdef __cplusplus
}
#endif
#endif /* NASR_B200_H */
