/*
 * msda_b200.h -- C ABI of libmsda_b200.so: multi-scale deformable attention
 * forward + backward, hand-written CUDA for sm_100a (NVIDIA B200).
 *
 * This is the drop-in boundary for the reference's only native component.
 * The two compute entry points replace, one for one, the two functions the
 * reference exports from its pybind11 extension `MultiScaleDeformableAttention`
 *   (/root/reference/models/ops/src/vision.cpp:13-16):
 *
 *   msda_forward   <->  ms_deform_attn_forward   (src/ms_deform_attn.h:21-39,
 *                        src/cuda/ms_deform_attn_cuda.cu:20-80)
 *   msda_backward  <->  ms_deform_attn_backward  (src/ms_deform_attn.h:42-61,
 *                        src/cuda/ms_deform_attn_cuda.cu:83-153)
 *
 * Differences a binder has to know about:
 *   - plain device pointers and sizes instead of at::Tensor; the CALLER
 *     allocates outputs (the reference allocates with at::zeros, :54,:121-123).
 *     Outputs need NOT be zero-initialised: every element is written exactly
 *     once, including grad_value.
 *   - the stream is an explicit argument (the reference takes
 *     at::cuda::getCurrentCUDAStream(), :65,:135); launches are asynchronous,
 *     the library never synchronises the stream or the device.
 *   - errors are RETURNED (status code + msda_last_error()); the reference
 *     swallows launch failures with a printf (ms_deform_im2col_cuda.cuh:948-952).
 *   - spatial_shapes / level_start_index are consumed on the device as int64,
 *     exactly as the reference builds them
 *     (/root/reference/models/deformable_transformer.py:164-165); no host copy.
 *   - backward needs a scratch workspace (size from
 *     msda_backward_workspace_bytes) because grad_value is accumulated
 *     deterministically instead of with floating-point atomics.
 *
 * Tensor layouts (contiguous, row-major), names as in the reference:
 *   value             [N][S][M][D]          value_dtype
 *   spatial_shapes    [L][2] int64 (H, W)   device memory
 *   level_start_index [L]    int64          device memory
 *   sampling_loc      [N][Lq][M][L][P][2]   aux_dtype, (x, y) normalised to [0,1]
 *   attn_weight       [N][Lq][M][L][P]      aux_dtype
 *   output, grad_output [N][Lq][M*D]        value_dtype
 *   grad_value        like value; grad_sampling_loc / grad_attn_weight like
 *                     sampling_loc / attn_weight
 *
 * dtypes: value_dtype in {F32, BF16, F16, F64}; aux_dtype is either equal to
 * value_dtype or F32 (the natural autocast mix: bf16 values, fp32 locations
 * and weights).  Arithmetic is fp32 (fp64 for F64).
 *
 * im2col_step: kept for signature parity.  The reference processes
 * min(N, im2col_step) frames per launch and rejects N not divisible by that
 * (ms_deform_attn_cuda.cu:50-52); this library always uses one launch and
 * keeps the divisibility check so error behaviour matches.
 *
 * Thread safety: no global mutable state besides the thread-local error
 * string; concurrent calls on different streams are independent as long as
 * their workspaces are distinct.
 */
#ifndef MSDA_B200_H
#define MSDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_VERSION 101 /* major*100 + minor */

enum msda_dtype { MSDA_F32 = 0, MSDA_BF16 = 1, MSDA_F16 = 2, MSDA_F64 = 3 };

enum msda_status {
    MSDA_OK = 0,
    MSDA_ERR_INVALID_ARGUMENT = 1, /* null pointer, non-positive size, bad dtype pair */
    MSDA_ERR_IM2COL_STEP = 2,      /* N % min(N, im2col_step) != 0 (reference :52,:119) */
    MSDA_ERR_WORKSPACE = 3,        /* workspace missing or too small */
    MSDA_ERR_CUDA = 4,             /* a CUDA runtime call or launch failed */
    MSDA_ERR_UNSUPPORTED = 5       /* shape outside what the kernels index (see msda_last_error) */
};

/* flags for msda_*_ex: tuning / A-B switches, never needed for correctness */
#define MSDA_FLAG_PYRAMID_TILES 1u /* A/B: 8x16 pixel query tiles when the queries are the pyramid's pixels */
#define MSDA_FLAG_GENERIC 2u       /* force the any-D scalar kernels */
#define MSDA_FLAG_ATOMIC_GRAD_VALUE 4u /* bench-only: fp32 red.global scatter (NOT deterministic) */
#define MSDA_FLAG_BF16_VEC4 8u     /* A/B: 64-byte bf16 rows on 8 lanes x 64 bit instead of 4 lanes x 128 bit */
#define MSDA_FLAG_BIN_KERNEL 32u   /* A/B: grad_value through the one-kernel bin pass (msda_bwd_bin.cuh: entries sorted and
                                      summed in shared memory, half the index traffic) instead of rank-sort + row walker;
                                      measured slower on B200, kept for comparison */
#define MSDA_FLAG_DIRECT_SPLIT 64u  /* A/B: decoder-shaped backward as memset + sample-gradient kernel + direct gather
                                      (three launches) instead of the one kernel that does all three */
#define MSDA_FLAG_UNORDERED 128u    /* opt-in: skip the rank sort of the inverse index.  grad_value is then summed in the
                                       order the index entries happened to be written (integer-atomic slots), i.e. it may
                                       differ in the last bits from run to run -- what the reference's atomicAdd scatter does
                                       all the time (ms_deform_im2col_cuda.cuh:116-153).  grad_sampling_loc and
                                       grad_attn_weight are unaffected.  Off by default: the default is bit-reproducible */
#define MSDA_FLAG_WALK_DENSE 16u   /* force the inverse-index pipeline for grad_value even when the call is small enough
                                      for the direct shared-memory gather (decoder-shaped calls, Lq * P <= 512) */

int msda_version(void);

/* Message of the last non-zero status returned on this thread ("" if none). */
const char *msda_last_error(void);

/* replaces ms_deform_attn_forward (vision.cpp:14) */
int msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                 const void *sampling_loc, const void *attn_weight, void *output,
                 int N, int S, int M, int D, int L, int Lq, int P,
                 int value_dtype, int aux_dtype, int im2col_step, void *cuda_stream);

int msda_forward_ex(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                    const void *sampling_loc, const void *attn_weight, void *output,
                    int N, int S, int M, int D, int L, int Lq, int P,
                    int value_dtype, int aux_dtype, int im2col_step, void *cuda_stream, unsigned flags);

/* Index handoff (optional, saves a pass in the backward).  A forward that will be followed by
 * a backward on the same inputs can count, on the way, how many samples fall into every
 * sub-bin of the inverse index the backward's grad_value gather uses; it leaves the exclusive
 * scan of those counts in `index` (msda_index_bytes bytes of device memory, 16-byte aligned,
 * owned by the caller, to be kept unchanged until the backward).  msda_backward_indexed then
 * skips its own counting pass and CONSUMES the buffer (it advances the offsets in place; pass a
 * copy to run a second backward).  With index == NULL both calls behave like the plain ones.
 * The reference's autograd function has no such state (it saves inputs only,
 * ms_deform_attn_func.py:27); the index is a pure function of the saved inputs.
 * msda_index_bytes returns 0 for calls with so few queries per frame (4 * Lq * P <= 2048: decoder
 * cross-attention) that the backward keeps no index at all; an index passed for such a call is ignored. */
size_t msda_index_bytes(int N, int S, int M, int D, int L, int Lq, int P);

int msda_forward_indexed(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                         const void *sampling_loc, const void *attn_weight, void *output,
                         void *index, size_t index_bytes,
                         int N, int S, int M, int D, int L, int Lq, int P,
                         int value_dtype, int aux_dtype, int im2col_step, void *cuda_stream, unsigned flags);

int msda_backward_indexed(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                          const void *sampling_loc, const void *attn_weight, const void *grad_output,
                          void *grad_value, void *grad_sampling_loc, void *grad_attn_weight,
                          void *workspace, size_t workspace_bytes, void *index, size_t index_bytes,
                          int N, int S, int M, int D, int L, int Lq, int P,
                          int value_dtype, int aux_dtype, int im2col_step, void *cuda_stream, unsigned flags);

/* Forward with the module's elementwise prologue fused in
 * (/root/reference/models/ops/modules/ms_deform_attn.py:99-106, 2-d reference points):
 *   attn_weight  = softmax over the L*P logits of every (query, head)
 *   sampling_loc = reference_points[n][q][l] + sampling_offsets / (W_l, H_l)
 * reference_points [N][Lq][L][2] fp32; sampling_offsets [N][Lq][M][L][P][2] and attn_logits
 * [N][Lq][M][L*P] in in_dtype (the value dtype or F32); the two results are written as fp32 to
 * sampling_loc_out / attn_weight_out (the module returns them, the backward reads them).
 * value_padding_mask (may be NULL): [N][S] bytes, non-zero on padded pixels -- the module's
 *   value = value.masked_fill(input_padding_mask[..., None], 0)      (ms_deform_attn.py:96-97)
 * without the pass over `value`: a padded pixel's row counts as zero wherever a sample touches it, and the backward
 * (which must be given the same mask) leaves its grad_value row zero.
 * Tile-kernel shapes with L*P <= 16 only: anything else returns MSDA_ERR_UNSUPPORTED and the caller
 * keeps the unfused sequence.  `index` as in msda_forward_indexed (may be NULL). */
int msda_forward_fused(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const unsigned char *value_padding_mask, const void *reference_points, const void *sampling_offsets, const void *attn_logits,
                       void *output, void *sampling_loc_out, void *attn_weight_out,
                       void *index, size_t index_bytes,
                       int N, int S, int M, int D, int L, int Lq, int P,
                       int value_dtype, int in_dtype, int im2col_step, void *cuda_stream, unsigned flags);

/* Backward of msda_forward_fused: like msda_backward_indexed on the saved sampling_loc / attn_weight (the fp32
 * tensors the fused forward wrote), but the chain rule of the prologue is applied on the way out --
 *   grad_sampling_offsets = grad_sampling_loc / (W_l, H_l)
 *   grad_attn_logits      = attn_weight * (grad_attn_weight - sum_over_L*P(attn_weight * grad_attn_weight))
 * written in aux_dtype with the layouts of sampling_loc / attn_weight.  (The gradient of the reference points, when
 * needed, is the sum of grad_sampling_offsets * (W_l, H_l) over heads and points.)  Same shape support as
 * msda_forward_fused; anything else returns MSDA_ERR_UNSUPPORTED. */
int msda_backward_fused(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const unsigned char *value_padding_mask, const void *sampling_loc, const void *attn_weight, const void *grad_output,
                        void *grad_value, void *grad_sampling_offsets, void *grad_attn_logits,
                        void *workspace, size_t workspace_bytes, void *index, size_t index_bytes,
                        int N, int S, int M, int D, int L, int Lq, int P,
                        int value_dtype, int aux_dtype, int im2col_step, void *cuda_stream, unsigned flags);

/* The same without ever materialising sampling_loc / attn_weight (the encoder drops them:
 * /root/reference/models/deformable_transformer.py:255 keeps only the first of the module's three results).
 * msda_forward_fused accepts NULL for BOTH sampling_loc_out and attn_weight_out; this backward then takes what that
 * forward took -- reference_points, the raw sampling_offsets and attn_logits (in_dtype) -- recomputes the prologue in
 * its staging threads and writes grad_sampling_offsets / grad_attn_logits in in_dtype (bf16 under a bf16 layer: no
 * fp32 round trip of 384 values per query).  Needs the index the forward left (index != NULL); only for calls that
 * keep one (msda_index_bytes != 0). */
int msda_backward_fused_raw(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const unsigned char *value_padding_mask, const void *reference_points, const void *sampling_offsets, const void *attn_logits,
                            const void *grad_output, void *grad_value, void *grad_sampling_offsets, void *grad_attn_logits,
                            void *workspace, size_t workspace_bytes, void *index, size_t index_bytes,
                            int N, int S, int M, int D, int L, int Lq, int P,
                            int value_dtype, int in_dtype, int im2col_step, void *cuda_stream, unsigned flags);

/* Bytes of device scratch msda_backward needs for this problem size. */
size_t msda_backward_workspace_bytes(int N, int S, int M, int D, int L, int Lq, int P,
                                     int value_dtype, int aux_dtype);

/* replaces ms_deform_attn_backward (vision.cpp:15) */
int msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                  const void *sampling_loc, const void *attn_weight, const void *grad_output,
                  void *grad_value, void *grad_sampling_loc, void *grad_attn_weight,
                  void *workspace, size_t workspace_bytes,
                  int N, int S, int M, int D, int L, int Lq, int P,
                  int value_dtype, int aux_dtype, int im2col_step, void *cuda_stream);

int msda_backward_ex(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                     const void *sampling_loc, const void *attn_weight, const void *grad_output,
                     void *grad_value, void *grad_sampling_loc, void *grad_attn_weight,
                     void *workspace, size_t workspace_bytes,
                     int N, int S, int M, int D, int L, int Lq, int P,
                     int value_dtype, int aux_dtype, int im2col_step, void *cuda_stream, unsigned flags);

/* Number of kernels the last msda_forward / msda_backward on this thread launched
 * (bench.py's gpu_launches is counted from this, not guessed). */
int msda_last_launch_count(void);

/* Per-kernel device timing for benchmarks: when enabled (thread-local), every kernel the
 * library launches is bracketed by CUDA events on the launching stream.  msda_profile_read
 * waits for the recorded events, writes up to `cap` durations in milliseconds to `ms` and the
 * kernel names, newline separated, to `names`; returns the number of records.  Enabling or
 * disabling clears the records.  Not for production use (two event records per launch). */
void msda_profile_enable(int on);
int msda_profile_read(char *names, size_t names_cap, float *ms, int cap);

/* ---- the memory-bound glue of the encoder layer around the op ------------------------------
 * The reference layer does  src = norm(src + branch)  twice per layer as separate PyTorch kernels
 * (/root/reference/models/deformable_transformer.py:253-263 and :247-251: add, nn.LayerNorm; in the backward the
 * LayerNorm input gradient, two parameter-gradient reductions and the add of the residual branch).  These two entry
 * points do each direction in one pass over the rows; they have no counterpart in the reference's FFI (it calls
 * ATen), the host side is neurips2023_soc_b200/modules/encoder_layer.py.
 *   out    = LayerNorm(branch + residual) * gamma + beta                 [rows][channels], dtype F32 or BF16
 *   presum = branch + residual (rounded to dtype; may alias `branch`)      kept for the backward
 *   mean, rstd                                                             [rows] fp32
 * backward: grad_in (the gradient of BOTH branch and residual), grad_gamma / grad_beta [channels] fp32, summed in a
 * fixed order (deterministic) through `workspace` (msda_add_layernorm_backward_workspace_bytes).
 * channels must be 256 (d_model of every SOC config); anything else returns MSDA_ERR_UNSUPPORTED. */
int msda_add_layernorm_forward(const void *branch, const void *residual, const float *gamma, const float *beta,
                               void *out, void *presum, float *mean, float *rstd,
                               long long rows, int channels, int dtype, float eps, void *cuda_stream);
size_t msda_add_layernorm_backward_workspace_bytes(long long rows, int channels);
int msda_add_layernorm_backward(const void *grad_out, const void *presum, const float *mean, const float *rstd,
                                const float *gamma, void *grad_in, float *grad_gamma, float *grad_beta,
                                void *workspace, size_t workspace_bytes,
                                long long rows, int channels, int dtype, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H */
