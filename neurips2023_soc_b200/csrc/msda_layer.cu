// msda_layer.cu -- the memory-bound glue of the deformable encoder layer around the op:
// residual add + LayerNorm, forward and backward, one pass each.
//
// The reference layer (/root/reference/models/deformable_transformer.py:253-263, 247-251) does
//     src = norm1(src + dropout(self_attn(...)))        src = norm2(src + dropout(linear2(act(linear1(src)))))
// as separate PyTorch kernels: an add, a LayerNorm forward, and in the backward a LayerNorm input gradient, two
// parameter-gradient reductions and an add for the residual branch -- under bf16 autocast with fp32 round trips in
// between (the residual stream stays fp32).  Here one kernel reads the branch output and the residual once and
// writes the normalised row (bf16 or fp32) plus the pre-norm sum the backward needs; the backward reads the
// upstream gradient and that sum once and writes the input gradient (shared by both branches of the residual) and
// per-CTA partial sums of dgamma / dbeta, which a second tiny kernel adds in a fixed order (deterministic).
//
// d_model = 256 only (every SOC config, configs/*.yaml: d_model 256): a warp owns a row, a lane 8 consecutive
// channels (one 128-bit access in bf16, two in fp32); statistics in fp32 by warp shuffles.  HBM-bound by
// construction: 3 row reads/writes forward, 3 backward.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_host.h"

namespace msda {

constexpr int kC = 256;          // channels per row
constexpr int kLnThreads = 256;  // 8 warps = 8 rows in flight per CTA
constexpr int kCpl = 8;          // channels per lane

template <typename T> struct Row8;
template <> struct Row8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    static __device__ __forceinline__ float round(float x) { return x; }
};
template <> struct Row8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 t;
        uint32_t* w = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = t;
    }
    static __device__ __forceinline__ float round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// y = LayerNorm(a + b) * gamma + beta;  s = a + b (rounded to T: what the backward will see);  mean, rstd per row
template <typename T>
__global__ void __launch_bounds__(kLnThreads) msda_add_layernorm_fwd_kernel(const T* a, const T* __restrict__ b,
                                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                            T* __restrict__ y, T* s_out /* may alias a */, float* __restrict__ mean,
                                                                            float* __restrict__ rstd, const long long rows, const float eps) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (kLnThreads / 32);
    float g[kCpl], bt[kCpl];
    Row8<float>::load(gamma + lane * kCpl, g);
    Row8<float>::load(beta + lane * kCpl, bt);
    for (long long r = (long long)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5); r < rows; r += warps) {
        float va[kCpl], vb[kCpl], s[kCpl];
        Row8<T>::load(a + r * kC + lane * kCpl, va);
        Row8<T>::load(b + r * kC + lane * kCpl, vb);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kCpl; ++i) { s[i] = Row8<T>::round(va[i] + vb[i]); sum += s[i]; }
        const float mu = warp_sum(sum) * (1.f / kC);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < kCpl; ++i) { const float d = s[i] - mu; sq = fmaf(d, d, sq); }
        const float rs = rsqrtf(warp_sum(sq) * (1.f / kC) + eps);
        float o[kCpl];
#pragma unroll
        for (int i = 0; i < kCpl; ++i) o[i] = fmaf((s[i] - mu) * rs, g[i], bt[i]);
        Row8<T>::store(y + r * kC + lane * kCpl, o);
        Row8<T>::store(s_out + r * kC + lane * kCpl, s);
        if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
    }
}

// dx = rstd * (dy*gamma - mean_c(dy*gamma) - xhat * mean_c(dy*gamma*xhat)),  xhat = (s - mean) * rstd;
// per-CTA partials of dgamma = sum_rows dy * xhat and dbeta = sum_rows dy
template <typename T>
__global__ void __launch_bounds__(kLnThreads) msda_add_layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ s,
                                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                            const float* __restrict__ gamma, T* __restrict__ dx,
                                                                            float* __restrict__ part, const long long rows) {
    __shared__ float red[2][kLnThreads / 32][kC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warps = (long long)gridDim.x * (kLnThreads / 32);
    float g[kCpl];
    Row8<float>::load(gamma + lane * kCpl, g);
    float dg[kCpl], db[kCpl];
#pragma unroll
    for (int i = 0; i < kCpl; ++i) dg[i] = db[i] = 0.f;
    for (long long r = (long long)blockIdx.x * (kLnThreads / 32) + warp; r < rows; r += warps) {
        float vdy[kCpl], vs[kCpl];
        Row8<T>::load(dy + r * kC + lane * kCpl, vdy);
        Row8<T>::load(s + r * kC + lane * kCpl, vs);
        const float mu = mean[r], rs = rstd[r];
        float xh[kCpl], t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < kCpl; ++i) {
            xh[i] = (vs[i] - mu) * rs;
            const float dyg = vdy[i] * g[i];
            t1 += dyg;
            t2 = fmaf(dyg, xh[i], t2);
            dg[i] = fmaf(vdy[i], xh[i], dg[i]);
            db[i] += vdy[i];
        }
        const float m1 = warp_sum(t1) * (1.f / kC), m2 = warp_sum(t2) * (1.f / kC);
        float o[kCpl];
#pragma unroll
        for (int i = 0; i < kCpl; ++i) o[i] = rs * (vdy[i] * g[i] - m1 - xh[i] * m2);
        Row8<T>::store(dx + r * kC + lane * kCpl, o);
    }
#pragma unroll
    for (int i = 0; i < kCpl; ++i) {
        red[0][warp][lane * kCpl + i] = dg[i];
        red[1][warp][lane * kCpl + i] = db[i];
    }
    __syncthreads();
    // thread c sums channel c over the CTA's warps in a fixed order
    const int c = threadIdx.x;
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < kLnThreads / 32; ++w) { sg += red[0][w][c]; sb += red[1][w][c]; }
    part[(size_t)blockIdx.x * 2 * kC + c] = sg;
    part[(size_t)blockIdx.x * 2 * kC + kC + c] = sb;
}

// dgamma / dbeta = the partials of all CTAs, added in a fixed order (deterministic): a CTA takes 32 channels of one
// of the two vectors, its 8 warps each add every 8th partial, thread-serially, and warp 0 adds the 8 sums in order
constexpr int kPgWarps = 8;
__global__ void __launch_bounds__(32 * kPgWarps) msda_layernorm_param_grad_kernel(const float* __restrict__ part, const int nparts,
                                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float red[kPgWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool beta = blockIdx.y == 1;
    const int c = blockIdx.x * 32 + lane;
    float acc = 0.f;
    for (int p = warp; p < nparts; p += kPgWarps) acc += part[(size_t)p * 2 * kC + (beta ? kC : 0) + c];
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kPgWarps; ++w) t += red[w][lane];
        (beta ? dbeta : dgamma)[c] = t;
    }
}

int ln_grid(long long rows) {
    const long long want = (rows + kLnThreads / 32 - 1) / (kLnThreads / 32);
    const long long cap = (long long)msda_host::num_sms() * 8;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace msda

using namespace msda;

extern "C" {

size_t msda_add_layernorm_backward_workspace_bytes(long long rows, int channels) {
    if (rows <= 0 || channels != kC) return 0;
    return (size_t)ln_grid(rows) * 2 * kC * sizeof(float);
}

int msda_add_layernorm_forward(const void* branch, const void* residual, const float* gamma, const float* beta, void* out,
                               void* presum, float* mean, float* rstd, long long rows, int channels, int dtype, float eps,
                               void* cuda_stream) {
    if (!branch || !residual || !gamma || !beta || !out || !presum || !mean || !rstd)
        return msda_host::fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    if (rows <= 0) return msda_host::fail(MSDA_ERR_INVALID_ARGUMENT, "non-positive row count %lld", rows);
    if (channels != kC) return msda_host::fail(MSDA_ERR_UNSUPPORTED, "add + LayerNorm is built for %d channels, got %d", kC, channels);
    if (dtype != MSDA_F32 && dtype != MSDA_BF16) return msda_host::fail(MSDA_ERR_UNSUPPORTED, "add + LayerNorm takes fp32 or bf16 rows");
    if (!(aligned16(branch) && aligned16(residual) && aligned16(gamma) && aligned16(beta) && aligned16(out) && aligned16(presum)))
        return msda_host::fail(MSDA_ERR_INVALID_ARGUMENT, "tensors must be 16-byte aligned");
    const msda_host::DeviceGuard guard(branch);
    if (!guard.ok()) return msda_host::fail(MSDA_ERR_INVALID_ARGUMENT, "branch must be device memory (a CUDA tensor)");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    msda_host::prof_begin(st, "msda_add_layernorm_fwd_kernel");
    if (dtype == MSDA_F32)
        msda_add_layernorm_fwd_kernel<float><<<ln_grid(rows), kLnThreads, 0, st>>>(
            static_cast<const float*>(branch), static_cast<const float*>(residual), gamma, beta, static_cast<float*>(out),
            static_cast<float*>(presum), mean, rstd, rows, eps);
    else
        msda_add_layernorm_fwd_kernel<__nv_bfloat16><<<ln_grid(rows), kLnThreads, 0, st>>>(
            static_cast<const __nv_bfloat16*>(branch), static_cast<const __nv_bfloat16*>(residual), gamma, beta,
            static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(presum), mean, rstd, rows, eps);
    msda_host::prof_end(st);
    MSDA_LAUNCHED("msda_add_layernorm_fwd_kernel");
    return MSDA_OK;
}

int msda_add_layernorm_backward(const void* grad_out, const void* presum, const float* mean, const float* rstd, const float* gamma,
                                void* grad_in, float* grad_gamma, float* grad_beta, void* workspace, size_t workspace_bytes,
                                long long rows, int channels, int dtype, void* cuda_stream) {
    if (!grad_out || !presum || !mean || !rstd || !gamma || !grad_in || !grad_gamma || !grad_beta)
        return msda_host::fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    if (rows <= 0) return msda_host::fail(MSDA_ERR_INVALID_ARGUMENT, "non-positive row count %lld", rows);
    if (channels != kC) return msda_host::fail(MSDA_ERR_UNSUPPORTED, "add + LayerNorm is built for %d channels, got %d", kC, channels);
    if (dtype != MSDA_F32 && dtype != MSDA_BF16) return msda_host::fail(MSDA_ERR_UNSUPPORTED, "add + LayerNorm takes fp32 or bf16 rows");
    const size_t need = msda_add_layernorm_backward_workspace_bytes(rows, channels);
    if (!workspace || workspace_bytes < need)
        return msda_host::fail(MSDA_ERR_WORKSPACE, "workspace of %zu bytes required, got %zu", need, workspace ? workspace_bytes : (size_t)0);
    if (!(aligned16(grad_out) && aligned16(presum) && aligned16(gamma) && aligned16(grad_in) && aligned16(workspace)))
        return msda_host::fail(MSDA_ERR_INVALID_ARGUMENT, "tensors must be 16-byte aligned");
    const msda_host::DeviceGuard guard(grad_out);
    if (!guard.ok()) return msda_host::fail(MSDA_ERR_INVALID_ARGUMENT, "grad_out must be device memory (a CUDA tensor)");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    float* part = static_cast<float*>(workspace);
    const int grid = ln_grid(rows);
    msda_host::prof_begin(st, "msda_add_layernorm_bwd_kernel");
    if (dtype == MSDA_F32)
        msda_add_layernorm_bwd_kernel<float><<<grid, kLnThreads, 0, st>>>(static_cast<const float*>(grad_out), static_cast<const float*>(presum),
                                                                           mean, rstd, gamma, static_cast<float*>(grad_in), part, rows);
    else
        msda_add_layernorm_bwd_kernel<__nv_bfloat16><<<grid, kLnThreads, 0, st>>>(
            static_cast<const __nv_bfloat16*>(grad_out), static_cast<const __nv_bfloat16*>(presum), mean, rstd, gamma,
            static_cast<__nv_bfloat16*>(grad_in), part, rows);
    msda_host::prof_end(st);
    MSDA_LAUNCHED("msda_add_layernorm_bwd_kernel");
    msda_host::prof_begin(st, "msda_layernorm_param_grad_kernel");
    msda_layernorm_param_grad_kernel<<<dim3(kC / 32, 2), 32 * kPgWarps, 0, st>>>(part, grid, grad_gamma, grad_beta);
    msda_host::prof_end(st);
    MSDA_LAUNCHED("msda_layernorm_param_grad_kernel");
    return MSDA_OK;
}

}  // extern "C"
