// msda_backward.cuh -- backward kernels.
//
// Replaces ms_deformable_col2im_gpu_kernel_* and ms_deform_attn_col2im_bilinear
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:87-159,301-920): one
// 32-thread block per (query, head), two block barriers and a serial 32-term sum by
// thread 0 per sample, and 4*D scalar fp32 atomicAdd per sample into grad_value
// (non-deterministic summation order, grad_value must be pre-zeroed).
//
// Here the backward is split by what is being produced:
//
//  A. grad_sampling_loc / grad_attn_weight  (msda_bwd_sample_*): same tile/descriptor
//     structure as the forward.  A group of G lanes owns a (query, head) row, keeps its
//     grad_output slice in registers and forms, per sample, the four corner dot products
//     d_k = <g, v_k> over its channels.  The 16 samples x 4 corners partials are reduced
//     across the G lanes with a log2(G)-step exchange in which every lane gives away half
//     of what it holds (warp shuffles, no shared memory, no barriers), after which lane r
//     holds the finished d_1..d_4 of 16/G samples and writes their
//        grad_attn = sum_k w_k d_k
//        grad_x    = W * a * (hh (d2-d1) + lh (d4-d3))
//        grad_y    = H * a * (hw (d3-d1) + lw (d4-d2))              (cuh:116-158)
//     with coalesced stores.
//
//  B. grad_value, deterministically and without floating-point atomics.  The scatter is
//     turned into a gather through an inverse index keyed by the sample's top-left corner
//     ("bin" = (h_lo+1, w_lo+1) in a (H+1)x(W+1) grid per level, frame and head):
//        count : every accepted sample takes a slot in its bin (integer atomics only --
//                integer addition commutes, so counts are exact and order-free)
//        scan  : per (frame, head) exclusive scan of the bin counts
//        fill  : every sample writes a 16-byte entry {query|sample id, lh, lw, a}
//        sort  : each bin's entries are put in ascending id order, which makes the
//                summation order below a pure function of the inputs
//        gather: a group of G lanes owns one grad_value row (pixel, head); the pixel is
//                corner 1/2/3/4 of the samples binned at (y+1,x+1)/(y+1,x)/(y,x+1)/(y,x);
//                it walks those four lists in order, accumulates
//                w_k * a * grad_output[query] in registers and writes the row once with a
//                128-bit store.  No zero-fill of grad_value is needed.
//     Results are bit-identical from run to run.
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr uint32_t kRejected = 0xffffffffu;
constexpr int kBigBin = 32;  // bins with more entries are sorted by msda_bin_sort_big_kernel

template <typename CT> struct Entry;
template <> struct __align__(16) Entry<float> {
    uint32_t id;
    float lh, lw, a;
};
template <> struct __align__(16) Entry<double> {
    uint32_t id, pad;
    double lh, lw, a;
};

// =========================================================================================
// A. grad_sampling_loc, grad_attn_weight
// =========================================================================================

// One exchange step of the cross-lane reduction: lanes whose bit DIST is clear keep the
// lower half of v[0..2*HALF) and receive the partner's lower half; the others the upper.
template <int HALF, int DIST>
__device__ __forceinline__ void exchange_halve(float* v, const bool upper) {
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const float keep = upper ? v[i + HALF] : v[i];
        const float send = upper ? v[i] : v[i + HALF];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, DIST);
    }
}

template <int G>
__device__ __forceinline__ void reduce_scatter(float* v, const int gl) {
    // v holds 4*kSC partials; afterwards v[0 .. 4*kSC/G) are complete sums for
    // indices [gl*4*kSC/G, (gl+1)*4*kSC/G).
    constexpr int NV = 4 * kSC;
    if constexpr (G >= 16) exchange_halve<NV * 8 / G, 8>(v, gl & 8);
    if constexpr (G >= 8) exchange_halve<NV * 4 / G, 4>(v, gl & 4);
    if constexpr (G >= 4) exchange_halve<NV * 2 / G, 2>(v, gl & 2);
    if constexpr (G >= 2) exchange_halve<NV * 1 / G, 1>(v, gl & 1);
}

template <int NG>
struct BwdSmem {
    int4 off[NG * kDescStride];
    float4 geo[NG * kDescStride];   // lh, lw, a, level index (int bits)
};

template <typename T, typename TA, int G, bool ATOMIC>
__global__ void __launch_bounds__(kThreads, 2) msda_bwd_sample_tile_kernel(const Params p, const int rounds) {
    constexpr int VEC = Elem<T>::kVec;
    constexpr int NG = kThreads / G;
    constexpr int DPT = NG * kSC / kThreads;
    constexpr int NSL = kSC / G;               // finished samples per lane after the reduction
    static_assert(G <= kSC, "a lane must end up with at least one whole sample");

    __shared__ Level lv[kMaxLevels];
    __shared__ TileMap tm;
    __shared__ int s_sb, s_sq;
    __shared__ BwdSmem<NG> sm;

    const int tile_q = NG * rounds;
    load_levels(p, lv, &s_sb, &s_sq);
    build_tile_map(p, lv, s_sq, tile_q, &tm);

    const T* __restrict__ value = static_cast<const T*>(p.value);
    const T* __restrict__ gout = static_cast<const T*>(p.grad_out);
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const TA* __restrict__ attn = static_cast<const TA*>(p.attn);
    TA* __restrict__ gloc = static_cast<TA*>(p.grad_loc);
    TA* __restrict__ gattn = static_cast<TA*>(p.grad_attn);

    const int tid = threadIdx.x;
    const int grp = tid / G, gl = tid % G;
    const int st_s = tid % kSC, st_j0 = tid / kSC;
    const int row_elems = p.M * p.D;
    const int total_tiles = p.N * p.M * tm.qtiles;

    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const Tile tl = decode_tile(p, lv, &tm, t, tile_q);
        const size_t frame_off = (size_t)tl.n * p.S * row_elems + tl.m * p.D + gl * VEC;
        const T* vbase = value + frame_off;

        for (int r = 0; r < rounds; ++r) {
            const int q_mine = tile_query(p, &tm, tl, r * NG + grp);
            const size_t qm_mine = ((size_t)tl.n * p.Lq + (q_mine < 0 ? 0 : q_mine)) * p.M + tl.m;
            float g[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) g[i] = 0.f;
            if (q_mine >= 0) load_vec(gout + qm_mine * p.D + gl * VEC, g);

            for (int c0 = 0; c0 < p.LP; c0 += kSC) {
                __syncthreads();
                const int sg = c0 + st_s;
                const bool s_ok = sg < p.LP;
                const int l = s_ok ? sg / p.P : 0;
                const Level L_ = lv[l];
#pragma unroll
                for (int k = 0; k < DPT; ++k) {
                    const int j = st_j0 + k * (kThreads / kSC);
                    const int q = tile_query(p, &tm, tl, r * NG + j);
                    int4 o = make_int4(-1, -1, -1, -1);
                    float4 d = make_float4(0.f, 0.f, 0.f, __int_as_float(l));
                    if (s_ok && q >= 0) {
                        const size_t si = (((size_t)tl.n * p.Lq + q) * p.M + tl.m) * p.LP + sg;
                        const XY<float> xy = load_xy(loc + 2 * si);
                        const Sample<float> s = locate(xy.x, xy.y, L_.H, L_.W);
                        if (s.ok) {
                            int pix[4];
                            corner_pixels(s, L_, pix);
                            o.x = pix[0] < 0 ? -1 : pix[0] * row_elems;
                            o.y = pix[1] < 0 ? -1 : pix[1] * row_elems;
                            o.z = pix[2] < 0 ? -1 : pix[2] * row_elems;
                            o.w = pix[3] < 0 ? -1 : pix[3] * row_elems;
                            d.x = s.lh; d.y = s.lw;
                            d.z = Elem<TA>::to_f(__ldg(attn + si));
                        }
                    }
                    sm.off[j * kDescStride + st_s] = o;
                    sm.geo[j * kDescStride + st_s] = d;
                }
                __syncthreads();

                // per-lane partial corner dot products for the chunk's 16 samples
                float part[4 * kSC];
#pragma unroll
                for (int s = 0; s < kSC; ++s) {
                    const int4 o = sm.off[grp * kDescStride + s];
                    float v0[VEC], v1[VEC], v2[VEC], v3[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) v0[i] = v1[i] = v2[i] = v3[i] = 0.f;
                    if (o.x >= 0) load_vec(vbase + o.x, v0);
                    if (o.y >= 0) load_vec(vbase + o.y, v1);
                    if (o.z >= 0) load_vec(vbase + o.z, v2);
                    if (o.w >= 0) load_vec(vbase + o.w, v3);
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        d0 = fmaf(g[i], v0[i], d0);
                        d1 = fmaf(g[i], v1[i], d1);
                        d2 = fmaf(g[i], v2[i], d2);
                        d3 = fmaf(g[i], v3[i], d3);
                    }
                    part[4 * s + 0] = d0; part[4 * s + 1] = d1;
                    part[4 * s + 2] = d2; part[4 * s + 3] = d3;
                    if constexpr (ATOMIC) {
                        // bench-only A/B arm: the reference's scatter with 128-bit fp32 reductions
                        // (order-dependent rounding => NOT deterministic; never the default)
                        static_assert(!ATOMIC || VEC == 4, "atomic arm is fp32 only");
                        const float4 d = sm.geo[grp * kDescStride + s];
                        const float hh = 1.f - d.x, hw = 1.f - d.y;
                        const float w[4] = {hh * hw * d.z, hh * d.y * d.z, d.x * hw * d.z, d.x * d.y * d.z};
                        const int oo[4] = {o.x, o.y, o.z, o.w};
                        float* gv = static_cast<float*>(p.grad_value) + frame_off;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (oo[k] >= 0 && q_mine >= 0)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gv + oo[k]),
                                             "f"(w[k] * g[0]), "f"(w[k] * g[1]), "f"(w[k] * g[2]), "f"(w[k] * g[3])
                                             : "memory");
                    }
                }
                reduce_scatter<G>(part, gl);

                if (q_mine >= 0) {
#pragma unroll
                    for (int i = 0; i < NSL; ++i) {
                        const int s = gl * NSL + i;
                        const int sgo = c0 + s;
                        if (sgo < p.LP) {
                            const float4 d = sm.geo[grp * kDescStride + s];
                            const int ls = __float_as_int(d.w);
                            const float Hf = (float)lv[ls].H, Wf = (float)lv[ls].W;
                            const float lh = d.x, lw = d.y, a = d.z;
                            const float hh = 1.f - lh, hw = 1.f - lw;
                            const float d0 = part[4 * i], d1 = part[4 * i + 1], d2 = part[4 * i + 2], d3 = part[4 * i + 3];
                            const float ga = hh * hw * d0 + hh * lw * d1 + lh * hw * d2 + lh * lw * d3;
                            const float gx = hh * (d1 - d0) + lh * (d3 - d2);
                            const float gy = hw * (d2 - d0) + lw * (d3 - d1);
                            const size_t si = qm_mine * p.LP + sgo;
                            gattn[si] = Elem<TA>::from_f(ga);
                            gloc[2 * si] = Elem<TA>::from_f(Wf * a * gx);
                            gloc[2 * si + 1] = Elem<TA>::from_f(Hf * a * gy);
                        }
                    }
                }
            }
        }
    }
}

// Any-D / any-dtype fallback for part A: one warp per (frame, query, head); lanes stride
// the channels, three partial sums per sample are reduced with warp shuffles.
template <typename T, typename TA, typename CT>
__global__ void __launch_bounds__(kThreads) msda_bwd_sample_generic_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    load_levels(p, lv, &s_sb, &s_sq);
    const T* __restrict__ value = static_cast<const T*>(p.value);
    const T* __restrict__ gout = static_cast<const T*>(p.grad_out);
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const TA* __restrict__ attn = static_cast<const TA*>(p.attn);
    TA* __restrict__ gloc = static_cast<TA*>(p.grad_loc);
    TA* __restrict__ gattn = static_cast<TA*>(p.grad_attn);
    const int lane = threadIdx.x & 31;
    const size_t warps = (size_t)gridDim.x * (kThreads / 32);
    const size_t rows = (size_t)p.N * p.Lq * p.M;
    const size_t row_elems = (size_t)p.M * p.D;
    for (size_t qm = (size_t)blockIdx.x * (kThreads / 32) + threadIdx.x / 32; qm < rows; qm += warps) {
        const int m = (int)(qm % p.M);
        const size_t n = qm / p.M / p.Lq;
        const T* vb = value + n * p.S * row_elems + (size_t)m * p.D;
        const T* gr = gout + qm * p.D;
        for (int l = 0; l < p.L; ++l) {
            const Level L_ = lv[l];
            for (int pt = 0; pt < p.P; ++pt) {
                const size_t si = qm * p.LP + l * p.P + pt;
                const XY<CT> xy = load_xy(loc + 2 * si);
                const Sample<CT> s = locate(xy.x, xy.y, L_.H, L_.W);
                CT ga = 0, gx = 0, gy = 0;
                CT a = 0;
                if (s.ok) {  // warp-uniform
                    a = (CT)Elem<TA>::to_f(attn[si]);
                    int pix[4];
                    corner_pixels(s, L_, pix);
                    const CT hh = (CT)1 - s.lh, hw = (CT)1 - s.lw;
                    for (int c = lane; c < p.D; c += 32) {
                        CT v[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            v[k] = pix[k] >= 0 ? (CT)Elem<T>::to_f(vb[(size_t)pix[k] * row_elems + c]) : (CT)0;
                        const CT gc = (CT)Elem<T>::to_f(gr[c]);
                        ga += gc * (hh * hw * v[0] + hh * s.lw * v[1] + s.lh * hw * v[2] + s.lh * s.lw * v[3]);
                        gx += gc * (hh * (v[1] - v[0]) + s.lh * (v[3] - v[2]));
                        gy += gc * (hw * (v[2] - v[0]) + s.lw * (v[3] - v[1]));
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        ga += __shfl_xor_sync(0xffffffffu, ga, d);
                        gx += __shfl_xor_sync(0xffffffffu, gx, d);
                        gy += __shfl_xor_sync(0xffffffffu, gy, d);
                    }
                }
                if (lane == 0) {
                    gattn[si] = Elem<TA>::from_f(ga);
                    gloc[2 * si] = Elem<TA>::from_f((CT)L_.W * a * gx);
                    gloc[2 * si + 1] = Elem<TA>::from_f((CT)L_.H * a * gy);
                }
            }
        }
    }
}

// =========================================================================================
// B. grad_value: count -> scan -> fill -> sort -> gather
// =========================================================================================

struct SampleRef {
    size_t si;     // linear sample index
    int n, q, m, sg, l;
};

__device__ __forceinline__ SampleRef sample_ref(const Params& p, size_t si) {
    SampleRef r;
    r.si = si;
    r.sg = (int)(si % p.LP);
    const size_t qm = si / p.LP;
    r.m = (int)(qm % p.M);
    const size_t nq = qm / p.M;
    r.q = (int)(nq % p.Lq);
    r.n = (int)(nq / p.Lq);
    r.l = r.sg / p.P;
    return r;
}

template <typename TA, typename CT>
__global__ void __launch_bounds__(kThreads) msda_bin_count_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    load_levels(p, lv, &s_sb, &s_sq);
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const size_t total = (size_t)p.N * p.Lq * p.M * p.LP;
    for (size_t si = (size_t)blockIdx.x * blockDim.x + threadIdx.x; si < total; si += (size_t)gridDim.x * blockDim.x) {
        const SampleRef r = sample_ref(p, si);
        const Level L_ = lv[r.l];
        const XY<CT> xy = load_xy(loc + 2 * si);
        const Sample<CT> s = locate(xy.x, xy.y, L_.H, L_.W);
        uint32_t slot = kRejected;
        if (s.ok) {
            const int bin = L_.bin_start + (s.h_lo + 1) * (L_.W + 1) + (s.w_lo + 1);
            slot = atomicAdd(p.bin_off + (size_t)(r.n * p.M + r.m) * (p.sb_max + 1) + bin, 1u);
        }
        p.pos[si] = slot;
    }
}

// One CTA per (frame, head): in-place exclusive scan of the bin counts; bins with more
// than kBigBin entries are appended to the big-bin list.
__global__ void __launch_bounds__(1024) msda_bin_scan_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    __shared__ uint32_t warp_tot[32];
    load_levels(p, lv, &s_sb, &s_sq);
    const int nm = blockIdx.x;
    uint32_t* data = p.bin_off + (size_t)nm * (p.sb_max + 1);
    const int SB = s_sb;
    const int ipt = (SB + blockDim.x - 1) / blockDim.x;
    const int beg = min(SB, (int)threadIdx.x * ipt), end = min(SB, beg + ipt);
    uint32_t sum = 0;
    for (int i = beg; i < end; ++i) {
        const uint32_t c = data[i];
        sum += c;
        if (c > (uint32_t)kBigBin) {
            const uint32_t k = atomicAdd(p.big_bins, 1u);
            if (k < (uint32_t)p.big_cap) {
                p.big_bins[1 + 2 * k] = (uint32_t)nm;
                p.big_bins[2 + 2 * k] = (uint32_t)i;
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = warp_tot[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        warp_tot[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    uint32_t run = warp_tot[wid] + inc - sum;
    for (int i = beg; i < end; ++i) {
        const uint32_t c = data[i];
        data[i] = run;
        run += c;
    }
    if (threadIdx.x == blockDim.x - 1) data[SB] = run;  // last thread's running total == grand total
}

template <typename TA, typename CT>
__global__ void __launch_bounds__(kThreads) msda_bin_fill_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    load_levels(p, lv, &s_sb, &s_sq);
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const TA* __restrict__ attn = static_cast<const TA*>(p.attn);
    Entry<CT>* __restrict__ entries = static_cast<Entry<CT>*>(p.entries);
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const size_t total = (size_t)p.N * p.Lq * p.M * p.LP;
    for (size_t si = (size_t)blockIdx.x * blockDim.x + threadIdx.x; si < total; si += (size_t)gridDim.x * blockDim.x) {
        const uint32_t slot = p.pos[si];
        if (slot == kRejected) continue;
        const SampleRef r = sample_ref(p, si);
        const Level L_ = lv[r.l];
        const XY<CT> xy = load_xy(loc + 2 * si);
        const Sample<CT> s = locate(xy.x, xy.y, L_.H, L_.W);
        const int bin = L_.bin_start + (s.h_lo + 1) * (L_.W + 1) + (s.w_lo + 1);
        const size_t nm = (size_t)r.n * p.M + r.m;
        const uint32_t base = p.bin_off[nm * (p.sb_max + 1) + bin];
        Entry<CT> e;
        e.id = ((uint32_t)r.q << p.id_shift) | (uint32_t)r.sg;
        e.lh = s.lh; e.lw = s.lw;
        e.a = (CT)Elem<TA>::to_f(attn[si]);
        entries[nm * per_nm + base + slot] = e;
    }
}

// ---- sort ---------------------------------------------------------------------------
// All compare-exchanges are ascending (lower index keeps the smaller id); each merge
// starts with a mirror step (partner = i ^ (k-1)) followed by half-cleaners
// (partner = i ^ j).  Missing elements (index >= cnt) act as +inf and never move.

template <typename CT, int WIDTH>
__device__ __forceinline__ void sort_in_lanes(Entry<CT>* base, const uint32_t cnt, const int sub, const uint32_t mask) {
    // `sub` = lane index inside a WIDTH-lane segment; one entry per lane.
    uint32_t key = sub < (int)cnt ? base[sub].id : 0xffffffffu;
    int src = sub;
#pragma unroll
    for (int k = 2; k <= WIDTH; k <<= 1) {
#pragma unroll
        for (int j = k - 1; j > 0; j = (j == k - 1) ? (k >> 2) : (j >> 1)) {
            const uint32_t ok = __shfl_xor_sync(mask, key, j, WIDTH);
            const int os = __shfl_xor_sync(mask, src, j, WIDTH);
            const bool lower = (sub & ((j == k - 1) ? (k >> 1) : j)) == 0;
            const bool take = lower ? (ok < key) : (ok > key);
            if (take) { key = ok; src = os; }
            if (j == 1 || (j == k - 1 && k == 2)) break;
        }
    }
    // lane `sub` now knows which original slot belongs at position `sub`
    Entry<CT> e;
    if (sub < (int)cnt) e = base[src];
    __syncwarp(mask);
    if (sub < (int)cnt) base[sub] = e;
}

template <typename CT>
__global__ void __launch_bounds__(kThreads) msda_bin_sort_small_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    load_levels(p, lv, &s_sb, &s_sq);
    Entry<CT>* __restrict__ entries = static_cast<Entry<CT>*>(p.entries);
    const int SB = s_sb;
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const size_t nbins = (size_t)p.N * p.M * SB;
    const int lane = threadIdx.x & 31;
    // phase 1: 8-lane segments, bins of 2..8 entries
    {
        const int sub = lane & 7;
        const uint32_t mask = 0xffu << (lane & 24);
        const size_t segs = (size_t)gridDim.x * (kThreads / 8);
        for (size_t b = (size_t)blockIdx.x * (kThreads / 8) + threadIdx.x / 8; b < nbins; b += segs) {
            const size_t nm = b / SB;
            const int bin = (int)(b - nm * SB);
            const uint32_t* off = p.bin_off + nm * (p.sb_max + 1) + bin;
            const uint32_t beg = off[0], cnt = off[1] - beg;
            if (cnt >= 2 && cnt <= 8) sort_in_lanes<CT, 8>(entries + nm * per_nm + beg, cnt, sub, mask);
        }
    }
    // phase 2: whole warps, bins of 9..32 entries
    {
        const size_t warps = (size_t)gridDim.x * (kThreads / 32);
        const size_t span = (nbins + 31) / 32;
        for (size_t w = (size_t)blockIdx.x * (kThreads / 32) + threadIdx.x / 32; w < span; w += warps) {
            const size_t b = w * 32 + lane;
            uint32_t beg = 0, cnt = 0;
            size_t nm = 0;
            if (b < nbins) {
                nm = b / SB;
                const int bin = (int)(b - nm * SB);
                const uint32_t* off = p.bin_off + nm * (p.sb_max + 1) + bin;
                beg = off[0];
                cnt = off[1] - beg;
            }
            uint32_t todo = __ballot_sync(0xffffffffu, cnt > 8 && cnt <= (uint32_t)kBigBin);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t bbeg = __shfl_sync(0xffffffffu, beg, src);
                const uint32_t bcnt = __shfl_sync(0xffffffffu, cnt, src);
                const size_t bnm = __shfl_sync(0xffffffffu, (unsigned long long)nm, src);
                sort_in_lanes<CT, 32>(entries + bnm * per_nm + bbeg, bcnt, lane, 0xffffffffu);
            }
        }
    }
}

// Bins with more than kBigBin entries: one CTA per bin, bitonic network over a
// shared-memory copy (or in place in global memory when the bin does not fit).
template <typename CT>
__global__ void __launch_bounds__(kThreads) msda_bin_sort_big_kernel(const Params p) {
    constexpr int CAP = 32768 / (int)sizeof(Entry<CT>);
    __shared__ Entry<CT> buf[CAP];
    Entry<CT>* __restrict__ entries = static_cast<Entry<CT>*>(p.entries);
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const uint32_t nbig = min(p.big_bins[0], (uint32_t)p.big_cap);
    for (uint32_t i = blockIdx.x; i < nbig; i += gridDim.x) {
        const size_t nm = p.big_bins[1 + 2 * i];
        const uint32_t bin = p.big_bins[2 + 2 * i];
        const uint32_t* off = p.bin_off + nm * (p.sb_max + 1) + bin;
        const uint32_t beg = off[0], cnt = off[1] - beg;
        Entry<CT>* g = entries + nm * per_nm + beg;
        const bool in_smem = cnt <= (uint32_t)CAP;
        Entry<CT>* a = in_smem ? buf : g;
        __syncthreads();
        if (in_smem)
            for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) buf[k] = g[k];
        __syncthreads();
        uint32_t n2 = 1;
        while (n2 < cnt) n2 <<= 1;
        for (uint32_t k = 2; k <= n2; k <<= 1) {
            for (uint32_t j = k - 1; j > 0; j = (j == k - 1) ? (k >> 2) : (j >> 1)) {
                const uint32_t half = (j == k - 1) ? (k >> 1) : j;  // distance class of this step
                for (uint32_t t = threadIdx.x; t < n2 / 2; t += blockDim.x) {
                    const uint32_t lo = ((t & ~(half - 1)) << 1) | (t & (half - 1));
                    const uint32_t hi = lo ^ j;
                    if (hi < cnt) {
                        const Entry<CT> x = a[lo], y = a[hi];
                        if (x.id > y.id) { a[lo] = y; a[hi] = x; }
                    }
                }
                __syncthreads();
                if (j == 1 || (j == k - 1 && k == 2)) break;
            }
        }
        if (in_smem)
            for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) g[k] = buf[k];
    }
}

// ---- gather ---------------------------------------------------------------------------

template <typename T, int G>
__global__ void __launch_bounds__(kThreads) msda_grad_value_tile_kernel(const Params p) {
    constexpr int VEC = Elem<T>::kVec;
    constexpr int NG = kThreads / G;           // grad_value rows per CTA pass
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    load_levels(p, lv, &s_sb, &s_sq);

    const T* __restrict__ gout = static_cast<const T*>(p.grad_out);
    T* __restrict__ gval = static_cast<T*>(p.grad_value);
    const Entry<float>* __restrict__ entries = static_cast<const Entry<float>*>(p.entries);

    const int tid = threadIdx.x, lane = tid & 31;
    const int grp = tid / G, gl = tid % G;
    const uint32_t gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - gl));
    const int chunks = (p.S + NG - 1) / NG;
    const int total_tiles = p.N * p.M * chunks;
    const size_t per_nm = (size_t)p.Lq * p.LP;

    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        // coarse levels (end of the pixel order) carry the longest lists: issue them first
        const int n = t / (chunks * p.M);
        const int r = t - n * chunks * p.M;
        const int chunk = chunks - 1 - r / p.M;
        const int m = r % p.M;
        const int s = chunk * NG + grp;
        if (s >= p.S) continue;               // whole group idle (group-uniform)
        const size_t nm = (size_t)n * p.M + m;
        const uint32_t* off = p.bin_off + nm * (p.sb_max + 1);
        const Entry<float>* ent = entries + nm * per_nm;
        const T* gbase = gout + ((size_t)n * p.Lq * p.M + m) * p.D + gl * VEC;
        const size_t qstride = (size_t)p.M * p.D;

        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

        int l = -1;
        for (int k = 0; k < p.L; ++k)
            if (s >= lv[k].start && s < lv[k].start + lv[k].H * lv[k].W) { l = k; break; }
        if (l >= 0) {
            const Level L_ = lv[l];
            const int y = (s - L_.start) / L_.W, x = (s - L_.start) % L_.W;
            const int b_hi = L_.bin_start + (y + 1) * (L_.W + 1) + x;   // (y+1, x): corner 2; +1: corner 1
            const int b_lo = L_.bin_start + y * (L_.W + 1) + x;         // (y,   x): corner 4; +1: corner 3
            const int bins[4] = {b_hi + 1, b_hi, b_lo + 1, b_lo};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t beg = off[bins[c]], end = off[bins[c] + 1];
                for (uint32_t e0 = beg; e0 < end; e0 += G) {
                    Entry<float> mine;
                    mine.id = 0; mine.lh = mine.lw = mine.a = 0.f;
                    if (e0 + gl < end) mine = ent[e0 + gl];
                    const float hh = 1.f - mine.lh, hw = 1.f - mine.lw;
                    const float wsel = (c == 0) ? hh * hw : (c == 1) ? hh * mine.lw : (c == 2) ? mine.lh * hw : mine.lh * mine.lw;
                    const float wa_mine = wsel * mine.a;
                    const int nb = min((uint32_t)G, end - e0);
                    for (int e = 0; e < nb; ++e) {
                        const uint32_t id = __shfl_sync(gmask, mine.id, e, G);
                        const float wa = __shfl_sync(gmask, wa_mine, e, G);
                        const uint32_t q = id >> p.id_shift;
                        float gv[VEC];
                        load_vec(gbase + (size_t)q * qstride, gv);
#pragma unroll
                        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(wa, gv[i], acc[i]);
                    }
                }
            }
        }
        store_vec(gval + ((size_t)n * p.S + s) * qstride + (size_t)m * p.D + gl * VEC, acc);
    }
}

// Any-D / any-dtype fallback: one warp per grad_value row, lanes stride the channels.
template <typename T, typename CT>
__global__ void __launch_bounds__(kThreads) msda_grad_value_generic_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    load_levels(p, lv, &s_sb, &s_sq);
    const T* __restrict__ gout = static_cast<const T*>(p.grad_out);
    T* __restrict__ gval = static_cast<T*>(p.grad_value);
    const Entry<CT>* __restrict__ entries = static_cast<const Entry<CT>*>(p.entries);
    const int lane = threadIdx.x & 31;
    const size_t warps = (size_t)gridDim.x * (kThreads / 32);
    const size_t rows = (size_t)p.N * p.S * p.M;
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const size_t qstride = (size_t)p.M * p.D;
    for (size_t row = (size_t)blockIdx.x * (kThreads / 32) + threadIdx.x / 32; row < rows; row += warps) {
        const int m = (int)(row % p.M);
        const size_t ns = row / p.M;
        const int s = (int)(ns % p.S);
        const size_t n = ns / p.S;
        const size_t nm = n * p.M + m;
        const uint32_t* off = p.bin_off + nm * (p.sb_max + 1);
        const Entry<CT>* ent = entries + nm * per_nm;
        const T* gbase = gout + (n * p.Lq * p.M + m) * p.D;
        T* dst = gval + row * p.D;
        int l = -1;
        for (int k = 0; k < p.L; ++k)
            if (s >= lv[k].start && s < lv[k].start + lv[k].H * lv[k].W) { l = k; break; }
        for (int c0 = 0; c0 < p.D; c0 += 32) {
            const int ch = c0 + lane;
            CT acc = 0;
            if (l >= 0) {
                const Level L_ = lv[l];
                const int y = (s - L_.start) / L_.W, x = (s - L_.start) % L_.W;
                const int b_hi = L_.bin_start + (y + 1) * (L_.W + 1) + x;
                const int b_lo = L_.bin_start + y * (L_.W + 1) + x;
                const int bins[4] = {b_hi + 1, b_hi, b_lo + 1, b_lo};
                for (int c = 0; c < 4; ++c) {
                    const uint32_t beg = off[bins[c]], end = off[bins[c] + 1];
                    for (uint32_t e = beg; e < end; ++e) {
                        const Entry<CT> en = ent[e];
                        const CT hh = (CT)1 - en.lh, hw = (CT)1 - en.lw;
                        const CT w = (c == 0) ? hh * hw : (c == 1) ? hh * en.lw : (c == 2) ? en.lh * hw : en.lh * en.lw;
                        const uint32_t q = en.id >> p.id_shift;
                        if (ch < p.D) acc += w * en.a * (CT)Elem<T>::to_f(gbase[(size_t)q * qstride + ch]);
                    }
                }
            }
            if (ch < p.D) dst[ch] = Elem<T>::from_f(acc);
        }
    }
}

}  // namespace msda
