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
//        sort  : each sub-bin's entries are put in ascending id order (every entry counts the
//                smaller ids of its sub-bin), which makes the summation order below a pure
//                function of the inputs
//        gather: a group of G lanes owns one grad_value row (pixel, head); the pixel is
//                corner 1/2/3/4 of the samples binned at (y+1,x+1)/(y+1,x)/(y,x+1)/(y,x);
//                it walks those four lists in order, accumulates
//                w_k * a * grad_output[query] in registers and writes the row once with a
//                128-bit store.  No zero-fill of grad_value is needed.
//     Results are bit-identical from run to run.
#pragma once

#include <cooperative_groups.h>

#include "msda_tiles.cuh"

namespace msda {


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

template <int G, int NV>
__device__ __forceinline__ void reduce_scatter(float* v, const int gl) {
    // v holds NV partials per lane; afterwards v[0 .. NV/G) are complete sums for
    // indices [gl*NV/G, (gl+1)*NV/G).
    static_assert(NV % G == 0, "every lane must end up with a whole number of values");
    if constexpr (G >= 16) exchange_halve<NV * 8 / G, 8>(v, gl & 8);
    if constexpr (G >= 8) exchange_halve<NV * 4 / G, 4>(v, gl & 4);
    if constexpr (G >= 4) exchange_halve<NV * 2 / G, 2>(v, gl & 2);
    if constexpr (G >= 2) exchange_halve<NV * 1 / G, 1>(v, gl & 1);
}

#ifndef MSDA_BWD_MIN_BLOCKS
#define MSDA_BWD_MIN_BLOCKS 3
#endif

// FILL: the staging threads also write each accepted sample's entry into the inverse index of
// part B (one integer atomic + one 16-byte store, overlapped with the gather).  ATOMIC:
// bench-only A/B arm that scatters grad_value with 128-bit fp32 reductions like the reference
// does with scalar ones.
// CHAIN: the chain rule of the module's prologue (/root/reference/models/ops/modules/ms_deform_attn.py:99-106) is
// applied on the way out: grad_loc / grad_attn then hold the gradients of the RAW sampling offsets and attention
// logits --  d/d offset = grad_loc / (W, H) = a * (dX, dY);  d/d logit_s = a_s * (g_s - sum_t a_t g_t)  (softmax) --
// for calls whose L*P samples are one chunk (the fused prologue's own restriction).
// RAWIN (with CHAIN): `loc` / `attn` are the RAW sampling offsets / attention logits and p.ref the reference points: the
// prologue is recomputed in the staging threads exactly as the fused forward does (msda_tiles.cuh: fused_prologue), so
// that neither the sampling locations nor the attention weights are ever materialised (msda_backward_fused_raw).
template <typename T, typename TA, int VEC, int G, int P, bool FILL, bool ATOMIC, int ROWB = 0, bool CHAIN = false, bool RAWIN = false>
__global__ void __launch_bounds__(kThreads, MSDA_BWD_MIN_BLOCKS) msda_bwd_sample_tile_kernel(const Params p, const int rounds) {
    using TS = TileShape<G>;
    constexpr int NG = TS::NG;
    constexpr int LPC = (P >= kSC) ? 1 : kSC / P;
    constexpr int PPC = (P >= kSC) ? kSC : P;
    static_assert(!ATOMIC || (VEC == 4 && sizeof(T) == 4), "atomic arm is fp32 only");
    static_assert(!RAWIN || CHAIN, "the raw-input variant always applies the chain rule");

    __shared__ Level lv[kMaxLevels];
    __shared__ TileMap tm;
    __shared__ int s_sb, s_sq;
    __shared__ uint4 desc[2][NG * kDescStride];
    // CHAIN: d/d weight of this lane's samples, kept unrounded until the (query, head)'s softmax sum is known
    __shared__ float s_ga[CHAIN ? 4 : 1][CHAIN ? kThreads : 1];      // (level step, sample of the lane): at most 4 per lane

    const int tile_q = NG * rounds;
    load_levels(p, lv, &s_sb, &s_sq);
    build_tile_map(p, lv, s_sq, tile_q, &tm);

    const T* __restrict__ value = static_cast<const T*>(p.value);
    const T* __restrict__ gout = static_cast<const T*>(p.grad_out);
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const TA* __restrict__ attn = static_cast<const TA*>(p.attn);
    TA* __restrict__ gloc = static_cast<TA*>(p.grad_loc);
    TA* __restrict__ gattn = static_cast<TA*>(p.grad_attn);

    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    const int row_elems = p.M * p.D;
    const int total_tiles = p.N * p.M * tm.qtiles;

    Work cur{(int)blockIdx.x, 0, 0};
    if (cur.t >= total_tiles) return;
    Tile tl = decode_tile(p, lv, &tm, cur.t, tile_q);
    Staged<TS::DPT> st;
    stage_load<TA, G, P, RAWIN>(st, p, &tm, tl, cur, loc, attn);
    stage_build<G, P, (FILL ? kIndexFill : kIndexNone), RAWIN, CHAIN>(st, p, lv, tl, cur, desc[0], index_usable(p, s_sb));
    __syncthreads();

    int buf = 0;
    using R = typename Raw<sizeof(T) * VEC>::type;
    constexpr bool BF16_DOT = std::is_same<T, __nv_bfloat16>::value && MSDA_BF16_MIXED_FMA;
    R graw = {};                                     // this row's grad_output slice: as loaded (BF16_DOT) ...
    float g[VEC];                                    // ... or unpacked
    while (true) {
        const Work nxt = next_work(cur, rounds, p.LP);
        const bool has_next = nxt.t < total_tiles;
        Tile ntl = tl;
        if (has_next) {
            if (nxt.t != cur.t) ntl = decode_tile(p, lv, &tm, nxt.t, tile_q);
            stage_load<TA, G, P, RAWIN>(st, p, &tm, ntl, nxt, loc, attn);
        }

        const int q_mine = tile_query(p, &tm, tl, cur.r * NG + grp);
        const size_t qm_mine = ((size_t)tl.n * p.Lq + (q_mine < 0 ? 0 : q_mine)) * p.M + tl.m;
        if (cur.c0 == 0) {
            if constexpr (BF16_DOT) {
                graw = load_raw_if<T, VEC>(q_mine >= 0, gout + qm_mine * p.D + gl * VEC);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) g[i] = 0.f;
                if (q_mine >= 0) load_row<T, VEC>(gout + qm_mine * p.D + gl * VEC, g);
            }
        }
        const size_t frame_off = (size_t)tl.n * p.S * row_elems + tl.m * p.D;
        const char* fb = reinterpret_cast<const char*>(value) + frame_off * sizeof(T);
        asm volatile("" : "+l"(fb));         // keep the base in registers (ptxas would re-read it per load)
        const uint32_t rowb = ROWB ? (uint32_t)ROWB : (uint32_t)row_elems * (uint32_t)sizeof(T);
        const uint32_t lane_off = (uint32_t)(gl * VEC * sizeof(T));
        const uint4* drow = desc[buf] + grp * kDescStride;
        const int l0 = cur.c0 / P;

        // One level (PPC samples) at a time: per-lane partial corner dot products d_k = <g, v_k>,
        // reduced across the G lanes by a halving exchange, then finished by the lanes that hold them.
        constexpr int NV = 4 * PPC;                  // partials per lane and level step
        constexpr int CPL = NV / G;                  // finished corner sums per lane
        constexpr int LPS = CPL >= 4 ? 1 : 4 / CPL;  // lanes that share one sample afterwards
        constexpr int SPL = CPL >= 4 ? CPL / 4 : 1;  // samples per lane afterwards
        static_assert(NV % G == 0 && (CPL >= 4 ? CPL % 4 == 0 : 4 % CPL == 0), "unsupported G / P combination");
        float chain_dot = 0.f;                       // CHAIN: this lane's share of sum_t a_t g_t
#pragma unroll 1
        for (int lc = 0; lc < LPC; ++lc) {
            const int l = l0 + lc;
            if (l >= p.L) break;                         // uniform: slots past L*P carry nothing
            const Level& L_ = lv[l];
            const LevelPitch lp = level_pitch(L_, rowb, lane_off);
            const int sbase = (P >= kSC) ? 0 : lc * PPC;
            float part[NV];
#pragma unroll
            for (int pp = 0; pp < PPC; ++pp) {
                const uint4 d = drow[sbase + pp];
                const int32_t o0 = lp.base + (int32_t)((d.x & 0x0fffffffu) * rowb);
                const int32_t o2 = o0 + lp.wrow;
                const char* p0 = fb + o0;                // top-left corner row (sign-extended offset)
                const char* p2 = p0 + lp.wrow;           // the row below
                const R r0 = load_raw_if<T, VEC>(d.x & (1u << 28), reinterpret_cast<const T*>(p0));
                const R r1 = load_raw_if<T, VEC>(d.x & (2u << 28), reinterpret_cast<const T*>(p0 + rowb));
                const R r2 = load_raw_if<T, VEC>(d.x & (4u << 28), reinterpret_cast<const T*>(p2));
                const R r3 = load_raw_if<T, VEC>(d.x & (8u << 28), reinterpret_cast<const T*>(p2 + rowb));
                float v0[VEC], v1[VEC], v2[VEC], v3[VEC];
                unpack_row<T, VEC>(r0, v0);
                unpack_row<T, VEC>(r1, v1);
                unpack_row<T, VEC>(r2, v2);
                unpack_row<T, VEC>(r3, v3);
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                if constexpr (BF16_DOT) {
                    // bf16 x bf16 -> fp32 FMAs straight from the packed registers, channel by channel
                    const uint32_t* gw = reinterpret_cast<const uint32_t*>(&graw);
                    const uint32_t* w0 = reinterpret_cast<const uint32_t*>(&r0);
                    const uint32_t* w1 = reinterpret_cast<const uint32_t*>(&r1);
                    const uint32_t* w2 = reinterpret_cast<const uint32_t*>(&r2);
                    const uint32_t* w3 = reinterpret_cast<const uint32_t*>(&r3);
#pragma unroll
                    for (int i = 0; i < VEC / 2; ++i) {
                        dot2_bf16(d0, gw[i], w0[i]);
                        dot2_bf16(d1, gw[i], w1[i]);
                        dot2_bf16(d2, gw[i], w2[i]);
                        dot2_bf16(d3, gw[i], w3[i]);
                    }
                } else if constexpr (use_packed_fma<T, 1>()) {
                    // even / odd channel partial sums on packed fp32x2 FMAs, added at the end
                    float o0_ = 0.f, o1_ = 0.f, o2_ = 0.f, o3_ = 0.f;
#pragma unroll
                    for (int i = 0; i < VEC; i += 2) {
                        fma2v(d0, o0_, g[i], g[i + 1], v0[i], v0[i + 1]);
                        fma2v(d1, o1_, g[i], g[i + 1], v1[i], v1[i + 1]);
                        fma2v(d2, o2_, g[i], g[i + 1], v2[i], v2[i + 1]);
                        fma2v(d3, o3_, g[i], g[i + 1], v3[i], v3[i + 1]);
                    }
                    d0 += o0_; d1 += o1_; d2 += o2_; d3 += o3_;
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        d0 = fmaf(g[i], v0[i], d0);
                        d1 = fmaf(g[i], v1[i], d1);
                        d2 = fmaf(g[i], v2[i], d2);
                        d3 = fmaf(g[i], v3[i], d3);
                    }
                }
                part[4 * pp + 0] = d0; part[4 * pp + 1] = d1;
                part[4 * pp + 2] = d2; part[4 * pp + 3] = d3;
                if constexpr (ATOMIC) {
                    const float lh = __uint_as_float(d.y), lw = __uint_as_float(d.z), a = __uint_as_float(d.w);
                    const float ah = a * (1.f - lh), al = a * lh, hw = 1.f - lw;
                    const float w[4] = {ah * hw, ah * lw, al * hw, al * lw};
                    char* gvb = reinterpret_cast<char*>(p.grad_value) + frame_off * sizeof(T);
                    const int32_t oo[4] = {o0, o0 + (int32_t)rowb, o2, o2 + (int32_t)rowb};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((d.x >> (28 + k)) & 1u)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gvb + oo[k]),
                                         "f"(w[k] * g[0]), "f"(w[k] * g[1]), "f"(w[k] * g[2]), "f"(w[k] * g[3])
                                         : "memory");
                }
            }
            reduce_scatter<G, NV>(part, gl);

            // Lane gl now holds the complete corner sums with indices [gl*CPL, gl*CPL + CPL): whole
            // samples (CPL >= 4), the top or bottom corner pair of one (CPL == 2), or one corner (CPL == 1).
            //   grad_attn = hh hw d0 + hh lw d1 + lh hw d2 + lh lw d3
            //   dX = hh (d1 - d0) + lh (d3 - d2)        dY = hw (d2 - d0) + lw (d3 - d1)      (cuh:116-158)
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const int s = (gl * CPL) / 4 + i;            // sample inside this level step
                const uint4 d = drow[sbase + s];
                const float lh = __uint_as_float(d.y), lw = __uint_as_float(d.z), a = __uint_as_float(d.w);
                const float hh = 1.f - lh, hw = 1.f - lw;
                float ga, gx, gy;
                if constexpr (CPL >= 4) {
                    const float d0 = part[4 * i], d1 = part[4 * i + 1], d2 = part[4 * i + 2], d3 = part[4 * i + 3];
                    ga = hh * (hw * d0 + lw * d1) + lh * (hw * d2 + lw * d3);
                    gx = hh * (d1 - d0) + lh * (d3 - d2);
                    gy = hw * (d2 - d0) + lw * (d3 - d1);
                } else if constexpr (CPL == 2) {
                    const bool bottom = (gl & 1) != 0;       // odd lanes hold (d2, d3), even lanes (d0, d1)
                    const float rowf = bottom ? lh : hh;
                    const float t = hw * part[0] + lw * part[1];
                    ga = rowf * t;
                    gx = rowf * (part[1] - part[0]);
                    gy = bottom ? t : -t;
                } else {
                    const bool bottom = (gl & 2) != 0, right = (gl & 1) != 0;   // corner k = gl & 3
                    const float rowf = bottom ? lh : hh, colf = right ? lw : hw;
                    const float dv = part[0];
                    ga = rowf * colf * dv;
                    gx = rowf * (right ? dv : -dv);
                    gy = colf * (bottom ? dv : -dv);
                }
                if constexpr (LPS > 1) {
#pragma unroll
                    for (int dd = 1; dd < LPS; dd <<= 1) {
                        ga += __shfl_xor_sync(0xffffffffu, ga, dd);
                        gx += __shfl_xor_sync(0xffffffffu, gx, dd);
                        gy += __shfl_xor_sync(0xffffffffu, gy, dd);
                    }
                }
                const int sgo = cur.c0 + sbase + s;
                if (q_mine >= 0 && ((gl * CPL) & 3) == 0 && sgo < p.LP) {
                    const size_t si = qm_mine * p.LP + sgo;
                    if constexpr (CHAIN) {
                        static_assert(LPC * SPL <= 4, "a lane finishes at most four samples per chunk");
                        s_ga[lc * SPL + i][threadIdx.x] = ga;
                        store_xy(gloc + 2 * si, a * gx, a * gy);
                        chain_dot = fmaf(a, ga, chain_dot);
                    } else {
                        gattn[si] = Elem<TA>::from_f(ga);
                        store_xy(gloc + 2 * si, (float)L_.W * a * gx, (float)L_.H * a * gy);
                    }
                }
            }
        }
        if constexpr (CHAIN) {
            // softmax backward: the lanes that wrote a sample's d/d weight come back to it once the (query, head)'s
            // sum is known (their own stores, re-read in program order)
#pragma unroll
            for (int dd = 1; dd < G; dd <<= 1) chain_dot += __shfl_xor_sync(0xffffffffu, chain_dot, dd, G);   // all lanes are here
#pragma unroll 1
            for (int lc = 0; lc < LPC; ++lc) {
                if (l0 + lc >= p.L) break;
                const int sbase = (P >= kSC) ? 0 : lc * PPC;
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const int s = (gl * CPL) / 4 + i;
                    const int sgo = cur.c0 + sbase + s;
                    if (q_mine >= 0 && ((gl * CPL) & 3) == 0 && sgo < p.LP) {
                        const size_t si = qm_mine * p.LP + sgo;
                        // the descriptor keeps the weight of rejected samples too (stage_build<KEEPW>): their logits
                        // still take  -a_s * sum_t a_t g_t
                        const float a = __uint_as_float(drow[sbase + s].w);
                        gattn[si] = Elem<TA>::from_f(a * (s_ga[lc * SPL + i][threadIdx.x] - chain_dot));
                    }
                }
            }
        }

        if (has_next) stage_build<G, P, (FILL ? kIndexFill : kIndexNone), RAWIN, CHAIN>(st, p, lv, ntl, nxt, desc[buf ^ 1], index_usable(p, s_sb));
        __syncthreads();
        if (!has_next) break;
        cur = nxt;
        tl = ntl;
        buf ^= 1;
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
    if (!index_usable(p, s_sb)) return;
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const size_t total = (size_t)p.N * p.Lq * p.M * p.LP;
    for (size_t si = (size_t)blockIdx.x * blockDim.x + threadIdx.x; si < total; si += (size_t)gridDim.x * blockDim.x) {
        const SampleRef r = sample_ref(p, si);
        const Level L_ = lv[r.l];
        const XY<CT> xy = load_xy(loc + 2 * si);
        const Sample<CT> s = locate(xy.x, xy.y, L_.H, L_.W);
        if (s.ok) {
            const int bin = sub_bin(L_, s.h_lo, s.w_lo, r.q);
            atomicAdd(p.bin_off + (size_t)(r.n * p.M + r.m) * (p.sb_max + 1) + bin + 1, 1u);
        }
    }
}

// One CTA per (frame, head): in-place exclusive scan of the sub-bin counts.  The count of sub-bin b
// sits in slot b+1 of the table row (slot 0 stays 0), so that after the scan slot b+1 holds the
// START of sub-bin b, and after the fill has advanced it by the sub-bin's population it holds its
// END, which is the start of sub-bin b+1: from then on (row[b], row[b+1]) is the range of sub-bin b
// and no second copy of the table is needed.
__global__ void __launch_bounds__(1024) msda_bin_scan_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    __shared__ uint32_t warp_tot[32];
    load_levels(p, lv, &s_sb, &s_sq);
    const int nm = blockIdx.x;
    uint32_t* data = p.bin_off + (size_t)nm * (p.sb_max + 1) + 1;
    const int SB = s_sb;
    // every warp owns one contiguous segment and walks it 32 counts at a time (coalesced), twice: totals first,
    // then the exclusive offsets on top of the scanned warp totals
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int seg = (((SB + 31) / 32) + 31) & ~31;          // per-warp segment, a multiple of 32
    const int beg = min(SB, wid * seg), end = min(SB, beg + seg);
    uint32_t sum = 0;
#pragma unroll 4
    for (int i = beg + lane; i < end; i += 32) sum += data[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) warp_tot[wid] = sum;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = warp_tot[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        warp_tot[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    uint32_t run = warp_tot[wid];
    for (int i0 = beg; i0 < end; i0 += 32) {
        const int i = i0 + lane;
        const uint32_t c = i < end ? data[i] : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (i < end) data[i] = run + inc - c;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
}

template <typename TA, typename CT>
__global__ void __launch_bounds__(kThreads) msda_bin_fill_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    load_levels(p, lv, &s_sb, &s_sq);
    if (!index_usable(p, s_sb)) return;
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const TA* __restrict__ attn = static_cast<const TA*>(p.attn);
    Entry<CT>* __restrict__ entries = static_cast<Entry<CT>*>(p.entries);
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const size_t total = (size_t)p.N * p.Lq * p.M * p.LP;
    for (size_t si = (size_t)blockIdx.x * blockDim.x + threadIdx.x; si < total; si += (size_t)gridDim.x * blockDim.x) {
        const SampleRef r = sample_ref(p, si);
        const Level L_ = lv[r.l];
        const XY<CT> xy = load_xy(loc + 2 * si);
        const Sample<CT> s = locate(xy.x, xy.y, L_.H, L_.W);
        if (!s.ok) continue;
        const int bin = sub_bin(L_, s.h_lo, s.w_lo, r.q);
        const size_t nm = (size_t)r.n * p.M + r.m;
        const uint32_t slot = atomicAdd(p.bin_off + nm * (p.sb_max + 1) + bin + 1, 1u);
        Entry<CT> e;
        e.id = ((uint32_t)r.q << p.id_shift) | (uint32_t)r.sg;
        e.lh = s.lh; e.lw = s.lw;
        e.a = (CT)Elem<TA>::to_f(attn[si]);
        entries[nm * per_nm + slot] = e;
    }
}

// ---- sort ---------------------------------------------------------------------------
// Every sub-bin is put in ascending id order, in place.  A CTA stages a run of consecutive
// sub-bins (their entries are contiguous) in shared memory; then every ENTRY finds its own rank
// inside its sub-bin by counting the smaller ids (a handful of shared-memory reads: sub-bins hold
// about six entries) and is written back to `first slot + rank`.  All lanes do the same amount
// of work whatever the sub-bin sizes are.  Sub-bins with more than kRankMax entries go to the
// big list instead (msda_bin_sort_big_kernel: bitonic network, one CTA per sub-bin).
#ifndef MSDA_RANK_SPAN
#define MSDA_RANK_SPAN 128
#endif
#ifndef MSDA_RANK_STAGE_BYTES
#define MSDA_RANK_STAGE_BYTES 16384
#endif
constexpr int kRankSpan = MSDA_RANK_SPAN;   // sub-bins staged per step (at most)

template <typename CT>
__global__ void __launch_bounds__(kThreads) msda_bin_rank_sort_kernel(const Params p) {
    constexpr int CAP = MSDA_RANK_STAGE_BYTES / (int)sizeof(Entry<CT>);     // entries staged per step
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    __shared__ Entry<CT> buf[CAP];
    __shared__ uint32_t ids[CAP];
    __shared__ uint16_t owner[CAP];
    __shared__ uint32_t soff[kRankSpan + 1];
    __shared__ int s_take;
    load_levels(p, lv, &s_sb, &s_sq);
    Entry<CT>* __restrict__ entries = static_cast<Entry<CT>*>(p.entries);
    const int SB = s_sb;
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const int spans = (SB + kRankSpan - 1) / kRankSpan;
    const int total = p.N * p.M * spans;
    const int tid = threadIdx.x;

    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int nm = item / spans;
        const int b_end = min(SB, (item - nm * spans + 1) * kRankSpan);
        const uint32_t* off = p.bin_off + (size_t)nm * (p.sb_max + 1);
        Entry<CT>* ent = entries + (size_t)nm * per_nm;
        int b = (item - nm * spans) * kRankSpan;
        while (b < b_end) {
            const int n = min(kRankSpan, b_end - b);
            for (int i = tid; i <= n; i += kThreads) soff[i] = off[b + i];
            __syncthreads();
            int take = n;                       // usually the whole span fits the staging buffer
            if (soff[n] - soff[0] > (uint32_t)CAP) {            // uniform: everybody reads the same two words
                if (tid == 0) {
                    // as many whole sub-bins as fit (at least one)
                    int lo = 1, hi = n;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (soff[mid] - soff[0] <= (uint32_t)CAP) lo = mid; else hi = mid - 1;
                    }
                    s_take = lo;
                }
                __syncthreads();
                take = s_take;
            }
            const uint32_t base = soff[0];
            const uint32_t cnt_e = soff[take] - base;
            if (cnt_e <= (uint32_t)CAP) {
                for (uint32_t e = tid; e < cnt_e; e += kThreads) {
                    const Entry<CT> t = ent[base + e];
                    buf[e] = t;
                    ids[e] = t.id;
                }
                for (int i = tid; i < take; i += kThreads) {
                    const uint32_t lo = soff[i] - base, hi = soff[i + 1] - base;
                    if (hi - lo > (uint32_t)kRankMax) {
                        const uint32_t k = atomicAdd(p.counts, 1u);
                        if (k < (uint32_t)p.big_cap) {
                            p.big_bins[2 * k] = (uint32_t)nm;
                            p.big_bins[2 * k + 1] = (uint32_t)(b + i);
                        }
                    }
                    for (uint32_t e = lo; e < hi; ++e) owner[e] = (uint16_t)i;
                }
                __syncthreads();
                for (uint32_t e = tid; e < cnt_e; e += kThreads) {
                    const int i = owner[e];
                    const uint32_t lo = soff[i] - base, hi = soff[i + 1] - base;
                    const uint32_t c = hi - lo;
                    if (c >= 2 && c <= (uint32_t)kRankMax) {
                        const uint32_t key = ids[e];
                        uint32_t r = 0;
                        for (uint32_t j = lo; j < hi; ++j) r += ids[j] < key;
                        if (lo + r != e) ent[base + lo + r] = buf[e];
                    }
                }
            } else if (tid == 0) {          // a single sub-bin larger than the buffer
                const uint32_t k = atomicAdd(p.counts, 1u);
                if (k < (uint32_t)p.big_cap) {
                    p.big_bins[2 * k] = (uint32_t)nm;
                    p.big_bins[2 * k + 1] = (uint32_t)b;
                }
            }
            __syncthreads();
            b += take;
        }
    }
}

// Sub-bins with more than kRankMax entries: one CTA per sub-bin, bitonic network over a
// shared-memory copy (or in place in global memory when the bin does not fit).
template <typename CT>
__global__ void __launch_bounds__(kThreads) msda_bin_sort_big_kernel(const Params p) {
    constexpr int CAP = 32768 / (int)sizeof(Entry<CT>);
    __shared__ Entry<CT> buf[CAP];
    Entry<CT>* __restrict__ entries = static_cast<Entry<CT>*>(p.entries);
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const uint32_t nbig = min(p.counts[0], (uint32_t)p.big_cap);
    for (uint32_t i = blockIdx.x; i < nbig; i += gridDim.x) {
        const size_t nm = p.big_bins[2 * i];
        const uint32_t bin = p.big_bins[2 * i + 1];
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
//
// Tile kernel.  Regrouping the scatter by BIN instead of by destination pixel means every
// sample's grad_output row is read once, not four times: for bin b
//      G_k[b] = sum_{e in b} w_k(e) a(e) grad_output[q(e)]        k = 1..4  (the four corners)
// and pixel (y, x) receives  G_1[(y+1,x+1)] + G_2[(y+1,x)] + G_3[(y,x+1)] + G_4[(y,x)].
// A CTA owns a TH x 8 pixel tile of one level, frame and head.  A group of G lanes walks one
// bin row of the tile left to right, keeps G_2/G_4 of the previous bin in registers and emits
//      T[y][x] = G_1[cur] + G_2[prev]  (bin row y+1)      B[y][x] = G_3[cur] + G_4[prev]  (bin row y)
// into two shared-memory tiles with plain stores -- every element is written exactly once, so
// there is nothing to zero and nothing to race on.  After one barrier the CTA stores T + B
// with vector stores.
// The entries of a bin (all its sub-bins) are contiguous and sorted; bin boundaries are read
// one bin ahead and entries one batch ahead (G at a time, one per lane; the owner lane forms
// the entry's four weight products once for the group), so that only the grad_output row
// loads of a batch are exposed.  All groups of a warp run the same bin loop.
// In dense levels (many entries per bin) the 32/G groups of a warp share each bin -- each takes
// a contiguous share of its entries -- and combine their sums with shuffles in a fixed order.
constexpr int kGThreads = 128;
constexpr int kGTileW = 8;

template <int VEC>
struct BinAcc {
    float g1[VEC], g2[VEC], g3[VEC], g4[VEC];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < VEC; ++i) g1[i] = g2[i] = g3[i] = g4[i] = 0.f;
    }
};

__device__ __forceinline__ Entry<float> load_entry(const Entry<float>* __restrict__ ent, uint32_t base, uint32_t pos, uint32_t end) {
    Entry<float> e;
    e.id = 0; e.lh = e.lw = e.a = 0.f;
    if (pos < end) {
        const uint4 t = ld_stream_b128(ent + (size_t)(base + pos));
        e.id = t.x; e.lh = __uint_as_float(t.y); e.lw = __uint_as_float(t.z); e.a = __uint_as_float(t.w);
    }
    return e;
}

#ifndef MSDA_WALK_MIN_BLOCKS
#define MSDA_WALK_MIN_BLOCKS 6
#endif
#ifndef MSDA_WALK_MB_BF16
#define MSDA_WALK_MB_BF16 7
#endif
#ifndef MSDA_WALK_MB_F32
#define MSDA_WALK_MB_F32 6
#endif
#ifndef MSDA_WALK_G4_MIN_BLOCKS
#define MSDA_WALK_G4_MIN_BLOCKS 5
#endif

// CTAs per SM, measured on B200 at the A2D shape (D = 32) with the lockstep loops: bf16 rows 6 / 7 / 8 -> 267 / 247 / 259 us,
// fp32 rows 4 / 5 / 6 -> 388 / 314 / 288 us (registers vs. rows in flight)
template <typename T, int VEC, int G>
constexpr int walk_min_blocks() {
    return (G == 8 && VEC == 4) ? (sizeof(T) == 2 ? MSDA_WALK_MB_BF16 : MSDA_WALK_MB_F32) : MSDA_WALK_MIN_BLOCKS;
}

// TWT / BRT: tile width in pixels and bin rows per tile (0: one bin row per 128-bit fp32 lane group, 512 / D).  The
// 4-lane variant for 64-byte bf16 rows (VEC = 8, G = 4) has 32 lane groups per CTA: 32 bin rows x 4 pixels.
template <typename T, int VEC, int G, int TWT = kGTileW, int BRT = 0>
__global__ void __launch_bounds__(kGThreads, (BRT ? MSDA_WALK_G4_MIN_BLOCKS : walk_min_blocks<T, VEC, G>())) msda_grad_value_walk_kernel(const Params p, const int thd, const int twd) {
    constexpr int D = VEC * G;
    constexpr int NGRP = kGThreads / G;          // groups per CTA
    constexpr int GW = 32 / G;                   // groups per warp
    constexpr int BR = BRT ? BRT : ((512 / D) < 2 ? 2 : (512 / D));   // bin rows per tile
    constexpr int TH = BR - 1, TW = TWT;
#ifndef MSDA_WALK_STEP
#define MSDA_WALK_STEP 4
#endif
    constexpr int STEP = VEC <= 4 ? MSDA_WALK_STEP : 4;       // grad_output rows in flight per lane
    constexpr int NWARP = kGThreads / 32;
    // dense levels (many entries per bin): a warp walks a bin row and its groups share every bin; their
    // tiles are a little smaller (12 x 8; measured 3 x 4 .. 12 x 8: the halo of re-walked bins costs more
    // than the coarser load balance), the other levels use the full TH x TW tile
#ifndef MSDA_WALK_THD
#define MSDA_WALK_THD 12
#endif
#ifndef MSDA_WALK_TWD
#define MSDA_WALK_TWD 8
#endif
    // thd x twd: the dense-level tile, chosen per launch (api: walk_dense_tile) -- it changes how the bins are
    // spread over CTAs, never the order in which a bin's entries or a pixel's four bins are summed
    const int TH_D = thd < TH ? thd : TH, TW_D = twd < TW ? twd : TW;

    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    __shared__ int tstart[kMaxLevels + 1];   // first tile of each level in the launch-wide order
    __shared__ int s_tile;
    __shared__ __align__(16) float sTB[2 * TH * TW * D];      // T tile, then B tile: one base, constant distance
    float* const sT = sTB;
    float* const sB = sTB + TH * TW * D;
    load_levels(p, lv, &s_sb, &s_sq);
    auto is_dense = [&](const int l) { return GW > 1 && lv[l].nch_log2 >= 3; };
    auto tiles_of = [&](const int l) {
        const int th = is_dense(l) ? TH_D : TH, tw = is_dense(l) ? TW_D : TW;
        return ((lv[l].H + th - 1) / th) * ((lv[l].W + tw - 1) / tw);
    };
    if (threadIdx.x == 0) {
        // launch-wide order: coarsest level first (its tiles carry the longest lists), all frames and
        // heads of a level before the next level
        int t = 0;
        for (int l = p.L - 1; l >= 0; --l) {
            tstart[l] = t;
            t += p.N * p.M * tiles_of(l);
        }
        tstart[p.L] = index_usable(p, s_sb) ? t : 0;     // total (no tiles at all when the index is unusable)
    }
    __syncthreads();

    const T* __restrict__ gout = static_cast<const T*>(p.grad_out);
    T* __restrict__ gval = static_cast<T*>(p.grad_value);
    const Entry<float>* __restrict__ entries = static_cast<const Entry<float>*>(p.entries);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = tid / G, gl = tid % G, gw = lane / G;    // gw: group index inside the warp
    const int total_tiles = tstart[p.L];
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const size_t qstride = (size_t)p.M * p.D;
    const uint32_t rowbytes = (uint32_t)(qstride * sizeof(T));     // < 2^31 (tile path: msda_api.cu)

    while (true) {
        // dynamic tile scheduler: heavy tiles first, no CTA is handed two of them while others idle
        if (tid == 0) s_tile = (int)atomicAdd(p.counts + 1, 1u);
        __syncthreads();
        const int t = s_tile;
        if (t >= total_tiles) break;
        int l = p.L - 1;
        while (l > 0 && t >= tstart[l - 1]) --l;
        const Level L_ = lv[l];
        const bool dense = is_dense(l);                    // the groups of a warp share each bin
        const int th_l = dense ? TH_D : TH, tw_l = dense ? TW_D : TW;
        const int tiles_x = (L_.W + tw_l - 1) / tw_l;
        const int tiles_l = tiles_x * ((L_.H + th_l - 1) / th_l);
        const int rem = t - tstart[l];
        const int n = rem / (tiles_l * p.M);
        const int r2 = rem - n * tiles_l * p.M;
        const int k = r2 / p.M, m = r2 - k * p.M;
        const int y0 = (k / tiles_x) * th_l, x0 = (k % tiles_x) * tw_l;
        const int nb = min(x0 + tw_l, L_.W) - x0 + 1;      // bins walked per row: x0 .. min(x0+tw, W)

        const size_t nm = (size_t)n * p.M + m;
        const uint32_t* off = p.bin_off + nm * (p.sb_max + 1);
        // entries are addressed from the launch-wide base with a 32-bit position (N*M*Lq*L*P < 2^32, msda_api.cu): one
        // register instead of a pointer ptxas would rather recompute than keep
        const Entry<float>* const ent = entries;
        const uint32_t ebase = (uint32_t)(nm * per_nm);
        // rows of grad_output by byte offset: query * rowbytes is a 32 x 32 -> 64 bit multiply-add onto the base
        const char* gbase = reinterpret_cast<const char*>(gout + ((size_t)n * p.Lq * p.M + m) * p.D + gl * VEC);

        // The lane groups of a warp run every loop below in lockstep, whatever their own rows and bins hold: trip
        // counts are warp-wide (votes), a group with nothing left works on empty batches (zero weights, no loads).
        // A warp that lets its groups leave a loop at different times pays for it at every shuffle afterwards
        // (ptxas routes shuffles of a diverged warp through a WARPSYNC.COLLECTIVE slow path, ~9 instructions each).
        constexpr uint32_t kFull = 0xffffffffu;
        const int row_first = dense ? warp : grp;
        const int row_step = dense ? NWARP : NGRP;
        for (int rbase = 0; rbase <= th_l; rbase += row_step) {
            const int row = rbase + row_first;
            const int by = y0 + row;
            const bool live = row <= th_l && by <= L_.H;               // this group has a bin row to walk
            if (GW > 1 ? !__any_sync(kFull, live) : !live) continue;   // no group of the warp has one
            const bool emit_t = live && row >= 1;                      // pixel row by-1 is in the tile
            const bool emit_b = live && row <= th_l - 1 && by <= L_.H - 1;     // pixel row by is in the tile
            // emit pointer: T[row-1][b-1] of this lane, advanced by one pixel per bin; B[row][b-1] sits at a constant
            // distance from it (kept as ONE running register: ptxas otherwise rebuilds both addresses from the thread
            // index at every bin)
            constexpr int kTtoB = TH * TW * D + TW * D;
            float* tcur = sT + (row - 1) * TW * D + gl * VEC - D;

            // Entry offsets at the bin boundaries of the row segment, read one bin ahead.
            const uint32_t* orow = off + L_.bin_start + (((live ? by : y0) * (L_.W + 1) + x0) << L_.nch_log2);
            auto bound = [&](const int j) { return live ? orow[min(j, nb) << L_.nch_log2] : 0u; };
            // the share of bin [lo, hi) this group works on: all of it, or 1/GW of it in dense levels
            auto share = [&](const uint32_t lo, const uint32_t hi, uint32_t& r0, uint32_t& r1) {
                if (!dense) { r0 = lo; r1 = hi; return; }
                const uint32_t len = (hi - lo + GW - 1) / GW;
                r0 = min(hi, lo + gw * len);
                r1 = min(hi, r0 + len);
            };

            BinAcc<VEC> acc;
            acc.clear();
            float p2[VEC], p4[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) p2[i] = p4[i] = 0.f;

            uint32_t b_lo = bound(0), b_hi = bound(1), b_nx = bound(2);
            uint32_t r0, r1, n0, n1;
            share(b_lo, b_hi, r0, r1);
            share(b_hi, b_nx, n0, n1);
            Entry<float> nxt = load_entry(ent, ebase, r0 + gl, r1);

#pragma unroll 1
            for (int b = 0; b < nb; ++b) {
                const uint32_t b_nn = bound(b + 3);                 // boundary needed two bins from now
                uint32_t e0 = r0;
                bool active = true;                                 // batches of this bin left for this group
                while (true) {
                    Entry<float> mine;
                    mine.id = 0; mine.lh = mine.lw = mine.a = 0.f;
                    int nbat = 0;
                    if (active) {
                        mine = nxt;
                        const bool more = e0 + G < r1;
                        nxt = more ? load_entry(ent, ebase, e0 + G + gl, r1) : load_entry(ent, ebase, n0 + gl, n1);
                        nbat = e0 < r1 ? (int)min((uint32_t)G, r1 - e0) : 0;
                        active = more;
                        e0 += G;
                    }
                    const float ah = mine.a * (1.f - mine.lh), al = mine.a * mine.lh, hw = 1.f - mine.lw;
                    const float myw[4] = {ah * hw, ah * mine.lw, al * hw, al * mine.lw};
                    const uint32_t myq = mine.id >> p.id_shift;
#pragma unroll
                    for (int c0 = 0; c0 < G; c0 += STEP) {
                        if (GW > 1 ? !__any_sync(kFull, c0 < nbat) : c0 >= nbat) break;
                        using R = typename Raw<sizeof(T) * VEC>::type;
                        R raw[STEP];
#pragma unroll
                        for (int e = 0; e < STEP; ++e) {
                            const uint32_t q = __shfl_sync(kFull, myq, c0 + e, G);
                            // entries past nbat carry zero weights and zero rows
                            raw[e] = load_raw_if<T, VEC>(c0 + e < nbat, reinterpret_cast<const T*>(gbase + (size_t)q * rowbytes));
                        }
#pragma unroll
                        for (int e = 0; e < STEP; ++e) {
                            const float w0 = __shfl_sync(kFull, myw[0], c0 + e, G);
                            const float w1 = __shfl_sync(kFull, myw[1], c0 + e, G);
                            const float w2 = __shfl_sync(kFull, myw[2], c0 + e, G);
                            const float w3 = __shfl_sync(kFull, myw[3], c0 + e, G);
                            float gv[VEC];
                            unpack_row<T, VEC>(raw[e], gv);
                            constexpr bool PK = use_packed_fma<T, 2>();
#pragma unroll
                            for (int i = 0; i < VEC; i += 2) {
                                axpy2<PK>(acc.g1[i], acc.g1[i + 1], w0, gv[i], gv[i + 1]);
                                axpy2<PK>(acc.g2[i], acc.g2[i + 1], w1, gv[i], gv[i + 1]);
                                axpy2<PK>(acc.g3[i], acc.g3[i + 1], w2, gv[i], gv[i + 1]);
                                axpy2<PK>(acc.g4[i], acc.g4[i + 1], w3, gv[i], gv[i + 1]);
                            }
                        }
                    }
                    if (GW > 1 ? !__any_sync(kFull, active) : !active) break;
                }
                if (dense) {
                    // fixed-order combine over the GW groups of the warp (lane bits >= log2 G)
#pragma unroll
                    for (int d = G; d < 32; d <<= 1) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            acc.g1[i] += __shfl_xor_sync(kFull, acc.g1[i], d);
                            acc.g2[i] += __shfl_xor_sync(kFull, acc.g2[i], d);
                            acc.g3[i] += __shfl_xor_sync(kFull, acc.g3[i], d);
                            acc.g4[i] += __shfl_xor_sync(kFull, acc.g4[i], d);
                        }
                    }
                }
                // finish the bin: the pixel to its left is complete; hand G2/G4 on
                if (b > 0 && (!dense || gw == 0)) {
                    if (emit_t) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) tcur[i] = acc.g1[i] + p2[i];
                    }
                    if (emit_b) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) tcur[kTtoB + i] = acc.g3[i] + p4[i];
                    }
                }
                tcur += D;
#pragma unroll
                for (int i = 0; i < VEC; ++i) { p2[i] = acc.g2[i]; p4[i] = acc.g4[i]; }
                acc.clear();
                b_lo = b_hi; b_hi = b_nx; b_nx = b_nn;
                r0 = n0; r1 = n1;
                share(b_hi, b_nx, n0, n1);
            }
        }
        __syncthreads();
        for (int i = tid; i < th_l * TW * G; i += kGThreads) {
            const int c = i % G, px = (i / G) % TW, py = i / (G * TW);
            const int y = y0 + py, x = x0 + px;
            if (px < tw_l && y < L_.H && x < L_.W) {
                float v[VEC];
                const float* a = sT + (py * TW + px) * D + c * VEC;
                const float* b = sB + (py * TW + px) * D + c * VEC;
                const size_t pixel = (size_t)n * p.S + L_.start + y * L_.W + x;
                const bool padded = p.value_mask != nullptr && p.value_mask[pixel] != 0;    // masked_fill's backward
#pragma unroll
                for (int j = 0; j < VEC; ++j) v[j] = padded ? 0.f : a[j] + b[j];
                store_row<T, VEC>(gval + pixel * qstride + (size_t)m * p.D + c * VEC, v);
            }
        }
        __syncthreads();
    }

    // value rows that belong to no level (level_start_index with gaps) get a zero gradient
    for (size_t row = (size_t)blockIdx.x * kGThreads + tid; row < (size_t)p.N * p.S; row += (size_t)gridDim.x * kGThreads) {
        const int s = (int)(row % p.S);
        bool covered = false;
        for (int k = 0; k < p.L; ++k) covered |= (s >= lv[k].start && s < lv[k].start + lv[k].H * lv[k].W);
        if (!covered) {
            T* dst = gval + row * qstride;
            for (size_t i = 0; i < qstride; ++i) dst[i] = Elem<T>::from_f(0.f);
        }
    }
}

// Direct gather for calls with few queries per frame (decoder cross-attention: Lq = 5 .. 100 over thousands of
// pixels, /root/reference/models/deformable_transformer.py:330-347).  The inverse index costs five launches that
// all scale with the number of BINS; here nothing does.  grad_value is cleared with a memset and one CTA per
// (frame, head, level) handles that level's Lq*P samples entirely in shared memory:
//   1. every sample yields up to four contributions (key, weight): key = pixel | sample | corner, weight =
//      bilinear weight * attention weight (same geometry rules as everywhere, msda_common.cuh: locate);
//   2. a bitonic network sorts them by key -- all contributions to one pixel become adjacent, in ascending
//      (sample, corner) order, which fixes the summation order as a pure function of the inputs;
//   3. a group of G lanes per touched pixel sums  weight * grad_output[query]  over its run and stores the row.
// No atomics on floating-point data, no workspace; bit-identical from run to run.
constexpr int kDirectMax = 2048;    // contributions (4 per sample) of one (frame, head, level) held in shared memory
constexpr int kDirectRows = 4096;   // grad_output elements of one (frame, head) staged in shared memory (fp32)

//
// ALL (the default for these calls; cooperative launch): the whole backward in this one launch.
//   * every CTA first zero-fills its share of grad_value (the reference's zeros_like, ms_deform_attn_cuda.cu:121) with
//     128-bit stores, evenly over the grid; one grid-wide barrier later the touched rows are written over the zeros;
//   * the CTA of (frame, head, level) also forms grad_sampling_loc / grad_attn_weight of the level's Lq*P samples
//     (cuh:116-158: a lane group per sample reads the four corner rows, three partial sums are reduced over its lanes),
// so that a decoder layer's backward is one kernel instead of memset + sample-gradient kernel + gather.
template <typename T, typename TA, int VEC, int G, bool ALL = false>
__global__ void __launch_bounds__(kThreads) msda_grad_value_direct_kernel(const Params p, const int K, const int id_bits) {
    constexpr int NG = kThreads / G;
    constexpr int D = VEC * G;
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    __shared__ unsigned long long s_kw[kDirectMax];       // key << 32 | weight bits: one 64-bit compare-exchange
    __shared__ uint16_t s_head[kDirectMax];               // first contribution of every touched pixel
    __shared__ __align__(16) float s_g[kDirectRows];      // this (frame, head)'s grad_output rows, fp32
    __shared__ int s_nhead;
    load_levels(p, lv, &s_sb, &s_sq);
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const TA* __restrict__ attn = static_cast<const TA*>(p.attn);
    const T* __restrict__ gout = static_cast<const T*>(p.grad_out);
    T* __restrict__ gval = static_cast<T*>(p.grad_value);
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    const int nsamp = p.Lq * p.P;                         // samples of one (frame, head, level)
    const int pshift = id_bits + 2;
    const uint32_t idmask = (1u << id_bits) - 1u;
    const bool staged = p.Lq * D <= kDirectRows;          // else the rows are read from global memory (L2)
    const int items = p.N * p.M * p.L;
    int nm_staged = -1;
    if constexpr (ALL) {
        uint4* gz = reinterpret_cast<uint4*>(gval);        // 16-byte aligned, a whole number of 16-byte words (tile path)
        const size_t n16 = (size_t)p.N * p.S * p.M * p.D * sizeof(T) / 16;
        for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n16; i += (size_t)gridDim.x * kThreads)
            gz[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    bool zero_fill_done = !ALL;     // ALL: the grid-wide barrier comes right before this CTA's first row write, so that its
                                    // first item's sample gradients and sort overlap the other CTAs' zero-fill
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int l = it % p.L;
        const int nm = it / p.L;
        const int m = nm % p.M, n = nm / p.M;
        const Level L_ = lv[l];
        const T* __restrict__ gbase = gout + ((size_t)n * p.Lq * p.M + m) * p.D;
        const size_t qstride = (size_t)p.M * p.D;
        if (threadIdx.x == 0) s_nhead = 0;
        if (staged && nm != nm_staged) {                  // consecutive items are the levels of one (frame, head)
            for (int i = grp; i < p.Lq; i += NG) {
                float g[VEC];
                load_row<T, VEC>(gbase + (size_t)i * qstride + gl * VEC, g);
#pragma unroll
                for (int c = 0; c < VEC; ++c) s_g[i * D + gl * VEC + c] = g[c];
            }
            nm_staged = nm;
        }
        if constexpr (ALL) {
            // sample gradients of the level's Lq * P samples: one lane group per sample
            const T* __restrict__ vb = static_cast<const T*>(p.value) + ((size_t)n * p.S * p.M + m) * p.D + gl * VEC;
            TA* __restrict__ gloc = static_cast<TA*>(p.grad_loc);
            TA* __restrict__ gattn = static_cast<TA*>(p.grad_attn);
            if (staged) __syncthreads();                        // s_g of this (frame, head) is complete
            // (the groups of a warp take the loop together: the reduction below runs with every lane present)
            for (int i0 = 0; i0 < nsamp; i0 += NG) {
                const int i = i0 + grp;
                const bool mine = i < nsamp;
                const int q = mine ? i / p.P : 0, pt = mine ? i - q * p.P : 0;
                const size_t si = (((size_t)n * p.Lq + q) * p.M + m) * p.LP + l * p.P + pt;
                const XY<float> xy = load_xy(loc + 2 * si);
                const float a = (float)ld_stream(attn + si);
                const Sample<float> sm = locate(xy.x, xy.y, L_.H, L_.W);
                float ga = 0.f, gx = 0.f, gy = 0.f;
                if (mine && sm.ok) {                             // uniform over the lane group
                    int pix[4];
                    corner_pixels(sm, L_, pix);
                    if (p.value_mask != nullptr) {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4)
                            if (pix[c4] >= 0 && p.value_mask[(size_t)n * p.S + pix[c4]]) pix[c4] = -1;
                    }
                    float v[4][VEC], g[VEC];
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
#pragma unroll
                        for (int c = 0; c < VEC; ++c) v[c4][c] = 0.f;
                        if (pix[c4] >= 0) load_row<T, VEC>(vb + (size_t)pix[c4] * qstride, v[c4]);
                    }
                    if (staged) {
#pragma unroll
                        for (int c = 0; c < VEC; ++c) g[c] = s_g[q * D + gl * VEC + c];
                    } else {
                        load_row<T, VEC>(gbase + (size_t)q * qstride + gl * VEC, g);
                    }
                    const float hh = 1.f - sm.lh, hw = 1.f - sm.lw;
#pragma unroll
                    for (int c = 0; c < VEC; ++c) {
                        ga = fmaf(g[c], hh * (hw * v[0][c] + sm.lw * v[1][c]) + sm.lh * (hw * v[2][c] + sm.lw * v[3][c]), ga);
                        gx = fmaf(g[c], hh * (v[1][c] - v[0][c]) + sm.lh * (v[3][c] - v[2][c]), gx);
                        gy = fmaf(g[c], hw * (v[2][c] - v[0][c]) + sm.lw * (v[3][c] - v[1][c]), gy);
                    }
                }
#pragma unroll
                for (int d = 1; d < G; d <<= 1) {
                    ga += __shfl_xor_sync(0xffffffffu, ga, d, G);
                    gx += __shfl_xor_sync(0xffffffffu, gx, d, G);
                    gy += __shfl_xor_sync(0xffffffffu, gy, d, G);
                }
                if (mine && gl == 0) {
                    gattn[si] = Elem<TA>::from_f(ga);
                    store_xy(gloc + 2 * si, (float)L_.W * a * gx, (float)L_.H * a * gy);
                }
            }
        }
        // 1. contributions
        for (int i = threadIdx.x; i < K / 4; i += kThreads) {
            uint32_t key[4] = {~0u, ~0u, ~0u, ~0u};
            float w[4] = {0.f, 0.f, 0.f, 0.f};
            if (i < nsamp) {
                const int q = i / p.P, pt = i - q * p.P;
                const size_t si = (((size_t)n * p.Lq + q) * p.M + m) * p.LP + l * p.P + pt;
                const XY<float> xy = load_xy(loc + 2 * si);
                const float a = (float)ld_stream(attn + si);
                const Sample<float> sm = locate(xy.x, xy.y, L_.H, L_.W);
                if (sm.ok) {
                    int pix[4];
                    corner_pixels(sm, L_, pix);           // level_start + h*W + w, or -1 outside the map
                    if (p.value_mask != nullptr) {         // padded pixels receive nothing (their rows stay zero)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (pix[c] >= 0 && p.value_mask[(size_t)n * p.S + pix[c]]) pix[c] = -1;
                    }
                    const float hh = 1.f - sm.lh, hw = 1.f - sm.lw;
                    const float cw[4] = {hh * hw, hh * sm.lw, sm.lh * hw, sm.lh * sm.lw};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (pix[c] >= 0) {
                            key[c] = ((uint32_t)(pix[c] - L_.start) << pshift) | ((uint32_t)i << 2) | (uint32_t)c;
                            w[c] = cw[c] * a;
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                s_kw[4 * i + c] = ((unsigned long long)key[c] << 32) | (unsigned long long)__float_as_uint(w[c]);
        }
        __syncthreads();
        // 2. bitonic sort by key (keys are unique, so the weight bits below them never decide).  One thread per
        //    compare-exchange, the same pairs of the same 64-element segments for a warp in every round at
        //    distance <= 32, so between two such rounds a warp barrier is enough.
        for (int k = 2; k <= K; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int pi = threadIdx.x; pi < K / 2; pi += kThreads) {
                    const int t = 2 * pi - (pi & (j - 1)), u = t + j;
                    const unsigned long long a0 = s_kw[t], a1 = s_kw[u];
                    if ((a0 > a1) == ((t & k) == 0)) {
                        s_kw[t] = a1;
                        s_kw[u] = a0;
                    }
                }
                // the next round exchanges at distance j/2 (or k, when this k is done): a block barrier unless
                // both this round's writes and the next round's reads stay inside the warps' own segments
                const int jn = j > 1 ? (j >> 1) : k;
                if (j > 32 || jn > 32) __syncthreads(); else __syncwarp();
            }
        }
        __syncthreads();
        // 3. runs of equal pixel (which group takes which run does not matter: runs are independent)
        for (int t = threadIdx.x; t < K; t += kThreads) {
            const uint32_t kt = (uint32_t)(s_kw[t] >> 32);
            if (kt != ~0u && (t == 0 || ((uint32_t)(s_kw[t - 1] >> 32) >> pshift) != (kt >> pshift)))
                s_head[atomicAdd(&s_nhead, 1)] = (uint16_t)t;
        }
        __syncthreads();
        if constexpr (ALL) {
            if (!zero_fill_done) {                         // uniform over the CTA; every CTA passes exactly one grid barrier
                cooperative_groups::this_grid().sync();
                zero_fill_done = true;
            }
        }
        const int nhead = s_nhead;
        for (int r = grp; r < nhead; r += NG) {
            int t = s_head[r];
            const uint32_t pix = (uint32_t)(s_kw[t] >> 32) >> pshift;
            float acc[VEC];
#pragma unroll
            for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
            while (t < K) {
                const unsigned long long kw = s_kw[t];
                const uint32_t kt = (uint32_t)(kw >> 32);
                if ((kt >> pshift) != pix) break;
                const float w = __uint_as_float((uint32_t)kw);
                const int q = (int)((kt >> 2) & idmask) / p.P;
                float g[VEC];
                if (staged) {
#pragma unroll
                    for (int c = 0; c < VEC; ++c) g[c] = s_g[q * D + gl * VEC + c];
                } else {
                    load_row<T, VEC>(gbase + (size_t)q * qstride + gl * VEC, g);
                }
#pragma unroll
                for (int c = 0; c < VEC; ++c) acc[c] = fmaf(w, g[c], acc[c]);
                ++t;
            }
            store_row<T, VEC>(gval + (((size_t)n * p.S + L_.start + pix) * p.M + m) * p.D + gl * VEC, acc);
        }
        __syncthreads();      // shared arrays are rewritten by the next item
    }
    if constexpr (ALL) {
        if (!zero_fill_done) cooperative_groups::this_grid().sync();       // a CTA without items
    }
}

// Any-D / any-dtype fallback: one warp per grad_value row, lanes stride the channels.
template <typename T, typename CT>
__global__ void __launch_bounds__(kThreads) msda_grad_value_generic_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    load_levels(p, lv, &s_sb, &s_sq);
    if (!index_usable(p, s_sb)) return;
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
                const int b_hi = (y + 1) * (L_.W + 1) + x;   // (y+1, x): corner 2; +1: corner 1
                const int b_lo = y * (L_.W + 1) + x;         // (y,   x): corner 4; +1: corner 3
                const int bins[4] = {b_hi + 1, b_hi, b_lo + 1, b_lo};
                for (int c = 0; c < 4; ++c) {
                    // all sub-bins of a bin are adjacent and each is sorted by id
                    const uint32_t beg = off[L_.bin_start + (bins[c] << L_.nch_log2)];
                    const uint32_t end = off[L_.bin_start + ((bins[c] + 1) << L_.nch_log2)];
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
