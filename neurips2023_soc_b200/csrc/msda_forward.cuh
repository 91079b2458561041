// msda_forward.cuh -- forward sampling kernels.
//
// Replaces ms_deformable_im2col_gpu_kernel
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299), which runs one
// thread per output element, re-reads every location/weight from global memory in all D
// lanes and loads value rows 4 bytes at a time.
//
// Design (tile kernel, the path taken when a value row is 32..256 bytes):
//   * persistent CTAs walk (frame, head, query-tile) tiles; a tile's queries are spatially
//     compact (msda_common.cuh: TileMap) so their corner rows are shared through L1;
//   * per chunk of 16 samples the CTA turns the tile's sampling locations and attention
//     weights (read once, coalesced, prefetched one work item ahead) into 16-byte sample
//     DESCRIPTORS in a shared-memory double buffer (msda_tiles.cuh);
//   * a group of G lanes owns one (query, head) output row; each lane reads its slice of
//     every corner row with one vector read-only load (128-bit; 64-bit for 64-byte bf16
//     rows so that a row is still spread over 8 lanes), so a warp load instruction fetches
//     32/G complete rows and every fetched byte is used;
//   * accumulation over the L*P samples is in registers, the output row leaves with one
//     128-bit store per lane.
#pragma once

#include "msda_tiles.cuh"

namespace msda {

#ifndef MSDA_FWD_MIN_BLOCKS
#define MSDA_FWD_MIN_BLOCKS 4
#endif
#ifndef MSDA_FWD_STALE
#define MSDA_FWD_STALE 0   // 1: corner registers are zeroed once per row instead of once per sample
#endif

// T value dtype, TA location/weight dtype, VEC channels per lane, G = D / VEC lanes per row,
// P points per level (compile time so that the per-level constants hoist out of the loop).
// COUNT: also count the accepted samples per sub-bin for a backward that will follow.
// FUSED: the softmax / sampling-location prologue of MSDeformAttn.forward runs in the staging
// threads (TA is then the dtype of the raw offsets / logits; the results are fp32).
template <typename T, typename TA, int VEC, int G, int P, bool COUNT, bool FUSED = false, int ROWB = 0>
__global__ void __launch_bounds__(kThreads, (VEC == 8 && G == 4) ? 3 : MSDA_FWD_MIN_BLOCKS)   // 4 wide lanes per row: 80 registers, no spill (222 -> 212 us)
msda_fwd_tile_kernel(const Params p, const int rounds) {
    constexpr int MODE = COUNT ? kIndexCount : kIndexNone;
    using TS = TileShape<G>;
    constexpr int NG = TS::NG;
    constexpr int LPC = (P >= kSC) ? 1 : kSC / P;   // levels per 16-sample chunk
    constexpr int PPC = (P >= kSC) ? kSC : P;       // points of one level per chunk
    static_assert(kSC % PPC == 0 && (P % PPC) == 0, "P must divide or be a multiple of 16");

    __shared__ Level lv[kMaxLevels];
    __shared__ TileMap tm;
    __shared__ int s_sb, s_sq;
    __shared__ uint4 desc[2][NG * kDescStride];

    const int tile_q = NG * rounds;
    load_levels(p, lv, &s_sb, &s_sq);
    build_tile_map(p, lv, s_sq, tile_q, &tm);

    const T* __restrict__ value = static_cast<const T*>(p.value);
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const TA* __restrict__ attn = static_cast<const TA*>(p.attn);
    T* __restrict__ out = static_cast<T*>(p.out);

    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    const int row_elems = p.M * p.D;
    const int total_tiles = p.N * p.M * tm.qtiles;

    Work cur{(int)blockIdx.x, 0, 0};
    if (cur.t >= total_tiles) return;
    Tile tl = decode_tile(p, lv, &tm, cur.t, tile_q);
    Staged<TS::DPT> st;
    stage_load<TA, G, P, FUSED>(st, p, &tm, tl, cur, loc, attn);
    stage_build<G, P, MODE, FUSED>(st, p, lv, tl, cur, desc[0], index_usable(p, s_sb));
    __syncthreads();

    int buf = 0;
    float acc[VEC];
    while (true) {
        const Work nxt = next_work(cur, rounds, p.LP);
        const bool has_next = nxt.t < total_tiles;
        Tile ntl = tl;
        if (has_next) {
            if (nxt.t != cur.t) ntl = decode_tile(p, lv, &tm, nxt.t, tile_q);
            stage_load<TA, G, P, FUSED>(st, p, &tm, ntl, nxt, loc, attn);      // HBM loads fly during the gather
        }

        const int q_mine = tile_query(p, &tm, tl, cur.r * NG + grp);
        if (cur.c0 == 0) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        }
        if (q_mine >= 0) {
            // all addresses of a tile are 32-bit byte offsets from one CTA-uniform base
            const char* fb = reinterpret_cast<const char*>(value) +
                             ((size_t)tl.n * p.S * row_elems + tl.m * p.D) * sizeof(T);
            asm volatile("" : "+l"(fb));     // keep the base in registers (ptxas would re-read it per load)
            // bytes between horizontally adjacent pixels; known at compile time for the usual
            // d_model (ROWB != 0), so that the second corner of a row pair is an immediate offset
            const uint32_t rowb = ROWB ? (uint32_t)ROWB : (uint32_t)row_elems * (uint32_t)sizeof(T);
            const uint32_t lane_off = (uint32_t)(gl * VEC * sizeof(T));
            const uint4* drow = desc[buf] + grp * kDescStride;
            const int l0 = cur.c0 / P;
#pragma unroll 1
            for (int lc = 0; lc < LPC; ++lc) {
                const int l = l0 + lc;
                if (l >= p.L) break;
                const LevelPitch lp = level_pitch(lv[l], rowb, lane_off);
                const int sbase = (P >= kSC) ? 0 : lc * PPC;
#pragma unroll
                for (int pp = 0; pp < PPC; ++pp) {
                    const uint4 d = drow[sbase + pp];
                    const char* p0 = fb + (lp.base + (int32_t)((d.x & 0x0fffffffu) * rowb));   // top-left corner row
                    const char* p2 = p0 + lp.wrow;                                   // the row below
                    using R = typename Raw<sizeof(T) * VEC>::type;
                    // predicated loads into zeroed raw registers: a corner outside the map contributes
                    // exactly 0 whatever its weight
                    const R r0 = load_raw_if<T, VEC>(d.x & (1u << 28), reinterpret_cast<const T*>(p0));
                    const R r1 = load_raw_if<T, VEC>(d.x & (2u << 28), reinterpret_cast<const T*>(p0 + rowb));
                    const R r2 = load_raw_if<T, VEC>(d.x & (4u << 28), reinterpret_cast<const T*>(p2));
                    const R r3 = load_raw_if<T, VEC>(d.x & (8u << 28), reinterpret_cast<const T*>(p2 + rowb));
                    const float lh = __uint_as_float(d.y), lw = __uint_as_float(d.z), a = __uint_as_float(d.w);
                    const float ah = a * (1.f - lh), al = a * lh, hw = 1.f - lw;
                    const float w0 = ah * hw, w1 = ah * lw, w2 = al * hw, w3 = al * lw;
                    float v0[VEC], v1[VEC], v2[VEC], v3[VEC];
                    unpack_row<T, VEC>(r0, v0);
                    unpack_row<T, VEC>(r1, v1);
                    unpack_row<T, VEC>(r2, v2);
                    unpack_row<T, VEC>(r3, v3);
                    constexpr bool PK = use_packed_fma<T, 0>();
#pragma unroll
                    for (int i = 0; i < VEC; i += 2) {
                        axpy2<PK>(acc[i], acc[i + 1], w0, v0[i], v0[i + 1]);
                        axpy2<PK>(acc[i], acc[i + 1], w1, v1[i], v1[i + 1]);
                        axpy2<PK>(acc[i], acc[i + 1], w2, v2[i], v2[i + 1]);
                        axpy2<PK>(acc[i], acc[i + 1], w3, v3[i], v3[i + 1]);
                    }
                }
            }
            if (cur.c0 + kSC >= p.LP)
                store_row<T, VEC>(out + (((size_t)tl.n * p.Lq + q_mine) * p.M + tl.m) * p.D + gl * VEC, acc);
        }

        if (has_next) stage_build<G, P, MODE, FUSED>(st, p, lv, ntl, nxt, desc[buf ^ 1], index_usable(p, s_sb));
        __syncthreads();
        if (!has_next) break;
        cur = nxt;
        tl = ntl;
        buf ^= 1;
    }
}

// Any-D, any-dtype fallback (D not a power-of-two multiple of the 128-bit vector, rows
// longer than 256 B, or fp64): one thread per output element, plain loops.  Same rules,
// no tiling; exists so that every channel count the reference's test walks through
// (/root/reference/models/ops/test.py:85) has a path.
template <typename T, typename TA, typename CT>
__global__ void __launch_bounds__(kThreads) msda_fwd_generic_kernel(const Params p) {
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    load_levels(p, lv, &s_sb, &s_sq);
    const T* __restrict__ value = static_cast<const T*>(p.value);
    const TA* __restrict__ loc = static_cast<const TA*>(p.loc);
    const TA* __restrict__ attn = static_cast<const TA*>(p.attn);
    T* __restrict__ out = static_cast<T*>(p.out);
    const size_t total = (size_t)p.N * p.Lq * p.M * p.D;
    const size_t row_elems = (size_t)p.M * p.D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % p.D);
        const size_t qm = i / p.D;               // (n*Lq + q)*M + m
        const int m = (int)(qm % p.M);
        const size_t n = qm / p.M / p.Lq;
        const T* vb = value + n * p.S * row_elems + (size_t)m * p.D + c;
        CT acc = 0;
        for (int l = 0; l < p.L; ++l) {
            const Level L_ = lv[l];
            for (int pt = 0; pt < p.P; ++pt) {
                const size_t si = qm * p.LP + l * p.P + pt;
                const XY<CT> xy = load_xy(loc + 2 * si);
                const Sample<CT> s = locate(xy.x, xy.y, L_.H, L_.W);
                if (!s.ok) continue;
                const CT a = (CT)Elem<TA>::to_f(attn[si]);
                int pix[4];
                corner_pixels(s, L_, pix);
                const CT hh = (CT)1 - s.lh, hw = (CT)1 - s.lw;
                const CT w[4] = {hh * hw, hh * s.lw, s.lh * hw, s.lh * s.lw};
                CT val = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (pix[k] >= 0) val += w[k] * (CT)Elem<T>::to_f(vb[(size_t)pix[k] * row_elems]);
                acc += val * a;
            }
        }
        out[i] = Elem<T>::from_f(acc);
    }
}

}  // namespace msda
