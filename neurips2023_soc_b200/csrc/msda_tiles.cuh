// msda_tiles.cuh -- the software pipeline shared by the forward and the sample-gradient tile
// kernels: persistent CTAs walk (tile, round, chunk) work items; while the CTA gathers value
// rows for the current item, the sampling locations / attention weights of the NEXT item
// are already in flight from HBM (registers), and are turned into 16-byte sample
// descriptors in the other half of a shared-memory double buffer -- one block barrier per
// item, no exposed DRAM latency.
//
// Sample descriptor (one uint4 per (query, sample), read with a single LDS.128 broadcast):
//   .x  bits 0..27  rel = (h_lo+1)*W + (w_lo+1)   pixel index of the top-left corner, shifted by
//                   one row and one column so that it is never negative (cuh:38-45 geometry)
//       bits 28..31 which of the four corners lie inside the map (cuh:56,62,68,74)
//   .y  lh   .z  lw   (fractional parts, fp32 bits)      .w  attention weight (fp32 bits)
// A rejected sample (cuh:288) or an empty slot has no corner bit set and weight 0.
// The gather derives the four row addresses from rel and the level's row pitch and the four
// bilinear weights from lh/lw: more ALU, a quarter of the shared-memory traffic of storing
// them.
#pragma once

#include "msda_common.cuh"

namespace msda {

struct Work {
    int t, r, c0;   // tile, round inside the tile, first sample of the chunk
};

template <int DPT>
struct Staged {     // what a staging thread carries from prefetch to descriptor build
    float x[DPT], y[DPT], a[DPT];
    float rx[DPT], ry[DPT];   // fused prologue only: the reference point of (query, level)
    int q[DPT];     // query index, -1: empty slot
};

template <int G>
struct TileShape {
    static constexpr int NG = kThreads / G;           // (query, head) rows per round
    static constexpr int DPT = NG * kSC / kThreads;   // descriptors per staging thread
    static_assert(NG * kSC % kThreads == 0, "descriptor staging must divide evenly");
};

__device__ __forceinline__ Work next_work(const Work w, const int rounds, const int LP) {
    Work n = w;
    n.c0 += kSC;
    if (n.c0 >= LP) {
        n.c0 = 0;
        if (++n.r >= rounds) {
            n.r = 0;
            n.t += gridDim.x;
        }
    }
    return n;
}

// Issue the global loads of one work item (no use of the results here).  FUSED: `loc` / `attn`
// are the raw sampling offsets / attention logits and the reference point comes along.
template <typename TA, int G, int P, bool FUSED = false>
__device__ __forceinline__ void stage_load(Staged<TileShape<G>::DPT>& st, const Params& p, const TileMap* tm,
                                           const Tile& tl, const Work& w, const TA* __restrict__ loc,
                                           const TA* __restrict__ attn) {
    constexpr int NG = TileShape<G>::NG, DPT = TileShape<G>::DPT;
    const int st_s = threadIdx.x % kSC, st_j0 = threadIdx.x / kSC;
    const int sg = w.c0 + st_s;
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const int j = st_j0 + k * (kThreads / kSC);
        int q = tile_query(p, tm, tl, w.r * NG + j);
        if (sg >= p.LP) q = -1;
        st.q[k] = q;
        st.x[k] = st.y[k] = st.a[k] = 0.f;
        if (q >= 0) {
            const size_t si = (((size_t)tl.n * p.Lq + q) * p.M + tl.m) * p.LP + sg;
            const XY<float> xy = load_xy(loc + 2 * si);
            st.x[k] = xy.x; st.y[k] = xy.y;
            st.a[k] = (float)ld_stream(attn + si);
            if constexpr (FUSED) {
                const float2 r = __ldg(reinterpret_cast<const float2*>(p.ref) +
                                       ((size_t)tl.n * p.Lq + q) * p.L + min(sg / P, p.L - 1));
                st.rx[k] = r.x; st.ry[k] = r.y;
            }
        }
    }
}

// Fused prologue of MSDeformAttn.forward (/root/reference/models/ops/modules/ms_deform_attn.py:99-106):
//   attention_weights = softmax over the L*P logits of (query, head)
//   sampling_location = reference_point[level] + offset / (W_level, H_level)
// The 16 samples of one (query, head) sit in 16 consecutive staging lanes, so the softmax is two
// 16-lane shuffle reductions.  Both results are also written out: the module returns them and the
// backward reads them.  Requires L*P <= 16 (one chunk).
template <int G, int P>
__device__ __forceinline__ void fused_prologue(Staged<TileShape<G>::DPT>& st, const Params& p, const Level* lv,
                                               const Tile& tl, const Work& w) {
    constexpr int DPT = TileShape<G>::DPT;
    const int st_s = threadIdx.x % kSC;
    const int sg = w.c0 + st_s;
    const Level L_ = lv[min(sg / P, p.L - 1)];
    // (every thread of the CTA stages: the warp is converged here, and stays so -- the skip below is warp-wide)
    constexpr uint32_t kFull = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const bool live = st.q[k] >= 0;                  // false for empty rows and for slots past L*P
        if (!__any_sync(kFull, live)) continue;          // both rows of the warp are empty
        float m = live ? st.a[k] : -INFINITY;
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, d, 16));
        const float e = live ? expf(st.a[k] - m) : 0.f;
        float sum = e;
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) sum += __shfl_xor_sync(kFull, sum, d, 16);
        if (live) {
            st.a[k] = e / sum;
            st.x[k] = st.rx[k] + st.x[k] / (float)L_.W;
            st.y[k] = st.ry[k] + st.y[k] / (float)L_.H;
            if (p.loc_out != nullptr) {      // null: the caller keeps only the raw projections (msda_backward_fused_raw)
                const size_t si = (((size_t)tl.n * p.Lq + st.q[k]) * p.M + tl.m) * p.LP + sg;
                reinterpret_cast<float2*>(p.loc_out)[si] = make_float2(st.x[k], st.y[k]);
                p.attn_out[si] = st.a[k];
            }
        }
    }
}

// What the staging threads do for the inverse index of the grad_value gather
// (msda_backward.cuh, part B) while they build descriptors:
//   kIndexNone   nothing (plain forward)
//   kIndexCount  count the accepted samples of every sub-bin (forward that will be followed by a
//                backward: the scan of these counts is handed to it)
//   kIndexFill   take a slot by advancing the sub-bin's offset in the table (which turns the row of
//                start offsets into the row of end offsets = the next sub-bin's start) and write
//                the sample's 16-byte entry there (backward)
// Only integer atomics are involved; they decide where an entry sits before sorting, never a
// floating-point result.
constexpr int kIndexNone = 0, kIndexCount = 1, kIndexFill = 2;

// KEEPW: a rejected sample's descriptor keeps its attention weight in .w (no corner bit set, so nothing is loaded
// for it): the in-kernel softmax backward needs the weight of every sample, accepted or not.
template <int G, int P, int MODE, bool FUSED = false, bool KEEPW = false>
__device__ __forceinline__ void stage_build(Staged<TileShape<G>::DPT>& st, const Params& p, const Level* lv,
                                            const Tile& tl, const Work& w, uint4* __restrict__ desc,
                                            const bool index_ok = true) {
    constexpr int DPT = TileShape<G>::DPT;
    if constexpr (FUSED) fused_prologue<G, P>(st, p, lv, tl, w);
    const int st_s = threadIdx.x % kSC, st_j0 = threadIdx.x / kSC;
    const int sg = w.c0 + st_s;
    const int l = min(sg / P, p.L - 1);
    const Level L_ = lv[l];
    const size_t nm = (size_t)tl.n * p.M + tl.m;
    // null entries: the direct gather follows, no index is kept; !index_ok: the level table does not fit the
    // index buffer (msda_common.cuh: load_levels), nothing may be written
    const bool fill = MODE == kIndexFill && p.entries != nullptr && index_ok;
    // With 4 descriptors per thread (4-lane rows) the slot atomics of a thread are issued together and the entries
    // written in a second pass: waiting for each atomic's return in turn was 23 % of the sample-gradient kernel's
    // stall samples (300 -> 288 us).  With 2 per thread (8-lane rows) the extra live registers cost more (fp32
    // 394 -> 405 us), so those store at once.
    constexpr bool TWO_PASS = MODE == kIndexFill && DPT >= 4;
    uint4 dsc[DPT];
    uint32_t slot[DPT];
    bool live[DPT];
    // pass 1: descriptors, and every accepted sample's integer atomic (count, or take a slot)
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        uint4 d = make_uint4(0u, 0u, 0u, 0u);
        live[k] = false;
        slot[k] = 0u;
        if (st.q[k] >= 0) {
            const Sample<float> s = locate(st.x[k], st.y[k], L_.H, L_.W);
            if constexpr (KEEPW) d.w = __float_as_uint(st.a[k]);
            if (s.ok) {
                const unsigned h0 = s.h_lo >= 0, w0 = s.w_lo >= 0;
                const unsigned h1 = s.h_lo + 1 <= L_.H - 1, w1 = s.w_lo + 1 <= L_.W - 1;
                unsigned flags = (h0 & w0) | ((h0 & w1) << 1) | ((h1 & w0) << 2) | ((h1 & w1) << 3);
                // padded pixels hold zeros (the module's masked_fill, fused); the mask only arrives through the fused
                // entry points, i.e. in the FUSED forward and the chain-rule (KEEPW) backward instances
                if ((FUSED || KEEPW) && p.value_mask != nullptr)
                    flags = mask_corners(flags, p.value_mask + (size_t)tl.n * p.S + L_.start + s.h_lo * L_.W + s.w_lo, L_.W);
                d.x = (flags << 28) | (unsigned)((s.h_lo + 1) * L_.W + (s.w_lo + 1));
                d.y = __float_as_uint(s.lh);
                d.z = __float_as_uint(s.lw);
                d.w = __float_as_uint(st.a[k]);
                if constexpr (MODE != kIndexNone) {
                    // slot b+1 of the table row belongs to sub-bin b (see msda_bin_scan_kernel)
                    const size_t b = nm * (p.sb_max + 1) + sub_bin(L_, s.h_lo, s.w_lo, st.q[k]) + 1;
                    if constexpr (MODE == kIndexCount) {
                        if (index_ok) atomicAdd(p.bin_off + b, 1u);
                    } else if (fill) {
                        slot[k] = atomicAdd(p.bin_off + b, 1u);
                        live[k] = true;
                        if constexpr (!TWO_PASS) {
                            uint4 e;
                            e.x = ((uint32_t)st.q[k] << p.id_shift) | (uint32_t)sg;
                            e.y = d.y; e.z = d.z; e.w = d.w;
                            static_cast<uint4*>(p.entries)[nm * ((size_t)p.Lq * p.LP) + slot[k]] = e;
                        }
                    }
                }
            }
        }
        dsc[k] = d;
        desc[(st_j0 + k * (kThreads / kSC)) * kDescStride + st_s] = d;
    }
    // pass 2: the index entries go to the slots taken above
    if constexpr (TWO_PASS) {
#pragma unroll
        for (int k = 0; k < DPT; ++k) {
            if (live[k]) {
                uint4 e;
                e.x = ((uint32_t)st.q[k] << p.id_shift) | (uint32_t)sg;
                e.y = dsc[k].y; e.z = dsc[k].z; e.w = dsc[k].w;
                static_cast<uint4*>(p.entries)[nm * ((size_t)p.Lq * p.LP) + slot[k]] = e;
            }
        }
    }
}

// Per-level constants of the gather.  Row addresses are SIGNED 32-bit byte offsets from the base
// of the tile's (frame, head) slice (S*M*D*sizeof(T) < 2^31 on this path): the top-left corner of a
// sample may be the virtual pixel (row -1 / col -1) with a negative offset; it is never
// dereferenced, but the other three corners are reached from it with 64-bit pointer adds.
struct LevelPitch {
    int32_t base;    // byte offset of the virtual pixel (row -1, col -1) of the level, plus the lane's slice
    int32_t wrow;    // bytes between vertically adjacent pixels
};

__device__ __forceinline__ LevelPitch level_pitch(const Level& L_, const uint32_t rowb, const uint32_t lane_off) {
    LevelPitch lp;
    lp.base = (L_.start - L_.W - 1) * (int32_t)rowb + (int32_t)lane_off;
    lp.wrow = L_.W * (int32_t)rowb;
    return lp;
}

}  // namespace msda
