// msda_bwd_bin.cuh -- the bin-major pass of the backward: grad_value (and, fused, grad_sampling_loc /
// grad_attn_weight) from the inverse index, sorted and summed inside one kernel.
//
// Replaces the scalar fp32 atomicAdd scatter of ms_deform_attn_col2im_bilinear
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:116-153) -- and, in this repo, the first-generation
// pair msda_bin_rank_sort_kernel + msda_grad_value_walk_kernel (msda_backward.cuh), which moved the 16-byte index
// entries through HBM three more times and walked them with dependent per-lane-group loads.
//
// The index (msda_tiles.cuh: stage_build<kIndexFill>, or msda_index_fill_kernel below) holds, per (frame, head), the
// entries {query|sample id, lh, lw, a} grouped by SUB-BIN (bin = top-left corner (h_lo+1, w_lo+1) of the sample in the
// (H+1)x(W+1) grid of its level; 2^k sub-bins per bin by the low bits of the query index), in arrival order inside a
// sub-bin.  For bin b
//      G_c[b] = sum_{e in b} w_c(e) a(e) grad_output[q(e)]        c = 0..3  (the four corners)
// and pixel (y, x) of the level receives  G_0[(y+1,x+1)] + G_1[(y+1,x)] + G_2[(y,x+1)] + G_3[(y,x)].
//
// A CTA owns a th x tw pixel tile of one (level, frame, head): four fp32 accumulator arrays [corner][pixel][D] in
// shared memory.  Its WARPS work independently (no block barrier until the tile's bins are done): a warp takes a UNIT
// -- 32/G/SH consecutive bins of one bin row, whose entries are contiguous in the index -- and
//   stage  loads the unit's entries with coalesced 16-byte streaming loads into registers, ids to its own slice of
//          shared memory;
//   rank   every entry counts the smaller ids of its own sub-bin (about six shared-memory reads) and is written to
//          `sub-bin start + rank`: the summation order below is (sub-bin, id) ascending, a pure function of the
//          inputs -- integer atomics decided only where an entry sat BEFORE this step.  Sub-bins with more than
//          kBPresort entries were sorted in place beforehand (msda_bin_presort_kernel) and keep their order;
//   walk   a group of G lanes (G * VEC = D channels) takes one bin -- or 1/SH of one in dense levels, the SH groups
//          combining with shuffles in a fixed order --, reads each entry with one broadcast shared-memory load,
//          gathers the grad_output row (one 128-bit load per lane, STEP rows in flight) and accumulates the four
//          corner sums in registers.  Units that do not fit the warp's staging area (dense bins) are streamed bin by
//          bin in chunks of whole sub-bins, all groups of the warp sharing the bin, sums carried in registers;
//   emit   G_c goes to the tile's c-th array with plain stores: every (pixel, corner) slot has exactly one producer
//          bin and a bin is handled by exactly one warp -- nothing is zeroed, nothing races.
// One barrier later the CTA adds the four arrays in a fixed order and writes the tile's grad_value rows with vector
// stores: every element of grad_value is written exactly once, no zero-fill, no floating-point atomics, bit-identical
// from run to run.
//
// GRADS (the fused backward): the same walk also forms grad_sampling_loc / grad_attn_weight.  All entries of a bin
// share their four corner value rows, so the group loads them ONCE per bin (instead of four row gathers per sample in
// msda_bwd_sample_tile_kernel), forms the corner dot products d_c = <g, v_c> of STEP entries, reduce-scatters them
// over its lanes and finishes
//      grad_attn = sum_c w_c d_c     grad_x = W a (hh (d1-d0) + lh (d3-d2))     grad_y = H a (hw (d2-d0) + lw (d3-d1))
// (cuh:116-158) in the lane that ends up with the entry.  Bins on the border between two tiles are walked by both;
// the tile that holds the bin's own pixel computes the sample gradients.  Rejected samples have no entry: the fill
// pass writes their zero gradients.
#pragma once

#include "msda_common.cuh"

namespace msda {

#ifndef MSDA_BIN_STEP
#define MSDA_BIN_STEP 4
#endif
#ifndef MSDA_BIN_MIN_BLOCKS
#define MSDA_BIN_MIN_BLOCKS 4
#endif
constexpr int kBThreads = 128;
constexpr int kBWarps = kBThreads / 32;
constexpr int kBCap = 256;                 // entries a warp stages per chunk
constexpr int kBEpl = kBCap / 32;          // ... per lane
constexpr int kBSub = 1 << kMaxSubLog2;    // sub-bins per chunk, at most: one whole bin always fits
constexpr int kBPresort = 128;             // sub-bins with more entries are sorted in place beforehand
constexpr int kBRows = 17;                 // bin rows of a tile, at most (th <= 16)
constexpr int kBPixBytes = 8192;           // fp32 bytes of ONE accumulator array: TPX * D * 4 (TPX = 2048 / D pixels)
constexpr int kBWarpBytes = kBCap * 16 + (kBSub + 4) * 4;   // per-warp staging: sorted entries (ids alias them) + sub-bin positions
static_assert(kBPresort <= kBCap, "a sub-bin the rank step takes must fit one chunk");

constexpr size_t bin_smem_bytes() { return 4 * (size_t)kBPixBytes + (size_t)kBWarps * kBWarpBytes; }

struct BinLevel {
    int th, tw;        // tile size in pixels
    int tiles_x;       // tiles per image row
    int tiles;         // tiles per (frame, head)
    int shl;           // log2 of the lane groups that share one bin (dense levels)
    int tstart;        // first tile of the level in the launch-wide order
    int pad[2];
};

// ---- presort ------------------------------------------------------------------------------------------
// Sub-bins with more than `limit` entries are rare (one location sampled by hundreds of queries with equal low
// index bits); ranking them by counting would cost c^2 shared-memory reads.  They are sorted by id in place
// beforehand, one CTA per sub-bin (bitonic network over a shared-memory copy, or in global memory when it does
// not fit); the bin kernel then keeps their order.  The kernel finds them itself by scanning the offset table.
template <typename CT>
__device__ __forceinline__ void sort_sub_bin_cta(Entry<CT>* __restrict__ g, const uint32_t cnt, Entry<CT>* buf, const int cap) {
    const bool in_smem = cnt <= (uint32_t)cap;
    Entry<CT>* a = in_smem ? buf : g;
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
    __syncthreads();
}

template <typename CT>
__global__ void __launch_bounds__(kThreads) msda_bin_presort_kernel(const Params p, const int limit) {
    constexpr int CAP = 32768 / (int)sizeof(Entry<CT>);
    constexpr int PER = 8;                             // sub-bins a thread looks at per step
    constexpr int SPAN = kThreads * PER;               // sub-bins scanned per work item
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    __shared__ Entry<CT> buf[CAP];
    __shared__ uint32_t s_list[SPAN];
    __shared__ int s_n;
    load_levels(p, lv, &s_sb, &s_sq);
    Entry<CT>* __restrict__ entries = static_cast<Entry<CT>*>(p.entries);
    const int SB = s_sb;
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const int spans = (SB + SPAN - 1) / SPAN;
    const int total = p.N * p.M * spans;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int nm = item / spans;
        const int b0 = (item - nm * spans) * SPAN;
        const uint32_t* off = p.bin_off + (size_t)nm * (p.sb_max + 1);
        Entry<CT>* ent = entries + (size_t)nm * per_nm;
        // thread t looks at sub-bins b0 + t + j * kThreads (coalesced)
        uint32_t prev[PER + 1];
        bool any = false;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int b = b0 + j * kThreads + threadIdx.x;
            const bool big = b < SB && (off[b + 1] - off[b]) > (uint32_t)limit;
            prev[j] = big;
            any |= big;
        }
        if (!__syncthreads_or(any)) continue;          // the usual case: nothing to do
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if (prev[j]) s_list[atomicAdd(&s_n, 1)] = (uint32_t)(b0 + j * kThreads + threadIdx.x);
        __syncthreads();
        const int nbig = s_n;
        for (int j = 0; j < nbig; ++j) {
            const uint32_t bb = s_list[j];
            const uint32_t beg = off[bb], cnt = off[bb + 1] - beg;
            sort_sub_bin_cta<CT>(ent + beg, cnt, buf, CAP);
        }
        __syncthreads();
    }
}

// ---- the bin kernel -----------------------------------------------------------------------------------

// Position (in floats) of 4-channel slice `h` of lane gl inside a pixel's D floats.  For VEC == 8 the two slices of
// a lane sit 4*G floats apart, swapped for odd pixels, so that the lane groups of a warp (which write different
// pixels) spread over all 32 banks.
template <int VEC, int G>
__device__ __forceinline__ int slice_pos(const int pix, const int gl, const int h) {
    if constexpr (VEC == 4) {
        return gl * 4;
    } else {
        return ((h ^ (pix & 1)) * (4 * G)) + gl * 4;
    }
}

template <typename T, int VEC, int G>
__global__ void __launch_bounds__(kBThreads, MSDA_BIN_MIN_BLOCKS)
msda_bwd_bin_kernel(const Params p, const int tile_w0, const int share_target) {
    constexpr int D = VEC * G;
    constexpr int NGRP = kBThreads / G;          // lane groups per CTA
    constexpr int GW = 32 / G;                   // lane groups per warp
    constexpr int GWL = (GW == 8) ? 3 : (GW == 4) ? 2 : (GW == 2) ? 1 : 0;
    constexpr int TPX = kBPixBytes / 4 / D;      // pixels per tile, at most
    constexpr int NS4 = VEC / 4;                 // 4-channel slices per lane
    constexpr int STEP = MSDA_BIN_STEP;          // grad_output rows in flight per lane
    static_assert(VEC == 4 || VEC == 8, "a lane holds 4 or 8 channels");
    static_assert(TPX >= 1 && G <= 32, "row too long for the bin kernel");
    using R = typename Raw<sizeof(T) * VEC>::type;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* const A = reinterpret_cast<float*>(smem_raw);                                   // [4][TPX * D]
    unsigned char* const wbase = smem_raw + 4 * kBPixBytes + (threadIdx.x >> 5) * kBWarpBytes;
    uint4* const sent = reinterpret_cast<uint4*>(wbase);                                   // [kBCap] sorted entries
    uint32_t* const ids = reinterpret_cast<uint32_t*>(wbase);                              // alias: staged ids
    uint32_t* const wspos = reinterpret_cast<uint32_t*>(wbase + kBCap * 16);               // [kBSub + 1]

    __shared__ Level lv[kMaxLevels];
    __shared__ BinLevel blv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    __shared__ int s_total_tiles, s_tile, s_unit;

    load_levels(p, lv, &s_sb, &s_sq);
    if (!index_usable(p, s_sb)) return;
    if (threadIdx.x == 0) {
        // launch-wide order: coarsest level first (its tiles carry the longest lists), all frames and heads
        // of a level before the next level
        int t = 0;
        for (int l = p.L - 1; l >= 0; --l) {
            const int H = lv[l].H, W = lv[l].W;
            const int nx = (W + tile_w0 - 1) / tile_w0;
            const int tw = (W + nx - 1) / nx;
            int th0 = TPX / tw;
            th0 = th0 < 1 ? 1 : (th0 > kBRows - 1 ? kBRows - 1 : th0);
            const int ny = (H + th0 - 1) / th0;
            const int th = (H + ny - 1) / ny;
            // expected entries per bin; SH lane groups share a bin once a share would still hold share_target entries
            const long long lam = (long long)p.Lq * p.P / ((long long)(H + 1) * (W + 1));
            int shl = 0;
            while (shl < GWL && lam >= 2LL * share_target * (1 << shl)) ++shl;
            blv[l].th = th; blv[l].tw = tw;
            blv[l].tiles_x = (W + tw - 1) / tw;
            blv[l].tiles = blv[l].tiles_x * ((H + th - 1) / th);
            blv[l].shl = shl;
            blv[l].tstart = t;
            t += p.N * p.M * blv[l].tiles;
        }
        s_total_tiles = t;
    }
    __syncthreads();

    const T* __restrict__ gout = static_cast<const T*>(p.grad_out);
    T* __restrict__ gval = static_cast<T*>(p.grad_value);
    const uint4* __restrict__ entries = static_cast<const uint4*>(p.entries);

    const int tid = threadIdx.x, lane = tid & 31;
    const int grp = tid / G, gl = tid % G, gw = lane / G;       // gw: lane group inside the warp
    const int total_tiles = s_total_tiles;
    const size_t per_nm = (size_t)p.Lq * p.LP;
    const size_t qstride = (size_t)p.M * p.D;
    const size_t rowbytes = qstride * sizeof(T);

    while (true) {
        if (tid == 0) {
            s_tile = (int)atomicAdd(p.counts + 1, 1u);
            s_unit = 0;
        }
        __syncthreads();
        const int t = s_tile;
        if (t >= total_tiles) break;
        int l = p.L - 1;
        while (l > 0 && t >= blv[l - 1].tstart) --l;
        const Level L_ = lv[l];
        const BinLevel BL = blv[l];
        const int rem = t - BL.tstart;
        const int n = rem / (BL.tiles * p.M);
        const int r2 = rem - n * BL.tiles * p.M;
        const int kt = r2 / p.M, m = r2 - kt * p.M;
        const int y0 = (kt / BL.tiles_x) * BL.th, x0 = (kt % BL.tiles_x) * BL.tw;
        const int th_e = min(BL.th, L_.H - y0), tw_e = min(BL.tw, L_.W - x0);
        const int nrows = th_e + 1, nbx = tw_e + 1;            // bin rows y0 .. y0+th_e, bins x0 .. x0+tw_e
        const int k = L_.nch_log2;
        const uint32_t kmask = (1u << k) - 1u;
        const int nbw = GW >> BL.shl;                          // bins of one unit
        const int upr = (nbx + nbw - 1) / nbw;                 // units per bin row
        const int nunits = nrows * upr;

        const size_t nm = (size_t)n * p.M + m;
        const uint32_t* __restrict__ off = p.bin_off + nm * (p.sb_max + 1) + L_.bin_start;
        const uint4* __restrict__ ent = entries + nm * per_nm;
        const char* gb = reinterpret_cast<const char*>(gout) + (((size_t)n * p.Lq * p.M + m) * p.D + gl * VEC) * sizeof(T);

        // ---- units: each warp on its own ---------------------------------------------------------------------
        while (true) {
            int u = 0;
            if (lane == 0) u = atomicAdd(&s_unit, 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= nunits) break;
            const int r = u / upr;
            const int b0 = (u - r * upr) * nbw;                // first bin of the unit (column inside the tile)
            const int nb_u = min(nbw, nbx - b0);
            const uint32_t* __restrict__ ro = off + ((size_t)((y0 + r) * (L_.W + 1) + x0 + b0) << k);
            const int nsub_u = nb_u << k;
            const uint32_t u_beg = ro[0], u_end = ro[nsub_u];  // same words for every lane: one broadcast each
            // all bins at once when they fit the staging area; else bin by bin, every group of the warp on the bin
            const bool multi = (u_end - u_beg) <= (uint32_t)kBCap && nsub_u <= kBSub;
            const int npass = multi ? 1 : nb_u;
            for (int pass = 0; pass < npass; ++pass) {
                const int pnb = multi ? nb_u : 1;              // bins of this pass
                const int shl_p = multi ? BL.shl : GWL;        // log2 of the groups that share one bin
                const int bsel = gw >> shl_p;                  // this group's bin inside the pass
                const int sidx = gw & ((1 << shl_p) - 1);      // its share of it
                const bool active = bsel < pnb;
                const int bcol = b0 + (multi ? bsel : pass);   // bin column inside the tile
                const int s_lo = multi ? 0 : (pass << k), s_hi = s_lo + (pnb << k);   // sub-bins of the pass, relative to ro
                const uint32_t p_end = multi ? u_end : ro[s_hi];

                float acc[4][VEC];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[c][i] = 0.f;

                int cur = s_lo;               // next sub-bin to stage
                uint32_t eoff = 0;            // entries of sub-bin `cur` already taken (slices of an oversized sub-bin)
                while (cur < s_hi) {
                    // ---- chunk: whole sub-bins [cur, cur + nsc) with at most kBCap entries, in even parts ------------
                    const int nld = min(kBSub, s_hi - cur);
                    const uint32_t base = ro[cur];
                    for (int i = lane; i <= nld; i += 32) wspos[i] = ro[cur + i] - base;
                    __syncwarp();
                    const uint32_t remaining = p_end - base - eoff;
                    const uint32_t nch = (remaining + kBCap - 1) / kBCap;
                    uint32_t limit = nch > 1 ? min((uint32_t)kBCap, (remaining + nch - 1) / nch + 16u) : (uint32_t)kBCap;
                    int fit = 0;
                    if (eoff == 0) {
                        for (int j0 = 0; j0 < nld; j0 += 32) {
                            const int idx = j0 + lane + 1;
                            const bool ok = idx <= nld && wspos[idx] <= limit;
                            fit += __popc(__ballot_sync(0xffffffffu, ok));
                        }
                    }
                    const bool slice = eoff > 0 || fit == 0;   // one oversized (presorted) sub-bin, kBCap entries at a time
                    uint32_t ctot, gstart;
                    int nsc;
                    uint32_t cnt_sb = 0;
                    if (slice) {
                        cnt_sb = wspos[1];
                        ctot = min((uint32_t)kBCap, cnt_sb - eoff);
                        gstart = base + eoff;
                        nsc = 1;
                    } else {
                        nsc = fit;
                        ctot = wspos[fit];
                        gstart = base;
                    }
                    // ---- stage: entries -> registers, ids -> shared memory ------------------------------------------
                    uint4 ev[kBEpl];
#pragma unroll
                    for (int i = 0; i < kBEpl; ++i) {
                        const uint32_t e = (uint32_t)(lane + 32 * i);
                        ev[i] = make_uint4(0u, 0u, 0u, 0u);
                        if (e < ctot) {
                            ev[i] = ld_stream_b128(ent + gstart + e);
                            ids[e] = ev[i].x;
                        }
                    }
                    __syncwarp();
                    // ---- rank ----------------------------------------------------------------------------------------
                    uint32_t slot[kBEpl];
#pragma unroll
                    for (int i = 0; i < kBEpl; ++i) {
                        const uint32_t e = (uint32_t)(lane + 32 * i);
                        slot[i] = e;
                        if (32 * i >= (int)ctot) break;        // uniform: the rest of the lanes' entries are empty
                        if (e < ctot && !slice) {
                            // the entry's sub-bin: its bin (the bin boundaries are few: count those at or before e),
                            // then the low bits of its query
                            int b = 0;
                            for (int bb = 1; bb < pnb; ++bb) b += (wspos[bb << k] <= e);
                            const int sr = ((multi ? b : 0) << k) + (int)((ev[i].x >> p.id_shift) & kmask) + (multi ? 0 : s_lo) - cur;
                            const uint32_t lo = wspos[sr], hi = wspos[sr + 1];
                            if (hi - lo <= (uint32_t)kBPresort) {   // larger ones were presorted
                                const uint32_t key = ev[i].x;
                                uint32_t rk = 0;
                                for (uint32_t x = lo; x < hi; ++x) rk += ids[x] < key;
                                slot[i] = lo + rk;
                            }
                        }
                    }
                    __syncwarp();          // ids are dead: the sorted entries take their place
#pragma unroll
                    for (int i = 0; i < kBEpl; ++i) {
                        const uint32_t e = (uint32_t)(lane + 32 * i);
                        if (e < ctot) sent[slot[i]] = ev[i];
                    }
                    __syncwarp();
                    // ---- walk: this group's entries of the chunk -----------------------------------------------------
                    uint32_t s0 = 0, s1 = 0;
                    if (active) {
                        const uint32_t lo_e = multi ? wspos[bsel << k] : 0u;
                        const uint32_t hi_e = multi ? wspos[(bsel + 1) << k] : ctot;
                        const uint32_t len = (hi_e - lo_e + (1u << shl_p) - 1u) >> shl_p;
                        s0 = min(hi_e, lo_e + (uint32_t)sidx * len);
                        s1 = min(hi_e, s0 + len);
                    }
#pragma unroll 1
                    for (uint32_t e = s0; e < s1; e += STEP) {
                        R raw[STEP];
                        uint4 en[STEP];
#pragma unroll
                        for (int i = 0; i < STEP; ++i) {
                            const bool ok = e + i < s1;
                            en[i] = make_uint4(0u, 0u, 0u, 0u);
                            if (ok) en[i] = sent[e + i];
                            raw[i] = load_raw_if<T, VEC>(ok, reinterpret_cast<const T*>(gb + (size_t)(en[i].x >> p.id_shift) * rowbytes));
                        }
#pragma unroll
                        for (int i = 0; i < STEP; ++i) {
                            float gv[VEC];
                            unpack_row<T, VEC>(raw[i], gv);
                            const float lh = __uint_as_float(en[i].y), lw = __uint_as_float(en[i].z), a = __uint_as_float(en[i].w);
                            const float ah = a * (1.f - lh), al = a * lh, hw = 1.f - lw;
                            const float w0 = ah * hw, w1 = ah * lw, w2 = al * hw, w3 = al * lw;
                            constexpr bool PK = use_packed_fma<T, 2>();
#pragma unroll
                            for (int c = 0; c < VEC; c += 2) {
                                axpy2<PK>(acc[0][c], acc[0][c + 1], w0, gv[c], gv[c + 1]);
                                axpy2<PK>(acc[1][c], acc[1][c + 1], w1, gv[c], gv[c + 1]);
                                axpy2<PK>(acc[2][c], acc[2][c + 1], w2, gv[c], gv[c + 1]);
                                axpy2<PK>(acc[3][c], acc[3][c + 1], w3, gv[c], gv[c + 1]);
                            }
                        }
                    }
                    __syncwarp();          // the staging area is rewritten by the next chunk
                    if (slice) {
                        eoff += ctot;
                        if (eoff >= cnt_sb) { eoff = 0; ++cur; }
                    } else {
                        cur += nsc;
                    }
                }
                // fixed-order combine over the lane groups that share the bin (lane bits log2 G .. log2 G + shl_p - 1)
                for (int d = G; d < (G << shl_p); d <<= 1) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < VEC; ++i) acc[c][i] += __shfl_xor_sync(0xffffffffu, acc[c][i], d);
                }
                // emit: corner c of bin (r, bcol) belongs to pixel (r - 1 + (c >> 1), bcol - 1 + (c & 1))
                if (active && sidx == 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int pr = r - 1 + (c >> 1), pc = bcol - 1 + (c & 1);
                        if (pr >= 0 && pr < th_e && pc >= 0 && pc < tw_e) {
                            const int pix = pr * tw_e + pc;
                            float* dst = A + c * (kBPixBytes / 4) + pix * D;
#pragma unroll
                            for (int h = 0; h < NS4; ++h)
                                *reinterpret_cast<float4*>(dst + slice_pos<VEC, G>(pix, gl, h)) =
                                    make_float4(acc[c][4 * h], acc[c][4 * h + 1], acc[c][4 * h + 2], acc[c][4 * h + 3]);
                        }
                    }
                }
            }
        }
        __syncthreads();          // every bin of the tile has been emitted

        // ---- the tile's rows: sum of the four corner arrays, in a fixed order ---------------------------------
        const int npix = th_e * tw_e;
        for (int px = grp; px < npix; px += NGRP) {
            const int py = px / tw_e, pxx = px - py * tw_e;
            float v[VEC];
#pragma unroll
            for (int h = 0; h < NS4; ++h) {
                const int o = px * D + slice_pos<VEC, G>(px, gl, h);
                const float4 a0 = *reinterpret_cast<const float4*>(A + 0 * (kBPixBytes / 4) + o);
                const float4 a1 = *reinterpret_cast<const float4*>(A + 1 * (kBPixBytes / 4) + o);
                const float4 a2 = *reinterpret_cast<const float4*>(A + 2 * (kBPixBytes / 4) + o);
                const float4 a3 = *reinterpret_cast<const float4*>(A + 3 * (kBPixBytes / 4) + o);
                v[4 * h + 0] = (a0.x + a1.x) + (a2.x + a3.x);
                v[4 * h + 1] = (a0.y + a1.y) + (a2.y + a3.y);
                v[4 * h + 2] = (a0.z + a1.z) + (a2.z + a3.z);
                v[4 * h + 3] = (a0.w + a1.w) + (a2.w + a3.w);
            }
            store_row<T, VEC>(gval + ((size_t)n * p.S + L_.start + (size_t)(y0 + py) * L_.W + (x0 + pxx)) * qstride +
                                  (size_t)m * p.D + gl * VEC, v);
        }
        // (the barrier at the top of the loop separates these reads from the next tile's writes)
    }

    // value rows that belong to no level (level_start_index with gaps) get a zero gradient
    for (size_t row = (size_t)blockIdx.x * kBThreads + tid; row < (size_t)p.N * p.S; row += (size_t)gridDim.x * kBThreads) {
        const int s = (int)(row % p.S);
        bool covered = false;
        for (int kk = 0; kk < p.L; ++kk) covered |= (s >= lv[kk].start && s < lv[kk].start + lv[kk].H * lv[kk].W);
        if (!covered) {
            T* dst = gval + row * qstride;
            for (size_t i = 0; i < qstride; ++i) dst[i] = Elem<T>::from_f(0.f);
        }
    }
}

}  // namespace msda
