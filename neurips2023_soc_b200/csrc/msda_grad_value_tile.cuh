// msda_grad_value_tile.cuh -- part B of the backward, second generation: grad_value from the inverse index,
// sorted and summed inside one kernel.
//
// Replaces the scalar fp32 atomicAdd scatter of ms_deform_attn_col2im_bilinear
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:116-153) -- and, in this repo, the pair
// msda_bin_rank_sort_kernel + msda_grad_value_walk_kernel (msda_backward.cuh), which moved the 16-byte index
// entries through HBM three more times (sort: read + write, walk: read) and walked them with dependent,
// per-lane-group loads.
//
// The index is what part A leaves behind (msda_tiles.cuh: stage_build<kIndexFill>): per (frame, head) the
// entries {query|sample id, lh, lw, a} grouped by SUB-BIN (bin = top-left corner (h_lo+1, w_lo+1) of the sample in
// the (H+1)x(W+1) grid of its level; 2^k sub-bins per bin by the low bits of the query index), in arrival order
// inside a sub-bin.  For bin b
//      G_k[b] = sum_{e in b} w_k(e) a(e) grad_output[q(e)]        k = 0..3  (the four corners)
// and pixel (y, x) of the level receives  G_0[(y+1,x+1)] + G_1[(y+1,x)] + G_2[(y,x+1)] + G_3[(y,x)].
//
// A CTA owns a th x tw pixel tile of one (level, frame, head) and works through its (th+1) x (tw+1) bins in
// ROUNDS of at most kTCap entries:
//   stage  the round's entries arrive with coalesced 16-byte streaming loads (whole bin rows of the tile are
//          contiguous in the index), ids go to shared memory;
//   rank   every entry counts the smaller ids of its own sub-bin (about six shared-memory reads) -- its rank --
//          and writes {w_0..w_3} and its query to position `sub-bin start + rank` of the sorted arrays.  The
//          summation order below is therefore a pure function of the inputs: (sub-bin, id) ascending.  Integer
//          atomics decided only where an entry sat BEFORE this step.  Sub-bins larger than kRankMax were sorted
//          in place beforehand (msda_bin_presort_kernel) and keep their order;
//   walk   a group of G lanes (G * VEC = D channels) takes one bin -- or, in dense levels, 1/SH of one, SH lane
//          groups of a warp combining with shuffles in a fixed order -- reads each entry's weights and query with
//          broadcast shared-memory loads, gathers the grad_output row (one 128-bit load per lane, several rows in
//          flight) and accumulates the four corner sums in registers;
//   emit   the group stores G_k into the pixel tile's k-th accumulator array in shared memory: each (pixel, corner)
//          slot has exactly one producer bin, so the stores are plain, nothing is zeroed and nothing races.  (Tiles
//          with a bin row too long for one round -- dense levels -- zero the arrays first and add instead; rounds
//          are separated by barriers and their cuts depend on the bin populations only.)
// After the last round the CTA adds the four arrays in a fixed order and writes the tile's grad_value rows with
// vector stores: every element of grad_value is written exactly once, no zero-fill, no floating-point atomics,
// bit-identical from run to run.
#pragma once

#include "msda_common.cuh"

namespace msda {

#ifndef MSDA_TILE_THREADS
#define MSDA_TILE_THREADS 128
#endif
#ifndef MSDA_TILE_CAP
#define MSDA_TILE_CAP 1024
#endif
#ifndef MSDA_TILE_STEP
#define MSDA_TILE_STEP 4
#endif
#ifndef MSDA_TILE_MIN_BLOCKS
#define MSDA_TILE_MIN_BLOCKS 4
#endif
constexpr int kTThreads = MSDA_TILE_THREADS;
constexpr int kTCap = MSDA_TILE_CAP;        // entries staged per round
constexpr int kTSub = 448;                  // sub-bins per round, at most
constexpr int kTRows = 17;                  // bin rows of a tile, at most (th <= 16)
constexpr int kTEpt = (kTCap + kTThreads - 1) / kTThreads;   // entries a thread stages per round
constexpr int kTPixBytes = 8192;            // fp32 bytes of ONE accumulator array: TPX * D * 4 (TPX = 2048 / D pixels)
static_assert(kTCap >= kRankMax, "a sub-bin the rank step takes must fit one round");
static_assert(kTCap * 16 >= kTCap * 4 + kTCap * 2, "ids + owner alias the sorted weight array");

constexpr size_t tile_smem_bytes() {
    return 4 * (size_t)kTPixBytes                 // A[4][TPX][D]
           + (size_t)kTCap * 16                   // sw (aliased by ids / owner while ranking)
           + (size_t)kTCap * 4                    // sq
           + (size_t)(kTSub + 2) * 4;             // spos
}

struct TileLevel {
    int th, tw;        // tile size in pixels
    int tiles_x;       // tiles per image row
    int tiles;         // tiles per (frame, head)
    int sh_log2;       // lane groups that share one bin (dense levels)
    int tstart;        // first tile of the level in the launch-wide order
    int pad[2];
};

// ---- presort ------------------------------------------------------------------------------------------
// Sub-bins with more than kRankMax entries are rare (one location sampled by hundreds of queries with equal low
// index bits); ranking them by counting would cost c^2 shared-memory reads.  They are sorted by id in place
// beforehand, one CTA per sub-bin (bitonic network over a shared-memory copy, or in global memory when it does
// not fit); the tile kernel then keeps their order.  The kernel finds them itself by scanning the offset table.
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
__global__ void __launch_bounds__(kThreads) msda_bin_presort_kernel(const Params p) {
    constexpr int CAP = 32768 / (int)sizeof(Entry<CT>);
    constexpr int SPAN = 2048;                         // sub-bins scanned per work item
    __shared__ Level lv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    __shared__ Entry<CT> buf[CAP];
    __shared__ uint32_t s_list[kThreads];
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
        for (int i0 = 0; i0 < SPAN; i0 += kThreads) {
            const int b = b0 + i0 + threadIdx.x;
            const bool big = b < SB && (off[b + 1] - off[b]) > (uint32_t)kRankMax;
            if (!__syncthreads_or(big)) continue;      // the usual case: nothing to do
            if (threadIdx.x == 0) s_n = 0;
            __syncthreads();
            if (big) s_list[atomicAdd(&s_n, 1)] = (uint32_t)b;
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
}

// ---- tile kernel --------------------------------------------------------------------------------------

// Position (in floats) of channel slice `h` (VEC/4 slices of 4 channels per lane) of lane gl inside a pixel's
// D floats.  For VEC == 8 the two slices of a lane sit 4*G floats apart, swapped for odd pixels, so that the
// lane groups of a warp (which write different pixels) spread over all 32 banks.
template <int VEC, int G>
__device__ __forceinline__ int slice_pos(const int pix, const int gl, const int h) {
    if constexpr (VEC == 4) {
        return gl * 4;
    } else {
        return ((h ^ (pix & 1)) * (4 * G)) + gl * 4;
    }
}

template <typename T, int VEC, int G>
__global__ void __launch_bounds__(kTThreads, MSDA_TILE_MIN_BLOCKS)
msda_grad_value_tile_kernel(const Params p, const int tile_w0, const int share_target) {
    constexpr int D = VEC * G;
    constexpr int NGRP = kTThreads / G;          // lane groups per CTA
    constexpr int GW = 32 / G;                   // lane groups per warp
    constexpr int NWARP = kTThreads / 32;
    constexpr int TPX = kTPixBytes / 4 / D;      // pixels per tile, at most
    constexpr int NS4 = VEC / 4;                 // 4-channel slices per lane
    constexpr int STEP = MSDA_TILE_STEP;         // grad_output rows in flight per lane
    static_assert(VEC == 4 || VEC == 8, "a lane holds 4 or 8 channels");
    static_assert(TPX >= 1 && G <= 32, "row too long for the tile kernel");
    using R = typename Raw<sizeof(T) * VEC>::type;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* const A = reinterpret_cast<float*>(smem_raw);                                   // [4][TPX * D]
    uint4* const sw = reinterpret_cast<uint4*>(smem_raw + 4 * kTPixBytes);                 // [kTCap] sorted weights
    uint32_t* const ids = reinterpret_cast<uint32_t*>(sw);                                 // alias: staged ids
    uint16_t* const owner = reinterpret_cast<uint16_t*>(ids + kTCap);                      // alias: sub-bin of an entry
    uint32_t* const sq = reinterpret_cast<uint32_t*>(sw + kTCap);                          // [kTCap] sorted queries
    uint32_t* const spos = sq + kTCap;                                                     // [kTSub + 2]

    __shared__ Level lv[kMaxLevels];
    __shared__ TileLevel tlv[kMaxLevels];
    __shared__ int s_sb, s_sq;
    __shared__ int s_total_tiles, s_tile, s_unit, s_cut;
    __shared__ uint32_t rowlo[kTRows], rowhi[kTRows], rowpre[kTRows + 1];

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
            th0 = th0 < 1 ? 1 : (th0 > kTRows - 1 ? kTRows - 1 : th0);
            const int ny = (H + th0 - 1) / th0;
            const int th = (H + ny - 1) / ny;
            // expected entries per bin; SH lane groups share a bin once a share would still hold share_target entries
            const long long lam = (long long)p.Lq * p.P / ((long long)(H + 1) * (W + 1));
            int shl = 0;
            while ((1 << shl) < GW && lam >= 2LL * share_target * (1 << shl)) ++shl;
            tlv[l].th = th; tlv[l].tw = tw;
            tlv[l].tiles_x = (W + tw - 1) / tw;
            tlv[l].tiles = tlv[l].tiles_x * ((H + th - 1) / th);
            tlv[l].sh_log2 = shl;
            tlv[l].tstart = t;
            t += p.N * p.M * tlv[l].tiles;
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
        if (tid == 0) s_tile = (int)atomicAdd(p.counts + 1, 1u);
        __syncthreads();
        const int t = s_tile;
        if (t >= total_tiles) break;
        int l = p.L - 1;
        while (l > 0 && t >= tlv[l - 1].tstart) --l;
        const Level L_ = lv[l];
        const TileLevel TL = tlv[l];
        const int rem = t - TL.tstart;
        const int n = rem / (TL.tiles * p.M);
        const int r2 = rem - n * TL.tiles * p.M;
        const int kt = r2 / p.M, m = r2 - kt * p.M;
        const int y0 = (kt / TL.tiles_x) * TL.th, x0 = (kt % TL.tiles_x) * TL.tw;
        const int th_e = min(TL.th, L_.H - y0), tw_e = min(TL.tw, L_.W - x0);
        const int nrows = th_e + 1, nbx = tw_e + 1;            // bin rows y0 .. y0+th_e, bins x0 .. x0+tw_e
        const int k = L_.nch_log2;
        const int nsr = nbx << k;                              // sub-bins of one bin row of the tile
        const int shl = TL.sh_log2;
        const int nbw = GW >> shl;                             // bins a warp takes per step

        const size_t nm = (size_t)n * p.M + m;
        const uint32_t* __restrict__ off = p.bin_off + nm * (p.sb_max + 1) + L_.bin_start;
        auto row_off = [&](const int r) { return off + ((size_t)((y0 + r) * (L_.W + 1) + x0) << k); };
        const uint4* __restrict__ ent = entries + nm * per_nm;
        const char* gb = reinterpret_cast<const char*>(gout) + (((size_t)n * p.Lq * p.M + m) * p.D + gl * VEC) * sizeof(T);

        if (tid < nrows) {
            const uint32_t* ro = row_off(tid);
            rowlo[tid] = ro[0];
            rowhi[tid] = ro[nsr];
        }
        __syncthreads();
        // does every bin row fit one round?  (uniform: everybody reads the same words)
        bool rmw = nsr > kTSub;
        for (int r = 0; r < nrows; ++r) rmw |= (rowhi[r] - rowlo[r]) > (uint32_t)kTCap;
        if (tid <= nrows) {
            uint32_t s = 0;
            for (int r = 0; r < tid; ++r) s += rowhi[r] - rowlo[r];
            rowpre[tid] = s;
        }
        if (rmw) {
            float4* a4 = reinterpret_cast<float4*>(A);
            for (int i = tid; i < kTPixBytes / 4; i += kTThreads) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // 4 arrays x kTPixBytes / 16
        }
        __syncthreads();

        int cr = 0, cc = 0;           // cursor: bin row of the tile, sub-bin inside the row
        uint32_t ceoff = 0;           // entries of sub-bin (cr, cc) already taken (slices of an oversized sub-bin)
        while (cr < nrows) {
            // ---- plan the round (uniform) ----------------------------------------------------------------
            const bool whole_rows = cc == 0 && ceoff == 0 && nsr <= kTSub && (rowhi[cr] - rowlo[cr]) <= (uint32_t)kTCap;
            int nr = 0, ns = 0, jbase;
            uint32_t total = 0, gstart = 0;
            bool slice = false;
            if (tid == 0) s_unit = 0;
            if (whole_rows) {
                while (cr + nr < nrows && (nr + 1) * nsr <= kTSub) {
                    const uint32_t c = rowhi[cr + nr] - rowlo[cr + nr];
                    if (total + c > (uint32_t)kTCap) break;
                    total += c;
                    ++nr;
                }
                ns = nr * nsr;
                jbase = cr * nsr;
                // sub-bin positions inside the round's staging area
                for (int idx = tid; idx <= ns; idx += kTThreads) {
                    uint32_t v = total;
                    if (idx < ns) {
                        const int ri = idx / nsr, i = idx - ri * nsr, r = cr + ri;
                        v = (rowpre[r] - rowpre[cr]) + (row_off(r)[i] - rowlo[r]);
                    }
                    spos[idx] = v;
                }
            } else {
                // part of one bin row: as many sub-bins as fit, whole bins when at least one fits; or a slice of one
                // oversized (presorted) sub-bin
                const uint32_t* ro = row_off(cr) + cc;
                const int n1 = min(kTSub, nsr - cc);
                const uint32_t g0 = ro[0];
                for (int idx = tid; idx <= n1; idx += kTThreads) spos[idx] = ro[idx] - g0;
                __syncthreads();
                for (int idx = tid; idx <= n1; idx += kTThreads) {
                    const bool fits = spos[idx] <= (uint32_t)kTCap;
                    const bool next_fits = idx < n1 && spos[idx + 1] <= (uint32_t)kTCap;
                    if (fits && !next_fits) s_cut = idx;        // spos is monotone, spos[0] = 0: exactly one writer
                }
                __syncthreads();
                int cut = s_cut;
                if (ceoff > 0 || cut == 0) {
                    slice = true;
                    const uint32_t cnt_sb = spos[1];
                    total = min((uint32_t)kTCap, cnt_sb - ceoff);
                    gstart = g0 + ceoff;
                    ns = 1;
                } else {
                    const int whole = (((cc + cut) >> k) << k) - cc;     // cut rounded down to a bin boundary
                    if (whole > 0) cut = whole;
                    ns = cut;
                    total = spos[cut];
                    gstart = g0;
                }
                jbase = cr * nsr + cc;
            }

            // ---- stage: entries -> registers, ids -> shared memory --------------------------------------
            // (the previous round's walk ended with a barrier: sw / sq / ids / owner are free)
            uint4 ev[kTEpt];
            {
                int ri = 0;
#pragma unroll
                for (int i = 0; i < kTEpt; ++i) {
                    const uint32_t e = (uint32_t)(tid + i * kTThreads);
                    ev[i] = make_uint4(0u, 0u, 0u, 0u);
                    if (e < total) {
                        uint32_t g;
                        if (whole_rows) {
                            while (rowpre[cr + ri + 1] - rowpre[cr] <= e) ++ri;
                            g = rowlo[cr + ri] + (e - (rowpre[cr + ri] - rowpre[cr]));
                        } else {
                            g = gstart + e;
                        }
                        ev[i] = ld_stream_b128(ent + g);
                    }
                }
            }
            __syncthreads();          // spos complete (and, in the partial-row branch, everybody has read s_cut / spos[1])
            if (slice) {
                if (tid == 0) { spos[0] = 0u; spos[1] = total; }
            } else {
                for (int j = tid; j < ns; j += kTThreads) {
                    const uint32_t lo = spos[j], hi = spos[j + 1];
                    for (uint32_t e = lo; e < hi; ++e) owner[e] = (uint16_t)j;
                }
            }
#pragma unroll
            for (int i = 0; i < kTEpt; ++i) {
                const uint32_t e = (uint32_t)(tid + i * kTThreads);
                if (e < total) ids[e] = ev[i].x;
            }
            __syncthreads();
            // ---- rank ---------------------------------------------------------------------------------------
            uint32_t slot[kTEpt];
#pragma unroll
            for (int i = 0; i < kTEpt; ++i) {
                const uint32_t e = (uint32_t)(tid + i * kTThreads);
                slot[i] = e;
                if (e < total && !slice) {
                    const int j = owner[e];
                    const uint32_t lo = spos[j], hi = spos[j + 1];
                    if (hi - lo <= (uint32_t)kRankMax) {            // larger ones were presorted
                        const uint32_t key = ev[i].x;
                        uint32_t r = 0;
                        for (uint32_t x = lo; x < hi; ++x) r += ids[x] < key;
                        slot[i] = lo + r;
                    }
                }
            }
            __syncthreads();          // ids / owner are dead: the sorted arrays take their place
#pragma unroll
            for (int i = 0; i < kTEpt; ++i) {
                const uint32_t e = (uint32_t)(tid + i * kTThreads);
                if (e < total) {
                    const float lh = __uint_as_float(ev[i].y), lw = __uint_as_float(ev[i].z), a = __uint_as_float(ev[i].w);
                    const float ah = a * (1.f - lh), al = a * lh, hw = 1.f - lw;
                    sw[slot[i]] = make_uint4(__float_as_uint(ah * hw), __float_as_uint(ah * lw),
                                             __float_as_uint(al * hw), __float_as_uint(al * lw));
                    sq[slot[i]] = ev[i].x >> p.id_shift;
                }
            }
            __syncthreads();

            // ---- walk ---------------------------------------------------------------------------------------
            const int fb0 = jbase >> k;                                  // first bin (flattened: row * nbx + x) of the round
            const int np = ((jbase + ns - 1) >> k) - fb0 + 1;            // bins (pieces of bins) in the round
            while (true) {
                int u = 0;
                if (lane == 0) u = atomicAdd(&s_unit, 1);
                u = __shfl_sync(0xffffffffu, u, 0);
                if (u * nbw >= np) break;
                const int pi = u * nbw + (gw >> shl);                    // this group's bin
                const int sidx = gw & ((1 << shl) - 1);                  // its share of it
                uint32_t s0 = 0, s1 = 0;
                int fb = fb0;
                if (pi < np) {
                    fb = fb0 + pi;
                    const int lo_j = max(jbase, fb << k) - jbase, hi_j = min(jbase + ns, (fb + 1) << k) - jbase;
                    const uint32_t lo_e = spos[lo_j], hi_e = spos[hi_j];
                    const uint32_t len = (hi_e - lo_e + (1u << shl) - 1u) >> shl;
                    s0 = min(hi_e, lo_e + (uint32_t)sidx * len);
                    s1 = min(hi_e, s0 + len);
                }
                float acc[4][VEC];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[c][i] = 0.f;
#pragma unroll 1
                for (uint32_t e = s0; e < s1; e += STEP) {
                    R raw[STEP];
                    uint4 wv[STEP];
#pragma unroll
                    for (int i = 0; i < STEP; ++i) {
                        const bool ok = e + i < s1;
                        uint32_t q = 0u;
                        wv[i] = make_uint4(0u, 0u, 0u, 0u);
                        if (ok) { q = sq[e + i]; wv[i] = sw[e + i]; }
                        raw[i] = load_raw_if<T, VEC>(ok, reinterpret_cast<const T*>(gb + (size_t)q * rowbytes));
                    }
#pragma unroll
                    for (int i = 0; i < STEP; ++i) {
                        float gv[VEC];
                        unpack_row<T, VEC>(raw[i], gv);
                        const float w0 = __uint_as_float(wv[i].x), w1 = __uint_as_float(wv[i].y);
                        const float w2 = __uint_as_float(wv[i].z), w3 = __uint_as_float(wv[i].w);
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
                __syncwarp();
                // fixed-order combine over the lane groups that share the bin (lane bits log2 G .. log2 G + shl - 1)
                for (int d = G; d < (G << shl); d <<= 1) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < VEC; ++i) acc[c][i] += __shfl_xor_sync(0xffffffffu, acc[c][i], d);
                }
                // emit: corner c of bin (row, bx) belongs to pixel (row - 1 + (c >> 1), bx - 1 + (c & 1))
                if (pi < np && sidx == 0) {
                    const int row = fb / nbx, bx = fb - row * nbx;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int pr = row - 1 + (c >> 1), pc = bx - 1 + (c & 1);
                        if (pr >= 0 && pr < th_e && pc >= 0 && pc < tw_e) {
                            const int pix = pr * tw_e + pc;
                            float* dst = A + c * (kTPixBytes / 4) + pix * D;
#pragma unroll
                            for (int h = 0; h < NS4; ++h) {
                                float4* d4 = reinterpret_cast<float4*>(dst + slice_pos<VEC, G>(pix, gl, h));
                                float4 v = make_float4(acc[c][4 * h], acc[c][4 * h + 1], acc[c][4 * h + 2], acc[c][4 * h + 3]);
                                if (rmw) {
                                    const float4 o = *d4;
                                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                                }
                                *d4 = v;
                            }
                        }
                    }
                }
            }
            __syncthreads();          // the round is done: staging buffers and s_unit may be reused

            // ---- advance the cursor (uniform) -----------------------------------------------------------------
            if (whole_rows) {
                cr += nr;
            } else if (slice) {
                const uint32_t cnt_sb = row_off(cr)[cc + 1] - row_off(cr)[cc];
                ceoff += total;
                if (ceoff >= cnt_sb) { ceoff = 0; ++cc; }
                if (cc >= nsr) { cc = 0; ++cr; }
            } else {
                cc += ns;
                if (cc >= nsr) { cc = 0; ++cr; }
            }
        }

        // ---- the tile's rows: sum of the four corner arrays, in a fixed order ---------------------------------
        const int npix = th_e * tw_e;
        for (int px = grp; px < npix; px += NGRP) {
            const int py = px / tw_e, pxx = px - py * tw_e;
            float v[VEC];
#pragma unroll
            for (int h = 0; h < NS4; ++h) {
                const int o = px * D + slice_pos<VEC, G>(px, gl, h);
                const float4 a0 = *reinterpret_cast<const float4*>(A + 0 * (kTPixBytes / 4) + o);
                const float4 a1 = *reinterpret_cast<const float4*>(A + 1 * (kTPixBytes / 4) + o);
                const float4 a2 = *reinterpret_cast<const float4*>(A + 2 * (kTPixBytes / 4) + o);
                const float4 a3 = *reinterpret_cast<const float4*>(A + 3 * (kTPixBytes / 4) + o);
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
    for (size_t row = (size_t)blockIdx.x * kTThreads + tid; row < (size_t)p.N * p.S; row += (size_t)gridDim.x * kTThreads) {
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
