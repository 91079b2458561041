// msda_common.cuh -- shared device-side definitions for the sm_100a MSDeformAttn kernels.
//
// Semantics implemented here are those of the reference op
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh):
//   sample position   h_im = loc_y*H - 0.5, w_im = loc_x*W - 0.5            (:285-286)
//   accept            -1 < h_im < H  and  -1 < w_im < W                       (:288)
//   corners           (h_lo,w_lo) (h_lo,w_hi) (h_hi,w_lo) (h_hi,w_hi), each dropped
//                     on its own when outside [0,H-1]x[0,W-1]                 (:56-79)
//   value row         value[n][level_start + h*W + w][m][0..D)                (:47-53,277)
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace msda {

constexpr int kMaxLevels = 16;   // levels held in shared memory per CTA
constexpr int kThreads = 256;    // CTA size of the tile kernels
constexpr int kSC = 16;          // samples staged per chunk (L*P = 16 in every SOC config)
constexpr int kDescStride = kSC + 1;  // 16 B slots per query in the descriptor arrays (+1: bank skew)
#ifndef MSDA_SUB_BIN_TARGET
#define MSDA_SUB_BIN_TARGET 6
#endif
constexpr int kSubBinTarget = MSDA_SUB_BIN_TARGET;   // aimed-at entries per sub-bin (see Level::nch_log2)
constexpr int kMaxSubLog2 = 8;
constexpr int kTileW = 16;       // pyramid query tile: kTileH x kTileW pixels of one level
constexpr int kTileH = 8;
constexpr int kTileQ = kTileW * kTileH;
constexpr int kRankMax = 512;    // largest sub-bin ranked by counting (larger ones are sorted in place beforehand)

// Everything a kernel needs; passed by value.
struct Params {
    const void* value;
    const int64_t* shapes;   // [L][2] (H, W), device
    const int64_t* lsi;      // [L], device
    const void* loc;         // [N][Lq][M][L][P][2]
    const void* attn;        // [N][Lq][M][L][P]
    const void* grad_out;    // [N][Lq][M*D]           (backward)
    // fused prologue (msda_tiles.cuh): loc/attn above then hold the raw sampling offsets / attention logits
    const uint8_t* value_mask;  // [N][S] or null: non-zero = padded pixel, its value row counts as zero and its
                                //   grad_value row is zero (value.masked_fill(mask, 0) of the module, ms_deform_attn.py:96-97)
    const float* ref;        // [N][Lq][L][2] reference points
    float* loc_out;          // [N][Lq][M][L][P][2]  sampling locations computed on the way
    float* attn_out;         // [N][Lq][M][L][P]     softmax of the logits
    void* out;               // [N][Lq][M*D]           (forward)
    void* grad_value;        // [N][S][M][D]
    void* grad_loc;
    void* grad_attn;
    // backward workspace (see msda_backward.cuh)
    uint32_t* bin_off;       // [N*M][sb_max + 1]  row[0] = 0, row[b+1]: count of sub-bin b -> its start offset
                             //                    (scan) -> its end offset (fill); readers use row[b], row[b+1]
    void* entries;           // [N*M][Lq*L*P]      per-bin contribution lists
    uint32_t* counts;        // [0] entries in big_bins (right behind bin_off so one memset clears both)
    uint32_t* big_bins;      // (nm, sub-bin) pairs of sub-bins with > kRankMax entries
    int N, S, M, D, L, Lq, P;
    int LP;                  // L*P
    int id_shift;            // entry id = (q << id_shift) | s,  1<<id_shift >= LP
    int sb_max;              // bound on the number of sub-bins per (frame, head), see ws_layout()
    int big_cap;             // capacity of big_bins in pairs
    unsigned flags;
};

// A "bin" is the top-left corner (h_lo+1, w_lo+1) of a sample in the (H+1)x(W+1) grid of a
// level; every bin is split into 2^nch_log2 sub-bins by the low bits of the query index so
// that a sub-bin holds about kSubBinTarget entries whatever the level's density.
struct Level {
    int H, W;
    int start;      // level_start_index[l]
    int bin_start;  // first sub-bin of the level
    int nch_log2;   // log2(sub-bins per bin)
    int pad[3];
};

// One entry of the inverse index of the grad_value gather (msda_backward.cuh, part B):
// query | sample id, the sample's fractional position and its attention weight.
template <typename CT> struct Entry;
template <> struct __align__(16) Entry<float> {
    uint32_t id;
    float lh, lw, a;
};
template <> struct __align__(16) Entry<double> {
    uint32_t id, pad;
    double lh, lw, a;
};

// ---------------------------------------------------------------------------------------
// element types
template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int kVec = 4;  // elements per 128-bit access
    __device__ static float to_f(float v) { return v; }
    __device__ static float from_f(float v) { return v; }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kVec = 8;
    __device__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Elem<__half> {
    static constexpr int kVec = 8;
    __device__ static float to_f(__half v) { return __half2float(v); }
    __device__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Elem<double> {
    static constexpr int kVec = 2;
    __device__ static double to_f(double v) { return v; }
    __device__ static double from_f(double v) { return v; }
};

// Row fragment of VEC consecutive channels <-> fp32 registers.  One access per lane:
// 128 bits (float x4, 16-bit x8) or 64 bits (16-bit x4); loads take the read-only path.
template <int BYTES> struct Raw;
template <> struct Raw<16> { using type = uint4; };
template <> struct Raw<8> { using type = uint2; };

template <typename T, int VEC>
__device__ __forceinline__ void load_row(const T* p, float (&v)[VEC]) {
    using R = typename Raw<sizeof(T) * VEC>::type;
    const R t = __ldg(reinterpret_cast<const R*>(p));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&t);
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = __uint_as_float(w[i]);
    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {  // bf16 -> fp32 is a 16-bit shift
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
}

// The same in two steps for the gathers: a predicated load into zeroed RAW registers, then an
// unconditional unpack -- cheaper than predicating the unpack of every channel.
template <typename T, int VEC>
__device__ __forceinline__ typename Raw<sizeof(T) * VEC>::type load_raw_if(const bool ok, const T* p) {
    using R = typename Raw<sizeof(T) * VEC>::type;
    R t;
    uint32_t* w = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(R) / 4); ++i) w[i] = 0u;
    if (ok) t = __ldg(reinterpret_cast<const R*>(p));
    return t;
}

template <typename T, int VEC>
__device__ __forceinline__ void unpack_row(const typename Raw<sizeof(T) * VEC>::type& t, float (&v)[VEC]) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&t);
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = __uint_as_float(w[i]);
    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
}

// (a0, a1) += w * (v0, v1) as ONE packed fp32x2 FMA (FFMA2 on sm_100): each half is an IEEE fma, so
// the result is bit-identical to two scalar fmaf calls at half the issue slots.
__device__ __forceinline__ void fma2(float& a0, float& a1, const float w, const float v0, const float v1) {
    uint64_t acc, vv, ww;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(vv) : "f"(v0), "f"(v1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ww), "l"(vv));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc));
}

// acc += a.lo * b.lo, then acc += a.hi * b.hi for two packed bf16 pairs, on the mixed-precision FMA of
// sm_100 (FHFMA.BF16 reads either half of a 32-bit register: no unpack).  A bf16 x bf16 product is exact
// in fp32, so each step equals fmaf(float(a), float(b), acc) bit for bit.
__device__ __forceinline__ void dot2_bf16(float& acc, const uint32_t a, const uint32_t b) {
    asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
        "mov.b32 {al, ah}, %1;\n\t"
        "mov.b32 {bl, bh}, %2;\n\t"
        "fma.rn.f32.bf16 %0, al, bl, %0;\n\t"
        "fma.rn.f32.bf16 %0, ah, bh, %0;\n\t}"
        : "+f"(acc) : "r"(a), "r"(b));
}

// Which kernels use the packed FMA (measured on B200: it pays for 16-bit rows, whose unpack leaves
// the channel pairs in adjacent registers anyway; for fp32 rows it is neutral to slower).
// bit 0: forward gather, bit 1: sample gradients, bit 2: grad_value walk
#ifndef MSDA_PACK_F32
#define MSDA_PACK_F32 0
#endif
#ifndef MSDA_PACK_16
#define MSDA_PACK_16 7
#endif
// sample gradients with bf16 rows: dot products on FHFMA.BF16 instead of unpack + fp32 FMAs
#ifndef MSDA_BF16_MIXED_FMA
#define MSDA_BF16_MIXED_FMA 1
#endif
template <typename T, int BIT>
__host__ __device__ constexpr bool use_packed_fma() { return (((sizeof(T) == 2) ? MSDA_PACK_16 : MSDA_PACK_F32) >> BIT) & 1; }

// (a0, a1) += w * (v0, v1): packed or two scalar FMAs, same bits either way
template <bool PACKED>
__device__ __forceinline__ void axpy2(float& a0, float& a1, const float w, const float v0, const float v1) {
    if constexpr (PACKED) {
        fma2(a0, a1, w, v0, v1);
    } else {
        a0 = fmaf(w, v0, a0);
        a1 = fmaf(w, v1, a1);
    }
}

// (a0, a1) += (u0, u1) * (v0, v1), packed
__device__ __forceinline__ void fma2v(float& a0, float& a1, const float u0, const float u1, const float v0,
                                      const float v1) {
    uint64_t acc, vv, uu;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(vv) : "f"(v0), "f"(v1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(uu) : "f"(u0), "f"(u1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(uu), "l"(vv));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc));
}

template <typename T, int VEC>
__device__ __forceinline__ void store_row(T* p, const float (&v)[VEC]) {
    using R = typename Raw<sizeof(T) * VEC>::type;
    R t;
    uint32_t* w = reinterpret_cast<uint32_t*>(&t);
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) w[i] = __float_as_uint(v[i]);
    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 2; ++i) {
            const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
    }
    *reinterpret_cast<R*>(p) = t;
}

// Streaming loads (data read exactly once: locations, weights, index entries): read-only path,
// no L1 allocation, so that L1 keeps the value / grad_output rows the gathers re-use.
__device__ __forceinline__ uint32_t ld_stream_b32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_b64(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_stream_b128(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream(const float* p) { return __uint_as_float(ld_stream_b32(p)); }
__device__ __forceinline__ float ld_stream(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
}
__device__ __forceinline__ float ld_stream(const __half* p) { return __half2float(__ldg(p)); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldg(p); }

// sampling location (x, y) and attention weight loads in the aux dtype
template <typename CT> struct XY { CT x, y; };
__device__ __forceinline__ XY<float> load_xy(const float* p) {
    const uint2 t = ld_stream_b64(p);
    return {__uint_as_float(t.x), __uint_as_float(t.y)};
}
__device__ __forceinline__ XY<float> load_xy(const __nv_bfloat16* p) {
    const uint32_t t = ld_stream_b32(p);
    return {__uint_as_float(t << 16), __uint_as_float(t & 0xffff0000u)};
}
__device__ __forceinline__ XY<float> load_xy(const __half* p) {
    const uint32_t t = ld_stream_b32(p);
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&t));
    return {f.x, f.y};
}
__device__ __forceinline__ XY<double> load_xy(const double* p) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    return {t.x, t.y};
}

// (x, y) pair store in the aux dtype: one 64-bit (32-bit for 16-bit types) store
__device__ __forceinline__ void store_xy(float* p, float x, float y) { *reinterpret_cast<float2*>(p) = make_float2(x, y); }
__device__ __forceinline__ void store_xy(__nv_bfloat16* p, float x, float y) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(x, y);
}
__device__ __forceinline__ void store_xy(__half* p, float x, float y) {
    *reinterpret_cast<__half2*>(p) = __floats2half2_rn(x, y);
}

// ---------------------------------------------------------------------------------------
// geometry of one sample
template <typename CT> struct Sample {
    bool ok;
    int h_lo, w_lo;
    CT lh, lw;
};

__device__ __forceinline__ Sample<float> locate(float x, float y, int H, int W) {
    Sample<float> s;
    const float h_im = fmaf(y, (float)H, -0.5f);   // one rounding, as nvcc contracts :285-286
    const float w_im = fmaf(x, (float)W, -0.5f);
    s.ok = h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;
    const float hf = floorf(h_im), wf = floorf(w_im);
    s.h_lo = (int)hf; s.w_lo = (int)wf;
    s.lh = h_im - hf; s.lw = w_im - wf;
    return s;
}
__device__ __forceinline__ Sample<double> locate(double x, double y, int H, int W) {
    Sample<double> s;
    const double h_im = fma(y, (double)H, -0.5);
    const double w_im = fma(x, (double)W, -0.5);
    s.ok = h_im > -1.0 && w_im > -1.0 && h_im < (double)H && w_im < (double)W;
    const double hf = floor(h_im), wf = floor(w_im);
    s.h_lo = (int)hf; s.w_lo = (int)wf;
    s.lh = h_im - hf; s.lw = w_im - wf;
    return s;
}

// Pixel index (inside the frame, i.e. level_start + h*W + w) of the four corners,
// -1 for a corner outside the map.
template <typename CT>
__device__ __forceinline__ void corner_pixels(const Sample<CT>& s, const Level& lv, int (&pix)[4]) {
    const int h_hi = s.h_lo + 1, w_hi = s.w_lo + 1;
    const bool h0 = s.h_lo >= 0, w0 = s.w_lo >= 0, h1 = h_hi <= lv.H - 1, w1 = w_hi <= lv.W - 1;
    const int base = lv.start + s.h_lo * lv.W + s.w_lo;
    pix[0] = (h0 && w0) ? base : -1;
    pix[1] = (h0 && w1) ? base + 1 : -1;
    pix[2] = (h1 && w0) ? base + lv.W : -1;
    pix[3] = (h1 && w1) ? base + lv.W + 1 : -1;
}

// Padding mask (fused module prologue): drops the corners that sit on padded pixels.  `flags` holds the four
// in-range bits (top-left, top-right, bottom-left, bottom-right); `mk` points at the mask byte of the top-left corner.
__device__ __forceinline__ unsigned mask_corners(unsigned flags, const uint8_t* __restrict__ mk, const int W) {
    if ((flags & 1u) && mk[0]) flags &= ~1u;
    if ((flags & 2u) && mk[1]) flags &= ~2u;
    if ((flags & 4u) && mk[W]) flags &= ~4u;
    if ((flags & 8u) && mk[W + 1]) flags &= ~8u;
    return flags;
}

// Sub-bin of a sample of query q whose top-left corner is (h_lo, w_lo).
__device__ __forceinline__ int sub_bin(const Level& lv, int h_lo, int w_lo, int q) {
    return lv.bin_start + ((((h_lo + 1) * (lv.W + 1)) + (w_lo + 1)) << lv.nch_log2) + (q & ((1 << lv.nch_log2) - 1));
}

// Read the level table into shared memory (threads 0..L-1) and prefix the bin starts.
// Returns the number of sub-bins through sb and sum_l H*W through sq.
__device__ __forceinline__ void load_levels(const Params& p, Level* lv, int* sb, int* sq) {
    if (threadIdx.x < p.L) {
        const int l = threadIdx.x;
        lv[l].H = (int)p.shapes[2 * l];
        lv[l].W = (int)p.shapes[2 * l + 1];
        lv[l].start = (int)p.lsi[l];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = 0, q = 0;
        for (int l = 0; l < p.L; ++l) {
            const long long nb = (long long)(lv[l].H + 1) * (lv[l].W + 1);
            const long long want = ((long long)p.Lq * p.P + kSubBinTarget * nb - 1) / (kSubBinTarget * nb);
            int k = 0;
            while ((1LL << k) < want && k < kMaxSubLog2) ++k;
            lv[l].nch_log2 = k;
            lv[l].bin_start = b;
            b += (int)(nb << k);
            q += lv[l].H * lv[l].W;
        }
        // The table holds p.sb_max sub-bins (host bound, msda_api.cu: sub_bin_bound, valid when sum_l H_l*W_l <= S).
        // Shapes that disagree with S would overflow it: such a call keeps no index at all (*sb = 0: nothing is
        // counted, filled or walked; grad_value is then left unwritten -- the reference reads out of bounds here).
        *sb = (p.sb_max > 0 && b > p.sb_max) ? 0 : b;
        *sq = q;
    }
    __syncthreads();
}

// false when load_levels() found the level table inconsistent with the index buffer (see there)
__device__ __forceinline__ bool index_usable(const Params& p, const int sb) { return p.sb_max <= 0 || sb > 0; }

// ---------------------------------------------------------------------------------------
// Query tiles.  A tile is a set of up to kTileQ queries of one (frame, head) handled by
// one CTA pass: kTileQ consecutive queries (for encoder self-attention, where the queries
// are the pyramid's own pixels, that is 1.6 image rows of the finest level -- already
// compact enough for L1).  MSDA_FLAG_PYRAMID_TILES switches to kTileH x kTileW pixel
// blocks of one level when Lq == sum_l H_l*W_l
// (/root/reference/models/deformable_transformer.py:273-285); measured slower on B200
// (partial tiles at the coarse levels idle 10 % of the lanes, L1 hit rate is the same), kept
// as an A/B switch.  The choice only affects locality, never results.
struct TileMap {
    int pyramid;          // 1: 2-D tiles per level, 0: linear
    int qtiles;           // tiles per (frame, head)
    int lvl_tile_start[kMaxLevels + 1];  // pyramid: prefix of tiles per level
    int lvl_q_start[kMaxLevels];         // pyramid: first query index of each level
};

__device__ __forceinline__ void build_tile_map(const Params& p, const Level* lv, int sq, int tile_q, TileMap* tm) {
    if (threadIdx.x == 0) {
        const bool pyr = (sq == p.Lq) && (p.flags & 1u) && tile_q == kTileQ;
        tm->pyramid = pyr;
        if (pyr) {
            int t = 0, q = 0;
            for (int l = 0; l < p.L; ++l) {
                tm->lvl_tile_start[l] = t;
                tm->lvl_q_start[l] = q;
                t += ((lv[l].H + kTileH - 1) / kTileH) * ((lv[l].W + kTileW - 1) / kTileW);
                q += lv[l].H * lv[l].W;
            }
            tm->lvl_tile_start[p.L] = t;
            tm->qtiles = t;
        } else {
            tm->qtiles = (p.Lq + tile_q - 1) / tile_q;
        }
    }
    __syncthreads();
}

struct Tile {
    int n, m;
    int q0;        // linear: first query
    int lvl, y0, x0, Wq, Hq, qbase;  // pyramid: level, tile origin, level dims, first query of level
};

__device__ __forceinline__ Tile decode_tile(const Params& p, const Level* lv, const TileMap* tm, int t, int tile_q) {
    Tile tl;
    const int per_frame = tm->qtiles * p.M;
    tl.n = t / per_frame;
    const int r = t - tl.n * per_frame;
    const int qt = r / p.M;
    tl.m = r - qt * p.M;
    tl.q0 = qt * tile_q;
    tl.lvl = 0; tl.y0 = tl.x0 = 0; tl.Wq = tl.Hq = 1; tl.qbase = 0;
    if (tm->pyramid) {
        int l = 0;
        while (l + 1 < p.L && qt >= tm->lvl_tile_start[l + 1]) ++l;
        const int k = qt - tm->lvl_tile_start[l];
        const int tw = (lv[l].W + kTileW - 1) / kTileW;
        tl.lvl = l;
        tl.y0 = (k / tw) * kTileH;
        tl.x0 = (k % tw) * kTileW;
        tl.Wq = lv[l].W; tl.Hq = lv[l].H;
        tl.qbase = tm->lvl_q_start[l];
    }
    return tl;
}

// Query index of slot `idx` (0..tile_q) of a tile, or -1 when the slot is empty.
__device__ __forceinline__ int tile_query(const Params& p, const TileMap* tm, const Tile& tl, int idx) {
    if (tm->pyramid) {
        const int y = tl.y0 + idx / kTileW, x = tl.x0 + idx % kTileW;
        return (y < tl.Hq && x < tl.Wq) ? tl.qbase + y * tl.Wq + x : -1;
    }
    const int q = tl.q0 + idx;
    return q < p.Lq ? q : -1;
}

}  // namespace msda
