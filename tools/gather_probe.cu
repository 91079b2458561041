// gather_probe.cu -- how much of the forward's time is the value layout?  (evidence for DESIGN.md section 7)
//
// The forward and sample-gradient kernels are bound by the L1 data pipe: one wavefront per 64-byte bf16 corner row,
// four rows per sample, and with value laid out [S][M][D] a 128-byte line holds two HEADS of one pixel -- half of
// every line a CTA (one head) pulls into L1 is dead weight, and the two x-adjacent corners of a sample are 512 bytes
// apart.  In a head-major copy [M][S][D] the corner pair (x, x+1) is 128 contiguous bytes: one 256-bit load per lane
// (LDG.E.256 exists on sm_100a) fetches a pair with 4 lanes, i.e. 1 (aligned) or 2 wavefronts instead of 2.
//
// This stand-alone probe times the same gather + bilinear accumulation both ways on the A2D encoder shape with the
// bench's location distribution (descriptors precomputed, so that only the gather differs):
//   A  [S][M][D], 4 lanes x 128 bit x 4 corners          (what msda_fwd_tile_kernel does today)
//   B  [M][S][D], 4 lanes x 256 bit x 2 corner pairs      (lane = column x channel half; one shuffle-add at the end)
//   C  as B for the finest level, the three coarse levels of a (frame, head) staged in shared memory
//   T  the transpose that layouts B / C need once per call
// and checks that A and B agree.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int M = 8, D = 32, L = 4, P = 4, LP = 16;
constexpr int HS[L] = {48, 24, 12, 6}, WS[L] = {80, 40, 20, 10};
__host__ __device__ constexpr int hs(int l) { return 48 >> l; }
__host__ __device__ constexpr int ws(int l) { return 80 >> l; }
constexpr int S = 5100, Lq = 5100;
constexpr int kThreads = 256, kRows = kThreads / 4;      // (query, head) rows per CTA pass

struct Desc { uint32_t pix_flags; float lh, lw, a; };   // pix = level_start + (h_lo+1)*W + (w_lo+1) - (W+1), flags << 28

__device__ __forceinline__ void fma2(float& a0, float& a1, float w, uint32_t packed) {
    a0 = fmaf(w, __uint_as_float(packed << 16), a0);
    a1 = fmaf(w, __uint_as_float(packed & 0xffff0000u), a1);
}

// A: value [n][s][m][d]
__global__ void __launch_bounds__(kThreads, 3) gather_a(const __nv_bfloat16* __restrict__ value, const Desc* __restrict__ desc,
                                                        const int* __restrict__ lvl_w, __nv_bfloat16* __restrict__ out, int N) {
    const int grp = threadIdx.x >> 2, gl = threadIdx.x & 3;
    const int tiles_q = (Lq + kRows - 1) / kRows;
    const int total = N * M * tiles_q;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int n = t / (M * tiles_q), r = t % (M * tiles_q), qt = r / M, m = r % M;
        const int q0 = qt * kRows + grp;
        const bool live = q0 < Lq;
        const int q = live ? q0 : 0;
        const char* fb = reinterpret_cast<const char*>(value) + (((size_t)n * S) * M + m) * D * 2 + gl * 16;
        const Desc* dq = desc + (((size_t)n * Lq + q) * M + m) * LP;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
        for (int s = 0; s < LP; ++s) {
            const Desc d = dq[s];
            const int W = lvl_w[s / P];
            const int pix = (int)(d.pix_flags & 0x0fffffffu) - 65536;     // may point one row / column before the level
            const char* p0 = fb + (long long)pix * (M * D * 2);
            const char* p2 = p0 + (long long)W * (M * D * 2);
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0;
            if (d.pix_flags & (1u << 28)) r0 = __ldg(reinterpret_cast<const uint4*>(p0));
            if (d.pix_flags & (2u << 28)) r1 = __ldg(reinterpret_cast<const uint4*>(p0 + M * D * 2));
            if (d.pix_flags & (4u << 28)) r2 = __ldg(reinterpret_cast<const uint4*>(p2));
            if (d.pix_flags & (8u << 28)) r3 = __ldg(reinterpret_cast<const uint4*>(p2 + M * D * 2));
            const float ah = d.a * (1.f - d.lh), al = d.a * d.lh, hw = 1.f - d.lw;
            const float w0 = ah * hw, w1 = ah * d.lw, w2 = al * hw, w3 = al * d.lw;
            const uint32_t* a0 = &r0.x; const uint32_t* a1 = &r1.x; const uint32_t* a2 = &r2.x; const uint32_t* a3 = &r3.x;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fma2(acc[2 * i], acc[2 * i + 1], w0, a0[i]);
                fma2(acc[2 * i], acc[2 * i + 1], w1, a1[i]);
                fma2(acc[2 * i], acc[2 * i + 1], w2, a2[i]);
                fma2(acc[2 * i], acc[2 * i + 1], w3, a3[i]);
            }
        }
        uint4 o;
        uint32_t* ow = &o.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
            ow[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        if (live) *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + ((((size_t)n * Lq + q) * M + m) * D) * 2 + gl * 16) = o;
    }
}

struct U8 { uint32_t w[8]; };
__device__ __forceinline__ U8 ld256(const void* p) {
    U8 r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}

// B: value_t [n][m][s][d]; lane gl: column = gl >> 1 (left / right corner), half = gl & 1 (channels 16*half ..)
__global__ void __launch_bounds__(kThreads, 3) gather_b(const __nv_bfloat16* __restrict__ value_t, const Desc* __restrict__ desc,
                                                        const int* __restrict__ lvl_w, __nv_bfloat16* __restrict__ out, int N) {
    const int grp = threadIdx.x >> 2, gl = threadIdx.x & 3;
    const int col = gl >> 1;
    const int tiles_q = (Lq + kRows - 1) / kRows;
    const int total = N * M * tiles_q;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int n = t / (M * tiles_q), r = t % (M * tiles_q), qt = r / M, m = r % M;
        const int q0 = qt * kRows + grp;
        const bool live = q0 < Lq;
        const int q = live ? q0 : 0;
        // the pair (x, x+1) of one image row is 128 contiguous bytes; this lane's 32 of them
        const char* fb = reinterpret_cast<const char*>(value_t) + (((size_t)n * M + m) * S) * D * 2 + gl * 32;
        const Desc* dq = desc + (((size_t)n * Lq + q) * M + m) * LP;
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll 4
        for (int s = 0; s < LP; ++s) {
            const Desc d = dq[s];
            const int W = lvl_w[s / P];
            const int pix = (int)(d.pix_flags & 0x0fffffffu) - 65536;
            const char* p0 = fb + (long long)pix * (D * 2);
            const char* p2 = p0 + (long long)W * (D * 2);
            U8 rt, rb;
#pragma unroll
            for (int i = 0; i < 8; ++i) rt.w[i] = rb.w[i] = 0u;
            if ((d.pix_flags >> (28 + col)) & 1u) rt = ld256(p0);
            if ((d.pix_flags >> (30 + col)) & 1u) rb = ld256(p2);
            const float cw = col ? d.lw : 1.f - d.lw;
            const float wt = d.a * (1.f - d.lh) * cw, wb = d.a * d.lh * cw;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                fma2(acc[2 * i], acc[2 * i + 1], wt, rt.w[i]);
                fma2(acc[2 * i], acc[2 * i + 1], wb, rb.w[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 2);
        if (col == 0 && live) {
            U8 o;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
                o.w[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + ((((size_t)n * Lq + q) * M + m) * D) * 2 + gl * 32);
            dst[0] = make_uint4(o.w[0], o.w[1], o.w[2], o.w[3]);
            dst[1] = make_uint4(o.w[4], o.w[5], o.w[6], o.w[7]);
        }
    }
}

// C: as B for level 0; levels 1..3 of the CTA's (frame, head) are staged once in shared memory with a zero border
// (26x42 + 14x22 + 8x12 pixels x 64 B = 95,744 B), so that their corner pairs are plain shared-memory reads: no L1
// tags, no L2 misses, no corner predicates.  A CTA of 512 threads (128 lane groups) takes one (frame, head, quarter of
// the queries); two CTAs fit an SM.
constexpr int kCThreads = 512, kCRanges = 4;
constexpr int kPadBytes = (26 * 42 + 14 * 22 + 8 * 12) * 64;

__global__ void __launch_bounds__(kCThreads, 2) gather_c(const __nv_bfloat16* __restrict__ value_t, const Desc* __restrict__ desc,
                                                          const int* __restrict__ ppad, __nv_bfloat16* __restrict__ out, int N) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int grp = threadIdx.x >> 2, gl = threadIdx.x & 3, col = gl >> 1;
    const int total = N * M * kCRanges;
    const int per = (Lq + kCRanges - 1) / kCRanges;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int nm = t / kCRanges, rg = t % kCRanges;
        const int n = nm / M, m = nm % M;
        const char* base = reinterpret_cast<const char*>(value_t) + ((size_t)nm * S) * D * 2;
        __syncthreads();
        // zero everything, then copy the three levels into their padded frames (16 B per thread and step)
        for (int i = threadIdx.x; i < kPadBytes / 16; i += kCThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        int soff = 0, pstart = hs(0) * ws(0);
#pragma unroll
        for (int l = 1; l < L; ++l) {
            const int H = hs(l), W = ws(l);
            for (int i = threadIdx.x; i < H * W * 4; i += kCThreads) {
                const int c = i & 3, pxl = i >> 2, y = pxl / W, x = pxl - y * W;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)(pstart + pxl)) * 64) + c);
                *reinterpret_cast<uint4*>(smem + soff + ((y + 1) * (W + 2) + x + 1) * 64 + c * 16) = v;
            }
            soff += (H + 2) * (W + 2) * 64;
            pstart += H * W;
        }
        __syncthreads();
        const char* fb = base + gl * 32;
        for (int it = 0; it < (per + kCThreads / 4 - 1) / (kCThreads / 4); ++it) {      // same trip count for every lane
            const int q0 = rg * per + it * (kCThreads / 4) + grp;
            const bool live = q0 < min(Lq, rg * per + per);
            const int q = live ? q0 : 0;
            const Desc* dq = desc + (((size_t)n * Lq + q) * M + m) * LP;
            const int* pq = ppad + (((size_t)n * Lq + q) * M + m) * LP;
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            // level 0: global, as in B
#pragma unroll
            for (int s = 0; s < P; ++s) {
                const Desc d = dq[s];
                const int pix = (int)(d.pix_flags & 0x0fffffffu) - 65536;
                const char* p0 = fb + (long long)pix * 64;
                const char* p2 = p0 + ws(0) * 64;
                U8 rt, rb;
#pragma unroll
                for (int i = 0; i < 8; ++i) rt.w[i] = rb.w[i] = 0u;
                if ((d.pix_flags >> (28 + col)) & 1u) rt = ld256(p0);
                if ((d.pix_flags >> (30 + col)) & 1u) rb = ld256(p2);
                const float cw = col ? d.lw : 1.f - d.lw;
                const float wt = d.a * (1.f - d.lh) * cw, wb = d.a * d.lh * cw;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    fma2(acc[2 * i], acc[2 * i + 1], wt, rt.w[i]);
                    fma2(acc[2 * i], acc[2 * i + 1], wb, rb.w[i]);
                }
            }
            // levels 1..3: shared memory, zero border instead of corner flags
#pragma unroll
            for (int l = 1; l < L; ++l) {
                const int W = ws(l);
#pragma unroll
                for (int pt = 0; pt < P; ++pt) {
                    const Desc d = dq[l * P + pt];
                    // byte offset of the top-left corner in the padded frames (precomputed: (h_lo+1)*(W+2) + (w_lo+1);
                    // a rejected sample has weight 0 and reads the frame's first pixels)
                    const unsigned char* p0 = smem + pq[l * P + pt] * 64 + gl * 32;
                    const unsigned char* p2 = p0 + (W + 2) * 64;
                    const uint4 t0 = *reinterpret_cast<const uint4*>(p0), t1 = *reinterpret_cast<const uint4*>(p0 + 16);
                    const uint4 b0 = *reinterpret_cast<const uint4*>(p2), b1 = *reinterpret_cast<const uint4*>(p2 + 16);
                    const float cw = col ? d.lw : 1.f - d.lw;
                    const float wt = d.a * (1.f - d.lh) * cw, wb = d.a * d.lh * cw;
                    const uint32_t tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                    const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        fma2(acc[2 * i], acc[2 * i + 1], wt, tw[i]);
                        fma2(acc[2 * i], acc[2 * i + 1], wb, bw[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 2);
            if (col == 0 && live) {
                U8 o;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
                    o.w[i] = *reinterpret_cast<const uint32_t*>(&hh);
                }
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + ((((size_t)n * Lq + q) * M + m) * D) * 2 + gl * 32);
                dst[0] = make_uint4(o.w[0], o.w[1], o.w[2], o.w[3]);
                dst[1] = make_uint4(o.w[4], o.w[5], o.w[6], o.w[7]);
            }
        }
    }
}

// T: [n][s][m][d] -> [n][m][s][d], 16 bytes per thread
__global__ void transpose_heads(const uint4* __restrict__ in, uint4* __restrict__ out, int N) {
    const size_t total = (size_t)N * S * M * 4;              // 4 x 16 B per 64-byte row
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i & 3);
        const size_t row = i >> 2;                            // (n*S + s)*M + m
        const int m = (int)(row % M);
        const size_t ns = row / M;
        const int s = (int)(ns % S);
        const size_t n = ns / S;
        out[(((n * M + m) * S + s) << 2) + c] = in[i];
    }
}

static float gauss() {
    const float u1 = (rand() + 1.f) / ((float)RAND_MAX + 2.f), u2 = (rand() + 1.f) / ((float)RAND_MAX + 2.f);
    return sqrtf(-2.f * logf(u1)) * cosf(6.2831853f * u2);
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 16;
    const int iters = 20;
    srand(1);
    int lstart[L], acc_s = 0;
    for (int l = 0; l < L; ++l) { lstart[l] = acc_s; acc_s += HS[l] * WS[l]; }
    // descriptors with the bench's "encoder" distribution: reference point = the query's own pixel centre, offsets =
    // compass direction of the head times (p+1) pixels + N(0, 1 px)
    std::vector<Desc> hd((size_t)N * Lq * M * LP);
    std::vector<int> hp((size_t)N * Lq * M * LP, 0);          // variant C: pixel offset in the padded shared-memory frames
    int padstart[L] = {0, 0, 0, 0};
    for (int l = 2; l < L; ++l) padstart[l] = padstart[l - 1] + (HS[l - 1] + 2) * (WS[l - 1] + 2);
    std::vector<float> rx(Lq), ry(Lq);
    for (int l = 0, q = 0; l < L; ++l)
        for (int y = 0; y < HS[l]; ++y)
            for (int x = 0; x < WS[l]; ++x, ++q) { rx[q] = (x + 0.5f) / WS[l]; ry[q] = (y + 0.5f) / HS[l]; }
    for (int n = 0; n < N; ++n)
        for (int q = 0; q < Lq; ++q)
            for (int m = 0; m < M; ++m) {
                const float th = m * (6.2831853f / M);
                float dx = cosf(th), dy = sinf(th);
                const float mx = fmaxf(fabsf(dx), fabsf(dy));
                dx /= mx; dy /= mx;
                for (int l = 0; l < L; ++l)
                    for (int p = 0; p < P; ++p) {
                        const float lx = rx[q] + (dx * (p + 1) + gauss()) / WS[l], ly = ry[q] + (dy * (p + 1) + gauss()) / HS[l];
                        const float him = ly * HS[l] - 0.5f, wim = lx * WS[l] - 0.5f;
                        Desc d = {65536u, 0.f, 0.f, 0.f};
                        if (him > -1.f && wim > -1.f && him < HS[l] && wim < WS[l]) {
                            const int h0 = (int)floorf(him), w0 = (int)floorf(wim);
                            const unsigned t = h0 >= 0, b = h0 + 1 <= HS[l] - 1, lft = w0 >= 0, rgt = w0 + 1 <= WS[l] - 1;
                            const unsigned flags = (t & lft) | ((t & rgt) << 1) | ((b & lft) << 2) | ((b & rgt) << 3);
                            d.pix_flags = (flags << 28) | (uint32_t)(lstart[l] + h0 * WS[l] + w0 + 65536);
                            d.lh = him - h0; d.lw = wim - w0; d.a = 1.f / LP;
                            if (l >= 1) hp[((((size_t)n * Lq + q) * M + m) * L + l) * P + p] = padstart[l] + (h0 + 1) * (WS[l] + 2) + (w0 + 1);
                        }
                        hd[((((size_t)n * Lq + q) * M + m) * L + l) * P + p] = d;
                    }
            }
    std::vector<__nv_bfloat16> hv((size_t)N * S * M * D);
    for (auto& v : hv) v = __float2bfloat16(gauss());

    __nv_bfloat16 *value, *value_t, *out_a, *out_b;
    Desc* desc;
    int* lw;
    char* flush;
    const size_t vbytes = hv.size() * 2, obytes = (size_t)N * Lq * M * D * 2;
    CK(cudaMalloc(&value, vbytes)); CK(cudaMalloc(&value_t, vbytes)); CK(cudaMalloc(&out_a, obytes)); CK(cudaMalloc(&out_b, obytes));
    CK(cudaMalloc(&desc, hd.size() * sizeof(Desc))); CK(cudaMalloc(&lw, sizeof(WS))); CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemcpy(value, hv.data(), vbytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(desc, hd.data(), hd.size() * sizeof(Desc), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(lw, WS, sizeof(WS), cudaMemcpyHostToDevice));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 3;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](auto launch, const char* name) {
        float best = 1e9f, sum = 0.f;
        for (int i = 0; i < iters + 2; ++i) {
            CK(cudaMemsetAsync(flush, i, 256 << 20));           // push the inputs out of L2
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (i >= 2) { best = fminf(best, ms); sum += ms; }
        }
        printf("{\"kernel\": \"%s\", \"N\": %d, \"us_mean\": %.1f, \"us_best\": %.1f}\n", name, N, sum / iters * 1e3f, best * 1e3f);
    };
    timeit([&] { transpose_heads<<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(value), reinterpret_cast<uint4*>(value_t), N); },
           "T transpose [S][M][D] -> [M][S][D]");
    timeit([&] { gather_a<<<grid, kThreads>>>(value, desc, lw, out_a, N); }, "A [S][M][D], 4 lanes x 128 bit x 4 corners");
    timeit([&] { gather_b<<<grid, kThreads>>>(value_t, desc, lw, out_b, N); }, "B [M][S][D], 4 lanes x 256 bit x 2 corner pairs");
    __nv_bfloat16* out_c;
    CK(cudaMalloc(&out_c, obytes));
    int* ppad;
    CK(cudaMalloc(&ppad, hp.size() * sizeof(int)));
    CK(cudaMemcpy(ppad, hp.data(), hp.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(gather_c, cudaFuncAttributeMaxDynamicSharedMemorySize, kPadBytes));
    timeit([&] { gather_c<<<sms * 2, kCThreads, kPadBytes>>>(value_t, desc, ppad, out_c, N); },
           "C [M][S][D], level 0 as B, levels 1-3 staged in shared memory (zero border)");
    CK(cudaGetLastError());
    {
        std::vector<__nv_bfloat16> hb2(obytes / 2), hc(obytes / 2);
        CK(cudaMemcpy(hb2.data(), out_b, obytes, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hc.data(), out_c, obytes, cudaMemcpyDeviceToHost));
        double md = 0;
        for (size_t i = 0; i < hc.size(); ++i) md = fmax(md, fabs((double)__bfloat162float(hb2[i]) - (double)__bfloat162float(hc[i])));
        printf("{\"max_abs_diff_B_vs_C\": %.3g}\n", md);
    }
    std::vector<__nv_bfloat16> ha(obytes / 2), hb(obytes / 2);
    CK(cudaMemcpy(ha.data(), out_a, obytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), out_b, obytes, cudaMemcpyDeviceToHost));
    double maxd = 0, maxv = 0;
    for (size_t i = 0; i < ha.size(); ++i) {
        maxd = fmax(maxd, fabs((double)__bfloat162float(ha[i]) - (double)__bfloat162float(hb[i])));
        maxv = fmax(maxv, fabs((double)__bfloat162float(ha[i])));
    }
    printf("{\"max_abs_diff_A_vs_B\": %.3g, \"max_abs_value\": %.3g}\n", maxd, maxv);
    return 0;
}
