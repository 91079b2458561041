/*
 * msda_oracle.c -- CPU restatement of the reference's multi-scale deformable
 * attention, forward and backward.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path
 * (neurips2023_soc_b200/) never links, imports or calls it; it fails loudly
 * when the CUDA extension is missing.
 *
 * Parity pin: the reference holds no golden vectors for this path
 * (SURVEY.md 8c).  This restatement is pinned against outputs of the
 * reference's own `ms_deform_attn_core_pytorch`
 * (/root/reference/models/ops/functions/ms_deform_attn_func.py:41-61) and of
 * autograd through it, generated in the build container by
 * tests/golden/make_golden.py and committed under tests/golden/.
 *
 * What is restated, and where it lives in the reference
 * (paths relative to /root/reference/models/ops/src/cuda/):
 *   - sample position:  h_im = loc_y*H - 0.5, w_im = loc_x*W - 0.5
 *                                   ms_deform_im2col_cuda.cuh:285-286
 *   - sample accepted iff -1 < h_im < H and -1 < w_im < W      :288
 *   - bilinear read with each corner dropped individually when it lies
 *     outside [0,H-1]x[0,W-1]                                  :33-84
 *   - value addressed as [n][level_start+h*W+w][m][c]          :47-53,277
 *   - output = sum_{l,p} attn * bilinear                       :272-297
 *   - backward: grad_value += w_k*attn*g  per corner           :116-153
 *               grad_attn   = sum_c g * bilinear               :155-156
 *               grad_loc.x  = W*attn*sum_c g*(-hh*v1+hh*v2-lh*v3+lh*v4)  :157
 *               grad_loc.y  = H*attn*sum_c g*(-hw*v1-lw*v2+hw*v3+lw*v4)  :158
 *     (v1..v4 = corners (lo,lo) (lo,hi) (hi,lo) (hi,hi) in (h,w) order,
 *      lh/lw = fractional parts, hh = 1-lh, hw = 1-lw)
 *   - rejected samples leave zero gradients                    :365-374
 *     (so at h_im == -1 or w_im == -1 exactly the kernels give zero grad_loc
 *      where autograd through grid_sample gives a one-sided slope; a
 *      measure-zero kink on which the reference's CUDA op and its Python path
 *      disagree -- this oracle follows the CUDA op, the thing being replaced)
 *
 * Layouts (all contiguous, row-major):
 *   value        [N][S][M][D]          spatial_shapes [L][2] int64 (H,W)
 *   sampling_loc [N][Lq][M][L][P][2]   (x,y) normalised to [0,1]
 *   attn_weight  [N][Lq][M][L][P]      level_start_index [L] int64
 *   output / grad_output [N][Lq][M*D]
 *
 * The REAL-typed body is instantiated twice (float, double).  The double
 * instance is the truth the fp32 CUDA kernels are checked against; the float
 * instance evaluates in the reference kernel's own arithmetic type.
 * Threads: OpenMP over (n,q) forward and over (n,m) backward -- each (n,m)
 * owns a disjoint slice of grad_value, so the result does not depend on the
 * thread count.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int msda_oracle_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void msda_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

#define FMA_f32 fmaf
#define FMA_f64 fma

#define DEFINE_ORACLE(REAL, SUFFIX)                                                          \
                                                                                             \
/* geometry of one accepted sample: corner row offsets (in pixels of the      */             \
/* level, -1 when that corner is outside the map) and the four weights.       */             \
typedef struct {                                                                             \
    int accepted;                                                                            \
    long pix[4];                                                                             \
    REAL w[4];                                                                               \
    REAL lh, lw, hh, hw;                                                                     \
} sample_##SUFFIX;                                                                           \
                                                                                             \
static sample_##SUFFIX locate_##SUFFIX(REAL loc_x, REAL loc_y, long H, long W)               \
{                                                                                            \
    sample_##SUFFIX s;                                                                       \
    memset(&s, 0, sizeof s);                                                                 \
    /* single-rounding fma: nvcc contracts cuh:285-286 to one FFMA/DFMA by default */     \
    const REAL h_im = (REAL)FMA_##SUFFIX(loc_y, (REAL)H, (REAL)-0.5);                        \
    const REAL w_im = (REAL)FMA_##SUFFIX(loc_x, (REAL)W, (REAL)-0.5);                        \
    if (!(h_im > (REAL)-1 && w_im > (REAL)-1 && h_im < (REAL)H && w_im < (REAL)W))           \
        return s;                                                                            \
    s.accepted = 1;                                                                          \
    const long h_lo = (long)floor((double)h_im), w_lo = (long)floor((double)w_im);           \
    const long h_hi = h_lo + 1, w_hi = w_lo + 1;                                             \
    s.lh = h_im - (REAL)h_lo;                                                                \
    s.lw = w_im - (REAL)w_lo;                                                                \
    s.hh = (REAL)1 - s.lh;                                                                   \
    s.hw = (REAL)1 - s.lw;                                                                   \
    s.w[0] = s.hh * s.hw; s.w[1] = s.hh * s.lw;                                              \
    s.w[2] = s.lh * s.hw; s.w[3] = s.lh * s.lw;                                              \
    s.pix[0] = (h_lo >= 0 && w_lo >= 0)         ? h_lo * W + w_lo : -1;                      \
    s.pix[1] = (h_lo >= 0 && w_hi <= W - 1)     ? h_lo * W + w_hi : -1;                      \
    s.pix[2] = (h_hi <= H - 1 && w_lo >= 0)     ? h_hi * W + w_lo : -1;                      \
    s.pix[3] = (h_hi <= H - 1 && w_hi <= W - 1) ? h_hi * W + w_hi : -1;                      \
    return s;                                                                                \
}                                                                                            \
                                                                                             \
void msda_oracle_forward_##SUFFIX(const REAL *value, const int64_t *shapes,                  \
                                  const int64_t *lsi, const REAL *loc, const REAL *attn,     \
                                  REAL *out, int N, int S, int M, int D, int L, int Lq,      \
                                  int P)                                                     \
{                                                                                            \
    const long row = (long)M * D;                                                            \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                 \
    for (int n = 0; n < N; ++n)                                                              \
        for (int q = 0; q < Lq; ++q)                                                         \
            for (int m = 0; m < M; ++m) {                                                    \
                REAL *o = out + (((long)n * Lq + q) * M + m) * D;                            \
                for (int c = 0; c < D; ++c) o[c] = 0;                                        \
                const long sidx = (((long)n * Lq + q) * M + m) * L * P;                      \
                for (int l = 0; l < L; ++l) {                                                \
                    const long H = shapes[2 * l], W = shapes[2 * l + 1];                     \
                    const REAL *vl = value + ((long)n * S + lsi[l]) * row + (long)m * D;     \
                    for (int p = 0; p < P; ++p) {                                            \
                        const long k = sidx + (long)l * P + p;                               \
                        sample_##SUFFIX s = locate_##SUFFIX(loc[2 * k], loc[2 * k + 1], H, W);\
                        if (!s.accepted) continue;                                           \
                        const REAL a = attn[k];                                              \
                        for (int c = 0; c < D; ++c) {                                        \
                            REAL v[4];                                                       \
                            for (int j = 0; j < 4; ++j)                                      \
                                v[j] = s.pix[j] >= 0 ? vl[s.pix[j] * row + c] : (REAL)0;     \
                            o[c] += (s.w[0] * v[0] + s.w[1] * v[1] + s.w[2] * v[2]           \
                                     + s.w[3] * v[3]) * a;                                   \
                        }                                                                    \
                    }                                                                        \
                }                                                                            \
            }                                                                                \
}                                                                                            \
                                                                                             \
void msda_oracle_backward_##SUFFIX(const REAL *value, const int64_t *shapes,                 \
                                   const int64_t *lsi, const REAL *loc, const REAL *attn,    \
                                   const REAL *grad_out, REAL *grad_value, REAL *grad_loc,   \
                                   REAL *grad_attn, int N, int S, int M, int D, int L,       \
                                   int Lq, int P)                                            \
{                                                                                            \
    const long row = (long)M * D;                                                            \
    memset(grad_value, 0, sizeof(REAL) * (size_t)N * S * row);                               \
    memset(grad_loc, 0, sizeof(REAL) * (size_t)N * Lq * M * L * P * 2);                      \
    memset(grad_attn, 0, sizeof(REAL) * (size_t)N * Lq * M * L * P);                         \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                 \
    for (int n = 0; n < N; ++n)                                                              \
        for (int m = 0; m < M; ++m)                                                          \
            for (int q = 0; q < Lq; ++q) {                                                   \
                const REAL *g = grad_out + (((long)n * Lq + q) * M + m) * D;                 \
                const long sidx = (((long)n * Lq + q) * M + m) * L * P;                      \
                for (int l = 0; l < L; ++l) {                                                \
                    const long H = shapes[2 * l], W = shapes[2 * l + 1];                     \
                    const long base = ((long)n * S + lsi[l]) * row + (long)m * D;            \
                    const REAL *vl = value + base;                                           \
                    REAL *gvl = grad_value + base;                                           \
                    for (int p = 0; p < P; ++p) {                                            \
                        const long k = sidx + (long)l * P + p;                               \
                        sample_##SUFFIX s = locate_##SUFFIX(loc[2 * k], loc[2 * k + 1], H, W);\
                        if (!s.accepted) continue;                                           \
                        const REAL a = attn[k];                                              \
                        REAL ga = 0, gx = 0, gy = 0;                                         \
                        for (int c = 0; c < D; ++c) {                                        \
                            REAL v[4];                                                       \
                            const REAL ag = g[c] * a;                                        \
                            for (int j = 0; j < 4; ++j) {                                    \
                                if (s.pix[j] >= 0) {                                         \
                                    v[j] = vl[s.pix[j] * row + c];                           \
                                    gvl[s.pix[j] * row + c] += s.w[j] * ag;                  \
                                } else                                                       \
                                    v[j] = 0;                                                \
                            }                                                                \
                            ga += g[c] * (s.w[0] * v[0] + s.w[1] * v[1] + s.w[2] * v[2]      \
                                          + s.w[3] * v[3]);                                  \
                            gx += ag * (-s.hh * v[0] + s.hh * v[1] - s.lh * v[2]             \
                                        + s.lh * v[3]);                                      \
                            gy += ag * (-s.hw * v[0] - s.lw * v[1] + s.hw * v[2]             \
                                        + s.lw * v[3]);                                      \
                        }                                                                    \
                        grad_attn[k] = ga;                                                   \
                        grad_loc[2 * k] = (REAL)W * gx;                                      \
                        grad_loc[2 * k + 1] = (REAL)H * gy;                                  \
                    }                                                                        \
                }                                                                            \
            }                                                                                \
}

DEFINE_ORACLE(float, f32)
DEFINE_ORACLE(double, f64)
