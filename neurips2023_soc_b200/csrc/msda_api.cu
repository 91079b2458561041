// msda_api.cu -- the C ABI of libmsda_b200.so (see include/msda_b200.h) and the host-side
// dispatch onto the sm_100a kernels.  No torch / ATen types anywhere: plain pointers,
// sizes, dtype enums and a cudaStream_t.
//
// Host logic mirrors ms_deform_attn_cuda_forward / ms_deform_attn_cuda_backward
// (/root/reference/models/ops/src/cuda/ms_deform_attn_cuda.cu:20-80, :83-153): argument
// validation, the im2col_step divisibility rule (:50-52), one launch sequence per call on
// the caller's stream.  Unlike the reference (cuh:948-952) launch errors are returned.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <type_traits>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>

#include "../../include/msda_b200.h"
#include "msda_backward.cuh"
#include "msda_forward.cuh"
#include "msda_host.h"

// ------------------------------------------------------------------------------------------
// helpers shared with the other translation units (msda_host.h)
namespace msda_host {

namespace {
thread_local std::string g_error;
std::atomic<int> g_launches{0};

// Optional per-kernel timing (bench.py's roofline leg): when enabled, every launch is bracketed
// by CUDA events on the launching stream; msda_profile_read() turns them into milliseconds.
// Process-wide (the backward of an autograd function runs on autograd's own thread).
struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
};
std::mutex g_prof_mu;
std::atomic<bool> g_prof{false};
std::vector<ProfRec> g_recs;
constexpr size_t kProfMax = 1u << 16;   // records (two CUDA events each) kept per msda_profile_enable(1)
}  // namespace

void prof_begin(cudaStream_t st, const char* name) {
    if (!g_prof.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_recs.size() >= kProfMax) {     // a forgotten profile stops recording instead of growing without bound
        g_prof.store(false);
        return;
    }
    ProfRec r{name, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
}
void prof_end(cudaStream_t st) {
    if (!g_prof.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, st);
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

int blocks_per_sm_cached(const void* kernel, int threads, size_t dyn_smem) {
    static std::mutex mu;
    static std::map<std::tuple<const void*, int, int, size_t>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_tuple(kernel, dev, threads, dyn_smem);
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
    }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, dyn_smem) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = 1;
    }
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = n;
    return n;
}

}  // namespace msda_host

namespace {

using namespace msda;
using msda_host::fail;
using msda_host::num_sms;
using msda_host::persistent_grid;
using msda_host::prof_begin;
using msda_host::prof_end;

size_t dtype_size(int dt) {
    switch (dt) {
        case MSDA_F32: return 4;
        case MSDA_BF16: return 2;
        case MSDA_F16: return 2;
        case MSDA_F64: return 8;
        default: return 0;
    }
}

int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int check_common(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                 const void* attn, int N, int S, int M, int D, int L, int Lq, int P, int vdt,
                 int adt, int im2col_step) {
    if (!value || !shapes || !lsi || !loc || !attn)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    if (N <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq <= 0 || P <= 0)
        return fail(MSDA_ERR_INVALID_ARGUMENT,
                    "non-positive size: N=%d S=%d M=%d D=%d L=%d Lq=%d P=%d", N, S, M, D, L, Lq, P);
    if (dtype_size(vdt) == 0 || dtype_size(adt) == 0)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "unknown dtype (value %d, aux %d)", vdt, adt);
    if (adt != vdt && !(adt == MSDA_F32 && vdt != MSDA_F64))
        return fail(MSDA_ERR_INVALID_ARGUMENT,
                    "aux dtype must equal the value dtype or be fp32 (value %d, aux %d)", vdt, adt);
    if (im2col_step <= 0)
        return fail(MSDA_ERR_IM2COL_STEP, "im2col_step must be positive, got %d", im2col_step);
    const int step = N < im2col_step ? N : im2col_step;
    if (N % step != 0)  // ms_deform_attn_cuda.cu:52
        return fail(MSDA_ERR_IM2COL_STEP, "batch(%d) must divide im2col_step(%d)", N, step);
    if (L > kMaxLevels)
        return fail(MSDA_ERR_UNSUPPORTED, "at most %d feature levels are supported, got %d", kMaxLevels, L);
    if ((long long)S * M * D >= (1LL << 31) || (long long)Lq * M * D >= (1LL << 31))
        return fail(MSDA_ERR_UNSUPPORTED, "one frame must hold fewer than 2^31 elements");
    if ((long long)N * M * ((long long)L * Lq * P / 3 + 4LL * S + 4LL * L + 2) >= (1LL << 31))
        return fail(MSDA_ERR_UNSUPPORTED, "N*M*(L*Lq*P/3 + 4S) must stay below 2^31");
    return MSDA_OK;
}

int id_shift_for(int LP) {
    int s = 0;
    while ((1 << s) < LP) ++s;
    return s;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------
// Which kernels a problem takes.  Tile kernels: fp32 rows of 64/128/256 B (D = 16/32/64),
// bf16 rows of 64/128/256 B (D = 32/64/128), P in {4, 8}.  Everything else (fp16, fp64, odd
// channel counts, other P) runs on the generic kernels.
struct Plan {
    bool tile;
    int vec, g;        // sample kernels (forward, grad_loc/grad_attn)
    int wvec, wg;      // grad_value walker
};

Plan make_plan(int D, int P, int vdt, unsigned flags) {
    Plan pl{false, 0, 0, 0, 0};
    if (flags & MSDA_FLAG_GENERIC) return pl;
    if (P != 4 && P != 8) return pl;
    if (vdt == MSDA_F32) {
        if (D != 16 && D != 32 && D != 64) return pl;
        pl = Plan{true, 4, D / 4, 4, D / 4};
    } else if (vdt == MSDA_BF16) {
        if (D == 32) {
            // 64-byte rows: 4 lanes x 128 bit -- the per-sample descriptor / address / weight instructions are
            // paid by half as many lanes (measured on B200: forward 261 -> 225 us, sample gradients 340 -> 299 us);
            // MSDA_FLAG_BF16_VEC4 selects 8 lanes x 64 bit instead (A/B switch).  The walker keeps 8 lanes.
            pl = (flags & MSDA_FLAG_BF16_VEC4) ? Plan{true, 4, 8, 4, 8} : Plan{true, 8, 4, 4, 8};
        } else if (D == 64) {
            pl = Plan{true, 8, 8, 4, 16};
        } else if (D == 128) {
            pl = Plan{true, 8, 16, 8, 16};
        }
    }
    return pl;
}

int rounds_for(int G, int Lq) {
    const int NG = kThreads / G;
    const int full = kTileQ / NG;
    const int need = ceil_div(Lq, NG);
    return need < full ? need : full;
}

// ------------------------------------------------------------------------------------------
// forward
template <typename T, typename TA, int VEC, int G, int P, int ROWB>
int launch_fwd_tile_rowb(const Params& p, cudaStream_t st) {
    const int rounds = rounds_for(G, p.Lq);
    const int tile_q = (kThreads / G) * rounds;
    const long long tiles = (long long)p.N * p.M * ceil_div(p.Lq, tile_q) * 2;  // pyramid tiling may need more passes
    const char* name = p.ref ? "msda_fwd_tile_kernel<fused>" : "msda_fwd_tile_kernel";
    prof_begin(st, name);
    if (p.ref) {   // fused prologue: softmax + sampling locations computed in the staging threads
        if (p.bin_off) {
            auto k = msda_fwd_tile_kernel<T, TA, VEC, G, P, true, true, ROWB>;
            k<<<persistent_grid(k, kThreads, tiles), kThreads, 0, st>>>(p, rounds);
        } else {
            auto k = msda_fwd_tile_kernel<T, TA, VEC, G, P, false, true, ROWB>;
            k<<<persistent_grid(k, kThreads, tiles), kThreads, 0, st>>>(p, rounds);
        }
    } else if (p.bin_off) {   // a backward will follow: count the sub-bin populations on the way
        auto k = msda_fwd_tile_kernel<T, TA, VEC, G, P, true, false, ROWB>;
        k<<<persistent_grid(k, kThreads, tiles), kThreads, 0, st>>>(p, rounds);
    } else {
        auto k = msda_fwd_tile_kernel<T, TA, VEC, G, P, false, false, ROWB>;
        k<<<persistent_grid(k, kThreads, tiles), kThreads, 0, st>>>(p, rounds);
    }
    prof_end(st);
    MSDA_LAUNCHED(name);
    return MSDA_OK;
}

// The row pitch M*D*sizeof(T) is a compile-time constant for d_model = 256 on 8-lane rows (SOC's case:
// 1024 B in fp32, 512 B in bf16); any other pitch takes the run-time variant.
template <typename T, typename TA, int VEC, int G, int P>
int launch_fwd_tile(const Params& p, cudaStream_t st) {
    if constexpr (G == 8 || (G == 4 && sizeof(T) == 2)) {      // D = 32: d_model = 256 gives a compile-time row pitch
        constexpr int kPitch = 256 * (int)sizeof(T);
        if (p.M * p.D * (int)sizeof(T) == kPitch) return launch_fwd_tile_rowb<T, TA, VEC, G, P, kPitch>(p, st);
    }
    return launch_fwd_tile_rowb<T, TA, VEC, G, P, 0>(p, st);
}

template <typename T, typename TA, typename CT>
int launch_fwd_generic(const Params& p, cudaStream_t st) {
    auto k = msda_fwd_generic_kernel<T, TA, CT>;
    const long long blocks = ceil_div((long long)p.N * p.Lq * p.M * p.D, kThreads);
    const long long cap = (long long)num_sms() * 16;
    prof_begin(st, "msda_fwd_generic_kernel");
    k<<<(int)(blocks < cap ? blocks : cap), kThreads, 0, st>>>(p);
    prof_end(st);
    MSDA_LAUNCHED("msda_fwd_generic_kernel");
    return MSDA_OK;
}

// ------------------------------------------------------------------------------------------
// backward
constexpr unsigned kFlagChain = 0x80000000u;   // internal: msda_backward_fused (never set through the public flags)

template <typename T, typename TA, int VEC, int G, int P, int ROWB>
int launch_bwd_sample_tile_rowb(const Params& p, cudaStream_t st) {
    const int rounds = rounds_for(G, p.Lq);
    const int tile_q = (kThreads / G) * rounds;
    const long long tiles = (long long)p.N * p.M * ceil_div(p.Lq, tile_q) * 2;
    if constexpr (std::is_same<T, float>::value) {
        if (p.flags & MSDA_FLAG_ATOMIC_GRAD_VALUE) {
            auto ka = msda_bwd_sample_tile_kernel<T, TA, VEC, G, P, false, true, ROWB>;
            prof_begin(st, "msda_bwd_sample_tile_kernel<atomic>");
            ka<<<persistent_grid(ka, kThreads, tiles), kThreads, 0, st>>>(p, rounds);
            prof_end(st);
            MSDA_LAUNCHED("msda_bwd_sample_tile_kernel<atomic>");
            return MSDA_OK;
        }
    }
    if ((p.flags & kFlagChain) && p.ref != nullptr) {      // raw offsets / logits in, prologue recomputed in the staging threads
        auto kr = msda_bwd_sample_tile_kernel<T, TA, VEC, G, P, true, false, ROWB, true, true>;
        prof_begin(st, "msda_bwd_sample_tile_kernel<chain,raw>");
        kr<<<persistent_grid(kr, kThreads, tiles), kThreads, 0, st>>>(p, rounds);
        prof_end(st);
        MSDA_LAUNCHED("msda_bwd_sample_tile_kernel<chain,raw>");
        return MSDA_OK;
    }
    if (p.flags & kFlagChain) {
        auto kc = msda_bwd_sample_tile_kernel<T, TA, VEC, G, P, true, false, ROWB, true>;
        prof_begin(st, "msda_bwd_sample_tile_kernel<chain>");
        kc<<<persistent_grid(kc, kThreads, tiles), kThreads, 0, st>>>(p, rounds);
        prof_end(st);
        MSDA_LAUNCHED("msda_bwd_sample_tile_kernel<chain>");
        return MSDA_OK;
    }
    auto k = msda_bwd_sample_tile_kernel<T, TA, VEC, G, P, true, false, ROWB>;   // FILL: writes the index entries
    prof_begin(st, "msda_bwd_sample_tile_kernel");
    k<<<persistent_grid(k, kThreads, tiles), kThreads, 0, st>>>(p, rounds);
    prof_end(st);
    MSDA_LAUNCHED("msda_bwd_sample_tile_kernel");
    return MSDA_OK;
}

template <typename T, typename TA, int VEC, int G, int P>
int launch_bwd_sample_tile(const Params& p, cudaStream_t st) {
    if constexpr (G == 8 || (G == 4 && sizeof(T) == 2)) {
        constexpr int kPitch = 256 * (int)sizeof(T);
        if (p.M * p.D * (int)sizeof(T) == kPitch) return launch_bwd_sample_tile_rowb<T, TA, VEC, G, P, kPitch>(p, st);
    }
    return launch_bwd_sample_tile_rowb<T, TA, VEC, G, P, 0>(p, st);
}

template <typename T, typename TA, typename CT>
int launch_bwd_sample_generic(const Params& p, cudaStream_t st) {
    auto k = msda_bwd_sample_generic_kernel<T, TA, CT>;
    const long long blocks = ceil_div((long long)p.N * p.Lq * p.M, kThreads / 32);
    const long long cap = (long long)num_sms() * 8;
    prof_begin(st, "msda_bwd_sample_generic_kernel");
    k<<<(int)(blocks < cap ? blocks : cap), kThreads, 0, st>>>(p);
    prof_end(st);
    MSDA_LAUNCHED("msda_bwd_sample_generic_kernel");
    return MSDA_OK;
}

int elementwise_grid(long long items) {
    const long long eb = ceil_div(items, kThreads);
    const long long cap = (long long)num_sms() * 16;
    return (int)(eb < cap ? eb : cap);
}

// count + scan of the sub-bin populations into p.bin_off (when no forward handed them over)
template <typename TA, typename CT>
int launch_count_scan(const Params& p, bool count, cudaStream_t st) {
    if (count) {
        const long long samples = (long long)p.N * p.Lq * p.M * p.LP;
        prof_begin(st, "msda_bin_count_kernel");
        msda_bin_count_kernel<TA, CT><<<elementwise_grid(samples), kThreads, 0, st>>>(p);
        prof_end(st);
        MSDA_LAUNCHED("msda_bin_count_kernel");
    }
    prof_begin(st, "msda_bin_scan_kernel");
    msda_bin_scan_kernel<<<p.N * p.M, 1024, 0, st>>>(p);
    prof_end(st);
    MSDA_LAUNCHED("msda_bin_scan_kernel");
    return MSDA_OK;
}

template <typename TA, typename CT>
int launch_fill(const Params& p, cudaStream_t st) {
    const long long samples = (long long)p.N * p.Lq * p.M * p.LP;
    prof_begin(st, "msda_bin_fill_kernel");
    msda_bin_fill_kernel<TA, CT><<<elementwise_grid(samples), kThreads, 0, st>>>(p);
    prof_end(st);
    MSDA_LAUNCHED("msda_bin_fill_kernel");
    return MSDA_OK;
}

template <typename CT>
int launch_sort(const Params& p, cudaStream_t st) {
    auto k = msda_bin_rank_sort_kernel<CT>;
    const long long items = (long long)p.N * p.M * ceil_div(p.sb_max, kRankSpan);
    prof_begin(st, "msda_bin_rank_sort_kernel");
    k<<<persistent_grid(k, kThreads, items), kThreads, 0, st>>>(p);
    prof_end(st);
    MSDA_LAUNCHED("msda_bin_rank_sort_kernel");
    if ((long long)p.Lq * p.LP <= kRankMax) return MSDA_OK;   // no sub-bin can hold more than the rank sort takes
    prof_begin(st, "msda_bin_sort_big_kernel");
    msda_bin_sort_big_kernel<CT><<<num_sms() * 2, kThreads, 0, st>>>(p);
    prof_end(st);
    MSDA_LAUNCHED("msda_bin_sort_big_kernel");
    return MSDA_OK;
}

// Dense-level tile of the walker.  A dense tile is one CTA's serial work (thousands of entries, ~150 us for a
// 12 x 8 tile of the coarsest A2D level): with many (frame, head) pairs the big tile wins (least halo re-walk),
// with few the launch is as long as one tile, so the tiles shrink until there are enough of them.
// MSDA_WALK_DENSE_TILE="HxW" overrides (tuning).
void walk_dense_tile(const Params& p, int& thd, int& twd) {
    thd = MSDA_WALK_THD; twd = MSDA_WALK_TWD;
    const long long nm = (long long)p.N * p.M;
    // measured on B200, A2D pyramid, walker us at N = 2 / 8 / 16 frames x 8 heads:
    //   12x8: 148 / 184 / 282    12x4: 100 / 144 / 290    6x4: 99 / 148 / 295    3x4: 65 / 152 / 308    2x2: 65 / 186 / 373
    if (nm < 96) { thd = 12; twd = 4; }
    if (nm < 48) { thd = 6; twd = 4; }
    if (nm < 24) { thd = 3; twd = 4; }
    static const char* env = getenv("MSDA_WALK_DENSE_TILE");
    if (env) { int a = 0, b = 0; if (sscanf(env, "%dx%d", &a, &b) == 2 && a > 0 && b > 0) { thd = a; twd = b; } }
}

template <typename T, int VEC, int G, int TWT = kGTileW, int BRT = 0>
int launch_grad_value_walk(const Params& p, cudaStream_t st) {
    auto k = msda_grad_value_walk_kernel<T, VEC, G, TWT, BRT>;
    int thd, twd;
    walk_dense_tile(p, thd, twd);
    // tiles per (frame, head) are only known on the device; S / 8 bounds them from above
    const long long tiles = (long long)p.N * p.M * (p.S / 8 + p.L);
    prof_begin(st, "msda_grad_value_walk_kernel");
    k<<<persistent_grid(k, kGThreads, tiles), kGThreads, 0, st>>>(p, thd, twd);
    prof_end(st);
    MSDA_LAUNCHED("msda_grad_value_walk_kernel");
    return MSDA_OK;
}

int ceil_log2(long long x) {
    int k = 0;
    while ((1LL << k) < x) ++k;
    return k;
}

constexpr unsigned kFlagDirectAll = 0x40000000u;   // internal: the direct kernel does the whole backward (see backward_typed)

template <typename T, typename TA, int VEC, int G>
int launch_grad_value_direct(const Params& p, cudaStream_t st) {
    const int K = 1 << ceil_log2(4LL * p.Lq * p.P);
    if (p.flags & kFlagDirectAll) {      // zero-fill, sample gradients and grad_value in one launch
        auto ka = msda_grad_value_direct_kernel<T, TA, VEC, G, true>;
        // cooperative launch (grid-wide barrier between the zero-fill and the row writes): the whole grid is resident
        const int grid = num_sms() * msda_host::blocks_per_sm_cached(reinterpret_cast<const void*>(ka), kThreads, 0);
        Params pa = p;
        int k_arg = K < 4 ? 4 : K, id_bits = ceil_log2((long long)p.Lq * p.P);
        void* args[] = {&pa, &k_arg, &id_bits};
        prof_begin(st, "msda_bwd_direct_kernel");
        MSDA_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(ka), dim3(grid), dim3(kThreads), args, 0, st));
        prof_end(st);
        MSDA_LAUNCHED("msda_bwd_direct_kernel");
        return MSDA_OK;
    }
    prof_begin(st, "memset(grad_value)");
    MSDA_CUDA(cudaMemsetAsync(p.grad_value, 0, (size_t)p.N * p.S * p.M * p.D * sizeof(T), st));
    prof_end(st);
    msda_host::count_launch();
    auto k = msda_grad_value_direct_kernel<T, TA, VEC, G>;
    prof_begin(st, "msda_grad_value_direct_kernel");
    k<<<persistent_grid(k, kThreads, (long long)p.N * p.M * p.L), kThreads, 0, st>>>(p, K < 4 ? 4 : K, ceil_log2((long long)p.Lq * p.P));
    prof_end(st);
    MSDA_LAUNCHED("msda_grad_value_direct_kernel");
    return MSDA_OK;
}

template <typename T, typename CT>
int launch_grad_value_generic(const Params& p, cudaStream_t st) {
    auto k = msda_grad_value_generic_kernel<T, CT>;
    const long long blocks = ceil_div((long long)p.N * p.S * p.M, kThreads / 32);
    const long long cap = (long long)num_sms() * 8;
    prof_begin(st, "msda_grad_value_generic_kernel");
    k<<<(int)(blocks < cap ? blocks : cap), kThreads, 0, st>>>(p);
    prof_end(st);
    MSDA_LAUNCHED("msda_grad_value_generic_kernel");
    return MSDA_OK;
}

// ---- (dtype, VEC, G, P) dispatch tables ------------------------------------------------------
#ifdef MSDA_SLIM   // variant builds for A/B runs (tools/build_variant.sh): D = 32, P = 4 only
#define MSDA_FOR_P(CALL, T, TA, VEC, G) (((VEC) * (G) == 32 && p.P == 4) ? CALL<T, TA, (VEC) * (G) == 32 ? (VEC) : 4, (VEC) * (G) == 32 ? (G) : 8, 4>(p, st) : fail(MSDA_ERR_UNSUPPORTED, "slim build"))
#else
#define MSDA_FOR_P(CALL, T, TA, VEC, G)                         \
    (p.P == 4 ? CALL<T, TA, VEC, G, 4>(p, st) : CALL<T, TA, VEC, G, 8>(p, st))
#endif

template <typename TA>
int dispatch_fwd_tile(const Params& p, const Plan& pl, int vdt, cudaStream_t st) {
    if (vdt == MSDA_F32) {
        if constexpr (std::is_same<TA, float>::value) {
            switch (pl.g) {
                case 4: return MSDA_FOR_P(launch_fwd_tile, float, float, 4, 4);
                case 8: return MSDA_FOR_P(launch_fwd_tile, float, float, 4, 8);
                default: return MSDA_FOR_P(launch_fwd_tile, float, float, 4, 16);
            }
        }
        return fail(MSDA_ERR_INVALID_ARGUMENT, "fp32 values need fp32 locations");
    }
    using B = __nv_bfloat16;
    if (pl.vec == 4) return MSDA_FOR_P(launch_fwd_tile, B, TA, 4, 8);
    switch (pl.g) {
        case 4: return MSDA_FOR_P(launch_fwd_tile, B, TA, 8, 4);
        case 8: return MSDA_FOR_P(launch_fwd_tile, B, TA, 8, 8);
        default: return MSDA_FOR_P(launch_fwd_tile, B, TA, 8, 16);
    }
}

template <typename TA>
int dispatch_bwd_sample_tile(const Params& p, const Plan& pl, int vdt, cudaStream_t st) {
    if (vdt == MSDA_F32) {
        if constexpr (std::is_same<TA, float>::value) {
            switch (pl.g) {
                case 4: return MSDA_FOR_P(launch_bwd_sample_tile, float, float, 4, 4);
                case 8: return MSDA_FOR_P(launch_bwd_sample_tile, float, float, 4, 8);
                default: return MSDA_FOR_P(launch_bwd_sample_tile, float, float, 4, 16);
            }
        }
        return fail(MSDA_ERR_INVALID_ARGUMENT, "fp32 values need fp32 locations");
    }
    using B = __nv_bfloat16;
    if (pl.vec == 4) return MSDA_FOR_P(launch_bwd_sample_tile, B, TA, 4, 8);
    switch (pl.g) {
        case 4: return MSDA_FOR_P(launch_bwd_sample_tile, B, TA, 8, 4);
        case 8: return MSDA_FOR_P(launch_bwd_sample_tile, B, TA, 8, 8);
        default: return MSDA_FOR_P(launch_bwd_sample_tile, B, TA, 8, 16);
    }
}

// Calls with few queries per frame (decoder cross-attention) skip the inverse index altogether: the
// contributions of one (frame, head, level) fit in shared memory (msda_grad_value_direct_kernel).  The choice
// depends on per-frame quantities only, so chunking a batch by frames never changes a result.
bool direct_call(int S, int Lq, int P, unsigned flags) {
    if (flags & (MSDA_FLAG_WALK_DENSE | MSDA_FLAG_ATOMIC_GRAD_VALUE | MSDA_FLAG_GENERIC)) return false;
    if (4LL * Lq * P > kDirectMax) return false;
    // pixel | sample | corner in one 32-bit key; strictly fewer than 32 bits so that no real key equals the
    // "no contribution" sentinel ~0u
    return ceil_log2(S) + ceil_log2((long long)Lq * P) + 2 < 32;
}

template <typename TA>
int dispatch_grad_value_direct(const Params& p, const Plan& pl, int vdt, cudaStream_t st) {
    if (vdt == MSDA_F32) {
        if constexpr (std::is_same<TA, float>::value) {
            switch (pl.wg) {
                case 4: return launch_grad_value_direct<float, float, 4, 4>(p, st);
                case 8: return launch_grad_value_direct<float, float, 4, 8>(p, st);
                default: return launch_grad_value_direct<float, float, 4, 16>(p, st);
            }
        }
        return fail(MSDA_ERR_INVALID_ARGUMENT, "fp32 values need fp32 locations");
    }
    using B = __nv_bfloat16;
    if (pl.wvec == 4) return pl.wg == 8 ? launch_grad_value_direct<B, TA, 4, 8>(p, st) : launch_grad_value_direct<B, TA, 4, 16>(p, st);
    return launch_grad_value_direct<B, TA, 8, 16>(p, st);
}

int dispatch_grad_value_walk(const Params& p, const Plan& pl, int vdt, cudaStream_t st) {
    if (vdt == MSDA_F32) {
        switch (pl.wg) {
            case 4: return launch_grad_value_walk<float, 4, 4>(p, st);
            case 8: return launch_grad_value_walk<float, 4, 8>(p, st);
            default: return launch_grad_value_walk<float, 4, 16>(p, st);
        }
    }
    using B = __nv_bfloat16;
    static const bool g4 = getenv("MSDA_WALK_G4") != nullptr;      // experiment: 64-byte bf16 rows on 4 lanes x 128 bit
    if (g4 && pl.wvec == 4 && pl.wg == 8) return launch_grad_value_walk<B, 8, 4, 4, 32>(p, st);
    if (pl.wvec == 4) return pl.wg == 8 ? launch_grad_value_walk<B, 4, 8>(p, st) : launch_grad_value_walk<B, 4, 16>(p, st);
    return launch_grad_value_walk<B, 8, 16>(p, st);
}

// Sub-bins per (frame, head) are only known on the device; bound them: sum_l nb_l * nch_l with
// nb_l = (H_l+1)(W_l+1) and nch_l < 2 * (Lq*P / (T nb_l) + 1), T = kSubBinTarget  =>  < 2*L*Lq*P/T + 2 * sum nb_l, and
// sum nb_l <= 2S + 2L.
int sub_bin_bound(int S, int L, int Lq, int P) {
    return (int)(2LL * L * Lq * P / kSubBinTarget + 4LL * S + 4LL * L + 1);
}

size_t index_bytes(int N, int S, int M, int L, int Lq, int P) {
    return (size_t)N * M * (sub_bin_bound(S, L, Lq, P) + 1) * sizeof(uint32_t);
}

// workspace layout (bytes, every region 256-byte aligned)
struct WsLayout {
    size_t bin_off, counts, big, entries, total;
    int sb_max, big_cap;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

WsLayout ws_layout(int N, int S, int M, int L, int Lq, int P, int vdt) {
    WsLayout w;
    const size_t samples = (size_t)N * Lq * M * L * P;
    w.sb_max = sub_bin_bound(S, L, Lq, P);
    w.big_cap = (int)(samples / (kRankMax + 1) + 1);
    const size_t entry = vdt == MSDA_F64 ? sizeof(Entry<double>) : sizeof(Entry<float>);
    const size_t table = index_bytes(N, S, M, L, Lq, P);
    w.bin_off = 0;
    w.counts = align256(table);
    w.big = align256(w.counts + 4 * sizeof(uint32_t));
    w.entries = align256(w.big + 2 * (size_t)w.big_cap * sizeof(uint32_t));
    w.total = align256(w.entries + samples * entry);
    return w;
}

template <typename T, typename TA, typename CT>
int backward_typed(Params& p, const Plan& pl, int vdt, void* index, size_t table_bytes, cudaStream_t st) {
    int rc;
    bool tile = pl.tile;
    if constexpr (std::is_same<T, double>::value || std::is_same<T, __half>::value) tile = false;
    const bool atomic_arm = (p.flags & MSDA_FLAG_ATOMIC_GRAD_VALUE) != 0;
    if (tile && direct_call(p.S, p.Lq, p.P, p.flags)) {
        if constexpr (!std::is_same<T, double>::value && !std::is_same<T, __half>::value) {
            p.entries = nullptr;                 // no index entries are kept
            if (!(p.flags & (kFlagChain | MSDA_FLAG_DIRECT_SPLIT))) {
                p.flags |= kFlagDirectAll;       // one launch: zero-fill + sample gradients + grad_value
                return dispatch_grad_value_direct<TA>(p, pl, vdt, st);
            }
            if ((rc = dispatch_bwd_sample_tile<TA>(p, pl, vdt, st))) return rc;
            return dispatch_grad_value_direct<TA>(p, pl, vdt, st);
        }
    }
    if (!atomic_arm) {
        // sub-bin offsets: handed over by the forward, or counted and scanned here
        if (index) {
            // consumed: the fill advances the start offsets in place into end offsets
            p.bin_off = static_cast<uint32_t*>(index);
        } else {
            prof_begin(st, "memset(bin table)");
            MSDA_CUDA(cudaMemsetAsync(p.bin_off, 0, table_bytes, st));
            prof_end(st);
            msda_host::count_launch();
            if ((rc = launch_count_scan<TA, CT>(p, true, st))) return rc;
        }
        MSDA_CUDA(cudaMemsetAsync(p.counts, 0, 4 * sizeof(uint32_t), st));
        msda_host::count_launch();
    }
    if (tile) {
        if constexpr (!std::is_same<T, double>::value && !std::is_same<T, __half>::value) {
            if ((rc = dispatch_bwd_sample_tile<TA>(p, pl, vdt, st))) return rc;
        }
    } else {
        if ((rc = launch_bwd_sample_generic<T, TA, CT>(p, st))) return rc;
        if (!atomic_arm && (rc = launch_fill<TA, CT>(p, st))) return rc;
    }
    if (atomic_arm) return MSDA_OK;
    if (tile && (p.flags & MSDA_FLAG_BIN_KERNEL)) {
        // A/B: sort + sum in one kernel over shared-memory pixel tiles (msda_bwd_bin.cuh)
        return msda_host::launch_grad_value_tile(p, vdt, vdt == MSDA_F32 ? 4 : 8, vdt == MSDA_F32 ? p.D / 4 : p.D / 8, st);
    }
    if (!(p.flags & MSDA_FLAG_UNORDERED) && (rc = launch_sort<CT>(p, st))) return rc;
    if (tile) return dispatch_grad_value_walk(p, pl, vdt, st);
    return launch_grad_value_generic<T, CT>(p, st);
}

}  // namespace

// ==========================================================================================
extern "C" {

int msda_version(void) { return MSDA_VERSION; }

const char* msda_last_error(void) { return msda_host::g_error.c_str(); }

int msda_last_launch_count(void) { return msda_host::g_launches.load(); }

void msda_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(msda_host::g_prof_mu);
    for (auto& r : msda_host::g_recs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    msda_host::g_recs.clear();
    msda_host::g_prof.store(on != 0);
}

int msda_profile_read(char* names, size_t names_cap, float* ms, int cap) {
    std::lock_guard<std::mutex> lk(msda_host::g_prof_mu);
    int n = 0;
    size_t used = 0;
    if (names && names_cap) names[0] = 0;
    for (auto& r : msda_host::g_recs) {
        if (n >= cap) break;
        float t = 0.f;
        if (!r.b || cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) t = -1.f;
        ms[n] = t;
        const size_t len = strlen(r.name);
        if (names && used + len + 2 <= names_cap) {
            memcpy(names + used, r.name, len);
            names[used + len] = '\n';
            names[used + len + 1] = 0;
            used += len + 1;
        }
        ++n;
    }
    return n;
}

size_t msda_index_bytes(int N, int S, int M, int D, int L, int Lq, int P) {
    (void)D;
    if (N <= 0 || S <= 0 || M <= 0 || L <= 0 || Lq <= 0 || P <= 0) return 0;
    if (direct_call(S, Lq, P, 0)) return 0;      // few queries per frame: the backward keeps no index at all
    return index_bytes(N, S, M, L, Lq, P);
}

static int forward_impl(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                        const void* sampling_loc, const void* attn_weight, void* output, void* index,
                        size_t index_size, const void* reference_points, void* loc_out, void* attn_out, int N, int S,
                        int M, int D, int L, int Lq, int P, int value_dtype, int aux_dtype, int im2col_step,
                        void* cuda_stream, unsigned flags, const void* value_mask = nullptr) {
    msda_host::g_launches.store(0);
    const msda_host::DeviceGuard guard(value);   // the device that owns `value` is current for this call
    const bool fused = reference_points != nullptr;
    if (fused) {   // the raw offsets / logits may be bf16 next to fp32 values only through the value dtype rule below
        if ((loc_out == nullptr) != (attn_out == nullptr))
            return fail(MSDA_ERR_INVALID_ARGUMENT, "sampling_loc / attn_weight outputs: pass both or neither");
        if (L * P > kSC)
            return fail(MSDA_ERR_UNSUPPORTED, "the fused prologue needs L*P <= %d, got %d", kSC, L * P);
    }
    int rc = check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, N, S, M, D, L,
                          Lq, P, value_dtype, aux_dtype, im2col_step);
    if (rc) return rc;
    if (!output) return fail(MSDA_ERR_INVALID_ARGUMENT, "null output pointer");
    if (!guard.ok()) return fail(MSDA_ERR_INVALID_ARGUMENT, "value must be device memory (a CUDA tensor)");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);

    Params p;
    memset(&p, 0, sizeof p);
    p.value = value; p.shapes = spatial_shapes; p.lsi = level_start_index;
    p.loc = sampling_loc; p.attn = attn_weight; p.out = output;
    p.ref = static_cast<const float*>(reference_points);
    p.loc_out = static_cast<float*>(loc_out); p.attn_out = static_cast<float*>(attn_out);
    p.value_mask = static_cast<const uint8_t*>(value_mask);
    p.N = N; p.S = S; p.M = M; p.D = D; p.L = L; p.Lq = Lq; p.P = P; p.LP = L * P;
    p.id_shift = id_shift_for(p.LP);
    p.flags = flags;
    p.sb_max = sub_bin_bound(S, L, Lq, P);
    if (index && msda_index_bytes(N, S, M, D, L, Lq, P) == 0) index = nullptr;   // no handoff for this shape
    if (index) {
        const size_t need = index_bytes(N, S, M, L, Lq, P);
        if (index_size < need || !aligned16(index))
            return fail(MSDA_ERR_WORKSPACE, "index buffer of %zu bytes (16-byte aligned) required, got %zu", need, index_size);
        p.bin_off = static_cast<uint32_t*>(index);
        prof_begin(st, "memset(index)");
        MSDA_CUDA(cudaMemsetAsync(index, 0, need, st));
        prof_end(st);
        msda_host::count_launch();
    }

    Plan pl = make_plan(D, P, value_dtype, flags);
    if (pl.tile && !(aligned16(value) && aligned16(output) && aligned16(sampling_loc) && aligned16(attn_weight)))
        pl.tile = false;
    if (pl.tile && (long long)S + 65536 >= (1LL << 28)) pl.tile = false;   // descriptor holds a 28-bit pixel index
    if (pl.tile && (long long)S * M * D * (long long)dtype_size(value_dtype) >= (1LL << 31)) pl.tile = false;  // signed 32-bit byte offsets

    const bool aux32 = aux_dtype == MSDA_F32;
    const bool tile = pl.tile && (value_dtype == MSDA_F32 || value_dtype == MSDA_BF16);
    if (fused && !(tile && aligned16(reference_points) && aligned16(loc_out) && aligned16(attn_out)))   /* null is aligned */
        return fail(MSDA_ERR_UNSUPPORTED, "the fused prologue exists for the tile kernels only (fp32/bf16, D and P as in DESIGN.md)");
    if (value_mask && (!tile || (flags & MSDA_FLAG_GENERIC)))
        return fail(MSDA_ERR_UNSUPPORTED, "the padding mask is applied by the tile kernels only");
    switch (value_dtype) {
        case MSDA_F32:
            rc = tile ? dispatch_fwd_tile<float>(p, pl, value_dtype, st) : launch_fwd_generic<float, float, float>(p, st);
            break;
        case MSDA_BF16:
            if (aux32)
                rc = tile ? dispatch_fwd_tile<float>(p, pl, value_dtype, st)
                          : launch_fwd_generic<__nv_bfloat16, float, float>(p, st);
            else
                rc = tile ? dispatch_fwd_tile<__nv_bfloat16>(p, pl, value_dtype, st)
                          : launch_fwd_generic<__nv_bfloat16, __nv_bfloat16, float>(p, st);
            break;
        case MSDA_F16:
            rc = aux32 ? launch_fwd_generic<__half, float, float>(p, st) : launch_fwd_generic<__half, __half, float>(p, st);
            break;
        default:
            rc = launch_fwd_generic<double, double, double>(p, st);
    }
    if (rc || !index) return rc;
    // the tile kernel counted on the way; the generic kernels leave that to a separate pass
    const bool count_now = !tile;
    switch (aux_dtype) {
        case MSDA_F32: return launch_count_scan<float, float>(p, count_now, st);
        case MSDA_BF16: return launch_count_scan<__nv_bfloat16, float>(p, count_now, st);
        case MSDA_F16: return launch_count_scan<__half, float>(p, count_now, st);
        default: return launch_count_scan<double, double>(p, count_now, st);
    }
}

int msda_forward_indexed(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                         const void* sampling_loc, const void* attn_weight, void* output, void* index,
                         size_t index_size, int N, int S, int M, int D, int L, int Lq, int P, int value_dtype,
                         int aux_dtype, int im2col_step, void* cuda_stream, unsigned flags) {
    return forward_impl(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, index, index_size,
                        nullptr, nullptr, nullptr, N, S, M, D, L, Lq, P, value_dtype, aux_dtype, im2col_step,
                        cuda_stream, flags);
}

int msda_forward_fused(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                       const unsigned char* value_padding_mask, const void* reference_points, const void* sampling_offsets, const void* attn_logits,
                       void* output, void* sampling_loc_out, void* attn_weight_out, void* index, size_t index_size,
                       int N, int S, int M, int D, int L, int Lq, int P, int value_dtype, int in_dtype,
                       int im2col_step, void* cuda_stream, unsigned flags) {
    if (!reference_points) return fail(MSDA_ERR_INVALID_ARGUMENT, "null reference_points");
    return forward_impl(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, output, index,
                        index_size, reference_points, sampling_loc_out, attn_weight_out, N, S, M, D, L, Lq, P,
                        value_dtype, in_dtype, im2col_step, cuda_stream, flags, value_padding_mask);
}

int msda_forward_ex(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                    const void* sampling_loc, const void* attn_weight, void* output, int N, int S, int M,
                    int D, int L, int Lq, int P, int value_dtype, int aux_dtype, int im2col_step,
                    void* cuda_stream, unsigned flags) {
    return msda_forward_indexed(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, nullptr,
                                0, N, S, M, D, L, Lq, P, value_dtype, aux_dtype, im2col_step, cuda_stream, flags);
}

int msda_forward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                 const void* sampling_loc, const void* attn_weight, void* output, int N, int S, int M, int D,
                 int L, int Lq, int P, int value_dtype, int aux_dtype, int im2col_step, void* cuda_stream) {
    return msda_forward_ex(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, N, S, M,
                           D, L, Lq, P, value_dtype, aux_dtype, im2col_step, cuda_stream, 0u);
}

size_t msda_backward_workspace_bytes(int N, int S, int M, int D, int L, int Lq, int P, int value_dtype,
                                     int aux_dtype) {
    (void)D; (void)aux_dtype;
    if (N <= 0 || S <= 0 || M <= 0 || L <= 0 || Lq <= 0 || P <= 0) return 0;
    return ws_layout(N, S, M, L, Lq, P, value_dtype).total;
}

static int backward_impl(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                         const void* sampling_loc, const void* attn_weight, const void* grad_output,
                         void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, void* workspace,
                         size_t workspace_bytes, void* index, size_t index_size, int N, int S, int M, int D,
                         int L, int Lq, int P, int value_dtype, int aux_dtype, int im2col_step, void* cuda_stream,
                         unsigned flags, bool chain, const void* reference_points = nullptr,
                         const void* value_mask = nullptr) {
    msda_host::g_launches.store(0);
    const msda_host::DeviceGuard guard(value);   // the device that owns `value` is current for this call
    flags &= ~(kFlagChain | kFlagDirectAll);
    int rc = check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, N, S, M, D, L,
                          Lq, P, value_dtype, aux_dtype, im2col_step);
    if (rc) return rc;
    if (!grad_output || !grad_value || !grad_sampling_loc || !grad_attn_weight)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "null gradient pointer");
    Plan pl = make_plan(D, P, value_dtype, flags);
    if ((flags & MSDA_FLAG_ATOMIC_GRAD_VALUE) && !(value_dtype == MSDA_F32 && pl.tile))
        return fail(MSDA_ERR_UNSUPPORTED, "the atomic A/B arm exists for fp32 tile shapes only");
    const int LP = L * P;
    const int shift = id_shift_for(LP);
    if (((long long)Lq << shift) >= (1LL << 31))
        return fail(MSDA_ERR_UNSUPPORTED, "Lq * next_pow2(L*P) must stay below 2^31");
    const WsLayout w = ws_layout(N, S, M, L, Lq, P, value_dtype);
    const bool need_ws = !(flags & MSDA_FLAG_ATOMIC_GRAD_VALUE);
    if (need_ws && (!workspace || workspace_bytes < w.total))
        return fail(MSDA_ERR_WORKSPACE, "workspace of %zu bytes required, got %zu", w.total,
                    workspace ? workspace_bytes : (size_t)0);
    if (need_ws && !aligned16(workspace))
        return fail(MSDA_ERR_WORKSPACE, "workspace must be 16-byte aligned");
    const size_t table_bytes = index_bytes(N, S, M, L, Lq, P);
    if (index && msda_index_bytes(N, S, M, D, L, Lq, P) == 0) index = nullptr;   // no handoff for this shape
    if (index && index_size < table_bytes)
        return fail(MSDA_ERR_WORKSPACE, "index buffer of %zu bytes required, got %zu", table_bytes, index_size);
    if (!guard.ok()) return fail(MSDA_ERR_INVALID_ARGUMENT, "value must be device memory (a CUDA tensor)");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);

    Params p;
    memset(&p, 0, sizeof p);
    p.value = value; p.shapes = spatial_shapes; p.lsi = level_start_index;
    p.loc = sampling_loc; p.attn = attn_weight; p.grad_out = grad_output;
    p.grad_value = grad_value; p.grad_loc = grad_sampling_loc; p.grad_attn = grad_attn_weight;
    p.ref = static_cast<const float*>(reference_points);     // raw-input chain variant only
    p.value_mask = static_cast<const uint8_t*>(value_mask);
    p.N = N; p.S = S; p.M = M; p.D = D; p.L = L; p.Lq = Lq; p.P = P; p.LP = LP;
    p.id_shift = shift;
    p.sb_max = w.sb_max; p.big_cap = w.big_cap;
    p.flags = flags;
    if (need_ws) {
        char* base = static_cast<char*>(workspace);
        p.bin_off = reinterpret_cast<uint32_t*>(base + w.bin_off);
        p.counts = reinterpret_cast<uint32_t*>(base + w.counts);
        p.big_bins = reinterpret_cast<uint32_t*>(base + w.big);
        p.entries = base + w.entries;
    } else {
        MSDA_CUDA(cudaMemsetAsync(grad_value, 0, (size_t)N * S * M * D * dtype_size(value_dtype), st));
        msda_host::count_launch();
    }

    if (pl.tile && !(aligned16(value) && aligned16(grad_output) && aligned16(grad_value) && aligned16(sampling_loc) &&
                     aligned16(attn_weight) && aligned16(grad_sampling_loc) && aligned16(grad_attn_weight)))
        pl.tile = false;
    if (pl.tile && (long long)S + 65536 >= (1LL << 28)) pl.tile = false;
    if (pl.tile && (long long)S * M * D * (long long)dtype_size(value_dtype) >= (1LL << 31)) pl.tile = false;  // signed 32-bit byte offsets
    if (pl.tile && (long long)N * M * Lq * LP >= (1LL << 32)) pl.tile = false;     // the walker's 32-bit entry positions
    if (chain) {
        if (!(pl.tile && (value_dtype == MSDA_F32 || value_dtype == MSDA_BF16)) || LP > kSC ||
            (flags & (MSDA_FLAG_ATOMIC_GRAD_VALUE | MSDA_FLAG_GENERIC)))
            return fail(MSDA_ERR_UNSUPPORTED, "the fused prologue exists for the tile kernels only (fp32/bf16, D and P as in DESIGN.md, L*P <= %d)", kSC);
        p.flags |= kFlagChain;
    }
    if (value_mask && (!(pl.tile && (value_dtype == MSDA_F32 || value_dtype == MSDA_BF16)) ||
                       (flags & (MSDA_FLAG_ATOMIC_GRAD_VALUE | MSDA_FLAG_GENERIC | MSDA_FLAG_BIN_KERNEL))))
        return fail(MSDA_ERR_UNSUPPORTED, "the padding mask is applied by the tile kernels only (default grad_value path)");

    const bool aux32 = aux_dtype == MSDA_F32;
    switch (value_dtype) {
        case MSDA_F32: return backward_typed<float, float, float>(p, pl, value_dtype, index, table_bytes, st);
        case MSDA_BF16:
            return aux32 ? backward_typed<__nv_bfloat16, float, float>(p, pl, value_dtype, index, table_bytes, st)
                         : backward_typed<__nv_bfloat16, __nv_bfloat16, float>(p, pl, value_dtype, index, table_bytes, st);
        case MSDA_F16:
            return aux32 ? backward_typed<__half, float, float>(p, pl, value_dtype, index, table_bytes, st)
                         : backward_typed<__half, __half, float>(p, pl, value_dtype, index, table_bytes, st);
        default: return backward_typed<double, double, double>(p, pl, value_dtype, index, table_bytes, st);
    }
}

int msda_backward_indexed(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const void* sampling_loc, const void* attn_weight, const void* grad_output,
                          void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, void* workspace,
                          size_t workspace_bytes, void* index, size_t index_size, int N, int S, int M, int D,
                          int L, int Lq, int P, int value_dtype, int aux_dtype, int im2col_step, void* cuda_stream,
                          unsigned flags) {
    return backward_impl(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, grad_value,
                         grad_sampling_loc, grad_attn_weight, workspace, workspace_bytes, index, index_size, N, S, M, D,
                         L, Lq, P, value_dtype, aux_dtype, im2col_step, cuda_stream, flags, false);
}

int msda_backward_fused(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                        const unsigned char* value_padding_mask, const void* sampling_loc, const void* attn_weight, const void* grad_output,
                        void* grad_value, void* grad_sampling_offsets, void* grad_attn_logits, void* workspace,
                        size_t workspace_bytes, void* index, size_t index_size, int N, int S, int M, int D,
                        int L, int Lq, int P, int value_dtype, int aux_dtype, int im2col_step, void* cuda_stream,
                        unsigned flags) {
    return backward_impl(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, grad_value,
                         grad_sampling_offsets, grad_attn_logits, workspace, workspace_bytes, index, index_size, N, S, M,
                         D, L, Lq, P, value_dtype, aux_dtype, im2col_step, cuda_stream, flags, true, nullptr, value_padding_mask);
}

int msda_backward_fused_raw(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                            const unsigned char* value_padding_mask, const void* reference_points, const void* sampling_offsets, const void* attn_logits,
                            const void* grad_output, void* grad_value, void* grad_sampling_offsets, void* grad_attn_logits,
                            void* workspace, size_t workspace_bytes, void* index, size_t index_size, int N, int S, int M,
                            int D, int L, int Lq, int P, int value_dtype, int in_dtype, int im2col_step, void* cuda_stream,
                            unsigned flags) {
    if (!reference_points) return fail(MSDA_ERR_INVALID_ARGUMENT, "null reference_points");
    if (!aligned16(reference_points)) return fail(MSDA_ERR_INVALID_ARGUMENT, "reference_points must be 16-byte aligned");
    if (direct_call(S, Lq, P, flags) || msda_index_bytes(N, S, M, D, L, Lq, P) == 0)
        return fail(MSDA_ERR_UNSUPPORTED, "msda_backward_fused_raw is for calls that keep the inverse index (many queries per frame)");
    if (!index)      // a backward that counts for itself would have to locate the samples from tensors that were never written
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_fused_raw needs the index the matching msda_forward_fused left");
    return backward_impl(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, grad_output, grad_value,
                         grad_sampling_offsets, grad_attn_logits, workspace, workspace_bytes, index, index_size, N, S, M, D,
                         L, Lq, P, value_dtype, in_dtype, im2col_step, cuda_stream, flags, true, reference_points,
                         value_padding_mask);
}

int msda_backward_ex(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                     const void* sampling_loc, const void* attn_weight, const void* grad_output,
                     void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, void* workspace,
                     size_t workspace_bytes, int N, int S, int M, int D, int L, int Lq, int P, int value_dtype,
                     int aux_dtype, int im2col_step, void* cuda_stream, unsigned flags) {
    return msda_backward_indexed(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                 grad_value, grad_sampling_loc, grad_attn_weight, workspace, workspace_bytes, nullptr,
                                 0, N, S, M, D, L, Lq, P, value_dtype, aux_dtype, im2col_step, cuda_stream, flags);
}

int msda_backward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                  const void* sampling_loc, const void* attn_weight, const void* grad_output, void* grad_value,
                  void* grad_sampling_loc, void* grad_attn_weight, void* workspace, size_t workspace_bytes, int N,
                  int S, int M, int D, int L, int Lq, int P, int value_dtype, int aux_dtype, int im2col_step,
                  void* cuda_stream) {
    return msda_backward_ex(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            grad_value, grad_sampling_loc, grad_attn_weight, workspace, workspace_bytes, N, S, M,
                            D, L, Lq, P, value_dtype, aux_dtype, im2col_step, cuda_stream, 0u);
}

}  // extern "C"
