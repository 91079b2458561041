// msda_bwd_bin.cu -- host side of the bin-major backward pass (see msda_bwd_bin.cuh).
// Its own translation unit so that the library builds in parallel.
#include <atomic>
#include <cstdlib>

#include "msda_bwd_bin.cuh"
#include "msda_host.h"

namespace msda_host {

namespace {

using namespace msda;

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    const int v = atoi(s);
    return v > 0 ? v : dflt;
}

template <typename T, int VEC, int G>
int launch_bin(const Params& p, cudaStream_t st) {
    auto k = msda_bwd_bin_kernel<T, VEC, G>;
    constexpr size_t smem = bin_smem_bytes();
    static std::atomic<unsigned long long> configured{0ull};     // per instantiation: one bit per device
    int dev = 0;
    MSDA_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_relaxed) & bit)) {
        MSDA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.fetch_or(bit, std::memory_order_relaxed);
    }
    constexpr int TPX = kBPixBytes / 4 / (VEC * G);
    // tuning knobs (defaults measured on B200 at the A2D shape): tile width in pixels (tile_w + 1 bins per row: a
    // whole number of 8-bin units), entries a lane group should at least keep when a bin is shared between groups
    static const int tile_w = env_int("MSDA_BIN_TILE_W", 7);
    static const int share_target = env_int("MSDA_BIN_SHARE", 24);
    const int tw0 = tile_w < TPX ? tile_w : TPX;
    // tiles per (frame, head) are only known on the device; every tile holds at least one pixel
    const long long tiles = (long long)p.N * p.M * ((long long)p.S + p.L);
    prof_begin(st, "msda_bwd_bin_kernel");
    k<<<persistent_grid(k, kBThreads, tiles, smem), kBThreads, smem, st>>>(p, tw0, share_target);
    prof_end(st);
    MSDA_LAUNCHED("msda_bwd_bin_kernel");
    return MSDA_OK;
}

}  // namespace

int launch_grad_value_tile(const Params& p, int vdt, int vec, int g, cudaStream_t st) {
    // sub-bins too large for the in-kernel rank step are sorted in place first (usually none: the kernel only scans
    // the offset table); impossible when a (frame, head) holds no more entries than the rank step takes
    if ((long long)p.Lq * p.LP > kBPresort) {
        auto ps = msda_bin_presort_kernel<float>;
        const long long items = (long long)p.N * p.M * ((p.sb_max + kThreads * 8 - 1) / (kThreads * 8));
        prof_begin(st, "msda_bin_presort_kernel");
        ps<<<persistent_grid(ps, kThreads, items), kThreads, 0, st>>>(p, kBPresort);
        prof_end(st);
        MSDA_LAUNCHED("msda_bin_presort_kernel");
    }
    using B = __nv_bfloat16;
    if (vdt == MSDA_F32 && vec == 4) {
        switch (g) {
            case 4: return launch_bin<float, 4, 4>(p, st);
            case 8: return launch_bin<float, 4, 8>(p, st);
            case 16: return launch_bin<float, 4, 16>(p, st);
        }
    } else if (vdt == MSDA_BF16 && vec == 8) {
        switch (g) {
            case 4: return launch_bin<B, 8, 4>(p, st);
            case 8: return launch_bin<B, 8, 8>(p, st);
            case 16: return launch_bin<B, 8, 16>(p, st);
        }
    }
    return fail(MSDA_ERR_UNSUPPORTED, "no bin kernel for dtype %d with %d lanes x %d channels", vdt, g, vec);
}

}  // namespace msda_host
