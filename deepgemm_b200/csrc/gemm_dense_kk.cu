// Kernel instances: dense GEMM with both operands K-major (fp8_gemm_nt): one CTA pair per tile (clusters 1, 2), the
// weight-multicast clusters (4, 8), cluster split-K (2, 4) and the TMA-store epilogue variant.
#include "launch.cuh"

namespace dgb200 {

template <int kCluster, typename out_t, bool kAcc>
static int launch_plain(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    return launch_kernel(fp8_gemm_kernel<kDense, kCluster, out_t, kAcc>, cfg, c.stream, maps, p);
}
template <int kCluster, typename out_t, bool kAcc>
static int launch_csplit(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    return launch_kernel(fp8_gemm_kernel<kDense, kCluster, out_t, kAcc, false, false, false, true>, cfg, c.stream, maps, p);
}

template <int kCluster>
static int by_output(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    const bool bf16 = c.d_dtype == DGB200_BF16;
    if constexpr (kCluster == 2 || kCluster == 4) {
        if (cfg.csplit) {
            if (bf16) return c.accumulate ? launch_csplit<kCluster, __nv_bfloat16, true>(c, cfg, maps, p)
                                          : launch_csplit<kCluster, __nv_bfloat16, false>(c, cfg, maps, p);
            return c.accumulate ? launch_csplit<kCluster, float, true>(c, cfg, maps, p)
                                : launch_csplit<kCluster, float, false>(c, cfg, maps, p);
        }
    }
    if constexpr (kCluster == 2) {
        if (cfg.tma_store)
            return launch_kernel(fp8_gemm_kernel<kDense, 2, __nv_bfloat16, false, false, false, false, false, true>, cfg,
                                 c.stream, maps, p);
    }
    if (bf16) return c.accumulate ? launch_plain<kCluster, __nv_bfloat16, true>(c, cfg, maps, p)
                                  : launch_plain<kCluster, __nv_bfloat16, false>(c, cfg, maps, p);
    return c.accumulate ? launch_plain<kCluster, float, true>(c, cfg, maps, p) : launch_plain<kCluster, float, false>(c, cfg, maps, p);
}

int dispatch_dense_kk(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    switch (cfg.cluster) {
        case 1: return by_output<1>(c, cfg, maps, p);
        case 2: return by_output<2>(c, cfg, maps, p);
        case 4: return by_output<4>(c, cfg, maps, p);
        case 8: return by_output<8>(c, cfg, maps, p);
        default: return host_fail(DGB200_ERR_INVALID_ARGUMENT, "unsupported cluster size %d for the dense GEMM", cfg.cluster);
    }
}

}  // namespace dgb200
