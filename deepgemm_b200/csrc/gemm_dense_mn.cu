// Kernel instances: dense GEMM with an MN-major token and / or weight operand (fp8_gemm_{nn,tn,tt}), clusters 1 and 2,
// plus the TMA-store epilogue variant of each.
#include "launch.cuh"

namespace dgb200 {

template <int kCluster, bool kXMn, bool kWMn>
static int by_output(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if constexpr (kCluster == 2) {
        if (cfg.tma_store)
            return launch_kernel(fp8_gemm_kernel<kDense, 2, __nv_bfloat16, false, kXMn, kWMn, false, 0, true>, cfg, c.stream,
                                 maps, p);
    }
    if (c.d_dtype == DGB200_BF16)
        return c.accumulate ? launch_kernel(fp8_gemm_kernel<kDense, kCluster, __nv_bfloat16, true, kXMn, kWMn>, cfg, c.stream, maps, p)
                            : launch_kernel(fp8_gemm_kernel<kDense, kCluster, __nv_bfloat16, false, kXMn, kWMn>, cfg, c.stream, maps, p);
    return c.accumulate ? launch_kernel(fp8_gemm_kernel<kDense, kCluster, float, true, kXMn, kWMn>, cfg, c.stream, maps, p)
                        : launch_kernel(fp8_gemm_kernel<kDense, kCluster, float, false, kXMn, kWMn>, cfg, c.stream, maps, p);
}

template <bool kXMn, bool kWMn>
static int by_cluster(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (cfg.cluster == 2) return by_output<2, kXMn, kWMn>(c, cfg, maps, p);
    if (cfg.cluster == 1) return by_output<1, kXMn, kWMn>(c, cfg, maps, p);
    return host_fail(DGB200_ERR_INVALID_ARGUMENT, "unsupported cluster size %d for MN-major operands", cfg.cluster);
}

int dispatch_dense_mn(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (c.x_mn && c.w_mn) return by_cluster<true, true>(c, cfg, maps, p);
    if (c.x_mn) return by_cluster<true, false>(c, cfg, maps, p);
    return by_cluster<false, true>(c, cfg, maps, p);
}

}  // namespace dgb200
