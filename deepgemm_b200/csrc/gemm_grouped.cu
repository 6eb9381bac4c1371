// Kernel instances: the grouped layouts.
//   contiguous / psum : BF16, no C, tokens K-major, weights {K, MN}-major, clusters 1, 2          (gemm.hpp:181,193)
//   masked            : BF16, no C, both K-major, clusters 1, 2                                   (gemm.hpp:263,275)
//   k-grouped (+psum) : FP32, accumulate into D, both MN-major, clusters 1, 2                     (gemm.hpp:325-328)
#include "launch.cuh"

namespace dgb200 {

template <int kType, bool kWMn>
static int m_grouped(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if constexpr (kType == kMContiguous) {
        if (cfg.cluster == 2 && cfg.tma_store)
            return launch_kernel(fp8_gemm_kernel<kType, 2, __nv_bfloat16, false, false, kWMn, false, 0, true>, cfg, c.stream,
                                 maps, p);
    }
    if (cfg.cluster == 2) return launch_kernel(fp8_gemm_kernel<kType, 2, __nv_bfloat16, false, false, kWMn>, cfg, c.stream, maps, p);
    if (cfg.cluster == 1) return launch_kernel(fp8_gemm_kernel<kType, 1, __nv_bfloat16, false, false, kWMn>, cfg, c.stream, maps, p);
    return host_fail(DGB200_ERR_INVALID_ARGUMENT, "unsupported cluster size %d for gemm type %d", cfg.cluster, kType);
}

template <int kType>
static int k_grouped(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (cfg.cluster == 2) return launch_kernel(fp8_gemm_kernel<kType, 2, float, true, true, true>, cfg, c.stream, maps, p);
    if (cfg.cluster == 1) return launch_kernel(fp8_gemm_kernel<kType, 1, float, true, true, true>, cfg, c.stream, maps, p);
    return host_fail(DGB200_ERR_INVALID_ARGUMENT, "unsupported cluster size %d for gemm type %d", cfg.cluster, kType);
}

int dispatch_grouped(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    switch (c.type) {
        case kMContiguous: return c.w_mn ? m_grouped<kMContiguous, true>(c, cfg, maps, p) : m_grouped<kMContiguous, false>(c, cfg, maps, p);
        case kMContiguousPsum:
            return c.w_mn ? m_grouped<kMContiguousPsum, true>(c, cfg, maps, p) : m_grouped<kMContiguousPsum, false>(c, cfg, maps, p);
        case kMMasked: return m_grouped<kMMasked, false>(c, cfg, maps, p);
        case kKGrouped: return k_grouped<kKGrouped>(c, cfg, maps, p);
        case kKGroupedPsum: return k_grouped<kKGroupedPsum>(c, cfg, maps, p);
        default: return host_fail(DGB200_ERR_INVALID_ARGUMENT, "unknown grouped gemm type %d", c.type);
    }
}

}  // namespace dgb200
