// Kernel instances: dense GEMM in the transposed-output orientation (tokens on the TMEM lanes, weights on the columns;
// fp8_gemm_kernel<..., kSwapD>): single CTA (up to 128 token rows per tile) or CTA pair, BF16 / FP32, optional accumulation.
#include "launch.cuh"

namespace dgb200 {

template <int kCluster>
static int by_output(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (cfg.tma_store)
        return launch_kernel(fp8_gemm_kernel<kDense, kCluster, __nv_bfloat16, false, false, false, false, 0, true, true>, cfg, c.stream, maps, p);
    if (c.d_dtype == DGB200_BF16)
        return c.accumulate
                   ? launch_kernel(fp8_gemm_kernel<kDense, kCluster, __nv_bfloat16, true, false, false, false, 0, false, true>, cfg, c.stream, maps, p)
                   : launch_kernel(fp8_gemm_kernel<kDense, kCluster, __nv_bfloat16, false, false, false, false, 0, false, true>, cfg, c.stream, maps, p);
    return c.accumulate ? launch_kernel(fp8_gemm_kernel<kDense, kCluster, float, true, false, false, false, 0, false, true>, cfg, c.stream, maps, p)
                        : launch_kernel(fp8_gemm_kernel<kDense, kCluster, float, false, false, false, false, 0, false, true>, cfg, c.stream, maps, p);
}

int dispatch_dense_swap(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (cfg.cluster == 1) return by_output<1>(c, cfg, maps, p);
    if (cfg.cluster == 2) return by_output<2>(c, cfg, maps, p);
    return host_fail(DGB200_ERR_INVALID_ARGUMENT, "unsupported cluster size %d for the transposed-output orientation", cfg.cluster);
}

}  // namespace dgb200
