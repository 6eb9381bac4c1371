// Kernel instances: dense GEMM with both operands K-major (fp8_gemm_nt): one CTA pair per tile (clusters 1, 2), the
// weight-multicast clusters (4, 8), cluster split-K (2, 4) and the TMA-store epilogue variant.
#include "launch.cuh"

namespace dgb200 {

template <int kCluster, typename out_t, bool kAcc>
static int launch_plain(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    return launch_kernel(fp8_gemm_kernel<kDense, kCluster, out_t, kAcc>, cfg, c.stream, maps, p);
}
template <int kCluster, int kSlices, typename out_t, bool kAcc>
static int launch_csplit(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    return launch_kernel(fp8_gemm_kernel<kDense, kCluster, out_t, kAcc, false, false, false, kSlices>, cfg, c.stream, maps, p);
}

template <int kCluster, int kSlices>
static int csplit_by_output(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (c.d_dtype == DGB200_BF16)
        return c.accumulate ? launch_csplit<kCluster, kSlices, __nv_bfloat16, true>(c, cfg, maps, p)
                            : launch_csplit<kCluster, kSlices, __nv_bfloat16, false>(c, cfg, maps, p);
    return c.accumulate ? launch_csplit<kCluster, kSlices, float, true>(c, cfg, maps, p)
                        : launch_csplit<kCluster, kSlices, float, false>(c, cfg, maps, p);
}

template <int kCluster>
static int by_output(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    const bool bf16 = c.d_dtype == DGB200_BF16;
    if constexpr (kCluster == 2) {
        if (cfg.tma_store)
            return launch_kernel(fp8_gemm_kernel<kDense, 2, __nv_bfloat16, false, false, false, false, 0, true>, cfg,
                                 c.stream, maps, p);
    }
    if (bf16) return c.accumulate ? launch_plain<kCluster, __nv_bfloat16, true>(c, cfg, maps, p)
                                  : launch_plain<kCluster, __nv_bfloat16, false>(c, cfg, maps, p);
    return c.accumulate ? launch_plain<kCluster, float, true>(c, cfg, maps, p) : launch_plain<kCluster, float, false>(c, cfg, maps, p);
}

int dispatch_dense_kk(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (cfg.csplit) {
        // cluster split-K: `csplit` slices, each a single CTA (cluster == csplit) or a CTA pair (cluster == 2 csplit)
        if (cfg.csplit == 2 && cfg.cluster == 2) return csplit_by_output<2, 2>(c, cfg, maps, p);
        if (cfg.csplit == 4 && cfg.cluster == 4) return csplit_by_output<4, 4>(c, cfg, maps, p);
        if (cfg.csplit == 2 && cfg.cluster == 4) return csplit_by_output<4, 2>(c, cfg, maps, p);
        if (cfg.csplit == 4 && cfg.cluster == 8) return csplit_by_output<8, 4>(c, cfg, maps, p);
        return host_fail(DGB200_ERR_INVALID_ARGUMENT, "unsupported cluster split-K: %d slices in a cluster of %d", cfg.csplit, cfg.cluster);
    }
    switch (cfg.cluster) {
        case 1: return by_output<1>(c, cfg, maps, p);
        case 2: return by_output<2>(c, cfg, maps, p);
        case 4: return by_output<4>(c, cfg, maps, p);
        case 8: return by_output<8>(c, cfg, maps, p);
        default: return host_fail(DGB200_ERR_INVALID_ARGUMENT, "unsupported cluster size %d for the dense GEMM", cfg.cluster);
    }
}

}  // namespace dgb200
