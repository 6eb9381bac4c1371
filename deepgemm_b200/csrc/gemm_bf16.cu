// Kernel instances: BF16 x BF16 GEMMs (no scale factors) on the FP8 kernel's skeleton (fp8_gemm_kernel<..., kBf16AB>), K-major
// operands, CTA pairs: dense (BF16 / FP32 out, optional accumulation; + the cluster split-K forms for small / medium M),
// m-grouped contiguous (+ psum) and masked.
// Reference: bf16_gemm_nt, m_grouped_bf16_gemm_nt_contiguous, m_grouped_bf16_gemm_nt_masked (csrc/apis/gemm.hpp:404-564,
// deep_gemm/include/deep_gemm/impls/sm100_bf16_gemm.cuh:34-420).
#include "launch.cuh"

namespace dgb200 {

template <int kType, typename out_t, bool kAcc>
static int launch(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    return launch_kernel(fp8_gemm_kernel<kType, 2, out_t, kAcc, false, false, false, 0, false, false, true>, cfg, c.stream, maps, p);
}

template <int kCluster, int kSlices>
static int launch_csplit(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (c.d_dtype == DGB200_BF16)
        return c.accumulate ? launch_kernel(fp8_gemm_kernel<kDense, kCluster, __nv_bfloat16, true, false, false, false, kSlices, false, false, true>, cfg, c.stream, maps, p)
                            : launch_kernel(fp8_gemm_kernel<kDense, kCluster, __nv_bfloat16, false, false, false, false, kSlices, false, false, true>, cfg, c.stream, maps, p);
    return c.accumulate ? launch_kernel(fp8_gemm_kernel<kDense, kCluster, float, true, false, false, false, kSlices, false, false, true>, cfg, c.stream, maps, p)
                        : launch_kernel(fp8_gemm_kernel<kDense, kCluster, float, false, false, false, false, kSlices, false, false, true>, cfg, c.stream, maps, p);
}

int dispatch_bf16(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (cfg.csplit && c.type == kDense) {
        // cluster split-K (small / medium M), as for the FP8 operands: four single-CTA slices, or two CTA-pair slices
        if (cfg.csplit == 4 && cfg.cluster == 4) return launch_csplit<4, 4>(c, cfg, maps, p);
        if (cfg.csplit == 2 && cfg.cluster == 4) return launch_csplit<4, 2>(c, cfg, maps, p);
        return host_fail(DGB200_ERR_UNSUPPORTED, "BF16 operands: cluster split-K with %d slices in a cluster of %d is not built", cfg.csplit, cfg.cluster);
    }
    if (cfg.cluster != 2) return host_fail(DGB200_ERR_UNSUPPORTED, "the BF16 GEMMs need at least 2 SMs (CTA pairs)");
    switch (c.type) {
        case kDense:
            if (c.d_dtype == DGB200_BF16)
                return c.accumulate ? launch<kDense, __nv_bfloat16, true>(c, cfg, maps, p) : launch<kDense, __nv_bfloat16, false>(c, cfg, maps, p);
            return c.accumulate ? launch<kDense, float, true>(c, cfg, maps, p) : launch<kDense, float, false>(c, cfg, maps, p);
        case kMContiguous: return launch<kMContiguous, __nv_bfloat16, false>(c, cfg, maps, p);
        case kMContiguousPsum: return launch<kMContiguousPsum, __nv_bfloat16, false>(c, cfg, maps, p);
        case kMMasked: return launch<kMMasked, __nv_bfloat16, false>(c, cfg, maps, p);
        default: return host_fail(DGB200_ERR_UNSUPPORTED, "BF16 operands: gemm type %d is not built", c.type);
    }
}

}  // namespace dgb200
