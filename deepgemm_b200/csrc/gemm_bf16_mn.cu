// Kernel instances: BF16 x BF16 GEMMs with at least one MN-major operand (fp8_gemm_kernel<..., kXMn, kWMn, ..., kBf16AB>), CTA
// pairs: dense nn / tn / tt (BF16 / FP32 out, optional accumulation), m-grouped contiguous with [G, K, N] weights (+ psum), and
// the k-grouped weight-gradient form with both operands MN-major; plus the batched form behind the BF16 `einsum`.
// Reference: bf16_gemm_{nn,tn,tt}, m_grouped_bf16_gemm_nn_contiguous, k_grouped_bf16_gemm_tn_contiguous
// (csrc/apis/gemm.hpp:440-462, 519-526, 566-608; deep_gemm/include/deep_gemm/impls/sm100_bf16_gemm.cuh:34-420).
#include "launch.cuh"

namespace dgb200 {

template <int kType, typename out_t, bool kAcc, bool kXMn, bool kWMn>
static int launch(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    return launch_kernel(fp8_gemm_kernel<kType, 2, out_t, kAcc, kXMn, kWMn, false, 0, false, false, true>, cfg, c.stream, maps, p);
}

template <bool kXMn, bool kWMn>
static int launch_dense(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (c.d_dtype == DGB200_BF16)
        return c.accumulate ? launch<kDense, __nv_bfloat16, true, kXMn, kWMn>(c, cfg, maps, p)
                            : launch<kDense, __nv_bfloat16, false, kXMn, kWMn>(c, cfg, maps, p);
    return c.accumulate ? launch<kDense, float, true, kXMn, kWMn>(c, cfg, maps, p) : launch<kDense, float, false, kXMn, kWMn>(c, cfg, maps, p);
}

int dispatch_bf16_mn(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (cfg.cluster != 2) return host_fail(DGB200_ERR_UNSUPPORTED, "the BF16 GEMMs need at least 2 SMs (CTA pairs)");
    switch (c.type) {
        case kDense:
            if (c.x_mn && c.w_mn) return launch_dense<true, true>(c, cfg, maps, p);
            return c.x_mn ? launch_dense<true, false>(c, cfg, maps, p) : launch_dense<false, true>(c, cfg, maps, p);
        case kMContiguous:
            if (!c.x_mn && c.w_mn) return launch<kMContiguous, __nv_bfloat16, false, false, true>(c, cfg, maps, p);
            break;
        case kMContiguousPsum:
            if (!c.x_mn && c.w_mn) return launch<kMContiguousPsum, __nv_bfloat16, false, false, true>(c, cfg, maps, p);
            break;
        case kBatched:    // the BF16 einsum forms (einsum.hpp:62-108): tokens K-major, weights [h, d, r] read either way
            if (!c.x_mn && c.d_dtype == DGB200_BF16 && !c.accumulate)
                return c.w_mn ? launch<kBatched, __nv_bfloat16, false, false, true>(c, cfg, maps, p)
                              : launch<kBatched, __nv_bfloat16, false, false, false>(c, cfg, maps, p);
            break;
        case kBatchReduce:   // einsum 'bmk,bnk->mn' (einsum.hpp:22-60): FP32 D accumulated in place
            if (!c.x_mn && !c.w_mn && c.d_dtype == DGB200_FP32 && c.accumulate) return launch<kBatchReduce, float, true, false, false>(c, cfg, maps, p);
            break;
        case kKGrouped:
            if (c.x_mn && c.w_mn) return launch<kKGrouped, float, true, true, true>(c, cfg, maps, p);
            break;
        case kKGroupedPsum:
            if (c.x_mn && c.w_mn) return launch<kKGroupedPsum, float, true, true, true>(c, cfg, maps, p);
            break;
        default: break;
    }
    return host_fail(DGB200_ERR_UNSUPPORTED, "BF16 operands: gemm type %d with majors %d%d is not built", c.type, (int)c.x_mn, (int)c.w_mn);
}

}  // namespace dgb200
