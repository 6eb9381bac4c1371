// Kernel instances: batched GEMM D[b] (+)= A[b] B[b]^T with arbitrary batch strides (3-D tensor maps), all four major
// combinations, BF16 / FP32 output, optional accumulation -- the kernel behind fp8_bmm / fp8_einsum
// (csrc/apis/einsum.hpp:137-214, csrc/jit_kernels/impls/sm100_fp8_fp4_gemm_1d1d.hpp:393-467). CTA pairs only.
#include "launch.cuh"

namespace dgb200 {

template <bool kXMn, bool kWMn>
static int by_output(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (c.d_dtype == DGB200_BF16)
        return c.accumulate ? launch_kernel(fp8_gemm_kernel<kBatched, 2, __nv_bfloat16, true, kXMn, kWMn>, cfg, c.stream, maps, p)
                            : launch_kernel(fp8_gemm_kernel<kBatched, 2, __nv_bfloat16, false, kXMn, kWMn>, cfg, c.stream, maps, p);
    return c.accumulate ? launch_kernel(fp8_gemm_kernel<kBatched, 2, float, true, kXMn, kWMn>, cfg, c.stream, maps, p)
                        : launch_kernel(fp8_gemm_kernel<kBatched, 2, float, false, kXMn, kWMn>, cfg, c.stream, maps, p);
}

int dispatch_batched(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p) {
    if (cfg.cluster != 2) return host_fail(DGB200_ERR_UNSUPPORTED, "the batched GEMM needs at least 2 SMs (CTA pairs)");
    if (c.x_mn && c.w_mn) return by_output<true, true>(c, cfg, maps, p);
    if (c.x_mn) return by_output<true, false>(c, cfg, maps, p);
    if (c.w_mn) return by_output<false, true>(c, cfg, maps, p);
    return by_output<false, false>(c, cfg, maps, p);
}

}  // namespace dgb200
