// Issue-only FP8 tensor-core peak probe: the denominator of every tensor-bound roofline fraction this library reports.
//
// BASELINE.md section 2 asks for a measured `tcgen05.mma kind::mxf8f6f4.block_scale` rate instead of the "2 x cuBLAS BF16"
// proxy. This kernel issues exactly the instruction the GEMM issues (same descriptors, cta_group::2, UMMA 256 x N x 32,
// UE8M0 scale factors in tensor memory, FP32 accumulate) on operand tiles that already sit in shared memory: no TMA, no
// global traffic, no epilogue. What it measures is therefore the tensor pipe itself (and the power/clock behaviour of the
// part while it runs flat out), the per-SM rate basis the reference quotes in impls/sm100_bf16_gemm.cuh:382.
//
// Operand bytes are pseudo-random finite E4M3 values (all-zero tiles would draw less power and overstate the clock).
#pragma once
#include "ptx.cuh"

namespace dgb200 {

// grid = 2 * pairs (cluster 2), block = 128. `iters` k-blocks of 4 UMMAs each per CTA pair.
__global__ void __launch_bounds__(128, 1)
fp8_mma_peak_kernel(uint32_t umma_n, uint32_t iters, uint32_t* sink) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    using namespace ptx;
    extern __shared__ __align__(1024) uint8_t smem[];
    // [W 16 KB | X (umma_n / 2) x 128 B <= 16 KB | SFW 512 B | SFX 2 x 512 B | barrier | tmem ptr]
    const uint32_t base = smem_u32(smem);
    const uint32_t off_x = 16384, off_sfw = 32768, off_sfx = off_sfw + 512, off_bar = off_sfx + 1024, off_ptr = off_bar + 8;
    const uint32_t cta_rank = cluster_ctarank();
    const uint32_t warp = threadIdx.x / 32, lane = threadIdx.x % 32;

    // operand tiles: finite pseudo-random E4M3 bytes (exponent field never all-ones); scale bytes 127 = 2^0
    for (uint32_t i = threadIdx.x; i < 32768 / 4; i += blockDim.x) {
        uint32_t h = (i + 1 + blockIdx.x * 8192u) * 2654435761u;
        h ^= h >> 15, h *= 2246822519u, h ^= h >> 13;
        h &= 0xBFBFBFBFu;                                // clear one exponent bit per byte: |value| <= 2^0 * 1.875, no NaN
        reinterpret_cast<uint32_t*>(smem)[i] = h;
    }
    for (uint32_t i = threadIdx.x; i < 1536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + off_sfw)[i] = 0x7F7F7F7Fu;
    if (threadIdx.x == 0) {
        mbar_init(base + off_bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    cluster_arrive_relaxed();
    cluster_wait();
    if (warp == 0) tmem_alloc<2>(base + off_ptr, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = ld_shared_u32(base + off_ptr);

    if (cta_rank == 0 && warp == 1) {
        if (elect_one()) {
            const uint64_t w_desc = make_smem_desc(base, 0, 1024, kLayoutSwizzle128B);
            const uint64_t x_desc = make_smem_desc(base + off_x, 0, 1024, kLayoutSwizzle128B);
            const uint64_t sfw_desc = make_smem_desc(base + off_sfw, 0, 128, kLayoutNoSwizzle);
            const uint64_t sfx_desc = make_smem_desc(base + off_sfx, 0, 128, kLayoutNoSwizzle);
            const uint32_t tmem_sfw = tmem_base + 496, tmem_sfx = tmem_base + 500;
            tmem_cp_sf<2>(tmem_sfw, sfw_desc);
            tmem_cp_sf<2>(tmem_sfx, sfx_desc);
            tmem_cp_sf<2>(tmem_sfx + 4, sfx_desc + 32);
            const uint32_t idesc = make_idesc(256, umma_n, 0, 0);
            // two accumulators when they fit beside the scale-factor columns (N <= 240), like the GEMM; else one
            const uint32_t acc_stride = umma_n <= 240 ? 256u : 0u;
            for (uint32_t it = 0; it < iters; ++it) {
                const uint32_t tmem_d = tmem_base + ((it >> 6) & 1) * acc_stride;   // a new "tile" every 64 k-blocks
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    mma_mxf8_block_scale<2>(tmem_d, w_desc + j * 2, x_desc + j * 2, idesc_with_sf_ids(idesc, it & 3, it & 3),
                                            tmem_sfw, tmem_sfx, ((it & 63) != 0 || j != 0) ? 1u : 0u);
            }
            mma_commit<2>(base + off_bar, 0b11);
        }
        __syncwarp();
    }
    mbar_wait(base + off_bar, 0);            // every MMA of the pair has retired (the commit is multicast to both CTAs)
    tcgen05_fence_after();
    if (warp == 0) {
        // read one accumulator word so that the work is observable (and keep the compiler / hardware honest)
        uint32_t v[8];
        tmem_ld_32x32b_x8(tmem_base, v);
        tmem_ld_wait();
        if (sink != nullptr && lane == 0 && v[0] == 0x12345678u) sink[0] = v[1];
    }
    tcgen05_fence_before();
    cluster_arrive_relaxed();
    cluster_wait();
    if (warp == 0) tmem_dealloc<2>(tmem_base, 512);
#endif
}

}  // namespace dgb200
