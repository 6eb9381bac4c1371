// Shared between the host API translation unit (dgb200_api.cu) and the kernel-instance translation units
// (gemm_*.cu): the launch record, the chosen configuration and the `cudaLaunchKernelEx` wrapper.
//
// The ahead-of-time instantiation menu (~100 kernels) is cut into several translation units so that the library builds in
// parallel; each unit exports ONE plain host function (`dispatch_*`) and keeps its kernels to itself, so no relocatable
// device code is needed. Replaces the reference's per-shape JIT (csrc/jit/*, csrc/jit_kernels/impls/sm100_fp8_fp4_gemm_1d1d.hpp:93-391).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "../../include/dgb200.h"
#include "fp8_gemm_kernel.cuh"

namespace dgb200 {

// ---- defined in dgb200_api.cu
int host_fail(int code, const char* fmt, ...);
int rt_device();
int rt_pdl();
void count_launches(int n);

#define DGB_REQUIRE(cond)                                                                                          \
    do {                                                                                                           \
        if (!(cond))                                                                                               \
            return host_fail(DGB200_ERR_INVALID_ARGUMENT, "Assertion error (%s:%d): %s", __FILE__, __LINE__, #cond); \
    } while (0)

#define DGB_CUDA(call)                                                                                                  \
    do {                                                                                                                \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess)                                                                                          \
            return host_fail(DGB200_ERR_CUDA, "CUDA runtime error (%s:%d): %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

constexpr int kSmemCapacity = 232448;  // 227 KB usable per CTA on sm_100 (heuristics/sm100.hpp:15)

struct Config {
    int block_m, cluster, stages, num_sms, smem_bytes, swizzle_group;
    int num_splits, kb_per_split;   // split-K (dense, small problems): K cut into num_splits ranges
    int csplit;                     // cluster split-K: `cluster` single-CTA MMAs share one tile (then num_splits == cluster)
    int grid, grid_y;               // grid == 0: persistent grid over num_sms; else exactly grid x grid_y CTAs
    int num_tall = 0, block_m_low = 0;   // dense wave balancing: first num_tall m-blocks block_m high, the rest block_m_low
    bool overlap_producer = false;  // launch as a programmatic dependent that does not wait for the preceding kernel
    int tma_store = 0;              // BF16 output staged through shared memory and written with TMA stores
    int swap_d = 0;                 // transposed-output orientation (tokens on the TMEM lanes)
};

struct GemmCall {
    int type;
    const void* a;
    const void* b;
    const int32_t* sfa;
    const int32_t* sfb;
    void* d;
    const int32_t* grouped_layout;
    int m, n, k, groups;
    int a_rows;  // total rows of the flattened A
    int64_t lda, ldb, ldd;
    bool x_mn = false, w_mn = false;  // operand is MN-major (M / N contiguous, K strided by lda / ldb)
    const uint32_t* arrival = nullptr;           // EP dispatch in flight: per-group arrival counters / their targets;
    const uint32_t* arrival_expected = nullptr;  // the launch overlaps the producer kernel (no griddepcontrol.wait)
    int sfa_krows = 0, sfb_krows = 0; // k-grouped: total packed SF rows (0: derive from k)
    int sfa_stride, sfb_stride, sfa_cols, sfb_cols;
    int gran_k_a, gran_k_b;
    int d_dtype, accumulate;
    int expected_m, alignment, zero_padding;
    void* workspace;
    size_t workspace_bytes;
    cudaStream_t stream;
    int64_t batch_stride_a = 0, batch_stride_b = 0, batch_stride_d = 0;   // batched (elements)
    int head_left = 0, head_mid = 0, head_right = 0;                     // fp8_gemm_nt_skip_head_mid
    bool swap_d = false;   // operands already exchanged by the caller (a = weights, b = tokens): the kernel writes D[lane][column]
    int forced_block_m = 0;
    bool bf16_ab = false;   // BF16 operands without scale factors: k, lda, ldb are in BYTES (2 x elements), sfa / sfb unused
};

struct Maps {
    CUtensorMap x, w, sfx, sfw, d;
};

template <typename Kernel>
int launch_kernel(Kernel kernel, const Config& cfg, cudaStream_t stream, const Maps& maps, const GemmParams& p) {
    // Opt in to > 48 KB dynamic smem once per instantiation and device
    // (all instantiations share one function-pointer type, so the memo is keyed by the kernel address)
    static std::mutex mu;
    static std::unordered_map<const void*, int> configured;  // kernel -> device it was configured on
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = configured.find(reinterpret_cast<const void*>(kernel));
        if (it == configured.end() || it->second != rt_device()) {
            DGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCapacity));
            configured[reinterpret_cast<const void*>(kernel)] = rt_device();
        }
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim = cfg.grid > 0 ? dim3(cfg.grid, cfg.grid_y, 1) : dim3(cfg.num_sms / cfg.cluster * cfg.cluster, 1, 1);
    lc.blockDim = dim3(kNumThreads, 1, 1);
    lc.dynamicSmemBytes = cfg.smem_bytes;
    lc.stream = stream;
    cudaLaunchAttribute attrs[3];
    int na = 0;
    if (cfg.cluster > 1) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = cfg.cluster;
        attrs[na].val.clusterDim.y = 1;
        attrs[na].val.clusterDim.z = 1;
        ++na;
    }
    if (cfg.cluster > 2 && cfg.grid == 0) {
        // 4/8-CTA clusters must sit inside one GPC: ask how many fit at once and size the persistent grid to that
        static std::mutex occ_mu;
        static std::unordered_map<const void*, int> resident;   // kernel (x cluster size, implied) -> clusters
        std::lock_guard<std::mutex> lock(occ_mu);
        auto it = resident.find(reinterpret_cast<const void*>(kernel));
        if (it == resident.end()) {
            lc.attrs = attrs, lc.numAttrs = na;
            int n = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kernel, &lc);
            if (e != cudaSuccess || n <= 0) n = cfg.num_sms / cfg.cluster;
            it = resident.emplace(reinterpret_cast<const void*>(kernel), n).first;
        }
        const int clusters = std::min(it->second, cfg.num_sms / cfg.cluster);
        lc.gridDim = dim3(clusters * cfg.cluster, 1, 1);
    }
    if (rt_pdl() || cfg.overlap_producer) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    lc.attrs = attrs;
    lc.numAttrs = na;
    DGB_CUDA(cudaLaunchKernelEx(&lc, kernel, maps.x, maps.w, maps.sfx, maps.sfw, maps.d, p));
    count_launches(1);
    return DGB200_OK;
}

// ---- one per kernel-instance translation unit
int dispatch_dense_kk(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p);      // gemm_dense_kk.cu
int dispatch_dense_mn(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p);      // gemm_dense_mn.cu
int dispatch_dense_splitk(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p);  // gemm_dense_splitk.cu
int dispatch_grouped(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p);       // gemm_grouped.cu
int dispatch_batched(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p);       // gemm_batched.cu
int dispatch_dense_swap(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p);    // gemm_dense_swap.cu
int dispatch_bf16(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p);          // gemm_bf16.cu
int dispatch_bf16_mn(const GemmCall& c, const Config& cfg, const Maps& maps, const GemmParams& p);       // gemm_bf16_mn.cu

}  // namespace dgb200
