// Scale-factor layout kernels (plain CUDA, HBM-trivial: a few hundred KB per call).
//
// Replaces the reference's `transpose_and_pack_fp32_into_ue8m0`, `pack_fp32_into_ue8m0` and `transpose_fp32`
// (deep_gemm/include/deep_gemm/impls/smxx_layout.cuh:12-246). One kernel covers every input stride pattern,
// fuses the "one 128x128-block scale -> 128 per-row scales" broadcast the reference does with a separate
// `index_select` (csrc/apis/layout.hpp:51-52), and takes all shapes at run time.
#pragma once
#include <cstdint>

namespace dgb200 {

// out[g][kp][mn] (int32, mn contiguous, row stride `aligned_mn`) : byte j = exponent of sf[g][mn / gran_mn][4 kp + j]
//   grid (ceil(aligned_mn / 128), num_kp, num_groups), block 128: one output word per thread, coalesced writes.
template <bool kPsum>
__global__ void __launch_bounds__(128)
pack_sf_ue8m0_kernel(const float* __restrict__ sf, uint32_t* __restrict__ out, uint32_t mn, uint32_t aligned_mn,
                     uint32_t sf_k, uint32_t num_kp, uint32_t gran_mn, int64_t stride_g, int64_t stride_mn,
                     int64_t stride_k, const int32_t* __restrict__ psum_layout, uint32_t num_psum_groups,
                     uint32_t m_alignment) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t r = blockIdx.x * 128 + threadIdx.x;
    const uint32_t kp = blockIdx.y, g = blockIdx.z;
    if (r >= aligned_mn) return;
    bool valid = r < mn;
    if constexpr (kPsum) {
        // rows inside [align(end_{i-1}), end_i) are real; alignment gaps hold uninitialised data (written as 0)
        if (valid) {
            uint32_t lo = 0, hi = num_psum_groups;  // first group whose end is > r
            while (lo < hi) {
                const uint32_t mid = (lo + hi) / 2;
                if (static_cast<uint32_t>(__ldg(psum_layout + mid)) > r) hi = mid; else lo = mid + 1;
            }
            if (lo >= num_psum_groups) {
                valid = false;
            } else {
                const uint32_t prev_end = lo == 0 ? 0u : static_cast<uint32_t>(__ldg(psum_layout + lo - 1));
                const uint32_t start = (prev_end + m_alignment - 1) / m_alignment * m_alignment;
                valid = r >= start;
            }
        }
    }
    uint32_t packed = 0;
    if (valid) {
        const float* row = sf + g * stride_g + static_cast<int64_t>(r / gran_mn) * stride_mn;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t kk = kp * 4 + j;
            if (kk < sf_k) {
                const uint32_t bits = __float_as_uint(__ldg(row + kk * stride_k));
                // UE8M0 keeps the exponent only: sign and mantissa must be zero (smxx_layout.cuh:131)
                if (bits & 0x807fffffu) {
                    printf("dgb200: scale factor %f is not a power of two (use_ue8m0=True required)\n", __uint_as_float(bits));
                    asm volatile("trap;");
                }
                packed |= (bits >> 23) << (8 * j);
            }
        }
    }
    out[(static_cast<size_t>(g) * num_kp + kp) * aligned_mn + r] = packed;
}

// out[g][k][mn] = sf[g][mn][k]  (fp32, MN-major, TMA aligned); padding columns are left untouched like the reference.
__global__ void __launch_bounds__(128)
transpose_sf_fp32_kernel(const float* __restrict__ sf, float* __restrict__ out, uint32_t mn, uint32_t aligned_mn,
                         uint32_t sf_k, int64_t stride_g, int64_t stride_mn, int64_t stride_k) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t r = blockIdx.x * 128 + threadIdx.x;
    const uint32_t k = blockIdx.y, g = blockIdx.z;
    if (r >= mn) return;
    out[(static_cast<size_t>(g) * sf_k + k) * aligned_mn + r] = __ldg(sf + g * stride_g + r * stride_mn + k * stride_k);
}

// K-grouped: input rows are K-granules of all groups back to back ([sum ceil(k_g/gran_k), mn], mn contiguous);
// every group is padded to a multiple of 4 granules on its own (smxx_layout.hpp:255-316).
//   grid (ceil(mn/128), total packed rows), block 128. `ks` (device) holds per-group K; the block finds its group
//   by a short uniform scan, so nothing but the grid size depends on host-side knowledge of `ks`.
__global__ void __launch_bounds__(128)
pack_sf_ue8m0_k_grouped_kernel(const float* __restrict__ sf, uint32_t* __restrict__ out, uint32_t mn,
                               const int32_t* __restrict__ ks, uint32_t num_groups, uint32_t gran_k,
                               uint32_t psum_alignment /* 0: ks[g] = K of group g; else ks[g] = end K, starts aligned */) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t c = blockIdx.x * 128 + threadIdx.x;
    const uint32_t pr = blockIdx.y;
    uint32_t in_row = 0, packed_row = 0, first = 0, count = 0, prev_end = 0;
    for (uint32_t g = 0; g < num_groups; ++g) {
        const uint32_t v = static_cast<uint32_t>(max(0, __ldg(ks + g)));
        uint32_t kg = v;
        if (psum_alignment) {
            const uint32_t start = (prev_end + psum_alignment - 1) / psum_alignment * psum_alignment;
            kg = v > start ? v - start : 0u;
            prev_end = max(start, v);
        }
        const uint32_t n_in = (kg + gran_k - 1) / gran_k, n_packed = (n_in + 3) / 4;
        if (pr < packed_row + n_packed) {
            first = in_row + (pr - packed_row) * 4;
            count = min(4u, n_in - (pr - packed_row) * 4);
            break;
        }
        in_row += n_in, packed_row += n_packed;
    }
    if (c >= mn) return;
    uint32_t packed = 0;
    for (uint32_t j = 0; j < count; ++j) {
        const uint32_t bits = __float_as_uint(__ldg(sf + static_cast<size_t>(first + j) * mn + c));
        if (bits & 0x807fffffu) asm volatile("trap;");
        packed |= (bits >> 23) << (8 * j);
    }
    out[static_cast<size_t>(pr) * mn + c] = packed;   // rows past the last group (upper-bound allocations) get 0
}

}  // namespace dgb200
