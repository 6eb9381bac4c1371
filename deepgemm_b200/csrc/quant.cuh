// Activation quantiser: BF16 rows -> FP8 E4M3 rows + packed UE8M0 scale factors, already in the MN-major wire format
// the GEMM consumes (int32 [M, ceil(K / (4 gran_k))], strides (1, align4(M))). One pass over the data, HBM bound.
//
// Replaces, for the step in front of the GEMM in inference, the reference's test-side
//   per_token_cast_to_fp8(x, use_ue8m0=True, gran_k, use_packed_ue8m0=True)      (deep_gemm/utils/math.py:26-38)
// followed by the layout transform of csrc/apis/layout.hpp:48-58 (transpose + pack kernels, impls/smxx_layout.cuh).
// Arithmetic is restated operation by operation so that the bytes are identical:
//   amax = max |x| over the 1 x gran_k block (FP32), clamped to 1e-4            math.py:33
//   sf   = amax / 448, rounded UP to a power of two, exponent in [1, 254]        math.py:13-16, :34-35
//   q    = e4m3_rn(float(x) * (1 / sf))                                          math.py:36
//   byte j of word w of row r = exponent of block 4w + j (0 past the end of K)   math.py:19-23, tests/test_layout.py:20-42
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp8.h>

#include <cstdint>

namespace dgb200 {

// One warp per (row, 512-element chunk): every lane owns 16 consecutive elements (32 B in, 16 B out). A scale block is
// gran_k / 16 consecutive lanes; a packed word covers 4 blocks. grid = (ceil(K / 512), ceil(M / warps per CTA)).
template <uint32_t kGranK>
__global__ void __launch_bounds__(256)
per_token_cast_to_fp8_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, uint8_t* __restrict__ q, int64_t ldq,
                             uint32_t* __restrict__ sf, uint32_t sf_stride, uint32_t m, uint32_t k) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    constexpr uint32_t kLanesPerBlock = kGranK / 16;            // 8 (gran 128) or 2 (gran 32)
    constexpr uint32_t kWordsPerChunk = 512 / (4 * kGranK);     // 1 or 4
    const uint32_t lane = threadIdx.x % 32;
    const uint32_t row = blockIdx.y * (blockDim.x / 32) + threadIdx.x / 32;
    if (row >= m) return;
    const uint32_t k0 = blockIdx.x * 512 + lane * 16;
    const __nv_bfloat16* src = x + static_cast<int64_t>(row) * ldx + k0;

    float v[16];
    const bool vec_in = ((reinterpret_cast<uintptr_t>(x) | (static_cast<uint64_t>(ldx) * 2)) & 15) == 0;
    if (vec_in && k0 + 16 <= k) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(src)), b = __ldg(reinterpret_cast<const uint4*>(src) + 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);              // BF16 -> FP32 is a shift
            v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    } else {
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i) v[i] = k0 + i < k ? __bfloat162float(src[i]) : 0.0f;   // K tail: zeros (math.py:30-31)
    }
    float amax = 0.0f;
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i) amax = fmaxf(amax, fabsf(v[i]));
#pragma unroll
    for (uint32_t d = 1; d < kLanesPerBlock; d *= 2) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
    const float sfv = __fdiv_rn(fmaxf(amax, 1e-4f), 448.0f);
    const uint32_t bits = __float_as_uint(sfv);
    uint32_t e = ((bits >> 23) & 0xFFu) + ((bits & 0x7FFFFFu) != 0 ? 1u : 0u);
    e = min(max(e, 1u), 254u);
    const float inv = __fdiv_rn(1.0f, __uint_as_float(e << 23));

    uint32_t out[4];
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        uint32_t word = 0;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            word |= static_cast<uint32_t>(__nv_cvt_float_to_fp8(v[4 * i + j] * inv, __NV_SATFINITE, __NV_E4M3)) << (8 * j);
        out[i] = word;
    }
    uint8_t* dst = q + static_cast<int64_t>(row) * ldq + k0;
    const bool vec_out = ((reinterpret_cast<uintptr_t>(q) | static_cast<uint64_t>(ldq)) & 15) == 0;
    if (vec_out && k0 + 16 <= k) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i)
            if (k0 + i < k) dst[i] = static_cast<uint8_t>(out[i / 4] >> (8 * (i % 4)));
    }

    // scale bytes: block b of this chunk lives in lane b * kLanesPerBlock; word w = blocks 4w .. 4w + 3
    const uint32_t block_in_chunk = lane / kLanesPerBlock;
    const bool block_valid = blockIdx.x * 512 + block_in_chunk * kGranK < k;       // blocks past K pack as 0
    uint32_t byte = block_valid ? e : 0u;
    byte <<= 8 * (block_in_chunk % 4);
#pragma unroll
    for (uint32_t d = kLanesPerBlock; d < 4 * kLanesPerBlock; d *= 2) byte |= __shfl_xor_sync(0xffffffffu, byte, d);
    if (lane % (4 * kLanesPerBlock) == 0) {
        const uint32_t word_idx = blockIdx.x * kWordsPerChunk + lane / (4 * kLanesPerBlock);
        const uint32_t num_words = (k + 4 * kGranK - 1) / (4 * kGranK);
        if (word_idx < num_words) sf[static_cast<uint64_t>(word_idx) * sf_stride + row] = byte;
    }
}

}  // namespace dgb200
