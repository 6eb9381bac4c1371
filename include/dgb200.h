/*
 * dgb200 -- C ABI of the B200-native FP8 blockwise-scaled GEMM library.
 *
 * This is the drop-in boundary for the FP8 GEMM hot path of deepseek-ai/DeepGEMM. The reference has no C ABI:
 * its boundary is the pybind11 module `deep_gemm._C` (csrc/python_api.cpp:17-28) whose functions take
 * torch.Tensors. Every entry point below is the torch-free core of one of those functions; the comment on each
 * names the reference function it replaces (file:line in the reference tree). `deepgemm_b200/` binds them with
 * ctypes; INTEGRATION.md shows the pybind/ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers unless the name ends in `_host`; sizes are element counts
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is enqueued, nothing syncs
 *   - A ("tokens", M side) and B ("weights", N side) are FP8 E4M3; D is BF16 (DGB200_BF16) or FP32 (DGB200_FP32)
 *   - scale factors are the reference's SM100 wire format: UE8M0 bytes, 4 consecutive K-granules packed in one
 *     int32, MN-major (element (mn, kp) at  sf[kp * sf_stride + mn],  sf_stride >= align(mn, 4);
 *     batched/grouped: group g starts at  g * num_kp * sf_stride)   (csrc/utils/layout.hpp:100-107)
 *   - return value: 0 on success, otherwise a DGB200_ERR_* code; dgb200_last_error() gives the message
 *     (the reference throws DGException -> Python RuntimeError, csrc/utils/exception.hpp:12-40)
 *   - no function reads device memory on the host: every call is CUDA-graph capturable
 */
#ifndef DGB200_H_
#define DGB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGB200_VERSION 100

enum { DGB200_BF16 = 0, DGB200_FP32 = 1 };
enum { DGB200_K_MAJOR = 0, DGB200_MN_MAJOR = 1 };
enum {
    DGB200_OK = 0,
    DGB200_ERR_INVALID_ARGUMENT = 1, /* host-side contract violation (DG_HOST_ASSERT in the reference) */
    DGB200_ERR_CUDA = 2,             /* CUDA runtime / driver failure */
    DGB200_ERR_UNSUPPORTED = 3       /* valid in the reference, not built here (FP4 operands, SM90-only paths) */
};

/* Message of the last failing call on this thread. */
const char* dgb200_last_error(void);
int dgb200_version(void);

/* ---- runtime knobs: csrc/apis/runtime.hpp:11-49, csrc/apis/layout.hpp:142-150 --------------------------- */
int dgb200_set_num_sms(int num_sms);      /* 0 <= n <= device SM count, 0 = all (jit/device_runtime.hpp:103-105); odd
                                             budgets are accepted and rounded down to whole CTA pairs at launch */
int dgb200_get_num_sms(void);             /* the value set; 0 -> all SMs of the device once one has been used */
int dgb200_set_tc_util(int percent);      /* accepted for API parity; only the reference's BF16 kernel consumes it */
int dgb200_get_tc_util(void);
int dgb200_set_pdl(int enabled);          /* programmatic dependent launch attribute on every kernel launch */
int dgb200_get_pdl(void);
/* Split-K (no reference equivalent). Small dense problems may cut K into slices -- across the CTAs of a cluster
 * (partials exchanged through distributed shared memory) or, with a workspace, across the grid -- so that all SMs
 * stream weights. The slices are added in a fixed order (run-to-run deterministic), but not in the reference's one
 * pass over K: results then agree with the reference to FP32 rounding instead of bit for bit. allow = 0 turns every
 * form of split-K off (default 1). */
int dgb200_set_split_k(int allow);
int dgb200_get_split_k(void);
int dgb200_set_mk_alignment_for_contiguous_layout(int alignment);
int dgb200_get_mk_alignment_for_contiguous_layout(void);
int dgb200_get_theoretical_mk_alignment_for_contiguous_layout(int expected_m /* <=0: none */);
int dgb200_get_tma_aligned_size(int x, int element_size);  /* csrc/utils/math.hpp:23-27 */

/* ---- scale-factor layout transforms: csrc/jit_kernels/impls/smxx_layout.hpp:120-316 ------------------------ */

/* FP32 power-of-two scale factors -> packed UE8M0 int32, MN-major, TMA aligned.
 * Replaces get_mn_major_tma_aligned_packed_ue8m0_tensor (smxx_layout.hpp:180-253) fused with the
 * `index_select` row broadcast of csrc/apis/layout.hpp:48-54 (gran_mn = 128 means one input row per 128 outputs).
 *   sf        : fp32 [num_groups, ceil(mn/gran_mn), sf_k] with element strides (stride_g, stride_mn, stride_k)
 *   out       : int32, num_groups * ceil(sf_k/4) * align(mn,4) words
 *   psum_layout: optional int32[num_psum_groups] end rows; rows in alignment gaps are written as 0 (safe, finite)
 * Scale factors that are not exact powers of two trap on the device, as in the reference (smxx_layout.cuh:131). */
int dgb200_pack_sf_ue8m0(const float* sf, int32_t* out, int mn, int sf_k, int num_groups, int gran_mn,
                         int64_t stride_g, int64_t stride_mn, int64_t stride_k,
                         const int32_t* psum_layout, int num_psum_groups, int m_alignment, void* stream);

/* FP32 scale factors -> FP32 MN-major TMA-aligned copy (get_mn_major_tma_aligned_tensor, smxx_layout.hpp:120-178).
 *   out: fp32, num_groups * sf_k * align(mn,4) elements, element (g, mn, k) at out[(g*sf_k + k)*align(mn,4) + mn] */
int dgb200_transpose_sf_fp32(const float* sf, float* out, int mn, int sf_k, int num_groups,
                             int64_t stride_g, int64_t stride_mn, int64_t stride_k, void* stream);

/* K-grouped variant (get_k_grouped_mn_major_tma_aligned_packed_ue8m0_tensor, smxx_layout.hpp:255-316):
 *   sf  : fp32 [sum_g ceil(k_g/gran_k), mn] contiguous;  out: int32 [packed_rows, mn] contiguous, where every group
 *   is padded to a multiple of 4 granules on its own (packed_rows = sum_g ceil(k_g/(4*gran_k)), computed by the caller
 *   from its host copy of the per-group K, the reference's `ks_cpu`, or an upper bound: extra rows are zero-filled).
 *   `ks_device`: int32[num_groups]: K of each group (psum_alignment == 0) or end K of each group, whose start is the
 *   previous end rounded up to `psum_alignment` (psum layout). */
int dgb200_pack_sf_ue8m0_k_grouped(const float* sf, int32_t* out, int mn, const int32_t* ks_device, int num_groups,
                                   int packed_rows, int gran_k, int psum_alignment, void* stream);

/* ---- GEMMs ------------------------------------------------------------------------------------------------ */

/* D[m,n] (= C +) sum_k A[m,k] B[n,k]          -- fp8_fp4_gemm_nt, csrc/apis/gemm.hpp:73-124 (+ nn/tn/tt :126-164
 * through the major flags: K_MAJOR means the K extent is contiguous, lda/ldb are the strides of the other extent).
 * accumulate != 0: D holds C on entry (the host wrapper copies C into D first, gemm.hpp:42-44). */
int dgb200_fp8_gemm_nt(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb, void* d,
                       int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd,
                       int major_a, int major_b, int sfa_stride, int sfb_stride, int gran_k_a, int gran_k_b,
                       int d_dtype, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);
/* `workspace` (optional, may be NULL): device scratch that lets small problems (fewer output tiles than SM pairs) cut K
 * into slices so that every SM streams its own part of B ("split-K"). Contract: 16-byte aligned, its first
 * DGB200_WORKSPACE_HEADER_BYTES are zero before the first use (the kernel leaves them zero again), and it is not
 * shared by GEMMs that may run concurrently (use one per stream). Results stay deterministic: partial sums are added
 * in slice order. dgb200_workspace_bytes() gives a size that is always sufficient for an (m, n) problem. */
#define DGB200_WORKSPACE_HEADER_BYTES 16384
int64_t dgb200_workspace_bytes(int m, int n);

/* Dense GEMM whose output skips a gap in every head  -- fp8_gemm_nt_skip_head_mid, csrc/apis/attention.hpp:19-74
 * (epilogue remap: deep_gemm/include/deep_gemm/epilogue/transform.cuh:15-22).
 *   a [m, k], b [n, k] both K-major, gran_k 128; n is a multiple of (head_left + head_right)
 *   d [m, n + n / (head_left + head_right) * head_mid]: GEMM column j lands at  j + (j + head_right) / (head_left +
 *   head_right) * head_mid, i.e. each head is written as [left | (mid columns left untouched) | right]. */
int dgb200_fp8_gemm_nt_skip_head_mid(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb, void* d,
                                     int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd,
                                     int head_left, int head_mid, int head_right,
                                     int sfa_stride, int sfb_stride, int d_dtype, void* stream);

/* Batched GEMM D[i] (= C[i] +) A[i] B[i]^T     -- fp8_bmm / fp8_einsum, csrc/apis/einsum.hpp:137-214 (launcher
 * csrc/jit_kernels/impls/sm100_fp8_fp4_gemm_1d1d.hpp:393-467).
 *   a: batch i at a + i * batch_stride_a; K_MAJOR: rows of k contiguous bytes, pitch lda; MN_MAJOR: m contiguous, k pitch lda
 *   b: likewise with n;  d: batch i at d + i * batch_stride_d (elements), row pitch ldd, columns contiguous
 *   Arbitrary (16-byte multiple) pitches, so permuted views such as the "bhr,hdr->bhd" operands need no copy.
 *   sfa / sfb: packed UE8M0 [batch, mn, ceil(k / (4 gran_k))], MN-major, batch i at i * num_kp * sf_stride words.
 *   accumulate != 0: D holds C on entry. */
int dgb200_fp8_bmm(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb, void* d,
                   int batch, int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd,
                   int64_t batch_stride_a, int64_t batch_stride_b, int64_t batch_stride_d,
                   int major_a, int major_b, int sfa_stride, int sfb_stride, int gran_k_a, int gran_k_b,
                   int d_dtype, int accumulate, void* stream);

/* Activation quantiser (the step in front of the GEMM in inference)  -- per_token_cast_to_fp8(x, use_ue8m0=True,
 * gran_k, use_packed_ue8m0=True), deep_gemm/utils/math.py:26-38, fused with the MN-major packing of
 * csrc/apis/layout.hpp:48-58: x [m, k] BF16 (row pitch ldx elements) -> q [m, k] E4M3 (row pitch ldq bytes) and
 * sf int32 [m, ceil(k / (4 gran_k))] MN-major (word (r, w) at sf[w * sf_stride + r], sf_stride >= align(m, 4)),
 * ready to be passed to the GEMMs above. Bit-identical to the reference's Python. gran_k: 32 or 128. */
int dgb200_per_token_cast_to_fp8(const void* x, int64_t ldx, void* q, int64_t ldq, int32_t* sf, int sf_stride,
                                 int m, int k, int gran_k, void* stream);

/* BF16 x BF16 GEMMs without scale factors on the same kernel skeleton (tcgen05.mma kind::f16):
 *   dgb200_bf16_gemm_nt                       -- bf16_gemm_{nt,nn,tn,tt}, csrc/apis/gemm.hpp:404-462 (majors as in dgb200_fp8_gemm_nt)
 *   dgb200_m_grouped_bf16_gemm_nt_contiguous  -- m_grouped_bf16_gemm_{nt,nn}_contiguous, gemm.hpp:464-526 (layouts as the FP8 form)
 *   dgb200_m_grouped_bf16_gemm_nt_masked      -- m_grouped_bf16_gemm_nt_masked, gemm.hpp:528-564
 *   dgb200_k_grouped_bf16_gemm_tn_contiguous  -- k_grouped_bf16_gemm_tn_contiguous, gemm.hpp:566-608
 * K-major: a [m, k], b [n, k] (grouped: [G, n, k]); MN-major: a [k, m], b [k, n] (grouped: [G, k, n]); lda / ldb = pitch of
 * the strided extent in ELEMENTS, contiguous extents and pitches multiples of 8 elements (16-byte rows for TMA).
 * k-grouped: a [sum_k, m], b [sum_k, n] BF16 contiguous, d [num_groups, m, n] FP32 holding C on entry (accumulated into),
 * grouped_layout int32[num_groups] = K of each group (psum: unaligned end K, group starts aligned to mk_alignment). */
int dgb200_bf16_gemm_nt(const void* a, const void* b, void* d, int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd,
                        int major_a, int major_b, int d_dtype, int accumulate, void* stream);
int dgb200_m_grouped_bf16_gemm_nt_contiguous(const void* a, const void* b, void* d, const int32_t* grouped_layout,
                                             int num_groups, int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd,
                                             int major_b, int use_psum_layout, int ensure_zero_padding,
                                             int expected_m_for_psum_layout, void* stream);
int dgb200_m_grouped_bf16_gemm_nt_masked(const void* a, const void* b, void* d, const int32_t* masked_m, int num_groups,
                                         int m_max, int n, int k, int expected_m, void* stream);
int dgb200_k_grouped_bf16_gemm_tn_contiguous(const void* a, const void* b, float* d, const int32_t* grouped_layout,
                                             int num_groups, int m, int n, int sum_k, int use_psum_layout, void* stream);
/* Batched BF16 GEMM D[i] = A[i] B[i]^T behind einsum('bhr,hdr->bhd' / 'bhd,hdr->bhr'), csrc/apis/einsum.hpp:62-108: a K-major,
 * b K-major [batch, n, k] or MN-major [batch, k, n], d BF16; pitches and batch strides in elements (as dgb200_fp8_bmm). */
/* Batch-reduction GEMM d[m, n] += sum_i a[i] b[i]^T behind einsum('bmk,bnk->mn'), csrc/apis/einsum.hpp:22-60
 * (sm100_bmn_bnk_mn_gemm): a [batch, m, k], b [batch, n, k] BF16 contiguous, d FP32 [m, n] contiguous, accumulated in place;
 * k % 64 == 0. The batches are cut into chunks over the grid, every chunk adds its partial tile with memory-side FP32 adds
 * (the order of those adds is not fixed: results agree to FP32 rounding from run to run, as the reference's do). */
int dgb200_bf16_bmk_bnk_mn(const void* a, const void* b, float* d, int batch, int m, int n, int k, void* stream);
int dgb200_bf16_bmm(const void* a, const void* b, void* d, int batch, int m, int n, int k, int64_t lda, int64_t ldb,
                    int64_t ldd, int64_t batch_stride_a, int64_t batch_stride_b, int64_t batch_stride_d, int major_b,
                    void* stream);

/* Rows of A grouped by expert          -- m_grouped_fp8_fp4_gemm_nt_contiguous, csrc/apis/gemm.hpp:166-232.
 *   a [m, k], b [num_groups, n, k], d [m, n] bf16
 *   use_psum_layout == 0: grouped_layout int32[m], expert id per row, -1 for padding rows
 *   use_psum_layout != 0: grouped_layout int32[num_groups], (unaligned) end row of each expert; expert g+1 starts
 *                         at align(end_g, mk_alignment); ensure_zero_padding writes zeros to the gap rows of D */
int dgb200_m_grouped_fp8_gemm_nt_contiguous(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb,
                                            void* d, const int32_t* grouped_layout, int num_groups,
                                            int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd, int major_b,
                                            int sfa_stride, int sfb_stride, int gran_k_a, int gran_k_b,
                                            int use_psum_layout, int ensure_zero_padding,
                                            int expected_m_for_psum_layout, void* stream);

/* Per-expert fixed slots, device-side counts   -- m_grouped_fp8_fp4_gemm_nt_masked, csrc/apis/gemm.hpp:250-297.
 *   a [num_groups, m_max, k], b [num_groups, n, k], d [num_groups, m_max, n] bf16, masked_m int32[num_groups]
 *   (device; never read by the host). Only rows < masked_m[g] of each group are written. */
int dgb200_m_grouped_fp8_gemm_nt_masked(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb,
                                        void* d, const int32_t* masked_m, int num_groups,
                                        int m_max, int n, int k, int expected_m,
                                        int sfa_stride, int sfb_stride, int gran_k_a, int gran_k_b, void* stream);

/* Weight gradient, K grouped           -- k_grouped_fp8_gemm_tn_contiguous, csrc/apis/gemm.hpp:299-346.
 *   a [sum_k, m], b [sum_k, n] e4m3 (both MN-major: m / n contiguous), d [num_groups, m, n] fp32 accumulated IN PLACE
 *   (D holds C on entry), grouped_layout int32[num_groups] on the device: K of each group (use_psum_layout == 0) or
 *   the end K of each group with group starts aligned to the mk alignment (use_psum_layout != 0). Per-group K must be
 *   a multiple of 32. sfa / sfb: k-grouped packed UE8M0 [sf_rows, m] / [sf_rows, n] (dgb200_pack_sf_ue8m0_k_grouped),
 *   gran_k 32 or 128 for both. Groups with K == 0 leave D untouched. */
int dgb200_k_grouped_fp8_gemm_tn_contiguous(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb,
                                            float* d, const int32_t* grouped_layout, int num_groups, int m, int n,
                                            int sum_k, int sf_rows, int gran_k, int use_psum_layout, void* stream);

/* ---- expert-parallel dispatch over NVLink peer memory ----------------------------------------------------------
 * The m-grouped contiguous GEMM sharded by experts (rank r of `world` owns experts [r*G/world, (r+1)*G/world)); the
 * reference pairs its grouped GEMM with an external dispatch library for this (tests/test_mega_moe.py:148-205).
 * Every rank owns one "dispatch buffer" of dgb200_ep_buffer_bytes() bytes, identical layout on all ranks, mapped into
 * every peer (dgb200_ep_export / _import use CUDA IPC; any other peer-mapping mechanism works as long as `buffers[p]`
 * is rank p's buffer as addressable from the calling device). dgb200_ep_dispatch() enqueues one persistent kernel that writes
 * the local tokens straight into the owners' buffers in the contiguous-grouped psum layout and return once all rows
 * destined to THIS rank have landed (in stream order) -- the GEMM then reads, inside the local buffer,
 *   a   = base + offsets[DGB200_EP_OFF_A]    uint8/e4m3 [capacity, k]
 *   sfa = base + offsets[DGB200_EP_OFF_SFA]  int32 [ceil(k/512), capacity] (MN-major, sfa_stride = capacity)
 *   psum layout = base + offsets[DGB200_EP_OFF_PSUM] int32 [G/world], m = capacity.
 * No host synchronisation, CUDA-graph capturable; all ranks must call it the same number of times. */
enum { DGB200_EP_OFF_A = 0, DGB200_EP_OFF_SFA = 1, DGB200_EP_OFF_PSUM = 2, DGB200_EP_OFF_COUNTS = 3,
       DGB200_EP_OFF_NUM_ROWS = 4, DGB200_EP_OFF_OVERFLOW = 5, DGB200_EP_NUM_OFFSETS = 6 };
int64_t dgb200_ep_buffer_bytes(int world, int num_experts, int capacity, int k);
int dgb200_ep_buffer_offsets(int world, int num_experts, int capacity, int k, int64_t* offsets /* [DGB200_EP_NUM_OFFSETS] */);
int dgb200_ep_alloc(int64_t bytes, void** ptr);                 /* cudaMalloc + zero fill: IPC exportable */
int dgb200_ep_free(void* ptr);
int dgb200_ep_export(void* ptr, void* handle_64_bytes);         /* cudaIpcGetMemHandle */
int dgb200_ep_import(const void* handle_64_bytes, void** ptr);  /* cudaIpcOpenMemHandle (enables peer access) */
int dgb200_ep_unimport(void* ptr);
/* x [num_tokens, k] e4m3 rows (pitch ldx bytes), sf: packed UE8M0 words of token t at sf[t*sf_stride_t + j*sf_stride_k],
 * j < ceil(k/512). Routing: expert_ids [num_tokens, topk] (row-major; int32 = id_bytes 4, int64 = id_bytes 8), values
 * outside [0, num_experts) = slot not routed; token t is copied once per routed slot ("entry" t * topk + j).
 * token_row int32[num_tokens * topk] (out): row of each entry inside its owner's buffer (-1: not routed / dropped on
 * overflow, which also sets the word at DGB200_EP_OFF_OVERFLOW). k % 16 == 0, ceil(k/512) <= 32, capacity % 4 == 0.
 * wait_for_all != 0 (the default path): ONE persistent kernel -- rank entries (O(T)), exchange counts with the peers,
 * scatter rows over NVLink, signal and wait for every source -- after which any consumer may follow in stream order.
 * wait_for_all == 0 (top-1 only): the count / exchange / order / scatter kernel chain without the final wait, for a
 * consumer that watches the per-expert arrival counters itself (dgb200_ep_grouped_gemm with overlap_dispatch); tokens
 * are then sent in expert order, so experts complete one by one. `order_scratch` is only used by that mode. */
int dgb200_ep_dispatch(const void* x, int64_t ldx, const int32_t* sf, int64_t sf_stride_t, int64_t sf_stride_k,
                       const void* expert_ids, int id_bytes, int num_tokens, int topk, int k, int num_experts, int rank,
                       int world, void* const* buffers, int capacity, int alignment, int32_t* token_row,
                       int32_t* order_scratch /* int32[num_tokens], device; may be NULL with wait_for_all */,
                       int wait_for_all, void* stream);
/* The grouped GEMM of this rank's experts over its dispatch buffer (psum layout, zero padding, BF16 D [capacity, n];
 * b [G/world, n, k] e4m3, sfb packed UE8M0). With overlap_dispatch != 0 it must directly follow a dgb200_ep_dispatch(...,
 * wait_for_all = 0, ...) on the same stream: the kernel is then launched as a programmatic dependent of the scatter
 * kernel, runs beside it, and its TMA producer waits expert by expert on the arrival counters the sources increment
 * (remote atomics) -- the transfer of expert g+1.. overlaps the math of expert g. With overlap_dispatch == 0 it is
 * dgb200_m_grouped_fp8_gemm_nt_contiguous on the buffer (use after wait_for_all = 1). */
int dgb200_ep_grouped_gemm(void* local_buffer, int world, int num_experts, int capacity, int k, const void* b,
                           const int32_t* sfb, void* d, int n, int64_t ldb, int64_t ldd, int major_b, int sfb_stride,
                           int gran_k_b, int expected_m, int overlap_dispatch, void* stream);

/* The way back (weighted top-k combine, the last step of the reference's baseline MoE layer, tests/test_mega_moe.py:196-202):
 *   out[t, 0:n] = sum_j weights[t, j] * D_owner(t, j)[token_row[t * topk + j], 0:n]      (slots with token_row < 0 skipped)
 * products and running sum in FP32 in slot order (separate multiply and add), rounded once to BF16; a token without any
 * routed slot gets zeros. weights == NULL means 1.0; with topk == 1 that is a pure gather (bit-exact copy, elt_bytes 2 or 4).
 * `d_buffers[p]` = rank p's grouped-GEMM output [capacity, n] (pitch ldd elements) as addressable from this device (peer
 * mapped, e.g. dgb200_ep_alloc + _export/_import); `buffers` = the dispatch buffers (their control blocks carry the
 * handshake). Enqueue on the stream that ran this rank's grouped GEMM: a first kernel tells every peer that this rank's D
 * is complete, the gather waits for every owner's message, then pulls the rows over NVLink. All ranks must call it once
 * per dispatch. D may be overwritten again after the next dgb200_ep_dispatch has returned control to the stream. */
int dgb200_ep_combine(void* out, int64_t ldo, const int32_t* token_row, const void* expert_ids, int id_bytes, int num_tokens,
                      int topk, const float* weights, int n, int elt_bytes, int num_experts, int rank, int world,
                      void* const* buffers, void* const* d_buffers, int64_t ldd, void* stream);

/* ---- introspection (bench / tests) -------------------------------------------------------------------------- */
typedef struct dgb200_config {
    int block_m;     /* token rows per tile */
    int cluster;     /* CTAs per MMA (1 or 2) */
    int num_stages;  /* TMA->MMA ring depth */
    int num_sms;     /* grid size */
    int smem_bytes;  /* dynamic shared memory per CTA */
    int num_tiles;   /* upper bound on (cluster) tiles */
    int num_splits;  /* split-K slices (1 = none) */
    int cluster_split; /* != 0: the slices are the CTAs of one cluster, reduced through distributed shared memory */
    int tma_store;   /* != 0: output tiles are staged in shared memory (TMA stores; plain coalesced stores in the transposed orientation) */
    int swap_ab;     /* != 0: transposed-output orientation (tokens on the TMEM lanes, block_m = weight rows per tile) */
} dgb200_config;
/* Pure function (no CUDA): the configuration the heuristics pick for a problem on `num_sms` SMs.
 * gemm_type: 0 dense, 1 m-grouped contiguous, 2 m-grouped masked, 3 m-grouped contiguous psum.
 * Replaces get_best_config<SM100ArchSpec> (csrc/jit_kernels/heuristics/common.hpp:13-52). */
int dgb200_plan(int gemm_type, int m, int n, int k, int num_groups, int expected_m, int alignment, int num_sms,
                dgb200_config* out);
/* Configuration the last GEMM call on this thread used (DG_PRINT_CONFIGS analogue, heuristics/common.hpp:39-50). */
int dgb200_last_config(dgb200_config* out);
/* Measurement aid: issue-only `tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale` loop,
 * UMMA 256 x umma_n x 32, operands resident in shared memory (no TMA, no epilogue), `iters` 128-deep k-blocks per CTA
 * pair on num_sms SMs (<= 0: all). FLOPs = (num_sms / 2) * iters * 4 * 2 * 256 * umma_n * 32; the caller times it. */
int dgb200_debug_fp8_peak(int umma_n, int iters, int num_sms, void* stream);
/* Development aid: when set to a device buffer of (16 + 2 * grid) int64, CTA 0 of every GEMM launch stamps clock64()
 * at ten points of its life (entry, setup done, first TMA, first data, first MMA, last MMA, accumulator ready, stores
 * issued, teardown begin/end) into [0,10) (cluster split-K: outbox written / barrier / copies issued / partials landed
 * into [10,14)), and every CTA b stamps %globaltimer (ns) at entry / exit into
 * [16 + 2b], [17 + 2b]. NULL (default) disables it. */
int dgb200_debug_set_timestamps(void* device_int64_buffer);
/* Number of kernels launched by this library since process start (all threads). */
int64_t dgb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DGB200_H_ */
