// FP8 (E4M3) x FP8 blockwise-scaled GEMM for sm_100a -- the one hot kernel of this library.
//
//   D[m, n] (+)= sum_k  A[m, k] * 2^(sfa[m, k/g]-127)  *  B[n, k] * 2^(sfb[n, k/g]-127)        (FP32 accumulate)
//
// Replaces the reference's `sm100_fp8_fp4_gemm_1d1d_impl`
// (deep_gemm/include/deep_gemm/impls/sm100_fp8_fp4_gemm_1d1d.cuh:33-534) together with its scheduler
// (scheduler/gemm.cuh:39-325) and epilogues (epilogue/sm100_store_cd*.cuh), but it is a different program:
//
//  * AOT + runtime shapes. M/N/K, group count, tile height `block_m`, pipeline depth and SM count are kernel
//    *arguments* (the reference bakes all of them into a JIT-compiled template instance per shape).
//  * One orientation only: the weight operand B always sits on the 128 TMEM lanes (UMMA "A" side), the token
//    operand A on the TMEM columns (UMMA "N" side, any multiple of 16 up to 240). A CTA pair (cta_group::2) owns
//    256 weight rows x block_m tokens and shares the token tile, half of it in each CTA's shared memory.
//  * No shared-memory staging of the output and no TMA store: with tokens on TMEM columns every warp-level
//    `tcgen05.ld` hands each lane one output column n and consecutive rows m, so a plain store instruction
//    writes 32 consecutive n of one row (64 B bf16 / 128 B fp32, sector aligned). All 227 KB of smem feed the
//    TMA->MMA ring instead, rows are predicated exactly (masked / psum layouts never write invalid rows).
//  * Scale factors arrive in the reference's documented wire format (packed UE8M0 int32, MN-major,
//    csrc/utils/layout.hpp:100-107) so user-packed tensors keep working; a helper warp re-tiles each 128-word
//    group into the `tcgen05.cp` 32x128b layout before the MMA warp copies it into TMEM.
//  * Things only run-time shapes allow: the UMMA N of every tile follows its valid rows (`tile_n`), dense problems
//    use two tile heights so that the m-block COUNT fills whole rounds of CTA pairs, and small-M problems cut K
//    across a cluster of single-CTA MMAs that reduce through distributed shared memory (`kCSplit`).
//  * The TMA producer free-runs ahead of the setup barriers (its first loads need nothing the other warps set up).
//
// Warp roles (384 threads): w0 TMA producer | w1 MMA issuer (leader CTA) | w2 TMEM alloc + SF re-tiler |
//                           w3 idle | w4-11 epilogue (TMEM -> registers -> global; two warps per lane quadrant,
//                           interleaved over 32-row chunks so the exposed tail of the last tile is halved).
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "ptx.cuh"

namespace dgb200 {

enum GemmType : int {
    kDense = 0,         // D[M,N] = A[M,K] B[N,K]^T
    kMContiguous = 1,   // rows of A grouped, grouped_layout[r] = group id (-1 = padding)      (gemm.hpp:166)
    kMMasked = 2,       // A[G,Mmax,K], rows < masked_m[g] valid                                 (gemm.hpp:250)
    kMContiguousPsum = 3,  // grouped_layout[g] = end row of group g, starts aligned              (scheduler/gemm.cuh:217-237)
    kKGrouped = 4,      // D[g] += A[k_g, :M]^T B[k_g, :N], grouped_layout[g] = K of group g        (gemm.hpp:299, sched :259-283)
    kKGroupedPsum = 5,  // same, grouped_layout[g] = end K of group g, starts aligned to m_alignment
    kBatched = 6,       // D[b] = A[b] B[b]^T with arbitrary batch strides (3-D tensor maps)      (einsum.hpp:137-175, fp8_bmm)
    kBatchReduce = 7,   // D += sum_b A[b] B[b]^T: the K loop walks (batch, k-block); the batches are cut into chunks over the
                        // grid and every chunk adds its partial tile into the FP32 D            (einsum.hpp:22-60, 'bmk,bnk->mn')
};

constexpr uint32_t kBlockN = 128;        // weight rows per CTA == TMEM lanes
constexpr uint32_t kBlockK = 128;        // K bytes per pipeline stage == one 128B swizzle atom
constexpr uint32_t kUmmaK = 32;          // K per tcgen05.mma for 8-bit operands
constexpr uint32_t kMaxBlockM = 240;     // 2 accumulator buffers + SF columns must fit 512 TMEM columns
constexpr uint32_t kAccumColStride = 256;
constexpr uint32_t kTmemColSFW = 496;    // weight scale factors: 4 columns
constexpr uint32_t kTmemColSFX = 500;    // token scale factors: up to 8 columns
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kNumThreads = 384;
constexpr uint32_t kNumEpilogueThreads = 256;   // two warps per TMEM lane quadrant
constexpr uint32_t kWTileBytes = kBlockN * kBlockK;  // 16 KB
constexpr uint32_t kStoreRows = 16;                  // output rows per TMA store (one staging buffer = 16 rows x 128 columns bf16)
constexpr uint32_t kStoreBufBytes = kStoreRows * kBlockN * 2;       // 4 KB: two 128B-swizzled boxes of 16 rows x 64 columns
constexpr uint32_t kStoreStagingBytes = 2 * kStoreBufBytes;         // two buffers, shared by the eight epilogue warps
constexpr uint32_t kSwapStoreCols = 32;              // transposed-output staged epilogue: columns per unit (64 B per output row)
constexpr uint32_t kSwapStoreBufBytes = 32 * kSwapStoreCols * 2;    // 2 KB: 32 output rows (one warp's lanes) x 32 columns bf16
constexpr uint32_t kSwapStagingBytes = 8 * kSwapStoreBufBytes;      // one buffer per epilogue warp

struct GemmParams {
    void* d;                    // output (bf16 or fp32), row stride ld_d elements
    const int* grouped_layout;  // see GemmType
    uint32_t m;                 // dense: M | contiguous: sum of aligned M | masked: M_max
    uint32_t n, k;
    uint32_t num_groups;
    uint32_t block_m;           // token rows per tile, multiple of 16, <= 240
    uint32_t num_stages;        // TMA->MMA ring depth
    uint32_t ld_d;
    uint32_t num_kp_x;          // packed SF words along K per group (tokens)
    uint32_t num_kp_w;          // packed SF words along K per group (weights)
    uint32_t sf_shift_x;        // log2(k-blocks covered by one packed SF word): 2 (gran_k 128) or 0 (gran_k 32)
    uint32_t sf_shift_w;
    uint32_t swizzle_group;     // L2 tile-order group width (in n-units)
    float* splitk_ws;           // split-K: fp32 partial tiles [num_splits][m][n]   (dense only, else nullptr)
    int* splitk_counters;       // split-K: one arrival counter per (m-block, 128-column block), zero between launches
    uint32_t num_splits;        // K is cut into this many ranges of `kb_per_split` k-blocks (1 = no split)
    uint32_t kb_per_split;
    const uint32_t* arrival;    // psum layout fed by the EP dispatch: rows landed per group (cumulative), or nullptr
    const uint32_t* arrival_expected;   // ... and the value each counter reaches when the group is complete
    uint64_t w_hint, x_hint;    // L2 cache policies of the weight / token TMA loads (ptx.cuh kEvict*)
    long long* debug_ts;        // optional (development): CTA 0 stamps clock64() at 10 points of its life
    uint32_t num_n_units;       // ceil(n / (128 * cluster))
    uint32_t num_m_blocks;      // dense / contiguous: number of m-blocks
    uint32_t num_tall, block_m_low;   // dense: the first `num_tall` m-blocks are block_m rows high, the others
                                      // block_m_low (= block_m or block_m - 16): lets the host pick the m-block COUNT
                                      // that fills whole waves of CTA pairs instead of the height (0: all block_m)
    uint32_t m_alignment;       // contiguous layouts: group start alignment (K alignment for k-grouped psum)
    uint32_t zero_padding;      // psum: write zeros to [end, aligned end)
    uint32_t x_swizzle;         // MN-major tokens: swizzle width in bytes (128 / 64 / 32) = rows of one TMA box
    uint32_t sf_k_span;         // k-grouped: K elements covered by one packed SF word (4 * gran_k)
    uint32_t k_shift;           // k-grouped: log2(bytes per operand element); the group sizes arrive in elements, K is walked in bytes
    uint32_t grid_tiles;        // dense, one wave: exactly one tile per cluster, addressed by the 2-D grid (blockIdx.x / cluster =
                                // n-unit, blockIdx.y = m-block) -- the first tile is known without the ~500 cycles of index arithmetic
    uint64_t d_batch_stride;    // batched: elements between consecutive batches of D (0 otherwise)
    // Head-split output remap (fp8_gemm_nt_skip_head_mid, attention.hpp:19-74 / epilogue/transform.cuh:15-22): output
    // column n is stored at n + (n + head_right) / head_lr * head_mid, i.e. every (left | right) head of the GEMM's N
    // leaves a gap of `mid` untouched columns between its halves. head_mid == 0: identity.
    uint32_t head_lr, head_mid, head_right;
};

// ------------------------------------------------------------------------------------------------ scheduler
struct Tile {
    uint32_t x_row;     // first token row of the tile in the (group-flattened) A / SFA-column space
    uint32_t d_row;     // first output row
    uint32_t sfx_col;   // column (mn index) in the token SF map
    uint32_t sfx_row;   // first k-row in the token SF map
    uint32_t w_row;     // first weight row of THIS CTA in the group-flattened B
    uint32_t sfw_col;
    uint32_t sfw_row;
    uint32_t n0;        // first output column of THIS CTA
    uint32_t valid_m;   // rows [0, valid_m) of the tile are real outputs
    uint32_t store_m;   // rows [0, store_m) are written (>= valid_m only when zero padding is requested)
    uint32_t kb_begin, kb_end;  // k-blocks this tile accumulates (a sub-range only under split-K)
    uint32_t split;             // split-K slice index
    uint32_t counter_idx;       // split-K arrival counter of this CTA's output block
    uint32_t k_base;            // K offset of the tile's group in A / B (k-grouped), else 0
    uint32_t wk_base;           // K-row offset of the tile's group in an MN-major grouped B ([G,K,N] flattened), else 0
    uint32_t last_umma;         // UMMAs (32 K-elements each) to issue in the last k-block (4 unless K ends inside it)
    uint32_t batch;             // batched: third tensor-map coordinate / D batch index (0 otherwise)
};

// kCluster = CTAs per cluster: 1 (single CTA MMA), 2 (one cta_group::2 pair) or 4 / 8 (2 / 4 pairs that work on
// consecutive m-blocks of the SAME weight panel and share its TMA loads by multicast; dense only).
// kCSplit (0 = off, else the number of K slices): the cluster is kCSplit MMA groups -- single CTAs (kCluster == kCSplit) or
// CTA pairs (kCluster == 2 kCSplit) -- that cut K of ONE output tile between them (split-K inside the cluster, reduced
// through distributed shared memory); `rank` is then the position inside the MMA group and `split_rank` the slice.
template <int kGemmType, int kCluster, bool kSplitK = false, int kCSplit = 0>
struct Scheduler {
    static constexpr uint32_t kPairs = (kCluster >= 2 && !kCSplit) ? kCluster / 2 : 1;
    static constexpr uint32_t kCtaGroup = kCSplit ? kCluster / kCSplit : (kCluster >= 2 ? 2 : 1);
    const GemmParams& p;
    uint32_t cta_rank, cluster_id, num_clusters;
    uint32_t num_n_units;
    uint32_t iter = 0;
    // grouped walk state
    uint32_t g = 0, unit_cum = 0, row_start = 0, row_end = 0;
    // k-grouped walk state: K range of the current group, packed-SF rows before it, non-empty groups before it
    uint32_t k_start = 0, k_end = 0, sf_rows_before = 0;
    bool kg_loaded = false;

    uint32_t split_rank;
    __device__ Scheduler(const GemmParams& p_, uint32_t rank, uint32_t split_rank_ = 0) : p(p_), cta_rank(rank), split_rank(split_rank_) {
        cluster_id = blockIdx.x / kCluster;
        num_clusters = gridDim.x / kCluster;
        num_n_units = p.num_n_units;
        if constexpr (kGemmType == kMContiguousPsum) row_end = static_cast<uint32_t>(max(0, __ldg(p.grouped_layout)));
    }

    // L2-friendly order inside one problem of `num_m` m-blocks: walk `swizzle_group` n-units at a time.
    __device__ void split(uint32_t local, uint32_t num_m, uint32_t& m_blk, uint32_t& n_unit) const {
        const uint32_t gw = p.swizzle_group;
        const uint32_t per_group = gw * num_m;
        const uint32_t grp = local < per_group ? 0u : local / per_group;     // (integer division is ~25 instructions)
        const uint32_t first = grp * gw;
        const uint32_t in = local - grp * per_group;
        const uint32_t width = min(gw, num_n_units - first);
        m_blk = in < width ? 0u : in / width;
        n_unit = first + in - m_blk * width;
    }

    __device__ bool next(Tile& t) {
        if constexpr (kCSplit) {
            // one tile per cluster, addressed by the 2-D grid: (blockIdx.x / kCluster, blockIdx.y) = (n-tile, m-block)
            if (iter++ != 0) return false;
            const uint32_t num_kb = (p.k + kBlockK - 1) / kBlockK;
            t.split = split_rank, t.counter_idx = 0, t.k_base = 0, t.wk_base = 0, t.batch = 0;
            t.kb_begin = split_rank * p.kb_per_split;
            t.kb_end = min(num_kb, t.kb_begin + p.kb_per_split);
            t.last_umma = t.kb_end != num_kb ? kBlockK / kUmmaK : ((p.k - (num_kb - 1) * kBlockK) + kUmmaK - 1) / kUmmaK;
            t.x_row = blockIdx.y * p.block_m;
            t.d_row = t.x_row, t.sfx_col = t.x_row, t.sfx_row = 0;
            t.valid_m = min(p.block_m, p.m - t.x_row);
            t.store_m = t.valid_m;
            t.n0 = (cluster_id * kCtaGroup + (cta_rank & 1)) * kBlockN;
            t.w_row = t.n0, t.sfw_col = t.n0, t.sfw_row = 0;
            return true;
        }
        const uint32_t idx = cluster_id + (iter++) * num_clusters;
        uint32_t m_blk, n_unit, group = 0;
        const uint32_t num_kb_total = (p.k + kBlockK - 1) / kBlockK;
        t.kb_begin = 0, t.kb_end = num_kb_total, t.split = 0, t.counter_idx = 0;
        t.k_base = 0, t.wk_base = 0, t.batch = 0;
        t.last_umma = ((p.k - (num_kb_total - 1) * kBlockK) + kUmmaK - 1) / kUmmaK;
        if constexpr (kGemmType == kKGrouped || kGemmType == kKGroupedPsum) {
            // groups are walked in order; every non-empty group contributes num_m_blocks * num_n_units tiles
            const uint32_t per_group = p.num_m_blocks * num_n_units;
            auto load_group = [&]() {
                const uint32_t v = static_cast<uint32_t>(max(0, __ldg(p.grouped_layout + g))) << p.k_shift;
                if constexpr (kGemmType == kKGroupedPsum) {
                    k_start = (k_end + p.m_alignment - 1) / p.m_alignment * p.m_alignment;
                    k_end = max(k_start, v);
                } else {
                    k_start = k_end;
                    k_end = k_start + v;
                }
            };
            if (!kg_loaded) {
                if (p.num_groups == 0) return false;
                load_group();
                kg_loaded = true;
            }
            while (true) {
                const bool empty = k_end == k_start;
                if (!empty && idx < (unit_cum + 1) * per_group) break;
                if (!empty) {
                    unit_cum += 1;
                    sf_rows_before += (k_end - k_start + p.sf_k_span - 1) / p.sf_k_span;
                }
                if (++g >= p.num_groups) return false;
                load_group();
            }
            split(idx - unit_cum * per_group, p.num_m_blocks, m_blk, n_unit);
            const uint32_t kg = k_end - k_start;
            t.kb_end = (kg + kBlockK - 1) / kBlockK;
            t.last_umma = ((kg - (t.kb_end - 1) * kBlockK) + kUmmaK - 1) / kUmmaK;
            t.k_base = k_start;
            t.x_row = m_blk * p.block_m;
            t.d_row = g * p.m + t.x_row;
            t.sfx_col = t.x_row;
            t.sfx_row = sf_rows_before;
            t.valid_m = min(p.block_m, p.m - t.x_row);
            t.store_m = t.valid_m;
            t.n0 = (n_unit * kCtaGroup + (cta_rank & 1)) * kBlockN;
            t.w_row = t.n0;
            t.sfw_col = t.n0;
            t.sfw_row = sf_rows_before;
            return true;
        }
        if constexpr (kGemmType == kDense || kGemmType == kMContiguous) {
            const uint32_t num_m = (p.num_m_blocks + kPairs - 1) / kPairs;   // m-blocks are handed out kPairs at a time
            uint32_t local = idx;
            if (kSplitK) {
                // split-K: slice index varies slowest, so the CTAs of one slice stream disjoint weight panels
                const uint32_t per_split = num_m * num_n_units;
                if (idx >= per_split * p.num_splits) return false;
                t.split = idx / per_split;
                local = idx - t.split * per_split;
                t.kb_begin = t.split * p.kb_per_split;
                t.kb_end = min(num_kb_total, t.kb_begin + p.kb_per_split);
                if (t.kb_end != num_kb_total) t.last_umma = kBlockK / kUmmaK;
            } else if (idx >= num_m * num_n_units) {
                return false;
            }

            if (kGemmType == kDense && !kSplitK && kPairs == 1 && p.grid_tiles) {
                if (iter != 1) return false;                                 // (iter was advanced above) one tile per cluster
                m_blk = blockIdx.y, n_unit = cluster_id;
            } else {
                split(local, num_m, m_blk, n_unit);
            }
            m_blk = m_blk * kPairs + (cta_rank >> 1);                        // this pair's m-block inside the group
            uint32_t height = p.block_m;
            t.x_row = m_blk * p.block_m;
            if (kGemmType == kDense && m_blk >= p.num_tall) {                // two tile heights (dense only)
                height = p.block_m_low;
                t.x_row = p.num_tall * p.block_m + (m_blk - p.num_tall) * p.block_m_low;
            }
            t.d_row = t.x_row;
            t.sfx_col = t.x_row;
            t.sfx_row = 0;
            t.valid_m = t.x_row < p.m ? min(height, p.m - t.x_row) : 0u;     // 0: a pair past the last m-block idles along
            t.store_m = t.valid_m;
            t.counter_idx = m_blk * (num_n_units * kCtaGroup) + n_unit * kCtaGroup + (cta_rank & 1);
            if constexpr (kGemmType == kMContiguous) {
                group = static_cast<uint32_t>(max(0, __ldg(p.grouped_layout + t.x_row)));
                // An expert's rows are a prefix of its aligned segment, the rest is padding (-1): count the 16-row groups that
                // start with a real row (independent loads, one L2 round trip) and leave the padding groups alone -- neither
                // multiplied (the UMMA N of the tile follows valid_m) nor written. With mean M = 128 and 128-row alignment a
                // third of the rows the reference multiplies are padding (its BLOCK_M is a compile-time constant).
                uint32_t groups = 0;
                const uint32_t max_groups = (t.valid_m + 15) / 16;
#pragma unroll
                for (uint32_t j = 1; j < kMaxBlockM / 16; ++j)
                    if (j < max_groups) groups += __ldg(p.grouped_layout + t.x_row + 16 * j) >= 0 ? 1u : 0u;
                t.valid_m = min(t.valid_m, 16 * (groups + 1));
                t.store_m = t.valid_m;
            }
        } else if constexpr (kGemmType == kBatchReduce) {
            // tile = (batch chunk, m-block, n-unit); num_splits chunks of kb_per_split batches each. The k-blocks of a tile are
            // VIRTUAL: kb counts (batch, k-block) pairs, so the MMA and re-tiling roles need not know; only the producer
            // turns kb back into coordinates.
            const uint32_t per_chunk = p.num_m_blocks * num_n_units;
            if (idx >= per_chunk * p.num_splits) return false;
            const uint32_t chunk = idx < per_chunk ? 0u : idx / per_chunk;
            split(idx - chunk * per_chunk, p.num_m_blocks, m_blk, n_unit);
            t.batch = chunk * p.kb_per_split;
            const uint32_t batches = min(p.num_groups - t.batch, p.kb_per_split);
            t.kb_begin = 0, t.kb_end = batches * num_kb_total;
            t.last_umma = kBlockK / kUmmaK;                       // (the host requires whole k-blocks per batch)
            t.x_row = m_blk * p.block_m;
            t.d_row = t.x_row;
            t.sfx_col = t.x_row, t.sfx_row = 0;
            t.valid_m = min(p.block_m, p.m - t.x_row);
            t.store_m = t.valid_m;
            t.n0 = (n_unit * kCtaGroup + (cta_rank & 1)) * kBlockN;
            t.w_row = t.n0;
            t.sfw_col = t.n0, t.sfw_row = 0;
            return true;
        } else if constexpr (kGemmType == kBatched) {
            // every batch is a full [m, n] problem; batches are walked in order (tensor-map coordinate 2 = batch)
            const uint32_t per_batch = p.num_m_blocks * num_n_units;
            if (idx >= per_batch * p.num_groups) return false;
            g = idx < per_batch ? 0u : idx / per_batch;
            split(idx - g * per_batch, p.num_m_blocks, m_blk, n_unit);
            t.batch = g;
            t.x_row = m_blk * p.block_m;
            t.d_row = t.x_row;
            t.sfx_col = t.x_row;
            t.sfx_row = g * p.num_kp_x;
            t.valid_m = min(p.block_m, p.m - t.x_row);
            t.store_m = t.valid_m;
            t.n0 = (n_unit * kCtaGroup + (cta_rank & 1)) * kBlockN;
            t.w_row = t.n0;
            t.sfw_col = t.n0;
            t.sfw_row = g * p.num_kp_w;
            return true;
        } else if constexpr (kGemmType == kMMasked) {
            uint32_t num_m;
            while (true) {
                if (g >= p.num_groups) return false;
                const uint32_t mg = min(static_cast<uint32_t>(max(0, __ldg(p.grouped_layout + g))), p.m);
                num_m = (mg + p.block_m - 1) / p.block_m;
                if (idx < (unit_cum + num_m) * num_n_units) {
                    row_end = mg;
                    break;
                }
                unit_cum += num_m, ++g;
            }
            split(idx - unit_cum * num_n_units, num_m, m_blk, n_unit);
            group = g;
            const uint32_t m0 = m_blk * p.block_m;
            t.x_row = g * p.m + m0;
            t.d_row = t.x_row;
            t.sfx_col = m0;
            t.sfx_row = g * p.num_kp_x;
            t.valid_m = min(p.block_m, row_end - m0);
            t.store_m = t.valid_m;
        } else {  // kMContiguousPsum
            uint32_t num_m, cover_end;
            while (true) {
                // rows [row_start, row_end) are real; with zero padding the gap up to the aligned end is written too.
                // Everything is clamped to the buffer: a caller whose `m` is not a multiple of the alignment (or whose
                // prefix sums run past it -- an overflowing EP dispatch) gets short segments, never rows >= m.
                row_start = min(row_start, p.m), row_end = min(max(row_end, row_start), p.m);
                cover_end = p.zero_padding ? min(p.m, (row_end + p.m_alignment - 1) / p.m_alignment * p.m_alignment) : row_end;
                cover_end = max(cover_end, row_start);
                num_m = (cover_end - row_start + p.block_m - 1) / p.block_m;
                if (idx < (unit_cum + num_m) * num_n_units) break;
                unit_cum += num_m;
                if (++g >= p.num_groups) return false;
                row_start = (row_end + p.m_alignment - 1) / p.m_alignment * p.m_alignment;
                row_end = max(row_start, static_cast<uint32_t>(max(0, __ldg(p.grouped_layout + g))));
            }
            split(idx - unit_cum * num_n_units, num_m, m_blk, n_unit);
            group = g;
            t.x_row = row_start + m_blk * p.block_m;
            t.d_row = t.x_row;
            t.sfx_col = t.x_row;
            t.sfx_row = 0;
            t.valid_m = t.x_row < row_end ? min(p.block_m, row_end - t.x_row) : 0u;   // 0: a pure padding tile
            t.store_m = min(p.block_m, cover_end - t.x_row);
        }
        t.n0 = (n_unit * kCtaGroup + (cta_rank & 1)) * kBlockN;
        t.w_row = group * p.n + t.n0;
        t.wk_base = group * p.k;
        t.sfw_col = t.n0;
        t.sfw_row = group * p.num_kp_w;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------ epilogue helpers
template <typename out_t>
__device__ __forceinline__ void store_out(out_t* ptr, float v, bool accumulate);

template <>
__device__ __forceinline__ void store_out<float>(float* ptr, float v, bool accumulate) {
    if (accumulate) v += *ptr;
    *ptr = v;
}
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* ptr, float v, bool accumulate) {
    __nv_bfloat16 r = __float2bfloat16_rn(v);
    // Same arithmetic as the reference's memory-side `cp.reduce.async.bulk ... add` on a BF16 tile
    // (epilogue/sm100_store_cd.cuh:126-128): round the accumulator to BF16 first, then one BF16 add.
    if (accumulate) r = __hadd(r, *ptr);
    *ptr = r;
}

__device__ __forceinline__ float as_f32(float v) { return v; }
__device__ __forceinline__ float as_f32(uint32_t v) { return __uint_as_float(v); }

// kRows consecutive output rows of one warp-wide column group (row r at `row + r * row_bytes`), the first `valid` of them
// written. Accumulating into C:
//  * FP32: a memory-side add, `red.global.add.f32` (one per element, 128 contiguous bytes per warp instruction): the add
//    happens in L2, nothing comes back to the SM -- the same thing the reference's TMA reduce-add does
//    (epilogue/sm100_store_cd.cuh:126-128, `cp.reduce.async.bulk ... add.f32`), incl. its flush of subnormals.
//  * BF16: round the accumulator, then one BF16 add (the reference's `add.noftz.bf16` reduce); all rows are READ first and
//    only then added and stored -- as one read-modify-write per row the loads serialise behind the (possibly aliasing)
//    stores, one DRAM round trip per row (the k-grouped weight-gradient GEMM ran 4.5x slower than the reference that way).
template <typename out_t, uint32_t kRows, bool kAccumulate, typename value_t>
__device__ __forceinline__ void store_rows(char* row, size_t row_bytes, const value_t* v, uint32_t valid) {
    if constexpr (kAccumulate && std::is_same_v<out_t, float>) {
        if (valid >= kRows) {
#pragma unroll
            for (uint32_t j = 0; j < kRows; ++j)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(row + j * row_bytes), "f"(as_f32(v[j])) : "memory");
        } else {
#pragma unroll
            for (uint32_t j = 0; j < kRows; ++j)
                if (j < valid) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(row + j * row_bytes), "f"(as_f32(v[j])) : "memory");
        }
    } else if constexpr (kAccumulate) {
        __nv_bfloat16 prev[kRows];
#pragma unroll
        for (uint32_t j = 0; j < kRows; ++j)
            if (j < valid) prev[j] = *reinterpret_cast<const __nv_bfloat16*>(row + j * row_bytes);
#pragma unroll
        for (uint32_t j = 0; j < kRows; ++j)
            if (j < valid) *reinterpret_cast<__nv_bfloat16*>(row + j * row_bytes) = __hadd(__float2bfloat16_rn(as_f32(v[j])), prev[j]);
    } else {
        if (valid >= kRows) {
#pragma unroll
            for (uint32_t j = 0; j < kRows; ++j) store_out<out_t>(reinterpret_cast<out_t*>(row + j * row_bytes), as_f32(v[j]), false);
        } else {
#pragma unroll
            for (uint32_t j = 0; j < kRows; ++j)
                if (j < valid) store_out<out_t>(reinterpret_cast<out_t*>(row + j * row_bytes), as_f32(v[j]), false);
        }
    }
}

// ------------------------------------------------------------------------------------------------ the kernel
// Shared memory: `num_stages` contiguous stage slots, then the barriers.
//   slot  = [ W 128x128 B | X load_m x 128 B | SFW 512 B | SFX groups x 512 B | pad to 1 KB ]
// Every role walks the ring with two running offsets (slot byte offset, barrier byte offset): the k-loops contain no
// multiplications, no generic->shared conversions and no 64-bit address arithmetic. (With runtime shapes the naive
// form cost ~450 cycles per k-block in the single-thread TMA producer and capped the small-M shapes.)
struct Ring {
    uint32_t slot, bar, phase;       // byte offset of the current slot, byte offset of its barrier, phase bit
    uint32_t slot_stride, bar_end;
    __device__ __forceinline__ Ring(uint32_t slot_stride_, uint32_t num_stages)
        : slot(0), bar(0), phase(0), slot_stride(slot_stride_), bar_end(num_stages * 8) {}
    __device__ __forceinline__ void advance() {
        slot += slot_stride, bar += 8;
        if (bar == bar_end) slot = 0, bar = 0, phase ^= 1;
    }
};

__host__ __device__ constexpr uint32_t slot_bytes(uint32_t block_m, uint32_t cluster) {
    return (kWTileBytes + (block_m / cluster) * kBlockK + 512 + ((block_m + 127) / 128) * 512 + 1023) / 1024 * 1024;
}

// Split-K finalisation by the last slice to arrive: D rows = sum over slices (in slice order) of the FP32 partials.
// All kRows x kSplits 16-byte loads of a batch are issued before the first add, so a batch costs one L2 round trip.
template <uint32_t kSplits, uint32_t kRows, typename out_t, bool kAccumulate>
__device__ __forceinline__ void splitk_finalize(const float* ws, size_t slice_elems, uint32_t n, out_t* d, uint32_t ld_d,
                                                uint32_t row0, uint32_t valid_m, uint32_t nc, uint32_t warp, uint32_t num_warps) {
    for (uint32_t r0 = warp * kRows; r0 < valid_m; r0 += num_warps * kRows) {
        float4 x[kRows][kSplits];
        const float* src = ws + static_cast<size_t>(row0 + r0) * n + nc;
#pragma unroll
        for (uint32_t i = 0; i < kRows; ++i)
#pragma unroll
            for (uint32_t sl = 0; sl < kSplits; ++sl)
                x[i][sl] = r0 + i < valid_m ? __ldcg(reinterpret_cast<const float4*>(src + static_cast<size_t>(i) * n + sl * slice_elems))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (uint32_t i = 0; i < kRows; ++i) {
            if (r0 + i >= valid_m) break;
            float4 acc = x[i][0];
#pragma unroll
            for (uint32_t sl = 1; sl < kSplits; ++sl) acc.x += x[i][sl].x, acc.y += x[i][sl].y, acc.z += x[i][sl].z, acc.w += x[i][sl].w;
            out_t* dst = d + static_cast<size_t>(row0 + r0 + i) * ld_d + nc;
            store_out<out_t>(dst + 0, acc.x, kAccumulate);
            store_out<out_t>(dst + 1, acc.y, kAccumulate);
            store_out<out_t>(dst + 2, acc.z, kAccumulate);
            store_out<out_t>(dst + 3, acc.w, kAccumulate);
        }
    }
}

#define DGB_STAMP(i)                                                            \
    do {                                                                        \
        if (p.debug_ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0) p.debug_ts[i] = clock64(); \
    } while (0)

// kXMn / kWMn: the token / weight operand is MN-major in global memory (its M / N extent is contiguous, K strided):
// fp8_gemm_{nn,tn,tt}, m_grouped nn, and both operands of the K-grouped weight-gradient GEMM.
// kSplitK: the dense split-K variant (K slices + finalising pass); kept out of the common instantiations because its
// epilogue doubles the code size, which a cold instruction cache charges to every short launch.
// kCSplit (number of slices): split-K inside a cluster of kCSplit MMA groups (single CTAs, or CTA pairs for taller tiles;
// dense, small / medium M): every group accumulates one K slice of the same output tile; the partial tiles are exchanged through distributed shared memory (reduce-scatter over the
// token columns, `st.async` + transaction barrier) and added in slice order, so each weight byte crosses L2->SM once
// instead of once per m-block and nothing goes through global memory. One tile per cluster (the host guarantees it).
// kTmaStore: BF16 output tiles leave through shared memory: TMEM -> registers (16x256b fragments) -> BF16 pairs ->
// `stmatrix.trans` into a 128B-swizzled staging box -> `cp.async.bulk.tensor` store, 16 output rows at a time, two
// 4 KB staging buffers (replaces the reference's epilogue/sm100_store_cd_swap_ab.cuh:22-128). The stores
// are asynchronous, so the epilogue warps are done with a tile once its accumulator has been read; rows / columns past
// the end of D are clipped by the tensor map. Used for tall tiles (dense, contiguous); the direct-store epilogue stays
// for small tiles (all shared memory feeds the ring) and for layouts that need exact row predication (masked, psum).
// kSwapD: the second orientation. The host hands the TOKENS to the lane side ("w": 128 rows per CTA) and the WEIGHTS to the
// column side ("x": block_m rows per tile, any multiple of 16), so TMEM lane = output row, TMEM column = output column, and
// the epilogue writes D[lane][column]: every thread owns one output row and stores 16 consecutive columns per TMEM load.
// kBf16AB: BF16 operands without scale factors (the reference's bf16_gemm family, impls/sm100_bf16_gemm.cuh:34-420) on the
// same skeleton: `tcgen05.mma.kind::f16`, 16 K-elements per instruction. K-major operands are addressed in BYTES (the host
// passes k = 2 K and UINT8 tensor maps), so a pipeline stage is again one 128-byte swizzle atom per row (64 BF16 of K) and
// nothing else in the kernel changes; the scale-factor loads, the re-tiling and the tcgen05.cp are compiled out.
// What it buys is tile-count freedom along N: with few token rows the number of tiles is N / block_m for ANY block_m, so the
// tiles can be cut to fill exactly one wave of SMs (the reference reaches the same through its non-swap-AB templates,
// csrc/jit_kernels/heuristics/sm100.hpp:27-92). Dense, K-major operands.
template <int kGemmType, int kCluster, typename out_t, bool kAccumulate, bool kXMn = false, bool kWMn = false,
          bool kSplitK = false, int kCSplit = 0, bool kTmaStore = false, bool kSwapD = false, bool kBf16AB = false>
__global__ void __launch_bounds__(kNumThreads, 1)
fp8_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_sfx, const __grid_constant__ CUtensorMap map_sfw,
                const __grid_constant__ CUtensorMap map_d, const __grid_constant__ GemmParams p) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
    using namespace ptx;
    extern __shared__ __align__(1024) uint8_t smem[];

    const uint32_t warp_idx = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0);
    const uint32_t lane = lane_id();
    if (threadIdx.x == 0) DGB_STAMP(0);
    if (threadIdx.x == 0 && p.debug_ts != nullptr) p.debug_ts[16 + 2 * blockIdx.x] = globaltimer_ns();   // per-CTA entry
    constexpr int kCtaGroup = kCSplit ? kCluster / kCSplit : (kCluster >= 2 ? 2 : 1);   // CTAs per UMMA (cta_group)
    static_assert(kCtaGroup == 1 || kCtaGroup == 2, "an MMA group is one CTA or a CTA pair");
    constexpr uint32_t kPairs = (kCluster >= 2 && !kCSplit) ? kCluster / 2 : 1;   // CTA pairs per cluster (share the weight loads)
    const uint32_t cluster_rank = kCluster == 1 ? 0u : cluster_ctarank();
    // position inside the MMA group(s): cluster split-K counts inside its own group, the others across the cluster
    const uint32_t cta_rank = kCSplit ? cluster_rank % kCtaGroup : cluster_rank;
    const uint32_t split_rank = kCSplit ? cluster_rank / kCtaGroup : 0u;              // K slice (cluster split-K)
    const uint32_t pair_idx = cta_rank >> 1;
    const uint32_t leader_rank = kCSplit ? cluster_rank - cta_rank : (cta_rank & ~1u);   // CLUSTER rank of this group's leader CTA
    const bool is_leader = (cta_rank & 1) == 0;

    // ---- shared memory carve-up (all sizes are runtime values), as 32-bit shared::cta addresses
    const uint32_t num_stages = p.num_stages;
    const uint32_t load_m = p.block_m / kCtaGroup;                          // token rows this CTA loads per stage
    const uint32_t x_tile_bytes = load_m * kBlockK;                         // multiple of 1024 (load_m % 8 == 0)
    const uint32_t num_sfx_groups = (p.block_m + 127) / 128;                // 128-row UTCCP groups of token SFs
    const uint32_t slot_stride = slot_bytes(p.block_m, kCtaGroup);
    static_assert(!kTmaStore || (std::is_same_v<out_t, __nv_bfloat16> && !kAccumulate && !kSplitK && !kCSplit && (kCluster == 2 || kSwapD)),
                  "the TMA-store epilogue is built for plain BF16 output tiles");
    static_assert(!kSwapD || (kGemmType == kDense && !kXMn && !kWMn && !kSplitK && !kCSplit && kCluster <= 2),
                  "the transposed-output orientation is built for plain dense K-major problems");
    static_assert(!kBf16AB || (!kSplitK && !kSwapD && (kCluster <= 2 || kCSplit)), "BF16 operands: plain and cluster split-K kernels");
    // Operand bytes per element, and what one pipeline stage / one UMMA covers along K in ELEMENTS (= rows of an MN-major
    // operand): FP8 128 / 32, BF16 64 / 16. Everything K-major is addressed in bytes and does not care.
    constexpr uint32_t kEl = kBf16AB ? 2 : 1;
    constexpr uint32_t kKRows = kBlockK / kEl, kUmmaKRows = kUmmaK / kEl;
    const uint32_t staging = smem_u32(smem);                               // kTmaStore: 2 buffers x 4 KB (transposed output: 8 x 2 KB)
    const uint32_t smem_base = staging + (kTmaStore ? (kSwapD ? kSwapStagingBytes : kStoreStagingBytes) : 0u);   // the TMA -> MMA ring starts here
    const uint32_t off_x = kWTileBytes, off_sfw = off_x + x_tile_bytes, off_sfx = off_sfw + 512;
    const uint32_t bars = smem_base + num_stages * slot_stride;
    const uint32_t full_bar = bars;                            // TMA bytes landed (per CTA)
    const uint32_t empty_bar = bars + num_stages * 8;          // MMAs that read the slot retired (per CTA, via commit)
    const uint32_t ready_bar = bars + num_stages * 16;         // slot landed in every CTA + SFs re-tiled (leader only)
    const uint32_t tmem_full_bar = bars + num_stages * 24;     // [2] accumulator complete (per CTA, via commit)
    const uint32_t tmem_empty_bar = tmem_full_bar + 16;        // [2] accumulator drained by all epilogue threads (leader)
    const uint32_t tmem_ptr_smem = tmem_empty_bar + 16;
    const uint32_t splitk_flag_smem = tmem_ptr_smem + 4;
    const uint32_t red_bar = tmem_ptr_smem + 8;                // cluster split-K: partial tiles of the peers have landed
    const uint32_t red_stage = (tmem_ptr_smem + 16 + 15) & ~15u;   // cluster split-K: float4 [kCluster-1][chunk/4][128]

    // Setup, arranged so that nothing waits that does not have to.
    //  * Warp 0 (TMA producer) initialises the barriers only it and the MMA commits touch (full / empty), warp 1 the
    //    rest; both publish them with `fence.mbarrier_init.release.cluster` + a relaxed arrive on the ONE cluster barrier
    //    of the prologue (a release-arrive would cost a MEMBAR.ALL.GPU). Passing that barrier therefore means: every
    //    CTA of the cluster is resident and all its mbarriers are live.
    //  * Warp 0 does not wait for it: its loads are local and the first thing a peer can do to it -- an MMA commit on
    //    `empty` -- happens only after the peer has passed the barrier, i.e. after warp 0's own arrive. So the producer
    //    starts streaming ~150 cycles into the kernel and collects the barrier phase after its loop. (Multi-pair
    //    clusters multicast into their peers' shared memory, so there it waits like everyone else.)
    //  * The others wait, warp 2 allocates tensor memory, and a CTA-local named barrier (warps 1..11) publishes the
    //    TMEM address. The peer's allocation is ordered before the leader's first MMA by the peer's re-tiler warp,
    //    which allocates first and arrives on the leader's `ready` barrier afterwards.
    // (Without a cluster there is no cluster barrier to carry warp 0's mbarrier initialisation to the other warps, so a
    // single-CTA launch takes the plain __syncthreads() prologue.)
    constexpr bool kEarlyProducer = kPairs == 1 && kCluster > 1;
    if constexpr (kCSplit != 0) {
        // One tile per cluster, known from the block index alone: the otherwise idle warp 3 asks L2 for this CTA's first
        // weight / token / scale tiles as its very first instructions. A cold launch (page walks, DRAM row opens) then has
        // ~1000 cycles of head start on the producer's first TMA load, which has to wait for the barrier setup.
        if (warp_idx == 3 && elect_one()) {
            const uint32_t n0 = ((blockIdx.x / kCluster) * kCtaGroup + (cta_rank & 1)) * kBlockN;
            const uint32_t kb0 = split_rank * p.kb_per_split;
            const uint32_t x_row0 = blockIdx.y * p.block_m + (cta_rank & 1) * (p.block_m / kCtaGroup);
            prefetch_tensormap(&map_w);
            prefetch_tensormap(&map_x);
            tma_prefetch_2d(&map_w, kb0 * kBlockK, n0);
            tma_prefetch_2d(&map_x, kb0 * kBlockK, x_row0);
            if constexpr (!kBf16AB) {
                tma_prefetch_2d(&map_sfw, n0, kb0 >> p.sf_shift_w);
                tma_prefetch_2d(&map_sfx, blockIdx.y * p.block_m, kb0 >> p.sf_shift_x);
            }
            const uint32_t kb_last = (p.k + kBlockK - 1) / kBlockK - 1;
#pragma unroll
            for (uint32_t j = 1; j < 4; ++j)
                if (kb0 + j <= kb_last) tma_prefetch_2d(&map_w, (kb0 + j) * kBlockK, n0);
        }
    }
    if constexpr (kCSplit == 0 && kGemmType == kDense && !kXMn && !kWMn && kPairs == 1) {
        // Same head start for the first tile of a persistent dense launch (its coordinates cost one scheduler step).
        if (warp_idx == 3 && elect_one()) {
            Scheduler<kGemmType, kCluster, kSplitK, kCSplit> sched(p, cta_rank, split_rank);
            Tile t;
            if (sched.next(t) && t.valid_m > 0) {
                const uint32_t rows = max(16u, min(p.block_m, (t.valid_m + 15u) & ~15u));
                const uint32_t x_row0 = t.x_row + (cta_rank & 1) * (rows / kCtaGroup);
                prefetch_tensormap(&map_w);
                prefetch_tensormap(&map_x);
                tma_prefetch_2d(&map_w, t.kb_begin * kBlockK, t.w_row);
                tma_prefetch_2d(&map_x, t.kb_begin * kBlockK, x_row0);
                if constexpr (!kBf16AB) {
                    tma_prefetch_2d(&map_sfw, t.sfw_col, t.sfw_row + (t.kb_begin >> p.sf_shift_w));
                    tma_prefetch_2d(&map_sfx, t.sfx_col, t.sfx_row + (t.kb_begin >> p.sf_shift_x));
                }
#pragma unroll
                for (uint32_t j = 1; j < 4; ++j)
                    if (t.kb_begin + j < t.kb_end) tma_prefetch_2d(&map_w, (t.kb_begin + j) * kBlockK, t.w_row);
            }
        }
    }
    bool producer_lane = false;
    if (warp_idx == 0) {
        for (uint32_t i = lane; i < num_stages; i += 32) {
            mbar_init(full_bar + i * 8, 1);
            mbar_init(empty_bar + i * 8, kPairs);      // one commit per pair: peers multicast weights into this slot too
        }
        fence_mbar_init();
        __syncwarp();
        producer_lane = elect_one();
        if (producer_lane) {
            prefetch_tensormap(&map_x);
            prefetch_tensormap(&map_w);
            if constexpr (!kBf16AB) {
                prefetch_tensormap(&map_sfx);
                prefetch_tensormap(&map_sfw);
            }
            if constexpr (kTmaStore && !kSwapD) prefetch_tensormap(&map_d);
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        for (uint32_t i = lane; i < num_stages; i += 32) mbar_init(ready_bar + i * 8, 32 * kCtaGroup);
        if (lane < 2) {
            mbar_init(tmem_full_bar + lane * 8, 1);
            mbar_init(tmem_empty_bar + lane * 8, kNumEpilogueThreads * kCtaGroup);
        }
        if (kCSplit && lane == 2) mbar_init(red_bar, 1);
        fence_mbar_init();
        __syncwarp();
    }
    if constexpr (kCluster > 1) cluster_arrive_relaxed();
    uint32_t tmem_base = 0;
    if (warp_idx != 0 || !kEarlyProducer) {
        if constexpr (kCluster > 1) cluster_wait();
        if (warp_idx == 2) tmem_alloc<kCtaGroup>(tmem_ptr_smem, kTmemCols);
        tcgen05_fence_before();
        if constexpr (kEarlyProducer)
            named_bar_sync(2, kNumThreads - 32);           // warps 1..11; warp 0 needs nothing that is set up here
        else
            __syncthreads();
        tcgen05_fence_after();
        tmem_base = ld_shared_u32(tmem_ptr_smem);
    }

    // Programmatic dependent launch: everything above overlaps the previous kernel's tail. A launch that synchronises
    // with its producer through the per-group arrival counters (EP dispatch still in flight) must not wait for it.
    // (The next kernel of a programmatically chained stream may start ITS prologue -- barrier setup, TMEM allocation, the
    // L2 prefetch of its weights -- as soon as every CTA of this grid has got here; it still waits for our results.)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (p.arrival == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) DGB_STAMP(1);

    const uint32_t sfw_mask = (1u << p.sf_shift_w) - 1, sfx_mask = (1u << p.sf_shift_x) - 1;
    Ring ring(slot_stride, num_stages);
    // UMMA N of a tile = its valid token rows rounded up to 16, not the tile height: a ragged last m-block (dense M not
    // a multiple of block_m, the tail of an expert's segment, a short masked group) costs tensor time in proportion to
    // its rows. With a CTA pair each CTA supplies N/2 token rows, so CTA 1 loads from row N/2 of the tile (not
    // block_m/2) and accumulator column j stays token row j. (Shapes are run-time values here; a kernel specialised
    // at compile time on block_m pays for the padding.) MN-major tokens keep the full height (swizzle-atom alignment).
    auto tile_n = [&](const Tile& t) -> uint32_t {
        if (kXMn || kCSplit) return p.block_m;
        return max(16u, min(p.block_m, (t.valid_m + 15u) & ~15u));
    };

    if (warp_idx == 0) {
        // =================================================================== TMA producer (one lane, every CTA)
        // (`elect_one()` right at the branch: ptxas then knows the region is single-threaded and keeps every TMA operand in
        // uniform registers; behind a plain bool it falls back to R2UR.BROADCAST waterfall loops, ~2x slower per k-block)
        if (elect_one()) {
            DGB_STAMP(4);
            Scheduler<kGemmType, kCluster, kSplitK, kCSplit> sched(p, cta_rank, split_rank);
            Tile t;
            const uint32_t ab_bytes = kWTileBytes + x_tile_bytes;
            const uint32_t sfw_tx = kBlockN * 4, sfx_tx = p.block_m * 4;
            constexpr uint32_t kWRows = kBlockN / kPairs;
            uint16_t w_mask = 0;
            for (uint32_t q = 0; q < kPairs; ++q) w_mask |= static_cast<uint16_t>(1u << (2 * q + (cta_rank & 1)));
            uint32_t fresh = num_stages;          // slots never used yet: nothing to wait for (a TRYWAIT costs ~90 cycles)
            uint32_t br_batch = 0, br_k0 = 0;     // kBatchReduce: the producer's position inside the current tile
            uint32_t landed_group = 0xffffffffu;
            while (sched.next(t)) {
                if (kGemmType == kMContiguousPsum && t.valid_m == 0) continue;
                if constexpr (kGemmType == kMContiguousPsum) {
                    // rows of this expert still travelling (dispatch kernels of the peers run concurrently)? wait for
                    // its arrival counter; groups are walked in order, the sources send them in order
                    if (p.arrival != nullptr && sched.g != landed_group) {
                        const uint32_t want = __ldg(p.arrival_expected + sched.g);
                        uint64_t t0 = 0;
                        uint32_t spins = 0, got;
                        while (true) {
                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(p.arrival + sched.g) : "memory");
                            if (static_cast<int32_t>(got - want) >= 0) break;
                            __nanosleep(64);
                            if ((++spins & 0x3FF) == 0) {
                                const uint64_t now = globaltimer_ns();
                                if (t0 == 0) t0 = now;
                                if (now - t0 > kSpinTimeoutNs) asm volatile("trap;");
                            }
                        }
                        asm volatile("fence.proxy.async;" ::: "memory");   // rows were written by generic stores, TMA reads them
                        landed_group = sched.g;
                    }
                }
                const uint32_t x_row = t.x_row + (cta_rank & 1) * (tile_n(t) / kCtaGroup);
                uint32_t k0 = t.kb_begin * kBlockK;
                for (uint32_t kb = t.kb_begin; kb < t.kb_end; ++kb, k0 += kBlockK, ring.advance()) {
                    const uint32_t full = full_bar + ring.bar, slot = smem_base + ring.slot;
                    if (fresh) --fresh; else mbar_wait(empty_bar + ring.bar, ring.phase ^ 1);
                    const bool first = kb == t.kb_begin;
                    if (first) DGB_STAMP(14);
                    const bool load_sfw = !kBf16AB && ((kb & sfw_mask) == 0 || first), load_sfx = !kBf16AB && ((kb & sfx_mask) == 0 || first);
                    mbar_arrive_expect_tx(full, ab_bytes + (load_sfw ? sfw_tx : 0u) + (load_sfx ? sfx_tx : 0u));
                    if constexpr (kGemmType == kBatchReduce) {
                        // virtual k-block -> (batch, byte offset inside the batch's K), kept incrementally
                        if (first) br_batch = t.batch, br_k0 = 0;
                        tma_load_3d(&map_w, full, slot, br_k0, t.w_row, br_batch, p.w_hint);
                        tma_load_3d(&map_x, full, slot + off_x, br_k0, x_row, br_batch, p.x_hint);
                        br_k0 += kBlockK;
                        if (br_k0 >= p.k) br_k0 = 0, ++br_batch;
                    } else if constexpr (kGemmType == kBatched) {
                        // 3-D maps {inner, outer, batch}: boxes are one batch deep, so the tiles land exactly like 2-D ones
                        if constexpr (kWMn) {
                            for (uint32_t j = 0; j < kEl; ++j)
                                tma_load_3d(&map_w, full, slot + j * (kKRows * 128), t.n0 * kEl + j * 128, k0 / kEl, t.batch, p.w_hint);
                        } else
                            tma_load_3d(&map_w, full, slot, k0, t.w_row, t.batch, p.w_hint);
                        if constexpr (kXMn) {
                            for (uint32_t i = 0, off = 0; i < load_m * kEl; i += p.x_swizzle, off += p.x_swizzle * kKRows)
                                tma_load_3d(&map_x, full, slot + off_x + off, x_row * kEl + i, k0 / kEl, t.batch, p.x_hint);
                        } else {
                            tma_load_3d(&map_x, full, slot + off_x, k0, x_row, t.batch, p.x_hint);
                        }
                    } else {
                    if constexpr (kWMn) {
                        // MN-major: the contiguous (byte-addressed) coordinate is N, the K rows are counted in elements
                        for (uint32_t j = 0; j < kEl; ++j)
                            tma_load_2d(&map_w, full, slot + j * (kKRows * 128), t.n0 * kEl + j * 128, (t.wk_base + t.k_base + k0) / kEl, p.w_hint);
                    } else if constexpr (kPairs == 1) {
                        tma_load_2d(&map_w, full, slot, t.k_base + k0, t.w_row, p.w_hint);
                    } else {
                        tma_load_2d_multicast(&map_w, full, slot + pair_idx * (kWRows * kBlockK), k0, t.w_row + pair_idx * kWRows,
                                              w_mask, p.w_hint);
                    }
                    if constexpr (kXMn) {
                        for (uint32_t i = 0, off = 0; i < load_m * kEl; i += p.x_swizzle, off += p.x_swizzle * kKRows)
                            tma_load_2d(&map_x, full, slot + off_x + off, x_row * kEl + i, (t.k_base + k0) / kEl, p.x_hint);
                    } else {
                        tma_load_2d(&map_x, full, slot + off_x, t.k_base + k0, x_row, p.x_hint);
                    }
                    }
                    if (load_sfw) tma_load_2d(&map_sfw, full, slot + off_sfw, t.sfw_col, t.sfw_row + (kb >> p.sf_shift_w), kEvictNormal);
                    if (load_sfx) tma_load_2d(&map_sfx, full, slot + off_sfx, t.sfx_col, t.sfx_row + (kb >> p.sf_shift_x), kEvictNormal);
                    if (first) DGB_STAMP(2);
                }
            }
        }
        if constexpr (kCluster > 1 && kEarlyProducer) {
            __syncwarp();
            cluster_wait();        // collect the prologue's cluster barrier phase (long complete)
        }
    } else if (warp_idx == 1) {
        // =================================================================== MMA issuer (leader CTA only)
        if (is_leader) {
            Scheduler<kGemmType, kCluster, kSplitK, kCSplit> sched(p, cta_rank, split_rank);
            Tile t;
            uint32_t idesc_base = 0;     // per tile: UMMA M = 128 x CTAs, N = tile_n(t)
            // descriptors of slot 0; a slot offset adds (bytes >> 4) to the 14-bit start-address field.
            //   K-major : 8-row x 128 B swizzle atoms stacked along MN (SBO 1024); +32 B per UMMA_K step
            //   MN-major: atoms of S bytes (MN) x 8 K-rows; SBO = 8*S between K groups, LBO = kKRows*S between MN atoms;
            //             + kUmmaKRows K-rows = kUmmaKRows*S bytes per UMMA_K step   (cf. reference mma/sm100.cuh:96-132)
            const uint32_t xs = kXMn ? p.x_swizzle : 128u;
            const uint32_t x_layout = xs == 128 ? kLayoutSwizzle128B : (xs == 64 ? kLayoutSwizzle64B : kLayoutSwizzle32B);
            const uint64_t w_desc0 = kWMn ? make_smem_desc(smem_base, kKRows * 128, 8 * 128, kLayoutSwizzle128B)
                                          : make_smem_desc(smem_base, 0, 1024, kLayoutSwizzle128B);
            const uint64_t x_desc0 = kXMn ? make_smem_desc(smem_base + off_x, kKRows * xs, 8 * xs, x_layout)
                                          : make_smem_desc(smem_base + off_x, 0, 1024, kLayoutSwizzle128B);
            const uint32_t w_kstep = kWMn ? (kUmmaKRows * 128) >> 4 : kUmmaK >> 4;
            const uint32_t x_kstep = kXMn ? (kUmmaKRows * xs) >> 4 : kUmmaK >> 4;
            const uint64_t sfw_desc0 = make_smem_desc(smem_base + off_sfw, 0, 128, kLayoutNoSwizzle);
            const uint64_t sfx_desc0 = make_smem_desc(smem_base + off_sfx, 0, 128, kLayoutNoSwizzle);
            const uint32_t tmem_sfw = tmem_base + kTmemColSFW, tmem_sfx = tmem_base + kTmemColSFX;
            const uint16_t pair_mask = static_cast<uint16_t>(0b11u << leader_rank);           // this pair only
            // `empty` barriers the retiring MMAs release: every CTA that holds operands of them (the whole cluster when pairs
            // share multicast weight tiles, else this MMA group)
            const uint16_t kEmptyMask = kCSplit ? pair_mask : static_cast<uint16_t>((1u << kCluster) - 1);
            // The issue loop below bounds every shape whose tiles are small (each k-block then costs its ~50
            // instructions, not its MMA time), so everything loop-invariant is hoisted: per-k-block work is one barrier
            // wait, <= 3 tcgen05.cp, 4 tcgen05.mma whose descriptors differ by immediates, and one commit.
            const uint32_t id_kb_mul = (sfw_mask ? (1u << 29) : 0u) | (sfx_mask ? (1u << 4) : 0u);   // gran_k 128: id = kb & 3
            const uint32_t id_j_mul = (sfw_mask ? 0u : (1u << 29)) | (sfx_mask ? 0u : (1u << 4));    // gran_k 32 : id = j
            const uint32_t sub_mask = sfw_mask | sfx_mask;                                          // 3 or 0
            auto issue_kblock = [&](auto n_const, uint32_t kb, bool first, uint32_t tmem_d) {
                constexpr uint32_t kNumUmma = decltype(n_const)::value;
                const uint32_t slot16 = ring.slot >> 4;
                if constexpr (!kBf16AB) {
                    if ((kb & sfw_mask) == 0 || first) tmem_cp_sf<kCtaGroup>(tmem_sfw, sfw_desc0 + slot16);
                    if ((kb & sfx_mask) == 0 || first) {
                        tmem_cp_sf<kCtaGroup>(tmem_sfx, sfx_desc0 + slot16);
                        if (num_sfx_groups > 1) tmem_cp_sf<kCtaGroup>(tmem_sfx + 4, sfx_desc0 + slot16 + 32);
                    }
                }
                const uint64_t w_desc = w_desc0 + slot16, x_desc = x_desc0 + slot16;
                const uint32_t idesc = idesc_base + (kb & sub_mask) * id_kb_mul;   // one UE8M0 byte per 32 K-elements
#pragma unroll
                for (uint32_t j = 0; j < kNumUmma; ++j) {
                    if constexpr (kBf16AB)
                        mma_f16<kCtaGroup>(tmem_d, w_desc + j * w_kstep, x_desc + j * x_kstep, idesc_base, (j != 0 || !first) ? 1u : 0u);
                    else
                        mma_mxf8_block_scale<kCtaGroup>(tmem_d, w_desc + j * w_kstep, x_desc + j * x_kstep, idesc + j * id_j_mul,
                                                       tmem_sfw, tmem_sfx, (j != 0 || !first) ? 1u : 0u);
                }
                // retire -> the smem slot may be overwritten (signals every CTA of the cluster)
                mma_commit<kCtaGroup>(empty_bar + ring.bar, kEmptyMask);
            };
            uint32_t tile_iter = 0;
            while (sched.next(t)) {
                if (kGemmType == kMContiguousPsum && t.valid_m == 0) continue;
                const uint32_t as = tile_iter & 1, aphase = (tile_iter >> 1) & 1;
                ++tile_iter;
                mbar_wait(tmem_empty_bar + as * 8, aphase ^ 1);
                tcgen05_fence_after();
                idesc_base = kBf16AB ? make_idesc_bf16(128 * kCtaGroup, tile_n(t), kWMn ? 1 : 0, kXMn ? 1 : 0) : make_idesc(128 * kCtaGroup, tile_n(t), kWMn ? 1 : 0, kXMn ? 1 : 0);
                const uint32_t tmem_d = tmem_base + as * kAccumColStride;
                uint32_t kb = t.kb_begin;
                for (; kb + 1 < t.kb_end; ++kb, ring.advance()) {        // every k-block but the last: 4 UMMAs
                    mbar_wait(ready_bar + ring.bar, ring.phase);
                    tcgen05_fence_after();
                    if (elect_one()) issue_kblock(std::integral_constant<uint32_t, 4>{}, kb, kb == t.kb_begin, tmem_d);
                    __syncwarp();
                }
                mbar_wait(ready_bar + ring.bar, ring.phase);               // last k-block: K may end inside it
                tcgen05_fence_after();
                if (elect_one()) {
                    const bool first = kb == t.kb_begin;
                    switch (t.last_umma) {
                        case 1: issue_kblock(std::integral_constant<uint32_t, 1>{}, kb, first, tmem_d); break;
                        case 2: issue_kblock(std::integral_constant<uint32_t, 2>{}, kb, first, tmem_d); break;
                        case 3: issue_kblock(std::integral_constant<uint32_t, 3>{}, kb, first, tmem_d); break;
                        default: issue_kblock(std::integral_constant<uint32_t, 4>{}, kb, first, tmem_d); break;
                    }
                    mma_commit<kCtaGroup>(tmem_full_bar + as * 8, pair_mask);   // accumulator complete -> epilogue
                }
                __syncwarp();
                ring.advance();
                if (tile_iter == 1 && lane == 0) DGB_STAMP(5);
            }
            // Drain: nobody may tear the CTA pair down while epilogue threads of the peer still arrive here
            if (tile_iter > 0) {
                const uint32_t last = tile_iter - 1;
                mbar_wait(tmem_empty_bar + (last & 1) * 8, (last >> 1) & 1);
            }
        }
    } else if (warp_idx == 2) {
        // =================================================================== SF re-tiler / slot forwarder
        Scheduler<kGemmType, kCluster, kSplitK, kCSplit> sched(p, cta_rank, split_rank);
        Tile t;
        // tcgen05.cp 32x128b wants word (row r of the 128-group) at [r % 32][r / 32]; TMA delivered it at [r]
        auto retile = [&](uint32_t base) {
            const uint32_t src = base + lane * 4;
            const uint32_t v0 = ld_shared_u32(src), v1 = ld_shared_u32(src + 128), v2 = ld_shared_u32(src + 256),
                           v3 = ld_shared_u32(src + 384);
            __syncwarp();
            st_shared_v4(base + lane * 16, v0, v1, v2, v3);
        };
        // barrier of the pair's leader CTA, as a shared::cluster address
        const uint32_t ready_dst = kCtaGroup > 1 ? mapa(ready_bar, leader_rank) : ready_bar;
        while (sched.next(t)) {
            if (kGemmType == kMContiguousPsum && t.valid_m == 0) continue;
            for (uint32_t kb = t.kb_begin; kb < t.kb_end; ++kb, ring.advance()) {
                mbar_wait(full_bar + ring.bar, ring.phase);
                const bool first = kb == t.kb_begin;
                if (first && lane == 0) DGB_STAMP(3);
                const bool do_w = !kBf16AB && ((kb & sfw_mask) == 0 || first), do_x = !kBf16AB && ((kb & sfx_mask) == 0 || first);
                if (do_w | do_x) {
                    const uint32_t slot = smem_base + ring.slot;
                    if (do_w) retile(slot + off_sfw);
                    if (do_x) {
                        retile(slot + off_sfx);
                        if (num_sfx_groups > 1) retile(slot + off_sfx + 512);
                    }
                    fence_proxy_async_smem();   // generic-proxy writes -> visible to tcgen05.cp
                }
                if constexpr (kCtaGroup > 1)
                    mbar_arrive_remote(ready_dst + ring.bar);
                else
                    mbar_arrive(ready_dst + ring.bar);
            }
        }
    } else if (warp_idx >= 4) {
        // =================================================================== epilogue: TMEM -> registers -> global
        Scheduler<kGemmType, kCluster, kSplitK, kCSplit> sched(p, cta_rank, split_rank);
        Tile t;
        const uint32_t quad = warp_idx & 3;                 // TMEM lane quadrant this warp may read
        const uint32_t half = (warp_idx - 4) >> 2;          // 0: chunks 0,2,4.. | 1: chunks 1,3,5..
        out_t* d = reinterpret_cast<out_t*>(p.d);
        const uint32_t tmem_empty_dst = kCtaGroup > 1 ? mapa(tmem_empty_bar, leader_rank) : tmem_empty_bar;
        const size_t row_bytes = static_cast<size_t>(p.ld_d) * sizeof(out_t);
        uint32_t tile_iter = 0;
        [[maybe_unused]] uint32_t store_iter = 0;              // kTmaStore: units this group has staged so far

        while (sched.next(t)) {
            const uint32_t n = t.n0 + quad * 32 + lane;
            const bool n_ok = n < p.n;
            const uint32_t n_store = p.head_mid ? n + (n + p.head_right) / p.head_lr * p.head_mid : n;   // head-split remap
            char* d_col = reinterpret_cast<char*>(d + static_cast<size_t>(t.batch) * p.d_batch_stride +
                                                  static_cast<size_t>(t.d_row) * p.ld_d + n_store);
            if (kGemmType == kMContiguousPsum && t.valid_m == 0) {
                if (n_ok && half == 0)
                    for (uint32_t r = 0; r < t.store_m; ++r) store_out<out_t>(reinterpret_cast<out_t*>(d_col + r * row_bytes), 0.0f, false);
                continue;
            }
            const uint32_t as = tile_iter & 1, aphase = (tile_iter >> 1) & 1;
            ++tile_iter;
            mbar_wait(tmem_full_bar + as * 8, aphase);
            tcgen05_fence_after();
            if (threadIdx.x == 128) DGB_STAMP(6);
            const uint32_t taddr = tmem_base + ((quad * 32) << 16) + as * kAccumColStride;
            const uint32_t load_cols = (max(t.valid_m, 1u) + 15) / 16 * 16;
            if constexpr (kCSplit) {
                // ---------------------------------------------------------------- cluster split-K epilogue
                // The tile's token columns are cut into kCluster chunks of `c` columns; CTA r owns chunk r. Work unit =
                // one 16-column piece (one TMEM load); the two warps of a lane quadrant take alternate pieces.
                //   foreign piece: TMEM -> registers -> local outbox -> bulk copy into the owner's staging buffer
                //   own piece    : (after the peers' bytes have landed) partials added in slice order -> D
                // Staging of owner q: float4 [source slot][c/4 column quads][128 weight rows]; a warp writes / reads 512
                // contiguous bytes per quad.
                constexpr uint32_t S = kCSplit;
                const uint32_t c = p.block_m / S, pieces_per_chunk = c / 16;
                const uint32_t chunk_bytes = c * 128 * 4;                      // one source's partial of one chunk
                auto peer_of = [&](uint32_t q) { return q * kCtaGroup + cta_rank; };   // cluster rank of slice q's CTA that holds my weight rows
                if (threadIdx.x == 4 * 32) mbar_arrive_expect_tx(red_bar, (S - 1) * chunk_bytes);
                const uint32_t row_n = quad * 32 + lane;                       // weight row of this thread inside the tile
                const uint32_t num_pieces = p.block_m / 16;
                // Foreign pieces go to a local outbox first (the stage ring is idle by now: this CTA's only tile has been
                // consumed), laid out exactly like the owner's staging slot, and travel as ONE bulk copy per owner:
                // SM-issued remote stores top out near 20 B/clk, the copy engine does not occupy the epilogue warps.
                for (uint32_t i = half; i < num_pieces; i += 2) {
                    const uint32_t q = i / pieces_per_chunk;
                    if (q == split_rank) continue;
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(taddr + i * 16, v);
                    tmem_ld_wait();
                    const uint32_t pi = i - q * pieces_per_chunk;
                    const uint32_t dst = smem_base + q * chunk_bytes + ((pi * 4) * 128 + row_n) * 16;
#pragma unroll
                    for (uint32_t g = 0; g < 4; ++g)
                        st_shared_f4(dst + g * 128 * 16, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                }
                fence_proxy_async_smem();                                     // generic writes -> visible to the copy engine
                if (threadIdx.x == 4 * 32) DGB_STAMP(10);
                named_bar_sync(1, kNumEpilogueThreads);
                if (threadIdx.x == 4 * 32) DGB_STAMP(11);
                if (threadIdx.x == 4 * 32) {
#pragma unroll
                    for (uint32_t q = 0; q < S; ++q) {
                        if (q == split_rank) continue;
                        const uint32_t slot = split_rank < q ? split_rank : split_rank - 1;
                        bulk_copy_to_peer(mapa(red_stage + slot * chunk_bytes, peer_of(q)), smem_base + q * chunk_bytes, chunk_bytes,
                                          mapa(red_bar, peer_of(q)));
                    }
                }
                // My own chunk in units of 8 token columns (the two warps of a quadrant alternate): own partial from TMEM,
                // the peers' partials from the staging buffer once they have landed, added in slice order, written to D.
                // (One tile per cluster: nothing waits for this accumulator, so it is read where it is needed.)
                const uint32_t units = c / 8;
                if (threadIdx.x == 4 * 32) DGB_STAMP(12);
                mbar_wait(red_bar, 0);
                if (threadIdx.x == 4 * 32) DGB_STAMP(13);
                for (uint32_t u = half; u < units; u += 2) {
                    uint32_t own[8];
                    tmem_ld_32x32b_x8(taddr + split_rank * c + u * 8, own);
                    tmem_ld_wait();
                    float acc[8];
#pragma unroll
                    for (uint32_t s = 0; s < S; ++s) {                          // slice order: deterministic
                        if (s == split_rank) {
#pragma unroll
                            for (uint32_t j = 0; j < 8; ++j)
                                acc[j] = s == 0 ? __uint_as_float(own[j]) : acc[j] + __uint_as_float(own[j]);
                        } else {
                            const uint32_t slot = s < split_rank ? s : s - 1;
                            const uint32_t src = red_stage + slot * chunk_bytes + ((u * 2) * 128 + row_n) * 16;
                            const float4 x0 = ld_shared_f4(src), x1 = ld_shared_f4(src + 128 * 16);
                            if (s == 0) {
                                acc[0] = x0.x, acc[1] = x0.y, acc[2] = x0.z, acc[3] = x0.w;
                                acc[4] = x1.x, acc[5] = x1.y, acc[6] = x1.z, acc[7] = x1.w;
                            } else {
                                acc[0] += x0.x, acc[1] += x0.y, acc[2] += x0.z, acc[3] += x0.w;
                                acc[4] += x1.x, acc[5] += x1.y, acc[6] += x1.z, acc[7] += x1.w;
                            }
                        }
                    }
                    const uint32_t r0 = split_rank * c + u * 8;               // first token row of this unit in the tile
                    if (n_ok && r0 < t.valid_m) {
                        char* row = d_col + static_cast<size_t>(r0) * row_bytes;
                        store_rows<out_t, 8, kAccumulate>(row, row_bytes, acc, t.valid_m - r0);
                    }
                }
                tcgen05_fence_before();
                if constexpr (kCtaGroup > 1)
                    mbar_arrive_remote(tmem_empty_dst + as * 8);             // last TMEM read of this tile
                else
                    mbar_arrive(tmem_empty_dst + as * 8);
                continue;
            }
            if constexpr (kSplitK) {
                // ---------------------------------------------------------------- split-K epilogue
                // (1) park this slice's FP32 partial tile in the workspace, (2) count arrivals per output block,
                // (3) the last slice to arrive adds the partials in slice order (deterministic) and writes D.
                float* ws_col = p.splitk_ws + (static_cast<size_t>(t.split) * p.m + t.d_row) * p.n + n;
                auto release = [&]() {
                    tcgen05_fence_before();
                    if constexpr (kCtaGroup > 1)
                        mbar_arrive_remote(tmem_empty_dst + as * 8);
                    else
                        mbar_arrive(tmem_empty_dst + as * 8);
                };
                if (half * 32 >= load_cols) release();
                for (uint32_t c0 = half * 32; c0 < load_cols; c0 += 64) {
                    uint32_t v[32];
                    const bool second = c0 + 16 < load_cols;
                    tmem_ld_32x32b_x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                    if (second) tmem_ld_32x32b_x16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                    tmem_ld_wait();
                    if (c0 + 64 >= load_cols) release();
                    if (n_ok) {
                        float* wrow = ws_col + static_cast<size_t>(c0) * p.n;
#pragma unroll
                        for (uint32_t h = 0; h < 2; ++h) {
                            const uint32_t r0 = c0 + h * 16;
                            if (r0 + 16 <= t.valid_m) {
#pragma unroll
                                for (uint32_t j = 0; j < 16; ++j) wrow[static_cast<size_t>(h * 16 + j) * p.n] = __uint_as_float(v[h * 16 + j]);
                            } else if (r0 < t.valid_m) {
#pragma unroll
                                for (uint32_t j = 0; j < 16; ++j)
                                    if (r0 + j < t.valid_m) wrow[static_cast<size_t>(h * 16 + j) * p.n] = __uint_as_float(v[h * 16 + j]);
                            }
                        }
                    }
                }
                // Publish: the CTA barrier orders every thread's partial stores before thread 0, whose gpu-scope fence
                // (cumulative) + counter increment release them; the increment's result tells who arrived last.
                const uint32_t et = threadIdx.x - 4 * 32;          // index among the epilogue threads
                named_bar_sync(1, kNumEpilogueThreads);
                if (et == 0) {
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");
                    const int prev = atomicAdd(p.splitk_counters + t.counter_idx, 1);
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(splitk_flag_smem), "r"(prev) : "memory");
                }
                named_bar_sync(1, kNumEpilogueThreads);
                const uint32_t prev = ld_shared_u32(splitk_flag_smem);
                if (prev == p.num_splits - 1) {
                    const uint32_t nc = t.n0 + lane * 4;           // 4 consecutive columns per lane, one row per warp
                    if (nc < p.n) {
                        const size_t slice = static_cast<size_t>(p.m) * p.n;
                        constexpr uint32_t kWarps = kNumEpilogueThreads / 32;
#define DGB_FINALIZE(S, R) \
    splitk_finalize<S, R, out_t, kAccumulate>(p.splitk_ws, slice, p.n, d, p.ld_d, t.d_row, t.valid_m, nc, et >> 5, kWarps)
                        switch (p.num_splits) {
                            case 2: DGB_FINALIZE(2, 4); break;
                            case 3: DGB_FINALIZE(3, 4); break;
                            case 4: DGB_FINALIZE(4, 4); break;
                            case 5: DGB_FINALIZE(5, 2); break;
                            case 6: DGB_FINALIZE(6, 2); break;
                            case 7: DGB_FINALIZE(7, 2); break;
                            default: DGB_FINALIZE(8, 2); break;
                        }
#undef DGB_FINALIZE
                    }
                    if (et == 0) p.splitk_counters[t.counter_idx] = 0;   // leave the counters clean for the next launch
                }
                named_bar_sync(1, kNumEpilogueThreads);                   // the flag word is reused by the next tile
                continue;
            }
            auto release_accumulator = [&]() {
                // last read of this accumulator buffer by this warp: hand it back before the stores drain
                tcgen05_fence_before();
                if constexpr (kCtaGroup > 1)
                    mbar_arrive_remote(tmem_empty_dst + as * 8);
                else
                    mbar_arrive(tmem_empty_dst + as * 8);
            };
            if constexpr (kSwapD && kTmaStore) {
                // ---------------------------------------------------------------- transposed output, staged through shared memory
                // lane = output row, TMEM column = output column. Every warp turns its own 32 rows x 32 columns around in a
                // private 2 KB buffer: each thread writes the 64 bytes of ITS row (4 x 16 B, XOR-swizzled), then lane L reads
                // the 16 bytes (row 8i + L/4, piece L%4) back and the warp stores 8 rows x 64 contiguous bytes per instruction
                // -- no cross-warp synchronisation, both shared-memory phases conflict-free. The two warps of a lane quadrant
                // take alternate 32-column units. (Per-warp TMA stores out of the same buffers cost 5.5 us of a 43.7 us launch
                // at 4096 x 7168 x 2048 -- 38.2 with the stores compiled out --, 64-column units 7.8 us: kept out.)
                const uint32_t num_units = (load_cols + kSwapStoreCols - 1) / kSwapStoreCols;
                const uint32_t buf = staging + (warp_idx - 4) * kSwapStoreBufBytes;
                const uint32_t row_off = buf + lane * 64, sw = (lane >> 1) & 3;
                const uint32_t rd_row = lane >> 2, rd_piece = lane & 3;                // read phase: row 8i + rd_row, 16-byte piece rd_piece
                const uint32_t rd_off = buf + rd_row * 64 + ((rd_piece ^ ((rd_row >> 1) & 3)) << 4);   // (+ 512 i: the swizzle term repeats every 8 rows)
                const uint32_t tok0 = t.n0 + quad * 32 + rd_row;
                if (half >= num_units) release_accumulator();
                for (uint32_t u = half; u < num_units; u += 2) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x16(taddr + u * kSwapStoreCols, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                    tmem_ld_32x32b_x16(taddr + u * kSwapStoreCols + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                    tmem_ld_wait();
                    if (u + 2 >= num_units) release_accumulator();
#pragma unroll
                    for (uint32_t piece = 0; piece < 4; ++piece)
                        st_shared_v4(row_off + ((piece ^ sw) << 4), pack_bf16x2(v[8 * piece + 0], v[8 * piece + 1]),
                                     pack_bf16x2(v[8 * piece + 2], v[8 * piece + 3]), pack_bf16x2(v[8 * piece + 4], v[8 * piece + 5]),
                                     pack_bf16x2(v[8 * piece + 6], v[8 * piece + 7]));
                    __syncwarp();
                    const uint32_t col = t.d_row + u * kSwapStoreCols + rd_piece * 8;   // output column of this lane's piece (D has a multiple of 8 columns)
                    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(d) + static_cast<size_t>(tok0) * p.ld_d + col;
#pragma unroll
                    for (uint32_t i = 0; i < 4; ++i) {
                        const uint4 x = ld_shared_u4(rd_off + i * 512);
                        if (tok0 + 8 * i < p.n && col < p.m) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(8 * i) * p.ld_d) = x;
                    }
                    __syncwarp();                                    // the buffer is rewritten by the next unit
                }
                continue;
            }
            if constexpr (kSwapD) {
                // ---------------------------------------------------------------- transposed-output epilogue
                // lane = output row (token), TMEM column = output column (weight): 16 consecutive columns per load, written
                // as 16-byte pieces when D allows it (base and row pitch multiples of 16 bytes; tile origins are multiples of
                // 16 columns), element by element on ragged edges and when accumulating into C.
                out_t* d_row = d + static_cast<size_t>(n) * p.ld_d + t.d_row;
                const bool vec_ok = !kAccumulate && ((reinterpret_cast<uintptr_t>(d) | (static_cast<uintptr_t>(p.ld_d) * sizeof(out_t))) & 15) == 0;
                if (half * 32 >= load_cols) release_accumulator();
                for (uint32_t c0 = half * 32; c0 < load_cols; c0 += 64) {
                    uint32_t v[32];
                    const bool second = c0 + 16 < load_cols;
                    tmem_ld_32x32b_x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                    if (second) tmem_ld_32x32b_x16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                    tmem_ld_wait();
                    if (c0 + 64 >= load_cols) release_accumulator();
                    if (n_ok) {
#pragma unroll
                        for (uint32_t h = 0; h < 2; ++h) {
                            const uint32_t r0 = c0 + h * 16;
                            out_t* dst = d_row + r0;
                            if (r0 + 16 <= t.valid_m && vec_ok) {
                                if constexpr (std::is_same_v<out_t, __nv_bfloat16>) {
                                    uint4 lo, hi;
                                    lo.x = pack_bf16x2(v[h * 16 + 0], v[h * 16 + 1]), lo.y = pack_bf16x2(v[h * 16 + 2], v[h * 16 + 3]);
                                    lo.z = pack_bf16x2(v[h * 16 + 4], v[h * 16 + 5]), lo.w = pack_bf16x2(v[h * 16 + 6], v[h * 16 + 7]);
                                    hi.x = pack_bf16x2(v[h * 16 + 8], v[h * 16 + 9]), hi.y = pack_bf16x2(v[h * 16 + 10], v[h * 16 + 11]);
                                    hi.z = pack_bf16x2(v[h * 16 + 12], v[h * 16 + 13]), hi.w = pack_bf16x2(v[h * 16 + 14], v[h * 16 + 15]);
                                    reinterpret_cast<uint4*>(dst)[0] = lo;
                                    reinterpret_cast<uint4*>(dst)[1] = hi;
                                } else {
#pragma unroll
                                    for (uint32_t q4 = 0; q4 < 4; ++q4)
                                        reinterpret_cast<uint4*>(dst)[q4] = make_uint4(v[h * 16 + 4 * q4], v[h * 16 + 4 * q4 + 1],
                                                                                       v[h * 16 + 4 * q4 + 2], v[h * 16 + 4 * q4 + 3]);
                                }
                            } else if (r0 < t.valid_m) {
#pragma unroll
                                for (uint32_t j = 0; j < 16; ++j)
                                    if (r0 + j < t.valid_m) store_out<out_t>(dst + j, __uint_as_float(v[h * 16 + j]), kAccumulate);
                            }
                        }
                    }
                }
                continue;
            }
            if constexpr (kTmaStore) {
                // ---------------------------------------------------------------- staged TMA-store epilogue
                // Work unit = 16 token rows x this CTA's 128 weight rows = one 4 KB staging buffer (two 128B-swizzled boxes of
                // 16 rows x 64 columns). All eight warps work on a unit: the two warps of a lane quadrant take its upper /
                // lower 8 token rows.
                //   tcgen05.ld 16x256b: thread t gets (TMEM lane t/4 (+8), columns 2(t%4), +1) = the stmatrix fragment, so
                //   one stmatrix.x4.trans writes 8 token rows x 32 weight columns as 16-byte pieces into the swizzled box.
                const uint32_t num_units = load_cols / kStoreRows;
                const uint32_t frag_row = lane & 7, frag_piece = (quad & 1) * 4 + (lane >> 3);
                const uint32_t frag_off = (quad >> 1) * (kStoreBufBytes / 2) + (half * 8 + frag_row) * 128 + ((frag_piece ^ frag_row) << 4);
                const bool issuer_warp = warp_idx == 4;
                for (uint32_t u = 0; u < num_units; ++u, ++store_iter) {
                    const uint32_t buf = staging + (store_iter & 1) * kStoreBufBytes;
                    if (issuer_warp) tma_store_wait_read<1>();     // the store that last used this buffer has read it out
                    named_bar_sync(3, kNumEpilogueThreads);
                    uint32_t v[8];
                    const uint32_t ta = taddr + u * kStoreRows + half * 8;
                    tmem_ld_16x256b(ta, &v[0]);
                    tmem_ld_16x256b(ta + (16u << 16), &v[4]);
                    tmem_ld_wait();
                    if (u + 1 == num_units) release_accumulator();
                    stmatrix_x4_trans(buf + frag_off, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                      pack_bf16x2(v[6], v[7]));
                    fence_proxy_async_smem();                      // generic writes -> visible to the TMA engine
                    named_bar_sync(3, kNumEpilogueThreads);
                    if (issuer_warp && lane == 0) {
                        const uint32_t row = t.d_row + u * kStoreRows;
                        if (t.n0 < p.n) tma_store_2d(&map_d, buf, t.n0, row);
                        if (t.n0 + 64 < p.n) tma_store_2d(&map_d, buf + kStoreBufBytes / 2, t.n0 + 64, row);
                        tma_store_commit();
                    }
                }
                if (num_units == 0) release_accumulator();
                continue;
            }
            if (half * 32 >= load_cols) release_accumulator();   // nothing to read for this warp
            // 32 token rows per iteration: two TMEM loads in flight, then 32 row stores (one instruction each,
            // 32 consecutive columns per warp). Full chunks take the branch-free path.
            for (uint32_t c0 = half * 32; c0 < load_cols; c0 += 64) {
                uint32_t v[32];
                const bool second = c0 + 16 < load_cols;
                tmem_ld_32x32b_x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                if (second) tmem_ld_32x32b_x16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                tmem_ld_wait();
                if (c0 + 64 >= load_cols) release_accumulator();
                char* row = d_col + static_cast<size_t>(c0) * row_bytes;
                if (n_ok) {
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h) {       // full 16-row halves take the path without per-row predicates
                        const uint32_t r0 = c0 + h * 16;
                        if (r0 < t.valid_m) store_rows<out_t, 16, kAccumulate>(row + h * 16 * row_bytes, row_bytes, &v[h * 16], t.valid_m - r0);
                    }
                }
            }
            // psum layout with zero padding: rows between the group's end and its aligned end are defined to be 0
            if (n_ok && half == 0)
                for (uint32_t r = t.valid_m; r < t.store_m; ++r)
                    store_out<out_t>(reinterpret_cast<out_t*>(d_col + r * row_bytes), 0.0f, false);
        }
        if constexpr (kTmaStore) {
            if (!kSwapD && warp_idx == 4) tma_store_wait_all();   // the staging buffers must outlive every store that reads them
        }
    }

    // ---- teardown
    if (threadIdx.x == 128) DGB_STAMP(7);
    if (threadIdx.x == 0) DGB_STAMP(8);
    __syncwarp();
    tcgen05_fence_before();
    if constexpr (kCluster > 1) {
        cluster_arrive_relaxed();   // (a release-arrive here would wait for every output store to become visible)
        cluster_wait();
    } else {
        __syncthreads();
    }
    if (warp_idx == 2) tmem_dealloc<kCtaGroup>(tmem_base, kTmemCols);
    if (threadIdx.x == 0) DGB_STAMP(9);
    if (threadIdx.x == 0 && p.debug_ts != nullptr) p.debug_ts[16 + 2 * blockIdx.x + 1] = globaltimer_ns();  // per-CTA exit
#endif
}

}  // namespace dgb200
