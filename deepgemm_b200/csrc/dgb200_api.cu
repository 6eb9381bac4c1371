// Host side of the dgb200 C ABI (include/dgb200.h): argument checks, tile/pipeline heuristics, tensor-map
// construction and kernel launch. No torch, no JIT: the kernels in this translation unit are compiled ahead of
// time for sm_100a and every shape parameter is a run-time argument.
//
// Replaces (reference file:line): csrc/apis/gemm.hpp:73-346 (checks), csrc/jit_kernels/heuristics/sm100.hpp
// (config choice), csrc/jit_kernels/impls/runtime_utils.hpp:113-267 (CUtensorMap builders),
// csrc/jit_kernels/impls/sm100_fp8_fp4_gemm_1d1d.hpp:93-391 (launchers), csrc/jit/handle.hpp:174-219 (launch attrs),
// csrc/jit/device_runtime.hpp:14-134 (device property cache + knobs).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>   // environ
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "launch.cuh"
#include "sf_layout.cuh"
#include "ep_dispatch.cuh"
#include "quant.cuh"
#include "peak_probe.cuh"

namespace dgb200 {
namespace {

// ------------------------------------------------------------------------------------------------ errors
thread_local std::string g_last_error;
thread_local dgb200_config g_last_config = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
std::atomic<int64_t> g_launch_count{0};
std::atomic<long long*> g_debug_ts{nullptr};

#define fail host_fail

// ------------------------------------------------------------------------------------------------ runtime state
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Runtime {
    std::mutex mu;
    bool device_ready = false;
    int device = -1;
    int sm_count = 0;
    int smem_optin = 0;
    int num_sms = 0;  // user override (0 = all)
    int tc_util = 100;
    int pdl = 0;
    int mk_alignment = 128;  // HeuristicsRuntime::kLegacyMKAlignmentForContiguousLayout
    int split_k = 1;         // 0: never cut K (bit-identical to the reference's K order); 1: heuristics may
    EncodeTiledFn encode = nullptr;
};
Runtime& rt() {
    static Runtime r;
    return r;
}

int ensure_device() {
    Runtime& r = rt();
    std::lock_guard<std::mutex> lock(r.mu);
    int dev = 0;
    DGB_CUDA(cudaGetDevice(&dev));
    if (r.device_ready && dev == r.device) return DGB200_OK;
    cudaDeviceProp prop;
    DGB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(DGB200_ERR_UNSUPPORTED, "dgb200 kernels are built for sm_100a only; device %d is sm_%d%d", dev,
                    prop.major, prop.minor);
    r.device = dev;
    r.sm_count = prop.multiProcessorCount;
    r.smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    if (!r.encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        DGB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || fn == nullptr)
            return fail(DGB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        r.encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    r.device_ready = true;
    return DGB200_OK;
}

int effective_num_sms() {
    Runtime& r = rt();
    int n = r.num_sms > 0 ? std::min(r.num_sms, r.sm_count) : r.sm_count;
    return n >= 2 ? (n & ~1) : n;  // CTA pairs need an even grid (a budget of 1 SM runs single-CTA MMAs)
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int align_up(int a, int b) { return ceil_div(a, b) * b; }
// Developer knobs are environment variables read per call. A launch consults ~20 of them, and every
// getenv() walks the whole environment: several microseconds of a host path that bounds small-M GEMMs in plain stream
// order. So the public entry points open an EnvScope: ONE pass over `environ` finds out whether any DGB200_* variable is
// set at all (in production: none), and the lookups below answer "unset" without searching.
thread_local int t_env_state = 0;           // 0: no scope open (plain getenv) | 1: scope, no DGB200_* variable | 2: scope, some are set
struct EnvScope {
    int saved;
    EnvScope() : saved(t_env_state) {
        if (saved != 0) return;              // nested: the outer scope has looked already
        int state = 1;
        for (char** e = ::environ; e != nullptr && *e != nullptr; ++e)
            if ((*e)[0] == 'D' && strncmp(*e, "DGB200_", 7) == 0) {
                state = 2;
                break;
            }
        t_env_state = state;
    }
    ~EnvScope() { t_env_state = saved; }
};
inline const char* env_str(const char* name) { return t_env_state == 1 ? nullptr : getenv(name); }
inline int env_int(const char* name, int dflt) {
    const char* v = env_str(name);
    return v && *v ? atoi(v) : dflt;
}

// ------------------------------------------------------------------------------------------------ tensor maps
struct MapKey {
    const void* ptr;
    uint64_t d0, d1, d2, stride, stride2;
    uint32_t b0, b1, dtype, swizzle;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && stride == o.stride && stride2 == o.stride2 &&
               b0 == o.b0 && b1 == o.b1 && dtype == o.dtype && swizzle == o.swizzle;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = reinterpret_cast<uint64_t>(k.ptr);
        auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.d0), mix(k.d1), mix(k.d2), mix(k.stride), mix(k.stride2), mix(k.b0), mix(k.b1), mix(k.dtype), mix(k.swizzle);
        return static_cast<size_t>(h);
    }
};

// Tiled map: inner dim d0 contiguous, outer dim d1 with `stride_bytes` pitch, optional batch dim d2 (0 = rank 2) with
// `stride2_bytes` pitch; box b0 x b1 (x 1). Encoding is a pure function of its arguments, so maps are memoised per thread
// (the reference re-encodes five maps on every call, impls/sm100_fp8_fp4_gemm_1d1d.hpp:117-135).
int make_map(CUtensorMap* out, const void* ptr, CUtensorMapDataType dtype, uint64_t d0, uint64_t d1, uint64_t stride_bytes,
             uint32_t b0, uint32_t b1, CUtensorMapSwizzle swizzle, uint64_t d2 = 0, uint64_t stride2_bytes = 0) {
    thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, d0, d1, d2, stride_bytes, stride2_bytes, b0, b1, static_cast<uint32_t>(dtype), static_cast<uint32_t>(swizzle)};
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return DGB200_OK;
    }
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {stride_bytes, stride2_bytes};
    const cuuint32_t box[3] = {b0, b1, 1};
    const cuuint32_t elem_strides[3] = {1, 1, 1};
    CUresult res = rt().encode(out, dtype, d2 > 0 ? 3 : 2, const_cast<void*>(ptr), dims, strides, box, elem_strides,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS)
        return fail(DGB200_ERR_CUDA,
                    "cuTensorMapEncodeTiled failed (%d): ptr=%p dims={%llu,%llu,%llu} strides={%llu,%llu} box={%u,%u} swizzle=%d",
                    static_cast<int>(res), ptr, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                    (unsigned long long)stride_bytes, (unsigned long long)stride2_bytes, b0, b1, static_cast<int>(swizzle));
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return DGB200_OK;
}
int make_map_2d(CUtensorMap* out, const void* ptr, CUtensorMapDataType dtype, uint64_t d0, uint64_t d1,
                uint64_t stride_bytes, uint32_t b0, uint32_t b1, CUtensorMapSwizzle swizzle) {
    return make_map(out, ptr, dtype, d0, d1, stride_bytes, b0, b1, swizzle);
}

// ------------------------------------------------------------------------------------------------ heuristics
struct Problem {
    int type;          // GemmType
    int m;             // dense M | contiguous total M | masked M_max
    int expected_m;    // rows per group the caller expects (== m for dense)
    int n, k, groups;
    int alignment;     // contiguous layouts: group start alignment
    int max_splits = 1;  // > 1 only for dense problems whose caller supplied a split-K workspace
    bool x_mn = false;   // MN-major tokens: the token tile is loaded in 32/64/128-byte swizzle atoms
    int el = 1;          // operand bytes per element (2: BF16 operands)
    bool any_mn = false; // any MN-major operand: no weight multicast
    bool tma_store_ok = false;   // the output may leave through the staged TMA-store epilogue (plain BF16 tiles)
    bool swapped = false;        // transposed-output orientation: `m` is the weight count (tiled freely), `n` the token count (lanes)
    int forced_block_m = 0;      // tile height fixed by the caller (the orientation rule): skips the cost model
    bool plain_only = false;     // no split-K of any kind, no multicast clusters (the BF16-operand instances)
};
int stage_bytes(int block_m, int cluster) { return static_cast<int>(slot_bytes(block_m, cluster)); }
int smem_bytes_for(int block_m, int cluster, int stages, int staging_bytes = 0) {
    // stage slots | barriers | tmem pointer + split-K flag | (cluster split-K: reduction barrier, 16-byte aligned staging)
    return stages * stage_bytes(block_m, cluster) + (3 * stages + 4) * 8 + 16 + (staging_bytes ? 32 + staging_bytes : 0);
}
int csplit_staging_bytes(int block_m, int splits) { return (splits - 1) * (block_m / splits) * 128 * 4; }

// Estimated cycles for the whole problem with a given tile height. All busy CTAs advance one k-block per "step";
// a step is bounded by the slowest of four resources (constants are B200 measurements / fits):
//   tensor pipe : a K=32 UMMA of a CTA pair retires in block_m/2 cycles -> 2*block_m cycles per 128-K block
//   SM ingest   : one SM pulls at most ~kSmIngest B/cycle through TMA
//   L2          : all busy CTAs together pull at most ~kL2Rate B/cycle
//   HBM         : bytes that are new to the chip in this step (weight tiles are shared by the m-blocks in flight,
//                 token tiles by the n-units in flight) at ~kHbmRate B/cycle
constexpr double kSmIngest = 45.0, kL2Rate = 8000.0, kHbmRate = 3400.0, kTileOverhead = 1500.0, kSplitOverhead = 2500.0;
constexpr int kMaxSplits = 8;
constexpr int kTmaStoreMinBlockM = 176;               // shorter tiles keep the direct-store epilogue
constexpr int kPairSplitMaxBlockM = 128;              // pair split-K by default: one m-block of at most this many rows
constexpr int kSplitKCounters = 4096;                 // ints at the start of the workspace
constexpr size_t kSplitKHeaderBytes = kSplitKCounters * sizeof(int);

double estimate_cycles(const Problem& pb, int block_m, int cluster_total, int num_sms, int splits) {
    const int cluster = std::min(cluster_total, 2);
    const int num_units = num_sms / cluster;
    const int n_units = ceil_div(pb.n, (int)kBlockN * cluster);
    const int num_kb = ceil_div(ceil_div(pb.k, (int)kBlockK), splits);
    int m_blocks;  // m-blocks that share one weight panel
    double tiles;
    if (pb.type == kMMasked || pb.type == kBatched) {
        m_blocks = ceil_div(std::max(pb.expected_m, 1), block_m);
        tiles = (double)pb.groups * m_blocks * n_units;
    } else if (pb.type == kDense) {
        m_blocks = ceil_div(pb.m, block_m);
        tiles = (double)m_blocks * n_units * splits;
    } else {
        m_blocks = ceil_div(std::max(pb.expected_m, 1), block_m);
        tiles = (double)ceil_div(pb.m, block_m) * n_units;
    }
    const double waves = std::ceil(tiles / num_units);
    const double busy_units = std::min<double>(tiles, num_units);
    const double busy_ctas = busy_units * cluster;
    const double cta_bytes = (128.0 + (double)block_m / cluster) * kBlockK;
    const double distinct_n = std::min<double>((double)n_units * splits, busy_units);   // distinct weight panels in flight
    const double distinct_m = std::max(1.0, busy_units / distinct_n);
    const double shared_m = std::min<double>(distinct_m, m_blocks);  // m-blocks in flight that reuse a weight tile
    const double new_bytes = (busy_ctas * 128.0 / shared_m + distinct_m * block_m) * kBlockK;
    const double step = std::max(std::max(2.0 * block_m, cta_bytes / kSmIngest),
                                 std::max(busy_ctas * cta_bytes / kL2Rate, new_bytes / kHbmRate));
    return waves * (num_kb * step + kTileOverhead + 6.0 * block_m) + (splits > 1 ? kSplitOverhead : 0.0);
}

Config choose_config(const Problem& pb, int num_sms_override = 0) {
    Config c{};
    c.num_sms = num_sms_override > 0 ? (num_sms_override & ~1) : effective_num_sms();
    c.cluster = c.num_sms >= 2 ? 2 : 1;
    if (pb.swapped && pb.n <= (int)kBlockN) c.cluster = 1;    // up to 128 token rows fit the lanes of one CTA: no pair needed
    if (int v = env_int("DGB200_CLUSTER", 0)) c.cluster = v;
    if (pb.plain_only && c.num_sms >= 2) c.cluster = 2;
    c.swap_d = pb.swapped;
    std::vector<int> candidates;
    if (pb.type == kDense || pb.type == kMMasked || pb.type == kKGrouped || pb.type == kKGroupedPsum || pb.type == kBatched || pb.type == kBatchReduce) {
        const int step = pb.x_mn ? 32 / pb.el * std::min(c.cluster, 2) : 16;   // MN-major tokens: a CTA's rows are whole 32-byte atoms
        for (int bm = step; bm <= (int)kMaxBlockM; bm += step) {
            // ... and narrow atoms are slow (k-grouped FP8, 4096 x 7168: 192-row tiles = 32 B atoms 280 us, 128-row = 64 B atoms
            // 262; BF16 160 / 224 rows = 64 / 32 B atoms 421 / 411 us against 356 at 192): keep the heights whose atoms are
            // at least 64 bytes wide, plus the shortest tile that covers a small problem
            if (pb.x_mn && (bm / std::min(c.cluster, 2) * pb.el) % 64 != 0 && bm < align_up(pb.m, step)) continue;
            candidates.push_back(bm);
        }
    } else {
        // a tile must not straddle two groups: block_m has to divide the group alignment
        for (int bm = 16; bm <= (int)kMaxBlockM; bm += 16)
            if (pb.alignment % bm == 0) candidates.push_back(bm);
        if (candidates.empty()) candidates.push_back(16);
    }
    const int num_kb = ceil_div(pb.k, (int)kBlockK);
    const int max_splits = (rt().split_k && !pb.plain_only) ? std::min({pb.max_splits, kMaxSplits, std::max(1, num_kb / 4)}) : 1;
    double best = 1e300;
    c.block_m = candidates[0];
    c.num_splits = 1;
    for (int bm : candidates) {
        if (pb.type != kMContiguous && pb.type != kMContiguousPsum && bm != candidates[0] && bm - 16 >= align_up(pb.m, 16))
            continue;  // taller than the problem
        for (int sp = 1; sp <= max_splits; ++sp) {
            if (sp > 1) {
                // split-K only pays for very small problems: the finalising pass costs ~3 us (measured), so it is used
                // when a handful of short tiles would otherwise leave most SMs idle (e.g. M=1, N=2112: 15 -> 10.8 us)
                if (c.cluster > 2 || bm > 32) break;
                if (ceil_div(pb.m, bm) * ceil_div(pb.n, (int)kBlockN * c.cluster) * 4 > c.num_sms / c.cluster) break;
                const int tiles = ceil_div(pb.m, bm) * ceil_div(pb.n, (int)kBlockN * c.cluster);
                if (tiles * sp > c.num_sms / c.cluster) break;
                if (ceil_div(num_kb, sp) * (sp - 1) >= num_kb) continue;    // an empty slice
            }
            const double t = estimate_cycles(pb, bm, c.cluster, c.num_sms, sp);
            if (t < best * 0.999) best = t, c.block_m = bm, c.num_splits = sp;  // ties -> smaller tile, fewer slices
        }
    }
    if (pb.forced_block_m > 0) c.block_m = pb.forced_block_m, c.num_splits = 1;
    if (int v = env_int("DGB200_BLOCK_M", 0)) c.block_m = v;
    if (const char* v = env_str("DGB200_SPLITS")) c.num_splits = std::max(1, std::min(atoi(v), max_splits));
    c.kb_per_split = ceil_div(num_kb, c.num_splits);
    c.num_splits = ceil_div(num_kb, c.kb_per_split);

    // Wave balancing (dense, tensor-bound sizes). A persistent grid of P CTA pairs runs ceil(tiles / P) rounds, so e.g.
    // M = N = 4096 with 240-row tiles is 288 tiles = 3.89 -> 4 rounds on 74 pairs although the work is worth 3.69. Tile
    // height barely matters above ~160 rows (measured), the tile COUNT does: pick the number of m-blocks that minimises
    // the simulated makespan of the round-robin schedule, and make the blocks as even as multiples of 16 allow (two
    // heights, the taller ones first).
    c.num_tall = 0, c.block_m_low = 0;
    if (pb.type == kDense && !pb.swapped && !pb.x_mn && c.num_splits == 1 && c.cluster == 2 && pb.m >= 1024 && c.block_m >= 160 &&
        !env_str("DGB200_BLOCK_M") && env_int("DGB200_BALANCE", 1)) {
        const int units = c.num_sms / 2, n_units = ceil_div(pb.n, (int)kBlockN * 2);
        const double overhead_rows = kTileOverhead / (2.0 * num_kb);      // per-tile fixed cost in units of token rows
        double best_ms = 1e300;
        const int force_nb = env_int("DGB200_M_BLOCKS", 0);      // (development: pin the number of m-blocks)
        for (int nb = ceil_div(pb.m, (int)kMaxBlockM); nb <= ceil_div(pb.m, 160) && nb * n_units <= 64 * units; ++nb) {
            if (force_nb && nb != force_nb) continue;
            const int hi = align_up(ceil_div(pb.m, nb), 16), lo = hi - 16;
            if (hi > (int)kMaxBlockM || lo < 16) continue;
            // tall blocks: smallest count with tall * hi + (nb - tall) * lo >= m
            const int tall = std::max(0, std::min(nb, ceil_div(pb.m - nb * lo, 16)));
            std::vector<double> load(units, 0.0);
            const int gw = std::max(1, env_int("DGB200_SWIZZLE_GROUP", 8));
            int idx = 0;
            for (int g0 = 0; g0 < n_units; g0 += gw) {
                const int width = std::min(gw, n_units - g0);
                for (int mb = 0; mb < nb; ++mb) {
                    const int row0 = mb < tall ? mb * hi : tall * hi + (mb - tall) * lo;
                    const int h = std::max(0, std::min(mb < tall ? hi : lo, pb.m - row0));
                    // taller tiles are worth more than their rows: fewer operand bytes per FLOP, which under the power cap is
                    // clock (measured, tools/tune.py mblocks / mblocks2: 4096 x 4096 x 7168 76.1 us with 18 m-blocks of 240 / 224
                    // rows against 76.9 with 23 of 192 / 176; 3000 x 4096 x 7168 58.6 against 60.6) -- ~1 % per 24 rows below 240
                    const double rows = align_up(std::max(h, 1), 16);
                    for (int j = 0; j < width; ++j, ++idx) load[idx % units] += rows * (1.0 + 0.10 * (240.0 - rows) / 240.0) + overhead_rows;
                }
            }
            const double ms = *std::max_element(load.begin(), load.end());
            if (ms < best_ms * 0.995) best_ms = ms, c.block_m = hi, c.block_m_low = lo, c.num_tall = tall;
        }
    }

    // Cluster split-K (dense, K-major, small M): S single-CTA MMAs share one output tile, each streams 1/S of K, and
    // the partial tiles are reduced through distributed shared memory. Every weight byte then crosses L2 -> SM once
    // (instead of once per m-block) and all SMs stream from HBM even when there are few output tiles.
    c.csplit = 0, c.grid = 0, c.grid_y = 1;
    const char* splits_env = env_str("DGB200_SPLITS");
    const bool pinned = env_str("DGB200_BLOCK_M") || env_str("DGB200_CLUSTER") || (splits_env && atoi(splits_env) <= 1);
    if (pb.type == kDense && !pb.swapped && !pb.plain_only && !pb.any_mn && c.cluster == 2 && pb.m > 0 && rt().split_k && !(pinned && !env_str("DGB200_CSPLIT"))) {
        const int want = env_int("DGB200_CSPLIT", -1);            // -1: heuristic, 0: off, 2/4: forced
        int pick = 0, pick_bm = 0;
        for (int sp : {4, 2}) {
            if (want == 0 || (want > 0 && want != sp)) continue;
            const int bm = std::min(align_up(std::min(pb.m, 128), 16 * sp), align_up(128, 16 * sp));
            const int tiles = ceil_div(pb.m, bm) * ceil_div(pb.n, (int)kBlockN);
            if (bm > (int)kMaxBlockM || ceil_div(num_kb, sp) * (sp - 1) >= num_kb || num_kb / sp < 2) continue;
            if (want > 0) {
                pick = sp, pick_bm = bm;
                break;
            }
            // Measured on B200 (tools/tune.py small): four slices win whenever the whole tile set fits the chip once
            // (tiles x 4 <= SMs) and K is long enough to amortise the exchange (~1.5 us): M <= 128 at N = 4096 / 2112,
            // K = 7168 runs 8-17% faster than one CTA pair per tile; two slices never paid.
            if (sp == 4 && pb.m <= 128 && tiles * sp <= c.num_sms && num_kb >= 16) pick = sp, pick_bm = bm;
            if (pick) break;
        }
        if (pick) {
            c.csplit = pick, c.cluster = pick, c.block_m = pick_bm;
            c.kb_per_split = ceil_div(num_kb, pick), c.num_splits = pick;
            c.grid = ceil_div(pb.n, (int)kBlockN) * pick, c.grid_y = ceil_div(pb.m, pick_bm);   // (n-tile x slice, m-block)
        }
    }
    // Pair split-K (dense, K-major, medium M): the same exchange between S CTA PAIRS (cluster of 2 S). A pair owns 256 weight
    // rows x block_m tokens of HALF (a quarter) of K, so a 192..480-row problem runs as few tall tiles on all SMs instead of
    // many short ones: the bytes every SM pulls through L2 per output drop by ~40 % (the mid-M shapes are bound by L2 -> SM
    // traffic, not by HBM or the tensor pipe). DGB200_PSPLIT = 0 / 2 / 4 pins it, DGB200_PSPLIT_BM the tile height.
    if (!c.csplit && pb.type == kDense && !pb.swapped && !pb.plain_only && !pb.any_mn && c.cluster == 2 && pb.m > 0 && rt().split_k) {
        const int want = env_int("DGB200_PSPLIT", -1);
        int pick = 0, pick_bm = 0;
        for (int sp : {2, 4}) {
            if (want == 0 || (want > 0 && want != sp) || (want < 0 && pinned)) continue;
            const int unit = 16 * sp, max_bm = (int)kMaxBlockM / unit * unit;
            // forced: tallest tiles that cover M evenly. Heuristic: the SHORTEST tiles (>= 64 rows, <= kPairSplitMaxBlockM) whose
            // clusters of 4 all run at once -- 8 GPCs x 4 clusters on a 148-SM part -- since more CTAs stream more bytes at a
            // time (256 x 2112 x 7168: 96-row tiles 11.9 us, 128-row 12.8; 64-row tiles = 36 clusters, a second wave: 18.8).
            const int max_clusters = c.num_sms * 7 / 32;
            const int cap = want > 0 ? max_bm : std::min(max_bm, kPairSplitMaxBlockM / unit * unit);
            int bm = align_up(ceil_div(pb.m, ceil_div(pb.m, cap)), unit);
            if (want <= 0)
                for (int cand = 64; cand <= cap; cand += unit)
                    if (ceil_div(pb.m, cand) * ceil_div(pb.n, 2 * (int)kBlockN) <= max_clusters) {
                        bm = cand;
                        break;
                    }
            if (int v = env_int("DGB200_PSPLIT_BM", 0)) bm = v;
            if (bm % unit != 0 || bm > max_bm || num_kb / sp < 2 || ceil_div(num_kb, sp) * (sp - 1) >= num_kb) continue;
            const int tiles = ceil_div(pb.m, bm) * ceil_div(pb.n, 2 * (int)kBlockN);
            if (want > 0) {
                pick = sp, pick_bm = bm;
                break;
            }
            // Measured (tools/tune.py mid, kineto us, ours plain / pair split / reference): 192 x 4096 x 7168 15.2 / 14.0 / 14.6,
            // 256 x 4096 x 7168 15.9 / 15.0 / 14.3, 256 x 2112 x 7168 14.2 / 12.1 / 11.8 -- it pays while all pair-slices run
            // as ONE wave and K is long; with more tiles than that (320+ rows at N = 4096, or N = 7168) the plain kernel wins.
            if (sp == 2 && pb.m > 128 && bm >= 64 && tiles <= max_clusters && num_kb >= 32 && c.num_splits == 1) pick = sp, pick_bm = bm;
            if (pick) break;
        }
        if (pick) {
            c.csplit = pick, c.cluster = 2 * pick, c.block_m = pick_bm;
            c.kb_per_split = ceil_div(num_kb, pick), c.num_splits = pick;
            c.grid = ceil_div(pb.n, 2 * (int)kBlockN) * 2 * pick, c.grid_y = ceil_div(pb.m, pick_bm);   // (n-unit x slice x 2, m-block)
            c.num_tall = 0, c.block_m_low = 0;
        }
    }
    const int cta_group = c.csplit ? c.cluster / c.csplit : std::min(c.cluster, 2);
    const int staging = c.csplit ? csplit_staging_bytes(c.block_m, c.csplit) : 0;
    // Staged TMA-store epilogue: pays when tiles are tall (the direct epilogue stores 64 B per instruction and keeps the
    // epilogue warps busy until the last row has left); small tiles keep all of shared memory for the TMA -> MMA ring.
    // DGB200_TMA_STORE = 0 / 1 pins the choice (development).
    auto max_stages = [&](int store_bytes) {
        int st = (kSmemCapacity - 64 - store_bytes - (staging ? 32 + staging : 0)) / stage_bytes(c.block_m, cta_group);
        while (st > 1 && smem_bytes_for(c.block_m, cta_group, st, staging) + store_bytes > kSmemCapacity) --st;
        return st;
    };
    c.tma_store = 0;
    if (pb.swapped) {
        // transposed output: per-warp 32 x 32 staging (16 KB), tile width a multiple of 32 columns. DGB200_TMA_STORE pins it.
        const int want = env_int("DGB200_TMA_STORE", -1);
        c.tma_store = pb.tma_store_ok && c.block_m % (int)kSwapStoreCols == 0 && c.cluster <= 2 && (want >= 0 ? want != 0 : c.block_m >= kTmaStoreMinBlockM);
    } else if (pb.tma_store_ok && c.cluster == 2 && !c.csplit && c.num_splits == 1) {
        // Measured (tools/tune.py store): the staged epilogue wins 1-4 % on tall tiles with a long enough K loop to hide it
        // behind (4096 x 4096 x 7168, 4096 x 7168 x 2048, 4096 x 24576 x 1536 at 240 rows); it loses when it costs a pipeline
        // stage, on short tiles, and when the kernel is epilogue-bound (K = 512: 8 warps of direct stores move more bytes per
        // cycle than one TMA issuer; K = 1536 at 240 rows: 106.7 us staged, 104.2 direct) -- hence K >= 2048.
        const int want = env_int("DGB200_TMA_STORE", -1);
        c.tma_store = want >= 0 ? (want != 0)
                                : (c.block_m >= kTmaStoreMinBlockM && num_kb >= 16 && max_stages((int)kStoreStagingBytes) == max_stages(0));
    }
    const int store_bytes = c.tma_store ? (pb.swapped ? (int)kSwapStagingBytes : (int)kStoreStagingBytes) : 0;
    int stages = max_stages(store_bytes);
    stages = std::min(stages, 32);
    if (int v = env_int("DGB200_STAGES", 0)) stages = std::min(v, stages);
    c.stages = std::max(stages, 1);
    c.smem_bytes = smem_bytes_for(c.block_m, cta_group, c.stages, staging) + store_bytes;
    c.swizzle_group = env_int("DGB200_SWIZZLE_GROUP", 8);
    return c;
}

// Which orientation for a dense K-major problem? DGB200_SWAP = 0 / 1 pins it (development). Returns the weight-tile width
// to use in the transposed-output orientation (0: stay with the default one; -1: transposed, width chosen by the cost model).
// Rule (tools/tune.py swap / swap2): the default orientation tiles N in fixed units of 256 weight rows per CTA pair, so a
// medium-M problem with a wide N needs two rounds of pairs where 224-row weight tiles under 256 tokens make ONE
// (512 x 7168 x 2048: 12.2 us default, 11.3 transposed with staged stores, 11.2 reference); everywhere else the default
// orientation was level or ahead, so nothing else selects it.
int want_swapped_orientation(const GemmCall& c, int num_sms_override = 0) {
    const int want = env_int("DGB200_SWAP", -1);
    if (want >= 0) return want != 0 ? -1 : 0;
    if (env_str("DGB200_BLOCK_M") || env_str("DGB200_CLUSTER") || env_str("DGB200_SPLITS")) return 0;   // pinned configurations
    if (c.d_dtype != DGB200_BF16 || c.accumulate || c.m <= (int)kBlockN || (reinterpret_cast<uintptr_t>(c.d) & 15) != 0 || (c.ldd * 2) % 16 != 0)
        return 0;
    Problem pb{kDense, c.m, c.m, c.n, c.k, 1, 1};
    const Config dflt = choose_config(pb, num_sms_override);
    if (dflt.csplit || dflt.num_splits > 1 || dflt.cluster != 2) return 0;
    const int pairs = dflt.num_sms / 2, num_kb = ceil_div(c.k, (int)kBlockK);
    const int default_tiles = ceil_div(c.m, dflt.block_m) * ceil_div(c.n, 2 * (int)kBlockN);
    if (default_tiles <= pairs || num_kb < 8) return 0;                       // already one round of pairs
    const int token_units = ceil_div(c.m, 2 * (int)kBlockN);
    for (int bn = 128; bn <= 224; bn += 32)                                    // the narrowest tiles that still make one round
        if (token_units * ceil_div(c.n, bn) <= pairs) return bn;
    return 0;
}

// ------------------------------------------------------------------------------------------------ launch
int run_gemm(const GemmCall& c) {
    EnvScope env_scope;
    if (int e = ensure_device()) return e;
    DGB_REQUIRE(c.gran_k_a == 32 || c.gran_k_a == 128);
    DGB_REQUIRE(c.gran_k_b == 32 || c.gran_k_b == 128);
    DGB_REQUIRE(c.lda % 16 == 0 && c.ldb % 16 == 0);  // TMA: 16-byte pitch
    DGB_REQUIRE((reinterpret_cast<uintptr_t>(c.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(c.b) & 15) == 0);
    if (!c.bf16_ab) {
        DGB_REQUIRE((reinterpret_cast<uintptr_t>(c.sfa) & 15) == 0 && (reinterpret_cast<uintptr_t>(c.sfb) & 15) == 0);
        DGB_REQUIRE(c.sfa_stride % 4 == 0 && c.sfb_stride % 4 == 0);
    }
    const bool k_grouped = c.type == kKGrouped || c.type == kKGroupedPsum;

    const bool batched = c.type == kBatched || c.type == kBatchReduce;   // rank-3 tensor maps (batch = third coordinate)
    const bool head_split = c.head_mid > 0;
    Problem pb{c.type, c.m, c.expected_m, c.n, c.k, c.groups, c.alignment};
    pb.x_mn = c.x_mn, pb.any_mn = c.x_mn || c.w_mn, pb.el = c.bf16_ab ? 2 : 1;
    pb.swapped = c.swap_d, pb.forced_block_m = c.forced_block_m;
    if (c.type == kBatchReduce)     // few, tall output tiles; the parallelism comes from the batch chunks
        pb.forced_block_m = align_up(ceil_div(c.m, ceil_div(c.m, (int)kMaxBlockM)), 16);
    pb.plain_only = c.bf16_ab && !(c.type == kDense && !c.x_mn && !c.w_mn);   // BF16: cluster split-K is built for dense K-major only
    // TMA stores need a 16-byte aligned base and row pitch; tiles that must not touch rows past `valid_m` (masked, psum),
    // accumulate into C or remap columns keep the predicated direct stores
    pb.tma_store_ok = (c.type == kDense || c.type == kMContiguous) && !c.bf16_ab && c.d_dtype == DGB200_BF16 && !c.accumulate && !head_split &&
                      (reinterpret_cast<uintptr_t>(c.d) & 15) == 0 && (c.ldd * 2) % 16 == 0 && c.arrival == nullptr &&
                      (!c.swap_d || c.m % 8 == 0);    // (transposed output: 16-byte pieces must not straddle the last column)
    // split-K needs scratch: [4096 arrival counters][num_splits x m x n fp32 partial tiles]
    if (c.type == kDense && !head_split && !c.swap_d && c.workspace != nullptr && c.n % 4 == 0 && c.workspace_bytes > kSplitKHeaderBytes &&
        (reinterpret_cast<uintptr_t>(c.workspace) & 15) == 0) {
        const size_t per_split = static_cast<size_t>(c.m) * c.n * sizeof(float);
        pb.max_splits = static_cast<int>(std::min<size_t>(kMaxSplits, (c.workspace_bytes - kSplitKHeaderBytes) / per_split));
    }
    Config cfg = choose_config(pb);
    if (cfg.num_splits > 1 && ceil_div(c.m, cfg.block_m) * ceil_div(c.n, (int)kBlockN) > kSplitKCounters)
        cfg.num_splits = 1, cfg.kb_per_split = ceil_div(c.k, (int)kBlockK);
    if (pb.any_mn && cfg.cluster > 2) cfg.cluster = 2;            // weight multicast is built for K-major tiles only
    const int cta_group = cfg.csplit ? cfg.cluster / cfg.csplit : (cfg.cluster >= 2 ? 2 : 1);
    const int pairs = (cfg.cluster >= 2 && !cfg.csplit) ? cfg.cluster / 2 : 1;
    const int load_m = cfg.block_m / cta_group;
    DGB_REQUIRE(cfg.block_m % 16 == 0 && cfg.block_m >= 16 && cfg.block_m <= (int)kMaxBlockM);
    DGB_REQUIRE(cfg.cluster == 1 || cfg.cluster == 2 || (c.type == kDense && (cfg.cluster == 4 || cfg.cluster == 8)));
    DGB_REQUIRE(cta_group == 1 || cta_group == 2);
    if (cfg.cluster > 2 && !cfg.csplit) cfg.tma_store = 0;
    DGB_REQUIRE(cfg.cluster <= 2 || cfg.num_splits == 1 || cfg.csplit);
    if (cfg.csplit) DGB_REQUIRE(cfg.block_m % (16 * cfg.csplit) == 0 && (cfg.cluster == cfg.csplit || cfg.cluster == 2 * cfg.csplit));
    if (c.type == kMContiguous || c.type == kMContiguousPsum) DGB_REQUIRE(c.alignment % cfg.block_m == 0);
    cfg.overlap_producer = c.arrival != nullptr;
    const int el = c.bf16_ab ? 2 : 1;             // operand bytes per element (K-major operands are addressed in bytes anyway)
    if (c.x_mn) DGB_REQUIRE((load_m * el) % 32 == 0);

    const int num_kp_a = ceil_div(c.k, c.gran_k_a * 4), num_kp_b = ceil_div(c.k, c.gran_k_b * 4);
    const int b_groups = (c.type == kDense || k_grouped || batched) ? 1 : c.groups;      // groups folded into the weight map's rows
    const int sfa_groups = (c.type == kMMasked || batched) ? c.groups : 1;
    const int sfb_groups = batched ? c.groups : b_groups;
    const uint64_t sfa_krows = c.sfa_krows > 0 ? (uint64_t)c.sfa_krows : (uint64_t)num_kp_a * sfa_groups;
    const uint64_t sfb_krows = c.sfb_krows > 0 ? (uint64_t)c.sfb_krows : (uint64_t)num_kp_b * sfb_groups;

    // Tensor maps. K-major operand [rows, K]: box 128 K-bytes x rows. MN-major operand [K rows, MN]: box S MN-bytes x
    // 128 / el K-rows, S = one swizzle atom (128 for the weights; the widest of 128/64/32 that divides load_m for tokens).
    // BF16 operands keep the UINT8 maps: whichever extent is contiguous is addressed in bytes.
    // Batched problems add the batch as a third dimension with its own pitch (box depth 1), so permuted views
    // (fp8_einsum) need no copy.
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    int x_swizzle = 128;
    const uint64_t nb = batched ? (uint64_t)c.groups : 0;     // rank-3 maps only for the batched type
    if (c.x_mn) {
        x_swizzle = (load_m * el) % 128 == 0 ? 128 : ((load_m * el) % 64 == 0 ? 64 : 32);
        const CUtensorMapSwizzle sw = x_swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                      : (x_swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        if (int e = make_map(&maps.x, c.a, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)c.m * el, c.a_rows, c.lda, x_swizzle, kBlockK / el, sw, nb,
                             c.batch_stride_a)) return e;
    } else {
        if (int e = make_map(&maps.x, c.a, CU_TENSOR_MAP_DATA_TYPE_UINT8, c.k, c.a_rows, c.lda, kBlockK, load_m,
                             CU_TENSOR_MAP_SWIZZLE_128B, nb, c.batch_stride_a)) return e;
    }
    if (c.w_mn) {
        const uint64_t k_rows = k_grouped ? (uint64_t)c.a_rows : (uint64_t)(c.k / el) * b_groups;
        if (int e = make_map(&maps.w, c.b, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)c.n * el, k_rows, c.ldb, kBlockN, kBlockK / el,
                             CU_TENSOR_MAP_SWIZZLE_128B, nb, c.batch_stride_b)) return e;
    } else {
        if (int e = make_map(&maps.w, c.b, CU_TENSOR_MAP_DATA_TYPE_UINT8, c.k, (uint64_t)c.n * b_groups, c.ldb, kBlockK,
                             kBlockN / pairs, CU_TENSOR_MAP_SWIZZLE_128B, nb, c.batch_stride_b)) return e;
    }
    if (!c.bf16_ab) {
        if (int e = make_map_2d(&maps.sfx, c.sfa, CU_TENSOR_MAP_DATA_TYPE_INT32, c.sfa_cols, sfa_krows, (uint64_t)c.sfa_stride * 4,
                                cfg.block_m, 1, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
        if (int e = make_map_2d(&maps.sfw, c.sfb, CU_TENSOR_MAP_DATA_TYPE_INT32, c.sfb_cols, sfb_krows, (uint64_t)c.sfb_stride * 4,
                                kBlockN, 1, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    }
    if (cfg.tma_store && c.swap_d) {
        // transposed output: staged through shared memory, written with plain 16-byte stores (no tensor map)
    } else if (cfg.tma_store) {
        // D [rows, N] BF16: box 64 columns (one 128 B swizzle atom) x 16 rows; rows / columns past the end are clipped
        if (int e = make_map_2d(&maps.d, c.d, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, c.n, c.m, (uint64_t)c.ldd * 2, 64, kStoreRows,
                                CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    }

    GemmParams p{};
    p.d = c.d;
    p.grouped_layout = c.grouped_layout;
    p.m = c.m, p.n = c.n, p.k = c.k;
    p.num_groups = c.groups;
    p.block_m = cfg.block_m;
    p.num_stages = cfg.stages;
    p.ld_d = static_cast<uint32_t>(c.ldd);
    p.num_kp_x = num_kp_a, p.num_kp_w = num_kp_b;
    p.sf_shift_x = c.gran_k_a == 128 ? 2 : 0;
    p.sf_shift_w = c.gran_k_b == 128 ? 2 : 0;
    p.swizzle_group = std::max(1, cfg.swizzle_group);
    p.num_splits = cfg.num_splits;
    p.kb_per_split = cfg.kb_per_split;
    const bool ws_split = cfg.num_splits > 1 && !cfg.csplit;
    p.splitk_counters = ws_split ? static_cast<int*>(c.workspace) : nullptr;
    p.splitk_ws = ws_split ? reinterpret_cast<float*>(static_cast<char*>(c.workspace) + kSplitKHeaderBytes) : nullptr;
    p.arrival = c.arrival, p.arrival_expected = c.arrival_expected;
    {
        const uint64_t policies[3] = {ptx::kEvictNormal, ptx::kEvictFirst, ptx::kEvictLast};
        // m-grouped GEMMs stream 4-8 GB of expert weights past token tiles that every n-unit of the expert re-reads:
        // tokens evict-last keeps them in L2 (measured on the 256-expert shapes: contiguous m=64 -3.4 %, masked -1..2 %;
        // weights evict-first made things worse)
        const bool m_grouped = c.type == kMContiguous || c.type == kMContiguousPsum || c.type == kMMasked;
        p.w_hint = policies[std::min(2, std::max(0, env_int("DGB200_W_HINT", 0)))];
        p.x_hint = policies[std::min(2, std::max(0, env_int("DGB200_X_HINT", m_grouped ? 2 : 0)))];
    }
    p.debug_ts = g_debug_ts.load();
    p.num_n_units = ceil_div(c.n, (int)kBlockN * cta_group);
    p.num_m_blocks = ceil_div(c.m, cfg.block_m);
    p.num_tall = 0xffffffffu, p.block_m_low = cfg.block_m;
    if (c.type == kDense && cfg.block_m_low > 0) {
        p.num_tall = cfg.num_tall, p.block_m_low = cfg.block_m_low;
        const int rest = std::max(0, c.m - cfg.num_tall * cfg.block_m);
        p.num_m_blocks = cfg.num_tall + ceil_div(rest, cfg.block_m_low);
    }
    p.k_shift = c.bf16_ab ? 1 : 0;
    p.m_alignment = std::max(1, c.alignment) << (k_grouped ? p.k_shift : 0);
    // Dense problems that make at most one tile per cluster run on a 2-D grid of exactly those clusters: every role knows its
    // tile from the block index (stamps: the first TMA load leaves ~450 cycles earlier than behind the persistent scheduler).
    p.grid_tiles = 0;
    if (c.type == kDense && !cfg.csplit && cfg.num_splits == 1 && cfg.cluster <= 2 && cfg.grid == 0 && env_int("DGB200_GRID_TILES", 1) &&
        (int)(p.num_m_blocks * p.num_n_units) <= cfg.num_sms / cfg.cluster) {
        cfg.grid = p.num_n_units * cfg.cluster, cfg.grid_y = p.num_m_blocks;
        p.grid_tiles = 1;
    }
    p.zero_padding = c.zero_padding;
    p.x_swizzle = x_swizzle;
    p.sf_k_span = 4 * c.gran_k_a;
    p.d_batch_stride = c.type == kBatched ? static_cast<uint64_t>(c.batch_stride_d) : 0;
    if (c.type == kBatchReduce) {
        // cut the batches into chunks so that every CTA pair gets about one tile (at least 2 batches per chunk)
        const int per_chunk = p.num_m_blocks * p.num_n_units, pairs = std::max(1, cfg.num_sms / cfg.cluster);
        const int want_chunks = std::max(1, std::min(ceil_div(c.groups, 2), std::max(1, pairs / per_chunk)));
        p.kb_per_split = ceil_div(c.groups, want_chunks);                  // batches per chunk
        p.num_splits = ceil_div(c.groups, (int)p.kb_per_split);            // chunks
    }
    p.head_lr = head_split ? c.head_left + c.head_right : 1, p.head_mid = head_split ? c.head_mid : 0, p.head_right = c.head_right;

    g_last_config = dgb200_config{cfg.block_m, cfg.cluster, cfg.stages, cfg.num_sms, cfg.smem_bytes, 0, cfg.num_splits, cfg.csplit,
                                  cfg.tma_store, cfg.swap_d};
    if (env_int("DGB200_PRINT_CONFIGS", 0))
        fprintf(stderr, "dgb200 config: type=%d m=%d n=%d k=%d groups=%d majors=%d%d -> block_m=%d cluster=%d stages=%d sms=%d smem=%d splits=%d tma_store=%d\n",
                c.type, c.m, c.n, c.k, c.groups, (int)c.x_mn, (int)c.w_mn, cfg.block_m, cfg.cluster, cfg.stages, cfg.num_sms,
                cfg.smem_bytes, cfg.csplit ? -cfg.num_splits : cfg.num_splits, cfg.tma_store);

    if (c.bf16_ab) return (c.x_mn || c.w_mn || batched) ? dispatch_bf16_mn(c, cfg, maps, p) : dispatch_bf16(c, cfg, maps, p);
    if (c.type == kBatchReduce) return fail(DGB200_ERR_UNSUPPORTED, "the batch-reduction GEMM is built for BF16 operands only");
    switch (c.type) {
        case kDense:
            if (c.swap_d) return dispatch_dense_swap(c, cfg, maps, p);
            if (cfg.num_splits > 1 && !cfg.csplit) return dispatch_dense_splitk(c, cfg, maps, p);
            return (c.x_mn || c.w_mn) ? dispatch_dense_mn(c, cfg, maps, p) : dispatch_dense_kk(c, cfg, maps, p);
        case kMContiguous: case kMMasked: case kMContiguousPsum: case kKGrouped: case kKGroupedPsum:
            return dispatch_grouped(c, cfg, maps, p);
        case kBatched: return dispatch_batched(c, cfg, maps, p);
        default: return fail(DGB200_ERR_INVALID_ARGUMENT, "unknown gemm type %d", c.type);
    }
}

}  // namespace

// ---- services for the kernel-instance translation units (launch.cuh)
int host_fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
int rt_device() { return rt().device; }
int rt_pdl() { return rt().pdl; }
void count_launches(int n) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

}  // namespace dgb200

// ================================================================================================ C ABI
using namespace dgb200;

extern "C" {

const char* dgb200_last_error(void) { return g_last_error.c_str(); }
int dgb200_version(void) { return DGB200_VERSION; }

int dgb200_set_num_sms(int num_sms) {
    // Like the reference (csrc/jit/device_runtime.hpp:103-105): any 0 <= n <= SM count, 0 = all SMs. An odd budget
    // (e.g. total - communication SMs) is accepted; the launch rounds it down to whole CTA pairs.
    DGB_REQUIRE(num_sms >= 0);
    if (rt().device_ready) DGB_REQUIRE(num_sms <= rt().sm_count);
    rt().num_sms = num_sms;
    return DGB200_OK;
}
int dgb200_get_num_sms(void) {
    if (rt().num_sms > 0 || !rt().device_ready) return rt().num_sms;   // what the caller set (reference semantics)
    return rt().sm_count;
}
int dgb200_set_tc_util(int percent) {
    DGB_REQUIRE(percent > 0 && percent <= 100);
    rt().tc_util = percent;
    return DGB200_OK;
}
int dgb200_get_tc_util(void) { return rt().tc_util; }
int dgb200_set_pdl(int enabled) {
    rt().pdl = enabled != 0;
    return DGB200_OK;
}
int dgb200_get_pdl(void) { return rt().pdl; }
int dgb200_set_split_k(int allow) {
    rt().split_k = allow ? 1 : 0;
    return DGB200_OK;
}
int dgb200_get_split_k(void) { return rt().split_k; }

int dgb200_set_mk_alignment_for_contiguous_layout(int alignment) {
    DGB_REQUIRE(alignment > 0 && alignment % 16 == 0);
    rt().mk_alignment = alignment;
    return DGB200_OK;
}
int dgb200_get_mk_alignment_for_contiguous_layout(void) { return rt().mk_alignment; }
int dgb200_get_theoretical_mk_alignment_for_contiguous_layout(int expected_m) {
    // Largest tile height this kernel supports, shrunk while it still covers `expected_m`
    // (the reference returns 224 on SM100 and shrinks in steps of 32, heuristics/runtime.hpp:47-57).
    int block_m = 224;
    if (expected_m > 0)
        for (; block_m > 32 && block_m - 32 >= expected_m; block_m -= 32) {
        }
    return block_m;
}
int dgb200_get_tma_aligned_size(int x, int element_size) {
    if (element_size <= 0 || 16 % element_size != 0) return -1;
    return align_up(x, 16 / element_size);
}

int dgb200_pack_sf_ue8m0(const float* sf, int32_t* out, int mn, int sf_k, int num_groups, int gran_mn,
                         int64_t stride_g, int64_t stride_mn, int64_t stride_k, const int32_t* psum_layout,
                         int num_psum_groups, int m_alignment, void* stream) {
    if (int e = ensure_device()) return e;
    DGB_REQUIRE(mn > 0 && sf_k > 0 && num_groups > 0 && gran_mn > 0);
    const int aligned_mn = align_up(mn, 4), num_kp = ceil_div(sf_k, 4);
    const dim3 grid(ceil_div(aligned_mn, 128), num_kp, num_groups);
    DGB_REQUIRE(num_kp <= 65535 && num_groups <= 65535);
    if (psum_layout != nullptr) {
        DGB_REQUIRE(num_psum_groups > 0 && m_alignment > 0);
        pack_sf_ue8m0_kernel<true><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
            sf, reinterpret_cast<uint32_t*>(out), mn, aligned_mn, sf_k, num_kp, gran_mn, stride_g, stride_mn, stride_k,
            psum_layout, num_psum_groups, m_alignment);
    } else {
        pack_sf_ue8m0_kernel<false><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
            sf, reinterpret_cast<uint32_t*>(out), mn, aligned_mn, sf_k, num_kp, gran_mn, stride_g, stride_mn, stride_k,
            nullptr, 0, 1);
    }
    DGB_CUDA(cudaGetLastError());
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return DGB200_OK;
}

int dgb200_transpose_sf_fp32(const float* sf, float* out, int mn, int sf_k, int num_groups, int64_t stride_g,
                             int64_t stride_mn, int64_t stride_k, void* stream) {
    if (int e = ensure_device()) return e;
    DGB_REQUIRE(mn > 0 && sf_k > 0 && num_groups > 0);
    DGB_REQUIRE(sf_k <= 65535 && num_groups <= 65535);
    const int aligned_mn = align_up(mn, 4);
    const dim3 grid(ceil_div(mn, 128), sf_k, num_groups);
    transpose_sf_fp32_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(sf, out, mn, aligned_mn, sf_k, stride_g,
                                                                                  stride_mn, stride_k);
    DGB_CUDA(cudaGetLastError());
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return DGB200_OK;
}

int dgb200_pack_sf_ue8m0_k_grouped(const float* sf, int32_t* out, int mn, const int32_t* ks_device, int num_groups,
                                   int packed_rows, int gran_k, int psum_alignment, void* stream) {
    if (int e = ensure_device()) return e;
    DGB_REQUIRE(mn > 0 && num_groups > 0 && (gran_k == 32 || gran_k == 128));
    if (packed_rows <= 0) return DGB200_OK;
    DGB_REQUIRE(packed_rows <= 65535);
    const dim3 grid(ceil_div(mn, 128), packed_rows);
    pack_sf_ue8m0_k_grouped_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        sf, reinterpret_cast<uint32_t*>(out), mn, ks_device, num_groups, gran_k, psum_alignment > 0 ? psum_alignment : 0);
    DGB_CUDA(cudaGetLastError());
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return DGB200_OK;
}

int dgb200_fp8_gemm_nt(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb, void* d, int m, int n,
                       int k, int64_t lda, int64_t ldb, int64_t ldd, int major_a, int major_b, int sfa_stride,
                       int sfb_stride, int gran_k_a, int gran_k_b, int d_dtype, int accumulate, void* workspace,
                       int64_t workspace_bytes, void* stream) {
    EnvScope env_scope;
    DGB_REQUIRE(m >= 0 && n >= 0 && k >= 0);
    if (m == 0 || n == 0) return DGB200_OK;  // gemm.hpp:22-23
    DGB_REQUIRE(k > 0);                      // k == 0 (D = C or 0) is handled by the host wrapper, gemm.hpp:36-40
    DGB_REQUIRE(d_dtype == DGB200_BF16 || d_dtype == DGB200_FP32);
    DGB_REQUIRE(major_a == DGB200_K_MAJOR || major_a == DGB200_MN_MAJOR);
    DGB_REQUIRE(major_b == DGB200_K_MAJOR || major_b == DGB200_MN_MAJOR);
    DGB_REQUIRE(lda >= (major_a == DGB200_K_MAJOR ? k : m) && ldb >= (major_b == DGB200_K_MAJOR ? k : n) && ldd >= n);
    DGB_REQUIRE(sfa_stride >= align_up(m, 4) && sfb_stride >= align_up(n, 4));
    GemmCall c{};
    c.type = kDense;
    c.a = a, c.b = b, c.sfa = sfa, c.sfb = sfb, c.d = d, c.grouped_layout = nullptr;
    c.x_mn = major_a == DGB200_MN_MAJOR, c.w_mn = major_b == DGB200_MN_MAJOR;
    c.m = m, c.n = n, c.k = k, c.groups = 1, c.a_rows = c.x_mn ? k : m;   // rows of the A tensor map's outer extent
    c.lda = lda, c.ldb = ldb, c.ldd = ldd;
    c.sfa_stride = sfa_stride, c.sfb_stride = sfb_stride;
    c.sfa_cols = align_up(m, 4), c.sfb_cols = align_up(n, 4);
    c.gran_k_a = gran_k_a, c.gran_k_b = gran_k_b;
    c.d_dtype = d_dtype, c.accumulate = accumulate != 0;
    c.expected_m = m, c.alignment = 1, c.zero_padding = 0;
    c.workspace = workspace, c.workspace_bytes = workspace_bytes > 0 ? static_cast<size_t>(workspace_bytes) : 0;
    c.stream = static_cast<cudaStream_t>(stream);
    if (int e = ensure_device()) return e;
    const int swap_bn = (!c.x_mn && !c.w_mn) ? want_swapped_orientation(c) : 0;
    if (swap_bn != 0) {
        // second orientation: tokens on the TMEM lanes, weights tiled freely along N (the kernel writes D[lane][column])
        std::swap(c.a, c.b), std::swap(c.sfa, c.sfb), std::swap(c.m, c.n), std::swap(c.lda, c.ldb);
        std::swap(c.sfa_stride, c.sfb_stride), std::swap(c.sfa_cols, c.sfb_cols), std::swap(c.gran_k_a, c.gran_k_b);
        c.a_rows = c.m, c.expected_m = c.m, c.swap_d = true;
        c.forced_block_m = swap_bn > 0 ? swap_bn : 0;
    }
    return run_gemm(c);
}

int dgb200_m_grouped_fp8_gemm_nt_contiguous(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb,
                                            void* d, const int32_t* grouped_layout, int num_groups, int m, int n,
                                            int k, int64_t lda, int64_t ldb, int64_t ldd, int major_b, int sfa_stride,
                                            int sfb_stride, int gran_k_a, int gran_k_b, int use_psum_layout,
                                            int ensure_zero_padding, int expected_m_for_psum_layout, void* stream) {
    DGB_REQUIRE(m >= 0);
    DGB_REQUIRE(n > 0 && k > 0 && num_groups > 0);  // gemm.hpp:192
    if (m == 0) return DGB200_OK;                    // gemm.hpp:210-211
    DGB_REQUIRE(grouped_layout != nullptr);
    DGB_REQUIRE(major_b == DGB200_K_MAJOR || major_b == DGB200_MN_MAJOR);
    DGB_REQUIRE(lda >= k && ldb >= (major_b == DGB200_K_MAJOR ? k : n) && ldd >= n);
    DGB_REQUIRE(sfa_stride >= align_up(m, 4) && sfb_stride >= align_up(n, 4));
    GemmCall c{};
    c.type = use_psum_layout ? kMContiguousPsum : kMContiguous;
    c.w_mn = major_b == DGB200_MN_MAJOR;
    c.a = a, c.b = b, c.sfa = sfa, c.sfb = sfb, c.d = d, c.grouped_layout = grouped_layout;
    c.m = m, c.n = n, c.k = k, c.groups = num_groups, c.a_rows = m;
    c.lda = lda, c.ldb = ldb, c.ldd = ldd;
    c.sfa_stride = sfa_stride, c.sfb_stride = sfb_stride;
    c.sfa_cols = align_up(m, 4), c.sfb_cols = align_up(n, 4);
    c.gran_k_a = gran_k_a, c.gran_k_b = gran_k_b;
    c.d_dtype = DGB200_BF16, c.accumulate = 0;
    c.expected_m = expected_m_for_psum_layout > 0 ? expected_m_for_psum_layout : ceil_div(m, num_groups);
    c.alignment = rt().mk_alignment;
    c.zero_padding = use_psum_layout && ensure_zero_padding;
    c.stream = static_cast<cudaStream_t>(stream);
    return run_gemm(c);
}

// ---- BF16 operands (no scale factors): the contiguous extent of each operand is byte-addressed (see fp8_gemm_kernel.cuh, kBf16AB)
static void bf16_common(GemmCall& c, int k, int64_t lda, int64_t ldb) {
    c.bf16_ab = true;
    c.sfa = c.sfb = nullptr, c.sfa_stride = c.sfb_stride = c.sfa_cols = c.sfb_cols = 4;
    c.gran_k_a = c.gran_k_b = 128;
    c.k = 2 * k, c.lda = 2 * lda, c.ldb = 2 * ldb;        // bytes
    c.batch_stride_a *= 2, c.batch_stride_b *= 2;
    c.workspace = nullptr, c.workspace_bytes = 0;
}

int dgb200_bf16_gemm_nt(const void* a, const void* b, void* d, int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd,
                        int major_a, int major_b, int d_dtype, int accumulate, void* stream) {
    DGB_REQUIRE(m >= 0 && n >= 0 && k >= 0);
    if (m == 0 || n == 0) return DGB200_OK;
    DGB_REQUIRE(k > 0);
    DGB_REQUIRE(d_dtype == DGB200_BF16 || d_dtype == DGB200_FP32);
    DGB_REQUIRE(major_a == DGB200_K_MAJOR || major_a == DGB200_MN_MAJOR);
    DGB_REQUIRE(major_b == DGB200_K_MAJOR || major_b == DGB200_MN_MAJOR);
    GemmCall c{};
    c.type = kDense;
    c.x_mn = major_a == DGB200_MN_MAJOR, c.w_mn = major_b == DGB200_MN_MAJOR;
    DGB_REQUIRE(lda >= (c.x_mn ? m : k) && ldb >= (c.w_mn ? n : k) && ldd >= n);
    DGB_REQUIRE((c.x_mn ? m : k) % 8 == 0 && (c.w_mn ? n : k) % 8 == 0);     // 16-byte rows for TMA
    c.a = a, c.b = b, c.d = d, c.grouped_layout = nullptr;
    c.m = m, c.n = n, c.groups = 1, c.a_rows = c.x_mn ? k : m, c.ldd = ldd;
    c.d_dtype = d_dtype, c.accumulate = accumulate != 0;
    c.expected_m = m, c.alignment = 1, c.zero_padding = 0;
    c.stream = static_cast<cudaStream_t>(stream);
    bf16_common(c, k, lda, ldb);
    return run_gemm(c);
}

int dgb200_m_grouped_bf16_gemm_nt_contiguous(const void* a, const void* b, void* d, const int32_t* grouped_layout, int num_groups,
                                             int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd, int major_b,
                                             int use_psum_layout, int ensure_zero_padding, int expected_m_for_psum_layout,
                                             void* stream) {
    DGB_REQUIRE(m >= 0);
    DGB_REQUIRE(n > 0 && k > 0 && k % 8 == 0 && num_groups > 0);
    if (m == 0) return DGB200_OK;
    DGB_REQUIRE(major_b == DGB200_K_MAJOR || major_b == DGB200_MN_MAJOR);
    GemmCall c{};
    c.w_mn = major_b == DGB200_MN_MAJOR;     // B [G, K, N]: N contiguous, ldb = pitch of a K row
    DGB_REQUIRE(grouped_layout != nullptr && lda >= k && ldb >= (c.w_mn ? n : k) && ldd >= n);
    DGB_REQUIRE(!c.w_mn || n % 8 == 0);
    c.type = use_psum_layout ? kMContiguousPsum : kMContiguous;
    c.a = a, c.b = b, c.d = d, c.grouped_layout = grouped_layout;
    c.m = m, c.n = n, c.groups = num_groups, c.a_rows = m, c.ldd = ldd;
    c.d_dtype = DGB200_BF16, c.accumulate = 0;
    c.expected_m = expected_m_for_psum_layout > 0 ? expected_m_for_psum_layout : ceil_div(m, num_groups);
    c.alignment = rt().mk_alignment;
    c.zero_padding = use_psum_layout && ensure_zero_padding;
    c.stream = static_cast<cudaStream_t>(stream);
    bf16_common(c, k, lda, ldb);
    return run_gemm(c);
}

int dgb200_m_grouped_bf16_gemm_nt_masked(const void* a, const void* b, void* d, const int32_t* masked_m, int num_groups, int m_max,
                                         int n, int k, int expected_m, void* stream) {
    DGB_REQUIRE(expected_m > 0 && m_max > 0 && n > 0 && k > 0 && k % 8 == 0 && num_groups > 0);
    DGB_REQUIRE(masked_m != nullptr);
    GemmCall c{};
    c.type = kMMasked;
    c.a = a, c.b = b, c.d = d, c.grouped_layout = masked_m;
    c.m = m_max, c.n = n, c.groups = num_groups, c.a_rows = num_groups * m_max, c.ldd = n;
    c.d_dtype = DGB200_BF16, c.accumulate = 0;
    c.expected_m = std::min(expected_m, m_max), c.alignment = 1, c.zero_padding = 0;
    c.stream = static_cast<cudaStream_t>(stream);
    bf16_common(c, k, k, k);
    return run_gemm(c);
}

int dgb200_bf16_bmm(const void* a, const void* b, void* d, int batch, int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd,
                    int64_t batch_stride_a, int64_t batch_stride_b, int64_t batch_stride_d, int major_b, void* stream) {
    DGB_REQUIRE(batch >= 0 && m >= 0 && n >= 0 && k > 0);
    if (batch == 0 || m == 0 || n == 0) return DGB200_OK;
    DGB_REQUIRE(major_b == DGB200_K_MAJOR || major_b == DGB200_MN_MAJOR);
    GemmCall c{};
    c.type = kBatched;
    c.w_mn = major_b == DGB200_MN_MAJOR;
    DGB_REQUIRE(lda >= k && ldb >= (c.w_mn ? n : k) && ldd >= n);
    DGB_REQUIRE(k % 8 == 0 && (!c.w_mn || n % 8 == 0));
    DGB_REQUIRE(batch_stride_a % 8 == 0 && batch_stride_b % 8 == 0 && batch_stride_a > 0 && batch_stride_b > 0 && batch_stride_d > 0);
    c.a = a, c.b = b, c.d = d, c.grouped_layout = nullptr;
    c.m = m, c.n = n, c.groups = batch, c.a_rows = m, c.ldd = ldd;
    c.batch_stride_a = batch_stride_a, c.batch_stride_b = batch_stride_b, c.batch_stride_d = batch_stride_d;
    c.d_dtype = DGB200_BF16, c.accumulate = 0;
    c.expected_m = m, c.alignment = 1, c.zero_padding = 0;
    c.stream = static_cast<cudaStream_t>(stream);
    bf16_common(c, k, lda, ldb);
    return run_gemm(c);
}

int dgb200_bf16_bmk_bnk_mn(const void* a, const void* b, float* d, int batch, int m, int n, int k, void* stream) {
    DGB_REQUIRE(batch >= 0 && m >= 0 && n >= 0 && k >= 0);
    if (batch == 0 || m == 0 || n == 0 || k == 0) return DGB200_OK;       // nothing to add
    DGB_REQUIRE(k % 64 == 0);                                            // whole 128-byte k-blocks per batch
    GemmCall c{};
    c.type = kBatchReduce;
    c.a = a, c.b = b, c.d = d, c.grouped_layout = nullptr;
    c.m = m, c.n = n, c.groups = batch, c.a_rows = m, c.ldd = n;
    c.batch_stride_a = static_cast<int64_t>(m) * k, c.batch_stride_b = static_cast<int64_t>(n) * k, c.batch_stride_d = 0;
    c.d_dtype = DGB200_FP32, c.accumulate = 1;
    c.expected_m = m, c.alignment = 1, c.zero_padding = 0;
    c.stream = static_cast<cudaStream_t>(stream);
    bf16_common(c, k, k, k);
    return run_gemm(c);
}

int dgb200_k_grouped_bf16_gemm_tn_contiguous(const void* a, const void* b, float* d, const int32_t* grouped_layout, int num_groups,
                                             int m, int n, int sum_k, int use_psum_layout, void* stream) {
    DGB_REQUIRE(num_groups > 0 && m >= 0 && n >= 0 && sum_k >= 0);
    if (m == 0 || n == 0 || sum_k == 0) return DGB200_OK;            // D already holds C (gemm.hpp:592-593)
    DGB_REQUIRE(grouped_layout != nullptr);
    DGB_REQUIRE(m % 8 == 0 && n % 8 == 0);                           // 16-byte rows of the [sum_k, m] / [sum_k, n] operands
    const int k_alignment = rt().mk_alignment;
    DGB_REQUIRE(k_alignment % 32 == 0);                              // gemm.hpp:580
    GemmCall c{};
    c.type = use_psum_layout ? kKGroupedPsum : kKGrouped;
    c.x_mn = c.w_mn = true;
    c.a = a, c.b = b, c.d = d, c.grouped_layout = grouped_layout;
    c.m = m, c.n = n, c.groups = num_groups, c.a_rows = sum_k, c.ldd = n;
    c.d_dtype = DGB200_FP32, c.accumulate = 1;
    c.expected_m = m, c.alignment = k_alignment, c.zero_padding = 0;
    c.stream = static_cast<cudaStream_t>(stream);
    bf16_common(c, sum_k, m, n);
    return run_gemm(c);
}

int dgb200_fp8_gemm_nt_skip_head_mid(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb, void* d, int m,
                                     int n, int k, int64_t lda, int64_t ldb, int64_t ldd, int head_left, int head_mid,
                                     int head_right, int sfa_stride, int sfb_stride, int d_dtype, void* stream) {
    DGB_REQUIRE(m >= 0 && n > 0 && k > 0);                                  // attention.hpp:42
    DGB_REQUIRE(head_left >= 0 && head_mid >= 0 && head_right >= 0 && head_left + head_right > 0);
    DGB_REQUIRE(n % (head_left + head_right) == 0);                         // attention.hpp:49
    const int n_out = n + n / (head_left + head_right) * head_mid;
    DGB_REQUIRE(d_dtype == DGB200_BF16 || d_dtype == DGB200_FP32);
    DGB_REQUIRE(lda >= k && ldb >= k && ldd >= n_out);
    if (m == 0) return DGB200_OK;                                           // attention.hpp:52-53
    DGB_REQUIRE(sfa_stride >= align_up(m, 4) && sfb_stride >= align_up(n, 4));
    GemmCall c{};
    c.type = kDense;
    c.a = a, c.b = b, c.sfa = sfa, c.sfb = sfb, c.d = d, c.grouped_layout = nullptr;
    c.m = m, c.n = n, c.k = k, c.groups = 1, c.a_rows = m;
    c.lda = lda, c.ldb = ldb, c.ldd = ldd;
    c.sfa_stride = sfa_stride, c.sfb_stride = sfb_stride;
    c.sfa_cols = align_up(m, 4), c.sfb_cols = align_up(n, 4);
    c.gran_k_a = 128, c.gran_k_b = 128;                                     // attention.hpp:59
    c.d_dtype = d_dtype, c.accumulate = 0;
    c.expected_m = m, c.alignment = 1, c.zero_padding = 0;
    c.workspace = nullptr, c.workspace_bytes = 0;
    c.head_left = head_left, c.head_mid = head_mid, c.head_right = head_right;
    c.stream = static_cast<cudaStream_t>(stream);
    return run_gemm(c);
}

int dgb200_fp8_bmm(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb, void* d, int batch, int m, int n,
                   int k, int64_t lda, int64_t ldb, int64_t ldd, int64_t batch_stride_a, int64_t batch_stride_b,
                   int64_t batch_stride_d, int major_a, int major_b, int sfa_stride, int sfb_stride, int gran_k_a,
                   int gran_k_b, int d_dtype, int accumulate, void* stream) {
    DGB_REQUIRE(batch >= 0 && m >= 0 && n >= 0 && k >= 0);
    if (batch == 0 || m == 0 || n == 0) return DGB200_OK;                    // einsum.hpp:160-161
    DGB_REQUIRE(k > 0);                                                     // k == 0 is the host wrapper's job (gemm.hpp:36-40)
    DGB_REQUIRE(d_dtype == DGB200_BF16 || d_dtype == DGB200_FP32);
    DGB_REQUIRE(major_a == DGB200_K_MAJOR || major_a == DGB200_MN_MAJOR);
    DGB_REQUIRE(major_b == DGB200_K_MAJOR || major_b == DGB200_MN_MAJOR);
    DGB_REQUIRE(lda >= (major_a == DGB200_K_MAJOR ? k : m) && ldb >= (major_b == DGB200_K_MAJOR ? k : n) && ldd >= n);
    DGB_REQUIRE(batch_stride_a % 16 == 0 && batch_stride_b % 16 == 0 && batch_stride_a > 0 && batch_stride_b > 0 && batch_stride_d > 0);
    DGB_REQUIRE(sfa_stride >= align_up(m, 4) && sfb_stride >= align_up(n, 4));
    GemmCall c{};
    c.type = kBatched;
    c.a = a, c.b = b, c.sfa = sfa, c.sfb = sfb, c.d = d, c.grouped_layout = nullptr;
    c.x_mn = major_a == DGB200_MN_MAJOR, c.w_mn = major_b == DGB200_MN_MAJOR;
    c.m = m, c.n = n, c.k = k, c.groups = batch, c.a_rows = c.x_mn ? k : m;
    c.lda = lda, c.ldb = ldb, c.ldd = ldd;
    c.batch_stride_a = batch_stride_a, c.batch_stride_b = batch_stride_b, c.batch_stride_d = batch_stride_d;
    c.sfa_stride = sfa_stride, c.sfb_stride = sfb_stride;
    c.sfa_cols = align_up(m, 4), c.sfb_cols = align_up(n, 4);
    c.gran_k_a = gran_k_a, c.gran_k_b = gran_k_b;
    c.d_dtype = d_dtype, c.accumulate = accumulate != 0;
    c.expected_m = m, c.alignment = 1, c.zero_padding = 0;
    c.workspace = nullptr, c.workspace_bytes = 0;
    c.stream = static_cast<cudaStream_t>(stream);
    return run_gemm(c);
}

int dgb200_per_token_cast_to_fp8(const void* x, int64_t ldx, void* q, int64_t ldq, int32_t* sf, int sf_stride, int m, int k,
                                 int gran_k, void* stream) {
    if (int e = ensure_device()) return e;
    DGB_REQUIRE(m >= 0 && k >= 0 && (gran_k == 32 || gran_k == 128));
    if (m == 0 || k == 0) return DGB200_OK;
    DGB_REQUIRE(x != nullptr && q != nullptr && sf != nullptr && ldx >= k && ldq >= k && sf_stride >= align_up(m, 4));
    DGB_REQUIRE(ceil_div(m, 8) <= 65535);
    const dim3 grid(ceil_div(k, 512), ceil_div(m, 8));
    const auto s = static_cast<cudaStream_t>(stream);
    if (gran_k == 128)
        per_token_cast_to_fp8_kernel<128><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx, static_cast<uint8_t*>(q), ldq,
                                                               reinterpret_cast<uint32_t*>(sf), sf_stride, m, k);
    else
        per_token_cast_to_fp8_kernel<32><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx, static_cast<uint8_t*>(q), ldq,
                                                              reinterpret_cast<uint32_t*>(sf), sf_stride, m, k);
    DGB_CUDA(cudaGetLastError());
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return DGB200_OK;
}

int dgb200_debug_fp8_peak(int umma_n, int iters, int num_sms, void* stream) {
    if (int e = ensure_device()) return e;
    DGB_REQUIRE(umma_n >= 16 && umma_n <= 256 && umma_n % 16 == 0 && iters > 0);
    const int sms = (num_sms > 0 ? std::min(num_sms, rt().sm_count) : rt().sm_count) & ~1;
    DGB_REQUIRE(sms >= 2);
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(sms, 1, 1), lc.blockDim = dim3(128, 1, 1);
    lc.dynamicSmemBytes = 36 * 1024;
    lc.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    lc.attrs = &attr, lc.numAttrs = 1;
    DGB_CUDA(cudaLaunchKernelEx(&lc, fp8_mma_peak_kernel, static_cast<uint32_t>(umma_n), static_cast<uint32_t>(iters),
                                static_cast<uint32_t*>(nullptr)));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return DGB200_OK;
}

int dgb200_ep_combine(void* out, int64_t ldo, const int32_t* token_row, const void* expert_ids, int id_bytes, int num_tokens,
                      int topk, const float* weights, int n, int elt_bytes, int num_experts, int rank, int world,
                      void* const* buffers, void* const* d_buffers, int64_t ldd, void* stream) {
    if (int e = ensure_device()) return e;
    DGB_REQUIRE(world > 0 && world <= static_cast<int>(ep::kMaxWorld) && rank >= 0 && rank < world);
    DGB_REQUIRE(num_experts > 0 && num_experts % world == 0 && num_tokens >= 0 && n > 0 && topk >= 1);
    DGB_REQUIRE(id_bytes == 4 || id_bytes == 8);
    DGB_REQUIRE(elt_bytes == 2 || (elt_bytes == 4 && topk == 1 && weights == nullptr));   // the weighted reduce is BF16 -> FP32 -> BF16
    DGB_REQUIRE(buffers != nullptr && d_buffers != nullptr);
    DGB_REQUIRE(num_tokens == 0 || (out != nullptr && token_row != nullptr && expert_ids != nullptr));
    DGB_REQUIRE((static_cast<int64_t>(n) * elt_bytes) % 16 == 0 && (ldo * elt_bytes) % 16 == 0 && (ldd * elt_bytes) % 16 == 0);
    DGB_REQUIRE(ldo >= n && ldd >= n && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
    ep::Peers ctrl, dbufs;
    for (int p = 0; p < world; ++p) {
        DGB_REQUIRE(buffers[p] != nullptr && d_buffers[p] != nullptr && (reinterpret_cast<uintptr_t>(d_buffers[p]) & 15) == 0);
        ctrl.base[p] = static_cast<uint8_t*>(buffers[p]);
        dbufs.base[p] = static_cast<uint8_t*>(d_buffers[p]);
    }
    const auto s = static_cast<cudaStream_t>(stream);
    ep::combine_publish_kernel<<<1, 32, 0, s>>>(ctrl, rank, world);
    const int grid = std::max(1, std::min(ceil_div(std::max(num_tokens, 1), 8), rt().sm_count * 8));
    if (id_bytes == 4)
        ep::combine_gather_kernel<int32_t><<<grid, 256, 0, s>>>(ctrl, dbufs, expert_ids, token_row, weights, topk,
                                                                static_cast<uint8_t*>(out), ldo * elt_bytes, ldd * elt_bytes,
                                                                n * elt_bytes, num_tokens, num_experts, rank, world);
    else
        ep::combine_gather_kernel<int64_t><<<grid, 256, 0, s>>>(ctrl, dbufs, expert_ids, token_row, weights, topk,
                                                                static_cast<uint8_t*>(out), ldo * elt_bytes, ldd * elt_bytes,
                                                                n * elt_bytes, num_tokens, num_experts, rank, world);
    DGB_CUDA(cudaGetLastError());
    g_launch_count.fetch_add(2, std::memory_order_relaxed);
    return DGB200_OK;
}

int dgb200_ep_grouped_gemm(void* local_buffer, int world, int num_experts, int capacity, int k, const void* b,
                           const int32_t* sfb, void* d, int n, int64_t ldb, int64_t ldd, int major_b, int sfb_stride,
                           int gran_k_b, int expected_m, int overlap_dispatch, void* stream) {
    DGB_REQUIRE(local_buffer != nullptr && world > 0 && num_experts > 0 && num_experts % world == 0 && capacity > 0);
    DGB_REQUIRE(n > 0 && k > 0 && b != nullptr && sfb != nullptr && d != nullptr);
    DGB_REQUIRE(major_b == DGB200_K_MAJOR || major_b == DGB200_MN_MAJOR);
    DGB_REQUIRE(ldb >= (major_b == DGB200_K_MAJOR ? k : n) && ldd >= n && sfb_stride >= align_up(n, 4));
    const ep::Layout l = ep::make_layout(world, num_experts, capacity, k);
    uint8_t* base = static_cast<uint8_t*>(local_buffer);
    GemmCall c{};
    c.type = kMContiguousPsum;
    c.w_mn = major_b == DGB200_MN_MAJOR;
    c.a = base + l.a_off, c.sfa = reinterpret_cast<const int32_t*>(base + l.sfa_off), c.b = b, c.sfb = sfb, c.d = d;
    c.grouped_layout = reinterpret_cast<const int32_t*>(base + l.psum_off);
    c.m = capacity, c.n = n, c.k = k, c.groups = num_experts / world, c.a_rows = capacity;
    c.lda = k, c.ldb = ldb, c.ldd = ldd;
    c.sfa_stride = capacity, c.sfb_stride = sfb_stride;
    c.sfa_cols = align_up(capacity, 4), c.sfb_cols = align_up(n, 4);
    c.gran_k_a = 128, c.gran_k_b = gran_k_b;
    c.d_dtype = DGB200_BF16, c.accumulate = 0;
    c.expected_m = expected_m > 0 ? expected_m : ceil_div(capacity, c.groups);
    c.alignment = rt().mk_alignment;
    c.zero_padding = 1;
    if (overlap_dispatch) {
        c.arrival = reinterpret_cast<const uint32_t*>(base + l.arrived_off);
        c.arrival_expected = reinterpret_cast<const uint32_t*>(base + l.expected_off);
    }
    c.stream = static_cast<cudaStream_t>(stream);
    return run_gemm(c);
}

int dgb200_m_grouped_fp8_gemm_nt_masked(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb, void* d,
                                        const int32_t* masked_m, int num_groups, int m_max, int n, int k,
                                        int expected_m, int sfa_stride, int sfb_stride, int gran_k_a, int gran_k_b,
                                        void* stream) {
    DGB_REQUIRE(expected_m > 0 && m_max > 0 && n > 0 && k > 0 && num_groups > 0);  // gemm.hpp:274
    DGB_REQUIRE(masked_m != nullptr);
    DGB_REQUIRE(sfa_stride >= align_up(m_max, 4) && sfb_stride >= align_up(n, 4));
    GemmCall c{};
    c.type = kMMasked;
    c.a = a, c.b = b, c.sfa = sfa, c.sfb = sfb, c.d = d, c.grouped_layout = masked_m;
    c.m = m_max, c.n = n, c.k = k, c.groups = num_groups, c.a_rows = num_groups * m_max;
    c.lda = k, c.ldb = k, c.ldd = n;
    c.sfa_stride = sfa_stride, c.sfb_stride = sfb_stride;
    c.sfa_cols = align_up(m_max, 4), c.sfb_cols = align_up(n, 4);
    c.gran_k_a = gran_k_a, c.gran_k_b = gran_k_b;
    c.d_dtype = DGB200_BF16, c.accumulate = 0;
    c.expected_m = std::min(expected_m, m_max), c.alignment = 1, c.zero_padding = 0;
    c.stream = static_cast<cudaStream_t>(stream);
    return run_gemm(c);
}

int dgb200_k_grouped_fp8_gemm_tn_contiguous(const void* a, const int32_t* sfa, const void* b, const int32_t* sfb,
                                            float* d, const int32_t* grouped_layout, int num_groups, int m, int n,
                                            int sum_k, int sf_rows, int gran_k, int use_psum_layout, void* stream) {
    DGB_REQUIRE(gran_k == 32 || gran_k == 128);                      // gemm.hpp:313
    DGB_REQUIRE(num_groups > 0 && m >= 0 && n >= 0 && sum_k >= 0);
    if (m == 0 || n == 0 || sum_k == 0) return DGB200_OK;            // D already holds C (gemm.hpp:334-335)
    DGB_REQUIRE(grouped_layout != nullptr && sf_rows > 0);
    DGB_REQUIRE(m % 4 == 0 && n % 4 == 0);                           // packed k-grouped SFs are [rows, mn] contiguous
    const int k_alignment = rt().mk_alignment;
    DGB_REQUIRE(k_alignment % 32 == 0);                              // gemm.hpp:315
    GemmCall c{};
    c.type = use_psum_layout ? kKGroupedPsum : kKGrouped;
    c.x_mn = c.w_mn = true;
    c.a = a, c.b = b, c.sfa = sfa, c.sfb = sfb, c.d = d, c.grouped_layout = grouped_layout;
    c.m = m, c.n = n, c.k = sum_k, c.groups = num_groups, c.a_rows = sum_k;
    c.lda = m, c.ldb = n, c.ldd = n;
    c.sfa_stride = m, c.sfb_stride = n, c.sfa_cols = m, c.sfb_cols = n;
    c.sfa_krows = sf_rows, c.sfb_krows = sf_rows;
    c.gran_k_a = gran_k, c.gran_k_b = gran_k;
    c.d_dtype = DGB200_FP32, c.accumulate = 1;
    c.expected_m = m, c.alignment = k_alignment, c.zero_padding = 0;
    c.workspace = nullptr, c.workspace_bytes = 0;
    c.stream = static_cast<cudaStream_t>(stream);
    return run_gemm(c);
}

int dgb200_plan(int gemm_type, int m, int n, int k, int num_groups, int expected_m, int alignment, int num_sms,
                dgb200_config* out) {
    EnvScope env_scope;
    DGB_REQUIRE(out != nullptr && num_sms >= 2);
    DGB_REQUIRE(gemm_type >= kDense && gemm_type <= kMContiguousPsum);
    DGB_REQUIRE(m > 0 && n > 0 && k > 0 && num_groups > 0);
    Problem pb{gemm_type, m, expected_m > 0 ? expected_m : m, n, k, num_groups, std::max(alignment, 1)};
    if (gemm_type == kDense && n % 4 == 0) pb.max_splits = kMaxSplits;   // as if a workspace were supplied
    pb.tma_store_ok = gemm_type == kDense || gemm_type == kMContiguous;   // as if D were an aligned BF16 tensor
    Config cfg = choose_config(pb, num_sms);
    if (gemm_type == kDense && n % 8 == 0) {
        // the orientation rule of dgb200_fp8_gemm_nt, for a BF16 output with an aligned row pitch
        GemmCall probe{};
        probe.m = m, probe.n = n, probe.k = k, probe.d_dtype = DGB200_BF16, probe.accumulate = 0, probe.d = nullptr, probe.ldd = n;
        const int bn = want_swapped_orientation(probe, num_sms);
        if (bn != 0) {
            Problem sw{kDense, n, n, m, k, 1, 1};
            sw.swapped = true, sw.tma_store_ok = true, sw.forced_block_m = bn > 0 ? bn : 0;
            cfg = choose_config(sw, num_sms);
            *out = dgb200_config{cfg.block_m, cfg.cluster, cfg.stages, cfg.num_sms, cfg.smem_bytes,
                                 ceil_div(n, cfg.block_m) * ceil_div(m, (int)kBlockN * std::min(cfg.cluster, 2)), 1, 0, cfg.tma_store, 1};
            return DGB200_OK;
        }
    }
    const int n_units = ceil_div(n, (int)kBlockN * (cfg.csplit ? cfg.cluster / cfg.csplit : std::min(cfg.cluster, 2)));
    int m_blocks = gemm_type == kMMasked ? num_groups * ceil_div(pb.expected_m, cfg.block_m) : ceil_div(m, cfg.block_m);
    if (gemm_type == kDense && cfg.block_m_low > 0)   // two tile heights (wave balancing)
        m_blocks = cfg.num_tall + ceil_div(std::max(0, m - cfg.num_tall * cfg.block_m), cfg.block_m_low);
    *out = dgb200_config{cfg.block_m, cfg.cluster, cfg.stages, cfg.num_sms, cfg.smem_bytes, m_blocks * n_units * cfg.num_splits,
                         cfg.num_splits, cfg.csplit, cfg.tma_store, cfg.swap_d};
    return DGB200_OK;
}

int dgb200_debug_set_timestamps(void* device_int64_buffer) {
    g_debug_ts.store(static_cast<long long*>(device_int64_buffer));
    return DGB200_OK;
}

int64_t dgb200_workspace_bytes(int m, int n) {
    if (m <= 0 || n <= 0) return 0;
    return static_cast<int64_t>(kSplitKHeaderBytes) + static_cast<int64_t>(kMaxSplits) * m * n * sizeof(float);
}

// ------------------------------------------------------------------------------------------------ expert-parallel dispatch
int64_t dgb200_ep_buffer_bytes(int world, int num_experts, int capacity, int k) {
    if (world <= 0 || num_experts <= 0 || capacity <= 0 || k <= 0) return 0;
    return static_cast<int64_t>(ep::make_layout(world, num_experts, capacity, k).total);
}

int dgb200_ep_buffer_offsets(int world, int num_experts, int capacity, int k, int64_t* offsets) {
    DGB_REQUIRE(world > 0 && num_experts > 0 && capacity > 0 && k > 0 && offsets != nullptr);
    const ep::Layout l = ep::make_layout(world, num_experts, capacity, k);
    offsets[DGB200_EP_OFF_A] = l.a_off;
    offsets[DGB200_EP_OFF_SFA] = l.sfa_off;
    offsets[DGB200_EP_OFF_PSUM] = l.psum_off;
    offsets[DGB200_EP_OFF_COUNTS] = l.counts_off;
    offsets[DGB200_EP_OFF_NUM_ROWS] = offsetof(ep::Control, num_rows);
    offsets[DGB200_EP_OFF_OVERFLOW] = offsetof(ep::Control, overflow);
    return DGB200_OK;
}

int dgb200_ep_alloc(int64_t bytes, void** ptr) {
    DGB_REQUIRE(bytes > 0 && ptr != nullptr);
    DGB_CUDA(cudaMalloc(ptr, bytes));
    DGB_CUDA(cudaMemset(*ptr, 0, bytes));
    DGB_CUDA(cudaDeviceSynchronize());
    return DGB200_OK;
}
int dgb200_ep_free(void* ptr) {
    DGB_CUDA(cudaFree(ptr));
    return DGB200_OK;
}
int dgb200_ep_export(void* ptr, void* handle_64_bytes) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DGB_REQUIRE(ptr != nullptr && handle_64_bytes != nullptr);
    cudaIpcMemHandle_t h;
    DGB_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_64_bytes, &h, 64);
    return DGB200_OK;
}
int dgb200_ep_import(const void* handle_64_bytes, void** ptr) {
    DGB_REQUIRE(ptr != nullptr && handle_64_bytes != nullptr);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64_bytes, 64);
    DGB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DGB200_OK;
}
int dgb200_ep_unimport(void* ptr) {
    DGB_CUDA(cudaIpcCloseMemHandle(ptr));
    return DGB200_OK;
}

int dgb200_ep_dispatch(const void* x, int64_t ldx, const int32_t* sf, int64_t sf_stride_t, int64_t sf_stride_k,
                       const void* expert_ids, int id_bytes, int num_tokens, int topk, int k, int num_experts, int rank,
                       int world, void* const* buffers, int capacity, int alignment, int32_t* token_row,
                       int32_t* order_scratch, int wait_for_all, void* stream) {
    DGB_REQUIRE(world > 0 && world <= static_cast<int>(ep::kMaxWorld) && rank >= 0 && rank < world);
    DGB_REQUIRE(num_experts > 0 && num_experts <= static_cast<int>(ep::kMaxExperts) && num_experts % world == 0);
    DGB_REQUIRE(num_tokens >= 0 && topk >= 1 && capacity > 0 && capacity % 4 == 0 && alignment > 0);   // SF pitch = capacity words (TMA: 16 B)
    DGB_REQUIRE(static_cast<int64_t>(num_tokens) * topk < (1ll << 31));
    DGB_REQUIRE(k > 0 && k % 16 == 0 && ceil_div(k, 512) <= 32);
    DGB_REQUIRE(id_bytes == 4 || id_bytes == 8);
    const int num_entries = num_tokens * topk;
    DGB_REQUIRE(buffers != nullptr && token_row != nullptr && (wait_for_all || num_tokens == 0 || order_scratch != nullptr));
    DGB_REQUIRE(num_tokens == 0 || (x != nullptr && sf != nullptr && expert_ids != nullptr));
    DGB_REQUIRE(ldx % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
    DGB_REQUIRE(wait_for_all || topk == 1);            // the dispatch || GEMM mode is built for top-1 routing
    if (int e = ensure_device()) return e;
    ep::Peers peers;
    for (int p = 0; p < world; ++p) {
        DGB_REQUIRE(buffers[p] != nullptr);
        peers.base[p] = static_cast<uint8_t*>(buffers[p]);
    }
    const ep::Layout l = ep::make_layout(world, num_experts, capacity, k);
    const auto s = static_cast<cudaStream_t>(stream);
    uint8_t* mine = peers.base[rank];
    const uint32_t kp = ceil_div(k, 512);
    const auto* xb = static_cast<const uint8_t*>(x);
    if (wait_for_all) {
        // One persistent kernel; every CTA must be resident (it synchronises through grid barriers), so the grid is cut to
        // what the occupancy API says fits, and to the work there is.
        static std::mutex mu;
        static int resident[2] = {0, 0};
        int fit;
        {
            std::lock_guard<std::mutex> lock(mu);
            int& r = resident[id_bytes == 8];
            if (r == 0) {
                int per_sm = 0;
                if (id_bytes == 4)
                    DGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ep::dispatch_fused_kernel<int32_t>, ep::kFusedThreads, 0));
                else
                    DGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ep::dispatch_fused_kernel<int64_t>, ep::kFusedThreads, 0));
                r = std::max(1, per_sm) * rt().sm_count;
            }
            fit = r;
        }
        const int warps = ep::kFusedThreads / 32;
        const int grid = std::max(1, std::min({fit, (int)ep::kMaxRankCtas, ceil_div(std::max(num_entries, 1), warps)}));
        const int slice = std::max((int)ep::kFusedThreads, align_up(ceil_div(std::max(num_entries, 1), grid), (int)ep::kFusedThreads));
        if (id_bytes == 4)
            ep::dispatch_fused_kernel<int32_t><<<grid, ep::kFusedThreads, 0, s>>>(peers, l, xb, ldx, sf, sf_stride_t, sf_stride_k, expert_ids,
                                                                                  token_row, num_entries, topk, slice, k, kp, rank, world,
                                                                                  num_experts, capacity, alignment);
        else
            ep::dispatch_fused_kernel<int64_t><<<grid, ep::kFusedThreads, 0, s>>>(peers, l, xb, ldx, sf, sf_stride_t, sf_stride_k, expert_ids,
                                                                                  token_row, num_entries, topk, slice, k, kp, rank, world,
                                                                                  num_experts, capacity, alignment);
        DGB_CUDA(cudaGetLastError());
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        return DGB200_OK;
    }
    // dispatch || GEMM: the round-1 chain. Expert-sorted send order, and a persistent single wave of 4 scatter CTAs per SM
    // with <= 32 registers, so that every CTA is resident from the start and the consumer's CTAs (384 threads, 1 per SM) fit beside.
    int32_t* counts = reinterpret_cast<int32_t*>(mine + l.counts_off);
    if (id_bytes == 4)
        ep::bucket_kernel<int32_t><<<num_experts, 1024, 0, s>>>(expert_ids, num_tokens, token_row, counts);
    else
        ep::bucket_kernel<int64_t><<<num_experts, 1024, 0, s>>>(expert_ids, num_tokens, token_row, counts);
    const int grid = std::max(1, std::min(ceil_div(num_tokens, 8), rt().sm_count * 4));
    ep::exchange_kernel<<<1, 1024, 0, s>>>(peers, l, rank, world, num_experts, capacity, alignment, true);
    if (num_tokens > 0) {
        if (id_bytes == 4)
            ep::order_kernel<int32_t><<<ceil_div(num_tokens, 256), 256, 0, s>>>(mine, l, expert_ids, num_tokens, num_experts, token_row, order_scratch);
        else
            ep::order_kernel<int64_t><<<ceil_div(num_tokens, 256), 256, 0, s>>>(mine, l, expert_ids, num_tokens, num_experts, token_row, order_scratch);
    }
    if (id_bytes == 4)
        ep::scatter_kernel<int32_t, true><<<grid, 256, 0, s>>>(peers, l, xb, ldx, sf, sf_stride_t, sf_stride_k, expert_ids, token_row,
                                                               order_scratch, num_tokens, k, kp, rank, world, num_experts, capacity);
    else
        ep::scatter_kernel<int64_t, true><<<grid, 256, 0, s>>>(peers, l, xb, ldx, sf, sf_stride_t, sf_stride_k, expert_ids, token_row,
                                                               order_scratch, num_tokens, k, kp, rank, world, num_experts, capacity);
    DGB_CUDA(cudaGetLastError());
    g_launch_count.fetch_add(num_tokens > 0 ? 4 : 3, std::memory_order_relaxed);
    return DGB200_OK;
}

int dgb200_last_config(dgb200_config* out) {
    if (!out) return DGB200_ERR_INVALID_ARGUMENT;
    *out = g_last_config;
    return DGB200_OK;
}
int64_t dgb200_launch_count(void) { return g_launch_count.load(); }

}  // extern "C"
