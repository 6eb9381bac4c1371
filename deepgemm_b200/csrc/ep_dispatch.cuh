// Expert-parallel dispatch over NVLink peer memory (plain CUDA, HBM / NVLink bound byte movement).
//
// Where the m-grouped contiguous GEMM sits inside expert parallelism (the reference's own baseline pairs it with a
// DeepEP dispatch, tests/test_mega_moe.py:148-205), the tokens must first travel to the rank that owns their expert and
// land in the contiguous-grouped layout (expert segments aligned to the M alignment, "psum" end rows per expert,
// MN-major packed UE8M0 scale factors). Instead of all-to-all + re-layout passes, every rank writes its rows STRAIGHT
// into the destination rank's GEMM input buffer with peer stores (NVSwitch: every peer at full bandwidth):
//
//   ONE persistent kernel (`dispatch_fused_kernel`, the default path), phases separated by grid barriers:
//     rank     : the first R CTAs each take a slice of the (token, slot) entries: stable rank of every entry inside its
//                expert within the slice + a per-slice histogram -- O(T) in total (round 1 had G CTAs re-read all T ids)
//     exchange : CTA 0 sums the histograms, publishes my per-expert counts into every peer's table (peer stores +
//                release flag), waits for theirs and derives, for every expert, the first destination row of MY tokens
//                (segment start on the owner + rows of lower source ranks) and the local psum layout; meanwhile the
//                ranking CTAs turn slice-local ranks into rank-within-source-GPU (prefix over the slice histograms)
//     scatter  : all CTAs, one warp per entry: 16-byte loads from local HBM, 16-byte stores into the owner's A buffer
//                over NVLink, scale-factor words into the owner's MN-major SF buffer
//     complete : the last CTA to finish tells every peer "my rows have landed" and waits for every peer's message, so
//                the kernel's end means the local GEMM input is complete
//   (The round-1 chain bucket -> exchange -> [order] -> scatter -> wait is kept for the dispatch || GEMM mode, where the
//   scatter kernel must be the direct programmatic predecessor of the GEMM.)
//
// No host synchronisation anywhere (counts never leave the devices), so the whole step is CUDA-graph capturable.
// Flags are monotonic epochs kept in device memory; buffer reuse across steps is ordered by the exchange of the next
// step (a rank publishes its counts only after its previous GEMM finished in stream order, and nobody scatters before
// it has seen every rank's counts).
#pragma once
#include <cuda_bf16.h>

#include <cstdint>

namespace dgb200 {
namespace ep {

constexpr uint32_t kMaxWorld = 16;
constexpr uint32_t kMaxExperts = 2048;
constexpr uint64_t kWaitTimeoutNs = 30ull * 1000 * 1000 * 1000;
constexpr uint32_t kMaxRankCtas = 512;       // fused dispatch: upper bound on the ranking CTAs (= on its grid)
constexpr uint32_t kFusedThreads = 512;

// Control block at the start of every rank's buffer (all offsets identical on all ranks).
struct Control {
    uint32_t epoch;               // last completed dispatch (local)
    uint32_t done_ctas;           // scatter CTAs finished (local)
    uint32_t num_rows;            // rows of the local A buffer in use after the last dispatch (aligned end of the last expert)
    uint32_t overflow;            // set when a dispatch would not fit `capacity`
    uint32_t num_routed;          // local tokens with a valid expert in the last dispatch
    uint32_t grid_bar;            // fused dispatch: grid barrier arrivals (reset by the last CTA of every launch)
    uint32_t pad[2];
    uint64_t dbg_ns[12];          // globaltimer stamps of the last fused dispatch [0..7] and combine [8..10] (development: tools/ep_phases.py)
    uint32_t counts_flag[kMaxWorld * 8];   // [s*8]: epoch of the counts source rank s published here (32 B apart)
    uint32_t data_flag[kMaxWorld * 8];     // [s*8]: epoch of the rows source rank s finished writing here
    uint32_t gemm_flag[kMaxWorld * 8];     // [o*8]: epoch of the grouped GEMM owner rank o has finished (combine)
};

struct Peers {
    uint8_t* base[kMaxWorld];
};

struct Layout {
    uint64_t table_off;     // int32 [2][world][num_experts] counts published by every source rank (double-buffered by
                            //                             epoch parity: a fast rank may publish step N+1 while a slow
                            //                             one still reads step N)
    uint64_t hist_off;      // int32 [kMaxRankCtas][num_experts] fused dispatch: per-slice histograms (local)
    uint64_t dst_base_off;  // int32 [num_experts]         first destination row of MY tokens for each expert (local)
    uint64_t psum_off;      // int32 [experts_per_rank]    end row of each local expert segment (the GEMM's psum layout)
    uint64_t counts_off;    // int32 [num_experts]         my own per-expert token counts (local)
    uint64_t sorted_off;    // int32 [num_experts]         position of my first token of each expert in expert-sorted order (local)
    uint64_t arrived_off;   // uint32 [experts_per_rank]   rows landed for each local expert, cumulative over all dispatches
                            //                             (incremented by the sources with remote atomics)
    uint64_t expected_off;  // uint32 [experts_per_rank]   value `arrived` reaches when the current dispatch is complete (local)
    uint64_t sfa_off;       // int32 [kp][capacity]        MN-major packed UE8M0 scale factors
    uint64_t a_off;         // uint8 [capacity][k]         FP8 rows
    uint64_t total;
};

__host__ __device__ inline uint64_t align_up64(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

inline Layout make_layout(uint32_t world, uint32_t num_experts, uint32_t capacity, uint32_t k) {
    Layout l;
    uint64_t off = align_up64(sizeof(Control), 1024);
    l.table_off = off, off += align_up64(2 * 4ull * world * num_experts, 1024);
    l.hist_off = off, off += align_up64(4ull * kMaxRankCtas * num_experts, 1024);
    l.dst_base_off = off, off += align_up64(4ull * num_experts, 1024);
    l.psum_off = off, off += align_up64(4ull * num_experts, 1024);
    l.counts_off = off, off += align_up64(4ull * num_experts, 1024);
    l.sorted_off = off, off += align_up64(4ull * num_experts, 1024);
    l.arrived_off = off, off += align_up64(4ull * num_experts, 1024);
    l.expected_off = off, off += align_up64(4ull * num_experts, 1024);
    l.sfa_off = off, off += align_up64(4ull * ((k + 511) / 512) * capacity, 1024);
    l.a_off = off, off += align_up64(1ull * capacity * k, 1024);
    l.total = off;
    return l;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ep_globaltimer() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag reaches `epoch` (flags only grow); trap instead of hanging the GPU if a peer never shows up
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch) {
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (static_cast<int32_t>(ld_acquire_sys(flag) - epoch) < 0) {
        if ((++spins & 0x3FF) == 0) {
            const uint64_t now = ep_globaltimer();
            if (t0 == 0) t0 = now;
            if (now - t0 > kWaitTimeoutNs) {
                printf("dgb200 ep: timed out waiting for a peer flag (epoch %u)\n", epoch);
                asm volatile("trap;");
            }
        }
    }
}

template <typename id_t>
__device__ __forceinline__ int64_t load_id(const void* ids, uint32_t t) {
    return static_cast<int64_t>(__ldg(reinterpret_cast<const id_t*>(ids) + t));
}

// grid = num_experts, block = 1024. slot[t] = number of earlier local tokens with the same expert (stable order).
// Every thread looks at 4 consecutive tokens per pass (one 16 / 32-byte load), so a pass covers 4096 tokens with one
// warp scan + one 32-entry block scan.
template <typename id_t>
__global__ void __launch_bounds__(1024)
bucket_kernel(const void* __restrict__ ids, uint32_t num_tokens, int32_t* __restrict__ slot, int32_t* __restrict__ counts) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __shared__ uint32_t warp_off[33];
    const uint32_t e = blockIdx.x, tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const id_t* p = reinterpret_cast<const id_t*>(ids);
    const bool vec_ok = (reinterpret_cast<uintptr_t>(ids) & 31) == 0;
    uint32_t base = 0;
    for (uint32_t t0 = 0; t0 < num_tokens; t0 += 4096) {
        const uint32_t t = t0 + tid * 4;
        bool m[4];
        if (vec_ok && t + 4 <= num_tokens) {
            if constexpr (sizeof(id_t) == 4) {
                const int4 v = __ldg(reinterpret_cast<const int4*>(p + t));
                m[0] = v.x == static_cast<int>(e), m[1] = v.y == static_cast<int>(e), m[2] = v.z == static_cast<int>(e), m[3] = v.w == static_cast<int>(e);
            } else {
                const longlong2 v0 = __ldg(reinterpret_cast<const longlong2*>(p + t)), v1 = __ldg(reinterpret_cast<const longlong2*>(p + t + 2));
                m[0] = v0.x == static_cast<long long>(e), m[1] = v0.y == static_cast<long long>(e);
                m[2] = v1.x == static_cast<long long>(e), m[3] = v1.y == static_cast<long long>(e);
            }
        } else {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) m[j] = t + j < num_tokens && load_id<id_t>(ids, t + j) == static_cast<int64_t>(e);
        }
        const uint32_t mine = m[0] + m[1] + m[2] + m[3];
        uint32_t incl = mine;                                   // inclusive scan over the warp
#pragma unroll
        for (uint32_t d = 1; d < 32; d *= 2) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) warp_off[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t c = warp_off[lane];
            uint32_t w = c;
#pragma unroll
            for (uint32_t d = 1; d < 32; d *= 2) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += v;
            }
            warp_off[lane] = w - c;
            if (lane == 31) warp_off[32] = w;
        }
        __syncthreads();
        uint32_t pos = base + warp_off[warp] + incl - mine;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            if (m[j]) slot[t + j] = static_cast<int32_t>(pos++);
        base += warp_off[32];
        __syncthreads();
    }
    if (tid == 0) counts[e] = static_cast<int32_t>(base);
}

// grid = 1, block = 1024.
__global__ void __launch_bounds__(1024)
exchange_kernel(Peers peers, Layout l, uint32_t rank, uint32_t world, uint32_t num_experts, uint32_t capacity,
                uint32_t alignment, bool want_order) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __shared__ uint32_t s_total[kMaxExperts], s_before[kMaxExperts], s_aligned[kMaxExperts], s_sorted[kMaxExperts];
    const uint32_t tid = threadIdx.x;
    uint8_t* mine = peers.base[rank];
    Control* ctrl = reinterpret_cast<Control*>(mine);
    const uint32_t epoch = ctrl->epoch + 1;
    const int32_t* counts = reinterpret_cast<const int32_t*>(mine + l.counts_off);

    // publish my counts into row `rank` of every peer's table (the half selected by the epoch's parity)
    const uint32_t table_half = (epoch & 1) * world * num_experts;
    for (uint32_t i = tid; i < world * num_experts; i += blockDim.x) {
        const uint32_t p = i / num_experts, e = i - p * num_experts;
        reinterpret_cast<int32_t*>(peers.base[p] + l.table_off)[table_half + rank * num_experts + e] = counts[e];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) st_release_sys(&reinterpret_cast<Control*>(peers.base[tid])->counts_flag[rank * 8], epoch);

    // while the counts travel: expert-sorted order of my own tokens (needs only my counts)
    // send order: local expert index first, owner rank second -- every owner receives its expert 0, then its expert 1,
    // ... at the same pace (plain expert order would serve the owners one after the other)
    if (want_order) {
        const uint32_t epr_ = num_experts / world;
        // s_sorted[key] = count of the expert with that send key; then an exclusive prefix over keys
        for (uint32_t e = tid; e < num_experts; e += blockDim.x) s_sorted[(e % epr_) * world + e / epr_] = static_cast<uint32_t>(counts[e]);
        __syncthreads();
        for (uint32_t e = tid; e < num_experts; e += blockDim.x) {
            const uint32_t key = (e % epr_) * world + e / epr_;
            uint32_t before = 0;
            for (uint32_t j = 0; j < key; ++j) before += s_sorted[j];
            reinterpret_cast<int32_t*>(mine + l.sorted_off)[e] = static_cast<int32_t>(before);
            if (key == num_experts - 1) ctrl->num_routed = before + s_sorted[key];
        }
    }
    if (tid < world) wait_flag(&ctrl->counts_flag[tid * 8], epoch);
    __syncthreads();

    const int32_t* table = reinterpret_cast<const int32_t*>(mine + l.table_off) + table_half;
    for (uint32_t e = tid; e < num_experts; e += blockDim.x) {
        uint32_t total = 0, before = 0;
        for (uint32_t s = 0; s < world; ++s) {
            const uint32_t c = static_cast<uint32_t>(__ldcv(table + s * num_experts + e));
            total += c;
            if (s < rank) before += c;
        }
        s_total[e] = total, s_before[e] = before;
        s_aligned[e] = (total + alignment - 1) / alignment * alignment;
    }
    __syncthreads();
    const uint32_t epr = num_experts / world;
    int32_t* dst_base = reinterpret_cast<int32_t*>(mine + l.dst_base_off);
    int32_t* psum = reinterpret_cast<int32_t*>(mine + l.psum_off);
    for (uint32_t e = tid; e < num_experts; e += blockDim.x) {
        const uint32_t owner = e / epr;
        uint32_t seg = 0;
        for (uint32_t j = owner * epr; j < e; ++j) seg += s_aligned[j];
        dst_base[e] = static_cast<int32_t>(seg + s_before[e]);
        if (owner == rank) {
            // cumulative, like `arrived`; only dispatches that signal arrivals count (the two kinds may alternate)
            if (want_order) reinterpret_cast<uint32_t*>(mine + l.expected_off)[e - rank * epr] += s_total[e];
            psum[e - rank * epr] = static_cast<int32_t>(min(seg + s_total[e], capacity));   // stays in bounds on overflow
            if (e == (rank + 1) * epr - 1) {
                const uint32_t rows = seg + s_aligned[e];
                ctrl->num_rows = rows;
                if (rows > capacity) ctrl->overflow = 1;
            }
        }
    }
    __syncthreads();
    if (tid == 0) ctrl->epoch = epoch;
}

// order[p] = the p-th of my routed tokens in expert-sorted (stable) order: the order the scatter walks them in, so that
// experts complete one after the other on their owners. grid = ceil(T / 256).
template <typename id_t>
__global__ void __launch_bounds__(256)
order_kernel(const uint8_t* __restrict__ mine, Layout l, const void* __restrict__ ids, uint32_t num_tokens,
             uint32_t num_experts, int32_t* __restrict__ token_row, int32_t* __restrict__ order) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t >= num_tokens) return;
    const int32_t* sorted = reinterpret_cast<const int32_t*>(mine + l.sorted_off);
    const int64_t e = load_id<id_t>(ids, t);
    if (e < 0 || e >= static_cast<int64_t>(num_experts))
        token_row[t] = -1;                                                   // routed nowhere (DeepEP uses -1)
    else
        order[static_cast<uint32_t>(__ldg(sorted + e)) + static_cast<uint32_t>(token_row[t])] = static_cast<int32_t>(t);
}

// One warp per token, tokens walked in expert-sorted order. `x` rows of `k` bytes (pitch ldx), `sf` [T][kp] int32 words
// with strides (sf_stride_t, sf_stride_k). After a row has been written its owner's per-expert arrival counter is
// incremented (remote atomic): a consumer may start on an expert as soon as all of its rows are in
// (dgb200_ep_grouped_gemm), without waiting for the whole dispatch.
// kSignal = false: plain token order, no counters (the consumer waits for the whole dispatch: wait_kernel).
template <typename id_t, bool kSignal>
__global__ void __launch_bounds__(256, kSignal ? 8 : 4)   // kSignal: <= 32 registers, so that 4 CTAs per SM leave room for a co-resident GEMM CTA
scatter_kernel(Peers peers, Layout l, const uint8_t* __restrict__ x, int64_t ldx, const int32_t* __restrict__ sf,
               int64_t sf_stride_t, int64_t sf_stride_k, const void* __restrict__ ids, int32_t* __restrict__ token_row,
               const int32_t* __restrict__ order, uint32_t num_tokens_plain, uint32_t k, uint32_t kp, uint32_t rank,
               uint32_t world, uint32_t num_experts, uint32_t capacity) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // a flag-synchronised consumer may start now
    asm volatile("griddepcontrol.wait;" ::: "memory");
    uint8_t* mine = peers.base[rank];
    Control* ctrl = reinterpret_cast<Control*>(mine);
    const uint32_t epoch = ctrl->epoch;                       // the exchange of this dispatch already bumped it
    const uint32_t num_routed = kSignal ? ctrl->num_routed : num_tokens_plain;
    const int32_t* dst_base = reinterpret_cast<const int32_t*>(mine + l.dst_base_off);
    const uint32_t epr = num_experts / world;
    const uint32_t lane = threadIdx.x % 32;
    const uint32_t warps_per_cta = blockDim.x / 32;
    const uint32_t chunks = k / 16;

    for (uint32_t pos = blockIdx.x * warps_per_cta + threadIdx.x / 32; pos < num_routed; pos += gridDim.x * warps_per_cta) {
        const uint32_t t = kSignal ? static_cast<uint32_t>(__ldg(order + pos)) : pos;
        const int64_t e64 = load_id<id_t>(ids, t);
        if (!kSignal && (e64 < 0 || e64 >= static_cast<int64_t>(num_experts))) {     // routed nowhere (DeepEP uses -1)
            if (lane == 0) token_row[t] = -1;
            continue;
        }
        const uint32_t e = static_cast<uint32_t>(e64);
        const uint32_t owner = e / epr;
        const uint32_t row = static_cast<uint32_t>(__ldg(dst_base + e)) + static_cast<uint32_t>(token_row[t]);
        uint32_t* arrived = reinterpret_cast<uint32_t*>(peers.base[owner] + l.arrived_off) + (e - owner * epr);
        if (row >= capacity) {                                            // the exchange flagged `overflow`; drop, but count
            __syncwarp();
            if (lane == 0) {
                token_row[t] = -1;
                if (kSignal) asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(arrived) : "memory");
            }
            continue;
        }
        const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(t) * ldx);
        uint4* dst = reinterpret_cast<uint4*>(peers.base[owner] + l.a_off + static_cast<uint64_t>(row) * k);
        uint32_t c = lane;
        if constexpr (!kSignal) {
            for (; c + 192 < chunks; c += 224) {                          // 7 x 16 B in flight per lane (K = 7168: 2 rounds)
                uint4 v[7];
#pragma unroll
                for (uint32_t j = 0; j < 7; ++j) v[j] = __ldg(src + c + 32 * j);
#pragma unroll
                for (uint32_t j = 0; j < 7; ++j) dst[c + 32 * j] = v[j];
            }
        }
        for (; c + 96 < chunks; c += 128) {                               // 4 x 16 B in flight per lane
            const uint4 v0 = __ldg(src + c), v1 = __ldg(src + c + 32), v2 = __ldg(src + c + 64), v3 = __ldg(src + c + 96);
            dst[c] = v0, dst[c + 32] = v1, dst[c + 64] = v2, dst[c + 96] = v3;
        }
        for (; c < chunks; c += 32) dst[c] = __ldg(src + c);
        if (lane < kp) {
            int32_t* sfa = reinterpret_cast<int32_t*>(peers.base[owner] + l.sfa_off);
            sfa[static_cast<uint64_t>(lane) * capacity + row] = __ldg(sf + t * sf_stride_t + lane * sf_stride_k);
        }
        __syncwarp();                                                     // every lane's stores precede lane 0's release
        if (lane == 0) {
            token_row[t] = static_cast<int32_t>(row);
            if (kSignal) {
                asm volatile("fence.acq_rel.sys;" ::: "memory");
                asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(arrived) : "memory");
            }
        }
    }

    // completion: the last CTA to finish tells every peer that all of this rank's rows have landed
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(&ctrl->done_ctas, 1u);
        if (prev == gridDim.x - 1) {
            ctrl->done_ctas = 0;
            __threadfence_system();
            for (uint32_t p = 0; p < world; ++p)
                st_release_sys(&reinterpret_cast<Control*>(peers.base[p])->data_flag[rank * 8], epoch);
        }
    }
}

// grid = 1, block = 32: returns once every source rank's rows of the current epoch are visible here.
__global__ void __launch_bounds__(32)
wait_kernel(uint8_t* mine, uint32_t world) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    Control* ctrl = reinterpret_cast<Control*>(mine);
    const uint32_t epoch = ctrl->epoch;
    if (threadIdx.x < world) wait_flag(&ctrl->data_flag[threadIdx.x * 8], epoch);
    __syncwarp();
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------- fused dispatch
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Barrier over all CTAs of a fully resident grid: arrivals are counted in `*bar`, which the launch leaves at zero again.
__device__ __forceinline__ void grid_barrier(uint32_t* bar, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        uint64_t t0 = 0;
        uint32_t spins = 0;
        while (ld_acquire_gpu(bar) < target) {
            if ((++spins & 0x3FF) == 0) {
                const uint64_t now = ep_globaltimer();
                if (t0 == 0) t0 = now;
                if (now - t0 > kWaitTimeoutNs) {
                    printf("dgb200 ep: grid barrier timed out (is the dispatch grid fully resident?)\n");
                    asm volatile("trap;");
                }
            }
        }
    }
    __syncthreads();
}

// The whole dispatch in one launch (see the header comment). grid <= kMaxRankCtas CTAs, all resident (the host sizes the
// grid from the occupancy API); block = kFusedThreads. Entries = (token, slot) pairs, entry i reads token i / topk.
//   slice_len   : entries per ranking CTA (multiple of the block size); the first ceil(num_entries / slice_len) CTAs rank
//   token_row[i]: row of entry i in its owner's buffer (-1: not routed / dropped); used as scratch for the ranks in between
template <typename id_t>
__global__ void __launch_bounds__(kFusedThreads, 2)
dispatch_fused_kernel(Peers peers, Layout l, const uint8_t* __restrict__ x, int64_t ldx, const int32_t* __restrict__ sf,
                      int64_t sf_stride_t, int64_t sf_stride_k, const void* __restrict__ ids, int32_t* __restrict__ token_row,
                      uint32_t num_entries, uint32_t topk, uint32_t slice_len, uint32_t k, uint32_t kp, uint32_t rank,
                      uint32_t world, uint32_t num_experts, uint32_t capacity, uint32_t alignment) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __shared__ uint32_t s_cnt[kMaxExperts];        // rank phase: per-expert counters | exchange: totals
    __shared__ uint32_t s_aux[kMaxExperts];        // rank phase: slice prefix        | exchange: rows of lower source ranks
    __shared__ uint32_t s_aligned[kMaxExperts];
    __shared__ int32_t s_ids[kFusedThreads];
    const uint32_t tid = threadIdx.x, cta = blockIdx.x, grid = gridDim.x;
    uint8_t* mine = peers.base[rank];
    Control* ctrl = reinterpret_cast<Control*>(mine);
    const uint32_t epoch = ctrl->epoch + 1;        // (written by CTA 0 only after the last grid barrier)
    int32_t* hist = reinterpret_cast<int32_t*>(mine + l.hist_off);
    const uint32_t num_slices = (num_entries + slice_len - 1) / slice_len;
    const bool ranker = cta < num_slices;
    const uint32_t slice_begin = cta * slice_len, slice_end = min(num_entries, slice_begin + slice_len);
    const bool stamper = cta == 0 && tid == 0;
    if (stamper) ctrl->dbg_ns[0] = ep_globaltimer();

    // ---- phase 1: stable rank of every entry inside (its expert, this slice) + the slice's histogram
    if (ranker) {
        for (uint32_t e = tid; e < num_experts; e += blockDim.x) s_cnt[e] = 0;
        __syncthreads();
        for (uint32_t base = slice_begin; base < slice_end; base += blockDim.x) {
            const uint32_t i = base + tid;
            int32_t e = -1;
            if (i < slice_end) {
                const int64_t e64 = load_id<id_t>(ids, i);
                if (e64 >= 0 && e64 < static_cast<int64_t>(num_experts)) e = static_cast<int32_t>(e64);
            }
            s_ids[tid] = e;
            __syncthreads();
            if (e >= 0) {
                uint32_t r = s_cnt[e];
                for (uint32_t j = 0; j < tid; ++j) r += s_ids[j] == e;     // shared-memory broadcast reads
                token_row[i] = static_cast<int32_t>(r);
            } else if (i < slice_end) {
                token_row[i] = -1;                                          // routed nowhere (DeepEP uses -1)
            }
            __syncthreads();
            if (e >= 0) atomicAdd(&s_cnt[e], 1u);
            __syncthreads();
        }
        for (uint32_t e = tid; e < num_experts; e += blockDim.x) hist[cta * num_experts + e] = static_cast<int32_t>(s_cnt[e]);
    }
    grid_barrier(&ctrl->grid_bar, grid);
    if (stamper) ctrl->dbg_ns[1] = ep_globaltimer();

    // ---- phase 2a (CTA 0): my counts -> every peer; theirs -> destination rows of my tokens + the local psum layout
    if (cta == 0) {
        int32_t* counts = reinterpret_cast<int32_t*>(mine + l.counts_off);
        const uint32_t table_half = (epoch & 1) * world * num_experts;
        for (uint32_t e = tid; e < num_experts; e += blockDim.x) {
            uint32_t c = 0;
            for (uint32_t sl = 0; sl < num_slices; ++sl) c += static_cast<uint32_t>(__ldcg(hist + sl * num_experts + e));
            counts[e] = static_cast<int32_t>(c);
            for (uint32_t p = 0; p < world; ++p)
                reinterpret_cast<int32_t*>(peers.base[p] + l.table_off)[table_half + rank * num_experts + e] = static_cast<int32_t>(c);
        }
        __threadfence_system();
        __syncthreads();
        if (tid < world) st_release_sys(&reinterpret_cast<Control*>(peers.base[tid])->counts_flag[rank * 8], epoch);
        if (stamper) ctrl->dbg_ns[2] = ep_globaltimer();
        if (tid < world) wait_flag(&ctrl->counts_flag[tid * 8], epoch);
        __syncthreads();
        if (stamper) ctrl->dbg_ns[3] = ep_globaltimer();
        const int32_t* table = reinterpret_cast<const int32_t*>(mine + l.table_off) + table_half;
        for (uint32_t e = tid; e < num_experts; e += blockDim.x) {
            uint32_t total = 0, before = 0;
            for (uint32_t sr = 0; sr < world; ++sr) {
                const uint32_t c = static_cast<uint32_t>(__ldcv(table + sr * num_experts + e));
                total += c;
                if (sr < rank) before += c;
            }
            s_cnt[e] = total, s_aux[e] = before;
            s_aligned[e] = (total + alignment - 1) / alignment * alignment;
        }
        __syncthreads();
        const uint32_t epr = num_experts / world;
        int32_t* dst_base = reinterpret_cast<int32_t*>(mine + l.dst_base_off);
        int32_t* psum = reinterpret_cast<int32_t*>(mine + l.psum_off);
        for (uint32_t e = tid; e < num_experts; e += blockDim.x) {
            const uint32_t owner = e / epr;
            uint32_t seg = 0;
            for (uint32_t j = owner * epr; j < e; ++j) seg += s_aligned[j];
            dst_base[e] = static_cast<int32_t>(seg + s_aux[e]);
            if (owner == rank) {
                psum[e - rank * epr] = static_cast<int32_t>(min(seg + s_cnt[e], capacity));   // stays in bounds on overflow
                if (e == (rank + 1) * epr - 1) {
                    const uint32_t rows = seg + s_aligned[e];
                    ctrl->num_rows = rows;
                    if (rows > capacity) ctrl->overflow = 1;
                }
            }
        }
    }
    // ---- phase 2b (ranking CTAs, CTA 0 after its exchange): slice-local rank -> rank among ALL my entries of the expert
    if (ranker) {
        __syncthreads();
        for (uint32_t e = tid; e < num_experts; e += blockDim.x) {
            uint32_t before = 0;
            for (uint32_t sl = 0; sl < cta; ++sl) before += static_cast<uint32_t>(__ldcg(hist + sl * num_experts + e));
            s_aux[e] = before;
        }
        __syncthreads();
        for (uint32_t i = slice_begin + tid; i < slice_end; i += blockDim.x) {
            const int32_t r = token_row[i];
            if (r >= 0) token_row[i] = r + static_cast<int32_t>(s_aux[static_cast<uint32_t>(load_id<id_t>(ids, i))]);
        }
    }
    grid_barrier(&ctrl->grid_bar, 2 * grid);
    if (stamper) ctrl->dbg_ns[4] = ep_globaltimer();

    // ---- phase 3: scatter, one warp per entry (grid-strided so that consecutive warps write consecutive rows' worth of bytes)
    {
        const int32_t* dst_base = reinterpret_cast<const int32_t*>(mine + l.dst_base_off);
        const uint32_t epr = num_experts / world;
        const uint32_t lane = tid % 32, warps_per_cta = blockDim.x / 32;
        const uint32_t chunks = k / 16;
        for (uint32_t i = cta * warps_per_cta + tid / 32; i < num_entries; i += grid * warps_per_cta) {
            const int32_t slot = __ldcg(token_row + i);
            if (slot < 0) continue;
            const uint32_t e = static_cast<uint32_t>(load_id<id_t>(ids, i));
            const uint32_t owner = e / epr;
            const uint32_t row = static_cast<uint32_t>(__ldcg(dst_base + e)) + static_cast<uint32_t>(slot);
            __syncwarp();
            if (row >= capacity) {                                            // the exchange flagged `overflow`: drop
                if (lane == 0) token_row[i] = -1;
                continue;
            }
            const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(i / topk) * ldx);
            uint4* dst = reinterpret_cast<uint4*>(peers.base[owner] + l.a_off + static_cast<uint64_t>(row) * k);
            uint32_t c = lane;
            for (; c + 192 < chunks; c += 224) {                              // 7 x 16 B in flight per lane (K = 7168: 2 rounds)
                uint4 v[7];
#pragma unroll
                for (uint32_t j = 0; j < 7; ++j) v[j] = __ldg(src + c + 32 * j);
#pragma unroll
                for (uint32_t j = 0; j < 7; ++j) dst[c + 32 * j] = v[j];
            }
            for (; c + 96 < chunks; c += 128) {
                const uint4 v0 = __ldg(src + c), v1 = __ldg(src + c + 32), v2 = __ldg(src + c + 64), v3 = __ldg(src + c + 96);
                dst[c] = v0, dst[c + 32] = v1, dst[c + 64] = v2, dst[c + 96] = v3;
            }
            for (; c < chunks; c += 32) dst[c] = __ldg(src + c);
            if (lane < kp) {
                int32_t* sfa = reinterpret_cast<int32_t*>(peers.base[owner] + l.sfa_off);
                sfa[static_cast<uint64_t>(lane) * capacity + row] = __ldg(sf + static_cast<int64_t>(i / topk) * sf_stride_t + lane * sf_stride_k);
            }
            if (lane == 0) token_row[i] = static_cast<int32_t>(row);
        }
    }

    // ---- phase 4: the last CTA tells every peer that all of this rank's rows have landed and waits for theirs
    if (stamper) ctrl->dbg_ns[5] = ep_globaltimer();
    __threadfence_system();
    __syncthreads();
    __shared__ uint32_t s_last;
    if (tid == 0) s_last = atomicAdd(&ctrl->done_ctas, 1u) == grid - 1;
    __syncthreads();
    if (s_last) {
        if (tid == 0) {
            ctrl->done_ctas = 0;
            ctrl->grid_bar = 0;                     // every CTA has passed both barriers: leave the counter clean
            ctrl->epoch = epoch;
            __threadfence_system();
        }
        __syncthreads();
        if (tid == 0) ctrl->dbg_ns[6] = ep_globaltimer();
        if (tid < world) st_release_sys(&reinterpret_cast<Control*>(peers.base[tid])->data_flag[rank * 8], epoch);
        if (tid < world) wait_flag(&ctrl->data_flag[tid * 8], epoch);
        __syncthreads();
        if (tid == 0) ctrl->dbg_ns[7] = ep_globaltimer();
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------------- combine (weighted top-k)
// The way back: every token's output row sits in its expert owner's D buffer [capacity, n] (peer mapped); the source
// rank pulls it over NVLink into token order. `publish` runs after the owner's grouped GEMM in stream order.
__global__ void __launch_bounds__(32)
combine_publish_kernel(Peers ctrl_bufs, uint32_t rank, uint32_t world) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t epoch = reinterpret_cast<const Control*>(ctrl_bufs.base[rank])->epoch;
    __threadfence_system();                      // (kernel boundary already ordered the GEMM's stores; belt and braces)
    if (threadIdx.x < world)
        st_release_sys(&reinterpret_cast<Control*>(ctrl_bufs.base[threadIdx.x])->gemm_flag[rank * 8], epoch);
}

// One warp per local token: out[t, :] = sum_j w[t, j] * D_owner(t, j)[token_row[t * topk + j], :], the products and the
// running sum in FP32 in slot order (separate multiply and add, so a plain torch loop reproduces the bits), rounded once
// to BF16; slots that were not routed / dropped contribute nothing, a token without any routed slot gets zeros.
// `weights` == nullptr means 1.0 (top-1 routing: a pure gather, bit-exact copy of the owner's row).
// The reverse path of the reference's baseline MoE step (tests/test_mega_moe.py:196-202: dispatch -> GEMMs -> combine).
template <typename id_t>
__global__ void __launch_bounds__(256)
combine_gather_kernel(Peers ctrl_bufs, Peers d_bufs, const void* __restrict__ ids, const int32_t* __restrict__ token_row,
                      const float* __restrict__ weights, uint32_t topk, uint8_t* __restrict__ out, int64_t ldo_bytes,
                      int64_t ldd_bytes, uint32_t row_bytes, uint32_t num_tokens, uint32_t num_experts, uint32_t rank,
                      uint32_t world) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const Control* ctrl = reinterpret_cast<const Control*>(ctrl_bufs.base[rank]);
    const uint32_t epoch = ctrl->epoch;
    uint64_t* dbg = const_cast<uint64_t*>(ctrl->dbg_ns);
    const bool stamper = blockIdx.x == 0 && threadIdx.x == 0;
    if (stamper) dbg[8] = ep_globaltimer();
    if (threadIdx.x < world) wait_flag(&ctrl->gemm_flag[threadIdx.x * 8], epoch);   // every owner's GEMM of this step is done
    __syncthreads();
    if (stamper) dbg[9] = ep_globaltimer();
    const uint32_t epr = num_experts / world;
    const uint32_t lane = threadIdx.x % 32, warps_per_cta = blockDim.x / 32;
    const uint32_t chunks = row_bytes / 16;
    for (uint32_t t = blockIdx.x * warps_per_cta + threadIdx.x / 32; t < num_tokens; t += gridDim.x * warps_per_cta) {
        uint4* dst = reinterpret_cast<uint4*>(out + static_cast<int64_t>(t) * ldo_bytes);
        if (topk == 1 && weights == nullptr) {                       // pure gather
            const int32_t row = __ldg(token_row + t);
            if (row < 0) {
                for (uint32_t c = lane; c < chunks; c += 32) dst[c] = make_uint4(0u, 0u, 0u, 0u);
                continue;
            }
            const uint32_t owner = static_cast<uint32_t>(load_id<id_t>(ids, t)) / epr;
            const uint4* src = reinterpret_cast<const uint4*>(d_bufs.base[owner] + static_cast<int64_t>(row) * ldd_bytes);
            // remote reads are a round trip over NVLink (~2-3 us): 8 x 16 B in flight per lane, an 8 KB row in two trips
            uint32_t c = lane;
            for (; c + 224 < chunks; c += 256) {
                uint4 v[8];
#pragma unroll
                for (uint32_t j = 0; j < 8; ++j) v[j] = __ldcv(src + c + 32 * j);
#pragma unroll
                for (uint32_t j = 0; j < 8; ++j) dst[c + 32 * j] = v[j];
            }
            for (; c + 96 < chunks; c += 128) {
                const uint4 v0 = __ldcv(src + c), v1 = __ldcv(src + c + 32), v2 = __ldcv(src + c + 64), v3 = __ldcv(src + c + 96);
                dst[c] = v0, dst[c + 32] = v1, dst[c + 64] = v2, dst[c + 96] = v3;
            }
            for (; c < chunks; c += 32) dst[c] = __ldcv(src + c);
            continue;
        }
        for (uint32_t c = lane; c < chunks; c += 32) {               // 8 BF16 outputs per lane and pass
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (uint32_t j0 = 0; j0 < topk; j0 += 8) {              // up to 8 slots at a time: their remote reads travel together
                uint4 v[8];
                float w[8];
#pragma unroll
                for (uint32_t jj = 0; jj < 8; ++jj) {
                    const uint32_t j = j0 + jj;
                    const int32_t row = j < topk ? __ldg(token_row + t * topk + j) : -1;
                    w[jj] = 0.f, v[jj] = make_uint4(0u, 0u, 0u, 0u);
                    if (row < 0) continue;
                    const uint32_t owner = static_cast<uint32_t>(load_id<id_t>(ids, t * topk + j)) / epr;
                    w[jj] = weights != nullptr ? __ldg(weights + t * topk + j) : 1.0f;
                    v[jj] = __ldcv(reinterpret_cast<const uint4*>(d_bufs.base[owner] + static_cast<int64_t>(row) * ldd_bytes) + c);
                }
#pragma unroll
                for (uint32_t jj = 0; jj < 8; ++jj) {                // slot order, as before: a slot that was not routed adds nothing
                    if (j0 + jj >= topk || __ldg(token_row + t * topk + j0 + jj) < 0) continue;
                    const uint32_t u[4] = {v[jj].x, v[jj].y, v[jj].z, v[jj].w};
#pragma unroll
                    for (uint32_t q = 0; q < 4; ++q) {
                        acc[2 * q] = __fadd_rn(acc[2 * q], __fmul_rn(w[jj], __uint_as_float(u[q] << 16)));
                        acc[2 * q + 1] = __fadd_rn(acc[2 * q + 1], __fmul_rn(w[jj], __uint_as_float(u[q] & 0xFFFF0000u)));
                    }
                }
            }
            uint32_t o[4];
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) {
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
                o[q] = *reinterpret_cast<const uint32_t*>(&b2);
            }
            dst[c] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    if (stamper) dbg[10] = ep_globaltimer();
}

}  // namespace ep
}  // namespace dgb200
