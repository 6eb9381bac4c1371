// Raw sm_100a PTX wrappers used by the FP8 blockwise GEMM kernel.
// No CUTLASS / CuTe: every instruction the kernel relies on is spelled out here.
//
// Reference cross-check (what each wrapper replaces in deepseek-ai/DeepGEMM):
//   tcgen05.mma ... block_scale      <- deep_gemm/include/deep_gemm/ptx/tcgen05.cuh:40-80
//   descriptor bitfields             <- third-party/cutlass/include/cute/arch/mma_sm100_desc.hpp:98-123,441-463
//   tcgen05.cp / ld / alloc / commit <- cute/arch/copy_sm100.hpp:482-517, tmem_allocator_sm100.hpp, cutlass/arch/barrier.h:766-820
//   TMA loads + mbarrier             <- cute/arch/copy_sm90_tma.hpp, cutlass/arch/barrier.h:395-500
#pragma once
#include <cstdint>
#include <cuda.h>

namespace dgb200 {
namespace ptx {

#define DGB_DEVICE __device__ __forceinline__

DGB_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

DGB_DEVICE uint32_t lane_id() {
    uint32_t v;
    asm("mov.u32 %0, %%laneid;" : "=r"(v));
    return v;
}

DGB_DEVICE uint32_t cluster_ctarank() {
    uint32_t v;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(v));
    return v;
}

DGB_DEVICE uint32_t cluster_id_x() {
    uint32_t v;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(v));
    return v;
}

DGB_DEVICE uint32_t num_clusters_x() {
    uint32_t v;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(v));
    return v;
}

DGB_DEVICE bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- cluster sync
DGB_DEVICE void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
DGB_DEVICE void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
DGB_DEVICE void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// per-thread forms (no .aligned): for a lane that joins the cluster barrier on its own schedule
DGB_DEVICE void cluster_arrive_relaxed_thread() { asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory"); }
DGB_DEVICE void cluster_wait_thread() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier (all addresses are 32-bit shared::cta)
DGB_DEVICE void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
DGB_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

DGB_DEVICE void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
DGB_DEVICE void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) as seen in CTA `cta` of the cluster
DGB_DEVICE uint32_t mapa(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
// Arrive on a barrier of another CTA of the cluster (`bar` is a shared::cluster address from mapa()).
// NOTE: default (.release.cta) semantics on purpose. Explicit `.release.cluster` / `.acquire.cluster` qualifiers make
// ptxas emit MEMBAR.ALL.GPU + CCTL.IVALL around every barrier operation, which serialised the whole k-loop
// (measured: 2.3x slower kernel). mbarrier objects in shared::cluster are coherent across the CTA pair.
DGB_DEVICE void mbar_arrive_remote(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
DGB_DEVICE bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n"   // hardware-suspended wait, bounded
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// A protocol bug must not hang the GPU: a wait that lasts longer than kSpinTimeoutNs traps (sticky CUDA error),
// the same policy and limit (60 s) as the reference's in-kernel barriers (comm/barrier.cuh:11-12).
constexpr uint64_t kSpinTimeoutNs = 60ull * 1000 * 1000 * 1000;
DGB_DEVICE uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
DGB_DEVICE void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xFF) == 0) {
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kSpinTimeoutNs) asm volatile("trap;");
        }
    }
}

// ---------------------------------------------------------------- proxies / fences
DGB_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DGB_DEVICE void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DGB_DEVICE void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

DGB_DEVICE void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// 2-D tiled load: global (via tensor map) -> this CTA's shared memory, completion counted on `bar`.
DGB_DEVICE void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t smem_dst, uint32_t c0, uint32_t c1,
                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}

// Ask L2 for one box of a tiled tensor (no shared-memory destination, no completion: a hint)
DGB_DEVICE void tma_prefetch_2d(const CUtensorMap* map, uint32_t c0, uint32_t c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0),
                 "r"(c1)
                 : "memory");
}

// 3-D tiled load (batched GEMM: coordinate 2 = batch, box depth 1)
DGB_DEVICE void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t smem_dst, uint32_t c0, uint32_t c1, uint32_t c2,
                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
        : "memory");
}

// Same, delivered to the same shared-memory offset of every CTA in `cta_mask` (and counted on each one's barrier).
DGB_DEVICE void tma_load_2d_multicast(const CUtensorMap* map, uint32_t bar, uint32_t smem_dst, uint32_t c0, uint32_t c1,
                                      uint16_t cta_mask, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5, %6;" ::"r"(smem_dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask), "l"(hint)
        : "memory");
}

// 2-D tiled store: this CTA's shared memory -> global (via tensor map), tracked by the thread's bulk async-group.
DGB_DEVICE void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, uint32_t c0, uint32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
DGB_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most kPending of this thread's bulk groups may still be READING their shared-memory source
template <int kPending>
DGB_DEVICE void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
DGB_DEVICE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// four 8x8 b16 matrices, transposed on the way: matrix i's stored row r (16 bytes at the address given by thread 8i + r)
// receives element r of every fragment row, i.e. the fragment's column r
DGB_DEVICE void stmatrix_x4_trans(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("stmatrix.sync.aligned.x4.m8n8.shared.b16.trans [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
                 "r"(d)
                 : "memory");
}
// {lo, hi} FP32 -> packed BF16 pair (round to nearest even), lo in bits [0, 16)
DGB_DEVICE uint32_t pack_bf16x2(uint32_t lo, uint32_t hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
    return r;
}

// ---------------------------------------------------------------- tensor memory
template <int kCtaGroup>
DGB_DEVICE void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    if constexpr (kCtaGroup == 1)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                     "r"(ncols)
                     : "memory");
    else
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                     "r"(ncols)
                     : "memory");
}
template <int kCtaGroup>
DGB_DEVICE void tmem_relinquish() {
    if constexpr (kCtaGroup == 1)
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    else
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCtaGroup>
DGB_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (kCtaGroup == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// smem -> TMEM copy of one 128-row scale-factor group (32 lanes x 128 bit, replicated to the 4 lane quadrants)
template <int kCtaGroup>
DGB_DEVICE void tmem_cp_sf(uint32_t taddr, uint64_t sdesc) {
    if constexpr (kCtaGroup == 1)
        asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
    else
        asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T with UE8M0 block scales held in TMEM (one scale per 32 K-elements, selected by sf ids)
template <int kCtaGroup>
DGB_DEVICE void mma_mxf8_block_scale(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t tmem_sfa,
                                     uint32_t tmem_sfb, uint32_t accumulate) {
    if constexpr (kCtaGroup == 1)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
            : "memory");
    else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
            : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T for 16-bit operands (kind::f16; BF16 here), 16 K-elements per instruction, FP32 accumulate
template <int kCtaGroup>
DGB_DEVICE void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (kCtaGroup == 1)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}

// Make all previously issued tcgen05 ops of this thread arrive on an mbarrier when they retire.
// cta_group::2 form multicasts the arrival to the barrier at the same offset in both CTAs of the pair.
template <int kCtaGroup>
DGB_DEVICE void mma_commit(uint32_t bar, uint16_t mask = 0b11) {
    if constexpr (kCtaGroup == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                     : "memory");
    } else {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                bar),
            "h"(mask)
            : "memory");
    }
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns.
DGB_DEVICE void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
DGB_DEVICE void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// TMEM -> registers, 16 lanes x 8 columns: thread t gets lane t/4 (regs 0,1) and t/4 + 8 (regs 2,3), columns 2(t%4), +1
// -- the accumulator-fragment layout stmatrix consumes. Lanes 16..31 of the warp's quadrant: add 16 << 16 to `taddr`.
DGB_DEVICE void tmem_ld_16x256b(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
DGB_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// 64-bit shared-memory matrix descriptor (version 1 = sm_100):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version | [61,64) swizzle (0 none, 2 = 128B)
DGB_DEVICE uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout_type & 0x7) << 61;
    return d;
}
constexpr uint32_t kLayoutNoSwizzle = 0;
constexpr uint32_t kLayoutSwizzle128B = 2;
constexpr uint32_t kLayoutSwizzle64B = 4;
constexpr uint32_t kLayoutSwizzle32B = 6;

// 32-bit instruction descriptor for kind::mxf8f6f4.block_scale, E4M3 x E4M3, UE8M0 scales, FP32 accumulate:
//   [4,6) b_sf_id | [7,10) a_fmt | [10,13) b_fmt | 15 a_major | 16 b_major | [17,23) N>>3 | 23 scale=UE8M0
//   [24,29) M>>4 | [29,31) a_sf_id
DGB_DEVICE constexpr uint32_t make_idesc(uint32_t umma_m, uint32_t umma_n, uint32_t a_mn_major, uint32_t b_mn_major) {
    return (a_mn_major << 15) | (b_mn_major << 16) | ((umma_n >> 3) << 17) | (1u << 23) | ((umma_m >> 4) << 24);
}
// 32-bit instruction descriptor for kind::f16 with BF16 x BF16 -> FP32 (bits 15 / 16: operand is MN-major)
// (cute/arch/mma_sm100_desc.hpp:411-432: [4,6) c_format 1 = F32 | [7,10) a_format 1 = BF16 | [10,13) b_format | [17,23) N>>3 | [24,29) M>>4)
DGB_DEVICE constexpr uint32_t make_idesc_bf16(uint32_t umma_m, uint32_t umma_n, uint32_t a_mn_major = 0, uint32_t b_mn_major = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((umma_n >> 3) << 17) | ((umma_m >> 4) << 24);
}
DGB_DEVICE uint32_t idesc_with_sf_ids(uint32_t idesc, uint32_t a_sf_id, uint32_t b_sf_id) {
    return idesc | (b_sf_id << 4) | (a_sf_id << 29);
}

// ---------------------------------------------------------------- misc
DGB_DEVICE void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
DGB_DEVICE uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// 16-byte store into a peer CTA's shared memory that completes 16 transaction bytes on the peer's mbarrier
// (`dst` and `bar` are shared::cluster addresses from mapa)
DGB_DEVICE void st_async_v4(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar)
                 : "memory");
}
// bulk copy of `bytes` (multiple of 16) from this CTA's shared memory into a peer's, completing transaction bytes on
// the peer's mbarrier (`dst`, `bar`: shared::cluster addresses from mapa; `src`: shared::cta)
DGB_DEVICE void bulk_copy_to_peer(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "r"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
DGB_DEVICE void st_shared_f4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
DGB_DEVICE uint4 ld_shared_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
DGB_DEVICE float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

DGB_DEVICE void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace ptx
}  // namespace dgb200
