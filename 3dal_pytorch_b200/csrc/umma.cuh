// sm_100a primitives used by the tensor-core kernels: mbarrier, TMA bulk copy (cp.async.bulk),
// TMEM allocation, tcgen05.mma / commit / ld, and the shared-memory operand layout.
//
// Operand layout ("KP", K-major, no swizzle): a tile of R rows x K columns of bf16 is stored as
// K/8 planes of R x 16 bytes:   byte(r, k) = (k / 8) * (R * 16) + r * 16 + (k % 8) * 2.
// One plane row is one 16-byte core-matrix row, 8 consecutive rows are one 8x16B core matrix, so a
// UMMA shared-memory descriptor (SWIZZLE_NONE, K-major) describes it with
//     leading-dimension byte offset (between the two K core matrices of one K=16 MMA) = R * 16
//     stride-dimension  byte offset (between 8-row groups along M/N)                   = 128
// Packing a (R, K) row-major matrix into this layout is `w.view(R, K/8, 8).permute(1, 0, 2)`.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace al3d {
namespace umma {

// Watchdog.  A wait that exceeds kWaitTimeoutCycles clock64 ticks records `code` in the status word of the launch
// (pinned, host-mapped memory owned by the library, one word per device: the host reads it without any CUDA call) and,
// unless the launch asked for diagnostics only, traps -- the launch then fails with a CUDA error at the caller's next
// synchronisation instead of handing back invalid outputs.  There is no device-global state: the word travels in the
// kernel parameters.
struct TcStatus {
    unsigned int *word;      // host-mapped status word (0 = ok)
    int trap;                // bit 0: __trap() after recording the code (default 1); bit 1: test hook (seg_pass1_kernel)
    int stress;              // > 0: upper bound (ns) of the pseudo-random delays injected per role (protocol stress tests)
};
constexpr long long kWaitTimeoutCycles = 1ll << 31;

static __device__ __noinline__ void watchdog_fire(const TcStatus &st, uint32_t code)
{
    if (st.word) {
        atomicCAS_system(st.word, 0u, code);
        __threadfence_system();
    }
    if (st.trap & 1) __trap();
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp, chosen by elect.sync.  Single-thread regions that issue tcgen05 / bulk-copy
// instructions should be entered through this and not through `lane == 0`: ptxas knows an elect.sync region has exactly
// one active lane and issues the uniform-datapath instructions (UTCHMMA, UTCBAR ...) directly, whereas in a region it
// cannot prove single-lane it wraps EVERY such instruction in an ELECT / branch loop over the active lanes.
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}

// Delay injection for the protocol stress tests (a no-op when st.stress == 0; only compiled into libal3d_stress.so).
// st.stress < 2^24: with probability 1/4 sleep a pseudo-random time below st.stress ns.  Bit 24 set: TARGETED mode --
// only the hooks of seg_pass1_kernel's front warps (salts 0x81, 0x82, 0x84) sleep, always, for the
// full (st.stress & 0xFFFFFF) ns: a consistently slow front is what laps the round-1 release protocol.
__device__ __forceinline__ uint32_t stress_ns_for(const TcStatus &st, uint32_t salt)
{
    if (st.stress & (1 << 24)) return (salt == 0x81u || salt == 0x82u || salt == 0x84u) ? ((uint32_t)st.stress & 0xFFFFFFu) : 0u;
    uint32_t h = ((uint32_t)clock64() ^ (blockIdx.x * 0x9E3779B9u) ^ ((salt + (threadIdx.x >> 5) * 131u) * 0x85EBCA6Bu)) * 0xC2B2AE35u;
    h ^= h >> 15;
    return ((h & 3u) == 0u) ? (h >> 8) % (uint32_t)st.stress : 0u;
}
__device__ __forceinline__ void stress_delay(const TcStatus &st, uint32_t salt)        // single thread
{
    if (st.stress > 0) {
        const uint32_t ns = stress_ns_for(st, salt);
        if (ns) __nanosleep(ns);
    }
}
__device__ __forceinline__ void stress_delay_warp(const TcStatus &st, uint32_t salt)   // whole converged warp, same delay
{
    if (st.stress > 0) {
        const uint32_t ns = __shfl_sync(0xFFFFFFFFu, stress_ns_for(st, salt), 0);
        if (ns) __nanosleep(ns);
        __syncwarp();
    }
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Returns false if the wait was abandoned (timeout; only reachable with st.trap == 0).
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, uint32_t code, const TcStatus &st)
{
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023) == 0 && clock64() - t0 > kWaitTimeoutCycles) {
            watchdog_fire(st, code);
            return false;
        }
    }
    return true;
}

// Same wait with cluster-scope acquire: for barriers whose arrivals come from the peer CTA of a pair.
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t *bar, uint32_t parity, uint32_t code, const TcStatus &st)
{
    const long long t0 = clock64();
    int spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
        if ((++spins & 1023) == 0 && clock64() - t0 > kWaitTimeoutCycles) {
            watchdog_fire(st, code);
            return false;
        }
    }
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMA bulk copy global -> shared
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- clusters (CTA pairs)
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    // default semantics (release at CTA scope), as for a local arrive: the data handed over lives in TMEM and is ordered by
    // tcgen05.wait::st + tcgen05.fence::before_thread_sync, so no cluster-scope release of generic memory is needed -- the
    // .release.cluster form stalled the arriving thread for ~250 cycles per hand-over (profiles/r1_pass2_peer_timeline.txt)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// same with a cluster-scope release: for hand-overs of generic-proxy data (shared / global memory) to the other CTA
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t *bar, uint32_t rank)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// ---------------------------------------------------------------- TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem)     // whole warp, in BOTH CTAs of the pair
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem)     // whole warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)       // whole warp
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (warp%4)*32+i.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns, registers -> TMEM (thread i of the warp writes lane (warp%4)*32+i).
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors + MMA
// Shared-memory matrix descriptor for a KP-layout tile of R rows, starting at byte address `saddr`
// (the first of the two K planes this K=16 MMA reads).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t rows, bool swap_lbo_sbo = false)
{
    uint32_t lbo = rows * 16u, sbo = 128u;
    if (swap_lbo_sbo) { uint32_t t = lbo; lbo = sbo; sbo = t; }
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                     // descriptor version 1 (sm_100)
    return d;                                   // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

// Descriptor of a sub-tile of rows inside a KP tile whose planes are `plane_bytes` apart (e.g. 64 rows of a 128-row tile).
__device__ __forceinline__ uint64_t make_desc_lbo(uint32_t saddr, uint32_t plane_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((plane_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((128u >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N)
{
    return (1u << 4)                 // D format fp32
         | (1u << 7) | (1u << 10)    // A, B format bf16
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// The same instruction with IEEE fp16 operands (A, B format 0): 11 significant bits instead of 8, range +-65504.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: A is a bf16 tile held in TMEM (row = lane, two consecutive K elements
// per 32-bit column, low half = even k); a K=16 slice is 8 columns starting at tmem_a.
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// CTA-pair MMA (cta_group::2, M = 256): issued by the leader CTA only.  Each CTA supplies its own 128 rows of A
// (from its TMEM, same column address) and N/2 rows of B (from its shared memory, same offset); each CTA's TMEM
// receives its 128 rows of D.
__device__ __forceinline__ void mma_bf16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// CTA-pair MMA with both operands in shared memory: each CTA supplies 128 rows of A and N/2 rows of B (same offsets).
__device__ __forceinline__ void mma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit for CTA-pair MMAs: arrives on the mbarrier at the same offset in every CTA of cta_mask
__device__ __forceinline__ void mma_commit_pair(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// mbarrier arrive once every tcgen05 op issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One K=64 block: 4 MMAs of K=16.  a_addr / b_addr: smem byte addresses of the first plane of the
// 64-wide K slice; rows_a / rows_b: row counts of the two KP tiles.
__device__ __forceinline__ void mma_block_k64(uint32_t tmem_d, uint32_t a_addr, uint32_t rows_a, uint32_t b_addr, uint32_t rows_b,
                                              uint32_t idesc, bool accumulate_first, bool swap = false)
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_desc(a_addr + k * 2 * rows_a * 16, rows_a, swap);
        const uint64_t db = make_desc(b_addr + k * 2 * rows_b * 16, rows_b, swap);
        mma_bf16(tmem_d, da, db, idesc, (accumulate_first || k > 0) ? 1u : 0u);
    }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

}  // namespace umma

// Host side (tcstatus.cu): the watchdog status of the CURRENT device for a launch (allocates the device's pinned,
// host-mapped status block on first use) and the SM count of the current device.  Return 0 on success.
int tc_launch_status(umma::TcStatus *out);
int tc_num_sms();
}  // namespace al3d
