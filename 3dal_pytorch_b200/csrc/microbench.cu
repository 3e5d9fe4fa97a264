// Development aid: tcgen05.mma issue / execution rate microbenchmark (one issuing thread per CTA, every SM busy).
// Answers the design questions DESIGN.md cites: cycles per MMA as a function of N and of the operand source
// (shared-memory descriptor vs TMEM), the depth of the issue queue, and the cost of interleaved commits.
#include "common.cuh"
#include "umma.cuh"

namespace al3d {
using namespace umma;

struct MicroSmem {
    uint8_t a[16384];      // 128 rows x 64 K
    uint8_t b[32768];      // 256 rows x 64 K
    uint8_t ring[2][16384]; // background TMA traffic lands here
    uint64_t bar[8];
    uint64_t done;
    uint64_t ring_full[2];
    uint32_t tmem_base;
    volatile uint32_t stop;
};

// MODE 0: A and B from shared memory; MODE 1: A from TMEM (columns 448..479), B from shared memory.
// CE: commits per group of 8 MMAs (0 = one commit at the very end, 1 = after each group, 2 = after every 4, 8 = every MMA).
template <int MODE, int CE>
__global__ void __launch_bounds__(256, 1)
mma_microbench_kernel(int N, int n_groups, int bg, const uint8_t *src, long long *out, const TcStatus wd)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    MicroSmem &s = *reinterpret_cast<MicroSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 256)
        reinterpret_cast<uint32_t *>(s.a)[i] = 0x3c003c00u + (uint32_t)(i * 2654435761u >> 28);   // small bf16 values
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&s.bar[i], 1);
        mbar_init(&s.done, 1);
        mbar_init(&s.ring_full[0], 1); mbar_init(&s.ring_full[1], 1);
        s.stop = 0;
        fence_barrier_init();
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<512>(&s.tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    auto issue = [&]() {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t a_addr = smem_u32(s.a), b_addr = smem_u32(s.b);
        uint64_t da[4], db[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { da[k] = make_desc(a_addr + k * 4096, 128); db[k] = make_desc(b_addr + k * 2 * N * 16, N); }
        const uint32_t d1 = tmem + (N > 128 ? 0 : 128);
        long long t_first = 0;
        const long long t0 = clock64();
#pragma unroll 1
        for (int g = 0; g < n_groups; ++g) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t d = (k < 4) ? tmem : d1;
                if (MODE == 0) mma_bf16(d, da[k & 3], db[k & 3], idesc, 1u);
                else           mma_bf16_ts(d, tmem + 448 + (k & 3) * 8, db[k & 3], idesc, 1u);
                if (CE == 8 || (CE == 2 && (k & 3) == 3) || (CE == 1 && k == 7)) mma_commit(&s.bar[k]);
            }
            if (g == 0) t_first = clock64();
        }
        mma_commit(&s.done);
        const long long t1 = clock64();
        mbar_wait(&s.done, 0, 0xB000, wd);
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;              // issue time of the whole sequence
            out[1] = t2 - t0;              // until the last MMA completed
            out[2] = t_first - t0;         // issue time of the first 8 MMAs (queue empty)
        }
        s.stop = 1;
    };
    if (warp == 1) {
        // the same loop entered through `lane == 0` or through elect.sync: ptxas wraps every tcgen05 instruction of the
        // former in an ELECT / branch loop (it cannot prove a single active lane)
        if (bg & 4) { if (elect_one_sync()) issue(); }
        else if (threadIdx.x == 32) issue();
    } else if (threadIdx.x == 0 && (bg & 2)) {
        // background: stream 16 KB blocks global -> shared like a weight ring (src: >= 1 MiB)
        uint32_t ph[2] = {0, 0};
        int i = 0;
        long long blocks = 0;
        while (!s.stop) {
            const int b = i & 1;
            mbar_arrive_expect_tx(&s.ring_full[b], 16384);
            bulk_g2s(s.ring[b], src + (size_t)((i * 148 + blockIdx.x) & 63) * 16384, 16384, &s.ring_full[b]);
            if (i > 0) { mbar_wait(&s.ring_full[b ^ 1], ph[b ^ 1], 0xB001, wd); ph[b ^ 1] ^= 1; }
            ++i; ++blocks;
        }
        mbar_wait(&s.ring_full[(i - 1) & 1], ph[(i - 1) & 1], 0xB002, wd);
        if (blockIdx.x == 0) out[3] = blocks;
    } else if (warp >= 4 && (bg & 1)) {
        // background: epilogue-like TMEM reads of the accumulator columns
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc = 0;
        long long loads = 0;
        while (!s.stop) {
            uint32_t v[32];
            tmem_ld32(tl + ((loads & 7) * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc ^= v[k];
            ++loads;
        }
        if (acc == 0x12345678u) out[5] = acc;
        if (blockIdx.x == 0 && threadIdx.x == 128) out[4] = loads;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int MODE, int CE>
static int launch_micro(int N, int n_groups, int n_ctas, int bg, const uint8_t *src, long long *out, cudaStream_t stream)
{
    const size_t smem = sizeof(MicroSmem) + 128;
    AL3D_CHECK_CUDA(cudaFuncSetAttribute(mma_microbench_kernel<MODE, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcStatus wd;
    if (tc_launch_status(&wd)) return 1;
    mma_microbench_kernel<MODE, CE><<<n_ctas, 256, smem, stream>>>(N, n_groups, bg, src, out, wd);
    AL3D_CHECK_LAUNCH("mma_microbench_kernel");
    return 0;
}

}  // namespace al3d

extern "C" int al3d_mma_microbench(int N, int n_mma, int commit_every, int mode, int n_ctas, int background, const void *src_1mib,
                                   long long *out, void *stream)
{
    using namespace al3d;
    AL3D_CHECK_ARG(out && (N == 64 || N == 128 || N == 256) && n_mma >= 8 && n_ctas >= 1, "al3d_mma_microbench: bad arguments");
    AL3D_CHECK_ARG(commit_every == 0 || commit_every == 1 || commit_every == 4 || commit_every == 8, "al3d_mma_microbench: commit_every");
    AL3D_CHECK_ARG(mode == 0 || mode == 1, "al3d_mma_microbench: mode");
    AL3D_CHECK_ARG(!(background & 2) || src_1mib, "al3d_mma_microbench: background TMA traffic needs a source buffer");
    const int g = n_mma / 8;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        if (commit_every == 0) return launch_micro<0, 0>(N, g, n_ctas, background, (const uint8_t *)src_1mib, out, st);
        if (commit_every == 8) return launch_micro<0, 1>(N, g, n_ctas, background, (const uint8_t *)src_1mib, out, st);
        if (commit_every == 4) return launch_micro<0, 2>(N, g, n_ctas, background, (const uint8_t *)src_1mib, out, st);
        return launch_micro<0, 8>(N, g, n_ctas, background, (const uint8_t *)src_1mib, out, st);
    }
    if (commit_every == 0) return launch_micro<1, 0>(N, g, n_ctas, background, (const uint8_t *)src_1mib, out, st);
    if (commit_every == 8) return launch_micro<1, 1>(N, g, n_ctas, background, (const uint8_t *)src_1mib, out, st);
    if (commit_every == 4) return launch_micro<1, 2>(N, g, n_ctas, background, (const uint8_t *)src_1mib, out, st);
    return launch_micro<1, 8>(N, g, n_ctas, background, (const uint8_t *)src_1mib, out, st);
}
