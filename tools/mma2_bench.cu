// mma2_bench.cu -- tcgen05.mma.cta_group::2 (CTA pair, M = 256) microbenchmark and semantics probe (bring-up tool for the
// planned pair version of the RRDB kernel, not part of the product).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma2_bench tools/mma2_bench.cu && build/mma2_bench
//
// Clusters of two CTAs.  Each CTA holds its own A rows (128 x K, SW128 K-major) and HALF of the B rows (N/2 x K) at the
// same shared-memory offsets; the leader (cluster rank 0) issues groups of 12 MMAs (M 256, N, K 16, kind::f16) and
// commits each group with a multicast arrive to a barrier ring in both CTAs.  Reports cycles per MMA, and checks where the
// two halves of B land: A = 1 everywhere, B = 1 in rank 0 and 2 in rank 1, so accumulator columns [0, N/2) must read
// 16 * k and [N/2, N) must read 32 * k in BOTH CTAs after k accumulating MMAs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x)                                                   \
    do {                                                           \
        cudaError_t e_ = (x);                                      \
        if (e_ != cudaSuccess) {                                   \
            printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); \
            exit(1);                                               \
        }                                                          \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(par)
                 : "memory");
    return ok;
}

constexpr int A_BYTES = 65536, B_BYTES = 98304;

template <int N, int PER_GROUP>
__global__ void __launch_bounds__(192, 1) mma2_bench_kernel(int groups, int a_row_step, int b_step, long long* out_cycles, float* out_vals) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (sbase - raw);
    const uint32_t a_s = sbase, b_s = sbase + A_BYTES;
    const uint32_t bar = b_s + B_BYTES;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(g + A_BYTES + B_BYTES + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    // A = 1.0 everywhere; B = 1.0 in rank 0, 2.0 in rank 1 (fp16 bit patterns 0x3c00 / 0x4000)
    for (int i = threadIdx.x; i < A_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0x3c003c00u;
    for (int i = threadIdx.x; i < B_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g + A_BYTES)[i] = rank ? 0x40004000u : 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // one warp in EACH CTA of the pair performs the pair allocation
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    long long cycles = 0;
    if (warp == 0 && rank == 0) {
        const uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
        const uint64_t hi64 = (uint64_t)hi << 32;
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);  // D f32, A = B = f16, M = 256
        const uint32_t a_lo0 = (a_s >> 4) | (1u << 16), b_lo = (b_s >> 4) | (1u << 16);
        const long long t0 = clock64();
        int arow = 0;
        for (int gi = 0; gi < groups; ++gi) {
            if (gi >= 4) {
                const uint32_t b = bar + 8 * (gi & 3), par = ((gi >> 2) - 1) & 1;
                while (!try_wait(b, par)) {
                }
            }
            const uint32_t d = tmem + (uint32_t)((gi % (512 / N)) * N);
            const uint32_t a_lo = a_lo0 + (uint32_t)(arow >> 4);
            arow += a_row_step;
            if (arow >= 32768) arow = 0;
            if (elect_one()) {
#pragma unroll
                for (int m = 0; m < PER_GROUP; ++m) {
                    const uint32_t ao = (uint32_t)(((m / 4) * 128 + (m % 4) * 32) >> 4);
                    umma2(d, hi64 | (a_lo + ao), hi64 | (b_lo + (uint32_t)((m * b_step) >> 4)), idesc, (gi >= 512 / N) || m != 0);
                }
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                                 bar + 8 * (gi & 3)),
                             "h"((uint16_t)3)
                             : "memory");
            }
            __syncwarp();
        }
        for (int c = groups > 4 ? groups - 4 : 0; c < groups; ++c) {
            while (!try_wait(bar + 8 * (c & 3), (c >> 2) & 1)) {
            }
        }
        cycles = clock64() - t0;
        if (lane == 0) out_cycles[blockIdx.x >> 1] = cycles;
    } else if (warp == 0 && rank == 1) {
        // the peer only watches the multicast commits arrive on ITS copy of the barrier ring
        for (int c = 0; c < groups; ++c) {
            long long t0 = clock64();
            while (!try_wait(bar + 8 * (c & 3), (c >> 2) & 1)) {
                if (clock64() - t0 > 4000000000LL) {
                    if (lane == 0) printf("peer: commit %d never arrived\n", c);
                    __trap();
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp >= 2 && warp < 6 && blockIdx.x < 2) {
        // each CTA reads accumulator block 0 of its own 128 rows: columns 0 .. N-1 of lane (warp & 3) * 32 + lane
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        for (int j = 0; j < N; j += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr + (uint32_t)j));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (lane == 5 && (warp & 3) == 1)
                for (int e = 0; e < 16; ++e) out_vals[rank * 256 + j + e] = __uint_as_float(r[e]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int N, int PG>
static void run(const char* name, int groups, int a_row_step, int b_step, int pairs) {
    long long* d;
    float* v;
    CHECK(cudaMalloc(&d, 256 * sizeof(long long)));
    CHECK(cudaMemset(d, 0, 256 * sizeof(long long)));
    CHECK(cudaMalloc(&v, 512 * sizeof(float)));
    CHECK(cudaMemset(v, 0, 512 * sizeof(float)));
    const int smem = 1024 + A_BYTES + B_BYTES + 256;
    auto kern = mma2_bench_kernel<N, PG>;
    CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs), cfg.blockDim = dim3(192), cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    CHECK(cudaLaunchKernelEx(&cfg, kern, groups, a_row_step, b_step, d, v));
    CHECK(cudaDeviceSynchronize());
    long long h[128];
    float hv[512];
    CHECK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(hv, v, sizeof hv, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < pairs; ++i) mean += h[i];
    mean /= pairs;
    const double n = (double)groups * PG;
    // block 0 received the groups gi with gi % (512 / N) == 0: the first of them overwrites (m = 0), the rest accumulate
    const int nblk = 512 / N, uses = (groups + nblk - 1) / nblk;
    const double k_acc = (double)PG * uses;
    printf("%-44s N=%3d x%2d pairs %2d cyc/MMA %6.1f (ideal %5.1f)  MAC/clk/SM %5.0f = %3.0f%% | rank0 D[0]=%g D[N/2]=%g rank1 D[0]=%g D[N/2]=%g (expect %g / %g)\n",
           name, N, PG, pairs, mean / n, N / 2.0, 256.0 * N * 16 * n / mean / 2, 100.0 * 256.0 * N * 16 * n / mean / 2 / 4096.0, hv[0], hv[N / 2],
           hv[256], hv[256 + N / 2], 16.0 * k_acc, 32.0 * k_acc);
    cudaFree(d);
    cudaFree(v);
}

int main() {
    const int G = 2000;
    for (int pairs : {1, 72}) {
        printf("---- %d pair(s) ----\n", pairs);
        run<96, 12>("pair N=96 (RRDB 32-ch layers)", G, 17408, 3072 / 2, pairs);
        run<96, 36>("pair N=96, 36 MMAs per group", G, 17408, 3072 / 2, pairs);
        run<192, 12>("pair N=192 (64-ch layers, 192->64)", G, 17408, 6144 / 2, pairs);
        run<192, 36>("pair N=192, 36 MMAs per group", G, 17408, 6144 / 2, pairs);
        run<256, 12>("pair N=256", G, 17408, 8192 / 2, pairs);
    }
    return 0;
}
