// mma_bench.cu -- tcgen05.mma throughput microbenchmark (bring-up tool, not part of the product).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_bench tools/mma_bench.cu && build/mma_bench
//
// One CTA per SM.  Warp 0 issues groups of 12 MMAs (M=128, N, K=16, kind::f16, SS operands, 128B swizzle) with a
// warp-uniform, fully unrolled issue loop (the shape of the real kernel's issuer) and commits each group to an
// mbarrier ring (4 groups in flight).  Optional interference: warps 2..5 stream tcgen05.ld over the accumulators
// (what the epilogue does), warps 6..9 stream LDS/STS over a scratch region (staging traffic).
// Reports cycles per MMA and MAC/clk/SM.
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
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
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

constexpr int A_BYTES = 65536, B_BYTES = 98304, X_BYTES = 32768;

// a_kx: bytes the A start moves per kx group of 4 MMAs (128 = one pixel); b_step: bytes between B tiles of successive MMAs
template <int N, int PER_GROUP>
__global__ void __launch_bounds__(320, 1) mma_bench_kernel(int groups, int a_row_step, int b_step, int tmem_readers, int lsu_level,
                                                            int commit_every, long long* out_cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (sbase - raw);
    const uint32_t a_s = sbase, b_s = sbase + A_BYTES, x_s = b_s + B_BYTES;
    const uint32_t bar = x_s + X_BYTES;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(g + A_BYTES + B_BYTES + X_BYTES + 64);
    volatile int* stop = reinterpret_cast<volatile int*>(g + A_BYTES + B_BYTES + X_BYTES + 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES + X_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *stop = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    if (warp == 0) {
        const uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
        const uint64_t hi64 = (uint64_t)hi << 32;
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_lo0 = (a_s >> 4) | (1u << 16), b_lo = (b_s >> 4) | (1u << 16);
        const long long t0 = clock64();
        int arow = 0;
        for (int gi = 0; gi < groups; ++gi) {
            if (gi >= 4 && (gi % commit_every) == 0) {
                const uint32_t b = bar + 8 * ((gi / commit_every) & 3), par = (((gi / commit_every) >> 2) - 1) & 1;
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
                    umma(d, hi64 | (a_lo + ao), hi64 | (b_lo + (uint32_t)((m * b_step) >> 4)), idesc, m != 0);
                }
                if ((gi % commit_every) == commit_every - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     bar + 8 * ((gi / commit_every) & 3))
                                 : "memory");
            }
            __syncwarp();
        }
        const int ncommit = groups / commit_every;
        for (int c = ncommit > 4 ? ncommit - 4 : 0; c < ncommit; ++c) {
            while (!try_wait(bar + 8 * (c & 3), (c >> 2) & 1)) {
            }
        }
        const long long t1 = clock64();
        if (lane == 0) {
            out_cycles[blockIdx.x] = t1 - t0;
            *stop = 1;
        }
    } else if (warp >= 2 && warp < 6) {
        if (tmem_readers) {
            // epilogue-like TMEM reads: 64 columns of this warp's lane quadrant per iteration
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            uint32_t acc = 0;
            int it = 0;
            while (!*stop) {
                uint32_t r[16];
#pragma unroll
                for (int j = 0; j < 64; j += 16) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr + (uint32_t)(((it & 7) * 64 + j) & 511)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    acc ^= r[3];
                }
                ++it;
                if (tmem_readers == 1) __nanosleep(600);  // paced roughly like one row per ~1000-2000 cycles
            }
            if (acc == 0x12345) out_cycles[1500] = acc;
        }
    } else if (warp >= 6) {
        if (lsu_level) {
            uint4* x = reinterpret_cast<uint4*>(g + A_BYTES + B_BYTES) + (warp - 6) * 512;
            uint4 v = make_uint4(lane, 1, 2, 3);
            while (!*stop) {
#pragma unroll 8
                for (int r = 0; r < 16; ++r) {
                    x[(r * 32 + lane) & 511] = v;
                    uint4 w = x[((r + 5) * 32 + lane) & 511];
                    v.x ^= w.y;
                }
                if (lsu_level == 1) __nanosleep(400);
            }
            if (v.x == 0x12345) out_cycles[1600] = v.x;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int N, int PG>
static void run(const char* name, int groups, int a_row_step, int b_step, int readers, int lsu, int commit_every, int grid) {
    long long* d;
    CHECK(cudaMalloc(&d, 2048 * sizeof(long long)));
    CHECK(cudaMemset(d, 0, 2048 * sizeof(long long)));
    const int smem = 1024 + A_BYTES + B_BYTES + X_BYTES + 256;
    CHECK(cudaFuncSetAttribute(mma_bench_kernel<N, PG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_bench_kernel<N, PG><<<grid, 320, smem>>>(groups, a_row_step, b_step, readers, lsu, commit_every, d);
    CHECK(cudaDeviceSynchronize());
    long long h[148];
    CHECK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
    double mx = 0, mean = 0;
    for (int i = 0; i < grid && i < 148; ++i) {
        mean += h[i];
        if (h[i] > mx) mx = h[i];
    }
    mean /= grid;
    const double n = (double)groups * PG;
    printf("%-64s N=%3d x%2d cyc/MMA %6.1f (max %6.1f) ideal %5.1f  MAC/clk/SM %5.0f = %3.0f%%\n", name, N, PG, mean / n, mx / n, N / 2.0,
           128.0 * N * 16 * n / mean, 100.0 * 128.0 * N * 16 * n / mean / 4096.0);
    cudaFree(d);
}

int main() {
    const int G = 4000;
    for (int grid : {1, 148}) {
        printf("---- grid %d ----\n", grid);
        run<64, 12>("K slabs, same B (b_step 32)", G, 0, 32, 0, 0, 1, grid);
        run<128, 12>("K slabs, same B (b_step 32)", G, 0, 32, 0, 0, 1, grid);
        run<192, 12>("K slabs, same B (b_step 32)", G, 0, 32, 0, 0, 1, grid);
        run<256, 12>("K slabs, same B (b_step 32)", G, 0, 32, 0, 0, 1, grid);
        run<64, 36>("conv-like N=64: 36 MMAs/row, B tile 2 KB apart", G, 17408, 2048, 0, 0, 1, grid);
        run<192, 12>("conv-like N=192: new A row/group, B slabs 6 KB apart", G, 17408, 6144, 0, 0, 1, grid);
        run<192, 12>("  same, commit every 4 groups", G, 17408, 6144, 0, 0, 4, grid);
        run<192, 12>("  + TMEM readers (paced)", G, 17408, 6144, 1, 0, 1, grid);
        run<192, 12>("  + TMEM readers (flat out)", G, 17408, 6144, 2, 0, 1, grid);
        run<192, 12>("  + LSU smem traffic (paced)", G, 17408, 6144, 0, 1, 1, grid);
        run<192, 12>("  + LSU smem traffic (flat out)", G, 17408, 6144, 0, 2, 1, grid);
        run<192, 12>("  + TMEM readers + LSU (paced)", G, 17408, 6144, 1, 1, 1, grid);
        run<192, 12>("  + TMEM readers + LSU (flat out)", G, 17408, 6144, 2, 2, 1, grid);
        // RRDB kernel shapes (NOUT = 32 -> N = 96): one 64-channel group of one input row = 12 MMAs, then a commit
        run<96, 12>("RRDB N=96: 12 MMAs/group, B slabs 3 KB apart", G, 17408, 3072, 0, 0, 1, grid);
        run<96, 12>("  same, commit every 2 groups", G, 17408, 3072, 0, 0, 2, grid);
        run<96, 12>("  same, commit every 4 groups", G, 17408, 3072, 0, 0, 4, grid);
        run<96, 12>("  + TMEM readers + LSU (paced)", G, 17408, 3072, 1, 1, 1, grid);
        run<96, 24>("RRDB N=96: 24 MMAs/group", G, 17408, 3072, 0, 0, 1, grid);
        run<96, 36>("RRDB N=96: 36 MMAs/group", G, 17408, 3072, 0, 0, 1, grid);
        run<128, 12>("N=128 conv-like", G, 17408, 4096, 0, 0, 1, grid);
        run<48, 12>("N=48 conv-like (final 64->3 layer)", G, 17408, 1536, 0, 0, 1, grid);
        run<256, 12>("conv-like N=256", G, 17408, 8192, 0, 0, 1, grid);
        run<256, 12>("  + TMEM readers + LSU (paced)", G, 17408, 8192, 1, 1, 1, grid);
    }
    return 0;
}
