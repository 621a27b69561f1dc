// tmem_wrap_probe.cu -- does a tcgen05.mma whose N columns run past TMEM column 511 wrap around to column 0?
// (bring-up probe, not part of the product)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_wrap_probe tools/tmem_wrap_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(float* out, int dcol, int N) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (sbase - raw);
    const uint32_t a_s = sbase, b_s = sbase + 16384, bar = sbase + 16384 + 32768;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(g + 16384 + 32768 + 64);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0x3c003c00u;  // 1.0h
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    // zero all 512 columns of this warp's lanes
    for (int c = 0; c < 512; c += 16)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
                         tmem + ((uint32_t)(warp * 32) << 16) + c),
                     "r"(0u)
                     : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t ad = ((uint64_t)hi << 32) | ((a_s >> 4) & 0x3fffu) | (1u << 16);
        const uint64_t bd = ((uint64_t)hi << 32) | ((b_s >> 4) & 0x3fffu) | (1u << 16);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)dcol),
            "l"(ad), "l"(bd), "r"(idesc), "r"(1u)
            : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(bar), "r"(0u)
                         : "memory");
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // dump lane (warp*32+lane) columns 0..511
    for (int c = 0; c < 512; c += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(size_t)threadIdx.x * 512 + c + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 512 * 4);
    static float h[128 * 512];
    const int smem = 1024 + 16384 + 32768 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int dcol : {0, 320, 384, 448}) {
        cudaMemset(d, 0, sizeof h);
        probe<<<1, 128, smem>>>(d, dcol, 192);
        cudaError_t e = cudaDeviceSynchronize();
        printf("dcol %3d N 192: %s\n", dcol, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        for (int lane : {0, 77}) {
            printf("  lane %3d nonzero column ranges:", lane);
            int start = -1;
            for (int c = 0; c <= 512; ++c) {
                const bool nz = c < 512 && h[lane * 512 + c] != 0.f;
                if (nz && start < 0) start = c;
                if (!nz && start >= 0) {
                    printf(" [%d,%d) = %.0f", start, c, h[lane * 512 + start]);
                    start = -1;
                }
            }
            printf("\n");
        }
    }
    return 0;
}
