// Developer microbenchmark (not part of the product): issue rate of tcgen05.mma on this part for
// kind::i8 (INT8 x INT8 -> INT32) and kind::f16 (BF16 -> FP32), M = 128, cta_group::1, operands
// resident in shared memory (canonical K-major no-swizzle layout), no loads in the loop.
// Gives the tensor-pipe denominators for wpe_gram_i8_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_probe tools/mma_rate_probe.cu && ./mma_rate_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

template <int KIND>   // 0: i8, 1: bf16
__global__ void __launch_bounds__(128, 1) probe(int n, int rounds, int per_round, unsigned long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 32 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        // A: 128 rows x 32 B, B: n rows x 32 B; core matrices 128 B, k halves 128 B apart is not possible here
        // (rows x 16 B per half): LBO = rows * 16, SBO = 128
        const uint64_t ad = umma_desc(smem_u32(smem), 128 * 16, 128);
        const uint64_t bd = umma_desc(smem_u32(smem) + 128 * 32, n * 16, 128);
        uint32_t idesc;
        if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
        else idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);   // F32 acc, BF16 x BF16
        const unsigned long long t0 = clock64();
        uint32_t parity = 0;
        for (int r = 0; r < rounds; ++r) {
            for (int i = 0; i < per_round; ++i) {
                const uint32_t acc = i > 0;
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + (uint32_t)((i & 1) * 256)), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + (uint32_t)((i & 1) * 256)), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done = 0;
            for (unsigned spin = 0; !done; ++spin) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
                if (spin > (1u << 24)) __trap();
            }
            parity ^= 1;
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

template <int KIND>
static void run(const char* name, int n, int kelems) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* cyc;
    cudaMalloc(&cyc, sms * sizeof(unsigned long long));
    const int rounds = 200, per = 64;
    const size_t smem = (128 + 256) * 32 + 1024;
    cudaFuncSetAttribute(probe<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<KIND><<<sms, 128, smem>>>(n, 10, per, cyc);
    cudaEventRecord(e0);
    probe<KIND><<<sms, 128, smem>>>(n, rounds, per, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double macs = (double)sms * rounds * per * 128.0 * n * kelems;
    printf("%-5s M=128 N=%3d K=%2d: %8.1f T(FL)OP/s dense, %6.1f clocks per MMA (SM 0)  [%s]\n", name, n, kelems,
           2.0 * macs / (ms * 1e-3) / 1e12, (double)h / (rounds * per), cudaGetErrorString(err));
    cudaFree(cyc);
}

int main() {
    for (int n : {256, 128, 96, 64}) run<0>("i8", n, 32);
    for (int n : {256, 128, 96}) run<1>("bf16", n, 16);
    return 0;
}
