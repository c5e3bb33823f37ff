// Microbenchmark: FP64 FMA pipe vs FP64 tensor (DMMA m8n8k4) vs both, per-SM throughput.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters) {
    double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma(double* out, int iters) {
    double c[8][2]; for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_both(double* out, int iters) {
    double c[4][2]; for (int i = 0; i < 4; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double f[8]; for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6, cc = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { dmma(c[i][0], c[i][1], a, b); f[2*i] = fma(f[2*i], b, cc); f[2*i+1] = fma(f[2*i+1], b, cc); }
    }
    double s = 0; for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1]; for (int i = 0; i < 8; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_cvt(double* out, const float* in, int iters) {
    float x[8]; for (int i = 0; i < 8; ++i) x[i] = in[threadIdx.x + i];
    double s = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s += (double)x[i]; x[i] += 1.0f; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    float* in; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        dim3 grid(sms * 2), block(warps * 16);   // 2 CTAs per SM
        float ms;
        k_dfma<<<grid, block>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dfma<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double)grid.x * block.x;
        printf("warps/SM %2d  DFMA  %.2f TFLOP/s", warps, fl / ms / 1e9);
        k_dmma<<<grid, block>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dmma<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 8 * 256 * iters * (double)grid.x * (block.x / 32);
        printf("   DMMA %.2f TFLOP/s", fl / ms / 1e9);
        k_both<<<grid, block>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_both<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 4 * 256 * iters * (double)grid.x * (block.x / 32) + 2.0 * 8 * iters * (double)grid.x * block.x;
        printf("   both %.2f TFLOP/s", fl / ms / 1e9);
        cudaEventRecord(e0); k_cvt<<<grid, block>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("   F2F.F64.F32+DADD %.1f Gconv/s  (%s)\n", 8.0 * iters * (double)grid.x * block.x / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
