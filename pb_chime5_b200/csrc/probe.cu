// Developer API (libgss_dev.so only): FP64 peak probe -- the denominator of the EM kernel's FP64
// roofline view, measured on the box the bench runs on (bench.py `roofline.fp64`).
#include "common.cuh"
#include "../../include/gss_dev.h"

namespace gss {

__global__ void __launch_bounds__(512) fp64_dfma_probe_kernel(double* out, int iters) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 2: the operand pattern of the EM kernel's inner loops -- every DFMA reads three DISTINCT
// registers (accumulator, a per-frame product, a per-class coefficient), 40 independent chains
__global__ void __launch_bounds__(512) fp64_dfma3_probe_kernel(double* out, int iters) {
    double acc[8][5], p[8], w[5];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        p[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-3;
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[i][k] = i + k;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = 1e-9 * (k + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i][k] = fma(w[k], p[i], acc[i][k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] += 1e-12;          // 8 DADD per 40 DFMA keep the products live
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 5; ++k) s += acc[i][k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// modes 6 / 7: the M-phase pattern -- 8 float32 operands widened per 40 DFMA -- with the hardware
// conversion (F2F.F64.F32) and with integer bit manipulation (exact for normal numbers and zero)
__device__ __forceinline__ double widen_bits(float f) {
    const unsigned x = __float_as_uint(f), m = x & 0x7fffffffu;
    const unsigned hi = (x & 0x80000000u) | (m ? (m >> 3) + 0x38000000u : 0u);
    return __hiloint2double((int)hi, (int)(x << 29));
}
template <bool BITS>
__global__ void __launch_bounds__(512) fp64_widen_probe_kernel(double* out, int iters) {
    double acc[8][5], w[5];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[i] = 1.0f + threadIdx.x * 1e-4f + i * 1e-2f;
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[i][k] = i + k;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = 1e-9 * (k + 1);
    for (int it = 0; it < iters; ++it) {
        double p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { f[i] += 1e-6f; p[i] = BITS ? widen_bits(f[i]) : (double)f[i]; }
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i][k] = fma(w[k], p[i], acc[i][k]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 5; ++k) s += acc[i][k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) fp64_dmma_probe_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
    const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace gss

extern "C" size_t gss_debug_fp64_peak_scratch_bytes(void) {
    return (size_t)gss::num_sms() * 2 * 512 * sizeof(double);
}

extern "C" int gss_debug_fp64_peak(int mode, int iters, void* scratch, double* flops_out, void* stream) {
    using namespace gss;
    GSS_REQUIRE(scratch && iters > 0 && mode >= 0 && mode <= 7, GSS_ERR_ARG, "gss_debug_fp64_peak: bad arguments");
    // 32 warps per SM (modes 0-2); mode 2 at 16 (mode 3: the EM kernel's occupancy), 8 (mode 4) and 4 (mode 5) warps per SM
    const int block = mode >= 6 ? 256 : mode == 3 ? 256 : mode == 4 ? 128 : mode == 5 ? 64 : 512, grid = num_sms() * 2;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) fp64_dfma_probe_kernel<<<grid, block, 0, st>>>((double*)scratch, iters);
    else if (mode == 1) fp64_dmma_probe_kernel<<<grid, block, 0, st>>>((double*)scratch, iters);
    else if (mode == 6) fp64_widen_probe_kernel<false><<<grid, block, 0, st>>>((double*)scratch, iters);
    else if (mode == 7) fp64_widen_probe_kernel<true><<<grid, block, 0, st>>>((double*)scratch, iters);
    else fp64_dfma3_probe_kernel<<<grid, block, 0, st>>>((double*)scratch, iters);
    GSS_LAUNCH_CHECK("fp64_probe_kernel");
    if (flops_out)
        *flops_out = mode == 0 ? 2.0 * 8 * iters * (double)grid * block
                   : mode == 1 ? 2.0 * 8 * 256 * iters * (double)grid * (block / 32)
                   : mode >= 6 ? 2.0 * 40 * iters * (double)grid * block
                               : (2.0 * 40 + 8) * iters * (double)grid * block;
    return GSS_OK;
}
