// Shared helpers for libgss (sm_100a).  Host side: error plumbing for the C ABI.
// Device side: complex128 arithmetic on double2 and pair indexing.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include "../../include/gss.h"

namespace gss {

// ---- host: thread-local last-error string ---------------------------------
char* last_error_buf();
int fail(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch();   // kernel launch counter (gss_launch_count)

#define GSS_REQUIRE(cond, code, ...) \
    do { if (!(cond)) return ::gss::fail((code), __VA_ARGS__); } while (0)
#define GSS_CUDA(expr) \
    do { int _rc = ::gss::check_cuda((expr), #expr); if (_rc) return _rc; } while (0)
#define GSS_LAUNCH_CHECK(name) \
    do { ::gss::count_launch(); int _rc = ::gss::check_cuda(cudaGetLastError(), name); if (_rc) return _rc; } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
    char* base; size_t size; size_t off;
    Arena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
    template <typename T> T* take(size_t n) {
        size_t bytes = align_up(n * sizeof(T));
        T* r = (T*)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= size; }
};

int num_sms();

// ---- device helpers ---------------------------------------------------------
typedef double2 cd;   // complex128: x = re, y = im

__host__ __device__ __forceinline__ cd cmake(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return cmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
__device__ __forceinline__ cd cmulc(cd a, cd b) { return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
// conj(a) * b
__device__ __forceinline__ cd ccmul(cd a, cd b) { return cmake(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ cd cscale(cd a, double s) { return cmake(a.x * s, a.y * s); }
__device__ __forceinline__ cd cconj(cd a) { return cmake(a.x, -a.y); }
__device__ __forceinline__ double cabs2(cd a) { return a.x * a.x + a.y * a.y; }
// acc += a * b
__device__ __forceinline__ void cfma(cd& acc, cd a, cd b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += a * conj(b)
__device__ __forceinline__ void cfmac(cd& acc, cd a, cd b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.y, b.x, acc.y); acc.y = fma(-a.x, b.y, acc.y);
}
// acc -= a * conj(b)
__device__ __forceinline__ void cfmsc(cd& acc, cd a, cd b) {
    acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(-a.y, b.x, acc.y); acc.y = fma(a.x, b.y, acc.y);
}
// acc -= a * b
__device__ __forceinline__ void cfms(cd& acc, cd a, cd b) {
    acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}

// Packed lower-triangular index (row d >= col e).
__host__ __device__ __forceinline__ constexpr int tri(int d, int e) { return d * (d + 1) / 2 + e; }

#define GSS_F64_TINY 2.2250738585072014e-308

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace gss
