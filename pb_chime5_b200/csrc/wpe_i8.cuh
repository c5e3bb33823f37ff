// Interface between wpe.cu and the INT8 tensor-core Gram path (wpe_gram_i8.cu).
#pragma once
#include "common.cuh"

namespace gss {

struct WpeDims { int F, D, T, L, delay, LD; const int* Tper; };   // T: frame stride; Tper: valid frames per utterance or null

__device__ __forceinline__ int wpe_valid_frames(const WpeDims& m, size_t bf) {
    return m.Tper ? min(max(m.Tper[bf / m.F], 0), m.T) : m.T;
}

// Scratch of the INT8 path (device pointers carved from the caller's workspace).
struct WpeI8Ws {
    int8_t* slices;     // [chunk_bins][NS][KB][NRp][16]  digit planes of one chunk of bins (L2 sized)
    double* mu;         // [chunk_bins][T]                sqrt(inv)
    int* ex;            // [chunk_bins][NRc]              power-of-two exponent of every complex row
    int chunk_bins;     // bins per chunk
};

// bytes of the INT8 scratch for these dimensions (0 if the path does not apply)
size_t wpe_i8_ws_bytes(int F, int D, int T, int L);
// carve the scratch out of p (may be null: sizes only); returns the bytes used
size_t wpe_i8_ws_layout(void* p, int F, int D, int T, int L, WpeI8Ws* out);
// true if the INT8 path is built for this shape
bool wpe_i8_applicable(int D, int T, int L);

// Raug[bf] (lower trapezoid, same contract as wpe_corr_kernel) for bf in [0, BF);
// rdiag[bf][i] = Re R_ii (may be null).  skip (device, [BF], may be null): bins whose flag is set are
// left untouched (they are on the float64 list of the a-posteriori check).
int wpe_gram_i8_run(const float2* Y, const double* inv, cd* Raug, double* rdiag, const WpeDims& m, int BF,
                    const WpeI8Ws& ws, const int* skip, cudaStream_t st);

}  // namespace gss
