/*
 * libgss_dev -- developer / measurement entry points.  NOT part of the product library: these
 * symbols exist only in pb_chime5_b200/csrc/libgss_dev.so (libgss.so + this API), which the tests,
 * tools/ and the roofline side measurements of bench.py load.  Nothing in pb_chime5_b200/ calls them.
 */
#ifndef GSS_DEV_H_
#define GSS_DEV_H_
#include "gss.h"
#ifdef __cplusplus
extern "C" {
#endif

/* one WPE correlation build: Y (B,F,D,T) c64, inv (B,F,T) f64 -> Raug (B,F,taps*D+D,taps*D) c128,
 * lower trapezoid (rows [0,LD): R, rows [LD,LD+D): P^H); mode 0 = float64 (DMMA), 1 = INT8 tensor
 * cores.  Workspace: gss_workspace_bytes(GSS_OP_WPE, ...). */
int gss_debug_wpe_gram(const gss_c64* Y, const double* inv, double* Raug, int mode, int variant,
                       int B, int F, int D, int T, int taps, int delay, const int* T_per_utt,
                       void* ws, size_t ws_bytes, void* stream);

/* EXPERIMENTAL building block (DESIGN.md "INT8 EM"), not used by any product entry point: the
 * CACGMM M-step covariance Phi[b,f,k] = sum_t w[b,f,k,t] y y^H
 * (complex_angular_central_gaussian.py:293-300) on the INT8 tensor cores with exact digit-split
 * arithmetic.  Y (B,F,D,T) c64, w (B,F,K,T) f64 >= 0, Phi (B,F,K,D,D) c128 (full Hermitian).
 * Built for D in {4, 8, 16, 24}, K * 2 D <= 256; workspace B F (ceil(T/32) 320 D + 4 (D + K)) + 1 KiB. */
int gss_debug_mstep_i8(const gss_c64* Y, const double* w, double* Phi, int B, int F, int D, int T, int K,
                       const int* T_per_utt, void* ws, size_t ws_bytes, void* stream);

/* FP64 peak probe (the roofline denominator of the EM kernel, measured on the bench box):
 * launches one kernel of dependent-chain-free FP64 work on every SM; mode 0 = DFMA with two
 * loop-invariant operands (the textbook peak), 1 = DMMA (mma.sync.m8n8k4.f64), 2 = DFMA with three
 * distinct register operands per instruction (the operand pattern of the EM kernel's M phase:
 * 40 accumulators += weight[k] * product[i]), 3 / 4 / 5 = mode 2 at 16 (the EM kernel's occupancy) / 8 / 4 warps per SM,
 * 6 / 7 = mode 3 with 8 float32 operands widened per 40 DFMA by F2F / by integer bit manipulation (DFMA flops only).
 * *flops_out (host) = floating point operations of the launch.
 * scratch: device, >= gss_debug_fp64_peak_scratch_bytes() bytes.  Time it with events on `stream`. */
size_t gss_debug_fp64_peak_scratch_bytes(void);
int gss_debug_fp64_peak(int mode, int iters, void* scratch, double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* GSS_DEV_H_ */
