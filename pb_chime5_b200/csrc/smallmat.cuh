// Small (n <= 32) complex128 Hermitian linear algebra in shared memory.
//
// Replaces the LAPACK calls on the reference's hot path:
//   np.linalg.eigh      complex_angular_central_gaussian.py:96      -> block_jacobi_eigh
//   zhegvd (ITYPE=1)    cythonized/get_gev_vector.pyx:118-121       -> cholesky + tri-inverse + block_jacobi_eigh
//   np.linalg.solve     math/solve.py:98 (zgesv)                    -> block_lu_solve
// Warp-level routines: one warp owns one matrix, lane i owns row/column i.
// Block-level routines: all threads of the CTA cooperate on one matrix.
#pragma once
#include "common.cuh"

namespace gss {

// In-place Cholesky A = L L^H on a packed lower-triangular matrix (index tri(i,j)).
// Returns false (uniformly across the warp) if a pivot is not > 0.
__device__ inline bool warp_cholesky_packed(cd* A, int n, int lane) {
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        cd s = cmake(0.0, 0.0);
        if (lane >= j && lane < n) {
            s = A[tri(lane, j)];
            const cd* ri = A + tri(lane, 0);
            const cd* rj = A + tri(j, 0);
            cd s2 = cmake(0.0, 0.0);
            int p = 0;
            for (; p + 1 < j; p += 2) { cfmsc(s, ri[p], rj[p]); cfmsc(s2, ri[p + 1], rj[p + 1]); }
            if (p < j) cfmsc(s, ri[p], rj[p]);
            s = cadd(s, s2);
        }
        double djj = __shfl_sync(0xffffffffu, s.x, j);
        if (!(djj > 0.0) || !isfinite(djj)) { ok = false; break; }
        double r = sqrt(djj);
        double inv = 1.0 / r;
        if (lane == j) A[tri(j, j)] = cmake(r, 0.0);
        else if (lane > j && lane < n) A[tri(lane, j)] = cscale(s, inv);
        __syncwarp();
    }
    return ok;
}

// M = L^{-1} (both packed lower).  Lane c produces column c.  L and M distinct.
__device__ inline void warp_tri_inverse_packed(const cd* L, cd* M, int n, int lane) {
    if (lane < n) {
        const int c = lane;
        M[tri(c, c)] = cmake(1.0 / L[tri(c, c)].x, 0.0);
        for (int i = c + 1; i < n; ++i) {
            cd s = cmake(0.0, 0.0), s2 = cmake(0.0, 0.0);
            const cd* ri = L + tri(i, 0);
            int p = c;
            for (; p + 1 < i; p += 2) { cfma(s, ri[p], M[tri(p, c)]); cfma(s2, ri[p + 1], M[tri(p + 1, c)]); }
            if (p < i) cfma(s, ri[p], M[tri(p, c)]);
            s = cadd(s, s2);
            M[tri(i, c)] = cscale(s, -1.0 / ri[i].x);
        }
    }
    __syncwarp();
}

// In-place L <- L^{-1} on a packed lower-triangular matrix with real positive
// diagonal.  Lane i owns row i; columns are processed right to left.
__device__ inline void warp_tri_inverse_inplace(cd* L, int n, int lane) {
    for (int j = n - 1; j >= 0; --j) {
        const double mjj = 1.0 / L[tri(j, j)].x;
        cd s = cmake(0.0, 0.0), s2 = cmake(0.0, 0.0);
        if (lane > j && lane < n) {
            const cd* ri = L + tri(lane, 0);          // row `lane`: entries (lane, p), p > j already inverted
            int pp = j + 1;
            for (; pp + 1 <= lane; pp += 2) { cfma(s, ri[pp], L[tri(pp, j)]); cfma(s2, ri[pp + 1], L[tri(pp + 1, j)]); }
            if (pp <= lane) cfma(s, ri[pp], L[tri(pp, j)]);
            s = cadd(s, s2);
        }
        __syncwarp();
        if (lane > j && lane < n) L[tri(lane, j)] = cscale(s, -mjj);
        else if (lane == j) L[tri(j, j)] = cmake(mjj, 0.0);
        __syncwarp();
    }
}

// Entry (d,e), d >= e, of M^H M for packed lower-triangular M.
__device__ inline cd mhm_entry(const cd* M, int n, int d, int e) {
    cd s = cmake(0.0, 0.0), s2 = cmake(0.0, 0.0);
    int p = d;
    for (; p + 1 < n; p += 2) {
        cfma(s, cconj(M[tri(p, d)]), M[tri(p, e)]);
        cfma(s2, cconj(M[tri(p + 1, d)]), M[tri(p + 1, e)]);
    }
    if (p < n) cfma(s, cconj(M[tri(p, d)]), M[tri(p, e)]);
    return cadd(s, s2);
}

// ---------------------------------------------------------------------------
// Parallel-order two-sided Jacobi for a complex Hermitian n x n matrix.
// A: full matrix, row-major with leading dimension ld; destroyed (diagonal ->
// eigenvalues).  V: output eigenvectors in columns (same ld).  rot: scratch of
// 16 * 6 doubles.  red: scratch of 64 doubles.  All in shared memory.
// Must be called by all threads of the block.  Returns #sweeps used, or -1 if
// the sweep limit was hit.
// ---------------------------------------------------------------------------
struct JacobiRot { int p, q; double c, s, ex, ey; };

__device__ inline double block_sum(double v, double* red, int tid, int nthreads) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (nthreads + 31) / 32; ++w) r += red[w];
    return r;
}

__device__ inline int block_jacobi_eigh(cd* A, cd* V, int n, int ld, JacobiRot* rot,
                                        double* red, int tid, int nthreads,
                                        bool init_v = true, int max_sweeps = 40) {
    if (init_v)
        for (int i = tid; i < n * n; i += nthreads) {
            int r = i / n, c = i % n;
            V[r * ld + c] = cmake(r == c ? 1.0 : 0.0, 0.0);
        }
    const int m = (n + 1) & ~1;         // even number of players
    const int half = m / 2;
    __syncthreads();
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        // convergence test
        double off = 0.0, dia = 0.0;
        for (int i = tid; i < n * n; i += nthreads) {
            int r = i / n, c = i % n;
            double a2 = cabs2(A[r * ld + c]);
            if (r == c) dia += a2; else off += a2;
        }
        off = block_sum(off, red, tid, nthreads);
        dia = block_sum(dia, red, tid, nthreads);
        // converged when the off-diagonal norm is at the rounding level of the matrix
        // (n * eps relative): below that the sweeps only shuffle rounding noise
        const double tol = 4.0 * (double)(n * n) * 4.93e-32;
        if (off <= tol * dia || off == 0.0) break;
        const double thresh = 1e-36 * dia;
        for (int round = 0; round < m - 1; ++round) {
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = m - 1; q = round; }
                else { p = (round + tid) % (m - 1); q = (round - tid + (m - 1)) % (m - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                JacobiRot r; r.p = p; r.q = q; r.c = 1.0; r.s = 0.0; r.ex = 1.0; r.ey = 0.0;
                if (q < n) {
                    cd apq = A[p * ld + q];
                    double a2 = cabs2(apq);
                    if (a2 > thresh) {
                        double a = sqrt(a2);
                        double app = A[p * ld + p].x, aqq = A[q * ld + q].x;
                        double tau = (aqq - app) / (2.0 * a);
                        double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        r.c = 1.0 / sqrt(1.0 + t * t);
                        r.s = t * r.c;
                        r.ex = apq.x / a; r.ey = apq.y / a;
                    }
                } else {
                    r.p = -1;
                }
                rot[tid] = r;
            }
            __syncthreads();
            // column update on A and V:  X[:, (p,q)] <- X[:, (p,q)] * J
            for (int i = tid; i < 2 * half * n; i += nthreads) {
                int which = i / (half * n);
                int rem = i - which * half * n;
                int k = rem / n, row = rem - k * n;
                JacobiRot r = rot[k];
                if (r.p < 0 || r.s == 0.0) continue;
                cd* X = which ? V : A;
                cd xp = X[row * ld + r.p], xq = X[row * ld + r.q];
                cd se = cmake(r.s * r.ex, r.s * r.ey);
                // xp' = c xp - s conj(e) xq ;  xq' = s e xp + c xq
                cd np_ = cscale(xp, r.c); cfms(np_, cconj(se), xq);
                cd nq_ = cscale(xq, r.c); cfma(nq_, se, xp);
                X[row * ld + r.p] = np_; X[row * ld + r.q] = nq_;
            }
            __syncthreads();
            // row update on A:  A[(p,q), :] <- J^H A[(p,q), :]
            for (int i = tid; i < half * n; i += nthreads) {
                int k = i / n, col = i - k * n;
                JacobiRot r = rot[k];
                if (r.p < 0 || r.s == 0.0) continue;
                cd xp = A[r.p * ld + col], xq = A[r.q * ld + col];
                cd se = cmake(r.s * r.ex, r.s * r.ey);
                // xp' = c xp - s e xq ;  xq' = s conj(e) xp + c xq
                cd np_ = cscale(xp, r.c); cfms(np_, se, xq);
                cd nq_ = cscale(xq, r.c); cfma(nq_, cconj(se), xp);
                A[r.p * ld + col] = np_; A[r.q * ld + col] = nq_;
            }
            __syncthreads();
            // clean the annihilated entries and keep the diagonal real
            if (tid < half) {
                JacobiRot r = rot[tid];
                if (r.p >= 0 && r.s != 0.0) {
                    A[r.p * ld + r.q] = cmake(0.0, 0.0);
                    A[r.q * ld + r.p] = cmake(0.0, 0.0);
                    A[r.p * ld + r.p].y = 0.0;
                    A[r.q * ld + r.q].y = 0.0;
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    return sweep < max_sweeps ? sweep : -1;
}

// ---------------------------------------------------------------------------
// Block-level LU with partial pivoting:  solve A X = Bm for n x nrhs (zgesv
// semantics, used for stable_solve, math/solve.py:98).  A (n x n, ld) and
// Bm (n x nrhs, ldb) are overwritten; result in Bm.  Returns false if an exact
// zero pivot is met (LAPACK's "singular" condition).  piv: int scratch[1].
// ---------------------------------------------------------------------------
__device__ inline bool block_lu_solve(cd* A, int ld, cd* Bm, int ldb, int n, int nrhs,
                                      int* piv, int tid, int nthreads) {
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        if (tid == 0) {
            int best = j; double bv = fabs(A[j * ld + j].x) + fabs(A[j * ld + j].y);
            for (int i = j + 1; i < n; ++i) {     // izamax uses |re|+|im|
                double v = fabs(A[i * ld + j].x) + fabs(A[i * ld + j].y);
                if (v > bv) { bv = v; best = i; }
            }
            piv[0] = (bv == 0.0) ? -1 : best;
        }
        __syncthreads();
        int pv = piv[0];
        if (pv < 0) { ok = false; break; }
        if (pv != j) {
            for (int c = tid; c < n + nrhs; c += nthreads) {
                cd* pa = c < n ? &A[j * ld + c] : &Bm[j * ldb + (c - n)];
                cd* pb = c < n ? &A[pv * ld + c] : &Bm[pv * ldb + (c - n)];
                cd t = *pa; *pa = *pb; *pb = t;
            }
        }
        __syncthreads();
        cd pj = A[j * ld + j];
        double den = 1.0 / cabs2(pj);
        cd pinv = cmake(pj.x * den, -pj.y * den);
        // multipliers
        for (int i = j + 1 + tid; i < n; i += nthreads) A[i * ld + j] = cmul(A[i * ld + j], pinv);
        __syncthreads();
        // trailing update (A and right-hand sides)
        int rows = n - j - 1, cols = n - j - 1 + nrhs;
        for (int idx = tid; idx < rows * cols; idx += nthreads) {
            int i = j + 1 + idx / cols, cc = idx % cols;
            cd l = A[i * ld + j];
            if (cc < n - j - 1) { int c = j + 1 + cc; cd v = A[i * ld + c]; cfms(v, l, A[j * ld + c]); A[i * ld + c] = v; }
            else { int c = cc - (n - j - 1); cd v = Bm[i * ldb + c]; cfms(v, l, Bm[j * ldb + c]); Bm[i * ldb + c] = v; }
        }
        __syncthreads();
    }
    if (!ok) { __syncthreads(); return false; }
    // back substitution, one thread per right-hand side column
    for (int c = tid; c < nrhs; c += nthreads) {
        for (int i = n - 1; i >= 0; --i) {
            cd s = Bm[i * ldb + c];
            for (int p = i + 1; p < n; ++p) cfms(s, A[i * ld + p], Bm[p * ldb + c]);
            cd pj = A[i * ld + i];
            double den = 1.0 / cabs2(pj);
            Bm[i * ldb + c] = cmul(s, cmake(pj.x * den, -pj.y * den));
        }
    }
    __syncthreads();
    return true;
}

}  // namespace gss
