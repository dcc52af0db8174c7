// k x k tridiagonal problem of expm_krylov on the device (shared by the stand-alone kernel of krylov.cu and the
// one-kernel local step of lanczos_small.cu).
// krylov.py:122-136: coeff = U (|vec| exp(dt w) * U[0, :]) with (w, U) the eigen-decomposition of the Lanczos
// tridiagonal matrix (krylov.py:142-150): implicit symmetric QL iteration (the classic tql2 recurrence).  Every
// worker thread runs the scalar recurrence redundantly (identical values, no communication) and applies the plane
// rotations to ITS row of the eigenvector matrix; thread r ends up holding U[r, :].  The breakdown rule of
// krylov.py:44-50 is applied to the betas first: k_eff = first j with beta[j] < thresh, plus one.
#pragma once
#include "common.cuh"

namespace ptb {

constexpr int TRIDIAG_MAX = 64;

// sqrt(a^2 + b^2) on the dependent chain of the QL recurrence: one FMA + one square root instead of the ~150
// instruction library hypot; the scaled form only when the squares leave the double range.
__device__ __forceinline__ double fast_hypot(double a, double b) {
    const double r = sqrt(fma(a, a, b * b));
    if (r > 1e-140 && r < 1e140) return r;
    return hypot(a, b);
}

// Block-wide device function: every thread of the CTA must call it (blockDim.x >= TRIDIAG_MAX, it contains CTA
// barriers); threads 0 .. TRIDIAG_MAX-1 do the work.  `scal` = [|vec|, alpha[0:numiter], beta[0:numiter-1]].
__device__ __forceinline__ void tridiag_expm_coeff(const double* scal, int numiter, double thresh, double dt_re,
                                                   double dt_im, double* coeff, int* keff_out) {
    // working arrays in shared memory (dynamically indexed: as thread-local arrays they would live in local memory
    // on the critical path of a serial recurrence): d, e once per warp (all lanes hold identical values, a store
    // of one value by 32 lanes is a single transaction), the eigenvector rows z[i][thread] conflict free
    __shared__ double d_s[TRIDIAG_MAX / 32][TRIDIAG_MAX], e_s[TRIDIAG_MAX / 32][TRIDIAG_MAX];
    __shared__ double z_s[TRIDIAG_MAX * TRIDIAG_MAX];
    __shared__ double u0[TRIDIAG_MAX], fre[TRIDIAG_MAX], fim[TRIDIAG_MAX];      // U[0, j]; |vec| exp(dt w_j) U[0, j]
    const int r = threadIdx.x;
    const bool worker = r < TRIDIAG_MAX;
    double* d = d_s[worker ? (r >> 5) : 0];
    double* e = e_s[worker ? (r >> 5) : 0];
    double* z = z_s + (worker ? r : 0);                        // z[i * TRIDIAG_MAX]
    const double nrm = scal[0];
    const double* alpha = scal + 1;
    const double* beta = alpha + numiter;
    int n = numiter;
    for (int j = 0; j < numiter - 1; j++)
        if (!(beta[j] >= thresh)) { n = j + 1; break; }      // also catches NaN of a speculative step
    if (worker) {
    for (int i = 0; i < n; i++) {
        d[i] = alpha[i]; e[i] = (i + 1 < n) ? beta[i] : 0.0; z[i * TRIDIAG_MAX] = (i == r) ? 1.0 : 0.0;
    }
    __syncwarp();
    // e[i] couples i and i+1 (already in tql2's shifted convention), e[n-1] = 0
    double f = 0.0, tst1 = 0.0;
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; l++) {
        tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
        int m = l;
        while (m < n - 1 && fabs(e[m]) > eps * tst1) m++;
        if (m > l) {
            int iter = 0;
            double el;
            do {
                iter++;
                double g = d[l];
                el = e[l];
                double p = (d[l + 1] - g) / (2.0 * el);
                double rr = fast_hypot(p, 1.0);
                if (p < 0) rr = -rr;
                const double dl = el / (p + rr);
                const double dl1 = el * (p + rr);
                double h = g - dl;
                __syncwarp();
                d[l] = dl;
                d[l + 1] = dl1;
                for (int i = l + 2; i < n; i++) d[i] -= h;
                __syncwarp();
                f += h;
                p = d[m];
                double c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
                const double el1 = e[l + 1];
                double zi1 = z[m * TRIDIAG_MAX];              // this thread's z[i + 1], carried in a register
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2; c2 = c; s2 = s;
                    const double ei = e[i], di = d[i];
                    g = c * ei;
                    h = c * p;
                    rr = fast_hypot(p, ei);
                    const double inv = 1.0 / rr;
                    const double enew = s * rr;
                    s = ei * inv;
                    c = p * inv;
                    p = c * di - s * g;
                    const double dnew = h + s * (c * g + s * di);
                    __syncwarp();
                    e[i + 1] = enew;
                    d[i + 1] = dnew;
                    const double zi = z[i * TRIDIAG_MAX];     // rotate columns i, i+1 of this thread's row
                    z[(i + 1) * TRIDIAG_MAX] = s * zi + c * zi1;
                    zi1 = c * zi - s * zi1;
                }
                z[l * TRIDIAG_MAX] = zi1;
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                __syncwarp();
                el = s * p;
                e[l] = el;
                d[l] = c * p;
                __syncwarp();
            } while (fabs(el) > eps * tst1 && iter < 60);
        }
        __syncwarp();
        d[l] = d[l] + f;
        e[l] = 0.0;
        __syncwarp();
    }
    }   // worker
    if (r == 0) {
        *keff_out = n;
        for (int j = 0; j < n; j++) u0[j] = z[j * TRIDIAG_MAX];          // thread 0 holds row 0 of U
    }
    __syncthreads();
    if (r < n) {
        // |vec| exp(dt w_j) U[0, j], one j per thread
        const double mag = nrm * exp(dt_re * d[r]) * u0[r];
        double sn, cs;
        sincos(dt_im * d[r], &sn, &cs);
        fre[r] = mag * cs;
        fim[r] = mag * sn;
    }
    __syncthreads();
    if (r < numiter && worker) {
        double cre = 0.0, cim = 0.0;
        if (r < n) {
            for (int j = 0; j < n; j++) {
                const double zj = z[j * TRIDIAG_MAX];
                cre += zj * fre[j];
                cim += zj * fim[j];
            }
        }
        coeff[2 * r] = cre;
        coeff[2 * r + 1] = cim;
    }
    __syncthreads();                                           // coeff / *keff_out are visible to the whole CTA
}


}  // namespace ptb
