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
static __device__ __noinline__ void tridiag_expm_coeff(const double* scal, int numiter, double thresh, double dt_re,
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

// ---- small Krylov spaces (k_eff <= 16): exp(dt T) e_0 without an eigen-decomposition ----------------------------
// The QL recurrence above is a serial chain of square roots and divisions (14 us at k = 5, 26 us at k = 8, measured
// on B200) -- in the launch-latency regime that is a third to a half of a whole local problem.  krylov.py:122-136
// only needs coeff = |vec| exp(dt T) e_0, and for the small k of that regime the exponential of the k x k matrix is
// cheaper than its eigenvectors: shift by the mean diagonal (exp(dt mu) factored out), scale by 2^-s until
// |dt| |T - mu| <= 1/2, sixteen Taylor terms of the scaled matrix (tridiagonal times dense, every thread owns up to
// four matrix elements, the running sum in registers), s squarings.  All arithmetic is backward stable at this
// norm; the result equals the eigen-decomposition formula to a few 2^s eps (tests: 1e-12).  Returns false --
// uniformly over the CTA, before any barrier -- when the space is larger than 16, the norm needs more than ten
// squarings, the real part of dt times the norm exceeds 1.5 (see below) or the scalars are not finite; the caller
// then runs the QL path.
constexpr int TAYLOR_MAX_K = 16;
constexpr int TAYLOR_TERMS = 16;
constexpr int TAYLOR_MAX_SQUARINGS = 10;
constexpr double TAYLOR_MAX_REAL_NORM = 1.5;                      // |Re dt| |T - mu|_inf allowed on the Taylor path
constexpr int TAYLOR_SLOTS = 2;                                   // matrix elements per thread (blockDim.x >= 128)
constexpr int TAYLOR_SCRATCH_DOUBLES = 3 * TAYLOR_MAX_K * TAYLOR_MAX_K * 2;
static __constant__ double TAYLOR_INV[TAYLOR_TERMS + 1] = {0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7,
                                                         1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13,
                                                         1.0 / 14, 1.0 / 15, 1.0 / 16};

// Block-wide device function (every thread of the CTA must call it, blockDim.x >= 128); `scratch`: shared memory,
// TAYLOR_SCRATCH_DOUBLES doubles, 16-byte aligned.
static __device__ __noinline__ bool tridiag_expm_taylor(const double* scal, int numiter, double thresh, double dt_re,
                                                    double dt_im, double* coeff, int* keff_out, double* scratch) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const double nrm = scal[0];
    const double* alpha = scal + 1;
    const double* beta = alpha + numiter;
    int n = numiter;
    for (int j = 0; j < numiter - 1; j++)
        if (!(beta[j] >= thresh)) { n = j + 1; break; }
    if (n > TAYLOR_MAX_K) return false;
    double mu = 0.0;
    for (int i = 0; i < n; i++) mu += alpha[i];
    mu = mu / (double)n;
    double nb = 0.0;
    for (int i = 0; i < n; i++)
        nb = fmax(nb, fabs(alpha[i] - mu) + (i > 0 ? fabs(beta[i - 1]) : 0.0) + (i < n - 1 ? fabs(beta[i]) : 0.0));
    double na = nb * (fabs(dt_re) + fabs(dt_im));                 // >= |dt| |T - mu|_inf
    if (!(na < ldexp(0.5, TAYLOR_MAX_SQUARINGS)) || !(fabs(mu) < 1e300)) return false;   // too large, or NaN
    // A real part of dt makes exp(dt (T - mu)) non-unitary: its largest eigen-component (<= e^{|Re dt| nb}) can
    // dominate the matrix while e_0 barely overlaps with it, and the rounding errors of the squarings -- relative to
    // the MATRIX norm -- would then exceed those of the eigenvector formula, which are relative to the result.
    // Bounded amplification only: e^{2 * 1.5} = 20.
    if (!(nb * fabs(dt_re) <= TAYLOR_MAX_REAL_NORM)) return false;
    int s = 0;
    while (na > 0.5) { na *= 0.5; s++; }
    const double sc = ldexp(1.0, -s);
    const double sre = dt_re * sc, sim = dt_im * sc;

    double2* X = reinterpret_cast<double2*>(scratch);
    double2* P = X + TAYLOR_MAX_K * TAYLOR_MAX_K;
    double2* Q = P + TAYLOR_MAX_K * TAYLOR_MAX_K;
    const int nn = n * n;
    // this thread's elements e = tid + q NT = (i, j); row i of (T - mu): bl, bd, bu
    int eidx[TAYLOR_SLOTS], elo[TAYLOR_SLOTS], ehi[TAYLOR_SLOTS];
    double bl[TAYLOR_SLOTS], bd[TAYLOR_SLOTS], bu[TAYLOR_SLOTS];
    double2 xacc[TAYLOR_SLOTS];
#pragma unroll
    for (int q = 0; q < TAYLOR_SLOTS; q++) {
        const int e = tid + q * NT;
        eidx[q] = -1;
        bl[q] = bd[q] = bu[q] = 0.0;
        elo[q] = ehi[q] = 0;
        xacc[q] = make_double2(0.0, 0.0);
        if (e < nn) {
            const int i = e / n, j = e - i * n;
            eidx[q] = e;
            bd[q] = alpha[i] - mu;
            if (i > 0) { bl[q] = beta[i - 1]; elo[q] = e - n; } else elo[q] = e;
            if (i < n - 1) { bu[q] = beta[i]; ehi[q] = e + n; } else ehi[q] = e;
            xacc[q] = make_double2(i == j ? 1.0 : 0.0, 0.0);
            P[e] = xacc[q];
        }
    }
    __syncthreads();
    double2* told = P;
    double2* tnew = Q;
#pragma unroll 1
    for (int m = 1; m <= TAYLOR_TERMS; m++) {                  // (not unrolled: the code must stay in the instruction cache)
        const double fr = sre * TAYLOR_INV[m], fi = sim * TAYLOR_INV[m];
#pragma unroll
        for (int q = 0; q < TAYLOR_SLOTS; q++) {
            if (eidx[q] >= 0) {
                const double2 lo = told[elo[q]], mid = told[eidx[q]], hi = told[ehi[q]];
                const double rr = fma(bl[q], lo.x, fma(bd[q], mid.x, bu[q] * hi.x));
                const double ri = fma(bl[q], lo.y, fma(bd[q], mid.y, bu[q] * hi.y));
                const double2 t = make_double2(fr * rr - fi * ri, fr * ri + fi * rr);
                tnew[eidx[q]] = t;
                xacc[q].x += t.x; xacc[q].y += t.y;
            }
        }
        __syncthreads();
        double2* sw = told; told = tnew; tnew = sw;
    }
#pragma unroll
    for (int q = 0; q < TAYLOR_SLOTS; q++)
        if (eidx[q] >= 0) X[eidx[q]] = xacc[q];
    __syncthreads();
    // exp(A) = (exp(A / 2^s))^(2^s)
    double2* cur = X;
    double2* nxt = P;
    for (int t = 0; t < s; t++) {
#pragma unroll
        for (int q = 0; q < TAYLOR_SLOTS; q++) {
            if (eidx[q] >= 0) {
                const int i = eidx[q] / n, j = eidx[q] - i * n;
                double re0 = 0.0, re1 = 0.0, im0 = 0.0, im1 = 0.0;
                for (int k = 0; k < n; k++) {
                    const double2 a = cur[i * n + k], b = cur[k * n + j];
                    re0 = fma(a.x, b.x, re0); re1 = fma(a.y, b.y, re1);
                    im0 = fma(a.x, b.y, im0); im1 = fma(a.y, b.x, im1);
                }
                nxt[eidx[q]] = make_double2(re0 - re1, im0 + im1);
            }
        }
        __syncthreads();
        double2* sw = cur; cur = nxt; nxt = sw;
    }
    // coeff_r = |vec| exp(dt mu) exp(dt (T - mu))[r, 0]; zero beyond the breakdown
    if (tid < numiter) {
        double cre = 0.0, cim = 0.0;
        if (tid < n) {
            const double mag = nrm * exp(dt_re * mu);
            double sn, cs;
            sincos(dt_im * mu, &sn, &cs);
            const double2 x = cur[tid * n];
            cre = mag * (cs * x.x - sn * x.y);
            cim = mag * (cs * x.y + sn * x.x);
        }
        coeff[2 * tid] = cre;
        coeff[2 * tid + 1] = cim;
    }
    if (tid == 0) *keff_out = n;
    __syncthreads();                                           // coeff / *keff_out are visible to the whole CTA
    return true;
}

// The k x k problem of expm_krylov: Taylor path for small spaces, QL otherwise (block-wide, blockDim.x >= 128).
__device__ __forceinline__ void tridiag_expm_solve(const double* scal, int numiter, double thresh, double dt_re,
                                                   double dt_im, double* coeff, int* keff_out, double* scratch) {
    if (!tridiag_expm_taylor(scal, numiter, thresh, dt_re, dt_im, coeff, keff_out, scratch))
        tridiag_expm_coeff(scal, numiter, thresh, dt_re, dt_im, coeff, keff_out);
}


}  // namespace ptb
