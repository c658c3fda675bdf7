/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the stencil sweep.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under drstencil_b200/ links, imports or calls it.
 *
 * It restates, in plain C, what the reference's *gold* kernels compute and how the emitted
 * host program drives them (the reference has no CPU loop of its own):
 *   sweep      /root/reference/drstencil_2d.hpp:164-178 + codegen_2d.hpp:666-688   (2D, K4)
 *              /root/reference/drstencil.hpp:182-196   + codegen.hpp:637-660      (3D, K5)
 *   schedule   /root/reference/codegen_2d.hpp:604-619,638-649 / codegen.hpp:575-589,609-620
 *   inputs     /root/reference/common.hpp:9-45
 *   metric     /root/reference/common.hpp:47-102
 *
 * Parity status: pinned.  tests/test_oracle_pinning.py checks the expression order, the
 * coefficient literals and the Halo/Dist/Range macros against text emitted by the reference
 * generator itself (oracle/_ref/drstencil_ref, built from /root/reference/main.cpp), and
 * tests/test_ref_gold_gpu.py checks the arithmetic bit-for-bit against the reference's
 * emitted gold_<name> kernel compiled with nvcc for sm_100a (oracle/_ref/libref_gold_*.so).
 *
 * Arithmetic rule (observed in the SASS nvcc 12.9 produces for the gold expression with the
 * reference's flags, default -fmad=true):  t1 + t2 + ... + tP  with t_q = c_q * in[p+q]
 * contracts to
 *      acc = c2*a2            (rounded product)
 *      acc = fma(c1, a1, acc)
 *      acc = fma(c3, a3, acc) ... acc = fma(cP, aP, acc)
 * and P == 1 is a single rounded product.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int drs_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's CPU legs: use n threads from now on, whatever OMP_NUM_THREADS said when the library was loaded
 * (torch.distributed.run exports OMP_NUM_THREADS=1 to every rank) */
void drs_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* common.hpp:9-11 -- rand()/(RAND_MAX-1), one draw per element, row-major.  `reseed` != 0
 * puts glibc's generator back in its never-seeded state (srand(1)). */
void drs_oracle_fill_rand(double *a, size_t n, int reseed) {
    if (reseed) srand(1);
    for (size_t x = 0; x < n; ++x) a[x] = (double)rand() / (double)(RAND_MAX - 1);
}

void drs_oracle_fill_rand_f32(float *a, size_t n, int reseed) {
    if (reseed) srand(1);
    for (size_t x = 0; x < n; ++x) a[x] = (float)((double)rand() / (double)(RAND_MAX - 1));
}

/* parallel first-touch fill for the throughput baseline (values do not matter there) */
void drs_oracle_fill_lcg(double *a, size_t n, unsigned seed) {
#pragma omp parallel for schedule(static)
    for (long long x = 0; x < (long long)n; ++x) {
        unsigned long long s = (unsigned long long)x * 6364136223846793005ULL + seed * 1442695040888963407ULL + 1;
        s ^= s >> 33; s *= 0xff51afd7ed558ccdULL; s ^= s >> 33;
        a[x] = (double)(s >> 11) * (1.0 / 9007199254740992.0);
    }
}
void drs_oracle_fill_lcg_f32(float *a, size_t n, unsigned seed) {
#pragma omp parallel for schedule(static)
    for (long long x = 0; x < (long long)n; ++x) {
        unsigned long long s = (unsigned long long)x * 6364136223846793005ULL + seed * 1442695040888963407ULL + 1;
        s ^= s >> 33; s *= 0xff51afd7ed558ccdULL; s ^= s >> 33;
        a[x] = (float)((double)(s >> 11) * (1.0 / 9007199254740992.0));
    }
}

/* The inner loop is blocked over i so that each term is a unit-stride pass the compiler can
 * vectorise; every output element still sees exactly the chain described above. */
#define DRS_IB 512
#define SWEEP_BODY(T, FMA)                                                                      \
    const long long sj = N, sk = M * N;                                                         \
    long long *lin = (long long *)malloc(sizeof(long long) * (size_t)P);                        \
    for (int q = 0; q < P; ++q) lin[q] = offs[3 * q] * sk + offs[3 * q + 1] * sj + offs[3 * q + 2]; \
    const long long k0 = dim == 3 ? H : 0, k1 = dim == 3 ? L - H : 1;                           \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                    \
    for (long long k = k0; k < k1; ++k)                                                         \
        for (long long j = H; j < M - H; ++j) {                                                 \
            const T *ip = in + k * sk + j * sj;                                                 \
            T *op = out + k * sk + j * sj;                                                      \
            T acc[DRS_IB];                                                                      \
            for (long long ib = H; ib < N - H; ib += DRS_IB) {                                  \
                const int n = (int)((N - H - ib) < DRS_IB ? (N - H - ib) : DRS_IB);             \
                if (P == 1) {                                                                   \
                    const T c = (T)coefs[0]; const T *a = ip + ib + lin[0];                     \
                    for (int x = 0; x < n; ++x) acc[x] = c * a[x];                              \
                } else if (contract) {                                                          \
                    { const T c = (T)coefs[1]; const T *a = ip + ib + lin[1];                   \
                      for (int x = 0; x < n; ++x) acc[x] = c * a[x]; }                          \
                    { const T c = (T)coefs[0]; const T *a = ip + ib + lin[0];                   \
                      for (int x = 0; x < n; ++x) acc[x] = FMA(c, a[x], acc[x]); }              \
                    for (int q = 2; q < P; ++q) {                                               \
                        const T c = (T)coefs[q]; const T *a = ip + ib + lin[q];                 \
                        for (int x = 0; x < n; ++x) acc[x] = FMA(c, a[x], acc[x]);              \
                    }                                                                           \
                } else {                                                                        \
                    { const T c = (T)coefs[0]; const T *a = ip + ib + lin[0];                   \
                      for (int x = 0; x < n; ++x) acc[x] = c * a[x]; }                          \
                    for (int q = 1; q < P; ++q) {                                               \
                        const T c = (T)coefs[q]; const T *a = ip + ib + lin[q];                 \
                        for (int x = 0; x < n; ++x) { T prod = c * a[x]; acc[x] = acc[x] + prod; } \
                    }                                                                           \
                }                                                                               \
                for (int x = 0; x < n; ++x) op[ib + x] = acc[x];                                \
            }                                                                                   \
        }                                                                                       \
    free(lin);

/* One gold sweep: out[p] = sum_q c_q * in[p+q] on [H, dim-H) in every axis; the H-wide ring
 * of `out` is left untouched.  offs = P triples (dk, dj, di) in evaluation order (ascending
 * std::map order); coefs are the doubles the printed literals denote.
 * contract = 1: nvcc's fused chain (header comment); 0: every product and sum rounded. */
void drs_oracle_sweep_f64(int dim, long long L, long long M, long long N, int H, int P,
                          const int *offs, const double *coefs, const double *in, double *out,
                          int contract) {
    SWEEP_BODY(double, fma)
}

/* fp32 has no reference counterpart (the reference is fp64 only); defined by extension:
 * same order and chain in single precision, coefficient = (float) of the same literal. */
void drs_oracle_sweep_f32(int dim, long long L, long long M, long long N, int H, int P,
                          const int *offs, const double *coefs, const float *in, float *out,
                          int contract) {
    SWEEP_BODY(float, fmaf)
}

/* Ping-pong schedule of the emitted host program: for (t = 0; t < iterations; t += 2*step)
 * { A->B; B->A; }  i.e. 2*ceil(iterations/(2*step)) sweeps, result in A.  Returns the sweep count. */
int drs_oracle_run_f64(int dim, long long L, long long M, long long N, int H, int P, const int *offs,
                       const double *coefs, double *A, double *B, int iterations, int step, int contract) {
    int sweeps = 0;
    for (int t = 0; t < iterations; t += 2 * step) {
        drs_oracle_sweep_f64(dim, L, M, N, H, P, offs, coefs, A, B, contract);
        drs_oracle_sweep_f64(dim, L, M, N, H, P, offs, coefs, B, A, contract);
        sweeps += 2;
    }
    return sweeps;
}
int drs_oracle_run_f32(int dim, long long L, long long M, long long N, int H, int P, const int *offs,
                       const double *coefs, float *A, float *B, int iterations, int step, int contract) {
    int sweeps = 0;
    for (int t = 0; t < iterations; t += 2 * step) {
        drs_oracle_sweep_f32(dim, L, M, N, H, P, offs, coefs, A, B, contract);
        drs_oracle_sweep_f32(dim, L, M, N, H, P, offs, coefs, B, A, contract);
        sweeps += 2;
    }
    return sweeps;
}

/* common.hpp:47-102 -- max |a-b| (floor 1e-13, as the reference initialises it) and RMS over
 * the interior box [lb, ub) per axis.  res[0] = max abs error, res[1] = rms,
 * res[2..4] = (k, j, i) of the max. */
void drs_oracle_check_error(int dim, long long L, long long M, long long N, int H, const double *outp,
                            const double *ref, double *res) {
    const long long k0 = dim == 3 ? H : 0, k1 = dim == 3 ? L - H : 1;
    double err = 0.0, mx = 1e-13;
    long long mk = 0, mj = 0, mi = 0;
    for (long long k = k0; k < k1; ++k)
        for (long long j = H; j < M - H; ++j)
            for (long long i = H; i < N - H; ++i) {
                double d = outp[(k * M + j) * N + i] - ref[(k * M + j) * N + i];
                d = d < 0.0 ? -d : d;
                err += d * d;
                if (d > mx) { mx = d; mk = k; mj = j; mi = i; }
            }
    double cnt = (double)(k1 - k0) * (double)(M - 2 * H) * (double)(N - 2 * H);
    res[0] = mx; res[1] = sqrt(err / cnt); res[2] = (double)mk; res[3] = (double)mj; res[4] = (double)mi;
}
