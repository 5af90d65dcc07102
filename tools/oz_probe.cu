// Bring-up / timing probe for the experimental int8-slice SYRK (csrc/ozaki_syrk.cu).
//   oz_probe onehot <layout>          n=128, m=32, S=1: prints which (row, col) pairs the MMA pairs up
//   oz_probe ints <layout> n m        S=1, integer data: exact comparison
//   oz_probe full <layout> n m S      random fp64 data with row scaling and H: error vs long double
//   oz_probe perf <layout> n m S reps timing (CUDA events) + sampled entries vs long double
// A watchdog prints the kernel's progress words and exits if the launch does not finish in 10 s.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <random>
#include <chrono>
#include <thread>
#include <unistd.h>

namespace cvxb {
size_t ozaki_workspace_bytes(int n, int m, int S);
int ozaki_syrk(int n, int m, const double *A, long long lda, const double *d, const double *D, long long ldd,
               double beta, double *C, long long ldc, int S, int layout, void *work, unsigned int *dbg,
               cudaStream_t st);
}
extern "C" const char *cvxb_last_error(void);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static unsigned int *g_dbg = nullptr;
static int wait_stream(cudaStream_t st, double seconds) {
    auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess) return 0;
        if (q != cudaErrorNotReady) { printf("stream error: %s\n", cudaGetErrorString(q)); return 2; }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) {
            printf("HANG: dbg words:");
            for (int i = 0; i < 8; ++i) printf(" %x", g_dbg[i]);
            printf("\n");
            fflush(stdout);
            _exit(3);
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
}

int main(int argc, char **argv) {
    if (argc < 3) { printf("usage: oz_probe mode layout [n m S reps]\n"); return 1; }
    const char *mode = argv[1];
    const int layout = atoi(argv[2]);
    int n = argc > 3 ? atoi(argv[3]) : 128, m = argc > 4 ? atoi(argv[4]) : 32, S = argc > 5 ? atoi(argv[5]) : 1;
    const int reps = argc > 6 ? atoi(argv[6]) : 3;
    const bool onehot = !strcmp(mode, "onehot"), ints = !strcmp(mode, "ints"), perf = !strcmp(mode, "perf");
    if (onehot) { n = 128; m = 32; S = 1; }
    if (ints) S = 1;
    CK(cudaSetDevice(0));
    CK(cudaHostAlloc(&g_dbg, 512, cudaHostAllocMapped));
    memset(g_dbg, 0, 512);
    unsigned int *ddbg = nullptr;
    CK(cudaHostGetDevicePointer(&ddbg, g_dbg, 0));
    cudaStream_t st; CK(cudaStreamCreate(&st));

    std::mt19937_64 rng(7);
    std::vector<double> A((size_t)m * n), d(m, 1.0), H((size_t)n * n, 0.0);
    const bool plain = onehot || ints;
    if (onehot) {
        for (int j = 0; j < n; ++j) A[(size_t)j * m + (j % 32)] = 63.0 / 64.0;
    } else if (ints) {
        std::uniform_int_distribution<int> u(-63, 63);
        for (int j = 0; j < n; ++j) {
            for (int k = 0; k < m; ++k) A[(size_t)j * m + k] = u(rng) / 64.0;
            A[(size_t)j * m + (j % m)] = 63.0 / 64.0;          // pins the column exponent at 0
        }
    } else {
        std::normal_distribution<double> g(0.0, 1.0);
        for (auto &v : A) v = g(rng);
        for (auto &v : d) v = std::exp(3.0 * g(rng));           // scaling spread ~ e^+-9
        for (int j = 0; j < n; ++j) for (int i = j; i < n; ++i) H[i + (size_t)j * n] = (i == j) ? 2.0 : 0.01 * g(rng);
    }
    double *dA, *dd, *dH, *dC; void *work;
    CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dd, d.size() * 8));
    CK(cudaMalloc(&dH, H.size() * 8)); CK(cudaMalloc(&dC, H.size() * 8));
    CK(cudaMalloc(&work, cvxb::ozaki_workspace_bytes(n, m, S)));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dd, d.data(), d.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xFF, H.size() * 8));                     // NaN pattern: unwritten entries show up

    auto run = [&](unsigned int *dbg) {
        return cvxb::ozaki_syrk(n, m, dA, m, plain ? nullptr : dd, plain ? nullptr : dH, n, 1.0, dC, n, S, layout, work, dbg, st);
    };
    int rc = run(ddbg);
    if (rc) { printf("ozaki_syrk rc=%d: %s\n", rc, cvxb_last_error()); return 2; }
    if (int w = wait_stream(st, 10.0)) return w;
    printf("launch finished; dbg:");
    for (int i = 0; i < 4; ++i) printf(" %x", g_dbg[i]);
    if (getenv("CVXB_OZ_TRACE_CTA")) {
        // pass timeline of the traced CTA (us after its start): issue start, issue end, accumulators complete, epilogue end
        printf("\n  trace:");
        for (int g = 0; g < 5 && g_dbg[16 + 1 + 4 * g]; ++g) {
            printf("  pass %d:", g);
            for (int q = 1; q <= 4; ++q) printf(" %.1f", (double)(g_dbg[16 + q + 4 * g] - g_dbg[16]) * 1e-3);
        }
    }
    printf("\n");
    std::vector<double> C((size_t)n * n);
    CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));

    if (onehot) {
        // expected: C[i,j]*4096 = 3969 iff i%32 == j%32 (j <= i)
        int bad = 0;
        for (int i = 0; i < n; ++i) {
            bool rowbad = false;
            for (int j = 0; j <= i; ++j) {
                const double e = (i % 32 == j % 32) ? 3969.0 / 4096.0 : 0.0;
                if (C[i + (size_t)j * n] != e) rowbad = true;
            }
            if (rowbad) {
                ++bad;
                if (bad <= 24) {
                    printf("row %3d: nonzero cols:", i);
                    int cnt = 0;
                    for (int j = 0; j <= i && cnt < 10; ++j) {
                        const double v = C[i + (size_t)j * n];
                        if (v != 0.0) { printf(" %d(%g)", j, v * 4096.0); ++cnt; }
                    }
                    printf("\n");
                }
            }
        }
        printf("onehot layout %d: %d bad rows of %d -> %s\n", layout, bad, n, bad ? "FAIL" : "PASS");
        return bad ? 1 : 0;
    }
    // reference in long double on sampled (or all) entries
    const size_t total = (size_t)n * (n + 1) / 2;
    const bool sample = total * (size_t)m > 400000000ull;
    std::uniform_int_distribution<int> ui(0, n - 1);
    double maxerr = 0, maxrel = 0; size_t nbad = 0, checked = 0;
    auto check = [&](int i, int j) {
        long double acc = plain ? 0.0L : (long double)H[i + (size_t)j * n];
        long double mag = fabsl(acc);
        for (int k = 0; k < m; ++k) {
            const double a = d[k] * A[(size_t)i * m + k], b = d[k] * A[(size_t)j * m + k];
            acc += (long double)a * b; mag += fabsl((long double)a * b);
        }
        const double got = C[i + (size_t)j * n];
        const double err = fabs((double)((long double)got - acc));
        const double rel = err / (double)(mag > 0 ? mag : 1);
        if (!(err <= 0) && ints) ++nbad;
        if (!(rel < 1e-13)) { if (nbad < 10 && !ints) printf("  (%d,%d): got %.17g want %.17Lg\n", i, j, got, acc); if (!ints) ++nbad; }
        if (err > maxerr) maxerr = err;
        if (rel > maxrel || rel != rel) maxrel = rel;
        ++checked;
    };
    if (sample) for (int q = 0; q < 20000; ++q) { int i = ui(rng), j = ui(rng); if (j > i) std::swap(i, j); check(i, j); }
    else for (int j = 0; j < n; ++j) for (int i = j; i < n; ++i) check(i, j);
    printf("%s layout %d n=%d m=%d S=%d: checked %zu entries, max abs err %.3e, max err/sum|terms| %.3e, bad %zu -> %s\n",
           mode, layout, n, m, S, checked, maxerr, maxrel, nbad, nbad ? "FAIL" : "PASS");
    if (perf) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int r = 0; r < reps; ++r) {
            CK(cudaEventRecord(e0, st));
            rc = run(nullptr);
            CK(cudaEventRecord(e1, st));
            if (rc) { printf("rc=%d %s\n", rc, cvxb_last_error()); return 2; }
            if (int w = wait_stream(st, 20.0)) return w;
            float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("  rep %d: %.3f ms  (%.1f TF/s fp64-equivalent, n^2 m flops)\n", r, ms, (double)n * n * m / ms * 1e-9);
        }
    }
    return nbad ? 1 : 0;
}
