// Dense fp64 Cholesky (lower) + triangular solves for sm_100a.
//
// Replaces lapack.potrf / lapack.potrs as called by the reference KKT solver
// (reference src/python/misc.py:1282 and :1327; bindings src/C/lapack.c:1471, :1553).
//
// potrf_lower: right-looking blocked factorisation with NB = 128.
//   step j:  (1) potf2_inv   : one CTA factors the 128x128 diagonal block in shared
//                              memory and also produces its inverse (kept for the solves)
//            (2) panel TRSM  : L21 = A21 * inv(L11)'  as a DMMA GEMM (gemm_dmma.cu)
//            (3) trailing    : A22 -= L21 L21'  (lower tiles) as a DMMA GEMM
//   Look-ahead: the first tile column of (3) — the next panel — runs on the
//   high-priority panel stream together with (1),(2); the rest of (3) runs on the
//   update stream, so the latency-bound panel work of step j+1 hides behind the
//   throughput-bound update of step j.
//
// trsv_lower: blocked substitution that re-uses the diagonal-block inverses; one CTA
//   per 128-row block, progress published through acquire/release flags so that the
//   whole triangular solve is a single kernel (no per-block launches).
#include "common.cuh"

namespace cvxb {

namespace {

constexpr int LDS = NB + 1;            // shared-memory leading dimension of the block
constexpr int POTF2_SMEM = (NB * LDS + 64 * 64 + NB) * 8;

// Factor the jb x jb diagonal block at A (lower) in place, write inv(L) (NB x NB,
// ld NB, zero upper, identity padding beyond jb) to inv.
__global__ void __launch_bounds__(256, 1)
potf2_inv_kernel(double *A, long long lda, int jb, double *inv, int *info, int joff) {
    extern __shared__ __align__(16) double sm[];
    double *As = sm;                   // NB x LDS, column-major
    double *T = As + NB * LDS;         // 64*64 temp for the inversion
    double *colj = T + 64 * 64;        // NB
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int e = tid; e < NB * NB; e += 256) {
        int i = e & (NB - 1), k = e >> 7;
        double v = (i == k) ? 1.0 : 0.0;
        if (i < jb && k < jb && i >= k) v = A[i + (long long)k * lda];
        As[i + k * LDS] = v;
    }
    __syncthreads();

    // ---- phase 1: unblocked right-looking Cholesky in shared memory ----
    for (int j = 0; j < jb; ++j) {
        const double ajj = As[j + j * LDS];
        if (!(ajj > 0.0)) {            // also catches NaN; uniform across the CTA
            if (tid == 0) atomicCAS(info, 0, joff + j + 1);
            break;
        }
        const double s = sqrt(ajj);
        const double rs = 1.0 / s;
        for (int i = j + tid; i < jb; i += 256) {
            double v = (i == j) ? s : As[i + j * LDS] * rs;
            colj[i] = v;
            As[i + j * LDS] = v;
        }
        __syncthreads();
        for (int k = j + 1 + warp; k < jb; k += 8) {
            const double lkj = colj[k];
            for (int i = k + lane; i < jb; i += 32) As[i + k * LDS] -= colj[i] * lkj;
        }
        __syncthreads();
    }
    __syncthreads();
    // write L back (lower part only)
    for (int e = tid; e < NB * NB; e += 256) {
        int i = e & (NB - 1), k = e >> 7;
        if (i < jb && k < jb && i >= k) A[i + (long long)k * lda] = As[i + k * LDS];
    }
    __syncthreads();

    // ---- phase 2: in-place inversion by recursive doubling ----
    // level s: for every aligned pair of s x s diagonal blocks (already inverted),
    //   X21 = - X22 * (L21 * X11)
    for (int i = tid; i < NB; i += 256) As[i + i * LDS] = 1.0 / As[i + i * LDS];
    __syncthreads();
    for (int s = 1; s < NB; s <<= 1) {
        const int npairs = NB / (2 * s);
        const int per = s * s;
        // T = L21 * X11   (X11 lower triangular)
        for (int e = tid; e < npairs * per; e += 256) {
            int pr = e / per, loc = e - pr * per;
            int i = loc % s, jj = loc / s;
            int o = pr * 2 * s;
            double acc = 0.0;
            for (int k = jj; k < s; ++k)
                acc += As[(o + s + i) + (o + k) * LDS] * As[(o + k) + (o + jj) * LDS];
            T[pr * per + loc] = acc;
        }
        __syncthreads();
        // X21 = - X22 * T   (X22 lower triangular)
        for (int e = tid; e < npairs * per; e += 256) {
            int pr = e / per, loc = e - pr * per;
            int i = loc % s, jj = loc / s;
            int o = pr * 2 * s;
            double acc = 0.0;
            for (int k = 0; k <= i; ++k)
                acc += As[(o + s + i) + (o + s + k) * LDS] * T[pr * per + k + jj * s];
            As[(o + s + i) + (o + jj) * LDS] = -acc;
        }
        __syncthreads();
    }
    for (int e = tid; e < NB * NB; e += 256) {
        int i = e & (NB - 1), k = e >> 7;
        inv[e] = (i >= k) ? As[i + k * LDS] : 0.0;
    }
}

// ---- blocked triangular solve with one right-hand side ------------------------
// forward:  L x = b ;  backward: L' x = b.   In place on b.  One CTA per row block.
// flags[i] == epoch  <=>  x_i is final in b.
template <bool TRANS>
__global__ void __launch_bounds__(256, 1)
trsv_kernel(int n, const double *__restrict__ L, long long ldl, const double *__restrict__ inv,
            double *b, int *flags, int epoch) {
    __shared__ double xs[NB];
    __shared__ double part[2][NB];
    __shared__ double tvec[NB];
    const int nblk = (n + NB - 1) / NB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bi = TRANS ? (nblk - 1 - (int)blockIdx.x) : (int)blockIdx.x;
    const int i0 = bi * NB;
    const int ni = min(NB, n - i0);
    const int r = tid & (NB - 1), half = tid >> 7;

    double acc = 0.0;
    if (!TRANS) {
        for (int j = 0; j < bi; ++j) {
            if (tid == 0) while (ld_acquire(flags + j) != epoch) { }
            __syncthreads();
            if (tid < NB) xs[tid] = __ldcg(b + j * NB + tid);
            __syncthreads();
            if (r < ni) {
                const double *Lp = L + (i0 + r) + (long long)(j * NB + half * 64) * ldl;
#pragma unroll 8
                for (int c = 0; c < 64; ++c) acc += Lp[(long long)c * ldl] * xs[half * 64 + c];
            }
        }
        part[half][r] = acc;
        __syncthreads();
        if (tid < NB) tvec[tid] = (tid < ni) ? (__ldcg(b + i0 + tid) - part[0][tid] - part[1][tid]) : 0.0;
        __syncthreads();
        // x_i = inv_ii * t   (inv lower: columns c <= r)
        double a2 = 0.0;
        {
            const double *ip = inv + (long long)bi * NB * NB + r + (long long)(half * 64) * NB;
            for (int c = 0; c < 64; ++c) {
                int cc = half * 64 + c;
                if (cc <= r) a2 += ip[c * NB] * tvec[cc];
            }
        }
        __syncthreads();
        part[half][r] = a2;
        __syncthreads();
        if (tid < ni) b[i0 + tid] = part[0][tid] + part[1][tid];
    } else {
        // accumulate t[c] = sum_{j>bi} sum_r L[j*NB + r, i0 + c] * x_j[r]; warp w owns
        // columns c = w*16 .. w*16+15
        double accc[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) accc[q] = 0.0;
        for (int j = nblk - 1; j > bi; --j) {
            if (tid == 0) while (ld_acquire(flags + j) != epoch) { }
            __syncthreads();
            const int nj = min(NB, n - j * NB);
            if (tid < NB) xs[tid] = (tid < nj) ? __ldcg(b + j * NB + tid) : 0.0;
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                int c = warp * 16 + q;
                if (c < ni) {
                    const double *Lp = L + (long long)(j * NB) + (long long)(i0 + c) * ldl;
                    double a = 0.0;
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        int row = lane + rr * 32;
                        if (row < nj) a += Lp[row] * xs[row];
                    }
                    accc[q] += a;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            double a = warp_sum(accc[q]);
            int c = warp * 16 + q;
            if (lane == 0) tvec[c] = (c < ni) ? (__ldcg(b + i0 + c) - a) : 0.0;
        }
        __syncthreads();
        // x_i[c] = sum_{r >= c} inv[r, c] * t[r]
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            int c = warp * 16 + q;
            const double *ip = inv + (long long)bi * NB * NB + (long long)c * NB;
            double a = 0.0;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                int row = lane + rr * 32;
                if (row >= c) a += ip[row] * tvec[row];
            }
            a = warp_sum(a);
            if (lane == 0 && c < ni) b[i0 + c] = a;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(flags + bi, epoch);
}

int g_trsv_epoch = 0;

}  // namespace

int chol_work_create(CholWork &w) {
    int least = 0, greatest = 0;
    CVXB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CVXB_CUDA(cudaStreamCreateWithPriority(&w.panel_stream, cudaStreamNonBlocking, greatest));
    CVXB_CUDA(cudaStreamCreateWithPriority(&w.update_stream, cudaStreamNonBlocking, least));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_start, cudaEventDisableTiming));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_panel, cudaEventDisableTiming));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_rest, cudaEventDisableTiming));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_end_p, cudaEventDisableTiming));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_end_u, cudaEventDisableTiming));
    CVXB_CUDA(cudaMalloc(&w.d_info, sizeof(int)));
    CVXB_CUDA(cudaMemset(w.d_info, 0, sizeof(int)));
    CVXB_CUDA(cudaMalloc(&w.d_flags, 4096 * sizeof(int)));
    CVXB_CUDA(cudaMemset(w.d_flags, 0, 4096 * sizeof(int)));
    CVXB_CUDA(cudaMalloc(&w.splitk_ws, (size_t)kNumSMs * NB * NB * sizeof(double)));
    CVXB_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   POTF2_SMEM));
    return 0;
}

void chol_work_destroy(CholWork &w) {
    if (w.panel_stream) cudaStreamDestroy(w.panel_stream);
    if (w.update_stream) cudaStreamDestroy(w.update_stream);
    if (w.ev_start) cudaEventDestroy(w.ev_start);
    if (w.ev_panel) cudaEventDestroy(w.ev_panel);
    if (w.ev_rest) cudaEventDestroy(w.ev_rest);
    if (w.ev_end_p) cudaEventDestroy(w.ev_end_p);
    if (w.ev_end_u) cudaEventDestroy(w.ev_end_u);
    if (w.d_info) cudaFree(w.d_info);
    if (w.d_flags) cudaFree(w.d_flags);
    if (w.splitk_ws) cudaFree(w.splitk_ws);
    w = CholWork();
}

int potrf_lower(int n, double *A, int lda, double *inv, CholWork &w, cudaStream_t st) {
    if (n <= 0) return 0;
    const int nblk = (n + NB - 1) / NB;
    cudaStream_t P = w.panel_stream, U = w.update_stream;
    CVXB_CUDA(cudaMemsetAsync(w.d_info, 0, sizeof(int), st));
    CVXB_CUDA(cudaEventRecord(w.ev_start, st));
    CVXB_CUDA(cudaStreamWaitEvent(P, w.ev_start, 0));
    CVXB_CUDA(cudaStreamWaitEvent(U, w.ev_start, 0));
    bool have_rest = false;
    for (int jb = 0; jb < nblk; ++jb) {
        const int j = jb * NB;
        const int wj = (n - j < NB) ? (n - j) : NB;
        const int m = n - j - wj;
        double *Ajj = A + j + (long long)j * lda;
        double *invj = inv + (long long)jb * NB * NB;
        potf2_inv_kernel<<<1, 256, POTF2_SMEM, P>>>(Ajj, lda, wj, invj, w.d_info, j);
        count_launch();
        CVXB_LAUNCH_CHECK();
        if (m <= 0) break;
        double *A21 = Ajj + wj;
        double *A22 = A21 + (long long)wj * lda;
        {   // panel TRSM as GEMM with the block inverse (in place: one CTA owns its rows)
            GemmDesc g;
            g.M = m; g.N = wj; g.K = wj;
            g.X = A21; g.ldx = lda; g.x_kmajor = false;
            g.Y = invj; g.ldy = NB; g.y_kmajor = false;
            g.C = A21; g.ldc = lda;
            CVXB_TRY(dmma_gemm(g, P));
        }
        CVXB_CUDA(cudaEventRecord(w.ev_panel, P));
        GemmDesc u;
        u.M = m; u.N = m; u.K = wj;
        u.X = A21; u.ldx = lda; u.x_kmajor = false;
        u.Y = A21; u.ldy = lda; u.y_kmajor = false;
        u.D = A22; u.ldd = lda; u.C = A22; u.ldc = lda;
        u.alpha = -1.0; u.beta = 1.0; u.lower_only = true;
        // tile column 0 (the next panel) on the panel stream
        if (have_rest) CVXB_CUDA(cudaStreamWaitEvent(P, w.ev_rest, 0));
        u.ct_begin = 0; u.ct_end = 1;
        CVXB_TRY(dmma_gemm(u, P));
        if (m > NB) {
            CVXB_CUDA(cudaStreamWaitEvent(U, w.ev_panel, 0));
            u.ct_begin = 1; u.ct_end = 1 << 30;
            CVXB_TRY(dmma_gemm(u, U));
            CVXB_CUDA(cudaEventRecord(w.ev_rest, U));
            have_rest = true;
        }
    }
    CVXB_CUDA(cudaEventRecord(w.ev_end_p, P));
    CVXB_CUDA(cudaEventRecord(w.ev_end_u, U));
    CVXB_CUDA(cudaStreamWaitEvent(st, w.ev_end_p, 0));
    CVXB_CUDA(cudaStreamWaitEvent(st, w.ev_end_u, 0));
    return 0;
}

int trsv_lower(int n, const double *L, int ldl, const double *inv, double *b, bool trans,
               CholWork &w, cudaStream_t st) {
    if (n <= 0) return 0;
    const int nblk = (n + NB - 1) / NB;
    if (nblk > 4096) {
        set_error("trsv_lower: n too large");
        return CVXB_E_ARG;
    }
    const int epoch = ++g_trsv_epoch;
    if (trans) trsv_kernel<true><<<nblk, 256, 0, st>>>(n, L, ldl, inv, b, w.d_flags, epoch);
    else       trsv_kernel<false><<<nblk, 256, 0, st>>>(n, L, ldl, inv, b, w.d_flags, epoch);
    count_launch();
    CVXB_LAUNCH_CHECK();
    return 0;
}

int potrs_lower(int n, const double *L, int ldl, const double *inv, double *b, CholWork &w,
                cudaStream_t st) {
    CVXB_TRY(trsv_lower(n, L, ldl, inv, b, false, w, st));
    CVXB_TRY(trsv_lower(n, L, ldl, inv, b, true, w, st));
    return 0;
}

}  // namespace cvxb
