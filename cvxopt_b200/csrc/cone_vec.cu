// The O(cdim) cone algebra of the IPM side, mirror of src/C/misc_solvers.c:
//   scale2 (:256-401), sprod (:634-767), sinv (:775-878), trisc (:887-935), triusc (:940-986),
//   sdot (:991-1039), max_step (:1052-1153).
// One CTA walks the whole cone vector; reductions inside a 'q' cone / for sdot are block-wide.
// The 's' part of max_step (reference: dsyevr_ for the smallest eigenvalue, dsyevd_ when sigma is
// given, :1099-1150) is a parallel-order cyclic Jacobi eigensolver: every round applies N/2 disjoint
// plane rotations at once, one thread per 2x2 block of J'AJ, ping-ponging between two copies so a
// round is a single race-free pass.
#include "cone.cuh"
#include <cooperative_groups.h>
#include <mutex>
#include <map>
#include <vector>
#include <algorithm>
#include <cfloat>

using namespace cvxb;

namespace {

__device__ __forceinline__ double blk_sum(double v, double *sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) t = warp_sum(t);
    if (threadIdx.x == 0) sh[0] = t;
    __syncthreads();
    return sh[0];
}

struct Cones {
    int nl;                     // mnl + ml
    int nq; const int *q;       // device arrays
    int ns; const int *s;
};

// x := H(lambda^{1/2}) x  or its inverse
__global__ void scale2_kernel(const double *lm, double *x, Cones c, int inverse) {
    __shared__ double sh[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < c.nl; i += nt) x[i] = inverse ? x[i] * lm[i] : x[i] / lm[i];
    int m = c.nl;
    for (int k = 0; k < c.nq; ++k) {
        const int mk = c.q[k];
        double n2 = 0, dot = 0;
        for (int i = 1 + tid; i < mk; i += nt) { n2 += lm[m + i] * lm[m + i]; dot += lm[m + i] * x[m + i]; }
        n2 = blk_sum(n2, sh); dot = blk_sum(dot, sh);
        const double nrm = sqrt(n2), l0 = lm[m], x0 = x[m];
        double a = sqrt(l0 + nrm) * sqrt(l0 - nrm);
        const double lx = inverse ? (l0 * x0 + dot) / a : (l0 * x0 - dot) / a;
        double b = (x0 + lx) / (l0 / a + 1.0) / a;
        if (!inverse) b = -b;
        const double sc = inverse ? a : 1.0 / a;
        __syncthreads();
        for (int i = 1 + tid; i < mk; i += nt) x[m + i] = (x[m + i] + b * lm[m + i]) * sc;
        if (tid == 0) x[m] = lx * sc;
        m += mk;
        __syncthreads();
    }
    int ind2 = m;
    for (int k = 0; k < c.ns; ++k) {
        const int mk = c.s[k];
        for (int e = tid; e < mk * mk; e += nt) {
            const int i = e % mk, j = e / mk;
            const double cc = sqrt(lm[ind2 + i]) * sqrt(lm[ind2 + j]);
            x[m + e] = inverse ? x[m + e] * cc : x[m + e] / cc;
        }
        m += mk * mk; ind2 += mk;
    }
}

// x := y o x for the l / q blocks and the 's' blocks with diagonal y (diag == 'D')
__global__ void sprod_kernel(double *x, const double *y, Cones c, int diag_d) {
    __shared__ double sh[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < c.nl; i += nt) x[i] *= y[i];
    int m = c.nl;
    for (int k = 0; k < c.nq; ++k) {
        const int mk = c.q[k];
        double d = 0;
        for (int i = tid; i < mk; i += nt) d += y[m + i] * x[m + i];
        d = blk_sum(d, sh);
        const double y0 = y[m], x0 = x[m];
        __syncthreads();
        for (int i = 1 + tid; i < mk; i += nt) x[m + i] = y0 * x[m + i] + x0 * y[m + i];
        if (tid == 0) x[m] = d;
        m += mk;
        __syncthreads();
    }
    if (!diag_d) return;
    int ind2 = m;
    for (int k = 0; k < c.ns; ++k) {
        const int mk = c.s[k];
        for (int e = tid; e < mk * mk; e += nt) {
            const int i = e % mk, j = e / mk;
            if (i >= j) x[m + e] *= 0.5 * (y[ind2 + i] + y[ind2 + j]);
        }
        m += mk * mk; ind2 += mk;
    }
}
// 's' blocks, full y: x_lower := 0.5 (T + T')  with T = sym(x) sym(y)
__global__ void sprod_s_finish_kernel(double *x, const double *T, int mk) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= mk * mk) return;
    const int i = e % mk, j = e / mk;
    if (i >= j) x[e] = 0.5 * (T[i + (long long)j * mk] + T[j + (long long)i * mk]);
}

__global__ void sinv_kernel(double *x, const double *y, Cones c) {
    __shared__ double sh[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < c.nl; i += nt) x[i] /= y[i];
    int m = c.nl;
    for (int k = 0; k < c.nq; ++k) {
        const int mk = c.q[k];
        double n2 = 0, d = 0;
        for (int i = 1 + tid; i < mk; i += nt) { n2 += y[m + i] * y[m + i]; d += x[m + i] * y[m + i]; }
        n2 = blk_sum(n2, sh); d = blk_sum(d, sh);
        const double nrm = sqrt(n2), y0 = y[m], cx = x[m];
        const double a = (y0 + nrm) * (y0 - nrm);
        const double al1 = a / y0, al2 = d / y0 - cx, ia = 1.0 / a;
        __syncthreads();
        for (int i = 1 + tid; i < mk; i += nt) x[m + i] = (x[m + i] * al1 + al2 * y[m + i]) * ia;
        if (tid == 0) x[m] = (cx * y0 - d) * ia;
        m += mk;
        __syncthreads();
    }
    int ind2 = m;
    for (int k = 0; k < c.ns; ++k) {
        const int mk = c.s[k];
        for (int e = tid; e < mk * mk; e += nt) {
            const int i = e % mk, j = e / mk;
            if (i >= j) x[m + e] /= 0.5 * (y[ind2 + i] + y[ind2 + j]);
        }
        m += mk * mk; ind2 += mk;
    }
}

// mode 0: trisc (upper := 0, strict lower *= 2); mode 1: triusc (strict lower *= 0.5)
__global__ void trisc_kernel(double *x, Cones c, int off, int mode) {
    int m = off;
    for (int k = 0; k < c.ns; ++k) {
        const int mk = c.s[k];
        for (int e = threadIdx.x; e < mk * mk; e += blockDim.x) {
            const int i = e % mk, j = e / mk;
            if (mode == 0) { if (i < j) x[m + e] = 0.0; else if (i > j) x[m + e] *= 2.0; }
            else if (i > j) x[m + e] *= 0.5;
        }
        m += mk * mk;
    }
}

__global__ void sdot_kernel(const double *x, const double *y, Cones c, int nlq, double *out) {
    __shared__ double sh[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    double a = 0;
    for (int i = tid; i < nlq; i += nt) a += x[i] * y[i];
    int m = nlq;
    for (int k = 0; k < c.ns; ++k) {
        const int mk = c.s[k];
        for (int e = tid; e < mk * mk; e += nt) {
            const int i = e % mk, j = e / mk;
            if (i == j) a += x[m + e] * y[m + e];
            else if (i > j) a += 2.0 * x[m + e] * y[m + e];
        }
        m += mk * mk;
    }
    a = blk_sum(a, sh);
    if (tid == 0) *out = a;
}

__global__ void max_step_kernel(const double *x, Cones c, double *out) {
    __shared__ double sh[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    double t = -FLT_MAX;
    for (int i = tid; i < c.nl; i += nt) t = fmax(t, -x[i]);
    // block max through the sum helper's scratch: do a max-reduction by hand
    for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    __syncthreads();
    if ((tid & 31) == 0) sh[tid >> 5] = t;
    __syncthreads();
    t = -FLT_MAX;
    for (int w = 0; w < (nt >> 5); ++w) t = fmax(t, sh[w]);
    __syncthreads();
    int m = c.nl;
    for (int k = 0; k < c.nq; ++k) {
        const int mk = c.q[k];
        double n2 = 0;
        for (int i = 1 + tid; i < mk; i += nt) n2 += x[m + i] * x[m + i];
        n2 = blk_sum(n2, sh);
        t = fmax(t, sqrt(n2) - x[m]);
        m += mk;
        __syncthreads();
    }
    if (tid == 0) *out = (m > 0) ? t : 0.0;
}


// ---- symmetric eigensolver for the 's' blocks (parallel-order cyclic Jacobi) ----------------
// Round r of the round-robin ("chess tournament") ordering on N (even) indices: N/2 disjoint pairs.
__device__ __forceinline__ void rr_pair(int N, int r, int k, int &p, int &q) {
    int a, b;
    if (k == 0) { a = N - 1; b = r; }
    else { a = (r + k) % (N - 1); b = (r - k + (N - 1)) % (N - 1); }
    p = min(a, b); q = max(a, b);
}

// rotation J = [c s; -s c] annihilating a_pq in J'[app apq; apq aqq]J;  t = tan of the angle
__device__ __forceinline__ void jac_rot(double app, double aqq, double apq, double &c, double &s, double &t) {
    if (apq == 0.0) { c = 1.0; s = 0.0; t = 0.0; return; }
    const double tau = (aqq - app) / (2.0 * apq);
    t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));     // tau*tau = inf -> t = 0
    c = 1.0 / sqrt(1.0 + t * t);
    s = t * c;
}

// One thread's share of a round: the 2x2 block (rows of pair I, columns of pair J) of
// dst = J'.src.J, and the same block of V := V.J.  Indices >= mk belong to the padding.
template <bool WITH_V>
__device__ __forceinline__ void jac_block(const double *src, double *dst, double *V, int mk, int N,
                                          int r, int I, int J) {
    int pi, qi, pj, qj;
    rr_pair(N, r, I, pi, qi);
    rr_pair(N, r, J, pj, qj);
    if (pi >= mk || pj >= mk) return;
    const bool vi = qi < mk, vj = qj < mk;
    double ci = 1, si = 0, ti = 0, cj = 1, sj = 0, tj = 0;
    if (vi) jac_rot(src[pi + (size_t)pi * mk], src[qi + (size_t)qi * mk], src[qi + (size_t)pi * mk], ci, si, ti);
    if (I == J) { cj = ci; sj = si; tj = ti; }
    else if (vj) jac_rot(src[pj + (size_t)pj * mk], src[qj + (size_t)qj * mk], src[qj + (size_t)pj * mk], cj, sj, tj);
    const double b00 = src[pi + (size_t)pj * mk];
    const double b01 = vj ? src[pi + (size_t)qj * mk] : 0.0;
    const double b10 = vi ? src[qi + (size_t)pj * mk] : 0.0;
    const double b11 = (vi && vj) ? src[qi + (size_t)qj * mk] : 0.0;
    double d00, d01, d10, d11;
    if (I == J) {
        d00 = b00 - ti * b10; d11 = b11 + ti * b10; d01 = 0.0; d10 = 0.0;
    } else {
        const double r00 = ci * b00 - si * b10, r01 = ci * b01 - si * b11;
        const double r10 = si * b00 + ci * b10, r11 = si * b01 + ci * b11;
        d00 = cj * r00 - sj * r01; d01 = sj * r00 + cj * r01;
        d10 = cj * r10 - sj * r11; d11 = sj * r10 + cj * r11;
    }
    dst[pi + (size_t)pj * mk] = d00;
    if (vj) dst[pi + (size_t)qj * mk] = d01;
    if (vi) dst[qi + (size_t)pj * mk] = d10;
    if (vi && vj) dst[qi + (size_t)qj * mk] = d11;
    if (WITH_V) {
        const double v00 = V[pi + (size_t)pj * mk];
        const double v01 = vj ? V[pi + (size_t)qj * mk] : 0.0;
        V[pi + (size_t)pj * mk] = cj * v00 - sj * v01;
        if (vj) V[pi + (size_t)qj * mk] = sj * v00 + cj * v01;
        if (vi) {
            const double v10 = V[qi + (size_t)pj * mk];
            const double v11 = vj ? V[qi + (size_t)qj * mk] : 0.0;
            V[qi + (size_t)pj * mk] = cj * v10 - sj * v11;
            if (vj) V[qi + (size_t)qj * mk] = sj * v10 + cj * v11;
        }
    }
}

struct JacArgs {
    const double *x;        // first 's' row of the cone vector (lower triangles significant)
    double *w0, *w1, *V;    // three work copies, each sum(mk^2), same block offsets as x
    const int *s, *soff, *sigoff;
    double *sigma;          // sum(mk): eigenvalues, ascending inside each block
    double *xout;           // eigenvectors -> 's' blocks of x (nullptr: eigenvalues only)
    double *stats;          // 2 per block: off-diagonal and total squared Frobenius norms
    int *perm;              // sum(mk): rank of each unsorted eigenvalue
    int N;                  // padded (even) order shared by all blocks of the launch
};

__global__ void jac_init_kernel(JacArgs a) {
    const int k = blockIdx.y, mk = a.s[k];
    const size_t o = a.soff[k];
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)mk * mk; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % mk), j = (int)(e / mk);
        a.w0[o + e] = (i >= j) ? a.x[o + e] : a.x[o + j + (size_t)i * mk];
        if (a.V) a.V[o + e] = (i == j) ? 1.0 : 0.0;
    }
}

template <bool WITH_V>
__global__ void jac_round_kernel(JacArgs a, int r, int flip) {
    const int k = blockIdx.z, mk = a.s[k];
    // I (the row pair) is the fast thread index: consecutive I are consecutive rows of the column-major blocks
    const int I = blockIdx.x * blockDim.x + threadIdx.x, J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= a.N / 2 || J >= a.N / 2) return;
    const size_t o = a.soff[k];
    jac_block<WITH_V>((flip ? a.w1 : a.w0) + o, (flip ? a.w0 : a.w1) + o, a.V + (WITH_V ? o : 0), mk, a.N, r, I, J);
}

// A whole sweep (N - 1 rounds) in one cooperative launch, a grid barrier between rounds; the buffers swap every round,
// so the result of the sweep is in the buffer the sweep did NOT start from (N - 1 is odd).
template <bool WITH_V>
__global__ void __launch_bounds__(256) jac_sweep_kernel(JacArgs a, int flip) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int k = blockIdx.z, mk = a.s[k];
    const int I = blockIdx.x * blockDim.x + threadIdx.x, J = blockIdx.y * blockDim.y + threadIdx.y;
    const bool live = I < a.N / 2 && J < a.N / 2;
    const size_t o = a.soff[k];
    for (int r = 0; r < a.N - 1; ++r) {
        if (live)
            jac_block<WITH_V>((flip ? a.w1 : a.w0) + o, (flip ? a.w0 : a.w1) + o, a.V + (WITH_V ? o : 0), mk, a.N, r, I, J);
        grid.sync();
        flip ^= 1;
    }
}

__global__ void jac_off_kernel(JacArgs a, int flip) {
    __shared__ double sh[32];
    const int k = blockIdx.x, mk = a.s[k];
    const double *S = (flip ? a.w1 : a.w0) + a.soff[k];
    double off = 0, tot = 0;
    for (size_t e = threadIdx.x; e < (size_t)mk * mk; e += blockDim.x) {
        const double v = S[e] * S[e];
        tot += v;
        if (e % mk != e / mk) off += v;
    }
    off = blk_sum(off, sh); tot = blk_sum(tot, sh);
    if (threadIdx.x == 0) { a.stats[2 * k] = off; a.stats[2 * k + 1] = tot; }
}

// eigenvalues = diagonal; rank them (stable), write sigma ascending
__global__ void jac_sort_kernel(JacArgs a, int flip) {
    const int k = blockIdx.x, mk = a.s[k];
    const double *S = (flip ? a.w1 : a.w0) + a.soff[k];
    for (int i = threadIdx.x; i < mk; i += blockDim.x) {
        const double di = S[i + (size_t)i * mk];
        int rank = 0;
        for (int j = 0; j < mk; ++j) {
            const double dj = S[j + (size_t)j * mk];
            rank += (dj < di) || (dj == di && j < i);
        }
        a.sigma[a.sigoff[k] + rank] = di;
        a.perm[a.sigoff[k] + i] = rank;
    }
}

__global__ void jac_vec_kernel(JacArgs a) {
    const int k = blockIdx.y, mk = a.s[k];
    const size_t o = a.soff[k];
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)mk * mk; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % mk), j = (int)(e / mk);
        a.xout[o + i + (size_t)a.perm[a.sigoff[k] + j] * mk] = a.V[o + e];
    }
}

// convergence of one block from its (off^2, total^2): at rounding level, or stagnating just above it
__host__ __device__ inline bool jac_done(double off2, double tot2, double prev_off2, int mk) {
    const double eps = 2.220446049250313e-16;
    if (!(off2 > eps * eps * (double)mk * tot2)) return off2 == off2;       // NaN never converges
    return off2 <= 1e-26 * tot2 && off2 >= 0.25 * prev_off2;
}

// All blocks of order <= 64: one CTA per block runs every sweep itself (block-level barriers only).
template <bool WITH_V>
__global__ void __launch_bounds__(1024) jac_small_kernel(JacArgs a, int max_sweeps, int *fail) {
    __shared__ double sh[32];
    const int k = blockIdx.x, mk = a.s[k];
    if (mk == 0) return;
    const size_t o = a.soff[k];
    double *w0 = a.w0 + o, *w1 = a.w1 + o, *V = WITH_V ? a.V + o : nullptr;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < mk * mk; e += nt) {
        const int i = e % mk, j = e / mk;
        w0[e] = (i >= j) ? a.x[o + e] : a.x[o + j + (size_t)i * mk];
        if (WITH_V) V[e] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int N = max(2, mk + (mk & 1)), h = N / 2;
    const int I = tid % h, J = tid / h;          // I fastest: rows of the column-major block
    double prev = 1e300;
    int sweep = 0;
    bool ok = false;
    for (;; ++sweep) {
        double off = 0, tot = 0;
        for (int e = tid; e < mk * mk; e += nt) {
            const double v = w0[e] * w0[e];
            tot += v;
            if (e % mk != e / mk) off += v;
        }
        off = blk_sum(off, sh); tot = blk_sum(tot, sh);
        if (jac_done(off, tot, prev, mk)) { ok = true; break; }
        if (sweep == max_sweeps) break;
        prev = off;
        for (int r = 0; r < N - 1; ++r) {
            if (J < h) jac_block<WITH_V>(w0, w1, V, mk, N, r, I, J);
            __syncthreads();
            double *t = w0; w0 = w1; w1 = t;
        }
    }
    if (!ok && tid == 0) atomicExch(fail, 1);
    for (int i = tid; i < mk; i += nt) {
        const double di = w0[i + (size_t)i * mk];
        int rank = 0;
        for (int j = 0; j < mk; ++j) {
            const double dj = w0[j + (size_t)j * mk];
            rank += (dj < di) || (dj == di && j < i);
        }
        a.sigma[a.sigoff[k] + rank] = di;
        if (WITH_V) a.perm[a.sigoff[k] + i] = rank;
    }
    if (WITH_V) {
        __syncthreads();
        for (int e = tid; e < mk * mk; e += nt) {
            const int i = e % mk, j = e / mk;
            a.xout[o + i + (size_t)a.perm[a.sigoff[k] + j] * mk] = V[e];
        }
    }
}

// t := max(t, -lambda_min) over the 's' blocks (sigma ascending per block)
__global__ void max_step_s_kernel(double *out, const double *sigma, const int *s, const int *sigoff,
                                  int ns, int any_lq) {
    double t = any_lq ? *out : -FLT_MAX;
    for (int k = 0; k < ns; ++k) if (s[k] > 0) t = fmax(t, -sigma[sigoff[k]]);
    *out = t;
}

struct VCtx { cudaStream_t st = nullptr; bool ok = false; };
VCtx g_v;
int vctx(cudaStream_t *st) {
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: cvxopt_b200 has no CPU fallback");
        return CVXB_E_NOGPU;
    }
    CVXB_CUDA(cudaSetDevice(0));
    static std::mutex mu;               // creation only; a CUDA stream itself may be shared by threads
    std::lock_guard<std::mutex> g(mu);
    if (!g_v.ok) { CVXB_CUDA(cudaStreamCreateWithFlags(&g_v.st, cudaStreamNonBlocking)); g_v.ok = true; }
    *st = g_v.st;
    return 0;
}

struct Buf {       // host buffer staged on the device (or a device pointer used in place)
    double *dev = nullptr, *host = nullptr; size_t n = 0; bool owned = false;
    ~Buf() { if (owned && dev) tmp_free(dev); }
    int in(const double *src, size_t count, int space, cudaStream_t st) {
        n = count; host = const_cast<double *>(src);
        if (space == CVXB_DEVICE) { dev = host; return 0; }
        CVXB_CUDA(tmp_malloc(&dev, (n ? n : 1) * sizeof(double)));
        owned = true;
        if (n) CVXB_CUDA(cudaMemcpyAsync(dev, src, n * sizeof(double), cudaMemcpyHostToDevice, st));
        return 0;
    }
    int out(cudaStream_t st) {
        if (owned && n) CVXB_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        return 0;
    }
};

struct Lay {
    ConeLayout c; Cones k;
    int init(const cvxb_dims *dims) {
        int rc = c.init(dims);
        if (rc) return rc;
        k.nl = c.mnl + c.ml; k.nq = c.nq; k.q = c.d_q; k.ns = c.ns; k.s = c.d_s;
        return 0;
    }
    ~Lay() { c.destroy(); }
};

}  // namespace

extern "C" {

int cvxb_scale2(const double *lmbda, double *x, const cvxb_dims *dims, int inverse, int space) {
    cudaStream_t st; CVXB_TRY(vctx(&st));
    Lay L; CVXB_TRY(L.init(dims));
    Buf l, xb;
    CVXB_TRY(l.in(lmbda, (size_t)L.c.mnl + L.c.ml + L.c.sumq + [&] { int t = 0; for (int v : L.c.s) t += v; return t; }(), space, st));
    CVXB_TRY(xb.in(x, L.c.cdim, space, st));
    scale2_kernel<<<1, 256, 0, st>>>(l.dev, xb.dev, L.k, inverse == 'I');
    count_launch(); CVXB_LAUNCH_CHECK();
    CVXB_TRY(xb.out(st));
    CVXB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int cvxb_sprod(double *x, const double *y, const cvxb_dims *dims, int diag, int space) {
    cudaStream_t st; CVXB_TRY(vctx(&st));
    Lay L; CVXB_TRY(L.init(dims));
    const bool dd = (diag == 'D');
    size_t ny = dd ? (size_t)L.c.mnl + L.c.ml + L.c.sumq + [&] { int t = 0; for (int v : L.c.s) t += v; return t; }()
                   : (size_t)L.c.cdim;
    Buf xb, yb;
    CVXB_TRY(xb.in(x, L.c.cdim, space, st));
    CVXB_TRY(yb.in(y, ny, space, st));
    sprod_kernel<<<1, 256, 0, st>>>(xb.dev, yb.dev, L.k, dd ? 1 : 0);
    count_launch(); CVXB_LAUNCH_CHECK();
    if (!dd && L.c.ns > 0) {
        // 0.5 (A Y + Y A) with A = sym(x_k), Y = sym(y_k): T = A Y on the DMMA GEMM
        const int nlq = L.c.mnl + L.c.ml + L.c.sumq;
        for (int k = 0; k < L.c.ns; ++k) {
            const int mk = L.c.s[k];
            if (mk == 0) continue;
            const long long m2 = (long long)mk * mk;
            double *tmp = nullptr;
            CVXB_CUDA(tmp_malloc(&tmp, 3 * m2 * sizeof(double)));
            double *As = tmp, *Ys = tmp + m2, *T = tmp + 2 * m2;
            int rc = 0;
            do {
                if (cudaMemcpyAsync(As, xb.dev + nlq + L.c.s_off[k], m2 * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
                    cudaMemcpyAsync(Ys, yb.dev + nlq + L.c.s_off[k], m2 * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess) { rc = CVXB_E_CUDA; break; }
                if ((rc = symmetrize_lower(mk, As, mk, 1, 0, st))) break;
                if ((rc = symmetrize_lower(mk, Ys, mk, 1, 0, st))) break;
                GemmDesc g;
                g.M = mk; g.N = mk; g.K = mk;
                g.X = As; g.ldx = mk; g.x_kmajor = false;
                g.Y = Ys; g.ldy = mk; g.y_kmajor = true;       // Y[c,k] = Ys[k, c]
                g.C = T; g.ldc = mk;
                if ((rc = dmma_gemm(g, st))) break;
                sprod_s_finish_kernel<<<(int)((m2 + 255) / 256), 256, 0, st>>>(xb.dev + nlq + L.c.s_off[k], T, mk);
                count_launch();
            } while (0);
            cudaStreamSynchronize(st);
            tmp_free(tmp);
            if (rc) return rc;
        }
    }
    CVXB_TRY(xb.out(st));
    CVXB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int cvxb_sinv(double *x, const double *y, const cvxb_dims *dims, int space) {
    cudaStream_t st; CVXB_TRY(vctx(&st));
    Lay L; CVXB_TRY(L.init(dims));
    size_t ny = (size_t)L.c.mnl + L.c.ml + L.c.sumq + [&] { int t = 0; for (int v : L.c.s) t += v; return t; }();
    Buf xb, yb;
    CVXB_TRY(xb.in(x, L.c.cdim, space, st));
    CVXB_TRY(yb.in(y, ny, space, st));
    sinv_kernel<<<1, 256, 0, st>>>(xb.dev, yb.dev, L.k);
    count_launch(); CVXB_LAUNCH_CHECK();
    CVXB_TRY(xb.out(st));
    CVXB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static int trisc_common(double *x, const cvxb_dims *dims, int space, int mode) {
    cudaStream_t st; CVXB_TRY(vctx(&st));
    Lay L; CVXB_TRY(L.init(dims));
    Buf xb;
    CVXB_TRY(xb.in(x, L.c.cdim, space, st));
    trisc_kernel<<<1, 256, 0, st>>>(xb.dev, L.k, L.c.mnl + L.c.ml + L.c.sumq, mode);
    count_launch(); CVXB_LAUNCH_CHECK();
    CVXB_TRY(xb.out(st));
    CVXB_CUDA(cudaStreamSynchronize(st));
    return 0;
}
int cvxb_trisc(double *x, const cvxb_dims *dims, int space) { return trisc_common(x, dims, space, 0); }
int cvxb_triusc(double *x, const cvxb_dims *dims, int space) { return trisc_common(x, dims, space, 1); }

int cvxb_sdot(const double *x, const double *y, const cvxb_dims *dims, double *result, int space) {
    if (!result) { set_error("sdot: result is NULL"); return CVXB_E_ARG; }
    cudaStream_t st; CVXB_TRY(vctx(&st));
    Lay L; CVXB_TRY(L.init(dims));
    Buf xb, yb;
    CVXB_TRY(xb.in(x, L.c.cdim, space, st));
    CVXB_TRY(yb.in(y, L.c.cdim, space, st));
    double *d = nullptr;
    CVXB_CUDA(tmp_malloc(&d, sizeof(double)));
    sdot_kernel<<<1, 256, 0, st>>>(xb.dev, yb.dev, L.k, L.c.mnl + L.c.ml + L.c.sumq, d);
    count_launch();
    cudaError_t e = cudaMemcpyAsync(result, d, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    tmp_free(d);
    if (e != cudaSuccess) { set_error("sdot: %s", cudaGetErrorString(e)); return CVXB_E_CUDA; }
    return 0;
}

// Eigen-decomposition of every 's' block of the device cone vector xs (first 's' row).
// sigma_dev: sum(mk) eigenvalues (ascending per block); with_vectors: the blocks of xs are
// overwritten by the eigenvectors (columns ordered like sigma), as dsyevd_ 'V' does in the reference.
static int sym_eig_blocks(const ConeLayout &c, double *xs, double *sigma_dev, const int *d_sigoff,
                          bool with_vectors, cudaStream_t st) {
    const int MAX_SWEEPS = 40;
    int sums = 0;
    for (int v : c.s) sums += v;
    if (c.maxs == 0) return 0;
    const size_t m2 = (size_t)c.sums2;
    double *work = nullptr, *stats = nullptr;
    int *perm = nullptr, *fail = nullptr;
    std::vector<double> hstats(2 * (size_t)c.ns), prev(c.ns, 1e300);
    int rc = 0;
    auto done = [&](int r) {
        cudaStreamSynchronize(st);
        tmp_free(work); tmp_free(stats); tmp_free(perm); tmp_free(fail);
        return r;
    };
    if (tmp_malloc(&work, (with_vectors ? 3 : 2) * m2 * sizeof(double)) != cudaSuccess ||
        tmp_malloc(&stats, 2 * (size_t)c.ns * sizeof(double)) != cudaSuccess ||
        tmp_malloc(&perm, (size_t)(sums ? sums : 1) * sizeof(int)) != cudaSuccess ||
        tmp_malloc(&fail, sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        set_error("max_step: out of device memory for the eigensolver workspace");
        return done(CVXB_E_NOMEM);
    }
    JacArgs a;
    a.x = xs; a.w0 = work; a.w1 = work + m2; a.V = with_vectors ? work + 2 * m2 : nullptr;
    a.s = c.d_s; a.soff = c.d_soff; a.sigoff = d_sigoff;
    a.sigma = sigma_dev; a.xout = with_vectors ? xs : nullptr; a.stats = stats; a.perm = perm;
    a.N = std::max(2, c.maxs + (c.maxs & 1));
    const int h = a.N / 2;
    if (c.maxs <= 64) {
        if (cudaMemsetAsync(fail, 0, sizeof(int), st) != cudaSuccess) return done(CVXB_E_CUDA);
        const int nt = std::max(32, (h * h + 31) / 32 * 32);
        if (with_vectors) jac_small_kernel<true><<<c.ns, nt, 0, st>>>(a, MAX_SWEEPS, fail);
        else jac_small_kernel<false><<<c.ns, nt, 0, st>>>(a, MAX_SWEEPS, fail);
        count_launch();
        int hfail = 0;
        cudaError_t e = cudaMemcpyAsync(&hfail, fail, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("max_step: %s", cudaGetErrorString(e)); return done(CVXB_E_CUDA); }
        if (hfail) { set_error("max_step: Jacobi eigensolver did not converge (non-finite input?)"); rc = 1; }
        return done(rc);
    }
    {
        const int gx = (int)std::min<size_t>(((size_t)c.maxs * c.maxs + 255) / 256, 1184);
        jac_init_kernel<<<dim3(gx, c.ns), 256, 0, st>>>(a);
        count_launch();
    }
    int flip = 0;
    bool ok = false;
    // one cooperative launch per sweep when all its CTAs can be resident at once (CVXB_JACOBI_COOP=0: per-round launches)
    bool coop = false;
    {
        static int coop_on = -1;
        if (coop_on < 0) { const char *e = getenv("CVXB_JACOBI_COOP"); coop_on = (e && e[0] == '0') ? 0 : 1; }
        int dev = 0, can = 0, per_sm = 0, sms = 0;
        const void *fn = with_vectors ? (const void *)jac_sweep_kernel<true> : (const void *)jac_sweep_kernel<false>;
        if (coop_on && cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && can &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0) == cudaSuccess)
            coop = (long long)per_sm * sms >= (long long)((h + 15) / 16) * ((h + 15) / 16) * c.ns;
    }
    for (int sweep = 0; sweep <= MAX_SWEEPS; ++sweep) {
        jac_off_kernel<<<c.ns, 256, 0, st>>>(a, flip);
        count_launch();
        cudaError_t e = cudaMemcpyAsync(hstats.data(), stats, hstats.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("max_step: %s", cudaGetErrorString(e)); return done(CVXB_E_CUDA); }
        ok = true;
        for (int k = 0; k < c.ns; ++k) {
            if (c.s[k] && !jac_done(hstats[2 * k], hstats[2 * k + 1], prev[k], c.s[k])) ok = false;
            prev[k] = hstats[2 * k];
        }
        if (ok || sweep == MAX_SWEEPS) break;
        const dim3 blk(16, 16), grd((h + 15) / 16, (h + 15) / 16, c.ns);
        if (coop) {
            void *args[] = {&a, &flip};
            const void *fn = with_vectors ? (const void *)jac_sweep_kernel<true> : (const void *)jac_sweep_kernel<false>;
            if (cudaLaunchCooperativeKernel(fn, grd, blk, args, 0, st) != cudaSuccess) {
                set_error("max_step: cooperative launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return done(CVXB_E_CUDA);
            }
            count_launch();
            flip ^= 1;                                   // N - 1 (odd) buffer swaps
        } else {
            for (int r = 0; r < a.N - 1; ++r) {
                if (with_vectors) jac_round_kernel<true><<<grd, blk, 0, st>>>(a, r, flip);
                else jac_round_kernel<false><<<grd, blk, 0, st>>>(a, r, flip);
                count_launch();
                flip ^= 1;
            }
        }
    }
    if (!ok) { set_error("max_step: Jacobi eigensolver did not converge (non-finite input?)"); return done(1); }
    jac_sort_kernel<<<c.ns, 256, 0, st>>>(a, flip);
    count_launch();
    if (with_vectors) {
        const int gx = (int)std::min<size_t>(((size_t)c.maxs * c.maxs + 255) / 256, 1184);
        jac_vec_kernel<<<dim3(gx, c.ns), 256, 0, st>>>(a);
        count_launch();
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("max_step: %s", cudaGetErrorString(e)); return done(CVXB_E_CUDA); }
    return done(0);
}

int cvxb_max_step(double *x, const cvxb_dims *dims, double *sigma, double *result, int space) {
    if (!result) { set_error("max_step: result is NULL"); return CVXB_E_ARG; }
    cudaStream_t st; CVXB_TRY(vctx(&st));
    Lay L; CVXB_TRY(L.init(dims));
    const int nlq = L.c.mnl + L.c.ml + L.c.sumq;
    int sums = 0;
    std::vector<int> sigoff(L.c.ns);
    for (int k = 0; k < L.c.ns; ++k) { sigoff[k] = sums; sums += L.c.s[k]; }
    Buf xb;
    CVXB_TRY(xb.in(x, L.c.cdim, space, st));
    double *d = nullptr, *dsig = nullptr;
    int *dsigoff = nullptr;
    auto done = [&](int r) { tmp_free(d); tmp_free(dsig); tmp_free(dsigoff); return r; };
    CVXB_CUDA(tmp_malloc(&d, sizeof(double)));
    max_step_kernel<<<1, 256, 0, st>>>(xb.dev, L.k, d);
    count_launch();
    if (L.c.maxs > 0) {
        // 's' blocks: lambda_min of each block (reference dsyevr_ range 'I' 1..1, or dsyevd_ 'V' when
        // sigma is given: eigenvalues -> sigma, eigenvectors -> x; misc_solvers.c:1099-1150)
        if (tmp_malloc(&dsig, (size_t)sums * sizeof(double)) != cudaSuccess ||
            tmp_malloc(&dsigoff, (size_t)L.c.ns * sizeof(int)) != cudaSuccess ||
            cudaMemcpyAsync(dsigoff, sigoff.data(), (size_t)L.c.ns * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) {
            cudaGetLastError();
            set_error("max_step: device allocation failed");
            return done(CVXB_E_NOMEM);
        }
        int rc = sym_eig_blocks(L.c, xb.dev + nlq, dsig, dsigoff, sigma != nullptr, st);
        if (rc) return done(rc);
        max_step_s_kernel<<<1, 1, 0, st>>>(d, dsig, L.c.d_s, dsigoff, L.c.ns, nlq > 0);
        count_launch();
        if (sigma) {
            if (cudaMemcpyAsync(sigma, dsig, (size_t)sums * sizeof(double),
                                space == CVXB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st) != cudaSuccess) {
                set_error("max_step: copy of sigma failed");
                return done(CVXB_E_CUDA);
            }
            int rc2 = xb.out(st);
            if (rc2) return done(rc2);
        }
    }
    cudaError_t e = cudaMemcpyAsync(result, d, sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("max_step: %s", cudaGetErrorString(e)); return done(CVXB_E_CUDA); }
    return done(0);
}

}  // extern "C"
