// The O(cdim) cone algebra of the IPM side, mirror of src/C/misc_solvers.c:
//   scale2 (:256-401), sprod (:634-767), sinv (:775-878), trisc (:887-935), triusc (:940-986),
//   sdot (:991-1039), max_step (:1052-1153; 'l' and 'q' cones — the 's' part needs a symmetric
//   eigensolver and returns CVXB_E_UNSUP).
// One CTA walks the whole cone vector; reductions inside a 'q' cone / for sdot are block-wide.
#include "cone.cuh"
#include <map>
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
    if (!g_v.ok) { CVXB_CUDA(cudaStreamCreateWithFlags(&g_v.st, cudaStreamNonBlocking)); g_v.ok = true; }
    *st = g_v.st;
    return 0;
}

struct Buf {       // host buffer staged on the device (or a device pointer used in place)
    double *dev = nullptr, *host = nullptr; size_t n = 0; bool owned = false;
    ~Buf() { if (owned && dev) cudaFree(dev); }
    int in(const double *src, size_t count, int space, cudaStream_t st) {
        n = count; host = const_cast<double *>(src);
        if (space == CVXB_DEVICE) { dev = host; return 0; }
        CVXB_CUDA(cudaMalloc(&dev, (n ? n : 1) * sizeof(double)));
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
            CVXB_CUDA(cudaMalloc(&tmp, 3 * m2 * sizeof(double)));
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
            cudaFree(tmp);
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
    CVXB_CUDA(cudaMalloc(&d, sizeof(double)));
    sdot_kernel<<<1, 256, 0, st>>>(xb.dev, yb.dev, L.k, L.c.mnl + L.c.ml + L.c.sumq, d);
    count_launch();
    cudaError_t e = cudaMemcpyAsync(result, d, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("sdot: %s", cudaGetErrorString(e)); return CVXB_E_CUDA; }
    return 0;
}

int cvxb_max_step(double *x, const cvxb_dims *dims, double *sigma, double *result, int space) {
    if (!result) { set_error("max_step: result is NULL"); return CVXB_E_ARG; }
    cudaStream_t st; CVXB_TRY(vctx(&st));
    Lay L; CVXB_TRY(L.init(dims));
    if (L.c.maxs > 0) {
        (void)sigma;
        set_error("max_step: 's' blocks need a symmetric eigensolver (reference dsyevr/dsyevd, "
                  "misc_solvers.c:1132-1143): not built on the device yet");
        return CVXB_E_UNSUP;
    }
    Buf xb;
    CVXB_TRY(xb.in(x, L.c.cdim, space, st));
    double *d = nullptr;
    CVXB_CUDA(cudaMalloc(&d, sizeof(double)));
    max_step_kernel<<<1, 256, 0, st>>>(xb.dev, L.k, d);
    count_launch();
    cudaError_t e = cudaMemcpyAsync(result, d, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("max_step: %s", cudaGetErrorString(e)); return CVXB_E_CUDA; }
    return 0;
}

}  // extern "C"
