// Device-resident primal-dual interior-point method for a BATCH of independent dense QPs
//
//      minimize  1/2 x'P x + q'x    subject to  G x + s = h,  s >= 0          ('l' cone, no A)
//
// run in lock-step, one problem per CTA-group, with per-problem convergence masks
// (BASELINE config 4; the reference has no batch API — its counterpart is a Python loop over
// solvers.qp).  The algorithm is a restatement of coneprog.coneqp for dims = {'l': m}
// (reference src/python/coneprog.py:1998-2547): same starting point (:2055-2106), residuals
// and stopping rule (:2169-2234), Nesterov-Todd scaling d = sqrt(s/z) (misc.py:284-287),
// Mehrotra predictor/corrector with STEP 0.99 / EXPON 3 (:2357-2456) and scaling update
// (misc.py:450-464).  Every KKT solve is the same path as cvxb_kkt_*: fused-scaling SYRK,
// Cholesky, GEMV/TRSV — here batched over the problems through blockIdx.z / blockIdx.y.
// Nothing leaves the device between iterations except one int ("how many are done").
#include "cone.cuh"
#include <cstdlib>

using namespace cvxb;

namespace {

struct Scal {                       // per-problem scalars, device resident
    double resx0, resz0, gap, mu, sigma, eta, step, dsdz;
    double xPxq, xq, resx, resz, zrz, pcost, dcost, relgap, pres, dres;
    int relgap_valid, done, iters, status;   // status: 0 running, 1 optimal, 2 maxiters, 3 singular
};

__device__ __forceinline__ double block_sum(double v, double *sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    if (warp == 0) t = warp_sum(t);
    if (threadIdx.x == 0) sh[0] = t;
    __syncthreads();
    return sh[0];
}
__device__ __forceinline__ double block_min(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : INFINITY;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = fmin(t, __shfl_xor_sync(0xffffffffu, t, o));
    }
    if (threadIdx.x == 0) sh[0] = t;
    __syncthreads();
    return sh[0];
}

struct Ptrs {
    int n, m;
    const double *q, *h;
    double *x, *s, *z, *rx, *rz, *dx, *ds, *dz, *lmbda, *lmbdasq, *d, *di, *di2, *ws3, *bzp;
    Scal *sc;
};
#define PB_SETUP                                                  \
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x; \
    const long long on = (long long)b * p.n, om = (long long)b * p.m; \
    __shared__ double sh[32];                                     \
    Scal &S = p.sc[b];

// starting point, part 1: rhs of [P G'; G -I][x; z] = [-q; h] with W = I   (coneprog.py:2076-2080)
__global__ void k_init_rhs(Ptrs p) {
    PB_SETUP
    double nq = 0, nh = 0;
    for (int i = tid; i < p.n; i += nt) { double v = p.q[on + i]; p.dx[on + i] = -v; nq += v * v; }
    for (int i = tid; i < p.m; i += nt) {
        double v = p.h[om + i];
        p.dz[om + i] = v;
        p.d[om + i] = 1.0; p.di[om + i] = 1.0; p.di2[om + i] = 1.0;
        nh += v * v;
    }
    nq = block_sum(nq, sh);
    nh = block_sum(nh, sh);
    if (tid == 0) {
        S.resx0 = fmax(1.0, sqrt(nq));                  // :1998
        S.resz0 = fmax(1.0, sqrt(nh));                  // :2000 (snrm2 == 2-norm for 'l')
        S.done = 0; S.iters = 0; S.status = 0; S.sigma = 0; S.eta = 0; S.step = 0;
    }
}
// bzp = di .* bz  (W^{-T} bz for the 'l' cone)
__global__ void k_scale_bz(Ptrs p, const double *bz) {
    PB_SETUP
    (void)sh; (void)S; (void)on;
    for (int i = tid; i < p.m; i += nt) p.bzp[om + i] = p.di[om + i] * bz[om + i];
}
// starting point, part 2: x = dx, z = dz (solution), s = -z, shifts (:2083-2106), gap (:2165)
__global__ void k_init_point(Ptrs p) {
    PB_SETUP
    double ns = 0, mins = INFINITY;
    for (int i = tid; i < p.n; i += nt) p.x[on + i] = p.dx[on + i];
    for (int i = tid; i < p.m; i += nt) {
        double zv = p.bzp[om + i];      // solve leaves W*uz in bzp
        p.z[om + i] = zv;
        p.s[om + i] = -zv;
        ns += zv * zv;
        mins = fmin(mins, -zv);
    }
    ns = sqrt(block_sum(ns, sh));
    mins = block_min(mins, sh);
    const double ts = -mins;                             // max_step(s) = -min(s) for 'l'
    double minz = INFINITY;
    for (int i = tid; i < p.m; i += nt) minz = fmin(minz, p.z[om + i]);
    minz = block_min(minz, sh);
    const double tz = -minz;
    const double as = (ts >= -1e-8 * fmax(ns, 1.0)) ? 1.0 + ts : 0.0;
    const double az = (tz >= -1e-8 * fmax(ns, 1.0)) ? 1.0 + tz : 0.0;   // nrmz == nrms here
    double gap = 0;
    for (int i = tid; i < p.m; i += nt) {
        double sv = p.s[om + i] + as, zv = p.z[om + i] + az;
        p.s[om + i] = sv; p.z[om + i] = zv;
        gap += sv * zv;
    }
    gap = block_sum(gap, sh);
    if (tid == 0) S.gap = gap;
}
// rx = q  (then rx += P x by GEMV)
__global__ void k_res_begin(Ptrs p) {
    PB_SETUP
    (void)sh; (void)S;
    for (int i = tid; i < p.n; i += nt) p.rx[on + i] = p.q[on + i];
    for (int i = tid; i < p.m; i += nt) p.rz[om + i] = p.s[om + i] - p.h[om + i];      // :2183-2184
}
// f0 pieces once rx = P x + q   (:2172)
__global__ void k_res_dots(Ptrs p) {
    PB_SETUP
    double a = 0, c = 0;
    for (int i = tid; i < p.n; i += nt) { double xv = p.x[on + i]; a += xv * p.rx[on + i]; c += xv * p.q[on + i]; }
    a = block_sum(a, sh); c = block_sum(c, sh);
    if (tid == 0) { S.xPxq = a; S.xq = c; }
}
// statistics + stopping rule (:2175-2234)
__global__ void k_stats(Ptrs p, int iter, int maxiters, double abstol, double reltol, double feastol,
                        int *ndone, int *doneflags) {
    PB_SETUP
    double rx2 = 0, rz2 = 0, zrz = 0;
    for (int i = tid; i < p.n; i += nt) { double v = p.rx[on + i]; rx2 += v * v; }
    for (int i = tid; i < p.m; i += nt) { double v = p.rz[om + i]; rz2 += v * v; zrz += p.z[om + i] * v; }
    rx2 = block_sum(rx2, sh); rz2 = block_sum(rz2, sh); zrz = block_sum(zrz, sh);
    if (tid == 0) {
        if (!S.done) {
            const double f0 = 0.5 * (S.xPxq + S.xq);
            S.resx = sqrt(rx2); S.resz = sqrt(rz2); S.zrz = zrz;
            S.pcost = f0;
            S.dcost = f0 + zrz - S.gap;
            if (S.pcost < 0.0) { S.relgap = S.gap / -S.pcost; S.relgap_valid = 1; }
            else if (S.dcost > 0.0) { S.relgap = S.gap / S.dcost; S.relgap_valid = 1; }
            else { S.relgap = 0.0; S.relgap_valid = 0; }
            S.pres = S.resz / S.resz0;
            S.dres = S.resx / S.resx0;
            const bool opt = S.pres <= feastol && S.dres <= feastol &&
                             (S.gap <= abstol || (S.relgap_valid && S.relgap <= reltol));
            if (opt || iter == maxiters) {
                S.done = 1; S.iters = iter; S.status = opt ? 1 : 2;
            }
        }
        if (S.done) atomicAdd(ndone, 1);
        doneflags[b] = S.done;
    }
}

// ---- compaction of finished problems ----
// The lock-step loop launches every batched kernel over the first `Bact` slots.  When problems finish, each finished
// slot below the new active count trades places with an active slot from the tail: everything a problem owns between
// iterations (P, G, its 17 vectors, its scalars; K / inv / info are rebuilt every iteration) is swapped, so the active
// problems stay a contiguous prefix and finished ones keep their final iterates in the tail.  ~6.3 MB per swap at
// n=512, m=1024, at most one swap per problem per solve.
struct SwapArgs {
    double *P, *G, *vecs; Scal *sc;
    long long sP, sG;
    int n, me, Btot;
};
__global__ void k_swap_slots(SwapArgs a, const int *pairs) {
    const int i = pairs[2 * blockIdx.y], j = pairs[2 * blockIdx.y + 1];
    const long long eP = a.sP, eG = a.sG, eN = 4LL * a.n, eM = 13LL * a.me, eS = (long long)(sizeof(Scal) / sizeof(double));
    const long long total = eP + eG + eN + eM + eS;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        double *x, *y;
        long long r = e;
        if (r < eP) { x = a.P + i * a.sP + r; y = a.P + j * a.sP + r; }
        else if ((r -= eP) < eG) { x = a.G + i * a.sG + r; y = a.G + j * a.sG + r; }
        else if ((r -= eG) < eN) {
            const long long arr = r / a.n, k = r % a.n;
            double *base = a.vecs + arr * (long long)a.Btot * a.n;
            x = base + (long long)i * a.n + k; y = base + (long long)j * a.n + k;
        } else if ((r -= eN) < eM) {
            const long long arr = r / a.me, k = r % a.me;
            double *base = a.vecs + 4LL * a.Btot * a.n + arr * (long long)a.Btot * a.me;
            x = base + (long long)i * a.me + k; y = base + (long long)j * a.me + k;
        } else {
            r -= eM;
            x = reinterpret_cast<double *>(a.sc + i) + r; y = reinterpret_cast<double *>(a.sc + j) + r;
        }
        const double t = *x; *x = *y; *y = t;
    }
}
// out[perm[slot], :] = in[slot, :]
__global__ void k_unpermute_rows(const double *in, double *out, const int *perm, int len) {
    const int slot = blockIdx.x;
    const double *src = in + (long long)slot * len;
    double *dst = out + (long long)perm[slot] * len;
    for (int k = threadIdx.x; k < len; k += blockDim.x) dst[k] = src[k];
}
static_assert(sizeof(Scal) % sizeof(double) == 0, "Scal is swapped as doubles");
// NT scaling at iteration 0 (misc.py:284-287) and lambda^2 (:2244)
__global__ void k_scaling(Ptrs p, int first) {
    PB_SETUP
    (void)sh; (void)on;
    if (S.done) return;
    for (int i = tid; i < p.m; i += nt) {
        if (first) {
            const double sv = p.s[om + i], zv = p.z[om + i];
            const double d = sqrt(sv / zv);
            p.d[om + i] = d;
            const double di = 1.0 / d;
            p.di[om + i] = di;
            p.di2[om + i] = di * di;
            p.lmbda[om + i] = sqrt(sv * zv);
        }
        const double l = p.lmbda[om + i];
        p.lmbdasq[om + i] = l * l;
    }
    if (tid == 0) { S.mu = S.gap / p.m; S.sigma = 0.0; S.eta = 0.0; }       // :2357-2358
}
// right-hand side of the i-th Newton system and the f4_no_ir preamble (:2376-2309)
__global__ void k_dir_prep(Ptrs p, int i) {
    PB_SETUP
    (void)sh;
    const double sm = S.sigma * S.mu, c = -1.0 + S.eta;
    for (int k = tid; k < p.n; k += nt) p.dx[on + k] = c * p.rx[on + k];
    for (int k = tid; k < p.m; k += nt) {
        double ds = -p.lmbdasq[om + k] + sm;
        if (i == 1) ds -= p.ws3[om + k];                 // Mehrotra correction
        ds = ds / p.lmbda[om + k];                       // sinv
        p.ds[om + k] = ds;
        const double dz = c * p.rz[om + k] - p.d[om + k] * ds;   // z := z - W' s
        p.dz[om + k] = dz;
        p.bzp[om + k] = p.di[om + k] * dz;               // W^{-T} bz for the solve
    }
}
// after the solve: dz = bzp (= W uz); ds := ds - dz; step length, sigma (:2316, :2423-2456)
__global__ void k_dir_post(Ptrs p, int i) {
    PB_SETUP
    double dsdz = 0, mins = INFINITY, minz = INFINITY;
    for (int k = tid; k < p.m; k += nt) {
        const double dz = p.bzp[om + k];
        const double ds = p.ds[om + k] - dz;
        dsdz += ds * dz;
        if (i == 0) p.ws3[om + k] = ds * dz;
        const double l = p.lmbda[om + k];
        const double dss = ds / l, dzs = dz / l;        // scale2
        p.ds[om + k] = dss; p.dz[om + k] = dzs;
        mins = fmin(mins, dss); minz = fmin(minz, dzs);
    }
    dsdz = block_sum(dsdz, sh);
    mins = block_min(mins, sh);
    minz = block_min(minz, sh);
    if (tid == 0) {
        const double t = fmax(0.0, fmax(-mins, -minz));
        double step;
        if (t == 0.0) step = 1.0;
        else step = (i == 0) ? fmin(1.0, 1.0 / t) : fmin(1.0, 0.99 / t);
        S.step = step; S.dsdz = dsdz;
        if (i == 0) {
            const double v = fmin(1.0, fmax(0.0, 1.0 - step + dsdz / S.gap * step * step));
            S.sigma = v * v * v;
            S.eta = 0.0;
        }
    }
}
// x += step dx; new scaled iterates, scaling update, unscaled s, z, gap (:2459-2547, misc.py:450-464)
__global__ void k_update(Ptrs p, const int *info, int iter) {
    PB_SETUP
    if (S.done) return;
    if (info[b] > 0) {      // non-positive pivot: "Terminated (singular KKT matrix)" (:2257-2275)
        if (tid == 0) { S.done = 1; S.status = 3; S.iters = iter; }
        return;
    }
    const double step = S.step;
    for (int k = tid; k < p.n; k += nt) p.x[on + k] += step * p.dx[on + k];
    double gap = 0;
    for (int k = tid; k < p.m; k += nt) {
        const double l = p.lmbda[om + k];
        const double ds = (1.0 + step * p.ds[om + k]) * l;      // scale2 inverse
        const double dz = (1.0 + step * p.dz[om + k]) * l;
        const double ss = sqrt(ds), sz = sqrt(dz);
        const double d = p.d[om + k] * ss / sz;
        const double di = 1.0 / d;
        const double ln = ss * sz;
        p.d[om + k] = d; p.di[om + k] = di; p.di2[om + k] = di * di;
        p.lmbda[om + k] = ln;
        p.s[om + k] = d * ln;                                   // W' lambda
        p.z[om + k] = di * ln;                                  // W^{-1} lambda
        gap += ln * ln;
    }
    gap = block_sum(gap, sh);
    if (tid == 0) S.gap = gap;
}

}  // namespace

struct cvxb_batch {
    int device = 0, B = 0, n = 0, m = 0;
    long long ldg = 0, ldp = 0, ldk = 0;
    long long sG = 0, sP = 0, sK = 0, sInv = 0;
    int nblk = 0;
    double *P = nullptr, *G = nullptr, *q = nullptr, *h = nullptr;
    double *K = nullptr, *inv = nullptr, *panel = nullptr, *gemv_ws = nullptr;
    double *vecs = nullptr;          // all n- and m-vectors
    Ptrs p;
    Scal *sc = nullptr;
    int *d_info = nullptr, *d_ndone = nullptr;
    int *d_done = nullptr, *d_pairs = nullptr, *d_perm = nullptr;     // compaction: done flags, swap list, slot -> problem
    std::vector<int> perm;           // slot -> original problem index (identity unless the last solve compacted)
    bool permuted = false;
    int compact = 1;                 // CVXB_BATCH_COMPACT=0 disables
    int Bact = 0;                    // slots [0, Bact) are launched by the lock-step loop
    CholWork cw;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool loaded = false;
    int iters_run = 0;
    double solve_ms = 0;
    // single large problem: SYRK on the int8 tensor path (ozaki_syrk.cu), same rule as cvxb_kkt_factor:
    // mode 1 (default) when B == 1, n >= 4096, m >= 8192; 2: whenever B == 1; 0: never.  CVXB_OZAKI (or the
    // older CVXB_OZAKI_IPM) = 0/1/2 read at create.
    int i8_mode = 1;
    int syrk_path = 0;
    void *oz_work = nullptr;
    size_t oz_bytes = 0;
};

namespace {

int batch_factor(cvxb_batch *b) {
    cudaStream_t st = b->st;
    bool i8 = b->B == 1 && b->m > 0 && (b->i8_mode == 2 || (b->i8_mode == 1 && b->n >= 4096 && b->m >= 8192));
    if (i8) {
        // K = P + G' diag(di)^2 G from nine int8 slices per entry (fp64-accurate, ~1.8x the DMMA SYRK);
        // same size rule and same fallback (workspace does not fit -> DMMA kernel) as cvxb_kkt_factor
        const size_t need = ozaki_workspace_bytes(b->n, b->m, 9);
        if (need > b->oz_bytes) {
            if (b->oz_work) cudaFree(b->oz_work);
            b->oz_work = nullptr; b->oz_bytes = 0;
            cudaError_t ae = cudaMalloc(&b->oz_work, need);
            if (ae == cudaErrorMemoryAllocation) { cudaGetLastError(); tmp_cache_release(); ae = cudaMalloc(&b->oz_work, need); }
            if (ae != cudaSuccess) {
                cudaGetLastError();
                b->oz_work = nullptr;
                i8 = false;
            } else {
                b->oz_bytes = need;
            }
        }
    }
    b->syrk_path = i8 ? 2 : 1;
    if (i8) {
        CVXB_TRY(ozaki_syrk(b->n, b->m, b->G, b->ldg, b->p.di, b->P, b->ldp, 1.0, b->K, b->ldk, 9, 0,
                            b->oz_work, nullptr, st));
        CVXB_TRY(potrf_lower(b->n, b->K, (int)b->ldk, b->inv, b->cw, st));
        CVXB_CUDA(cudaMemcpyAsync(b->d_info, b->cw.d_info, sizeof(int), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    GemmDesc g;
    g.M = b->n; g.N = b->n; g.K = b->m;
    g.X = b->G; g.ldx = (int)b->ldg; g.x_kmajor = true; g.sX = b->sG;
    g.Y = b->G; g.ldy = (int)b->ldg; g.y_kmajor = true; g.sY = b->sG;
    g.w = b->p.di2; g.sW = b->m;
    g.D = b->P; g.ldd = (int)b->ldp; g.sD = b->sP; g.beta = 1.0;
    g.C = b->K; g.ldc = (int)b->ldk; g.sC = b->sK;
    g.lower_only = true; g.batch = b->Bact;
    if (b->B == 1) g.splitk_ws = b->cw.splitk_ws;
    CVXB_TRY(dmma_gemm(g, st));
    if (b->B == 1) {
        CVXB_TRY(potrf_lower(b->n, b->K, (int)b->ldk, b->inv, b->cw, st));
        CVXB_CUDA(cudaMemcpyAsync(b->d_info, b->cw.d_info, sizeof(int), cudaMemcpyDeviceToDevice, st));
    } else {
        CVXB_TRY(potrf_lower_batched(b->n, b->K, (int)b->ldk, b->sK, b->inv, b->sInv, b->Bact, b->d_info,
                                     b->panel, (b->n + 1) & ~1, st));
    }
    return 0;
}

// (dx, bzp) := solution of the reduced KKT system; on entry dx = bx, bzp = W^{-T} bz
int batch_solve(cvxb_batch *b) {
    cudaStream_t st = b->st;
    const int n = b->n, m = b->m, B = b->Bact;
    GemvBatch gt; gt.batch = B; gt.sA = b->sG; gt.sw = m; gt.sx = m; gt.sy = n;
    // x := x + G' (di .* bzp)
    CVXB_TRY(gemv_t(m, n, b->G, b->ldg, b->p.di, b->p.bzp, 1.0, 1.0, b->p.dx, st, gt));
    CVXB_TRY(potrs_lower(n, b->K, (int)b->ldk, b->inv, b->p.dx, b->cw, st, B, b->sK, b->sInv, n));
    // bzp := di .* (G x) - bzp
    GemvBatch gn; gn.batch = B; gn.sA = b->sG; gn.sw = m; gn.sx = n; gn.sy = m;
    CVXB_TRY(gemv_n(m, n, b->G, b->ldg, b->p.di, b->p.dx, 1.0, -1.0, b->p.bzp, b->gemv_ws, st, gn));
    return 0;
}

}  // namespace

extern "C" {

int cvxb_batch_create(cvxb_batch **out, int nprob, int n, int m, int device) {
    if (!out || nprob <= 0 || n <= 0 || m < 0) { set_error("batch_create: bad sizes"); return CVXB_E_ARG; }
    *out = nullptr;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: cvxopt_b200 has no CPU fallback");
        return CVXB_E_NOGPU;
    }
    if (device < 0 || device >= cnt) { set_error("device out of range"); return CVXB_E_ARG; }
    CVXB_CUDA(cudaSetDevice(device));
    cvxb_batch *b = new cvxb_batch();
    b->device = device; b->B = nprob; b->n = n; b->m = m;
    if (const char *e = getenv("CVXB_OZAKI")) b->i8_mode = (e[0] == '0') ? 0 : (e[0] == '2') ? 2 : 1;
    if (const char *e = getenv("CVXB_OZAKI_IPM")) b->i8_mode = (e[0] == '1') ? 1 : (e[0] == '2') ? 2 : 0;
    b->ldg = ((m + 1) & ~1) > 2 ? ((m + 1) & ~1) : 2;
    b->ldp = b->ldk = (n + 1) & ~1;
    b->sG = b->ldg * n; b->sP = b->ldp * n; b->sK = b->ldk * n;
    b->nblk = (n + NB - 1) / NB;
    b->sInv = (long long)2 * b->nblk * NB * NB;
    auto fail = [&](int r) { cvxb_batch_destroy(b); return r; };
#define BCUDA(expr) do { cudaError_t _e = (expr); \
        /* out of memory: give the scratch-buffer cache (common.cuh) back to the driver and try once more */ \
        if (_e == cudaErrorMemoryAllocation) { cudaGetLastError(); tmp_cache_release(); _e = (expr); } \
        if (_e != cudaSuccess) { \
        set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
        return fail(_e == cudaErrorMemoryAllocation ? CVXB_E_NOMEM : CVXB_E_CUDA); } } while (0)
    const size_t B = nprob;
    BCUDA(cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking));
    BCUDA(cudaEventCreate(&b->e0)); BCUDA(cudaEventCreate(&b->e1));
    { int r = chol_work_create(b->cw); if (r) return fail(r); }
    BCUDA(cudaMalloc(&b->P, B * b->sP * sizeof(double)));
    BCUDA(cudaMalloc(&b->G, B * b->sG * sizeof(double)));
    BCUDA(cudaMalloc(&b->K, B * b->sK * sizeof(double)));
    BCUDA(cudaMalloc(&b->inv, B * b->sInv * sizeof(double)));
    BCUDA(cudaMalloc(&b->panel, B * (size_t)((n + 1) & ~1) * NB * sizeof(double)));
    BCUDA(cudaMalloc(&b->gemv_ws, B * (size_t)(m > 0 ? m : 1) * gemv_n_chunks(n) * sizeof(double)));
    // vectors: n-sized: q x rx dx ; m-sized: h s z rz ds dz lmbda lmbdasq d di di2 ws3 bzp
    const size_t nv = 4, mv = 13;
    const size_t me = (size_t)(m > 0 ? m : 1);
    BCUDA(cudaMalloc(&b->vecs, B * (nv * n + mv * me) * sizeof(double)));
    BCUDA(cudaMemset(b->vecs, 0, B * (nv * n + mv * me) * sizeof(double)));
    double *v = b->vecs;
    auto take = [&](size_t len) { double *r = v; v += B * len; return r; };
    b->q = take(n); b->p.x = take(n); b->p.rx = take(n); b->p.dx = take(n);
    b->h = take(me); b->p.s = take(me); b->p.z = take(me); b->p.rz = take(me); b->p.ds = take(me);
    b->p.dz = take(me); b->p.lmbda = take(me); b->p.lmbdasq = take(me); b->p.d = take(me);
    b->p.di = take(me); b->p.di2 = take(me); b->p.ws3 = take(me); b->p.bzp = take(me);
    b->p.q = b->q; b->p.h = b->h; b->p.n = n; b->p.m = m;
    BCUDA(cudaMalloc(&b->sc, B * sizeof(Scal)));
    BCUDA(cudaMemset(b->sc, 0, B * sizeof(Scal)));
    b->p.sc = b->sc;
    BCUDA(cudaMalloc(&b->d_info, B * sizeof(int)));
    BCUDA(cudaMalloc(&b->d_ndone, sizeof(int)));
    BCUDA(cudaMalloc(&b->d_done, B * sizeof(int)));
    BCUDA(cudaMalloc(&b->d_pairs, 2 * B * sizeof(int)));
    BCUDA(cudaMalloc(&b->d_perm, B * sizeof(int)));
    b->perm.resize(B);
    for (size_t i = 0; i < B; ++i) b->perm[i] = (int)i;
    if (const char *e = getenv("CVXB_BATCH_COMPACT")) b->compact = (e[0] == '0') ? 0 : 1;
#undef BCUDA
    *out = b;
    return 0;
}

void cvxb_batch_destroy(cvxb_batch *b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->st) cudaStreamSynchronize(b->st);
    double *bufs[] = {b->P, b->G, b->K, b->inv, b->panel, b->gemv_ws, b->vecs};
    for (double *x : bufs) if (x) cudaFree(x);
    if (b->sc) cudaFree(b->sc);
    if (b->oz_work) cudaFree(b->oz_work);
    if (b->d_info) cudaFree(b->d_info);
    if (b->d_ndone) cudaFree(b->d_ndone);
    if (b->d_done) cudaFree(b->d_done);
    if (b->d_pairs) cudaFree(b->d_pairs);
    if (b->d_perm) cudaFree(b->d_perm);
    chol_work_destroy(b->cw);
    if (b->e0) cudaEventDestroy(b->e0);
    if (b->e1) cudaEventDestroy(b->e1);
    if (b->st) cudaStreamDestroy(b->st);
    delete b;
}

// P: nprob x (n x n, ld n) ; q: nprob x n ; G: nprob x (m x n column-major, ld m) ; h: nprob x m
int cvxb_batch_load(cvxb_batch *b, const double *P, const double *q, const double *G,
                    const double *h, int space) {
    if (!b || !P || !q || (b->m > 0 && (!G || !h))) { set_error("batch_load: NULL argument"); return CVXB_E_ARG; }
    CVXB_CUDA(cudaSetDevice(b->device));
    const cudaMemcpyKind kind = (space == CVXB_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t B = b->B, n = b->n, m = b->m;
    // one strided 2-D copy per operand: rows of the "matrix of columns" are the matrix columns
    CVXB_CUDA(cudaMemcpy2DAsync(b->P, b->ldp * sizeof(double), P, n * sizeof(double), n * sizeof(double),
                                n * B, kind, b->st));
    if (m > 0) {
        CVXB_CUDA(cudaMemcpy2DAsync(b->G, b->ldg * sizeof(double), G, m * sizeof(double),
                                    m * sizeof(double), n * B, kind, b->st));
        CVXB_CUDA(cudaMemcpyAsync(const_cast<double *>(b->h), h, B * m * sizeof(double), kind, b->st));
    }
    CVXB_CUDA(cudaMemcpyAsync(const_cast<double *>(b->q), q, B * n * sizeof(double), kind, b->st));
    // only tril(P) is significant in the reference; make the resident copies symmetric
    CVXB_TRY(symmetrize_lower(b->n, b->P, b->ldp, b->B, b->sP, b->st));
    CVXB_CUDA(cudaStreamSynchronize(b->st));
    b->loaded = true;
    for (size_t i = 0; i < B; ++i) b->perm[i] = (int)i;
    b->permuted = false;
    return 0;
}

// swap the slots of each pair (disjoint pairs: one launch)
static int swap_slots(cvxb_batch *b, const std::vector<int> &pairs) {
    const int np = (int)pairs.size() / 2;
    if (np == 0) return 0;
    CVXB_CUDA(cudaMemcpyAsync(b->d_pairs, pairs.data(), pairs.size() * sizeof(int), cudaMemcpyHostToDevice, b->st));
    SwapArgs a;
    a.P = b->P; a.G = b->G; a.vecs = b->vecs; a.sc = b->sc; a.sP = b->sP; a.sG = b->sG;
    a.n = b->n; a.me = b->m > 0 ? b->m : 1; a.Btot = b->B;
    k_swap_slots<<<dim3(96, np), 256, 0, b->st>>>(a, b->d_pairs);
    count_launch();
    // `pairs` is pageable host memory: the copy above is staged before cudaMemcpyAsync returns
    return 0;
}

// put every problem back into its own slot (a solve that compacted left them permuted)
static int restore_order(cvxb_batch *b) {
    if (!b->permuted) return 0;
    std::vector<int> pr(2);
    for (int i = 0; i < b->B; ++i) {
        while (b->perm[i] != i) {
            const int j = b->perm[i];                // the problem in slot i belongs to slot j
            pr[0] = i; pr[1] = j;
            CVXB_TRY(swap_slots(b, pr));
            std::swap(b->perm[i], b->perm[j]);
        }
    }
    CVXB_CUDA(cudaStreamSynchronize(b->st));
    b->permuted = false;
    return 0;
}

int cvxb_batch_solve(cvxb_batch *b, int maxiters, double abstol, double reltol, double feastol) {
    if (!b || !b->loaded) { set_error("batch_solve: load the problems first"); return CVXB_E_ARG; }
    CVXB_CUDA(cudaSetDevice(b->device));
    cudaStream_t st = b->st;
    CVXB_TRY(restore_order(b));
    const int n = b->n, m = b->m, T = 256;
    int B = b->B;                                 // active slots: shrinks as problems finish (compaction)
    b->Bact = B;
    Ptrs &p = b->p;
    GemvBatch gP; gP.batch = B; gP.sA = b->sP; gP.sx = n; gP.sy = n;
    GemvBatch gGt; gGt.batch = B; gGt.sA = b->sG; gGt.sx = m; gGt.sy = n;
    GemvBatch gGn; gGn.batch = B; gGn.sA = b->sG; gGn.sx = n; gGn.sy = m;
    CVXB_CUDA(cudaMemsetAsync(b->sc, 0, (size_t)B * sizeof(Scal), st));
    CVXB_CUDA(cudaEventRecord(b->e0, st));
    // ---- starting point: W = I ----
    k_init_rhs<<<B, T, 0, st>>>(p); count_launch();
    CVXB_TRY(batch_factor(b));
    k_scale_bz<<<B, T, 0, st>>>(p, p.dz); count_launch();
    CVXB_TRY(batch_solve(b));
    k_init_point<<<B, T, 0, st>>>(p); count_launch();
    CVXB_LAUNCH_CHECK();
    int info_fail = 0;
    {
        // a singular first factorisation is the reference's "Rank([P; G]) < n" ValueError
        std::vector<int> info(B);
        CVXB_CUDA(cudaMemcpyAsync(info.data(), b->d_info, B * sizeof(int), cudaMemcpyDeviceToHost, st));
        CVXB_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < B; ++i) if (info[i] > 0) { info_fail = i + 1; break; }
        if (info_fail) {
            set_error("batch_solve: problem %d: Rank([P; G]) < n (singular KKT matrix at the start)", info_fail - 1);
            return CVXB_E_ARG;
        }
    }
    std::vector<int> flags(B), pairs;
    int it = 0;
    for (it = 0; it <= maxiters; ++it) {
        // residuals (:2169-2186)
        k_res_begin<<<B, T, 0, st>>>(p); count_launch();
        CVXB_TRY(gemv_t(n, n, b->P, b->ldp, nullptr, p.x, 1.0, 1.0, p.rx, st, gP));
        k_res_dots<<<B, T, 0, st>>>(p); count_launch();
        if (m > 0) {
            CVXB_TRY(gemv_t(m, n, b->G, b->ldg, nullptr, p.z, 1.0, 1.0, p.rx, st, gGt));
            CVXB_TRY(gemv_n(m, n, b->G, b->ldg, nullptr, p.x, 1.0, 1.0, p.rz, b->gemv_ws, st, gGn));
        }
        CVXB_CUDA(cudaMemsetAsync(b->d_ndone, 0, sizeof(int), st));
        k_stats<<<B, T, 0, st>>>(p, it, maxiters, abstol, reltol, feastol, b->d_ndone, b->d_done); count_launch();
        int ndone = 0;
        CVXB_CUDA(cudaMemcpyAsync(&ndone, b->d_ndone, sizeof(int), cudaMemcpyDeviceToHost, st));
        CVXB_CUDA(cudaMemcpyAsync(flags.data(), b->d_done, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
        CVXB_CUDA(cudaStreamSynchronize(st));
        if (ndone >= B) break;
        if (ndone > 0 && b->compact && b->B > 1) {
            // finished slots below the new active count trade places with active slots from the tail
            const int nb = B - ndone;
            pairs.clear();
            int j = B - 1;
            for (int i = 0; i < nb; ++i) {
                if (!flags[i]) continue;
                while (flags[j]) --j;             // an active slot in [nb, B): there are as many as finished ones below nb
                pairs.push_back(i); pairs.push_back(j);
                std::swap(b->perm[i], b->perm[j]);
                --j;
            }
            CVXB_TRY(swap_slots(b, pairs));
            b->permuted = true;
            B = nb;
            b->Bact = B;
            gP.batch = gGt.batch = gGn.batch = B;
        }
        k_scaling<<<B, T, 0, st>>>(p, it == 0 ? 1 : 0); count_launch();
        CVXB_TRY(batch_factor(b));
        for (int i = 0; i < 2; ++i) {
            k_dir_prep<<<B, T, 0, st>>>(p, i); count_launch();
            CVXB_TRY(batch_solve(b));
            k_dir_post<<<B, T, 0, st>>>(p, i); count_launch();
        }
        k_update<<<B, T, 0, st>>>(p, b->d_info, it); count_launch();
        CVXB_LAUNCH_CHECK();
    }
    b->iters_run = it;
    b->Bact = b->B;
    CVXB_CUDA(cudaEventRecord(b->e1, st));
    CVXB_CUDA(cudaStreamSynchronize(st));
    float t = 0;
    cudaEventElapsedTime(&t, b->e0, b->e1);
    b->solve_ms = t;
    return 0;
}

int cvxb_batch_results(cvxb_batch *b, double *x, double *s, double *z, int *status, int *iters,
                       double *pobj, double *dobj, int space) {
    if (!b) { set_error("batch is NULL"); return CVXB_E_ARG; }
    CVXB_CUDA(cudaSetDevice(b->device));
    const cudaMemcpyKind kind = (space == CVXB_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const size_t B = b->B;
    // slot -> problem (identity unless the solve compacted finished problems away)
    if (b->permuted) CVXB_CUDA(cudaMemcpy(b->d_perm, b->perm.data(), B * sizeof(int), cudaMemcpyHostToDevice));
    auto give = [&](double *dst, const double *src, int len) -> int {
        if (!b->permuted) { CVXB_CUDA(cudaMemcpy(dst, src, B * len * sizeof(double), kind)); return 0; }
        double *tmp = (space == CVXB_DEVICE) ? dst : nullptr;
        if (!tmp) CVXB_CUDA(tmp_malloc(&tmp, B * len * sizeof(double)));
        k_unpermute_rows<<<(unsigned)B, 256, 0, b->st>>>(src, tmp, b->d_perm, len);
        count_launch();
        cudaError_t e = cudaStreamSynchronize(b->st);
        if (e == cudaSuccess && tmp != dst) e = cudaMemcpy(dst, tmp, B * len * sizeof(double), kind);
        if (tmp != dst) tmp_free(tmp);
        CVXB_CUDA(e);
        return 0;
    };
    if (x) CVXB_TRY(give(x, b->p.x, b->n));
    if (s && b->m) CVXB_TRY(give(s, b->p.s, b->m));
    if (z && b->m) CVXB_TRY(give(z, b->p.z, b->m));
    if (status || iters || pobj || dobj) {
        if (space == CVXB_DEVICE) { set_error("batch_results: scalars are returned to host memory only"); return CVXB_E_ARG; }
        std::vector<Scal> sc(B);
        CVXB_CUDA(cudaMemcpy(sc.data(), b->sc, B * sizeof(Scal), cudaMemcpyDeviceToHost));
        for (size_t slot = 0; slot < B; ++slot) {
            const size_t i = (size_t)b->perm[slot];
            if (status) status[i] = sc[slot].status;
            if (iters) iters[i] = sc[slot].iters;
            if (pobj) pobj[i] = sc[slot].pcost;
            if (dobj) dobj[i] = sc[slot].dcost;
        }
    }
    return 0;
}

int cvxb_batch_syrk_path(cvxb_batch *b) { return b ? b->syrk_path : CVXB_E_ARG; }

int cvxb_batch_stats(cvxb_batch *b, double *solve_ms, int *iterations) {
    if (!b) return CVXB_E_ARG;
    if (solve_ms) *solve_ms = b->solve_ms;
    if (iterations) *iterations = b->iters_run;
    return 0;
}

}  // extern "C"
