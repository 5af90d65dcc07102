// fp64 tensor-core GEMM/SYRK for sm_100a.
//
//   C[r,c] = alpha * sum_k X[r,k] * w[k] * Y[c,k] + beta * D[r,c]
//
// This one kernel is the flop carrier of the whole KKT path:
//   * normal-equations assembly  K = H + G' diag(di^2) G        (X=Y=G, K-major, w=di^2,
//     lower tiles only; replaces scale(Gs)+blas.syrk+`K += H`,   reference misc.py:1268-1276)
//   * Cholesky panel TRSM  L21 = A21 * L11^{-T}                  (X=A21, Y=inv(L11), M-major)
//   * Cholesky trailing update A22 -= L21 L21'                   (X=Y=L21, M-major, lower)
//   * the 's'-cone congruences r' X r                            (general GEMMs)
//
// B200 has no tcgen05 kind for fp64 (ptxas rejects kind::f64), so the fp64 tensor
// path is warp-level DMMA.8x8x4 (mma.sync.m8n8k4.f64).  Measured pipe peak on this
// pool: 37.2 TF/s = 64 FMA/clk/SM * 148 SMs * 1.965 GHz (tools/fp64_peak.cu).  At
// 64 FMA/clk a 128x128x16 tile step keeps an SM busy for 4096 cycles while moving
// only 32 KB, so operand traffic is trivial; the design goal is simply to keep the
// DMMA pipe issuing: 8 warps, 64x32 warp tiles (64 accumulator doubles / thread,
// 32 independent DMMAs between dependent ones), 4-stage cp.async pipeline, padded
// shared-memory layouts that make every fragment LDS.64 bank-conflict free.
//
// MMA roles are swapped w.r.t. the matrix: the MMA "m" index runs over C's columns
// (Y operand), the "n" index over C's rows (X operand), so each thread's accumulator
// pair is two consecutive ROWS of a column-major C (contiguous).
#include "common.cuh"

namespace cvxb {

namespace {

constexpr int BR = 128, BC = 128, BK = 16, STAGES = 4;
constexpr int SK = BK + 4;       // row stride (doubles) of a K-major operand tile  [idx][k]
constexpr int SMJ = BR + 4;      // row stride (doubles) of an M-major operand tile [k][idx]
constexpr int OPER_STAGE = 128 * SK;  // 2560 doubles >= 16*132
constexpr int STAGE_DOUBLES = 2 * OPER_STAGE + BK;
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;
constexpr int TILE_ELEMS = BR * BC;
constexpr int THREADS = 256;

struct KParams {
    int M, N, K;
    const double *X; long long ldx;
    const double *Y; long long ldy;
    const double *w;
    const double *D; long long ldd;
    double *C; long long ldc;
    double alpha, beta;
    int lower_only;
    int ct_begin, ct_end;   // c-tile window (already clipped to the tile grid)
    int nTr;                // number of r tiles
    long long sX, sY, sW, sD, sC;
    // split-K
    int full_tiles;         // units [0, full_tiles) are whole tiles
    int S;                  // splits per remainder tile (1 = none)
    int kchunk;             // K elements per split (multiple of BK)
    double *ws;
};

__device__ __forceinline__ void decode_tile(const KParams &p, int t, int &tr, int &tc) {
    // tiles are enumerated c-tile by c-tile (column-major over the tile grid)
    int c = p.ct_begin;
    if (p.lower_only) {
        while (true) {
            int cnt = p.nTr - c;
            if (t < cnt) break;
            t -= cnt;
            ++c;
        }
        tr = c + t;
        tc = c;
    } else {
        tc = c + t / p.nTr;
        tr = t % p.nTr;
    }
}

// ---- operand tile loaders ----------------------------------------------------
// K-major source: element (idx, k) at src[k + idx*ld]; smem [idx*SK + k]
template <bool VEC>
__device__ __forceinline__ void load_kmajor(double *s, const double *src, long long ld,
                                            int nrows, int kvalid, int tid) {
    if (VEC) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int q = tid + i * THREADS;
            int idx = q >> 3, kc = (q & 7) * 2;
            int rem = kvalid - kc;
            int bytes = (idx < nrows) ? (rem >= 2 ? 16 : (rem == 1 ? 8 : 0)) : 0;
            const double *g = bytes ? (src + kc + (long long)idx * ld) : src;
            cp_async16(s + idx * SK + kc, g, bytes);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int q = tid + i * THREADS;
            int idx = q >> 4, kc = q & 15;
            int bytes = (idx < nrows && kc < kvalid) ? 8 : 0;
            const double *g = bytes ? (src + kc + (long long)idx * ld) : src;
            cp_async8(s + idx * SK + kc, g, bytes);
        }
    }
}
// M-major source: element (idx, k) at src[idx + k*ld]; smem [k*SMJ + idx]
template <bool VEC>
__device__ __forceinline__ void load_mmajor(double *s, const double *src, long long ld,
                                            int nrows, int kvalid, int tid) {
    if (VEC) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int q = tid + i * THREADS;
            int k = q >> 6, ic = (q & 63) * 2;
            int rem = nrows - ic;
            int bytes = (k < kvalid) ? (rem >= 2 ? 16 : (rem == 1 ? 8 : 0)) : 0;
            const double *g = bytes ? (src + ic + (long long)k * ld) : src;
            cp_async16(s + k * SMJ + ic, g, bytes);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int q = tid + i * THREADS;
            int k = q >> 7, ic = q & 127;
            int bytes = (k < kvalid && ic < nrows) ? 8 : 0;
            const double *g = bytes ? (src + ic + (long long)k * ld) : src;
            cp_async8(s + k * SMJ + ic, g, bytes);
        }
    }
}

template <bool XK, bool YK, bool VEC>
__global__ void __launch_bounds__(THREADS, 1) dmma_gemm_kernel(const KParams p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp & 1, wc = warp >> 1;
    const int g4 = lane >> 2, t4 = lane & 3;

    // ---- which unit am I? ----
    const int u = blockIdx.x;
    int tile, split = 0;
    if (u < p.full_tiles) {
        tile = u;
    } else {
        tile = p.full_tiles + (u - p.full_tiles) / p.S;
        split = (u - p.full_tiles) % p.S;
    }
    int tr, tc;
    decode_tile(p, tile, tr, tc);
    const bool is_split = (u >= p.full_tiles) && (p.S > 1);
    int kbeg = 0, kend = p.K;
    if (is_split) {
        kbeg = split * p.kchunk;
        kend = min(p.K, kbeg + p.kchunk);
    }
    const int r0 = tr * BR, c0 = tc * BC;
    const int nr = min(BR, p.M - r0), nc = min(BC, p.N - c0);

    const long long b = blockIdx.z;
    const double *X = p.X + b * p.sX;
    const double *Y = p.Y + b * p.sY;
    const double *w = p.w ? p.w + b * p.sW : nullptr;

    const double *Xt = XK ? X + (long long)r0 * p.ldx : X + r0;
    const double *Yt = YK ? Y + (long long)c0 * p.ldy : Y + c0;

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int ktiles = (kend - kbeg + BK - 1) / BK;

    auto load_stage = [&](int kt, int stage) {
        double *sx = smem + stage * STAGE_DOUBLES;
        double *sy = sx + OPER_STAGE;
        double *sw = sy + OPER_STAGE;
        const int k0 = kbeg + kt * BK;
        const int kvalid = min(BK, kend - k0);
        if (XK) load_kmajor<VEC>(sx, Xt + k0, p.ldx, nr, kvalid, tid);
        else    load_mmajor<VEC>(sx, Xt + (long long)k0 * p.ldx, p.ldx, nr, kvalid, tid);
        if (YK) load_kmajor<VEC>(sy, Yt + k0, p.ldy, nc, kvalid, tid);
        else    load_mmajor<VEC>(sy, Yt + (long long)k0 * p.ldy, p.ldy, nc, kvalid, tid);
        if (w != nullptr && tid < BK) {
            int bytes = (tid < kvalid) ? 8 : 0;
            cp_async8(sw + tid, bytes ? (w + k0 + tid) : w, bytes);
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ktiles) load_stage(s, s);
        cp_async_commit();
    }

    const bool has_w = (w != nullptr);
    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < ktiles) load_stage(nk, nk % STAGES);
            cp_async_commit();
        }
        const double *sx = smem + (kt % STAGES) * STAGE_DOUBLES;
        const double *sy = sx + OPER_STAGE;
        const double *sw = sy + OPER_STAGE;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            const int k = kk * 4 + t4;
            double a[4], bf[8];
#pragma unroll
            for (int cf = 0; cf < 4; ++cf) {
                int idx = wc * 32 + cf * 8 + g4;
                a[cf] = YK ? sy[idx * SK + k] : sy[k * SMJ + idx];
            }
#pragma unroll
            for (int rf = 0; rf < 8; ++rf) {
                int idx = wr * 64 + rf * 8 + g4;
                bf[rf] = XK ? sx[idx * SK + k] : sx[k * SMJ + idx];
            }
            if (has_w) {
                double wv = sw[k];
#pragma unroll
                for (int cf = 0; cf < 4; ++cf) a[cf] *= wv;
            }
#pragma unroll
            for (int cf = 0; cf < 4; ++cf)
#pragma unroll
                for (int rf = 0; rf < 8; ++rf) dmma(acc[cf][rf][0], acc[cf][rf][1], a[cf], bf[rf]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ----
    if (is_split) {
        double *ws = p.ws + ((long long)(tile - p.full_tiles) * p.S + split) * TILE_ELEMS;
#pragma unroll
        for (int cf = 0; cf < 4; ++cf)
#pragma unroll
            for (int rf = 0; rf < 8; ++rf) {
                int rl = wr * 64 + rf * 8 + t4 * 2;
                int cl = wc * 32 + cf * 8 + g4;
                *reinterpret_cast<double2 *>(ws + rl + cl * BR) =
                    make_double2(acc[cf][rf][0], acc[cf][rf][1]);
            }
        return;
    }
    double *C = p.C + b * p.sC;
    const double *D = p.D ? p.D + b * p.sD : nullptr;
    const bool diag = p.lower_only && (tr == tc);
#pragma unroll
    for (int cf = 0; cf < 4; ++cf) {
        const int cl = wc * 32 + cf * 8 + g4;
        if (cl >= nc) continue;
        const long long c = c0 + cl;
#pragma unroll
        for (int rf = 0; rf < 8; ++rf) {
            const int rl = wr * 64 + rf * 8 + t4 * 2;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (rl + e >= nr) continue;
                if (diag && (rl + e) < cl) continue;
                const long long r = r0 + rl + e;
                double v = p.alpha * acc[cf][rf][e];
                if (p.beta != 0.0) v += p.beta * D[r + c * p.ldd];
                C[r + c * p.ldc] = v;
            }
        }
    }
}

// sums the S split-K partials of one remainder tile in a fixed order (deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const KParams p) {
    const int tile = p.full_tiles + blockIdx.x;
    int tr, tc;
    decode_tile(p, tile, tr, tc);
    const int r0 = tr * BR, c0 = tc * BC;
    const int nr = min(BR, p.M - r0), nc = min(BC, p.N - c0);
    const bool diag = p.lower_only && (tr == tc);
    const double *ws = p.ws + (long long)blockIdx.x * p.S * TILE_ELEMS;
    for (int e = threadIdx.x; e < TILE_ELEMS; e += blockDim.x) {
        int rl = e % BR, cl = e / BR;
        if (rl >= nr || cl >= nc) continue;
        if (diag && rl < cl) continue;
        double s = 0.0;
        for (int k = 0; k < p.S; ++k) s += ws[(long long)k * TILE_ELEMS + e];
        long long r = r0 + rl, c = c0 + cl;
        double v = p.alpha * s;
        if (p.beta != 0.0) v += p.beta * p.D[r + c * p.ldd];
        p.C[r + c * p.ldc] = v;
    }
}

template <bool XK, bool YK, bool VEC>
int launch_inst(const KParams &p, dim3 grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        CVXB_CUDA(cudaFuncSetAttribute(dmma_gemm_kernel<XK, YK, VEC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    dmma_gemm_kernel<XK, YK, VEC><<<grid, THREADS, SMEM_BYTES, st>>>(p);
    count_launch();
    CVXB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int dmma_gemm(const GemmDesc &g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return 0;
    if (g.K < 0 || !g.X || !g.Y || !g.C) {
        set_error("dmma_gemm: bad arguments");
        return CVXB_E_ARG;
    }
    KParams p;
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.X = g.X; p.ldx = g.ldx; p.Y = g.Y; p.ldy = g.ldy; p.w = g.w;
    p.D = g.D; p.ldd = g.ldd; p.C = g.C; p.ldc = g.ldc;
    p.alpha = g.alpha; p.beta = (g.D ? g.beta : 0.0);
    p.lower_only = g.lower_only ? 1 : 0;
    p.nTr = (g.M + BR - 1) / BR;
    const int nTc = (g.N + BC - 1) / BC;
    p.ct_begin = g.ct_begin < 0 ? 0 : g.ct_begin;
    p.ct_end = g.ct_end > nTc ? nTc : g.ct_end;
    if (p.ct_begin >= p.ct_end) return 0;
    p.sX = g.sX; p.sY = g.sY; p.sW = g.sW; p.sD = g.sD; p.sC = g.sC;
    long long T = 0;
    if (p.lower_only) {
        for (int c = p.ct_begin; c < p.ct_end; ++c) T += (p.nTr - c > 0 ? p.nTr - c : 0);
    } else {
        T = (long long)p.nTr * (p.ct_end - p.ct_begin);
    }
    if (T <= 0) return 0;
    // split-K of the remainder wave (deterministic: partials + ordered reduce)
    p.full_tiles = (int)T; p.S = 1; p.kchunk = g.K; p.ws = nullptr;
    if (g.splitk_ws && g.batch == 1 && g.K >= 1024) {
        int full = (int)(T / kNumSMs) * kNumSMs;
        int rem = (int)T - full;
        if (rem > 0) {
            int S = kNumSMs / rem;
            int maxS = g.K / 512;
            if (S > maxS) S = maxS;
            if (S >= 2) {
                p.full_tiles = full;
                p.S = S;
                int kc = (g.K + S - 1) / S;
                p.kchunk = ((kc + BK - 1) / BK) * BK;
                p.ws = g.splitk_ws;
            }
        }
    }
    const int rem_tiles = (int)T - p.full_tiles;
    const int units = p.full_tiles + rem_tiles * p.S;
    dim3 grid(units, 1, g.batch);
    auto aligned = [](const void *ptr, long long ld) {
        return ((uintptr_t)ptr % 16 == 0) && (ld % 2 == 0);
    };
    const bool vec = aligned(g.X, g.ldx) && aligned(g.Y, g.ldy) &&
                     (g.batch == 1 || (g.sX % 2 == 0 && g.sY % 2 == 0));
    int rc;
#define DISPATCH(XK, YK)                                                   \
    rc = vec ? launch_inst<XK, YK, true>(p, grid, st) : launch_inst<XK, YK, false>(p, grid, st)
    if (g.x_kmajor && g.y_kmajor) DISPATCH(true, true);
    else if (g.x_kmajor && !g.y_kmajor) DISPATCH(true, false);
    else if (!g.x_kmajor && g.y_kmajor) DISPATCH(false, true);
    else DISPATCH(false, false);
#undef DISPATCH
    if (rc) return rc;
    if (p.S > 1) {
        splitk_reduce_kernel<<<rem_tiles, 256, 0, st>>>(p);
        count_launch();
        CVXB_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace cvxb
