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
#include <cstdlib>

namespace cvxb {

namespace {

constexpr int PB = 8;                  // inner block width of the in-CTA factorisation
constexpr int SP = 12;                 // row stride (doubles) of the panel buffer: conflict-free frags
constexpr int LDM = NB + 4;            // column stride (doubles) of the shared-memory block
constexpr int POTF2_SMEM = (NB * LDM + NB * SP + 4 * 80) * 8;

// acc[cf][rf] -= sum_k Lt[c, k] Lt[r, k] over k = 0..127 for the tiles rf >= S_cf of each column
// fragment cf (Lt in shared memory, column-major with stride LDM; Ma/Mb already offset by lane)
template <int PAT>      // 0: every tile, 1: rf >= cf, 2: rf >= 4 + cf
__device__ __forceinline__ void sym_update(double (&acc)[4][8][2], const double *Ma, const double *Mb) {
#pragma unroll 2
    for (int kk = 0; kk < NB / 4; ++kk) {
        double a[4], bfr[8];
#pragma unroll
        for (int cf = 0; cf < 4; ++cf) a[cf] = -Ma[cf * 8 + kk * 4 * LDM];
#pragma unroll
        for (int rf = 0; rf < 8; ++rf) bfr[rf] = Mb[rf * 8 + kk * 4 * LDM];
#pragma unroll
        for (int cf = 0; cf < 4; ++cf)
#pragma unroll
            for (int rf = 0; rf < 8; ++rf) {
                constexpr bool dummy = true; (void)dummy;
                if (PAT == 1 && rf < cf) continue;           // compile-time after unrolling
                if (PAT == 2 && rf < 4 + cf) continue;
                dmma(acc[cf][rf][0], acc[cf][rf][1], a[cf], bfr[rf]);
            }
    }
}

// Factor the jb x jb diagonal block at A (lower) in place, write inv(L) (NB x NB,
// ld NB, zero upper, identity padding beyond jb) to inv.
//
// The 128x128 block lives in REGISTERS as DMMA accumulator fragments (64 doubles per
// thread, same 8-warp 64x32 layout as the GEMM).  Right-looking with an 8-wide inner
// block: the owning warps drop the current 128x8 column block into shared memory,
// every warp re-derives the 8x8 Cholesky factor with shuffles (no CTA barrier for it),
// 128 row-threads do the 8-step substitution, then all warps apply the rank-8 update to
// their fragments with two DMMA k-steps per tile.  The inverse is then built in shared
// memory by recursive doubling (X21 = -X22 (L21 X11)) with DMMA tile products.
__global__ void __launch_bounds__(256, 1)
potf2_inv_kernel(double *A, long long lda, int jb, double *inv, double *invT, int *info, int joff,
                 long long sA, long long sInv, const double *Tprev, long long ldt,
                 const double *invprev, unsigned long long *trace) {
    extern __shared__ __align__(16) double sm[];
    if (trace && threadIdx.x == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace[0] = t;
    }
#define CVXB_STAMP(SLOT)                                                                    \
    if (trace && threadIdx.x == 0) {                                                        \
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));       \
        trace[8 * 2048 + (SLOT)] = t_;                                                      \
    }
    // Programmatic dependent launch (CVXB_CHOL_PDL): the next diagonal-block kernel of the chain may be scheduled while
    // this one runs; it blocks here until its predecessor has completed and flushed.  No-ops for ordinary launches.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    A += (long long)blockIdx.x * sA; inv += (long long)blockIdx.x * sInv;
    invT += (long long)blockIdx.x * sInv; info += blockIdx.x;
    double *M = sm;                    // NB x LDM, column-major
    double *P = M + NB * LDM;          // NB x SP, row-major panel
    double *Dw = P + NB * SP;          // 4 private copies of the 8x8 factor (+ reciprocal diagonal)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Warp -> 64x32 fragment block.  Only the lower triangle is live, so the blocks carry very
    // different amounts of DMMA work; warps w and w+4 share an SM sub-partition, so pair a heavy
    // block with a light one: SMSP0 (1,0)+(0,2), SMSP1 (1,1)+(0,3), SMSP2 (1,2)+(0,1), SMSP3 (0,0)+(1,3)
    const int role = (0x72640531 >> (warp * 4)) & 7;     // role = wr + 2*wc, one nibble per warp
    const int wr = role & 1, wc = role >> 1;
    const int g4 = lane >> 2, t4 = lane & 3;

    double acc[4][8][2];
    // ---- optional prologue (look-ahead of the blocked factorisation): this diagonal block still
    // misses the previous step's update.  Tprev = A(j, j-1) before its TRSM, invprev = inv(L_{j-1,j-1}):
    //   Lt = Tprev * invprev'   (private copy of L(j, j-1)),   C -= Lt Lt'
    // so the chain of diagonal factorisations never waits for the full-width TRSM / update.
    if (Tprev != nullptr) {
#pragma unroll
        for (int cf = 0; cf < 4; ++cf)
#pragma unroll
            for (int rf = 0; rf < 8; ++rf) acc[cf][rf][0] = acc[cf][rf][1] = 0.0;
        const int kkmax = (wc + 1) * 8;                 // invprev[c, k] = 0 for k > c
        // Tprev (jb x 128) goes to shared memory M in one burst of async copies: read straight from global memory in
        // the k loop, the loop was a chain of L2 round trips (6 us for 0.5 MFLOP per warp)
        {
            const bool v16 = ((ldt & 1) == 0) && ((reinterpret_cast<uintptr_t>(Tprev) & 15) == 0);
            if (v16) {
                for (int q = tid; q < NB * NB / 2; q += 256) {
                    const int c = q >> 6, r = (q & 63) * 2;
                    const int nb = (r + 1 < jb) ? 16 : (r < jb ? 8 : 0);
                    cp_async16(M + r + c * LDM, nb ? Tprev + r + (long long)c * ldt : Tprev, nb);
                }
            } else {
                for (int q = tid; q < NB * NB; q += 256) {
                    const int c = q >> 7, r = q & 127;
                    cp_async8(M + r + c * LDM, (r < jb) ? Tprev + r + (long long)c * ldt : Tprev, (r < jb) ? 8 : 0);
                }
            }
            cp_async_commit();
        }
        const double *ip = invprev + (wc * 32 + g4) + t4 * NB;
        // first fragments of inv(L_prev) while the copies are in flight
        double a0[4];
#pragma unroll
        for (int cf = 0; cf < 4; ++cf) a0[cf] = ip[cf * 8];
        cp_async_wait<0>();
        __syncthreads();
        const double *tq = M + (wr * 64 + g4) + t4 * LDM;
#pragma unroll 2
        for (int kk = 0; kk < kkmax; ++kk) {
            double a[4], bfr[8];
#pragma unroll
            for (int cf = 0; cf < 4; ++cf) a[cf] = a0[cf];
            if (kk + 1 < kkmax) {
#pragma unroll
                for (int cf = 0; cf < 4; ++cf) a0[cf] = ip[cf * 8 + (kk + 1) * 4 * NB];
            }
#pragma unroll
            for (int rf = 0; rf < 8; ++rf) bfr[rf] = tq[rf * 8 + kk * 4 * LDM];
#pragma unroll
            for (int cf = 0; cf < 4; ++cf)
#pragma unroll
                for (int rf = 0; rf < 8; ++rf) dmma(acc[cf][rf][0], acc[cf][rf][1], a[cf], bfr[rf]);
        }
        __syncthreads();                                // every warp is done reading Tprev from M
#pragma unroll
        for (int cf = 0; cf < 4; ++cf)
#pragma unroll
            for (int rf = 0; rf < 8; ++rf)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    M[(wr * 64 + rf * 8 + t4 * 2 + e) + (wc * 32 + cf * 8 + g4) * LDM] = acc[cf][rf][e];
    }
    CVXB_STAMP(0)
    // the block itself: straight into the fragment registers (identity padding beyond jb)
#pragma unroll
    for (int cf = 0; cf < 4; ++cf)
#pragma unroll
        for (int rf = 0; rf < 8; ++rf)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = wr * 64 + rf * 8 + t4 * 2 + e, c = wc * 32 + cf * 8 + g4;
                double v = (r == c) ? 1.0 : 0.0;
                if (r < jb && c < jb && r >= c) v = A[r + (long long)c * lda];
                acc[cf][rf][e] = v;
            }
    __syncthreads();
    CVXB_STAMP(4)
    if (Tprev != nullptr) {
        // which 8x8 tiles of this warp's 64x32 block touch the lower triangle depends only on
        // d = wc - 2*wr: all (d < 0), rf >= cf (d == 0), rf >= 4 + cf (d == 1), none (d > 1).
        // Static patterns keep the DMMA stream free of per-tile branches.
        const int d = wc - 2 * wr;
        const double *Ma = M + (wc * 32 + g4) + t4 * LDM;
        const double *Mb = M + (wr * 64 + g4) + t4 * LDM;
        if (d < 0) sym_update<0>(acc, Ma, Mb);
        else if (d == 0) sym_update<1>(acc, Ma, Mb);
        else if (d == 1) sym_update<2>(acc, Ma, Mb);
        __syncthreads();
    }
    CVXB_STAMP(1)
#pragma unroll 1
    for (int t = 0; t < NB / PB; ++t) {
      {
        const int c0 = t * PB;
        // 1. owners of column block t publish it: P[r][k] = C[r, c0 + k]
        //    (switch keeps the fragment index static without unrolling the whole step 4x)
        if (wc == (t >> 2)) {
#define CVXB_PUBLISH(CF)                                                          \
    _Pragma("unroll") for (int rf = 0; rf < 8; ++rf)                              \
        _Pragma("unroll") for (int e = 0; e < 2; ++e)                             \
            P[(wr * 64 + rf * 8 + t4 * 2 + e) * SP + g4] = acc[CF][rf][e];
            switch (t & 3) {
                case 0: CVXB_PUBLISH(0) break;
                case 1: CVXB_PUBLISH(1) break;
                case 2: CVXB_PUBLISH(2) break;
                default: CVXB_PUBLISH(3) break;
            }
#undef CVXB_PUBLISH
        }
        __syncthreads();
        // 2. 8x8 Cholesky of the diagonal block by a single lane; warps 0-3 pick the factor up from
        //    shared memory behind a 128-thread named barrier, warps 4-7 skip to the CTA barrier.
        if (warp == 0 && lane == 0) {
            // one lane, 36 registers, no shuffles: the critical path per pivot is one rsqrt, one
            // multiply and one FMA (the 32-lane shuffle version spent ~200 cycles per pivot)
            double a[PB][PB];
#pragma unroll
            for (int i = 0; i < PB; ++i)
#pragma unroll
                for (int k = 0; k <= i; ++k) a[i][k] = P[(c0 + i) * SP + k];
            bool bad = false;
#pragma unroll
            for (int j = 0; j < PB; ++j) {
                const double piv = a[j][j];
                if (!(piv > 0.0) && !bad) {
                    atomicCAS(info, 0, joff + c0 + j + 1);
                    bad = true;
                }
                const double rs = rsqrt(piv);
                a[j][j] = piv * rs;
                Dw[64 + j] = rs;
#pragma unroll
                for (int i = j + 1; i < PB; ++i) a[i][j] *= rs;
#pragma unroll
                for (int k = j + 1; k < PB; ++k)
#pragma unroll
                    for (int i = k; i < PB; ++i) a[i][k] -= a[i][j] * a[k][j];
            }
#pragma unroll
            for (int i = 0; i < PB; ++i)
#pragma unroll
                for (int k = 0; k < PB; ++k) Dw[i * 8 + k] = (k <= i) ? a[i][k] : 0.0;
        }
        if (warp < 4) asm volatile("bar.sync 1, 128;" ::: "memory");
        // 3. substitution on the rows below (one row per thread), zero rows above
        if (warp < 4) {
            const double *D = Dw;
            const int r = warp * 32 + lane;
            double x[PB];
            if (r >= c0 + PB) {
#pragma unroll
                for (int j = 0; j < PB; ++j) x[j] = P[r * SP + j];
#pragma unroll
                for (int j = 0; j < PB; ++j) {
                    double v = x[j];
#pragma unroll
                    for (int k = 0; k < j; ++k) v -= x[k] * D[j * 8 + k];
                    x[j] = v * D[64 + j];
                }
            } else if (r >= c0) {
#pragma unroll
                for (int j = 0; j < PB; ++j) x[j] = (j <= r - c0) ? D[(r - c0) * 8 + j] : 0.0;
            } else {
#pragma unroll
                for (int j = 0; j < PB; ++j) x[j] = 0.0;
            }
#pragma unroll
            for (int j = 0; j < PB; ++j) P[r * SP + j] = x[j];
        }
        __syncthreads();
        // 4. rank-8 update of the fragments right/below the block: C -= P P'.  Skipping is
        //    warp-uniform and coarse (whole warp / whole column fragment): per-tile tests put a
        //    branch in front of every DMMA and cost more than the few dead tiles they save.
        if (wr * 8 + 7 > t && wc * 4 + 3 > t) {
            double a[4][2], bfr[8][2];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                for (int cf = 0; cf < 4; ++cf) a[cf][kk] = -P[(wc * 32 + cf * 8 + g4) * SP + kk * 4 + t4];
#pragma unroll
                for (int rf = 0; rf < 8; ++rf) bfr[rf][kk] = P[(wr * 64 + rf * 8 + g4) * SP + kk * 4 + t4];
            }
#pragma unroll
            for (int cf = 0; cf < 4; ++cf) {
                if (wc * 4 + cf <= t) continue;            // column block already final
#pragma unroll
                for (int rf = 0; rf < 8; ++rf) {
                    dmma(acc[cf][rf][0], acc[cf][rf][1], a[cf][0], bfr[rf][0]);
                    dmma(acc[cf][rf][0], acc[cf][rf][1], a[cf][1], bfr[rf][1]);
                }
            }
        }
        // 5. the finished column block of L goes to shared M and to global memory
        for (int e = tid; e < NB * PB; e += 256) {
            const int r = e & (NB - 1), j = e >> 7;
            const double v = P[r * SP + j];
            M[r + (c0 + j) * LDM] = v;
            if (r >= c0 + j && r < jb && c0 + j < jb) A[r + (long long)(c0 + j) * lda] = v;
        }
        __syncthreads();
      }
    }

    CVXB_STAMP(2)
    // ---- inverse of L in shared memory ----
    // base: the 16 8x8 diagonal blocks, one per half-warp; lane (lane&15) < 8 owns column j
    {
        const int blk = warp * 2 + (lane >> 4);
        const int j = lane & 15;
        const int o = blk * PB;
        double x[PB];
        if (j < PB) {
#pragma unroll
            for (int i = 0; i < PB; ++i) {
                double v = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k)
                    if (k >= j) v -= M[(o + i) + (o + k) * LDM] * x[k];
                x[i] = (i >= j) ? v / M[(o + i) + (o + i) * LDM] : 0.0;
            }
        }
        __syncwarp();
        if (j < PB) {
#pragma unroll
            for (int i = 0; i < PB; ++i) M[(o + i) + (o + j) * LDM] = x[i];
        }
    }
    __syncthreads();
    for (int s = PB; s < NB; s <<= 1) {
        const int sb = s / PB;                   // 8-blocks per side
        const int tp = sb * sb;                  // tiles per pair
        const int ntiles = (NB / (2 * s)) * tp;
        // phase 0: T = L21 * X11 -> stored in the (unused) upper block of the pair
        // phase 1: X21 = - X22 * T
        // each warp walks its tiles four at a time so four DMMA chains are in flight
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            for (int base = warp * 4; base < ntiles; base += 32) {
                double c[4][2];
                int oo[4], ti[4], tj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = min(base + u, ntiles - 1);
                    const int pr = idx / tp, loc = idx - pr * tp;
                    ti[u] = loc % sb; tj[u] = loc / sb; oo[u] = pr * 2 * s;
                    c[u][0] = c[u][1] = 0.0;
                }
                for (int kb = 0; kb < sb; ++kb) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        // X11 / X22 are lower triangular: skip the zero k blocks
                        const bool live = phase == 0 ? (kb >= tj[u]) : (kb <= ti[u]);
                        if (!live) continue;
                        const int o = oo[u];
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const int k = kb * 8 + kk * 4 + t4;
                            double af, bf;
                            if (phase == 0) {
                                af = M[(o + s + ti[u] * 8 + g4) + (o + k) * LDM];
                                bf = M[(o + k) + (o + tj[u] * 8 + g4) * LDM];
                            } else {
                                af = M[(o + s + ti[u] * 8 + g4) + (o + s + k) * LDM];
                                bf = M[(o + k) + (o + s + tj[u] * 8 + g4) * LDM];
                            }
                            dmma(c[u][0], c[u][1], af, bf);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (base + u >= ntiles) continue;
                    const int o = oo[u];
                    if (phase == 0) {
                        M[(o + ti[u] * 8 + g4) + (o + s + tj[u] * 8 + t4 * 2) * LDM] = c[u][0];
                        M[(o + ti[u] * 8 + g4) + (o + s + tj[u] * 8 + t4 * 2 + 1) * LDM] = c[u][1];
                    } else {
                        M[(o + s + ti[u] * 8 + g4) + (o + tj[u] * 8 + t4 * 2) * LDM] = -c[u][0];
                        M[(o + s + ti[u] * 8 + g4) + (o + tj[u] * 8 + t4 * 2 + 1) * LDM] = -c[u][1];
                    }
                }
            }
            __syncthreads();
        }
    }
    CVXB_STAMP(3)
    for (int e = tid; e < NB * NB; e += 256) {
        int i = e & (NB - 1), k = e >> 7;
        inv[e] = (i >= k) ? M[i + k * LDM] : 0.0;
        invT[e] = (k >= i) ? M[k + i * LDM] : 0.0;     // invT[i + k*NB] = inv[k + i*NB]
    }
    if (trace && threadIdx.x == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace[1] = t;
    }
}

// ---- trailing update of the blocked factorisation:  A22(lower) -= L21 L21'  for one 128-wide panel ----
// The generic DMMA GEMM runs this K = 128 update at 21-25 TF/s (57-67 % of the DMMA peak, profiles/r02v): every 128x64
// tile is its own CTA with a pipeline prologue and an epilogue for only eight k steps.  Here a CTA (one per SM at a
// time) walks a few consecutive 128x128 tiles strip by strip (strip = block column cb of the trailing matrix, rb >= cb):
//   * the strip's column operand L21[cb rows, 0:128] (128 KB) stays in shared memory for all tiles of the strip;
//   * the row operand streams through two 32-column buffers, the copy of the next chunk -- of the next TILE after
//     the last chunk -- in flight under the DMMAs of the current one, so there is no per-tile prologue;
//   * the tile of A22 goes straight into the accumulator fragments (acc = C, a-fragments negated: acc = C - L L')
//     and is stored from them: no shared-memory staging, no epilogue barrier.
// STATUS: validated correct (tests/test_kkt_gpu.py, tests/test_fullsize_gpu.py with CVXB_CHOL_TU=1) but SLOWER than the
// generic kernel on B200 (17.6 vs 24 TF/s at the first step of n=8192: one 8-warp CTA per SM without register
// double-buffering of the fragments and with the C tile's load/store latency exposed at every tile boundary does not
// keep the DMMA pipe as busy as two 4-warp CTAs of the generic kernel do).  Opt-in: CVXB_CHOL_TU=1.
// W = L21 (m x 128, column-major, ld ldw); tiles of block column 0 are left to the caller (the panel stream updates
// the next block column itself).  Requires even ldc / 16-byte aligned C (the caller falls back to dmma_gemm otherwise).
constexpr int TU_LD = NB + 4;                   // 132: row stride of the [k][idx] operand buffers (conflict-free LDS.64)
constexpr int TU_CH = 32;                       // k columns per streamed chunk
constexpr int TU_NCH = NB / TU_CH;
constexpr int TU_SMEM = (NB * TU_LD + 2 * TU_CH * TU_LD) * 8;      // 202752 B

__global__ void __launch_bounds__(256, 1)
chol_trailing_kernel(int m, const double *__restrict__ W, long long ldw, double *C, long long ldc, int nbk, int cb0,
                     long long T, unsigned long long *trace) {
    extern __shared__ __align__(16) double sm[];
    double *Bs = sm;                            // [128 k][TU_LD]  rows of block column cb
    double *As = sm + NB * TU_LD;               // 2 x [32 k][TU_LD]  rows of block rb
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp & 1, wc = warp >> 1;
    const int g4 = lane >> 2, t4 = lane & 3;
    if (trace && tid == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicCAS(trace, 0ULL, t);
    }
    const long long t0 = T * blockIdx.x / gridDim.x, t1 = T * (blockIdx.x + 1) / gridDim.x;
    if (t0 >= t1) return;
    // tile t -> (cb, rb): strips cb = cb0 .. nbk-1, strip cb holds rb = cb .. nbk-1
    int cb = cb0, rb;
    {
        long long t = t0;
        while (t >= nbk - cb) { t -= nbk - cb; ++cb; }
        rb = cb + (int)t;
    }
    auto issue_rows = [&](double *dst, int blk, int k0, int nk) {       // dst[kk][r] = W[blk*128 + r, k0 + kk]
        for (int q = tid; q < nk * (NB / 2); q += 256) {
            const int kk = q >> 6, rr = (q & 63) * 2;
            const long long row = (long long)blk * NB + rr;
            const int bytes = (row + 1 < m) ? 16 : (row < m ? 8 : 0);
            cp_async16(dst + kk * TU_LD + rr, bytes ? W + row + (long long)(k0 + kk) * ldw : W, bytes);
        }
    };
    int cur_cb = -1, buf = 0;
    issue_rows(As, rb, 0, TU_CH);
    cp_async_commit();
    for (long long t = t0; t < t1; ++t) {
        if (cb != cur_cb) {                     // new strip: every warp is past the previous tile's last barrier
            issue_rows(Bs, cb, 0, NB);
            cp_async_commit();
            cur_cb = cb;
        }
        // ---- the tile of A22 -> accumulators ----
        const long long r0 = (long long)rb * NB, c0 = (long long)cb * NB;
        const bool interior = (rb > cb) && (r0 + NB <= m);
        double acc[4][8][2];
        double *Ct = C + r0 + c0 * ldc;
#pragma unroll
        for (int cf = 0; cf < 4; ++cf) {
            const int cl = wc * 32 + cf * 8 + g4;
#pragma unroll
            for (int rf = 0; rf < 8; ++rf) {
                const int rl = wr * 64 + rf * 8 + t4 * 2;
                if (interior) {
                    const double2 v = *reinterpret_cast<const double2 *>(Ct + rl + cl * ldc);
                    acc[cf][rf][0] = v.x; acc[cf][rf][1] = v.y;
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const long long r = r0 + rl + e, c = c0 + cl;
                        acc[cf][rf][e] = (r < m && c < m && r >= c) ? Ct[rl + e + cl * ldc] : 0.0;
                    }
                }
            }
        }
        // next tile (for the cross-tile prefetch)
        int ncb = cb, nrb = rb + 1;
        if (nrb >= nbk) { ++ncb; nrb = ncb; }
        const bool has_next = (t + 1 < t1);
#pragma unroll 1
        for (int ch = 0; ch < TU_NCH; ++ch) {
            if (ch + 1 < TU_NCH) issue_rows(As + (buf ^ 1) * TU_CH * TU_LD, rb, (ch + 1) * TU_CH, TU_CH);
            else if (has_next) issue_rows(As + (buf ^ 1) * TU_CH * TU_LD, nrb, 0, TU_CH);
            cp_async_commit();
            cp_async_wait<1>();                 // everything but the chunk just issued has landed (incl. a new Bs)
            __syncthreads();
            const double *Ab = As + buf * TU_CH * TU_LD + (wr * 64 + g4) + t4 * TU_LD;
            const double *Bb = Bs + (ch * TU_CH) * TU_LD + (wc * 32 + g4) + t4 * TU_LD;
#pragma unroll
            for (int kk = 0; kk < TU_CH / 4; ++kk) {
                double a[4], bf[8];
#pragma unroll
                for (int cf = 0; cf < 4; ++cf) a[cf] = -Bb[cf * 8 + kk * 4 * TU_LD];
#pragma unroll
                for (int rf = 0; rf < 8; ++rf) bf[rf] = Ab[rf * 8 + kk * 4 * TU_LD];
#pragma unroll
                for (int cf = 0; cf < 4; ++cf)
#pragma unroll
                    for (int rf = 0; rf < 8; ++rf) dmma(acc[cf][rf][0], acc[cf][rf][1], a[cf], bf[rf]);
            }
            __syncthreads();                    // the buffer may be refilled by the next issue
            buf ^= 1;
        }
        // ---- accumulators -> A22 ----
#pragma unroll
        for (int cf = 0; cf < 4; ++cf) {
            const int cl = wc * 32 + cf * 8 + g4;
#pragma unroll
            for (int rf = 0; rf < 8; ++rf) {
                const int rl = wr * 64 + rf * 8 + t4 * 2;
                if (interior) {
                    *reinterpret_cast<double2 *>(Ct + rl + cl * ldc) = make_double2(acc[cf][rf][0], acc[cf][rf][1]);
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const long long r = r0 + rl + e, c = c0 + cl;
                        if (r < m && c < m && r >= c) Ct[rl + e + cl * ldc] = acc[cf][rf][e];
                    }
                }
            }
        }
        cb = ncb; rb = nrb;
    }
    cp_async_wait<0>();
    if (trace && tid == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(trace + 1, t);
    }
}

// Second form of the strip kernel (CVXB_CHOL_TU=2; validated, also slower: potrf 11.8 ms, profiles/r02x): the generic GEMM's geometry -- 128x64 tiles, four warps, two CTAs per
// SM so that one CTA's tile boundary (store C, load the next C into the accumulators) runs under the other's DMMAs --
// but a CTA walks several consecutive tiles of a 64-column strip and its 3-stage operand ring never drains: the chunks
// of the next tile follow the last chunk of the current one.
constexpr int T2_BC = 64, T2_LDA = NB + 4, T2_LDB = T2_BC + 4, T2_CH = 16, T2_ST = 3;
constexpr int T2_STAGE = T2_CH * (T2_LDA + T2_LDB);            // doubles per stage: A chunk then B chunk
constexpr int T2_SMEM = T2_ST * T2_STAGE * 8;                  // 76800 B
constexpr int T2_NCH = NB / T2_CH;                             // 8 chunks per tile

__global__ void __launch_bounds__(128, 2)
chol_trailing2_kernel(int m, const double *__restrict__ W, long long ldw, double *C, long long ldc, int nbk, long long T,
                      unsigned long long *trace) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp & 1, wc = warp >> 1;
    const int g4 = lane >> 2, t4 = lane & 3;
    if (trace && tid == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicCAS(trace, 0ULL, t);
    }
    const long long t0 = T * blockIdx.x / gridDim.x, t1 = T * (blockIdx.x + 1) / gridDim.x;
    if (t0 >= t1) return;
    // tile t -> (cs, rb): 64-column strips cs = 2 .. 2*nbk-1 (the first 128 columns belong to the panel stream),
    // strip cs holds the 128-row blocks rb = cs/2 .. nbk-1
    int cs = 2, rb;
    {
        long long t = t0;
        while (t >= nbk - (cs >> 1)) { t -= nbk - (cs >> 1); ++cs; }
        rb = (cs >> 1) + (int)t;
    }
    auto issue = [&](int stage, int rblk, int cstrip, int k0) {
        double *As = sm + stage * T2_STAGE, *Bs = As + T2_CH * T2_LDA;
        for (int q = tid; q < T2_CH * (NB / 2); q += 128) {           // A: rows of block rblk
            const int kk = q >> 6, rr = (q & 63) * 2;
            const long long row = (long long)rblk * NB + rr;
            const int bytes = (row + 1 < m) ? 16 : (row < m ? 8 : 0);
            cp_async16(As + kk * T2_LDA + rr, bytes ? W + row + (long long)(k0 + kk) * ldw : W, bytes);
        }
        for (int q = tid; q < T2_CH * (T2_BC / 2); q += 128) {        // B: rows of strip cstrip
            const int kk = q >> 5, rr = (q & 31) * 2;
            const long long row = (long long)cstrip * T2_BC + rr;
            const int bytes = (row + 1 < m) ? 16 : (row < m ? 8 : 0);
            cp_async16(Bs + kk * T2_LDB + rr, bytes ? W + row + (long long)(k0 + kk) * ldw : W, bytes);
        }
    };
    // the chunk stream of this CTA: chunk q of tile i; `pf_*` walks two chunks ahead of the compute position
    int pf_cs = cs, pf_rb = rb, pf_ch = 0;
    long long pf_t = t0;
    int pf_stage = 0;
    auto prefetch = [&]() {
        if (pf_t < t1) {
            issue(pf_stage, pf_rb, pf_cs, pf_ch * T2_CH);
            if (++pf_ch == T2_NCH) {
                pf_ch = 0; ++pf_t;
                if (++pf_rb >= nbk) { ++pf_cs; pf_rb = pf_cs >> 1; }
            }
        }
        cp_async_commit();
        if (++pf_stage == T2_ST) pf_stage = 0;
    };
    prefetch();
    prefetch();
    int stage = 0;
    for (long long t = t0; t < t1; ++t) {
        const long long r0 = (long long)rb * NB, c0 = (long long)cs * T2_BC;
        const bool interior = (r0 >= c0 + T2_BC) && (r0 + NB <= m);
        double acc[4][8][2];
        double *Ct = C + r0 + c0 * ldc;
#pragma unroll
        for (int cf = 0; cf < 4; ++cf) {
            const int cl = wc * 32 + cf * 8 + g4;
#pragma unroll
            for (int rf = 0; rf < 8; ++rf) {
                const int rl = wr * 64 + rf * 8 + t4 * 2;
                if (interior) {
                    const double2 v = *reinterpret_cast<const double2 *>(Ct + rl + cl * ldc);
                    acc[cf][rf][0] = v.x; acc[cf][rf][1] = v.y;
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const long long r = r0 + rl + e, c = c0 + cl;
                        acc[cf][rf][e] = (r < m && c < m && r >= c) ? Ct[rl + e + cl * ldc] : 0.0;
                    }
                }
            }
        }
#pragma unroll 1
        for (int ch = 0; ch < T2_NCH; ++ch) {
            cp_async_wait<T2_ST - 2>();         // this chunk has landed (one newer group may be in flight)
            __syncthreads();                    // ... for every thread; the stage read last iteration is free
            prefetch();                         // two chunks ahead, into the stage freed by the barrier above
            const double *As = sm + stage * T2_STAGE, *Bs = As + T2_CH * T2_LDA;
            const double *Ab = As + (wr * 64 + g4) + t4 * T2_LDA;
            const double *Bb = Bs + (wc * 32 + g4) + t4 * T2_LDB;
#pragma unroll
            for (int kk = 0; kk < T2_CH / 4; ++kk) {
                double a[4], bf[8];
#pragma unroll
                for (int cf = 0; cf < 4; ++cf) a[cf] = -Bb[cf * 8 + kk * 4 * T2_LDB];
#pragma unroll
                for (int rf = 0; rf < 8; ++rf) bf[rf] = Ab[rf * 8 + kk * 4 * T2_LDA];
#pragma unroll
                for (int cf = 0; cf < 4; ++cf)
#pragma unroll
                    for (int rf = 0; rf < 8; ++rf) dmma(acc[cf][rf][0], acc[cf][rf][1], a[cf], bf[rf]);
            }
            if (++stage == T2_ST) stage = 0;
        }
#pragma unroll
        for (int cf = 0; cf < 4; ++cf) {
            const int cl = wc * 32 + cf * 8 + g4;
#pragma unroll
            for (int rf = 0; rf < 8; ++rf) {
                const int rl = wr * 64 + rf * 8 + t4 * 2;
                if (interior) {
                    *reinterpret_cast<double2 *>(Ct + rl + cl * ldc) = make_double2(acc[cf][rf][0], acc[cf][rf][1]);
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const long long r = r0 + rl + e, c = c0 + cl;
                        if (r < m && c < m && r >= c) Ct[rl + e + cl * ldc] = acc[cf][rf][e];
                    }
                }
            }
        }
        if (++rb >= nbk) { ++cs; rb = cs >> 1; }
    }
    cp_async_wait<0>();
    if (trace && tid == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(trace + 1, t);
    }
}

// ---- blocked triangular solve with one right-hand side ------------------------
// forward:  L x = b ;  backward: L' x = b.   In place on b.  One CTA per 128-row block.
// flags[i] == epoch  <=>  x_i is final in b.
//
// Forward-progress assumption: CTA `bi` spins on the flags of the blocks it depends on, which are produced by
// CTAs with a SMALLER blockIdx.x (forward; the backward kernel reverses the block order so the same holds).  The
// hardware dispatches the CTAs of a grid in ascending linear block index, so a spinning CTA's producers are always
// resident or already finished; with at most ceil(n/128) CTAs per problem the whole grid is usually co-resident.
// CTA `bi` streams its block row (forward) / block column (backward) of L through a
// cp.async ring of 128x32 chunks that runs ahead of the dependency chain (L is static,
// only x arrives late), keeps inv(L_ii) in registers, and waits on the flag of block j
// only when it reaches that block's columns.  The critical path per block is then
// flag -> 128x128 smem GEMV -> 128x128 register GEMV -> flag.
constexpr int TR_CH = 32;                       // columns per ring chunk
constexpr int TR_R = 5;                         // ring depth
constexpr int TRSV_SMEM = (TR_R * NB * TR_CH + 4 * NB + 32 * 129) * 8;

template <bool TRANS, bool VEC>
__global__ void __launch_bounds__(256, 1)
trsv_kernel(int n, const double *__restrict__ L, long long ldl, const double *__restrict__ inv,
            const double *__restrict__ invT, double *b, int *flags, int epoch, long long sL,
            long long sInv, long long sb) {
    extern __shared__ __align__(16) double sm[];
    {   // batched: blockIdx.y selects the problem (blocks of one problem keep ascending order)
        const long long pb = blockIdx.y;
        L += pb * sL; inv += pb * sInv; invT += pb * sInv; b += pb * sb;
        flags += pb * ((n + NB - 1) / NB);
    }
    double *ring = sm;                          // TR_R x (32 cols x 128 rows), column-contiguous
    double *xs = ring + TR_R * NB * TR_CH;      // 128
    double *part = xs + NB;                     // 2 x 128
    double *tvec = part + 2 * NB;               // 128
    double *part2 = tvec + NB;                  // 32 x 129: per-lane column partials (backward)
    const int nblk = (n + NB - 1) / NB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bi = TRANS ? (nblk - 1 - (int)blockIdx.x) : (int)blockIdx.x;
    const int i0 = bi * NB;
    const int ni = min(NB, n - i0);
    const int r = tid & (NB - 1), half = tid >> 7;

    // inverse of the diagonal block -> registers (row r of inv, or row r of inv' for TRANS)
    double ireg[64];
    {
        const double *ip = (TRANS ? invT : inv) + (long long)bi * NB * NB + r + (long long)(half * 64) * NB;
#pragma unroll
        for (int k = 0; k < 64; ++k) ireg[k] = ip[k * NB];
    }

    // this block's right-hand side is input data nobody else writes: fetch it before the chain
    const double bown = (tid < ni) ? b[i0 + tid] : 0.0;
    const int nb_dep = TRANS ? (nblk - 1 - bi) : bi;         // blocks this CTA depends on
    const int nchunks = nb_dep * (NB / TR_CH);
    // chunk q -> (block j, column chunk cc)
    auto issue = [&](int q) {
        const int jj = q / (NB / TR_CH), cc = q % (NB / TR_CH);
        const int j = TRANS ? (nblk - 1 - jj) : jj;
        double *dst = ring + (q % TR_R) * (NB * TR_CH);
        const double *src;
        int nrows, ncols;
        if (!TRANS) { src = L + i0 + (long long)(j * NB + cc * TR_CH) * ldl; nrows = ni; ncols = TR_CH; }
        else { src = L + (long long)j * NB + (long long)(i0 + cc * TR_CH) * ldl; nrows = min(NB, n - j * NB); ncols = min(TR_CH, ni - cc * TR_CH); }
        if (VEC) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int e = tid + it * 256;          // 2048 16-byte pieces
                const int c = e >> 6, rr = (e & 63) * 2;
                const int rem = nrows - rr;
                const int bytes = (c < ncols) ? (rem >= 2 ? 16 : (rem == 1 ? 8 : 0)) : 0;
                cp_async16(dst + c * NB + rr, bytes ? (src + rr + (long long)c * ldl) : L, bytes);
            }
        } else {
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int e = tid + it * 256;
                const int c = e >> 7, rr = e & 127;
                const int bytes = (c < ncols && rr < nrows) ? 8 : 0;
                cp_async8(dst + c * NB + rr, bytes ? (src + rr + (long long)c * ldl) : L, bytes);
            }
        }
    };
#pragma unroll
    for (int q = 0; q < TR_R - 1; ++q) {
        if (q < nchunks) issue(q);
        cp_async_commit();
    }

    double acc = 0.0;           // forward: partial of row r over this thread's columns
    double accc[16];            // backward: per-lane partials of the warp's 16 columns
#pragma unroll
    for (int q = 0; q < 16; ++q) accc[q] = 0.0;

    for (int jj = 0; jj < nb_dep; ++jj) {
        const int j = TRANS ? (nblk - 1 - jj) : jj;
        if (tid == 0) while (ld_acquire(flags + j) != epoch) { }
        __syncthreads();
        if (tid < NB) xs[tid] = (j * NB + tid < n) ? __ldcg(b + j * NB + tid) : 0.0;
#pragma unroll
        for (int cc = 0; cc < NB / TR_CH; ++cc) {
            const int q = jj * (NB / TR_CH) + cc;
            cp_async_wait<TR_R - 2>();
            __syncthreads();
            {
                const int nq = q + TR_R - 1;
                if (nq < nchunks) issue(nq);
                cp_async_commit();
            }
            const double *ch = ring + (q % TR_R) * (NB * TR_CH);
            if (!TRANS) {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    acc += ch[(half * 16 + c) * NB + r] * xs[cc * TR_CH + half * 16 + c];
            } else {
                // warp w owns chunk columns w*4 .. w*4+3; lanes stride the 128 rows
#pragma unroll
                for (int qc = 0; qc < 4; ++qc) {
                    const double *col = ch + (warp * 4 + qc) * NB;
                    double a = 0.0;
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) a += col[lane + rr * 32] * xs[lane + rr * 32];
                    accc[cc * 4 + qc] += a;
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (!TRANS) {
        part[half * NB + r] = acc;
        __syncthreads();
        if (tid < NB) tvec[tid] = (tid < ni) ? (bown - part[tid] - part[NB + tid]) : 0.0;
        __syncthreads();
    } else {
        // cross-lane reduction of the 16 column partials through shared memory (16 warp-shuffle
        // trees cost ~1.6 us on the critical path of every block step)
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int c = (q >> 2) * TR_CH + warp * 4 + (q & 3);
            part2[lane * 129 + c] = accc[q];
        }
        __syncthreads();
        if (tid < NB) {
            double sacc = 0.0;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) sacc += part2[l * 129 + tid];
            tvec[tid] = (tid < ni) ? (bown - sacc) : 0.0;
        }
        __syncthreads();
    }
    // x_i = inv_ii * t  (forward)  /  inv_ii' * t (backward): row r of the (transposed) inverse
    double a2 = 0.0;
#pragma unroll
    for (int k = 0; k < 64; ++k) a2 += ireg[k] * tvec[half * 64 + k];
    part[half * NB + r] = a2;
    __syncthreads();
    if (tid < ni) b[i0 + tid] = part[tid] + part[NB + tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(flags + bi, epoch);
}

std::atomic<int> g_trsv_epoch{0};   // flags written by an earlier launch never equal a later epoch

}  // namespace

int chol_work_create(CholWork &w) {
    int least = 0, greatest = 0;
    CVXB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CVXB_CUDA(cudaStreamCreateWithPriority(&w.panel_stream, cudaStreamNonBlocking, greatest));
    CVXB_CUDA(cudaStreamCreateWithPriority(&w.trsm_stream, cudaStreamNonBlocking, greatest));
    CVXB_CUDA(cudaStreamCreateWithPriority(&w.update_stream, cudaStreamNonBlocking, least));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_end_t, cudaEventDisableTiming));
    w.r_valid[0] = w.r_valid[1] = -1;
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_start, cudaEventDisableTiming));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_panel, cudaEventDisableTiming));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_rest, cudaEventDisableTiming));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_end_p, cudaEventDisableTiming));
    CVXB_CUDA(cudaEventCreateWithFlags(&w.ev_end_u, cudaEventDisableTiming));
    CVXB_CUDA(cudaMalloc(&w.d_info, sizeof(int)));
    CVXB_CUDA(cudaMemset(w.d_info, 0, sizeof(int)));
    w.flags_cap = 4096;
    CVXB_CUDA(cudaMalloc(&w.d_flags, 4096 * sizeof(int)));
    CVXB_CUDA(cudaMemset(w.d_flags, 0, 4096 * sizeof(int)));
    CVXB_CUDA(cudaMalloc(&w.splitk_ws, dmma_gemm_splitk_ws_doubles() * sizeof(double)));
    CVXB_CUDA(cudaFuncSetAttribute(chol_trailing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TU_SMEM));
    CVXB_CUDA(cudaFuncSetAttribute(chol_trailing2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
    CVXB_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   POTF2_SMEM));
    CVXB_CUDA(cudaFuncSetAttribute(trsv_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CVXB_CUDA(cudaFuncSetAttribute(trsv_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CVXB_CUDA(cudaFuncSetAttribute(trsv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CVXB_CUDA(cudaFuncSetAttribute(trsv_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    return 0;
}

void chol_work_destroy(CholWork &w) {
    if (w.panel_stream) cudaStreamDestroy(w.panel_stream);
    if (w.update_stream) cudaStreamDestroy(w.update_stream);
    if (w.trsm_stream) cudaStreamDestroy(w.trsm_stream);
    for (auto &g : w.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    w.graphs.clear();
    if (w.ev_end_t) cudaEventDestroy(w.ev_end_t);
    for (auto *v : {&w.ev_dg, &w.ev_tr, &w.ev_c0, &w.ev_r})
        for (cudaEvent_t e : *v) cudaEventDestroy(e);
    if (w.ev_start) cudaEventDestroy(w.ev_start);
    if (w.ev_panel) cudaEventDestroy(w.ev_panel);
    if (w.ev_rest) cudaEventDestroy(w.ev_rest);
    if (w.ev_end_p) cudaEventDestroy(w.ev_end_p);
    if (w.ev_end_u) cudaEventDestroy(w.ev_end_u);
    if (w.d_info) cudaFree(w.d_info);
    if (w.d_flags) cudaFree(w.d_flags);
    if (w.splitk_ws) cudaFree(w.splitk_ws);
    if (w.panel[0]) cudaFree(w.panel[0]);
    if (w.panel[1]) cudaFree(w.panel[1]);
    w = CholWork();
}

// Look-ahead schedule (three streams, per-step events):
//   D  (diag chain)   Dg(j): potf2_inv of block (j,j); for j>0 its prologue first applies the
//                     step j-1 update to the block from the RAW tile A(j,j-1) and inv(j-1), so the
//                     chain Dg(j-1) -> Dg(j) never waits for a full-width kernel.
//   T  (panel)        Tr(j): Wp = A(j+1:, j) inv(j)'           needs Dg(j), C0(j-1)
//                     C0(j): block column j+1 below the diagonal block -= Wp Wp'   needs R(j-1)
//   U  (bulk)         R(j):  all later block columns (lower tiles) -= Wp Wp'        needs Tr(j)
//   Dg(j) needs C0(j-2) and R(j-2) (they produced the tiles it reads).  L21 is copied from Wp
//   back into A on D after Dg(j+1) has consumed the raw tile.
static int potrf_enqueue(int n, double *A, int lda, double *inv, CholWork &w, cudaStream_t st) {
    if (n <= 0) return 0;
    const int nblk = (n + NB - 1) / NB;
    if (w.panel_rows < n) {
        // captured graphs bake in the panel addresses and their leading dimension: drop them
        for (auto &g : w.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        w.graphs.clear();
        for (int i = 0; i < 2; ++i) {
            if (w.panel[i]) CVXB_CUDA(cudaFree(w.panel[i]));
            w.panel[i] = nullptr;
        }
        // two group buffers, each holding the TRSM results of a PAIR of consecutive panels side by side
        // (rows x 2 NB), so that the bulk trailing update of a pair is one K = 256 product
        const int rows = (n + 1) & ~1;
        for (int i = 0; i < 2; ++i) CVXB_CUDA(cudaMalloc(&w.panel[i], (size_t)rows * 2 * NB * sizeof(double)));
        w.panel_rows = rows;
    }
    while ((int)w.ev_dg.size() < nblk) {
        cudaEvent_t e[4];
        for (int i = 0; i < 4; ++i) CVXB_CUDA(cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming));
        w.ev_dg.push_back(e[0]); w.ev_tr.push_back(e[1]); w.ev_c0.push_back(e[2]); w.ev_r.push_back(e[3]);
    }
    if (getenv("CVXB_TRACE") && !w.trace) {
        CVXB_CUDA(cudaMalloc(&w.trace, 8 * 4096 * sizeof(unsigned long long)));
    }
    if (w.trace) CVXB_CUDA(cudaMemsetAsync(w.trace, 0, 8 * 4096 * sizeof(unsigned long long), st));
    const int ldw = w.panel_rows;
    const int panel_tiles = NB / dmma_gemm_tile_cols();      // c tiles that make up one block column
    cudaStream_t D = w.panel_stream, T = w.trsm_stream, U = w.update_stream;
    CVXB_CUDA(cudaMemsetAsync(w.d_info, 0, sizeof(int), st));
    CVXB_CUDA(cudaEventRecord(w.ev_start, st));
    CVXB_CUDA(cudaStreamWaitEvent(D, w.ev_start, 0));
    CVXB_CUDA(cudaStreamWaitEvent(T, w.ev_start, 0));
    CVXB_CUDA(cudaStreamWaitEvent(U, w.ev_start, 0));
    // Pair aggregation of the bulk update (opt-in, CVXB_CHOL_PAIR=1; measured slower on B200 — potrf 10.99 ms
    // against 10.31 ms at n=8192, 3.13 against 2.92 ms at n=4096, profiles/r02e — because the K = 256 column
    // update of the odd steps lengthens the panel path while the chain, not the bulk rate, sets the pace):
    //   even step 2g  : TRSM -> group buffer columns [0, NB);  C0 (K = NB) completes block column 2g+1;
    //                   D0: diagonal tile (2g+2, 2g+2) -= its panel-2g part (the only tile of block column 2g+2
    //                   the diagonal chain needs before the pair's bulk update exists); NO bulk update.
    //   odd step 2g+1 : TRSM -> columns [NB, 2 NB);  C0 with K = 2 NB (panels 2g and 2g+1) completes block
    //                   column 2g+2 below its diagonal tile;  bulk R with K = 2 NB on block columns >= 2g+3.
    // A K = 128 update keeps the DMMA pipe 63 % busy (prologue + read-modify-write epilogue per 8 k steps,
    // profiles/r01e); K = 256 halves the C traffic and the per-tile overhead of the bulk flops.
    // CVXB_CHOL_PAIR = 0 off (default), 1 every step, k >= 2: the first k (even) steps only — the pair form helps
    // where the bulk update sets the pace (large trailing matrix) and hurts once the diagonal chain does.
    static int pair_mode = -1;
    if (pair_mode < 0) {
        const char *e = getenv("CVXB_CHOL_PAIR");
        pair_mode = e ? atoi(e) : 0;
        if (pair_mode < 0) pair_mode = 0;
    }
    const int pair_limit = pair_mode == 1 ? (1 << 30) : (pair_mode & ~1);
    int last_r = -1;                       // last step that recorded ev_r
    int prev_r = -1;                       // the one before
    for (int jb = 0; jb < nblk; ++jb) {
        const int j = jb * NB;
        const int wj = (n - j < NB) ? (n - j) : NB;
        const int m = n - j - wj;
        double *Ajj = A + j + (long long)j * lda;
        double *invj = inv + (long long)jb * NB * NB;
        double *invTj = inv + (long long)(nblk + jb) * NB * NB;
        const bool pair = jb < pair_limit;
        const bool pair_prev = (jb - 1) < pair_limit;
        const bool odd = pair && (jb & 1);
        // Group buffer of this step.  Its row r is global row R0 + r, R0 = first row below the EVEN panel's
        // diagonal block, for both panels of the pair: the odd panel's TRSM result therefore sits at column
        // offset NB and row offset NB.  `Wo` = buffer row of this step's first trailing row (A22).
        double *Wg = pair ? w.panel[(jb >> 1) & 1] : w.panel[jb & 1];
        double *Wp = Wg + (odd ? (long long)NB * ldw + NB : 0);
        double *Wo = Wg + (odd ? NB : 0);
        // ---- D: diagonal block.  Needs A(jb,jb) updated through panel jb-2 and the raw tile A(jb,jb-1)
        // (block column jb-1 complete through panel jb-2): C0(jb-2) [which follows D0(jb-2) on T] and the
        // last bulk update that touched block column jb.
        static int nowait = -1;      // timing experiment only (WRONG results): chain stream without its cross-stream waits
        if (nowait < 0) { const char *e = getenv("CVXB_CHOL_NOWAIT_EXPERIMENT"); nowait = (e && e[0] == '1') ? 1 : 0; }
        if (jb >= 2 && !nowait) {
            CVXB_CUDA(cudaStreamWaitEvent(D, w.ev_c0[jb - 2], 0));
            // the latest bulk update issued at a step <= jb-2 (a later one belongs to panels the prologue
            // applies itself, waiting for it would serialise the chain behind the bulk work)
            // (pair region, odd step: the pair's own bulk update (step jb-2) does not touch this step's diagonal
            // tile — the odd step's column update C1 did — so the one before it is the one to wait for)
            const int rd = (pair_prev && (jb & 1) && last_r == jb - 2) ? prev_r : ((last_r <= jb - 2) ? last_r : prev_r);
            if (rd >= 0) CVXB_CUDA(cudaStreamWaitEvent(D, w.ev_r[rd], 0));
        }
        const double *Tprev = jb > 0 ? A + j + (long long)(j - NB) * lda : nullptr;
        const double *invprev = jb > 0 ? inv + (long long)(jb - 1) * NB * NB : nullptr;
        static int pdl = -1;
        if (pdl < 0) { const char *e = getenv("CVXB_CHOL_PDL"); pdl = (e && e[0] == '1') ? 1 : 0; }
        if (pdl && jb > 0) {
            // the edge potf2(jb-1) -> potf2(jb) on the chain stream becomes a programmatic dependency: the launch
            // latency (~10 us of the ~107 us per chain step, profiles/r01_potrf_timeline.md) overlaps the predecessor
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(1); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = POTF2_SMEM; cfg.stream = D;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            unsigned long long *tr = w.trace ? w.trace + 8 * jb : nullptr;
            CVXB_CUDA(cudaLaunchKernelEx(&cfg, potf2_inv_kernel, Ajj, (long long)lda, wj, invj, invTj, w.d_info, j,
                                         (long long)0, (long long)0, Tprev, (long long)lda, invprev, tr));
        } else {
            potf2_inv_kernel<<<1, 256, POTF2_SMEM, D>>>(Ajj, lda, wj, invj, invTj, w.d_info, j, 0, 0, Tprev,
                                                         lda, invprev, w.trace ? w.trace + 8 * jb : nullptr);
        }
        count_launch();
        CVXB_LAUNCH_CHECK();
        CVXB_CUDA(cudaEventRecord(w.ev_dg[jb], D));
        // L(j:, j-1) goes back into A once Dg(jb) has consumed the raw tile; the copy rides on the
        // panel stream T behind this step's TRSM / column update
        auto copy_back_prev = [&]() -> int {
            if (jb == 0) return 0;
            const int jp = j - NB, mp = n - j;
            const int pb = jb - 1;
            const double *src = pair_prev ? w.panel[(pb >> 1) & 1] + ((pb & 1) ? (long long)NB * ldw + NB : 0) : w.panel[pb & 1];
            CVXB_CUDA(cudaMemcpy2DAsync(A + j + (long long)jp * lda, (size_t)lda * sizeof(double),
                                        src, (size_t)ldw * sizeof(double),
                                        (size_t)mp * sizeof(double), NB, cudaMemcpyDeviceToDevice, T));
            return 0;
        };
        if (m <= 0) {
            CVXB_CUDA(cudaStreamWaitEvent(T, w.ev_dg[jb], 0));
            CVXB_TRY(copy_back_prev());
            break;
        }
        double *A21 = Ajj + wj;
        double *A22 = A21 + (long long)wj * lda;
        // ---- T: panel TRSM as a GEMM with the block inverse (out of place) ----
        CVXB_CUDA(cudaStreamWaitEvent(T, w.ev_dg[jb], 0));
        // the group buffer about to be overwritten was read by the bulk update two groups (steps) ago
        if (!odd && prev_r >= 0) CVXB_CUDA(cudaStreamWaitEvent(T, w.ev_r[prev_r], 0));
        if (!pair && pair_prev && last_r >= 0) CVXB_CUDA(cudaStreamWaitEvent(T, w.ev_r[last_r], 0));   // mode switch
        {
            GemmDesc g;
            g.M = m; g.N = wj; g.K = wj;
            g.X = A21; g.ldx = lda; g.x_kmajor = false;
            g.Y = invj; g.ldy = NB; g.y_kmajor = false;
            g.C = Wp; g.ldc = ldw;
            g.trace = w.trace ? w.trace + 8 * jb + 2 : nullptr;
            CVXB_TRY(dmma_gemm(g, T));
        }
        CVXB_CUDA(cudaEventRecord(w.ev_tr[jb], T));
        // every later write of this step into block columns >= jb+1 is ordered behind the last bulk update
        if (last_r >= 0) CVXB_CUDA(cudaStreamWaitEvent(T, w.ev_r[last_r], 0));
        const int wn = (m < NB) ? m : NB;          // width of block column jb+1
        // ---- T (even step of a pair): diagonal tile of block column jb+2 gets its panel-jb part now ----
        if (pair && !odd && m > NB) {
            const int w2 = (m - NB < NB) ? (m - NB) : NB;
            GemmDesc d0;
            d0.M = w2; d0.N = w2; d0.K = wj;
            d0.X = Wp + NB; d0.ldx = ldw; d0.x_kmajor = false;
            d0.Y = Wp + NB; d0.ldy = ldw; d0.y_kmajor = false;
            double *tile = A22 + NB + (long long)NB * lda;
            d0.D = tile; d0.ldd = lda; d0.C = tile; d0.ldc = lda;
            d0.alpha = -1.0; d0.beta = 1.0; d0.lower_only = true;
            CVXB_TRY(dmma_gemm(d0, T));
        }
        // ---- T: next block column, rows below its diagonal block ----
        if (m > wn) {
            GemmDesc c;
            c.M = m - wn; c.N = wn;
            if (odd) {          // both panels of the pair: rows of the group buffer, K = NB + wj
                c.K = NB + wj;
                c.X = Wo + wn; c.Y = Wo;
            } else {
                c.K = wj;
                c.X = Wp + wn; c.Y = Wp;
            }
            c.ldx = ldw; c.x_kmajor = false;
            c.ldy = ldw; c.y_kmajor = false;
            c.D = A22 + wn; c.ldd = lda; c.C = A22 + wn; c.ldc = lda;
            c.alpha = -1.0; c.beta = 1.0;
            c.trace = w.trace ? w.trace + 8 * jb + 4 : nullptr;
            CVXB_TRY(dmma_gemm(c, T));
        }
        // ---- T (odd step of a pair): block column jb+2 INCLUDING its diagonal tile gets both panels now, so that
        // the diagonal chain two steps ahead does not wait for this pair's bulk update
        if (odd && m > NB) {
            GemmDesc c1;
            c1.M = m; c1.N = m; c1.K = NB + wj;
            c1.X = Wo; c1.Y = Wo; c1.ldx = ldw; c1.x_kmajor = false; c1.ldy = ldw; c1.y_kmajor = false;
            c1.D = A22; c1.ldd = lda; c1.C = A22; c1.ldc = lda;
            c1.alpha = -1.0; c1.beta = 1.0; c1.lower_only = true;
            c1.ct_begin = panel_tiles; c1.ct_end = 2 * panel_tiles;
            CVXB_TRY(dmma_gemm(c1, T));
        }
        CVXB_CUDA(cudaEventRecord(w.ev_c0[jb], T));
        CVXB_TRY(copy_back_prev());
        // ---- U: the rest of the trailing matrix ----
        if ((pair ? (odd && m > 2 * NB) : (m > NB))) {
            CVXB_CUDA(cudaStreamWaitEvent(U, w.ev_tr[jb], 0));
            GemmDesc u;
            u.M = m; u.N = m;
            if (odd) { u.K = NB + wj; u.X = Wo; u.Y = Wo; }
            else { u.K = wj; u.X = Wp; u.Y = Wp; }
            u.ldx = ldw; u.x_kmajor = false;
            u.ldy = ldw; u.y_kmajor = false;
            u.D = A22; u.ldd = lda; u.C = A22; u.ldc = lda;
            u.alpha = -1.0; u.beta = 1.0; u.lower_only = true;
            u.ct_begin = odd ? 2 * panel_tiles : panel_tiles; u.ct_end = 1 << 30;
            u.trace = w.trace ? w.trace + 8 * jb + 6 : nullptr;
            static int tu_on = -1;
            // measured (profiles/r02x): correct, but ~35 us per 128x128 tile against 16.8 us of DMMA work -- potrf 12.0 ms
            // with it, 10.5 ms without -- so it is an opt-in experiment (CVXB_CHOL_TU=1), the generic GEMM stays default
            if (tu_on < 0) { const char *e = getenv("CVXB_CHOL_TU"); tu_on = (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0; }
            const bool tu = tu_on && !odd && !pair && wj == NB && (lda & 1) == 0 && (ldw & 1) == 0 &&
                            ((reinterpret_cast<uintptr_t>(A22) & 15) == 0) && ((reinterpret_cast<uintptr_t>(Wp) & 15) == 0);
            if (tu) {
                // persistent strip kernel (chol_trailing_kernel): block columns 1 .. nbk-1 of the trailing matrix
                const int nbk = (m + NB - 1) / NB;
                const long long Tt = (long long)(nbk - 1) * nbk / 2;
                if (tu_on == 2) {
                    // 128 x 64 tiles, two CTAs per SM, ~6 tiles (~3 of the 128 x 128 ones) per CTA
                    long long T2 = 0;
                    for (int c2 = 2; c2 < 2 * nbk; ++c2) T2 += nbk - (c2 >> 1);
                    if (T2 > 0) {
                        static int tpc2 = -1;
                        if (tpc2 < 0) { const char *e = getenv("CVXB_CHOL_TU_TILES"); tpc2 = e ? std::max(1, atoi(e)) : 6; }
                        const long long slots = 2LL * kNumSMs;
                        const long long waves = (T2 + slots * tpc2 - 1) / (slots * tpc2);
                        const int grid = (int)std::min<long long>(T2, waves * slots);
                        chol_trailing2_kernel<<<grid, 128, T2_SMEM, U>>>(m, Wp, ldw, A22, lda, nbk, T2, u.trace);
                        count_launch();
                        CVXB_LAUNCH_CHECK();
                    }
                } else if (Tt > 0) {
                    // CTAs of ~3 tiles (~50 us): long enough to amortise the strip operand and the pipeline fill, short
                    // enough that the SMs keep coming free for the chain / panel streams' kernels (a CTA of this
                    // kernel fills an SM's shared memory: nothing else can be resident beside it)
                    static int tpc = -1;
                    if (tpc < 0) { const char *e = getenv("CVXB_CHOL_TU_TILES"); tpc = e ? std::max(1, atoi(e)) : 3; }
                    const long long waves = (Tt + (long long)kNumSMs * tpc - 1) / ((long long)kNumSMs * tpc);
                    const int grid = (int)std::min<long long>(Tt, waves * kNumSMs);
                    chol_trailing_kernel<<<grid, 256, TU_SMEM, U>>>(m, Wp, ldw, A22, lda, nbk, 1, Tt, u.trace);
                    count_launch();
                    CVXB_LAUNCH_CHECK();
                }
            } else {
                CVXB_TRY(dmma_gemm(u, U));
            }
            CVXB_CUDA(cudaEventRecord(w.ev_r[jb], U));
            prev_r = last_r;
            last_r = jb;
        }
    }
    w.r_valid[0] = w.r_valid[1] = -1;
    CVXB_CUDA(cudaEventRecord(w.ev_end_p, D));
    CVXB_CUDA(cudaEventRecord(w.ev_end_u, U));
    CVXB_CUDA(cudaEventRecord(w.ev_end_t, T));
    CVXB_CUDA(cudaStreamWaitEvent(st, w.ev_end_p, 0));
    CVXB_CUDA(cudaStreamWaitEvent(st, w.ev_end_u, 0));
    CVXB_CUDA(cudaStreamWaitEvent(st, w.ev_end_t, 0));
    return 0;
}

// The ~450 launches / event operations of one factorisation are captured once per (n, A, lda,
// inv) into a CUDA graph (three-stream fork/join included) and replayed afterwards: dependent
// kernels then start without host-side launch latency, which is what the diagonal chain of
// small kernels is sensitive to.  First call with a new key runs eagerly (allocations, function
// attributes), the second one is captured; any capture failure falls back to eager launches.
// CVXB_GRAPH=0 disables.
int potrf_lower(int n, double *A, int lda, double *inv, CholWork &w, cudaStream_t st) {
    if (n <= 0) return 0;
    static int enabled = -1;
    if (enabled < 0) {
        const char *e = getenv("CVXB_GRAPH");
        enabled = (e && e[0] == '0') ? 0 : 1;
    }
    // measured (B200): replay is 1-2 % faster up to n = 4096 and removes ~450 host API calls per
    // factorisation; at n = 8192 it is 9 % slower (graph kernel nodes do not keep the panel streams'
    // priority over the bulk update), so large factorisations stay on the eager path
    if (!enabled || w.graph_failed || n > 4096) return potrf_enqueue(n, A, lda, inv, w, st);
    CholWork::GraphEntry *ent = nullptr;
    for (auto &g : w.graphs)
        if (g.n == n && g.A == A && g.lda == lda && g.inv == inv) { ent = &g; break; }
    if (ent && ent->exec) {
        CVXB_CUDA(cudaGraphLaunch(ent->exec, st));
        count_launch(ent->launches);
        return 0;
    }
    if (!ent) {                                        // first sight: eager, remember the key
        if (w.graphs.size() >= 4) {
            for (auto &g : w.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
            w.graphs.clear();
        }
        CholWork::GraphEntry g;
        g.n = n; g.A = A; g.lda = lda; g.inv = inv;
        w.graphs.push_back(g);
        return potrf_enqueue(n, A, lda, inv, w, st);
    }
    // second call with this key: capture
    const unsigned long long l0 = g_launches.load();
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        w.graph_failed = true;
        return potrf_enqueue(n, A, lda, inv, w, st);
    }
    const int rc = potrf_enqueue(n, A, lda, inv, w, st);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    const int captured = (int)(g_launches.load() - l0);
    g_launches.fetch_sub((unsigned long long)captured);    // nothing has run yet
    if (rc != 0 || ce != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        w.graph_failed = true;
        return potrf_enqueue(n, A, lda, inv, w, st);
    }
    cudaGraphExec_t exec = nullptr;
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess || !exec) {
        cudaGetLastError();
        cudaGraphDestroy(graph);
        w.graph_failed = true;
        return potrf_enqueue(n, A, lda, inv, w, st);
    }
    cudaGraphDestroy(graph);
    ent->exec = exec;
    ent->launches = captured;
    CVXB_CUDA(cudaGraphLaunch(exec, st));
    count_launch(captured);
    return 0;
}

int trsv_lower(int n, const double *L, int ldl, const double *inv, double *b, bool trans,
               CholWork &w, cudaStream_t st, int batch, long long sL, long long sInv, long long sb) {
    if (n <= 0 || batch <= 0) return 0;
    const int nblk = (n + NB - 1) / NB;
    if ((long long)nblk * batch > w.flags_cap) {
        if (w.d_flags) CVXB_CUDA(cudaFree(w.d_flags));
        w.flags_cap = (long long)nblk * batch;
        CVXB_CUDA(cudaMalloc(&w.d_flags, (size_t)w.flags_cap * sizeof(int)));
        CVXB_CUDA(cudaMemset(w.d_flags, 0, (size_t)w.flags_cap * sizeof(int)));
    }
    const int epoch = g_trsv_epoch.fetch_add(1) + 1;
    const double *invT = inv + (long long)nblk * NB * NB;
    const bool vec = ((uintptr_t)L % 16 == 0) && (ldl % 2 == 0) && (batch == 1 || sL % 2 == 0);
    dim3 grid(nblk, batch);
#define TRSV_LAUNCH(T, V) trsv_kernel<T, V><<<grid, 256, TRSV_SMEM, st>>>(n, L, ldl, inv, invT, b, w.d_flags, epoch, sL, sInv, sb)
    if (trans) { if (vec) TRSV_LAUNCH(true, true); else TRSV_LAUNCH(true, false); }
    else       { if (vec) TRSV_LAUNCH(false, true); else TRSV_LAUNCH(false, false); }
#undef TRSV_LAUNCH
    count_launch();
    CVXB_LAUNCH_CHECK();
    return 0;
}

int potrs_lower(int n, const double *L, int ldl, const double *inv, double *b, CholWork &w,
                cudaStream_t st, int batch, long long sL, long long sInv, long long sb) {
    CVXB_TRY(trsv_lower(n, L, ldl, inv, b, false, w, st, batch, sL, sInv, sb));
    CVXB_TRY(trsv_lower(n, L, ldl, inv, b, true, w, st, batch, sL, sInv, sb));
    return 0;
}

namespace {
__global__ void copy2d_batched_kernel(const double *src, long long lds, long long ssrc, double *dst,
                                      long long ldd, long long sdst, int rows, int cols) {
    const long long pb = blockIdx.z;
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r < rows && c < cols) dst[pb * sdst + r + (long long)c * ldd] = src[pb * ssrc + r + (long long)c * lds];
}
}  // namespace

// Batched Cholesky of `batch` independent n x n matrices (stride sA, inverse blocks stride
// sInv).  The batch itself fills the machine, so the steps run back to back on one stream
// (no look-ahead).  d_info: one int per problem.  panel: batch * ldw * NB doubles.
int potrf_lower_batched(int n, double *A, int lda, long long sA, double *inv, long long sInv,
                        int batch, int *d_info, double *panel, int ldw, cudaStream_t st) {
    if (n <= 0 || batch <= 0) return 0;
    const int nblk = (n + NB - 1) / NB;
    CVXB_CUDA(cudaMemsetAsync(d_info, 0, (size_t)batch * sizeof(int), st));
    const long long sW = (long long)ldw * NB;
    for (int jb = 0; jb < nblk; ++jb) {
        const int j = jb * NB;
        const int wj = (n - j < NB) ? (n - j) : NB;
        const int m = n - j - wj;
        double *Ajj = A + j + (long long)j * lda;
        double *invj = inv + (long long)jb * NB * NB;
        double *invTj = inv + (long long)(nblk + jb) * NB * NB;
        potf2_inv_kernel<<<batch, 256, POTF2_SMEM, st>>>(Ajj, lda, wj, invj, invTj, d_info, j, sA, sInv, nullptr, 0,
                                                        nullptr, nullptr);
        count_launch();
        CVXB_LAUNCH_CHECK();
        if (m <= 0) break;
        double *A21 = Ajj + wj;
        double *A22 = A21 + (long long)wj * lda;
        GemmDesc g;
        g.M = m; g.N = wj; g.K = wj;
        g.X = A21; g.ldx = lda; g.x_kmajor = false; g.sX = sA;
        g.Y = invj; g.ldy = NB; g.y_kmajor = false; g.sY = sInv;
        g.C = panel; g.ldc = ldw; g.sC = sW; g.batch = batch;
        CVXB_TRY(dmma_gemm(g, st));
        GemmDesc u;
        u.M = m; u.N = m; u.K = wj;
        u.X = panel; u.ldx = ldw; u.x_kmajor = false; u.sX = sW;
        u.Y = panel; u.ldy = ldw; u.y_kmajor = false; u.sY = sW;
        u.D = A22; u.ldd = lda; u.sD = sA; u.C = A22; u.ldc = lda; u.sC = sA;
        u.alpha = -1.0; u.beta = 1.0; u.lower_only = true; u.batch = batch;
        CVXB_TRY(dmma_gemm(u, st));
        dim3 cg((m + 255) / 256, wj, batch);
        copy2d_batched_kernel<<<cg, 256, 0, st>>>(panel, ldw, sW, A21, lda, sA, m, wj);
        count_launch();
        CVXB_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace cvxb
