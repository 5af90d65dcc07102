// fp64 SYRK  C = A' diag(d)^2 A (+ H)  on the int8 tensor path (tcgen05.mma kind::i8, int32
// accumulators in TMEM) by error-free slicing (Ozaki scheme).  EXPERIMENTAL, opt-in: the product
// path (kkt_api.cu) stays on the DMMA kernel unless CVXB_OZAKI=1.  Replaces the same reference call
// as the DMMA SYRK: blas.syrk(Gs, K, trans='T') in misc.kkt_chol.factor (misc.py:1275, blas.c:3039).
//
// Arithmetic.  Column j of Gs = diag(d) A is written  Gs[k,j] = 2^e_j * sum_s q_s[k,j] 2^-(6+7s),
// q_s integers in [-64, 64] (round-to-nearest digits, radix 2^7, e_j from the column maximum), so
//   C[i,j] = 2^(e_i+e_j-12) * sum_d 2^(-7d) * ( sum_{s+t=d} q_s[:,i] . q_t[:,j] ),   d = 0 .. S-1.
// Every inner product is exact in int32 ((d+1) * 4096 * K < 2^31 for K <= 32768 rows per drain);
// levels d >= S are dropped (S = 9: 62 bits below the column maximum, see
// profiles/r01_ozaki_slice_accuracy_study.md).  Four levels live in TMEM at a time (4 x 128 columns),
// so an output tile takes ceil(S/4) passes over K; inside a pass the four levels are combined
// exactly in fp64 (30 + 21 bits), scaled by powers of two and added to C.
//
// Data layout.  The slicing kernel writes the digits directly as the shared-memory image the MMA
// reads: a "unit" is the 128-column x 32-row (K) block of one slice, 4 KB, K-major with the 32-byte
// swizzle; units are ordered [column block][k step][slice] so that the first nS slices of one
// (block, k step) are one contiguous bulk copy.  One CTA per 128x128 lower tile: warp 0 producer
// (cp.async.bulk + mbarrier ring), warp 1 MMA issuer, warps 2-5 epilogue (TMEM -> fp64 -> C).
#include "common.cuh"
#include <cstdint>
#include <cstdlib>
#include <cmath>

namespace cvxb {

namespace {

constexpr int OZ_SMAX = 9;
constexpr int OZ_T = 128;                 // tile edge (rows and columns of C)
constexpr int OZ_KS = 32;                 // K rows per MMA (32 bytes of int8)
constexpr int OZ_UNIT = OZ_T * OZ_KS;     // 4096 B
constexpr int OZ_STAGES = 3;
constexpr int OZ_STAGE_BYTES = 2 * OZ_SMAX * OZ_UNIT;          // A slices | B slices
constexpr int OZ_SMEM = OZ_STAGES * OZ_STAGE_BYTES + 1024 + 256;
constexpr int OZ_THREADS = 192;
constexpr int OZ_KRANGE = 32768 / OZ_KS;  // k steps per drain (int32 overflow bound)

// byte offset of (row r of the unit = column of A, k byte kb) inside a 4 KB unit
__host__ __device__ inline int oz_unit_off(int layout, int r, int kb) {
    if (layout == 0)        // K-major, 32-byte swizzle: 16-byte chunk index ^= address bit 7
        return r * 32 + ((((kb >> 4) ^ (r >> 2)) & 1) << 4) + (kb & 15);
    // no swizzle, 8x16-byte core matrices: [row group][k chunk][row in group][16 B]
    return (r >> 3) * 256 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15);
}

struct OzParams {
    const uint8_t *Q;          // units, [nblk][nk][S][4096]
    const double *cs;          // 2^e_j per column
    const double *D; long long ldd;
    double *C; long long ldc;
    double beta;
    int n, nblk, nk, S;
    int layout;                // 0: SW32; 1: none (SBO 256, LBO 128); 2: none (SBO 128, LBO 256)
    unsigned int *dbg;         // optional progress words (mapped host memory) or nullptr
};

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]', int8 x int8 -> int32, M = N = 128, K = 32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void dbg_put(const OzParams &p, int slot, unsigned int v) {
    if (p.dbg && blockIdx.x == 0) {
        *reinterpret_cast<volatile unsigned int *>(p.dbg + slot) = v;
        __threadfence_system();
    }
}

// instruction descriptor (cute::UMMA::InstrDescriptor bit layout): dense, no saturate,
// D = s32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, N/8 << 17, M/16 << 24
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((OZ_T / 8) << 17) | ((OZ_T / 16) << 24);

__global__ void __launch_bounds__(OZ_THREADS, 1) oz_mma_kernel(OzParams p) {
    extern __shared__ uint8_t oz_smem_raw[];
    __shared__ uint32_t tmem_base_sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OZ_STAGES * OZ_STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + OZ_STAGES, *acc_full = bars + 2 * OZ_STAGES,
             *acc_empty = bars + 2 * OZ_STAGES + 1;

    // lower-triangular tile (I >= J) of this CTA
    const int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((long long)I * (I + 1) / 2 > t) --I;
    while ((long long)(I + 1) * (I + 2) / 2 <= t) ++I;
    const int J = t - (int)((long long)I * (I + 1) / 2);

    if (tid == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    dbg_put(p, 0, 0x100u + (tid == 0));

    const int S = p.S;
    const int npass = (S + 3) / 4;
    const int nrange = (p.nk + OZ_KRANGE - 1) / OZ_KRANGE;

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer: one bulk copy per operand per k step =====
            int st = 0; uint32_t ph = 0;
            for (int rg = 0; rg < nrange; ++rg) {
                const int k0 = rg * OZ_KRANGE, k1 = min(p.nk, k0 + OZ_KRANGE);
                for (int ps = 0; ps < npass; ++ps) {
                    const int nS = min(S, 4 * ps + 4);
                    const uint32_t bytes = (uint32_t)nS * OZ_UNIT;
                    for (int kc = k0; kc < k1; ++kc) {
                        mbar_wait(empty + st, ph ^ 1);
                        mbar_expect_tx(full + st, 2 * bytes);
                        uint8_t *sa = smem + st * OZ_STAGE_BYTES;
                        bulk_g2s(sa, p.Q + ((size_t)I * p.nk + kc) * (size_t)S * OZ_UNIT, bytes, full + st);
                        bulk_g2s(sa + OZ_SMAX * OZ_UNIT, p.Q + ((size_t)J * p.nk + kc) * (size_t)S * OZ_UNIT, bytes, full + st);
                        if (++st == OZ_STAGES) { st = 0; ph ^= 1; }
                    }
                }
            }
            dbg_put(p, 1, 0x200u);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            uint32_t hi, lbo16;
            if (p.layout == 0) { hi = (256u >> 4) | (1u << 14) | (6u << 29); lbo16 = 0; }
            else if (p.layout == 1) { hi = (256u >> 4) | (1u << 14); lbo16 = 128u >> 4; }
            else { hi = (128u >> 4) | (1u << 14); lbo16 = 256u >> 4; }
            int st = 0; uint32_t ph = 0;
            int g = 0;                                   // global pass counter (for the accumulator barriers)
            for (int rg = 0; rg < nrange; ++rg) {
                const int k0 = rg * OZ_KRANGE, k1 = min(p.nk, k0 + OZ_KRANGE);
                for (int ps = 0; ps < npass; ++ps, ++g) {
                    const int nS = min(S, 4 * ps + 4), d0 = 4 * ps, d1 = min(S - 1, d0 + 3);
                    if (g > 0) { mbar_wait(acc_empty, (uint32_t)(g - 1) & 1); tc_fence_after(); }
                    uint32_t touched = 0;
                    for (int kc = k0; kc < k1; ++kc) {
                        mbar_wait(full + st, ph);
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(smem + st * OZ_STAGE_BYTES);
                        const uint32_t b0 = a0 + OZ_SMAX * OZ_UNIT;
                        for (int s = 0; s < nS; ++s) {
                            const int tlo = max(0, d0 - s), thi = min(nS - 1, d1 - s);
                            const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)((((a0 + s * OZ_UNIT) >> 4) & 0x3FFFu) | (lbo16 << 16));
                            for (int tt = tlo; tt <= thi; ++tt) {
                                const int lev = s + tt - d0;
                                const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)((((b0 + tt * OZ_UNIT) >> 4) & 0x3FFFu) | (lbo16 << 16));
                                tc_mma_i8(tmem_base + lev * OZ_T, da, db, OZ_IDESC, (touched >> lev) & 1u);
                                touched |= 1u << lev;
                            }
                        }
                        tc_commit(empty + st);           // frees the stage when these MMAs have read it
                        if (++st == OZ_STAGES) { st = 0; ph ^= 1; }
                    }
                    tc_commit(acc_full);
                    dbg_put(p, 2, 0x300u + g);
                }
            }
        }
    } else {
        // ===== epilogue: 4 warps, warp w owns TMEM lanes 32*(w%4) .. +31 (rows of the tile) =====
        const int quad = warp & 3;
        const int row = I * OZ_T + quad * 32 + lane;
        const bool row_ok = row < p.n;
        const double rsc = row_ok ? p.cs[row] : 0.0;
        int g = 0;
        for (int rg = 0; rg < nrange; ++rg) {
            for (int ps = 0; ps < npass; ++ps, ++g) {
                const int d0 = 4 * ps, nlev = min(4, S - d0);
                mbar_wait(acc_full, (uint32_t)g & 1);
                tc_fence_after();
                // 2^(-12 - 7*(d0 + nlev - 1)): scale of the last level of the pass
                const double lsc = ldexp(1.0, -12 - 7 * (d0 + nlev - 1));
                const bool first = (g == 0);
                for (int c0 = 0; c0 < OZ_T; c0 += 16) {
                    uint32_t r[4][16];
                    const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + c0;
                    tc_ld16(ta, r[0]);
                    if (nlev > 1) tc_ld16(ta + OZ_T, r[1]);
                    if (nlev > 2) tc_ld16(ta + 2 * OZ_T, r[2]);
                    if (nlev > 3) tc_ld16(ta + 3 * OZ_T, r[3]);
                    tc_wait_ld();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int col = J * OZ_T + c0 + c;
                        double v = (double)(int)r[0][c];
                        if (nlev > 1) v = fma(v, 128.0, (double)(int)r[1][c]);
                        if (nlev > 2) v = fma(v, 128.0, (double)(int)r[2][c]);
                        if (nlev > 3) v = fma(v, 128.0, (double)(int)r[3][c]);
                        if (row_ok && col < p.n && col <= row) {
                            v = (v * lsc) * rsc * p.cs[col];
                            double *dst = p.C + row + (size_t)col * p.ldc;
                            if (first) v += (p.D ? p.beta * p.D[row + (size_t)col * p.ldd] : 0.0);
                            else v += *dst;
                            *dst = v;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);
                if (quad == 0 && lane == 0) dbg_put(p, 3, 0x400u + g);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// amax[j] = max_k |d[k] * A[k,j]|  ->  cs[j] = 2^e_j, sinv[j] = 64 * 2^-e_j
__global__ void oz_colscale_kernel(int m, int n, const double *A, long long lda, const double *d, double *cs, double *sinv) {
    __shared__ double sh[32];
    const int j = blockIdx.x;
    const double *a = A + (size_t)j * lda;
    double mx = 0.0;
    for (int k = threadIdx.x; k < m; k += blockDim.x) mx = fmax(mx, fabs((d ? d[k] : 1.0) * a[k]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmax(mx, sh[w]);
        int e = 0;
        if (mx > 0.0 && mx < INFINITY) frexp(mx, &e);         // mx = f * 2^e, f in [0.5, 1): |x| < 2^e
        cs[j] = ldexp(1.0, e);
        sinv[j] = ldexp(64.0, -e);
    }
}

// digits of column block cb, k steps [kc0, kc0 + gridDim.x chunk): thread r = column cb*128 + r
__global__ void __launch_bounds__(OZ_T) oz_slice_kernel(int m, int n, const double *A, long long lda, const double *d,
                                                        const double *sinv, uint8_t *Q, int nk, int S, int layout) {
    const int cb = blockIdx.y, kc = blockIdx.x, r = threadIdx.x;
    const int j = cb * OZ_T + r;
    const bool col_ok = j < n;
    const double sc = col_ok ? sinv[j] : 0.0;
    const double *a = A + (size_t)(col_ok ? j : 0) * lda;
    uint8_t *unit0 = Q + ((size_t)cb * nk + kc) * (size_t)S * OZ_UNIT;
    for (int half = 0; half < 2; ++half) {
        uint32_t pk[OZ_SMAX][4];
#pragma unroll
        for (int s = 0; s < OZ_SMAX; ++s) pk[s][0] = pk[s][1] = pk[s][2] = pk[s][3] = 0u;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int k = kc * OZ_KS + half * 16 + e;
            double x = 0.0;
            if (col_ok && k < m) x = ((d ? d[k] : 1.0) * a[k]) * sc;      // |x| < 64
#pragma unroll
            for (int s = 0; s < OZ_SMAX; ++s) {
                if (s < S) {
                    const double q = rint(x);
                    x = (x - q) * 128.0;                                   // exact
                    pk[s][e >> 2] |= ((uint32_t)(int)q & 0xFFu) << ((e & 3) * 8);
                }
            }
        }
        const int off = oz_unit_off(layout, r, half * 16);
#pragma unroll
        for (int s = 0; s < OZ_SMAX; ++s)
            if (s < S)
                *reinterpret_cast<uint4 *>(unit0 + (size_t)s * OZ_UNIT + off) = make_uint4(pk[s][0], pk[s][1], pk[s][2], pk[s][3]);
    }
}

}  // namespace

size_t ozaki_workspace_bytes(int n, int m, int S) {
    const size_t nblk = (n + OZ_T - 1) / OZ_T, nk = (m + OZ_KS - 1) / OZ_KS;
    return nblk * nk * (size_t)S * OZ_UNIT + 2 * (size_t)n * sizeof(double) + 256;
}

// C(lower) = A' diag(d)^2 A + beta * D.  A: m x n (lda), d: m (or nullptr).  work: ozaki_workspace_bytes.
int ozaki_syrk(int n, int m, const double *A, long long lda, const double *d, const double *D, long long ldd,
               double beta, double *C, long long ldc, int S, int layout, void *work, unsigned int *dbg,
               cudaStream_t st) {
    if (n <= 0) return 0;
    if (S < 1 || S > OZ_SMAX || !A || !C || !work || m < 0) { set_error("ozaki_syrk: bad arguments"); return CVXB_E_ARG; }
    static bool attr = false;
    if (!attr) {
        CVXB_CUDA(cudaFuncSetAttribute(oz_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        attr = true;
    }
    const int nblk = (n + OZ_T - 1) / OZ_T, nk = std::max(1, (m + OZ_KS - 1) / OZ_KS);
    double *cs = reinterpret_cast<double *>(work);
    double *sinv = cs + n;
    uint8_t *Q = reinterpret_cast<uint8_t *>(((uintptr_t)(sinv + n) + 255) & ~uintptr_t(255));
    oz_colscale_kernel<<<n, 256, 0, st>>>(m, n, A, lda, d, cs, sinv);
    count_launch();
    oz_slice_kernel<<<dim3(nk, nblk), OZ_T, 0, st>>>(m, n, A, lda, d, sinv, Q, nk, S, layout);
    count_launch();
    OzParams p;
    p.Q = Q; p.cs = cs; p.D = D; p.ldd = ldd; p.C = C; p.ldc = ldc; p.beta = beta;
    p.n = n; p.nblk = nblk; p.nk = nk; p.S = S; p.layout = layout; p.dbg = dbg;
    const long long tiles = (long long)nblk * (nblk + 1) / 2;
    oz_mma_kernel<<<(unsigned)tiles, OZ_THREADS, OZ_SMEM, st>>>(p);
    count_launch();
    CVXB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cvxb
