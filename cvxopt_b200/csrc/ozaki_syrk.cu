// fp64 SYRK  C = A' diag(d)^2 A (+ H)  on the int8 tensor path (tcgen05.mma kind::i8, int32
// accumulators in TMEM) by error-free slicing (Ozaki scheme).  kkt_api.cu uses it for the 'l'-row SYRK
// of large problems (n >= 4096, ml >= 8192; CVXB_OZAKI=0 keeps the DMMA kernel, =2 forces it at any
// size).  Replaces the same reference call as the DMMA SYRK: blas.syrk(Gs, K, trans='T') in
// misc.kkt_chol.factor (misc.py:1275, blas.c:3039).  Measured: 14.1 ms at n=8192, m=16384 (18.5 ms at the end of
// round 1) against 33.7 ms on the fp64 DMMA pipe, result within 4e-16 * sum|terms| of an 80-bit evaluation.
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
// (cp.async.bulk + mbarrier ring), warp 1 MMA issuer, warps 2-9 epilogue (TMEM -> fp64 -> C).
#include "common.cuh"
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cmath>
#include <vector>
#include <algorithm>

namespace cvxb {

namespace {

constexpr int OZ_SMAX = 9;
constexpr int OZ_T = 128;                 // tile edge (rows and columns of C)
constexpr int OZ_KS = 32;                 // K rows per MMA (32 bytes of int8)
constexpr int OZ_UNIT = OZ_T * OZ_KS;     // 4096 B
constexpr int OZ_RING = 54;                // 4 KB units in the shared-memory ring (216 KB)
constexpr int OZ_MAXST = 8;                // most stages a pass may split the ring into
constexpr int OZ_MAXPASS = 5;
constexpr int OZ_SMEM = OZ_RING * OZ_UNIT + 1024 + 256;
constexpr int OZ_THREADS = 192;            // two-SM kernels: producer, issuer, 4 epilogue warps
constexpr int OZ_THREADS1 = 320;           // one-SM kernel: producer, issuer, 8 epilogue warps (two per TMEM lane quadrant)
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
    int npass;                 // level groups: pass ps accumulates levels pd0[ps] .. pd1[ps] (<= 4 of them)
    int pd0[OZ_MAXPASS], pd1[OZ_MAXPASS];
    const unsigned int *tiles; // (I << 16) | J per CTA, in launch order
    unsigned int *dbg;         // optional progress words (mapped host memory) or nullptr
    // split-K launch of the tail wave: CTA b works on tile b / nchunk over k steps [c*kper, (c+1)*kper), c = b % nchunk,
    // and writes its 128 x 128 partial (tile-local, column-major) to part + b * 128*128; oz_tail_reduce_kernel sums them
    int nchunk, kper;
    double *part;
    int kb2, kb5;              // k steps per ring stage for passes with <= 2 / <= 5 slices (CVXB_OZ_KB=a,b)
    int trace_cta;             // diagnostics: CTA whose pass timeline goes to dbg (or -1)
    int ablate;                // diagnostics (CVXB_OZ_ABLATE): 1 = no operand copies, 2 = no MMAs (results are then garbage)
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
// one lane of a converged warp (the pattern the compiler recognises for single-thread tcgen05 issue)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// pass timeline of one CTA (CVXB_OZ_TRACE_CTA=<launch index>): globaltimer ns (low 32 bits) in dbg[16 + slot]
__device__ __forceinline__ void dbg_stamp(const OzParams &p, int slot) {
    if (p.dbg && p.trace_cta >= 0 && (int)blockIdx.x == p.trace_cta && slot < 48) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        *reinterpret_cast<volatile unsigned int *>(p.dbg + 16 + slot) = (unsigned int)t;
    }
}
__device__ __forceinline__ void dbg_put(const OzParams &p, int slot, unsigned int v) {
    if (p.dbg && blockIdx.x == 0) {
        *reinterpret_cast<volatile unsigned int *>(p.dbg + slot) = v;
        __threadfence_system();
    }
}

// instruction descriptor (cute::UMMA::InstrDescriptor bit layout): dense, no saturate,
// D = s32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, N/8 << 17, M/16 << 24
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((OZ_T / 8) << 17) | ((OZ_T / 16) << 24);

// One k step of a pass whose level range is known at compile time: a straight-line MMA sequence
// (the tensor pipe retires a 128x128x32 int8 MMA every 64 cycles; a runtime pair loop issues too slowly).
// The first k step of a pass overwrites each accumulator with its first product (s == 0), later ones add.
// N = 256 instruction descriptor: one MMA multiplies A slice s with TWO consecutive B slices (t, t+1), which
// sit back to back in shared memory and whose products belong to the adjacent levels s+t, s+t+1 = adjacent
// 128-column accumulators.  The tensor pipe reads its operands from shared memory: a 128x128x32 int8 MMA
// reads 8 KB per 64 cycles = the full 128 B/cycle of the SM, so the bulk copies that refill the ring compete
// with it (the one-slice-pair kernel was 64 % busy, profiles/r01k); a 128x256x32 MMA reads 12 KB per 128 cycles.
constexpr uint32_t OZ_IDESC_N256 = (2u << 4) | (1u << 7) | (1u << 10) | ((2 * OZ_T / 8) << 17) | ((OZ_T / 16) << 24);

// k steps per ring stage: the producer / issuer handshake (barrier wait, commit, barrier wait) costs ~500 cycles per
// stage (measured with copies and MMAs disabled, profiles/r02p), more than the 192 cycles of tensor work one k step
// of pass {0,1} holds, so the passes with few slices put several k steps behind one handshake
__device__ __forceinline__ int oz_kb(const OzParams &p, int nS) { return nS <= 2 ? p.kb2 : (nS <= 5 ? p.kb5 : 1); }

template <int D0, int D1, bool FIRST>
__device__ __forceinline__ void oz_issue_step(uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t tmem_base) {
    constexpr int NS = D1 + 1;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)(a_lo + s * (OZ_UNIT >> 4));
        // B slices t with D0 <= s + t <= D1, 0 <= t < NS: a compile-time range, walked two at a time
        const int tlo = (D0 - s) > 0 ? (D0 - s) : 0;
        const int thi = (D1 - s) < (NS - 1) ? (D1 - s) : (NS - 1);
#pragma unroll
        for (int tt = 0; tt < NS; tt += 1) {
            if (tt < tlo || tt > thi) continue;
            if (((tt - tlo) & 1) != 0) continue;                 // second half of a pair: already issued
            const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)(b_lo + tt * (OZ_UNIT >> 4));
            const uint32_t acc = (FIRST && s == 0) ? 0u : 1u;
            if (tt + 1 <= thi)
                tc_mma_i8(tmem_base + (s + tt - D0) * OZ_T, da, db, OZ_IDESC_N256, acc);
            else
                tc_mma_i8(tmem_base + (s + tt - D0) * OZ_T, da, db, OZ_IDESC, acc);
        }
    }
}

__global__ void __launch_bounds__(OZ_THREADS1, 1) oz_mma_kernel(OzParams p) {
    extern __shared__ uint8_t oz_smem_raw[];
    __shared__ uint32_t tmem_base_sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OZ_RING * OZ_UNIT);
    uint64_t *full = bars, *empty = bars + OZ_MAXST, *acc_full = bars + 2 * OZ_MAXST,
             *acc_empty = bars + 2 * OZ_MAXST + 1;

    // lower-triangular tile (I >= J) of this CTA
    const bool split = p.nchunk > 1;
    const unsigned int tl = p.tiles[split ? blockIdx.x / p.nchunk : blockIdx.x];
    const int I = (int)(tl >> 16), J = (int)(tl & 0xFFFFu);
    const int kb = split ? (int)(blockIdx.x % p.nchunk) * p.kper : 0;
    const int ke = split ? min(p.nk, kb + p.kper) : p.nk;

    if (tid == 0) {
        for (int s = 0; s < OZ_MAXST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
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
    if (tid == 0) dbg_stamp(p, 0);

    const int S = p.S;
    const int npass = p.npass;
    const int nrange = (ke - kb + OZ_KRANGE - 1) / OZ_KRANGE;
    // Stage geometry of a pass: nS slices of A then nS slices of B per k step; the ring is cut into
    // as many such stages as fit (deeper prefetch for the passes that need few slices).

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer: one bulk copy per operand per k step =====
            uint32_t filled = 0, epar = 0;               // per stage: ever filled / parity of its last release
            for (int rg = 0; rg < nrange; ++rg) {
                const int k0 = kb + rg * OZ_KRANGE, k1 = min(ke, k0 + OZ_KRANGE);
                for (int ps = 0; ps < npass; ++ps) {
                    const int nS = min(S, p.pd1[ps] + 1);
                    const int KB = oz_kb(p, nS);
                    const int nst = min(OZ_MAXST, OZ_RING / (2 * nS * KB));
                    const uint32_t bytes = (uint32_t)nS * OZ_UNIT;
                    // the new geometry overlaps the old stages: wait until every one of them is released
                    for (int s2 = 0; s2 < OZ_MAXST; ++s2)
                        if ((filled >> s2) & 1u) mbar_wait(empty + s2, (epar >> s2) & 1u);
                    int st = 0;
                    for (int kc = k0; kc < k1; kc += KB) {
                        const int kn = min(KB, k1 - kc);
                        if ((filled >> st) & 1u) { mbar_wait(empty + st, (epar >> st) & 1u); epar ^= 1u << st; }
                        else filled |= 1u << st;
                        if (p.ablate & 1) { mbar_arrive(full + st); if (++st == nst) st = 0; continue; }
                        mbar_expect_tx(full + st, 2 * bytes * (uint32_t)kn);
                        uint8_t *sa = smem + (size_t)st * 2 * bytes * KB;
                        for (int i = 0; i < kn; ++i, sa += 2 * bytes) {
                            bulk_g2s(sa, p.Q + ((size_t)I * p.nk + kc + i) * (size_t)S * OZ_UNIT, bytes, full + st);
                            bulk_g2s(sa + bytes, p.Q + ((size_t)J * p.nk + kc + i) * (size_t)S * OZ_UNIT, bytes, full + st);
                        }
                        if (++st == nst) st = 0;
                    }
                }
            }
            dbg_put(p, 1, 0x200u);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop (uniform control flow and addresses),
        // lane 0 issues the tcgen05 instructions =====
        uint32_t hi, lbo16;
        if (p.layout == 0) { hi = (256u >> 4) | (1u << 14) | (6u << 29); lbo16 = 0; }
        else if (p.layout == 1) { hi = (256u >> 4) | (1u << 14); lbo16 = 128u >> 4; }
        else { hi = (128u >> 4) | (1u << 14); lbo16 = 256u >> 4; }
        uint32_t cpar = 0;                           // per stage: parity of its next fill
        int g = 0;                                   // global pass counter (for the accumulator barriers)
        const uint32_t sbase = smem_u32(smem);
        for (int rg = 0; rg < nrange; ++rg) {
            const int k0 = kb + rg * OZ_KRANGE, k1 = min(ke, k0 + OZ_KRANGE);
            for (int ps = 0; ps < npass; ++ps, ++g) {
                const int d0 = p.pd0[ps], d1 = p.pd1[ps];
                const int nS = min(S, d1 + 1);
                const int KB = oz_kb(p, nS);
                const int nst = min(OZ_MAXST, OZ_RING / (2 * nS * KB));
                const uint32_t bytes = (uint32_t)nS * OZ_UNIT;
                const int code = (p.layout == 0 && nS == d1 + 1) ? d0 * 16 + d1 : -1;
                if (g > 0) { mbar_wait(acc_empty, (uint32_t)(g - 1) & 1); tc_fence_after(); }
                if (lane == 0) dbg_stamp(p, 1 + 4 * g);
                uint32_t touched = 0;
                int st = 0;
                for (int kc = k0; kc < k1; kc += KB) {
                    const int kn = min(KB, k1 - kc);
                    mbar_wait(full + st, (cpar >> st) & 1u);
                    cpar ^= 1u << st;
                    tc_fence_after();
                    if (elect_one()) {
                        for (int i = 0; i < kn; ++i) {
                            const uint32_t a0 = sbase + ((uint32_t)st * KB + (uint32_t)i) * 2u * bytes;
                            const uint32_t b0 = a0 + bytes;
                            const bool firstk = (kc + i == k0);
#define OZ_CASE(D0, D1)                                                                              \
    case (D0) * 16 + (D1):                                                                           \
        if (firstk) oz_issue_step<D0, D1, true>(a0 >> 4, b0 >> 4, hi, tmem_base);                    \
        else oz_issue_step<D0, D1, false>(a0 >> 4, b0 >> 4, hi, tmem_base);                          \
        break;
                            switch ((p.ablate & 2) ? -2 : code) {
                                case -2: break;
                                OZ_CASE(0, 0) OZ_CASE(1, 4) OZ_CASE(5, 8)          // S = 9: {0} {1..4} {5..8}
                                OZ_CASE(0, 1) OZ_CASE(2, 4)                         // S = 9: {0,1} {2..4} {5..8}
                                OZ_CASE(0, 3) OZ_CASE(4, 7)                         // S = 8: {0..3} {4..7}
                                OZ_CASE(2, 5) OZ_CASE(6, 8) OZ_CASE(8, 8)
                            default:
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
                            }
#undef OZ_CASE
                        }
                        tc_commit(empty + st);           // frees the stage when these MMAs have read it
                    }
                    __syncwarp();
                    if (++st == nst) st = 0;
                }
                if (lane == 0) dbg_stamp(p, 2 + 4 * g);
                if (elect_one()) { tc_commit(acc_full); dbg_put(p, 2, 0x300u + g); }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: 8 warps, warp w owns TMEM lanes 32*(w%4) .. +31 (rows of the tile) and one half of the
        // columns: the values C already holds come from HBM / L2 under the full load of the operand stream (2-8 us per
        // round trip, profiles/r02r), so the epilogue is bound by how many of those loads are in flight =====
        const int quad = warp & 3;
        const int cbeg = ((warp - 2) >> 2) * (OZ_T / 2), cend = cbeg + OZ_T / 2;
        const int row = I * OZ_T + quad * 32 + lane;
        const bool row_ok = row < p.n;
        const double rsc = row_ok ? p.cs[row] : 0.0;
        int g = 0;
        for (int rg = 0; rg < nrange; ++rg) {
            for (int ps = 0; ps < npass; ++ps, ++g) {
                const int d0 = p.pd0[ps], nlev = p.pd1[ps] - d0 + 1;
                // 2^(-12 - 7*(d0 + nlev - 1)): scale of the last level of the pass
                const double lsc = ldexp(1.0, -12 - 7 * (d0 + nlev - 1));
                const bool first = (g == 0);
                // The pass adds to what earlier passes left in C (or in the split-K partial tile): those values are
                // loaded 16 columns ahead, as one batch of independent loads issued before the TMEM loads are waited
                // for -- a load / add / store chain per column costs a global-memory latency per column.
                double *const part0 = split ? p.part + (size_t)blockIdx.x * (OZ_T * OZ_T) + (quad * 32 + lane) : nullptr;
                auto load_old = [&](int c0, double (&o)[16]) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int col = J * OZ_T + c0 + c;
                        double v = 0.0;
                        if (split) { if (!first) v = part0[(size_t)(c0 + c) * OZ_T]; }
                        else if (row_ok && col < p.n && col <= row) {
                            if (!first) v = p.C[row + (size_t)col * p.ldc];
                            else if (p.D) v = p.beta * p.D[row + (size_t)col * p.ldd];
                        }
                        o[c] = v;
                    }
                };
                double oldv[16], nxtv[16];
                load_old(cbeg, oldv);                    // in flight while the pass's last MMAs run
                mbar_wait(acc_full, (uint32_t)g & 1);
                tc_fence_after();
                if (warp == 4 && lane == 0) dbg_stamp(p, 3 + 4 * g);
                for (int c0 = cbeg; c0 < cend; c0 += 16) {
                    uint32_t r[4][16];
                    const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + c0;
                    tc_ld16(ta, r[0]);
                    if (nlev > 1) tc_ld16(ta + OZ_T, r[1]);
                    if (nlev > 2) tc_ld16(ta + 2 * OZ_T, r[2]);
                    if (nlev > 3) tc_ld16(ta + 3 * OZ_T, r[3]);
                    if (c0 + 16 < cend) load_old(c0 + 16, nxtv);
                    tc_wait_ld();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int col = J * OZ_T + c0 + c;
                        double v = (double)(int)r[0][c];
                        if (nlev > 1) v = fma(v, 128.0, (double)(int)r[1][c]);
                        if (nlev > 2) v = fma(v, 128.0, (double)(int)r[2][c]);
                        if (nlev > 3) v = fma(v, 128.0, (double)(int)r[3][c]);
                        if (split) {
                            v = (col < p.n) ? (v * lsc) * rsc * p.cs[col] : 0.0;
                            part0[(size_t)(c0 + c) * OZ_T] = v + oldv[c];
                        } else if (row_ok && col < p.n && col <= row) {
                            v = (v * lsc) * rsc * p.cs[col];
                            p.C[row + (size_t)col * p.ldc] = v + oldv[c];
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 16; ++c) oldv[c] = nxtv[c];
                }
                tc_fence_before();
                __syncwarp();
                if (warp == 4 && lane == 0) dbg_stamp(p, 4 + 4 * g);
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


// ------------------------------------------------------------------------------------------------
// Two-SM variant (tcgen05 cta_group::2): a cluster of two CTAs (the two SMs of one TPC) computes the
// 256 x 128 block { (I0, J), (I1 = I0 + 1, J) } with M = 256 MMAs issued by the leader (cluster rank 0).
// Each CTA stages its own 128 rows of the M operand (nS units of 4 KB per k step) and HALF of the N
// operand (rows 64*rank .. 64*rank+63 of every unit of column block J: nS pieces of 2 KB), so a pair moves
// 12 KB * nS per k step where two one-SM tiles move 16 KB * nS: the kernel is bound by operand ingest
// (profiles/r01k), this is the -25 % of DESIGN.md's round-2 plan.  Conventions follow the vendored CUTLASS
// sm100 headers (cute/arch/tmem_allocator_sm100.hpp Allocator2Sm, mma_sm100_umma.hpp SM100_MMA_S8_2x1SM_SS,
// cutlass/arch/barrier.h umma_arrive_multicast_2x1SM):
//   * tcgen05.alloc / dealloc / relinquish .cta_group::2 by the same warp (1) of BOTH CTAs;
//   * the leader alone issues tcgen05.mma.cta_group::2 (idesc M = 256); descriptors hold the leader's
//     shared-memory offsets, the peer's operands sit at the same offsets of its own shared memory;
//     D rows 0-127 land in the leader's TMEM, rows 128-255 in the peer's, same columns;
//   * stage-full: every CTA's bulk copies complete on its own mbarrier; the peer's warp 1 relays each
//     completion with a remote mbarrier.arrive on the leader's `pfull` barrier (shared::cluster address with
//     the CTA-rank bit cleared, cute's Sm100MmaPeerBitMask);
//   * stage-empty / accumulator-full: tcgen05.commit.cta_group::2 ... multicast::cluster with mask 0b11 arrives
//     on the barrier at the same offset in both CTAs; accumulator-empty: the 4 epilogue warps of both CTAs
//     arrive on the leader's barrier (count 8).
constexpr uint32_t OZ_PEER_MASK = 0xFEFFFFFFu;
constexpr uint32_t OZ_IDESC2 = (2u << 4) | (1u << 7) | (1u << 10) | ((OZ_T / 8) << 17) | ((2 * OZ_T / 16) << 24);
constexpr int OZ_HALF = OZ_UNIT / 2;       // 2 KB: 64 rows of a unit

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr) {     // addr: shared::cluster address
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_commit2_mc(uint64_t *bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void tc_mma2_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// WIDE form (N = 256): the pair computes the 256 x 256 block of row blocks {2a, 2a+1} x column blocks {2b, 2b+1};
// each CTA stages its 128 rows of the M operand and ITS 128 rows of the N operand (block 2b + rank, whole 4 KB
// units), the MMA is 256x256x32 (idesc N = 256) and an accumulator level takes 256 TMEM columns, so a pass
// holds two levels.  Per 128x128 of output a CTA ingests 25 x 4 KB per k step (passes {0} {1,2} {3,4} {5,6} {7,8})
// where the one-SM kernel ingests 32 x 4 KB (passes {0,1} {2..4} {5..8}).
constexpr uint32_t OZ_IDESC2W = (2u << 4) | (1u << 7) | (1u << 10) | ((2 * OZ_T / 8) << 17) | ((2 * OZ_T / 16) << 24);

// one k step of a pass (levels D0..D1): A slices are 4 KB units, B slices 2 KB half units (4 KB units if WIDE)
template <int D0, int D1, bool FIRST, bool WIDE>
__device__ __forceinline__ void oz_issue_step2(uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t tmem_base) {
    constexpr int NS = D1 + 1;
    constexpr int BU = WIDE ? OZ_UNIT : OZ_HALF;
    constexpr int LW = WIDE ? 2 * OZ_T : OZ_T;
    constexpr uint32_t ID = WIDE ? OZ_IDESC2W : OZ_IDESC2;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)(a_lo + s * (OZ_UNIT >> 4));
#pragma unroll
        for (int tt = 0; tt < NS; ++tt) {
            if (s + tt >= D0 && s + tt <= D1) {
                const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)(b_lo + tt * (BU >> 4));
                tc_mma2_i8(tmem_base + (s + tt - D0) * LW, da, db, ID, (FIRST && s == 0) ? 0u : 1u);
            }
        }
    }
}

template <bool WIDE>
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_mma2_kernel(OzParams p) {
    constexpr int BU = WIDE ? OZ_UNIT : OZ_HALF;          // bytes of one staged B slice
    constexpr int LW = WIDE ? 2 * OZ_T : OZ_T;            // TMEM columns of one accumulator level
    constexpr uint32_t ID = WIDE ? OZ_IDESC2W : OZ_IDESC2;
    extern __shared__ uint8_t oz_smem_raw[];
    __shared__ uint32_t tmem_base_sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OZ_RING * OZ_UNIT);
    uint64_t *full = bars, *empty = bars + OZ_MAXST, *pfull = bars + 2 * OZ_MAXST, *acc_full = bars + 3 * OZ_MAXST,
             *acc_empty = bars + 3 * OZ_MAXST + 1;

    // pair (a, J): row blocks 2a (leader) and 2a + 1 (peer), column block J  (WIDE: column blocks 2J, 2J + 1)
    const unsigned int tl = p.tiles[blockIdx.x >> 1];
    const int a = (int)(tl >> 16), J = (int)(tl & 0xFFFFu);
    int I = 2 * a + (int)rank;
    const bool live = I < p.nblk;                // odd nblk: the last pair has no second row block
    if (!live) I = 2 * a;                        // load valid data, store nothing
    const int Jb = WIDE ? min(2 * J + (int)rank, p.nblk - 1) : J;     // block whose rows this CTA stages for N
    const int Jc = WIDE ? 2 * J : J;                                  // first column block of the accumulators

    if (tid == 0) {
        for (int s = 0; s < OZ_MAXST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); mbar_init(pfull + s, 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();                          // both CTAs' barriers exist before anyone signals them
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    dbg_put(p, 0, 0x100u + (tid == 0));

    const int S = p.S;
    const int npass = p.npass;
    const int nrange = (p.nk + OZ_KRANGE - 1) / OZ_KRANGE;
    // stage of a pass: nS units of A (4 KB) then nS half units of B (2 KB)

    if (warp == 0) {
        if (lane == 0) {
            uint32_t filled = 0, epar = 0;
            for (int rg = 0; rg < nrange; ++rg) {
                const int k0 = rg * OZ_KRANGE, k1 = min(p.nk, k0 + OZ_KRANGE);
                for (int ps = 0; ps < npass; ++ps) {
                    const int nS = min(S, p.pd1[ps] + 1);
                    const uint32_t sbytes = (uint32_t)nS * (OZ_UNIT + BU);
                    const int nst = min(OZ_MAXST, (OZ_RING * OZ_UNIT) / (int)sbytes);
                    for (int s2 = 0; s2 < OZ_MAXST; ++s2)
                        if ((filled >> s2) & 1u) mbar_wait(empty + s2, (epar >> s2) & 1u);
                    int st = 0;
                    for (int kc = k0; kc < k1; ++kc) {
                        if ((filled >> st) & 1u) { mbar_wait(empty + st, (epar >> st) & 1u); epar ^= 1u << st; }
                        else filled |= 1u << st;
                        if (p.ablate & 1) { mbar_arrive(full + st); if (++st == nst) st = 0; continue; }
                        mbar_expect_tx(full + st, sbytes);
                        uint8_t *sa = smem + (size_t)st * sbytes;
                        bulk_g2s(sa, p.Q + ((size_t)I * p.nk + kc) * (size_t)S * OZ_UNIT, (uint32_t)nS * OZ_UNIT, full + st);
                        uint8_t *sb = sa + (size_t)nS * OZ_UNIT;
                        if (WIDE) {
                            bulk_g2s(sb, p.Q + ((size_t)Jb * p.nk + kc) * (size_t)S * OZ_UNIT, (uint32_t)nS * OZ_UNIT, full + st);
                        } else {
                            const uint8_t *qb = p.Q + ((size_t)J * p.nk + kc) * (size_t)S * OZ_UNIT + (size_t)rank * OZ_HALF;
                            for (int s = 0; s < nS; ++s) bulk_g2s(sb + (size_t)s * OZ_HALF, qb + (size_t)s * OZ_UNIT, OZ_HALF, full + st);
                        }
                        if (++st == nst) st = 0;
                    }
                }
            }
            dbg_put(p, 1, 0x200u + rank);
        }
    } else if (warp == 1) {
        uint32_t cpar = 0;
        int g = 0;
        if (!leader) {
            // ===== peer: relay "my stage is full" to the leader =====
            const uint32_t pf0 = smem_u32(pfull) & OZ_PEER_MASK;
            for (int rg = 0; rg < nrange; ++rg) {
                const int k0 = rg * OZ_KRANGE, k1 = min(p.nk, k0 + OZ_KRANGE);
                for (int ps = 0; ps < npass; ++ps) {
                    const int nS = min(S, p.pd1[ps] + 1);
                    const uint32_t sbytes = (uint32_t)nS * (OZ_UNIT + BU);
                    const int nst = min(OZ_MAXST, (OZ_RING * OZ_UNIT) / (int)sbytes);
                    int st = 0;
                    for (int kc = k0; kc < k1; ++kc) {
                        mbar_wait(full + st, (cpar >> st) & 1u);
                        cpar ^= 1u << st;
                        if (lane == 0) mbar_arrive_cluster(pf0 + (uint32_t)st * 8u);
                        __syncwarp();
                        if (++st == nst) st = 0;
                    }
                }
            }
        } else {
            // ===== leader: MMA issue for both SMs =====
            const uint32_t hi = (256u >> 4) | (1u << 14) | (6u << 29);      // SW32, K-major, 8-row groups 256 B apart
            const uint32_t sbase = smem_u32(smem);
            for (int rg = 0; rg < nrange; ++rg) {
                const int k0 = rg * OZ_KRANGE, k1 = min(p.nk, k0 + OZ_KRANGE);
                for (int ps = 0; ps < npass; ++ps, ++g) {
                    const int d0 = p.pd0[ps], d1 = p.pd1[ps];
                    const int nS = min(S, d1 + 1);
                    const uint32_t sbytes = (uint32_t)nS * (OZ_UNIT + BU);
                    const int nst = min(OZ_MAXST, (OZ_RING * OZ_UNIT) / (int)sbytes);
                    const int code = (nS == d1 + 1) ? d0 * 16 + d1 : -1;
                    if (g > 0) { mbar_wait_cluster(acc_empty, (uint32_t)(g - 1) & 1); tc_fence_after(); }
                    uint32_t touched = 0;
                    int st = 0;
                    for (int kc = k0; kc < k1; ++kc) {
                        mbar_wait(full + st, (cpar >> st) & 1u);
                        mbar_wait_cluster(pfull + st, (cpar >> st) & 1u);
                        cpar ^= 1u << st;
                        tc_fence_after();
                        const uint32_t a0 = sbase + (uint32_t)st * sbytes;
                        const uint32_t b0 = a0 + (uint32_t)nS * OZ_UNIT;
                        const bool firstk = (kc == k0);
                        if (elect_one()) {
#define OZ_CASE2(D0, D1)                                                                             \
    case (D0) * 16 + (D1):                                                                           \
        if (firstk) oz_issue_step2<D0, D1, true, WIDE>(a0 >> 4, b0 >> 4, hi, tmem_base);             \
        else oz_issue_step2<D0, D1, false, WIDE>(a0 >> 4, b0 >> 4, hi, tmem_base);                   \
        break;
                            switch ((p.ablate & 2) ? -2 : (WIDE ? code + 4096 : code)) {
                                case -2: break;
                                OZ_CASE2(0, 0) OZ_CASE2(1, 4) OZ_CASE2(5, 8)
                                OZ_CASE2(0, 1) OZ_CASE2(2, 4)
                                OZ_CASE2(0, 3) OZ_CASE2(4, 7)
                                OZ_CASE2(2, 5) OZ_CASE2(6, 8) OZ_CASE2(8, 8)
#define OZ_CASE2W(D0, D1)                                                                            \
    case 4096 + (D0) * 16 + (D1):                                                                    \
        if (firstk) oz_issue_step2<D0, D1, true, true>(a0 >> 4, b0 >> 4, hi, tmem_base);             \
        else oz_issue_step2<D0, D1, false, true>(a0 >> 4, b0 >> 4, hi, tmem_base);                   \
        break;
                                OZ_CASE2W(0, 0) OZ_CASE2W(1, 2) OZ_CASE2W(3, 4) OZ_CASE2W(5, 6) OZ_CASE2W(7, 8)
                                OZ_CASE2W(0, 1) OZ_CASE2W(2, 3) OZ_CASE2W(4, 5) OZ_CASE2W(6, 7)
#undef OZ_CASE2W
                            default:
                                for (int s = 0; s < nS; ++s) {
                                    const int tlo = max(0, d0 - s), thi = min(nS - 1, d1 - s);
                                    const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)(((a0 + s * OZ_UNIT) >> 4) & 0x3FFFu);
                                    for (int tt = tlo; tt <= thi; ++tt) {
                                        const int lev = s + tt - d0;
                                        const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)(((b0 + tt * BU) >> 4) & 0x3FFFu);
                                        tc_mma2_i8(tmem_base + lev * LW, da, db, ID, (touched >> lev) & 1u);
                                        touched |= 1u << lev;
                                    }
                                }
                            }
#undef OZ_CASE2
                            tc_commit2_mc(empty + st);       // frees the stage in BOTH CTAs once these MMAs have read it
                        }
                        __syncwarp();
                        if (++st == nst) st = 0;
                    }
                    if (elect_one()) { tc_commit2_mc(acc_full); dbg_put(p, 2, 0x300u + g); }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue (both CTAs): warp w owns TMEM lanes 32*(w%4) .. +31 of ITS OWN SM =====
        const int quad = warp & 3;
        const int row = I * OZ_T + quad * 32 + lane;
        const bool row_ok = live && row < p.n;
        const double rsc = row_ok ? p.cs[row] : 0.0;
        const uint32_t ae = smem_u32(acc_empty) & OZ_PEER_MASK;      // the leader's barrier
        int g = 0;
        for (int rg = 0; rg < nrange; ++rg) {
            for (int ps = 0; ps < npass; ++ps, ++g) {
                const int d0 = p.pd0[ps], nlev = p.pd1[ps] - d0 + 1;
                mbar_wait(acc_full, (uint32_t)g & 1);
                tc_fence_after();
                const double lsc = ldexp(1.0, -12 - 7 * (d0 + nlev - 1));
                const bool first = (g == 0);
                auto load_old = [&](int c0, double (&o)[16]) {          // see oz_mma_kernel
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int col = Jc * OZ_T + c0 + c;
                        double v = 0.0;
                        if (row_ok && col < p.n && col <= row) {
                            if (!first) v = p.C[row + (size_t)col * p.ldc];
                            else if (p.D) v = p.beta * p.D[row + (size_t)col * p.ldd];
                        }
                        o[c] = v;
                    }
                };
                double oldv[16], nxtv[16];
                load_old(0, oldv);
                for (int c0 = 0; c0 < LW; c0 += 16) {
                    uint32_t r[4][16];
                    const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + c0;
                    tc_ld16(ta, r[0]);
                    if (nlev > 1) tc_ld16(ta + LW, r[1]);
                    if (!WIDE && nlev > 2) tc_ld16(ta + 2 * LW, r[2]);
                    if (!WIDE && nlev > 3) tc_ld16(ta + 3 * LW, r[3]);
                    if (c0 + 16 < LW) load_old(c0 + 16, nxtv);
                    tc_wait_ld();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int col = Jc * OZ_T + c0 + c;
                        double v = (double)(int)r[0][c];
                        if (nlev > 1) v = fma(v, 128.0, (double)(int)r[1][c]);
                        if (!WIDE && nlev > 2) v = fma(v, 128.0, (double)(int)r[2][c]);
                        if (!WIDE && nlev > 3) v = fma(v, 128.0, (double)(int)r[3][c]);
                        if (row_ok && col < p.n && col <= row) {
                            v = (v * lsc) * rsc * p.cs[col];
                            p.C[row + (size_t)col * p.ldc] = v + oldv[c];
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 16; ++c) oldv[c] = nxtv[c];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(ae);
                if (quad == 0 && lane == 0) dbg_put(p, 3, 0x400u + g);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's shared memory / TMEM stay alive until the leader's MMAs are done
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// C tile = beta * D + sum of the nchunk partial tiles of the split-K tail launch (fixed order: deterministic)
__global__ void oz_tail_reduce_kernel(OzParams p, int ntail) {
    const int t = blockIdx.y;
    const unsigned int tl = p.tiles[t];
    const int I = (int)(tl >> 16), J = (int)(tl & 0xFFFFu);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= OZ_T * OZ_T) return;
    const int r = e % OZ_T, c = e / OZ_T;
    const int row = I * OZ_T + r, col = J * OZ_T + c;
    if (row >= p.n || col >= p.n || col > row) return;
    double v = p.D ? p.beta * p.D[row + (size_t)col * p.ldd] : 0.0;
    const double *src = p.part + (size_t)t * p.nchunk * (OZ_T * OZ_T) + e;
    for (int ch = 0; ch < p.nchunk; ++ch) v += src[(size_t)ch * (OZ_T * OZ_T)];
    p.C[row + (size_t)col * p.ldc] = v;
}

// amax[j] = max_k |d[k] * A[k,j]|  ->  cs[j] = 2^e_j, sinv[j] = 64 * 2^-e_j
__global__ void oz_colscale_kernel(int m, int n, const double *A, long long lda, const double *d, double *cs, double *sinv) {
    __shared__ double sh[32];
    const int j = blockIdx.x;
    const double *a = A + (size_t)j * lda;
    double mx = 0.0;
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        double v = fabs((d ? d[k] : 1.0) * a[k]);
        if (!(v <= 1.7976931348623157e308)) v = INFINITY;      // NaN or Inf entry: poison the column (below)
        mx = fmax(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmax(mx, sh[w]);
        int e = 0;
        if (mx > 0.0 && mx < INFINITY) frexp(mx, &e);         // mx = f * 2^e, f in [0.5, 1): |x| < 2^e
        // a non-finite entry makes row and column j of C NaN, as the fp64 kernel would (potrf then reports it)
        cs[j] = (mx < INFINITY) ? ldexp(1.0, e) : __longlong_as_double(0x7ff8000000000000LL);
        sinv[j] = (mx < INFINITY) ? ldexp(64.0, -e) : 0.0;
    }
}

// digits of column block cb, k step kc (one 128-column x 32-row block of Gs = 9 units).  The block is read with
// coalesced 256-byte rows (lane = k, one column per load instruction) into shared memory and sliced from there by one
// thread per column; reading it with one thread per column straight from global memory (32 distinct lines per load
// instruction) ran at 3 TB/s (742 us at n=8192, m=16384, profiles/r02n launch list).
__global__ void __launch_bounds__(OZ_T) oz_slice_kernel(int m, int n, const double *A, long long lda, const double *d,
                                                        const double *sinv, uint8_t *Q, int nk, int S, int layout) {
    __shared__ double tile[OZ_T][OZ_KS + 1];
    const int cb = blockIdx.y, kc = blockIdx.x, r = threadIdx.x, lane = r & 31, warp = r >> 5;
    {
        const int k = kc * OZ_KS + lane;
        const double dk = (k < m) ? (d ? d[k] : 1.0) : 0.0;
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
            const int cc = warp * 32 + c, j = cb * OZ_T + cc;
            double x = 0.0;
            if (j < n && k < m) x = dk * A[(size_t)j * lda + k];
            tile[cc][lane] = x;
        }
    }
    __syncthreads();
    const int j = cb * OZ_T + r;
    const double sc = (j < n) ? sinv[j] : 0.0;
    uint8_t *unit0 = Q + ((size_t)cb * nk + kc) * (size_t)S * OZ_UNIT;
    for (int half = 0; half < 2; ++half) {
        uint32_t pk[OZ_SMAX][4];
#pragma unroll
        for (int s = 0; s < OZ_SMAX; ++s) pk[s][0] = pk[s][1] = pk[s][2] = pk[s][3] = 0u;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            double x = tile[r][half * 16 + e] * sc;                          // |x| < 64 (0 outside the matrix)
#pragma unroll
            for (int s = 0; s < OZ_SMAX; ++s) {
                if (s < S) {
                    // round to nearest-even integer by adding 1.5 * 2^52 (|x| <= 8192 here): the sum's low mantissa word
                    // IS the integer in two's complement, so there is neither a FRND nor an F2I (both quarter rate) per
                    // digit: this kernel was bound by them (0.71 ms for 2.3 GB of traffic)
                    const double t = x + 6755399441055744.0;
                    const double q = t - 6755399441055744.0;               // == rint(x)
                    x = (x - q) * 128.0;                                   // exact
                    pk[s][e >> 2] |= ((uint32_t)__double2loint(t) & 0xFFu) << ((e & 3) * 8);
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

// level groups of <= 4 (TMEM holds four 128-column accumulators) minimising the modelled time
// sum_p max(operand bytes / L2 rate, MMA cycles); CVXB_OZ_GROUPS="2,4,3" overrides
static void oz_choose_groups(int S, int *npass, int *pd0, int *pd1) {
    if (const char *e = getenv("CVXB_OZ_GROUPS")) {
        int d = 0, np = 0;
        const char *q = e;
        while (*q && np < OZ_MAXPASS) {
            const int len = atoi(q);
            if (len < 1 || len > 4 || d + len > S) break;
            pd0[np] = d; pd1[np] = d + len - 1; d += len; ++np;
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        if (d == S) { *npass = np; return; }
    }
    if (S == 9) {        // measured on B200 (n=8192, m=16384): {0,1} {2..4} {5..8} 18.5 ms, {0} {1..4} {5..8} 18.9, {0..3} {4..7} {8} 18.9
        *npass = 3; pd0[0] = 0; pd1[0] = 1; pd0[1] = 2; pd1[1] = 4; pd0[2] = 5; pd1[2] = 8;
        return;
    }
    double best = 1e300;
    int bestcode = 0, bestn = 0;
    // compositions of S into parts 1..4, at most OZ_MAXPASS parts: code = base-5 digits
    for (int np = 1; np <= OZ_MAXPASS; ++np) {
        int lim = 1;
        for (int i = 0; i < np; ++i) lim *= 4;
        for (int code = 0; code < lim; ++code) {
            int c = code, d = 0;
            double cost = 0;
            for (int i = 0; i < np; ++i) {
                const int len = 1 + c % 4; c /= 4;
                const int d0 = d, d1 = d + len - 1;
                d += len;
                if (d > S) { cost = 1e300; break; }
                const int nS = d1 + 1;
                double mm = 0;
                for (int l = d0; l <= d1; ++l) mm += l + 1;
                cost += std::max(2.0 * nS * OZ_UNIT / 42.5, mm * 64.0) + 30.0;
            }
            if (d != S || cost >= best) continue;
            best = cost; bestcode = code; bestn = np;
        }
    }
    int c = bestcode, d = 0;
    for (int i = 0; i < bestn; ++i) { const int len = 1 + c % 4; c /= 4; pd0[i] = d; pd1[i] = d + len - 1; d += len; }
    *npass = bestn;
}

// Optional timing of the MMA launches alone (without the two slicing kernels): the NEXT ozaki_syrk call of this thread
// records `a` before its first MMA launch and `b` after its last one (bench.py's roofline line via cvxb_kkt_syrk_mma_ms).
static thread_local cudaEvent_t g_oz_ev0 = nullptr, g_oz_ev1 = nullptr;
void ozaki_time_mma(cudaEvent_t a, cudaEvent_t b) { g_oz_ev0 = a; g_oz_ev1 = b; }

size_t ozaki_workspace_bytes(int n, int m, int S) {
    const size_t nblk = (n + OZ_T - 1) / OZ_T, nk = std::max(1, (m + OZ_KS - 1) / OZ_KS);
    return nblk * nk * (size_t)S * OZ_UNIT + 2 * (size_t)n * sizeof(double) + nblk * (nblk + 1) / 2 * sizeof(unsigned int) + 1024
           + (size_t)kNumSMs * OZ_T * OZ_T * sizeof(double) + 256;         // partial tiles of the split-K tail wave
}

// C(lower) = A' diag(d)^2 A + beta * D.  A: m x n (lda), d: m (or nullptr).  work: ozaki_workspace_bytes.
int ozaki_syrk(int n, int m, const double *A, long long lda, const double *d, const double *D, long long ldd,
               double beta, double *C, long long ldc, int S, int layout, void *work, unsigned int *dbg,
               cudaStream_t st) {
    if (n <= 0) return 0;
    if (S < 1 || S > OZ_SMAX || !A || !C || !work || m < 0) { set_error("ozaki_syrk: bad arguments"); return CVXB_E_ARG; }
    static DeviceOnce once;
    if (const unsigned long long bit = once.pending()) {
        CVXB_CUDA(cudaFuncSetAttribute(oz_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        CVXB_CUDA(cudaFuncSetAttribute(oz_mma2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        CVXB_CUDA(cudaFuncSetAttribute(oz_mma2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        once.mark(bit);
    }
    const int nblk = (n + OZ_T - 1) / OZ_T, nk = std::max(1, (m + OZ_KS - 1) / OZ_KS);
    if (nblk > 65535) { set_error("ozaki_syrk: n too large"); return CVXB_E_ARG; }
    const long long tiles = (long long)nblk * (nblk + 1) / 2;
    double *cs = reinterpret_cast<double *>(work);
    double *sinv = cs + n;
    unsigned int *dtiles = reinterpret_cast<unsigned int *>(sinv + n);
    uint8_t *Q = reinterpret_cast<uint8_t *>(((uintptr_t)(dtiles + tiles) + 255) & ~uintptr_t(255));
    // launch order: bands of tile rows, column by column inside a band, so that the ~148 tiles in
    // flight form a compact block (few distinct operand streams -> L2 hits instead of HBM reads)
    // (rebuilt per call: a few microseconds, and no shared mutable state between threads / handles; the
    // pageable-source copy is staged before cudaMemcpyAsync returns)
    // CVXB_OZ_2SM=1 selects the two-SM (cta_group::2) kernel, =2 its WIDE form (256x256x32 MMAs, 256x256 blocks of C
    // per SM pair).  Both are correct and both are slower than the one-SM kernel (23.1 / 18.7 ms against 14.1 ms at
    // n=8192, m=16384): not for the MMA rate (tools/int8_peak.cu: cta_group::2 256x128x32 = 64 cycles, 256x256x32 = 128)
    // but for their copy pipeline -- every stage needs the peer's "full" relayed to the leader and a multicast commit
    // back, and the ring holds 3 such stages; with MMAs disabled they still take 20-24 ms (profiles/r02p).  Opt-in.
    int two_sm = 0;
    if (const char *e = getenv("CVXB_OZ_2SM")) two_sm = (layout == 0 && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
    std::vector<unsigned int> order;
    order.reserve((size_t)tiles);
    int band = 12;
    if (const char *e = getenv("CVXB_OZ_BAND")) band = std::max(1, atoi(e));
    static int tail_on = -1;
    if (tail_on < 0) { const char *e = getenv("CVXB_OZ_TAIL"); tail_on = (e && e[0] == '0') ? 0 : 1; }
    const int npairs_hw = kNumSMs / 2;
    long long npairs = 0, nmain2 = 0;
    if (two_sm == 2) {
        // 256 x 256 blocks (a, b), a >= b, band-major; the blocks of the last partial wave of SM pairs are handed to
        // the one-SM kernel as 128 x 128 tiles split along K (below)
        const int na = (nblk + 1) / 2, pband = std::max(1, band / 2);
        for (int a0 = 0; a0 < na; a0 += pband) {
            const int a1 = std::min(na, a0 + pband);
            for (int b = 0; b < a1; ++b)
                for (int a = std::max(a0, b); a < a1; ++a) order.push_back(((unsigned)a << 16) | (unsigned)b);
        }
        npairs = (long long)order.size();
        nmain2 = npairs;
        const long long rem = npairs % npairs_hw;
        if (tail_on && npairs > npairs_hw && rem > 0 && 4 * rem <= kNumSMs / 2) nmain2 = npairs - rem;
        std::vector<unsigned int> t128;
        for (long long q = nmain2; q < npairs; ++q) {
            const int a = (int)(order[(size_t)q] >> 16), b = (int)(order[(size_t)q] & 0xFFFFu);
            for (int di = 0; di < 2; ++di)
                for (int dj = 0; dj < 2; ++dj) {
                    const int I = 2 * a + di, J = 2 * b + dj;
                    if (I < nblk && J < nblk && I >= J) t128.push_back(((unsigned)I << 16) | (unsigned)J);
                }
        }
        order.resize((size_t)nmain2);
        order.insert(order.end(), t128.begin(), t128.end());
    } else if (two_sm == 1) {
        // pairs of row blocks (2a, 2a+1) x column block J with 2a+1 >= J (the tile (2a, J) of a pair that
        // straddles the diagonal is computed and masked), same band-major launch order
        const int na = (nblk + 1) / 2, pband = std::max(1, band / 2);
        for (int a0 = 0; a0 < na; a0 += pband) {
            const int a1 = std::min(na, a0 + pband);
            for (int J = 0; J < nblk && J <= 2 * a1 - 1; ++J)
                for (int a = std::max(a0, J / 2); a < a1; ++a)
                    if (2 * a + 1 >= J) order.push_back(((unsigned)a << 16) | (unsigned)J);
        }
        npairs = nmain2 = (long long)order.size();
    } else {
        for (int r0 = 0; r0 < nblk; r0 += band) {
            const int r1 = std::min(nblk, r0 + band);
            for (int J = 0; J < r1; ++J)
                for (int I = std::max(r0, J); I < r1; ++I) order.push_back(((unsigned)I << 16) | (unsigned)J);
        }
    }
    if ((long long)order.size() > tiles) { set_error("ozaki_syrk: tile list overflow"); return CVXB_E_ARG; }
    CVXB_CUDA(cudaMemcpyAsync(dtiles, order.data(), order.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, st));
    oz_colscale_kernel<<<n, 256, 0, st>>>(m, n, A, lda, d, cs, sinv);
    count_launch();
    oz_slice_kernel<<<dim3(nk, nblk), OZ_T, 0, st>>>(m, n, A, lda, d, sinv, Q, nk, S, layout);
    count_launch();
    OzParams p;
    p.Q = Q; p.cs = cs; p.D = D; p.ldd = ldd; p.C = C; p.ldc = ldc; p.beta = beta;
    p.n = n; p.nblk = nblk; p.nk = nk; p.S = S; p.layout = layout; p.dbg = dbg; p.tiles = dtiles;
    p.nchunk = 1; p.kper = nk; p.part = nullptr;
    p.ablate = 0;
    p.trace_cta = -1;
    if (const char *e = getenv("CVXB_OZ_TRACE_CTA")) p.trace_cta = atoi(e);
    p.kb2 = 4; p.kb5 = 1;
    if (const char *e = getenv("CVXB_OZ_KB")) {
        int a = 0, b = 0;
        if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 1 && a <= 6 && b >= 1 && b <= 2) { p.kb2 = a; p.kb5 = b; }
    }
    if (const char *e = getenv("CVXB_OZ_ABLATE")) p.ablate = atoi(e);
    oz_choose_groups(S, &p.npass, p.pd0, p.pd1);
    cudaEvent_t tev0 = g_oz_ev0, tev1 = g_oz_ev1;
    g_oz_ev0 = g_oz_ev1 = nullptr;
    if (tev0) CVXB_CUDA(cudaEventRecord(tev0, st));
    long long t1_first = 0, t1_count = tiles;            // 128 x 128 tiles left to the one-SM kernel
    if (two_sm) {
        OzParams p2 = p;
        if (two_sm == 2) {                                // two levels per pass: {0} {1,2} {3,4} ... (S odd) / {0,1} {2,3} ...
            int np = 0, dlev = 0;
            if (S & 1) { p2.pd0[np] = 0; p2.pd1[np] = 0; ++np; dlev = 1; }
            for (; dlev < S; dlev += 2, ++np) { p2.pd0[np] = dlev; p2.pd1[np] = dlev + 1; }
            p2.npass = np;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * nmain2), 1, 1);
        cfg.blockDim = dim3(OZ_THREADS, 1, 1);
        cfg.dynamicSmemBytes = OZ_SMEM;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (nmain2 > 0) {
            if (two_sm == 2) CVXB_CUDA(cudaLaunchKernelEx(&cfg, oz_mma2_kernel<true>, p2));
            else CVXB_CUDA(cudaLaunchKernelEx(&cfg, oz_mma2_kernel<false>, p2));
        }
        t1_first = nmain2;
        t1_count = (long long)order.size() - nmain2;
    }
    if (t1_count > 0) {
        // The last, partial wave of tiles (2080 = 14 x 148 + 8 at n = 8192) would leave most SMs idle for a whole
        // tile time: its tiles are split along K over all SMs (partials + ordered reduce).  CVXB_OZ_TAIL=0 disables.
        long long tmain = t1_count;
        int ntail = (int)(t1_count % kNumSMs), nchunk = 1;
        if (tail_on && (two_sm || t1_count > kNumSMs) && ntail > 0 && ntail <= kNumSMs / 2) {
            nchunk = std::min(kNumSMs / ntail, nk);
            if (nchunk >= 2) tmain = t1_count - ntail; else nchunk = 1;
        }
        OzParams pm = p;
        pm.tiles = dtiles + t1_first;
        if (tmain > 0) { oz_mma_kernel<<<(unsigned)tmain, OZ_THREADS1, OZ_SMEM, st>>>(pm); if (two_sm) count_launch(); }
        if (nchunk >= 2) {
            OzParams pt = p;
            pt.tiles = dtiles + t1_first + tmain;
            pt.nchunk = nchunk;
            pt.kper = (nk + nchunk - 1) / nchunk;
            pt.nchunk = (nk + pt.kper - 1) / pt.kper;                 // no empty chunk
            pt.part = reinterpret_cast<double *>(((uintptr_t)(Q + (size_t)nblk * nk * S * OZ_UNIT) + 255) & ~uintptr_t(255));
            oz_mma_kernel<<<(unsigned)(ntail * pt.nchunk), OZ_THREADS1, OZ_SMEM, st>>>(pt);
            oz_tail_reduce_kernel<<<dim3(OZ_T * OZ_T / 256, ntail), 256, 0, st>>>(pt, ntail);
            count_launch(2);
        }
    }
    if (tev1) CVXB_CUDA(cudaEventRecord(tev1, st));
    count_launch();
    CVXB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cvxb
