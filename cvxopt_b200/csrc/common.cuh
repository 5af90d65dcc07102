// Shared device/host helpers for the cvxopt_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <atomic>
#include "../../include/cvxopt_b200.h"

namespace cvxb {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;   // kernels launched by this library
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// Function attributes (dynamic shared-memory opt-in ...) are PER DEVICE: a call site keeps one of
// these and sets its attributes the first time it runs on each device of the process.
struct DeviceOnce {
    std::atomic<unsigned long long> done{0};
    // returns the bit of the current device if its attributes are still to be set, else 0
    unsigned long long pending() {
        int d = 0;
        cudaGetDevice(&d);
        const unsigned long long bit = 1ull << (d & 63);
        return (done.load(std::memory_order_acquire) & bit) ? 0ull : bit;
    }
    void mark(unsigned long long bit) { done.fetch_or(bit, std::memory_order_release); }
};

#define CVXB_CUDA(expr)                                                            \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            cvxb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,          \
                            cudaGetErrorString(_e));                               \
            return CVXB_E_CUDA;                                                    \
        }                                                                          \
    } while (0)

#define CVXB_TRY(expr)                                                             \
    do {                                                                           \
        int _r = (expr);                                                           \
        if (_r != 0) return _r;                                                    \
    } while (0)

#define CVXB_LAUNCH_CHECK()                                                        \
    do {                                                                           \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) {                                                   \
            cvxb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,      \
                            cudaGetErrorString(_e));                               \
            return CVXB_E_CUDA;                                                    \
        }                                                                          \
    } while (0)

// ---- scratch-buffer cache ----------------------------------------------------
// The cone / scaling entry points (cvxb_scale, cvxb_max_step, cvxb_update_scaling ...) need device temporaries per call;
// cudaMalloc + cudaFree cost 0.1-1 ms each (cudaFree also synchronises the device), which was most of a call
// (profiles/r02t: 10.6 ms per cvxb_scale on a 512 x 512 's' block whose GEMMs take 0.1 ms).  tmp_malloc / tmp_free keep
// freed blocks per device and hand them out again (best fit within +25 %); callers synchronise their stream before
// freeing, as they did for cudaFree.  The cache is bounded (CVXB_TMP_CACHE_MB, default 4096) and is dropped when a
// real allocation fails.
cudaError_t tmp_malloc_bytes(void **p, size_t bytes);
void tmp_free(void *p);
void tmp_cache_release();          // free every cached block of every device
template <class T> inline cudaError_t tmp_malloc(T **p, size_t bytes) { return tmp_malloc_bytes(reinterpret_cast<void **>(p), bytes); }

constexpr int kNumSMs = 148;       // B200: 2 dies x 74 SMs
constexpr int NB = 128;            // Cholesky block size == GEMM tile edge

// ---- device helpers ---------------------------------------------------------
#ifdef __CUDACC__
// fp64 tensor-core MMA: D(8x8) += A(8x4,row) * B(4x8,col).  SASS: DMMA.8x8x4.
// lane -> A[lane>>2][lane&3], B[k=lane&3][n=lane>>2], D[lane>>2][2*(lane&3)+{0,1}]
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// 16-byte async copy global->shared, zero-filling (16 - src_bytes) trailing bytes.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem)), "l"(gmem),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#endif

// ---- host-side kernels' launch API (device pointers, explicit stream) --------
struct GemmDesc {
    // C[r,c] = alpha * sum_k X[r,k] * w[k] * Y[c,k] + beta * D[r,c]      (col-major C/D)
    // X[r,k] = x_kmajor ? X[k + r*ldx] : X[r + k*ldx]; same for Y.
    int M = 0, N = 0, K = 0;
    const double *X = nullptr; int ldx = 0; bool x_kmajor = false;
    const double *Y = nullptr; int ldy = 0; bool y_kmajor = false;
    const double *w = nullptr;          // optional K-vector applied inside the contraction
    const double *D = nullptr; int ldd = 0;
    double *C = nullptr; int ldc = 0;
    double alpha = 1.0, beta = 0.0;
    bool lower_only = false;            // only tiles with r-tile >= c-tile (M == N)
    // tile-column window (look-ahead in the Cholesky): only c-tiles in [ct_begin, ct_end)
    int ct_begin = 0, ct_end = 1 << 30;
    // batching (blockIdx.z): element strides between consecutive problems
    int batch = 1;
    long long sX = 0, sY = 0, sW = 0, sD = 0, sC = 0;
    // split-K remainder workspace (>= kNumSMs * 128*128 doubles) or nullptr to disable
    double *splitk_ws = nullptr;
    // debug timeline (CVXB_TRACE): [0] = min CTA start, [1] = max CTA end (globaltimer ns)
    unsigned long long *trace = nullptr;
};
int dmma_gemm(const GemmDesc &g, cudaStream_t st);
size_t dmma_gemm_splitk_ws_doubles();   // workspace size for GemmDesc::splitk_ws
int dmma_gemm_tile_cols();              // width of a c tile (units of ct_begin / ct_end)

// ozaki_syrk.cu (experimental): C(lower) = A' diag(d)^2 A + beta*D through int8 slices on tcgen05
size_t ozaki_workspace_bytes(int n, int m, int S);
void ozaki_time_mma(cudaEvent_t a, cudaEvent_t b);   // events recorded around the MMA launches of this thread's next call
int ozaki_syrk(int n, int m, const double *A, long long lda, const double *d, const double *D,
               long long ldd, double beta, double *C, long long ldc, int S, int layout, void *work,
               unsigned int *dbg, cudaStream_t st);

// Cholesky (lower) of the n x n matrix A in place; inv receives the inverses of the
// NB x NB diagonal blocks of L (block j at inv + j*NB*NB, leading dimension NB).
// info (device int) = first non-positive pivot (1-based) or 0.
struct CholWork {
    cudaStream_t panel_stream = nullptr;   // high-priority: diagonal block, panel, next-panel update
    cudaStream_t trsm_stream = nullptr;    // high-priority: panel TRSM + next block column update
    cudaStream_t update_stream = nullptr;  // low-priority: bulk trailing update
    cudaEvent_t ev_end_t = nullptr;
    std::vector<cudaEvent_t> ev_dg, ev_tr, ev_c0, ev_r;   // one per block step
    unsigned long long *trace = nullptr;  // CVXB_TRACE=1: per step {Dg, Tr, C0, R} x {start, end}
    struct GraphEntry {                    // captured factorisation, keyed by its arguments
        int n = 0, lda = 0, launches = 0;
        const void *A = nullptr, *inv = nullptr;
        cudaGraphExec_t exec = nullptr;
    };
    std::vector<GraphEntry> graphs;
    bool graph_failed = false;
    int r_valid[2] = {-1, -1};            // which steps recorded ev_r (steps without bulk work do not)
    cudaEvent_t ev_start = nullptr, ev_panel = nullptr, ev_rest = nullptr, ev_end_p = nullptr,
                ev_end_u = nullptr;
    int *d_info = nullptr;                // device flag
    int *d_flags = nullptr;               // trsv progress flags (batch * ceil(n/NB) ints)
    long long flags_cap = 0;
    double *splitk_ws = nullptr;
    double *panel[2] = {nullptr, nullptr};   // out-of-place TRSM results (double-buffered)
    int panel_rows = 0;
};
int chol_work_create(CholWork &w);
void chol_work_destroy(CholWork &w);
int potrf_lower(int n, double *A, int lda, double *inv, CholWork &w, cudaStream_t st);
// b := L^{-T} L^{-1} b  (potrs with one right-hand side)
// batched variants: problem p uses L + p*sL, inv + p*sInv, b + p*sb
int potrs_lower(int n, const double *L, int ldl, const double *inv, double *b, CholWork &w,
                cudaStream_t st, int batch = 1, long long sL = 0, long long sInv = 0,
                long long sb = 0);
int trsv_lower(int n, const double *L, int ldl, const double *inv, double *b, bool trans,
               CholWork &w, cudaStream_t st, int batch = 1, long long sL = 0, long long sInv = 0,
               long long sb = 0);
int potrf_lower_batched(int n, double *A, int lda, long long sA, double *inv, long long sInv,
                        int batch, int *d_info, double *panel, int ldw, cudaStream_t st);

}  // namespace cvxb
