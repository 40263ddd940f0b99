// Host side of libexauq_b200.so: device state behind opaque handles, the recursive
// potrf+trtri driver, chunked prediction / PEI pipeline and the C ABI of
// include/exauq_b200.h.  sm_100a only; no CPU fallback.
#include "../../include/exauq_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "exq_common.cuh"
#include "gemm_dmma.cuh"
#include "kernels.cuh"
#include "leaf.cuh"
#include "fused_node.cuh"
#include "ozaki_i8.cuh"
#include "ozaki_gemm.cuh"
#include "pei_fused.cuh"
#include "pivot.cuh"

using namespace exq;

// -------------------------------------------------------------------------------------
// global context
// -------------------------------------------------------------------------------------
namespace {

struct ProfClass {
    std::vector<std::pair<int, int>> pairs;   // indices into the event pool
    int64_t launches = 0;
    double flops = 0.0, bytes = 0.0;
};

struct Ctx {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    double bm64_factor = 1e9;   // 64-row tiles (2 CTAs/SM) measured faster than 128-row tiles at every level of the
                                // recursion (34.8 vs 36.6 ms per batched evaluation); EXQ_BM64_FACTOR=0 restores 128
    int variance_bm = 64;       // row tile of the variance GEMM (EXQ_VARIANCE_BM)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int64_t launches = 0;
    unsigned prof_mask = 0;
    std::vector<cudaEvent_t> pool;
    size_t pool_used = 0;
    ProfClass prof[EXQ_PROF_NCLASS];
    bool attrs_set = false;
    // prediction scratch shared by all GP handles (one stream => no overlap): covariance chunk
    // KcT[chunk][Np] and the partial column sums P[Np/128][chunk]
    double *KcT = nullptr, *P = nullptr;
    size_t kct_elems = 0, p_elems = 0;
    double* PMR = nullptr;           // [2][Np/64][chunk] partial means / repulsion products of the fused PEI pipeline
    size_t pmr_elems = 0;
    int oz_min_tiles = 16;           // fewest 128 x 64 output tiles (over the batch) for the int8 kernel INSIDE exq_gp_logpost_batch
                                     // (EXQ_OZAKI_MIN_TILES; -1: the SM count).  Even 32 CTAs finish a k = 512 product of one system faster
                                     // than the FP64 DMMA kernel does (9 us + slicing against 46 us): a round of ONE system 4.72 -> 4.36 ms.
                                     // A fit (one system, feeds predictions and the LOO sweep) keeps the SM-count rule: with all three
                                     // levels on 7-digit products its LOO means at cond 4e7 sit 8 x further from the oracle than FP64's.
    bool in_logpost = false;         // set while exq_gp_logpost_batch enqueues work
    bool oz_shallow = false;         // set while a system whose kappa reached guard_widen is re-factored with the int8 kernel on
                                     // the top levels only (k >= 1024, chip-filling batches): deeper 7-digit levels cost accuracy as
                                     // conditioning worsens (kappa 1e8: log-posterior 5e-6 from the oracle against 2e-8)
    int oz_colmax = 1;               // EXQ_OZ_COLMAX=0: separate exponent pass over T instead of column maxima from the epilogue of T's product
    int oz_sched = 1;                // EXQ_OZ_SCHED=0: arithmetic tile walk of the int8 GEMM instead of the balanced work lists
    int grad_fused = 1;              // EXQ_GRAD_FUSED=0: K^-1 stored by LAUUM and contracted by a separate gradient kernel
    int pei_fused = 1;               // EXQ_PEI_FUSED=0: round-1 pipeline (FP64 chunk -> slicer -> per-candidate epilogue)
    // int8 tcgen05 variance path (ozaki_i8.cuh): digits per operand (EXQ_OZAKI_SLICES = 0 keeps the FP64 DMMA
    // GEMM, 5..7 selects the number of base-256 digits; 6 = 48 bits is ~2e-12 absolute on the variance)
    int oz_slices = 6;
    int8_t* cand_planes = nullptr;   // [S][chunk][Np] digits of the covariance chunk
    size_t cand_planes_bytes = 0;
    // digit planes of Linv are per GP handle; a design batch creates and destroys a GP per step, so freed
    // buffers are parked here (cudaMalloc / cudaFree of 100 MB synchronise the device)
    std::vector<std::pair<size_t, int8_t*>> plane_pool;
    // int8 path for the large GEMMs of the potrf+trtri recursion and LAUUM (ozaki_gemm.cuh): 7 digits = 56 bits
    // keeps the factorisation at the FP64 level (EXQ_OZAKI_FACTOR_SLICES = 0: FP64 DMMA everywhere)
    int oz_factor_slices = 7;
    int oz_min_k = 512;              // smallest inner dimension that goes to the int8 kernel (EXQ_OZAKI_MIN_K); with the balanced
                                     // work lists the 1024-row nodes (k = 512) are faster there than on FP64 DMMA (47 -> 5.6 + 16 + 13 ms
                                     // of DMMA / int8 / slicing time per design batch)
    int fast_corr = 1;               // branch-free exp / sqrt + compile-time kernel id in the assembly kernel (EXQ_FAST_CORR)
    int oz_lauum_slices = 6;         // K^-1 = Linv^T Linv for the gradient (parity 1e-7): 48 bits unless the guard widens
    // ---- conditioning guard of the int8 paths ----------------------------------------------------------------
    // kappa = (max L_ii / min L_ii)^2 is a lower bound of cond_2(K), read back with every factorisation.  The int8
    // products round every operand row to 8 S bits relative to the row maximum, so their error relative to an
    // FP64 product grows with the dynamic range inside the rows of L / Linv, i.e. with kappa:
    //   kappa >= guard_widen    : 6-digit products (variance, LAUUM, kinv) are run with 7 digits (56 bits)
    //   kappa >= guard_fallback : the factorisation itself is redone on the FP64 DMMA path and everything
    //                             downstream (variance, LAUUM) stays on FP64 DMMA for that system
    double guard_widen = 65536.0;          // 2^16   (EXQ_GUARD_WIDEN)
    double guard_fallback = 1.0e12;        //        (EXQ_GUARD_FALLBACK)
    bool oz_off = false;                   // set while a guarded system is re-evaluated on FP64 DMMA
    // fused sub-tree kernel (fused_node.cuh): diagonal blocks of up to fuse_n rows are factored + inverted by ONE launch
    // (a cluster of fuse_g CTAs per system); 0 switches it off (EXQ_FUSE_N / EXQ_FUSE_G; fuse_g = 0: 16 CTAs for up to 8
    // systems, 8 otherwise)
    int fuse_n = 512;
    int fuse_g = 0;
    bool fuse_np16 = false;                // non-portable cluster size 16 accepted by the driver
    int fuse_clusters16 = 0, fuse_clusters8 = 0;   // co-resident clusters of 16 / 8 CTAs (cudaOccupancyMaxActiveClusters)
    int refine_steps = 1;                  // iterative-refinement steps of alpha after an int8 factorisation (EXQ_REFINE_STEPS)
    int64_t guard_widened = 0, guard_fallbacks = 0;
} g;

// Every extern "C" entry point takes this lock first: the library keeps its state (error string, scratch
// buffers, workspace pools, the one stream) in globals, ctypes releases the GIL during calls and Python may
// run a handle's destructor on any thread.  Recursive because entry points call each other.
std::recursive_mutex g_api_mutex;
#define API_LOCK() std::lock_guard<std::recursive_mutex> _api_lock(g_api_mutex)

thread_local std::string g_err;   // per calling thread: exq_last_error() reports the caller's own failure
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail(EXQ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    } while (0)

#define REQUIRE_READY()                                                           \
    API_LOCK();                                                                   \
    if (!g.ready) return fail(EXQ_ERR_CUDA, "exq_init has not succeeded (no GPU?)"); \
    cudaSetDevice(g.device) /* the calling host thread may be new (per-thread current device) */

inline int pad128(int64_t n) { return (int)(((n + 127) / 128) * 128); }

struct ProfMark {
    int cls;
    int e0 = -1;
};

ProfMark prof_begin(int cls, double flops, double bytes) {
    ProfMark m{cls};
    g.launches++;
    g.prof[cls].launches++;
    g.prof[cls].flops += flops;
    g.prof[cls].bytes += bytes;
    if (((g.prof_mask >> cls) & 1u) && g.pool_used + 2 <= g.pool.size()) {
        m.e0 = (int)g.pool_used;
        g.pool_used += 2;
        cudaEventRecord(g.pool[m.e0], g.stream);
    }
    return m;
}
void prof_end(const ProfMark& m) {
    if (m.e0 >= 0) {
        cudaEventRecord(g.pool[m.e0 + 1], g.stream);
        g.prof[m.cls].pairs.emplace_back(m.e0, m.e0 + 1);
    }
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(EXQ_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return EXQ_OK;
}
#define LAUNCH_CHECK(what)                     \
    do {                                       \
        int _rc = check_launch(what);          \
        if (_rc != EXQ_OK) return _rc;         \
    } while (0)

int set_kernel_attrs() {
    if (g.attrs_set) return EXQ_OK;
#define EXQ_SET_GEMM_ATTR(AL, BL, BM)                                                                     \
    CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_kernel<AL, BL, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  GemmCfg<BM>::SMEM_BYTES))
    EXQ_SET_GEMM_ATTR(LAYOUT_K, LAYOUT_K, 128);
    EXQ_SET_GEMM_ATTR(LAYOUT_K, LAYOUT_K, 64);
    EXQ_SET_GEMM_ATTR(LAYOUT_K, LAYOUT_MN, 128);
    EXQ_SET_GEMM_ATTR(LAYOUT_K, LAYOUT_MN, 64);
    EXQ_SET_GEMM_ATTR(LAYOUT_MN, LAYOUT_MN, 128);
    EXQ_SET_GEMM_ATTR(LAYOUT_MN, LAYOUT_MN, 64);
#undef EXQ_SET_GEMM_ATTR
    CUDA_TRY(cudaFuncSetAttribute(leaf_potrf_trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  LEAF_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(fused_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM_BYTES));
    g.fuse_np16 = cudaFuncSetAttribute(fused_factor_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    cudaGetLastError();
    {   // can the device co-schedule the clusters at all?  (16 CTAs need 16 free SMs inside one GPC)
        auto clusters = [&](int G) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G, 1, 1);
            cfg.blockDim = dim3(FUSED_THREADS, 1, 1);
            cfg.dynamicSmemBytes = FUSED_SMEM_BYTES;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, fused_factor_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
            return n;
        };
        g.fuse_clusters16 = g.fuse_np16 ? clusters(16) : 0;
        g.fuse_clusters8 = clusters(8);
        if (g.fuse_clusters16 < 1) g.fuse_np16 = false;
        if (g.fuse_clusters8 < 1) g.fuse_n = 0;   // no clusters on this device: one launch per step, as before
    }
    CUDA_TRY(oz::set_attrs<5>());
    CUDA_TRY(oz::set_attrs<6>());
    CUDA_TRY(oz::set_attrs<7>());
    CUDA_TRY(oz::set_gemm_attrs<6>());
    CUDA_TRY(oz::set_gemm_attrs<7>());
    const int big = 220 * 1024;
    CUDA_TRY(cudaFuncSetAttribute(assemble_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(assemble_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(assemble_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(assemble_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(grad_fast_kernel<EXQ_KERNEL_MATERN52, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(grad_fast_kernel<EXQ_KERNEL_MATERN52, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(grad_fast_kernel<EXQ_KERNEL_SQEXP, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(grad_fast_kernel<EXQ_KERNEL_SQEXP, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(grad_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(grad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(grad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_TRY(cudaFuncSetAttribute(grad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    g.attrs_set = true;
    return EXQ_OK;
}

// -------------------------------------------------------------------------------------
// workspace for B same-sized systems
// -------------------------------------------------------------------------------------
struct Sys {
    int B = 0, N = 0, Np = 0, d = 0;
    double *A = nullptr, *Linv = nullptr, *W = nullptr;     // [B][Np][Np]
    double* theta = nullptr;                                // [B][d+2]
    double *tvec = nullptr, *alpha = nullptr, *dinv = nullptr, *kvec = nullptr;   // [B][Np]
    double* tpart = nullptr;                                // [B][TRMV_T_SPLITS][2][Np] partial sums
    double* ldq = nullptr;                                  // [B][2] {log det K, y^T K^-1 y}
    double* dstat = nullptr;                                // [B][2] {min, max} of diag(L): conditioning proxy for the int8 guard
    int* info = nullptr;                                    // [B]
    double *Xb = nullptr, *yb = nullptr;                    // own data (LOO refits) or null
    int* skip = nullptr;
    double* gpartial = nullptr;                             // gradient partials
    double* grad = nullptr;                                 // [B][d+2]
    int8_t* oz_planes = nullptr;                            // digit planes of the int8 GEMM operands (lazy)
    int* oz_exp = nullptr;                                  // [2][B][Np] operand row exponents
    uint32_t* oz_colmax = nullptr;                          // [B][Np] column maxima (high words) of the T products, by global column
    size_t oz_bytes = 0, oz_per_sys = 0;                    // total / per-system bytes of oz_planes
    int64_t sq() const { return (int64_t)Np * Np; }
};

void sys_free(Sys& s) {
    cudaFree(s.A); cudaFree(s.Linv); cudaFree(s.W); cudaFree(s.theta); cudaFree(s.tvec);
    cudaFree(s.alpha); cudaFree(s.dinv); cudaFree(s.kvec); cudaFree(s.tpart); cudaFree(s.ldq); cudaFree(s.dstat); cudaFree(s.info);
    cudaFree(s.Xb); cudaFree(s.yb); cudaFree(s.skip); cudaFree(s.gpartial); cudaFree(s.grad);
    cudaFree(s.oz_planes); cudaFree(s.oz_exp); cudaFree(s.oz_colmax);
    s = Sys();
}

// Freed workspaces are parked here and handed back to the next request of the same shape: a
// design batch creates and destroys several GP handles per step, and cudaMalloc / cudaFree of
// multi-GB buffers would otherwise cost more than the LOO sweep itself.
std::vector<Sys> g_sys_pool;
const size_t SYS_POOL_LIMIT_BYTES = (size_t)48 << 30;

size_t sys_bytes(const Sys& s) {
    return (size_t)s.B * s.sq() * sizeof(double) * (s.W ? 3 : 2) + s.oz_bytes;
}

// a workspace is complete when every buffer its shape calls for exists (a failed cudaMalloc leaves later ones null)
bool sys_complete(const Sys& s) {
    return s.A && s.Linv && s.theta && s.tvec && s.alpha && s.dinv && s.kvec && s.tpart && s.ldq && s.info && s.grad &&
           s.dstat && (!s.W || s.gpartial) && (s.oz_bytes == 0 || (s.oz_planes && s.oz_exp && s.oz_colmax)) && (!s.Xb || (s.yb && s.skip));
}

void sys_release(Sys& s) {
    if (!s.A) { sys_free(s); return; }
    if (!sys_complete(s)) { sys_free(s); return; }   // never park a half-allocated workspace
    size_t held = 0;
    for (auto& q : g_sys_pool) held += sys_bytes(q);
    if (held + sys_bytes(s) <= SYS_POOL_LIMIT_BYTES && g_sys_pool.size() < 16) {
        g_sys_pool.push_back(s);
        s = Sys();
    } else {
        sys_free(s);
    }
}

void sys_pool_clear() {
    for (auto& q : g_sys_pool) sys_free(q);
    g_sys_pool.clear();
}

void plane_pool_clear() {
    for (auto& pp : g.plane_pool) cudaFree(pp.second);
    g.plane_pool.clear();
}

cudaError_t sys_alloc_buffers(Sys& s, int B, int N, int d, bool need_w, bool own_data) {
    s.B = B; s.N = N; s.Np = pad128(N); s.d = d;
    cudaError_t e = cudaSuccess;
    auto A = [&](void* pp, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc((void**)pp, bytes); };
    const size_t mat = (size_t)B * s.sq() * sizeof(double);
    const size_t vec = (size_t)B * s.Np * sizeof(double);
    A(&s.A, mat);
    A(&s.Linv, mat);
    if (need_w) A(&s.W, mat);
    A(&s.theta, (size_t)B * (d + 2) * sizeof(double));
    A(&s.tvec, vec);
    A(&s.alpha, vec);
    A(&s.dinv, vec);
    A(&s.kvec, vec);
    A(&s.tpart, vec * TRMV_T_SPLITS * 2);
    A(&s.ldq, (size_t)B * 2 * sizeof(double));
    A(&s.dstat, (size_t)B * 2 * sizeof(double));
    A(&s.info, (size_t)B * sizeof(int));
    A(&s.grad, (size_t)B * (d + 2) * sizeof(double));
    if (need_w) {
        const int nt = s.Np / 128;
        A(&s.gpartial, (size_t)B * nt * nt * (64 + 2) * sizeof(double));
    }
    if (g.oz_factor_slices > 0 && s.Np >= 2 * g.oz_min_k && s.Np <= 16384) {
        // digit planes + exponents of the int8 GEMM operands (LAUUM needs S planes of a whole matrix)
        s.oz_per_sys = (size_t)g.oz_factor_slices * s.sq();
        s.oz_bytes = s.oz_per_sys * B;
        A(&s.oz_planes, s.oz_bytes);
        A(&s.oz_exp, (size_t)2 * B * s.Np * sizeof(int));
        A(&s.oz_colmax, (size_t)B * s.Np * sizeof(uint32_t));
    }
    if (own_data) {
        A(&s.Xb, (size_t)B * s.Np * d * sizeof(double));
        A(&s.yb, vec);
        A(&s.skip, (size_t)B * sizeof(int));
    }
    return e;
}

int sys_alloc(Sys& s, int B, int N, int d, bool need_w, bool own_data) {
    if (s.B >= B && s.N == N && s.d == d && (!need_w || s.W) && (!own_data || s.Xb) && sys_complete(s)) return EXQ_OK;
    sys_release(s);
    for (size_t i = 0; i < g_sys_pool.size(); i++) {
        Sys& q = g_sys_pool[i];
        if (q.B >= B && q.B <= 2 * B && q.Np == pad128(N) && q.d == d && (!need_w || q.W) && (!own_data || q.Xb)) {
            s = q;
            s.N = N;
            g_sys_pool.erase(g_sys_pool.begin() + i);
            return EXQ_OK;
        }
    }
    cudaError_t e = sys_alloc_buffers(s, B, N, d, need_w, own_data);
    if (e != cudaSuccess) {
        // out of memory: nothing partial survives; give back everything parked and try once more
        sys_free(s);
        cudaGetLastError();
        sys_pool_clear();
        plane_pool_clear();
        e = sys_alloc_buffers(s, B, N, d, need_w, own_data);
        if (e != cudaSuccess) {
            sys_free(s);
            cudaGetLastError();
            return fail(EXQ_ERR_CUDA, std::string("workspace allocation (B=") + std::to_string(B) + ", N=" + std::to_string(N) +
                                          "): " + cudaGetErrorString(e));
        }
    }
    return EXQ_OK;
}

}  // namespace

struct exq_gp_s {
    int64_t N = 0;
    int d = 0, kernel_id = 0, Np = 0;
    double *X = nullptr, *y = nullptr;     // [Np][d] zero padded, [Np]
    Sys fit;                                // B = 1: fitted state
    bool fitted = false;
    double cov = 0.0, nugget = 0.0;
    double kappa = 1.0;                     // (max L_ii / min L_ii)^2 of the fit: conditioning proxy of the int8 guard
    bool fit_on_dmma = false;               // the guard sent this fit to the FP64 DMMA path
    std::vector<double> corr_raw;
    Sys batch;                              // MLE restarts / LOO refits
    // prediction scratch
    double *KcT = nullptr, *P = nullptr;   // views of the global scratch
    double* zsmall = nullptr;              // [8][Np] for the small-M variance path
    double* small_x = nullptr;             // [128][d] + [2][128]: resident scratch for predict / cross_cov calls with a
    double* small_out = nullptr;           //   handful of points (the unchanged designers call them one point at a time)
    int8_t* linv_planes = nullptr;         // [S][Np][Np] base-256 digits of Linv (int8 variance path), built lazily
    size_t linv_planes_bytes = 0;
    int* linv_exp = nullptr;               // [Np] row exponents of Linv
    int planes_S = 0;                      // digits held in linv_planes (0: stale)
    int64_t chunk = 0;
    double* rep = nullptr;
    int rep_cap = 0;
};

struct exq_cand_s {
    int64_t M = 0, Mpad = 0;
    int d = 0;
    double *Xc = nullptr, *mean = nullptr, *var = nullptr, *pei = nullptr;
    ArgPair *partial = nullptr, *best = nullptr, *chosen = nullptr;
    double* xnew = nullptr;
    int nblocks = 0, chosen_cap = 0;
    bool evaluated = false;
};

namespace {

// -------------------------------------------------------------------------------------
// kernel launch helpers
// -------------------------------------------------------------------------------------
template <int AL, int BL, int BM>
int launch_gemm_bm(const GemmArgs& a, int batch, int cls, double fl) {
    using Cfg = GemmCfg<BM>;
    dim3 grid(a.N / GEMM_BN, a.M / BM, batch);
    if (a.raster > 0) grid = dim3((a.M / BM) * (a.N / GEMM_BN), 1, batch);
    ProfMark m = prof_begin(cls, fl, 0.0);
    gemm_dmma_kernel<AL, BL, BM><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, g.stream>>>(a);
    prof_end(m);
    return check_launch("gemm_dmma_kernel");
}

// Row tile 128 (1 CTA/SM) when there are enough tiles to fill the chip, else 64 (twice the CTAs,
// 2 per SM) so that the small GEMMs low in the recursion spread over more SMs.
template <int AL, int BL>
int launch_gemm(const GemmArgs& a, int batch, int cls) {
    // algorithmic flops (triangular k ranges / lower-only halve the work)
    double fl = 2.0 * a.M * (double)a.N * a.K * batch;
    if (a.krange != KR_FULL) fl *= 0.5;                          // triangular operand
    if (a.lower_only && a.krange == KR_FULL) fl *= 0.5;          // SYRK: n^2 k
    if (a.lower_only && a.krange != KR_FULL) fl *= 1.0 / 3.0;    // LAUUM: n^3 / 3
    double tiles128 = (double)(a.M / 128) * (a.N / 128) * batch * (a.lower_only ? 0.5 : 1.0);
    if (a.cmode == C_COLSUMSQ ? g.variance_bm == 64 : tiles128 < g.bm64_factor * std::max(1, g.sm_count))
        return launch_gemm_bm<AL, BL, 64>(a, batch, cls, fl);
    return launch_gemm_bm<AL, BL, 128>(a, batch, cls, fl);
}

int launch_assemble(const AsmArgs& a, int rows_pad, int cols_pad, int batch) {
    size_t smem = ((size_t)(128 + ASM_BN) * a.d + (size_t)a.d * ASM_XT_LD + a.d + 2) * sizeof(double);
    dim3 grid(cols_pad / ASM_BN, rows_pad / 128, batch);
    double entries = (double)rows_pad * cols_pad * batch * (a.sym ? 0.5 : 1.0);
    ProfMark m = prof_begin(EXQ_PROF_ASSEMBLE, entries * (3.0 * a.d + 30.0), entries * 8.0);
    if (!g.fast_corr) assemble_kernel<-1><<<grid, ASM_THREADS, smem, g.stream>>>(a);
    else if (a.kernel_id == EXQ_KERNEL_MATERN52) assemble_kernel<EXQ_KERNEL_MATERN52><<<grid, ASM_THREADS, smem, g.stream>>>(a);
    else if (a.kernel_id == EXQ_KERNEL_SQEXP) assemble_kernel<EXQ_KERNEL_SQEXP><<<grid, ASM_THREADS, smem, g.stream>>>(a);
    else assemble_kernel<EXQ_KERNEL_PRODMAT52><<<grid, ASM_THREADS, smem, g.stream>>>(a);
    prof_end(m);
    return check_launch("assemble_kernel");
}

// ---- large GEMMs of the recursion on the int8 tensor cores (ozaki_gemm.cuh) --------------------
struct OzOperand {
    const double* p;   // sub-matrix origin inside a system (leading dimension Np, system stride Np^2)
    int trans;         // 0: p[row][k], 1: p[k][row]
    int mask;          // oz::MASK_*: which side of the operand's diagonal holds data
    const double* sumsq = nullptr;   // transposed operands: sum_k p[k][row]^2 per system (stride Np) if already known
    const uint32_t* colmax = nullptr;   // transposed operands: high word of max_k |p[k][row]| per system (stride Np) if already known
};

// digit planes of both operands (A in the first half of the per-system workspace, B in the second; one operand
// when A == B) must fit what sys_alloc sized for oz_factor_slices digits of one whole matrix
bool oz_planes_fit(const Sys& s, int digits, int M, int N, int K, bool same) {
    if (!s.oz_planes || s.oz_bytes < s.oz_per_sys * s.B) return false;
    if (same) return (size_t)digits * M * K <= s.oz_per_sys;
    return (size_t)digits * M * K <= s.oz_per_sys / 2 && (size_t)digits * N * K <= s.oz_per_sys / 2;
}

bool oz_gemm_wanted(const Sys& s, int M, int N, int K, bool same = false) {
    const bool deep = g.in_logpost && !g.oz_shallow;
    return !g.oz_off && (g.oz_factor_slices == 6 || g.oz_factor_slices == 7) && K >= (deep ? g.oz_min_k : std::max(g.oz_min_k, 1024)) &&
           K <= 16384 &&
           (M % 128) == 0 && (N % 64) == 0 && (K % oz::OZ_KB) == 0 && (int64_t)(M / 128) * (N / 64) * s.B >= (deep && g.oz_min_tiles >= 0 ? g.oz_min_tiles : g.sm_count) &&
           oz_planes_fit(s, g.oz_factor_slices, M, N, K, same);
}

// does the recursion over this workspace send any product to the int8 kernel?  (the top node has the largest ones)
bool factor_uses_int8(const Sys& s) {
    const int nb = s.Np / 128, n1 = ((nb + 1) / 2) * 128, n2 = s.Np - n1;
    return n2 > 0 && (oz_gemm_wanted(s, n2, n1, n1) || oz_gemm_wanted(s, n2, n2, n1, true) || oz_gemm_wanted(s, n2, n1, n2));
}

// ---- balanced work lists of the int8 GEMM (oz::make_schedule), one per launch shape, kept on the device ----------------
struct SchedEntry {
    const uint32_t* dev = nullptr;
    int n = 0, grid = 0;
    int built_for = 0, kept = 0;      // systems the list was built for, kept tiles per system
};
std::map<std::array<int, 6>, SchedEntry> g_sched;
uint32_t* g_sched_arena = nullptr;
size_t g_sched_cap = 0, g_sched_used = 0;          // in entries

// false: no list (arena unavailable) -- the kernel then walks the tiles arithmetically
bool sched_get(int M, int N, int K, int B, int krange, int lower_only, SchedEntry& e) {
    if (!g.oz_sched) return false;
    const int items = (M / oz::BLK_C) * (N / oz::BLK_R) * B;
    const int grid = std::min(items, g.sm_count);
    const std::array<int, 6> key{M, N, K, krange, lower_only, grid};
    auto it = g_sched.find(key);
    if (it == g_sched.end() || it->second.built_for < B) {
        if (!g_sched_arena) {
            g_sched_cap = (size_t)8 << 20;             // 32 MB of 4-byte entries
            if (cudaMalloc(&g_sched_arena, g_sched_cap * sizeof(uint32_t)) != cudaSuccess) {
                cudaGetLastError();
                g_sched_arena = nullptr; g_sched_cap = 0;
                return false;
            }
        }
        // built for the next power of two: the batches of a MAP fit shrink restart by restart, and every prefix of a
        // list is a balanced list (make_schedule), so one construction (1 ms for LAUUM at n = 4096, 16 systems) serves them all
        int Bq = 1;
        while (Bq < B) Bq <<= 1;
        if (grid < g.sm_count) Bq = B;                 // fewer items than SMs: the grid depends on B itself
        std::vector<uint32_t> list;
        SchedEntry ne;
        ne.kept = oz::make_schedule(M, N, K, Bq, krange, lower_only, grid, list);
        if (list.size() > g_sched_cap) return false;
        if (g_sched_used + list.size() > g_sched_cap) {   // arena full: every launch that reads the old lists must be done
            cudaStreamSynchronize(g.stream);
            g_sched.clear();
            g_sched_used = 0;
        }
        uint32_t* dst = g_sched_arena + g_sched_used;
        if (cudaMemcpyAsync(dst, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, g.stream) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        cudaStreamSynchronize(g.stream);               // once per shape; the host vector goes out of scope
        g_sched_used += list.size();
        ne.dev = dst; ne.n = (int)list.size(); ne.grid = grid; ne.built_for = Bq;
        g_sched[key] = ne;
        it = g_sched.find(key);
    }
    e = it->second;
    // the rounds that hold the first B systems
    const int64_t need = ((int64_t)B * e.kept + grid - 1) / grid * grid;
    e.n = (int)std::min<int64_t>(e.n, need);
    return true;
}

template <int S>
int oz_gemm_s(Sys& s, OzOperand A, OzOperand Bm, bool same, double* C, int M, int N, int K, int krange, int lower_only,
              int cmode, double fl, bool reuse_a, const oz::GradArgs* ga = nullptr, int grad_kernel_id = -1,
              uint32_t* colmax_out = nullptr) {
    int rc;
    const int64_t ld = s.Np, sq = s.sq();
    int* expA = s.oz_exp;
    int* expB = same ? expA : s.oz_exp + (int64_t)s.B * s.Np;
    int8_t* planesA = s.oz_planes;
    int8_t* planesB = same ? planesA : s.oz_planes + (s.oz_per_sys / 2) * s.B;
    // reuse_a: the digit planes of A are still in place from the previous product of this node
    ProfMark m = prof_begin(EXQ_PROF_SLICE, 0.0, (double)s.B * K * ((reuse_a ? 0 : M) + (same ? 0 : N)) * (16.0 + S));
    if (!reuse_a) oz::slice_operand<S>(A.p, ld, sq, M, K, s.B, A.trans, A.mask, expA, planesA, g.stream, A.sumsq, s.Np);
    if (!same) oz::slice_operand<S>(Bm.p, ld, sq, N, K, s.B, Bm.trans, Bm.mask, expB, planesB, g.stream, nullptr, 0, Bm.colmax, s.Np);
    prof_end(m);
    g.launches += (reuse_a ? 0 : (A.trans ? 2 : 1)) + (same ? 0 : (Bm.trans ? 2 : 1)) - 1;
    LAUNCH_CHECK("ozaki operand slicing");
    oz::GArgs a{};
    a.expA = expA; a.expB = expB; a.C = C; a.ldc = ld; a.strideC = sq;
    a.M = M; a.N = N; a.K = K; a.B = s.B; a.krange = krange; a.lower_only = lower_only; a.cmode = cmode;
    a.heavy_last = (krange == KR_LE_J || krange == KR_LE_I) ? 1 : 0;
    if (ga) a.g = *ga;
    if (colmax_out && cmode == C_STORE) {       // column maxima of C for the product that slices it transposed
        cudaMemset2DAsync(colmax_out, (size_t)s.Np * sizeof(uint32_t), 0, (size_t)N * sizeof(uint32_t), (size_t)s.B, g.stream);
        a.colmax = colmax_out; a.colmax_stride = s.Np;
    }
    SchedEntry se;
    if (sched_get(M, N, K, s.B, krange, lower_only, se)) { a.sched = se.dev; a.n_sched = se.n; a.sched_grid = se.grid; }
    a.sys_slow = 1;   // also for SYRK / LAUUM: at n = 4096 LAUUM streams 12.5 GB per launch otherwise (bench: 240 -> 230 ms)
    // int8 operations issued: every non-skipped (128 x 64) tile runs S(S+1)/2 digit products over its k range --
    // counted by the same tile walk the kernel does (round 1 counted LAUUM twice: 2/3 instead of 1/3 of the cube)
    m = prof_begin(ga ? EXQ_PROF_LAUUM_GRAD : EXQ_PROF_GEMM_FACTOR_I8, fl, oz::gemm_issued_ops(M, N, K, s.B, krange, lower_only, S));
    if (ga) {
        a.g.slots = std::min(g.sm_count, (M / oz::BLK_C) * (N / oz::BLK_R) * s.B) * oz::EPI_WARPS;
        cudaMemsetAsync(a.g.partial, 0, (size_t)s.B * a.g.slots * (oz::GRAD_D + 2) * sizeof(double), g.stream);
    }
    rc = oz::launch_gemm<S>(planesA, planesB, a, g.sm_count, g.stream, grad_kernel_id);
    prof_end(m);
    if (!rc && ga) {
        oz::grad_reduce_slots_kernel<<<s.B, 32, 0, g.stream>>>(a.g.partial, a.g.slots, a.g.d, s.grad);
        g.launches++;
    }
    if (rc)
        return fail(EXQ_ERR_CUDA, "cuTensorMapEncodeTiled failed for the int8 GEMM: CUresult " + std::to_string(rc) + " M=" +
                                      std::to_string(M) + " N=" + std::to_string(N) + " K=" + std::to_string(K) + " B=" +
                                      std::to_string(s.B) + " planes=" + std::to_string((unsigned long long)(uintptr_t)planesA) +
                                      "/" + std::to_string((unsigned long long)(uintptr_t)planesB));
    return check_launch("oz::gemm_kernel");
}

int oz_gemm(Sys& s, OzOperand A, OzOperand Bm, bool same, double* C, int M, int N, int K, int krange, int lower_only,
            int cmode, bool reuse_a = false, int digits = 0, const oz::GradArgs* ga = nullptr, int grad_kernel_id = -1,
            uint32_t* colmax_out = nullptr) {
    double fl = 2.0 * M * (double)N * K * s.B;
    if (krange != KR_FULL) fl *= 0.5;
    if (lower_only && krange == KR_FULL) fl *= 0.5;
    if (lower_only && krange != KR_FULL) fl *= 1.0 / 3.0;
    if (digits == 0) digits = g.oz_factor_slices;
    if ((digits != 6 && digits != 7) || !oz_planes_fit(s, digits, M, N, K, same))
        return fail(EXQ_ERR_STATE, "int8 GEMM: " + std::to_string(digits) + " digits of a " + std::to_string(M) + " x " +
                                       std::to_string(K) + " operand do not fit the digit-plane workspace");
    if (digits == 6) return oz_gemm_s<6>(s, A, Bm, same, C, M, N, K, krange, lower_only, cmode, fl, reuse_a, ga, grad_kernel_id, colmax_out);
    return oz_gemm_s<7>(s, A, Bm, same, C, M, N, K, krange, lower_only, cmode, fl, reuse_a, ga, grad_kernel_id, colmax_out);
}

// potrf + trtri of the diagonal block [r0, r0+n) of every system in s by ONE launch of fused_factor_kernel
int launch_fused(Sys& s, int r0, int n) {
    FusedArgs a{};
    a.A = s.A; a.Linv = s.Linv; a.ld = s.Np; a.stride = s.sq(); a.info = s.info;
    int G = g.fuse_g;
    if (G <= 0) G = (g.fuse_np16 && s.B <= g.fuse_clusters16) ? 16 : 8;
    if (G > 8 && !g.fuse_np16) G = 8;
    a.n_ops = 0;
    if (!fused_build(a.ops, a.n_ops, r0, n, G)) return fail(EXQ_ERR_STATE, "fused sub-tree: step list overflow");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G, s.B, 1);
    cfg.blockDim = dim3(FUSED_THREADS, 1, 1);
    cfg.dynamicSmemBytes = FUSED_SMEM_BYTES;
    cfg.stream = g.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // potrf + trtri of an n x n block: 2/3 n^3 flop
    ProfMark m = prof_begin(EXQ_PROF_FUSED, s.B * (2.0 / 3.0) * (double)n * n * n, 0.0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, fused_factor_kernel, a);
    prof_end(m);
    if (e != cudaSuccess) return fail(EXQ_ERR_CUDA, std::string("fused_factor_kernel: ") + cudaGetErrorString(e));
    return check_launch("fused_factor_kernel");
}

// fused recursive potrf + trtri on the diagonal block [r0, r0+n) of every system in s
int factor_rec(Sys& s, int r0, int n) {
    const int64_t ld = s.Np, sq = s.sq();
    if (n > 128 && n <= g.fuse_n) return launch_fused(s, r0, n);
    if (n == 128) {
        ProfMark m = prof_begin(EXQ_PROF_LEAF, s.B * (128.0 * 128 * 128), 0.0);
        leaf_potrf_trtri_kernel<<<s.B, LEAF_THREADS, LEAF_SMEM_BYTES, g.stream>>>(s.A, s.Linv, ld, sq, sq, r0,
                                                                               s.info);
        prof_end(m);
        return check_launch("leaf_potrf_trtri_kernel");
    }
    const int nb = n / 128;
    const int n1 = ((nb + 1) / 2) * 128, n2 = n - n1;
    int rc = factor_rec(s, r0, n1);
    if (rc) return rc;
    double* A21 = s.A + (int64_t)(r0 + n1) * ld + r0;
    double* A22 = s.A + (int64_t)(r0 + n1) * ld + (r0 + n1);
    double* Li11 = s.Linv + (int64_t)r0 * ld + r0;
    double* Li21 = s.Linv + (int64_t)(r0 + n1) * ld + r0;
    double* Li22 = s.Linv + (int64_t)(r0 + n1) * ld + (r0 + n1);
    GemmArgs a{};
    a.lda = a.ldb = a.ldc = ld;
    a.strideA = a.strideB = a.strideC = sq;
    // (1) L21 := A21 * Linv11^T, parked in the Linv21 slot
    a.A = A21; a.B = Li11; a.C = Li21; a.M = n2; a.N = n1; a.K = n1;
    a.krange = KR_LE_J; a.lower_only = 0; a.cmode = C_STORE;
    const bool oz1 = oz_gemm_wanted(s, n2, n1, n1), oz2 = oz_gemm_wanted(s, n2, n1, n2);
    if (oz1) {
        if ((rc = oz_gemm(s, {A21, 0, oz::MASK_NONE}, {Li11, 0, oz::MASK_LE}, false, Li21, n2, n1, n1, KR_LE_J, 0, C_STORE)))
            return rc;
    } else
    if ((rc = launch_gemm<LAYOUT_K, LAYOUT_K>(a, s.B, EXQ_PROF_GEMM_FACTOR))) return rc;
    // (2) A22 -= L21 * L21^T (lower tiles)
    a.A = Li21; a.B = Li21; a.C = A22; a.M = n2; a.N = n2; a.K = n1;
    a.krange = KR_FULL; a.lower_only = 1; a.cmode = C_SUB;
    const bool oz_syrk = oz_gemm_wanted(s, n2, n2, n1, true);
    if (oz_syrk) {
        if ((rc = oz_gemm(s, {Li21, 0, oz::MASK_NONE}, {Li21, 0, oz::MASK_NONE}, true, A22, n2, n2, n1, KR_FULL, 1, C_SUB)))
            return rc;
    } else
    if ((rc = launch_gemm<LAYOUT_K, LAYOUT_K>(a, s.B, EXQ_PROF_GEMM_FACTOR))) return rc;
    // (3) T := L21 * Linv11, parked in the (now free) A21 slot.  Needs only L21 and Linv11, so it runs before the
    // recursion on A22: the digit planes of L21 from (2) are then still in place for the int8 kernel.
    a.A = Li21; a.B = Li11; a.C = A21; a.M = n2; a.N = n1; a.K = n1;
    a.krange = KR_GE_J; a.lower_only = 0; a.cmode = C_STORE;
    if (oz1) {
        // the column maxima of T (by global column r0 + j: the nodes whose T is alive at the same time own disjoint columns)
        // are collected by this product's epilogue for (4), which slices T transposed
        if ((rc = oz_gemm(s, {Li21, 0, oz::MASK_NONE}, {Li11, 1, oz::MASK_GE}, false, A21, n2, n1, n1, KR_GE_J, 0, C_STORE,
                          /*reuse_a=*/oz_syrk, 0, nullptr, -1, (oz2 && g.oz_colmax) ? s.oz_colmax + r0 : nullptr)))
            return rc;
    } else
    if ((rc = launch_gemm<LAYOUT_K, LAYOUT_MN>(a, s.B, EXQ_PROF_GEMM_FACTOR))) return rc;
    if ((rc = factor_rec(s, r0 + n1, n2))) return rc;
    // (4) Linv21 := -Linv22 * T
    a.A = Li22; a.B = A21; a.C = Li21; a.M = n2; a.N = n1; a.K = n2;
    a.krange = KR_LE_I; a.lower_only = 0; a.cmode = C_NEG;
    if (oz2) {
        OzOperand Tt{A21, 1, oz::MASK_NONE};
        if (oz1 && g.oz_colmax) Tt.colmax = s.oz_colmax + r0;      // T came out of the int8 kernel: its column maxima are known
        if ((rc = oz_gemm(s, {Li22, 0, oz::MASK_LE}, Tt, false, Li21, n2, n1, n2, KR_LE_I, 0, C_NEG)))
            return rc;
    } else
    if ((rc = launch_gemm<LAYOUT_K, LAYOUT_MN>(a, s.B, EXQ_PROF_GEMM_FACTOR))) return rc;
    return EXQ_OK;
}

// alpha = K^-1 y, dinv = diag K^-1, {logdet, quad}; y given per system with stride
int solve_vectors(Sys& s, const double* y, int64_t strideY, int nvalid) {
    const int64_t ld = s.Np, sq = s.sq();
    ProfMark m = prof_begin(EXQ_PROF_STREAM, (double)s.B * s.Np * s.Np, (double)s.B * s.Np * s.Np * 4.0);
    trmv_lower_kernel<<<dim3(s.Np / 8, 1, s.B), 256, 0, g.stream>>>(s.Linv, ld, sq, y, strideY, s.tvec, s.Np, s.Np);
    prof_end(m);
    LAUNCH_CHECK("trmv_lower_kernel");
    m = prof_begin(EXQ_PROF_STREAM, 2.0 * s.B * s.Np * s.Np, (double)s.B * s.Np * s.Np * 4.0);
    trmv_lower_t_kernel<<<dim3(s.Np / 32, TRMV_T_SPLITS, s.B), 256, 0, g.stream>>>(s.Linv, ld, sq, s.tvec, s.Np, s.tpart,
                                                                                  s.Np);
    prof_end(m);
    LAUNCH_CHECK("trmv_lower_t_kernel");
    m = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
    trmv_t_reduce_kernel<<<dim3((s.Np + 255) / 256, 1, s.B), 256, 0, g.stream>>>(s.tpart, s.alpha, s.dinv, s.Np, s.Np);
    prof_end(m);
    LAUNCH_CHECK("trmv_t_reduce_kernel");
    m = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
    logdet_quad_kernel<<<s.B, 256, 0, g.stream>>>(s.A, ld, sq, y, strideY, s.alpha, s.Np, nvalid, s.ldq, s.dstat);
    prof_end(m);
    LAUNCH_CHECK("logdet_quad_kernel");
    return EXQ_OK;
}

// `steps` steps of iterative refinement of alpha = K^-1 y on the systems of s (X, y as in solve_vectors), then
// logdet / quad / diag(L) range again with the refined alpha.  r = y - K alpha from freshly evaluated covariance
// entries; the correction Linv^T (Linv r) reuses the triangular mat-vec kernels (kvec / tvec as scratch).
int refine_alpha(Sys& s, const double* X, int64_t strideX, const double* y, int64_t strideY, int nvalid, int kernel_id,
                 int steps) {
    const int64_t ld = s.Np, sq = s.sq();
    const size_t smem = ((size_t)32 * s.d + s.d + 2) * sizeof(double);
    for (int it = 0; it < steps; it++) {
        ProfMark m = prof_begin(EXQ_PROF_STREAM, (double)s.B * nvalid * nvalid * (3.0 * s.d + 32.0), 0.0);
        kresidual_kernel<<<dim3(s.Np / 32, 1, s.B), 256, smem, g.stream>>>(X, strideX, s.theta, s.d + 2, y, strideY, s.alpha, s.Np,
                                                                       s.kvec, nvalid, s.Np, s.d, kernel_id);
        prof_end(m);
        LAUNCH_CHECK("kresidual_kernel");
        m = prof_begin(EXQ_PROF_STREAM, (double)s.B * s.Np * s.Np, (double)s.B * s.Np * s.Np * 4.0);
        trmv_lower_kernel<<<dim3(s.Np / 8, 1, s.B), 256, 0, g.stream>>>(s.Linv, ld, sq, s.kvec, s.Np, s.tvec, s.Np, s.Np);
        prof_end(m);
        m = prof_begin(EXQ_PROF_STREAM, 2.0 * s.B * s.Np * s.Np, (double)s.B * s.Np * s.Np * 4.0);
        trmv_lower_t_kernel<<<dim3(s.Np / 32, TRMV_T_SPLITS, s.B), 256, 0, g.stream>>>(s.Linv, ld, sq, s.tvec, s.Np, s.tpart, s.Np);
        prof_end(m);
        m = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
        // the reduce kernel rewrites dinv with the same column sums of squares as before (they do not depend on the vector)
        trmv_t_reduce_kernel<<<dim3((s.Np + 255) / 256, 1, s.B), 256, 0, g.stream>>>(s.tpart, s.kvec, s.dinv, s.Np, s.Np);
        prof_end(m);
        m = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
        axpy_kernel<<<dim3((nvalid + 255) / 256, 1, s.B), 256, 0, g.stream>>>(s.alpha, s.kvec, s.Np, nvalid);
        prof_end(m);
        LAUNCH_CHECK("alpha refinement");
    }
    ProfMark m = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
    logdet_quad_kernel<<<s.B, 256, 0, g.stream>>>(s.A, ld, sq, y, strideY, s.alpha, s.Np, nvalid, s.ldq, s.dstat);
    prof_end(m);
    LAUNCH_CHECK("logdet_quad_kernel");
    return EXQ_OK;
}

// K^-1 (lower tiles) into s.W; digits: base-256 digits of the int8 product (0: the configured LAUUM digits);
// have_dinv: s.dinv holds the column sums of squares of Linv (solve_vectors ran on these systems)
// ga != nullptr: K^-1 is not stored -- the int8 kernel's epilogue contracts it with dK/dtheta (C_GRAD) and s.grad
// receives the gradient; only when lauum_fuses_grad() said so
int lauum(Sys& s, int digits = 0, bool have_dinv = false, const oz::GradArgs* ga = nullptr, int grad_kernel_id = -1) {
    GemmArgs a{};
    a.lda = a.ldb = a.ldc = s.Np;
    a.strideA = a.strideB = a.strideC = s.sq();
    a.A = s.Linv; a.B = s.Linv; a.C = s.W; a.M = a.N = a.K = s.Np;
    a.krange = KR_GE_I; a.lower_only = 1; a.cmode = C_STORE;
    if (oz_gemm_wanted(s, s.Np, s.Np, s.Np, true))
        return oz_gemm(s, {s.Linv, 1, oz::MASK_GE, have_dinv ? s.dinv : nullptr}, {s.Linv, 1, oz::MASK_GE}, true, s.W, s.Np, s.Np,
                       s.Np, KR_GE_I, 1, ga ? C_GRAD : C_STORE,
                       false, digits ? std::min(digits, g.oz_factor_slices) : std::min(g.oz_lauum_slices, g.oz_factor_slices),
                       ga, grad_kernel_id);
    if (ga) return fail(EXQ_ERR_STATE, "lauum: fused gradient needs the int8 kernel");
    return launch_gemm<LAYOUT_MN, LAYOUT_MN>(a, s.B, EXQ_PROF_GEMM_FACTOR);
}

int assemble_sym(Sys& s, const double* X, int64_t strideX, int nvalid, int kernel_id) {
    AsmArgs a{};
    a.Xi = X; a.Xj = X; a.theta = s.theta; a.out = s.A;
    a.ld = s.Np; a.strideOut = s.sq(); a.strideXi = strideX; a.strideXj = strideX; a.strideTheta = s.d + 2;
    a.n_i = nvalid; a.n_j = nvalid; a.d = s.d; a.kernel_id = kernel_id; a.sym = 1; a.scale_sigma2 = 1;
    return launch_assemble(a, s.Np, s.Np, s.B);
}

// kct_rows: rows of the FP64 covariance chunk actually needed (the fused PEI pipeline only keeps 128 for its <= 8-point tail)
int ensure_pred_scratch(exq_gp_t gp, int64_t want_chunk, int64_t kct_rows = -1) {
    const size_t need_k = (size_t)(kct_rows < 0 ? want_chunk : std::min(want_chunk, kct_rows)) * gp->Np, need_p = (size_t)(gp->Np / 32) * want_chunk;   // int8: 2 partials per 64 rows
    if (g.kct_elems < need_k) {
        cudaFree(g.KcT);
        g.KcT = nullptr; g.kct_elems = 0;
        CUDA_TRY(cudaMalloc(&g.KcT, need_k * sizeof(double)));
        g.kct_elems = need_k;
    }
    if (g.p_elems < need_p) {
        cudaFree(g.P);
        g.P = nullptr; g.p_elems = 0;
        CUDA_TRY(cudaMalloc(&g.P, need_p * sizeof(double)));
        g.p_elems = need_p;
    }
    gp->KcT = g.KcT;
    gp->P = g.P;
    gp->chunk = want_chunk;
    return EXQ_OK;
}

int ensure_small_scratch(exq_gp_t gp) {
    if (!gp->small_x) CUDA_TRY(cudaMalloc(&gp->small_x, (size_t)128 * gp->d * sizeof(double)));
    if (!gp->small_out) CUDA_TRY(cudaMalloc(&gp->small_out, (size_t)2 * 128 * sizeof(double)));
    return EXQ_OK;
}

int64_t default_chunk(const exq_gp_t gp, int64_t Mpad) {
    // keep the KcT chunk around 1 GiB and at least a few waves of tiles
    int64_t c = ((int64_t)1 << 27) / gp->Np;
    c = std::max<int64_t>(128, (c / 128) * 128);
    c = std::min<int64_t>(c, 32768);
    return std::min(c, Mpad);
}

void planes_park(int8_t* p, size_t bytes) {
    if (!p) return;
    if (g.plane_pool.size() < 4) g.plane_pool.emplace_back(bytes, p);
    else cudaFree(p);
}

// ---- variance product on the int8 tensor cores (ozaki_i8.cuh) ---------------------------------
int ensure_cand_planes(size_t need) {
    if (g.cand_planes_bytes < need) {
        cudaFree(g.cand_planes);
        g.cand_planes = nullptr; g.cand_planes_bytes = 0;
        CUDA_TRY(cudaMalloc(&g.cand_planes, need));
        g.cand_planes_bytes = need;
    }
    return EXQ_OK;
}

// sliced = true: the digit planes of this chunk are already in g.cand_planes (assemble_digits_kernel)
template <int S>
int variance_ozaki_s(exq_gp_t gp, int mc_pad, bool sliced = false) {
    const int Np = gp->Np;
    if (gp->planes_S != S) {
        // digits of Linv: once per fit
        if (!gp->linv_exp) CUDA_TRY(cudaMalloc(&gp->linv_exp, (size_t)Np * sizeof(int)));
        const size_t need_l = (size_t)S * Np * Np;
        if (gp->linv_planes_bytes < need_l) {
            planes_park(gp->linv_planes, gp->linv_planes_bytes);
            gp->linv_planes = nullptr; gp->linv_planes_bytes = 0;
            for (size_t i = 0; i < g.plane_pool.size(); i++)
                if (g.plane_pool[i].first >= need_l && g.plane_pool[i].first <= 2 * need_l) {
                    gp->linv_planes = g.plane_pool[i].second; gp->linv_planes_bytes = g.plane_pool[i].first;
                    g.plane_pool.erase(g.plane_pool.begin() + i);
                    break;
                }
            if (!gp->linv_planes) {
                CUDA_TRY(cudaMalloc(&gp->linv_planes, need_l));
                gp->linv_planes_bytes = need_l;
            }
        }
        ProfMark m = prof_begin(EXQ_PROF_SLICE, 0.0, (double)Np * Np * (8.0 + 0.5 * S));
        oz::row_exponent_lower_kernel<<<(Np + 7) / 8, 256, 0, g.stream>>>(gp->fit.Linv, Np, Np, gp->linv_exp);
        oz::launch_slice<S>(gp->fit.Linv, Np, Np, Np, gp->linv_exp, 0, 1, gp->linv_planes, g.stream);
        prof_end(m);
        LAUNCH_CHECK("ozaki slice Linv");
        g.launches++;
        gp->planes_S = S;
    }
    int rcp = ensure_cand_planes((size_t)S * gp->chunk * Np);
    if (rcp) return rcp;
    // covariance entries lie in [0, sigma^2]: one global exponent for the candidate side
    const int fexp = oz::scale_exponent(gp->cov);
    ProfMark m{};
    if (!sliced) {
        m = prof_begin(EXQ_PROF_SLICE, 0.0, (double)mc_pad * Np * (8.0 + S));
        oz::launch_slice<S>(gp->KcT, Np, mc_pad, Np, nullptr, fexp, 0, g.cand_planes, g.stream);
        prof_end(m);
        LAUNCH_CHECK("ozaki slice candidates");
    }
    // flops: FP64-equivalent Np^2 Mc; bytes field carries the int8 operations actually issued
    m = prof_begin(EXQ_PROF_GEMM_PREDICT, (double)Np * Np * mc_pad,
                   (double)Np * (Np + oz::BLK_R) * mc_pad * (S * (S + 1) / 2));
    int rc = oz::launch_colsumsq<S>(g.cand_planes, mc_pad, fexp, gp->linv_planes, gp->linv_exp, Np, gp->P, mc_pad,
                                    g.sm_count, g.stream);
    prof_end(m);
    if (rc) return fail(EXQ_ERR_CUDA, "cuTensorMapEncodeTiled failed for the int8 variance kernel: CUresult " + std::to_string(rc));
    LAUNCH_CHECK("oz::colsumsq_kernel");
    return EXQ_OK;
}
// digits of the variance product for this GP: the configured count, widened to 7 by the conditioning guard;
// 0 = FP64 DMMA GEMM (int8 switched off, K beyond the exact-accumulation bound, or the guard's fallback)
int variance_digits(const exq_gp_t gp) {
    if (g.oz_slices <= 0 || gp->Np > 16384 || gp->fit_on_dmma || gp->kappa >= g.guard_fallback) return 0;
    if (g.oz_slices < 7 && gp->kappa >= g.guard_widen) return 7;
    return g.oz_slices;
}
int variance_ozaki(exq_gp_t gp, int mc_pad, int digits, bool sliced = false) {
    switch (digits) {
        case 5: return variance_ozaki_s<5>(gp, mc_pad, sliced);
        case 7: return variance_ozaki_s<7>(gp, mc_pad, sliced);
        default: return variance_ozaki_s<6>(gp, mc_pad, sliced);
    }
}

// fused cross-covariance assembly: digit planes + partial means / repulsion products, no FP64 chunk (pei_fused.cuh)
template <int KID, int S>
int launch_assemble_digits(const AsmDigArgs& a) {
    const size_t smem = ((size_t)(128 + ASM_BN) * a.d + (size_t)a.d * ASM_XT_LD + a.d + 2 + ASM_BN) * sizeof(double);
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(assemble_digits_kernel<KID, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done = true;
    }
    dim3 grid(a.Np / ASM_BN, a.mc_pad / 128, 1);
    const double entries = (double)a.mc_pad * a.Np;
    ProfMark m = prof_begin(EXQ_PROF_ASSEMBLE, entries * (3.0 * a.d + 30.0), entries * S);
    assemble_digits_kernel<KID, S><<<grid, ASM_THREADS, smem, g.stream>>>(a);
    prof_end(m);
    return check_launch("assemble_digits_kernel");
}
template <int S>
int launch_assemble_digits_k(const AsmDigArgs& a) {
    if (a.kernel_id == EXQ_KERNEL_MATERN52) return launch_assemble_digits<EXQ_KERNEL_MATERN52, S>(a);
    if (a.kernel_id == EXQ_KERNEL_SQEXP) return launch_assemble_digits<EXQ_KERNEL_SQEXP, S>(a);
    return launch_assemble_digits<EXQ_KERNEL_PRODMAT52, S>(a);
}

// mean / var / (pei) of the fitted GP at device-resident candidates Xc[Mpad][d]
int predict_device(exq_gp_t gp, const double* Xc, int64_t M, int64_t Mpad, const double* rep, int n_rep,
                   double tmax, int want_pei, double* mean, double* var, double* pei) {
    const int Np = gp->Np;
    int64_t chunk = default_chunk(gp, Mpad);
    // int8 digits bound |D_t| <= S 2^14 K < 2^31 only up to K = 16384
    const int vdigits = variance_digits(gp);
    const bool use_oz = vdigits > 0;
    if (use_oz && vdigits != g.oz_slices && Mpad > 8) g.guard_widened++;
    const bool fused_pei = g.pei_fused && use_oz && g.fast_corr && (vdigits == 6 || vdigits == 7);
    int rc = ensure_pred_scratch(gp, chunk, fused_pei ? 128 : -1);
    if (rc) return rc;
    for (int64_t c0 = 0; c0 < M; c0 += chunk) {
        const int64_t mc_pad = std::min(chunk, Mpad - c0);
        const int64_t mc = std::min(mc_pad, M - c0);
        if (fused_pei && mc > 8) {
            const int nM = Np / ASM_BN;
            const size_t need = (size_t)2 * nM * chunk;
            if (g.pmr_elems < need) {
                cudaFree(g.PMR);
                g.PMR = nullptr; g.pmr_elems = 0;
                CUDA_TRY(cudaMalloc(&g.PMR, need * sizeof(double)));
                g.pmr_elems = need;
            }
            if ((rc = ensure_cand_planes((size_t)vdigits * gp->chunk * Np))) return rc;
            AsmDigArgs ad{};
            ad.Xc = Xc + c0 * gp->d; ad.X = gp->X; ad.theta = gp->fit.theta; ad.alpha = gp->fit.alpha;
            ad.planes = g.cand_planes; ad.PM = g.PMR; ad.PR = g.PMR + (size_t)nM * chunk; ad.ldp = mc_pad;
            ad.mc_pad = (int)mc_pad; ad.m_valid = (int)mc; ad.N = (int)gp->N; ad.Np = Np; ad.d = gp->d;
            ad.kernel_id = gp->kernel_id; ad.cand_exp = oz::scale_exponent(gp->cov);
            rc = (vdigits == 7) ? launch_assemble_digits_k<7>(ad) : launch_assemble_digits_k<6>(ad);
            if (rc) return rc;
            if ((rc = variance_ozaki(gp, (int)mc_pad, vdigits, /*sliced=*/true))) return rc;
            FinLightArgs f{};
            f.PM = ad.PM; f.PR = ad.PR; f.P = gp->P; f.ldp = mc_pad; f.nM = nM; f.nP = 2 * (Np / oz::BLK_R);
            f.theta = gp->fit.theta; f.Xc = Xc + c0 * gp->d; f.Rp = rep; f.n_rep = n_rep; f.d = gp->d;
            f.kernel_id = gp->kernel_id; f.m_valid = (int)mc; f.tmax = tmax; f.want_pei = want_pei;
            f.mean = mean ? mean + c0 : nullptr; f.var = var ? var + c0 : nullptr; f.pei = pei ? pei + c0 : nullptr;
            const size_t rp_bytes = (size_t)n_rep * gp->d * sizeof(double);
            f.rp_in_smem = (want_pei && rp_bytes <= 160 * 1024) ? 1 : 0;
            const size_t fsmem = (gp->d + 2) * sizeof(double) + (f.rp_in_smem ? rp_bytes : 0);
            static bool fl_attr = false;
            if (!fl_attr) {
                CUDA_TRY(cudaFuncSetAttribute(fin_light_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                fl_attr = true;
            }
            ProfMark mf = prof_begin(EXQ_PROF_STREAM, (double)mc * n_rep * (3.0 * gp->d + 30.0), (double)mc * (2.0 * nM + f.nP) * 8.0);
            fin_light_kernel<<<(unsigned)((mc * FIN_LANES + 127) / 128), 128, fsmem, g.stream>>>(f);
            prof_end(mf);
            LAUNCH_CHECK("fin_light_kernel");
            continue;
        }
        AsmArgs a{};
        a.Xi = Xc + c0 * gp->d; a.Xj = gp->X; a.theta = gp->fit.theta; a.out = gp->KcT;
        a.ld = Np; a.n_i = (int)mc; a.n_j = (int)gp->N; a.d = gp->d; a.kernel_id = gp->kernel_id;
        a.sym = 0; a.scale_sigma2 = 1;
        if ((rc = launch_assemble(a, (int)mc_pad, Np, 1))) return rc;
        if (mc <= 8) {
            // a handful of points (scalar predict calls from the unchanged designers): z = Linv k by
            // one triangular mat-vec per point instead of a one-column-tile GEMM
            if (!gp->zsmall) CUDA_TRY(cudaMalloc(&gp->zsmall, (size_t)8 * Np * sizeof(double)));
            ProfMark m0 = prof_begin(EXQ_PROF_STREAM, (double)mc * Np * Np, (double)mc * Np * Np * 4.0);
            trmv_lower_kernel<<<dim3(Np / 8, 1, (unsigned)mc), 256, 0, g.stream>>>(gp->fit.Linv, Np, 0, gp->KcT, Np, gp->zsmall,
                                                                               Np, Np);
            prof_end(m0);
            LAUNCH_CHECK("trmv_lower_kernel");
            m0 = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
            sumsq_rows_kernel<<<(unsigned)mc, 256, 0, g.stream>>>(gp->zsmall, Np, Np, gp->P);
            prof_end(m0);
            LAUNCH_CHECK("sumsq_rows_kernel");
        } else if (use_oz) {
            if ((rc = variance_ozaki(gp, (int)mc_pad, vdigits))) return rc;
        } else {
        GemmArgs ga{};
        ga.A = gp->fit.Linv; ga.B = gp->KcT; ga.C = gp->P;
        ga.lda = Np; ga.ldb = Np; ga.ldc = mc_pad;
        ga.M = Np; ga.N = (int)mc_pad; ga.K = Np;
        ga.krange = KR_LE_I; ga.lower_only = 0; ga.cmode = C_COLSUMSQ; ga.raster = 8;
        if ((rc = launch_gemm<LAYOUT_K, LAYOUT_K>(ga, 1, EXQ_PROF_GEMM_PREDICT))) return rc;
        }
        FinArgs f{};
        f.KcT = gp->KcT; f.ld = Np; f.P = gp->P; f.ldp = mc_pad; f.nrb = (mc <= 8) ? 1 : (use_oz ? 2 * (Np / oz::BLK_R) : Np / g.variance_bm);
        f.alpha = gp->fit.alpha; f.theta = gp->fit.theta; f.Xc = Xc + c0 * gp->d; f.Rp = rep; f.n_rep = n_rep;
        f.N = (int)gp->N; f.Np = Np; f.d = gp->d; f.kernel_id = gp->kernel_id; f.m_valid = (int)mc;
        f.tmax = tmax; f.want_pei = want_pei;
        f.mean = mean ? mean + c0 : nullptr; f.var = var ? var + c0 : nullptr; f.pei = pei ? pei + c0 : nullptr;
        ProfMark m = prof_begin(EXQ_PROF_STREAM, (double)mc * gp->N * 4.0, (double)mc * Np * 8.0);
        predict_finalize_kernel<<<(unsigned)((mc + 7) / 8), 256, 0, g.stream>>>(f);
        prof_end(m);
        LAUNCH_CHECK("predict_finalize_kernel");
    }
    return EXQ_OK;
}

int upload_theta(Sys& s, const std::vector<double>& host) {
    CUDA_TRY(cudaMemcpyAsync(s.theta, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    return EXQ_OK;
}

// one assemble + factor pass over the first `B` systems of s; returns info on the host
int assemble_and_factor(Sys& s, const double* X, int64_t strideX, int nvalid, int kernel_id,
                        std::vector<int>& info_host) {
    int rc;
    CUDA_TRY(cudaMemsetAsync(s.info, 0, s.B * sizeof(int), g.stream));
    if ((rc = assemble_sym(s, X, strideX, nvalid, kernel_id))) return rc;
    if ((rc = factor_rec(s, 0, s.Np))) return rc;
    info_host.resize(s.B);
    CUDA_TRY(cudaMemcpyAsync(info_host.data(), s.info, s.B * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return EXQ_OK;
}

// systems [b0, b0 + B) of a workspace, as a workspace of their own
Sys sys_view(const Sys& s, int b0, int B, int dmax) {
    Sys v = s;
    const int64_t sq = s.sq();
    const int nt = s.Np / 128;
    v.B = B;
    v.A += (int64_t)b0 * sq; v.Linv += (int64_t)b0 * sq;
    if (v.W) v.W += (int64_t)b0 * sq;
    v.theta += (int64_t)b0 * (s.d + 2);
    v.tvec += (int64_t)b0 * s.Np; v.alpha += (int64_t)b0 * s.Np; v.dinv += (int64_t)b0 * s.Np; v.kvec += (int64_t)b0 * s.Np;
    v.tpart += (int64_t)b0 * TRMV_T_SPLITS * 2 * s.Np;
    v.ldq += 2 * b0; v.dstat += 2 * b0; v.info += b0; v.grad += (int64_t)b0 * (s.d + 2);
    if (v.gpartial) v.gpartial += (int64_t)b0 * nt * nt * (dmax + 2);
    if (v.oz_planes) {
        v.oz_planes += (size_t)b0 * s.oz_per_sys;
        v.oz_exp += (int64_t)2 * b0 * s.Np;
        v.oz_colmax += (int64_t)b0 * s.Np;
        v.oz_bytes = s.oz_per_sys * B;
    }
    return v;
}

// enqueue (no host synchronisation) the first half of an evaluation of the systems in s: assembly, potrf+trtri,
// alpha / diag K^-1 / logdet / quad / diag(L) range
int enqueue_factor(exq_gp_t gp, Sys& s, const double* theta_host) {
    const int N = (int)gp->N, d = gp->d, B = s.B;
    int rc;
    CUDA_TRY(cudaMemcpyAsync(s.theta, theta_host, (size_t)B * (d + 2) * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaMemsetAsync(s.info, 0, B * sizeof(int), g.stream));
    if ((rc = assemble_sym(s, gp->X, 0, N, gp->kernel_id))) return rc;
    if ((rc = factor_rec(s, 0, s.Np))) return rc;
    return solve_vectors(s, gp->y, 0, N);
}

// second half: K^-1 = Linv^T Linv (LAUUM) and the gradient contraction
int enqueue_grad(exq_gp_t gp, Sys& s, int dmax, int lauum_digits) {
    const int N = (int)gp->N, d = gp->d, Np = gp->Np, nt = Np / 128, B = s.B;
    int rc;
    // LAUUM on the int8 kernel, stationary kernel, d <= 8: the gradient contraction rides in the GEMM's epilogue
    if (g.grad_fused && d <= oz::GRAD_D && gp->kernel_id != EXQ_KERNEL_PRODMAT52 && g.fast_corr &&
        oz_gemm_wanted(s, Np, Np, Np, true) &&
        (size_t)g.sm_count * oz::EPI_WARPS * (oz::GRAD_D + 2) <= (size_t)nt * nt * (64 + 2)) {
        oz::GradArgs ga{};
        ga.X = gp->X; ga.theta = s.theta; ga.alpha = s.alpha; ga.ldv = Np; ga.n_valid = N; ga.d = d; ga.partial = s.gpartial;
        return lauum(s, lauum_digits, /*have_dinv=*/true, &ga, gp->kernel_id);
    }
    if ((rc = lauum(s, lauum_digits, /*have_dinv=*/true))) return rc;
    size_t smem = ((size_t)128 * d + (size_t)d * GRAD_XT_LD + d + 2 + 256 + 8 * (dmax + 2)) * sizeof(double);
    dim3 grid(nt, nt, B);
    ProfMark m = prof_begin(EXQ_PROF_STREAM, (double)B * N * N * 0.5 * (5.0 * d + 40.0), (double)B * N * N * 4.0);
#define EXQ_GRAD_ARGS gp->X, 0, s.theta, d + 2, s.W, Np, s.sq(), s.alpha, Np, N, d
    const bool fast = g.fast_corr && gp->kernel_id != EXQ_KERNEL_PRODMAT52 && dmax <= 16;
    if (fast && gp->kernel_id == EXQ_KERNEL_MATERN52 && dmax == 8)
        grad_fast_kernel<EXQ_KERNEL_MATERN52, 8><<<grid, 256, smem, g.stream>>>(EXQ_GRAD_ARGS, s.gpartial, nt);
    else if (fast && gp->kernel_id == EXQ_KERNEL_MATERN52)
        grad_fast_kernel<EXQ_KERNEL_MATERN52, 16><<<grid, 256, smem, g.stream>>>(EXQ_GRAD_ARGS, s.gpartial, nt);
    else if (fast && dmax == 8)
        grad_fast_kernel<EXQ_KERNEL_SQEXP, 8><<<grid, 256, smem, g.stream>>>(EXQ_GRAD_ARGS, s.gpartial, nt);
    else if (fast)
        grad_fast_kernel<EXQ_KERNEL_SQEXP, 16><<<grid, 256, smem, g.stream>>>(EXQ_GRAD_ARGS, s.gpartial, nt);
    else if (dmax == 8)
        grad_kernel<8><<<grid, 256, smem, g.stream>>>(EXQ_GRAD_ARGS, gp->kernel_id, s.gpartial, nt);
    else if (dmax == 16)
        grad_kernel<16><<<grid, 256, smem, g.stream>>>(EXQ_GRAD_ARGS, gp->kernel_id, s.gpartial, nt);
    else if (dmax == 32)
        grad_kernel<32><<<grid, 256, smem, g.stream>>>(EXQ_GRAD_ARGS, gp->kernel_id, s.gpartial, nt);
    else
        grad_kernel<64><<<grid, 256, smem, g.stream>>>(EXQ_GRAD_ARGS, gp->kernel_id, s.gpartial, nt);
#undef EXQ_GRAD_ARGS
    prof_end(m);
    LAUNCH_CHECK("grad_kernel");
    m = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
    grad_reduce_kernel<<<B, 128, 0, g.stream>>>(s.gpartial, nt, dmax, d, s.grad);
    prof_end(m);
    LAUNCH_CHECK("grad_reduce_kernel");
    return EXQ_OK;
}

int pei_argmax_launch(exq_gp_t gp, exq_cand_t c, const double* xnew_dev) {
    size_t smem = 2 * c->d * sizeof(double);
    ProfMark m = prof_begin(EXQ_PROF_STREAM, (double)c->M * (3.0 * c->d + 35.0),
                            xnew_dev ? (double)c->M * (c->d + 2) * 8.0 : (double)c->M * 8.0);
    pei_update_argmax_kernel<<<c->nblocks, 256, smem, g.stream>>>(c->pei, c->Xc, c->M, c->d, gp ? gp->kernel_id : 0,
                                                                  gp ? gp->fit.theta : nullptr, xnew_dev, c->partial);
    prof_end(m);
    LAUNCH_CHECK("pei_update_argmax_kernel");
    m = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
    argmax_final_kernel<<<1, 256, 0, g.stream>>>(c->partial, c->nblocks, c->best);
    prof_end(m);
    LAUNCH_CHECK("argmax_final_kernel");
    return EXQ_OK;
}



// one parked candidate set (device buffers only), reused by the next exq_cand_create of the same shape
exq_cand_s* g_cand_parked = nullptr;

void cand_free_buffers(exq_cand_s* c) {
    cudaFree(c->Xc); cudaFree(c->mean); cudaFree(c->var); cudaFree(c->pei);
    cudaFree(c->partial); cudaFree(c->best); cudaFree(c->chosen); cudaFree(c->xnew);
}

// lambda_max of a symmetric positive definite operator by power iteration; apply(x -> y) enqueues y = Op x
template <class Apply>
int power_iteration(Apply apply, double* x, double* y, double* norm_dev, int n, int nvalid, int iters, double* lambda) {
    // positive, non-constant start vector on the real rows, zero on the padding (which the operators keep at zero)
    std::vector<double> ones(n, 0.0);
    for (int i = 0; i < nvalid; i++) ones[i] = 1.0 + 0.37 * ((i * 2654435761u) % 1024) / 1024.0;
    CUDA_TRY(cudaMemcpyAsync(x, ones.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    normalize_kernel<<<1, 256, 0, g.stream>>>(x, n, norm_dev);
    for (int it = 0; it < iters; it++) {
        int rc = apply(x, y);
        if (rc) return rc;
        normalize_kernel<<<1, 256, 0, g.stream>>>(y, n, norm_dev);   // |Op x| with |x| = 1 -> lambda_max from below
        std::swap(x, y);
    }
    g.launches += 1 + iters;
    LAUNCH_CHECK("power iteration");
    CUDA_TRY(cudaMemcpyAsync(lambda, norm_dev, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return EXQ_OK;
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

const char* exq_last_error(void) { return g_err.c_str(); }

int exq_init(int device) {
    API_LOCK();
    if (g.ready && g.device == device) return EXQ_OK;
    if (g.ready) return fail(EXQ_ERR_STATE, "already initialised on another device");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(EXQ_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(EXQ_ERR_ARG, "bad device index");
    if (const char* e = getenv("EXQ_BLOCKING_SYNC"))
        if (atoi(e)) cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync);   // sleep instead of spinning in synchronisations
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(EXQ_ERR_CUDA, "libexauq_b200 is built for sm_100a only; found sm_" +
                                      std::to_string(prop.major) + std::to_string(prop.minor));
    g.sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    if (const char* e = getenv("EXQ_BM64_FACTOR")) g.bm64_factor = atof(e);
    if (const char* e = getenv("EXQ_VARIANCE_BM")) g.variance_bm = atoi(e) == 128 ? 128 : 64;
    if (const char* e = getenv("EXQ_OZAKI_FACTOR_SLICES")) {
        const int v = atoi(e);
        g.oz_factor_slices = (v == 6 || v == 7) ? v : 0;   // only S = 6 and 7 are instantiated (as exq_set_int8_digits)
    }
    if (const char* e = getenv("EXQ_FAST_CORR")) g.fast_corr = atoi(e) != 0;
    if (const char* e = getenv("EXQ_OZAKI_MIN_K")) g.oz_min_k = std::max(128, atoi(e));
    if (const char* e = getenv("EXQ_OZ_SCHED")) g.oz_sched = atoi(e) != 0;
    if (const char* e = getenv("EXQ_OZ_COLMAX")) g.oz_colmax = atoi(e) != 0;
    if (const char* e = getenv("EXQ_OZAKI_MIN_TILES")) g.oz_min_tiles = atoi(e);
    if (const char* e = getenv("EXQ_PEI_FUSED")) g.pei_fused = atoi(e) != 0;
    if (const char* e = getenv("EXQ_GRAD_FUSED")) g.grad_fused = atoi(e) != 0;
    if (const char* e = getenv("EXQ_FUSE_N")) { const int v = atoi(e); g.fuse_n = (v >= 256 && v <= 1024) ? (v / 128) * 128 : 0; }
    if (const char* e = getenv("EXQ_FUSE_G")) { const int v = atoi(e); g.fuse_g = (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) ? v : 0; }
    if (const char* e = getenv("EXQ_REFINE_STEPS")) g.refine_steps = std::max(0, std::min(4, atoi(e)));
    if (const char* e = getenv("EXQ_GUARD_WIDEN")) g.guard_widen = std::max(1.0, atof(e));
    if (const char* e = getenv("EXQ_GUARD_FALLBACK")) g.guard_fallback = std::max(g.guard_widen, atof(e));
    if (const char* e = getenv("EXQ_OZAKI_LAUUM_SLICES")) g.oz_lauum_slices = atoi(e) == 7 ? 7 : 6;
    if (const char* e = getenv("EXQ_OZAKI_SLICES")) {
        const int v = atoi(e);
        g.oz_slices = (v >= 5 && v <= 7) ? v : 0;
    }
    CUDA_TRY(cudaEventCreate(&g.t0));
    CUDA_TRY(cudaEventCreate(&g.t1));
    g.device = device;
    int rc = set_kernel_attrs();
    if (rc) return rc;
    g.ready = true;
    return EXQ_OK;
}

int exq_shutdown(void) {
    API_LOCK();
    if (!g.ready) return EXQ_OK;
    cudaStreamSynchronize(g.stream);
    sys_pool_clear();
    if (g_cand_parked) { cand_free_buffers(g_cand_parked); delete g_cand_parked; g_cand_parked = nullptr; }
    cudaFree(g.KcT); cudaFree(g.P); cudaFree(g.cand_planes); cudaFree(g.PMR);
    cudaFree(g_sched_arena);
    g_sched_arena = nullptr; g_sched_cap = g_sched_used = 0;
    g_sched.clear();
    plane_pool_clear();
    for (auto ev : g.pool) cudaEventDestroy(ev);
    g.pool.clear();
    cudaEventDestroy(g.t0);
    cudaEventDestroy(g.t1);
    cudaStreamDestroy(g.stream);
    g = Ctx();
    return EXQ_OK;
}

int exq_device_sm_count(void) { return g.sm_count; }
int exq_fused_info(int* max_rows, int* clusters8, int* clusters16) {
    REQUIRE_READY();
    if (max_rows) *max_rows = g.fuse_n;
    if (clusters8) *clusters8 = g.fuse_clusters8;
    if (clusters16) *clusters16 = g.fuse_clusters16;
    return EXQ_OK;
}
int exq_variance_slices(void) { return g.oz_slices; }
int exq_factor_slices(void) { return g.oz_factor_slices; }
int exq_set_int8_digits(int variance_digits, int factor_digits) {
    REQUIRE_READY();
    auto ok = [](int v) { return v == 0 || (v >= 5 && v <= 7); };
    if (!ok(variance_digits) || !(factor_digits == 0 || factor_digits == 6 || factor_digits == 7))
        return fail(EXQ_ERR_ARG, "exq_set_int8_digits: digits must be 0 (FP64 DMMA) or 5..7 (factorisation: 6 or 7)");
    cudaStreamSynchronize(g.stream);
    g.oz_slices = variance_digits;
    if (factor_digits != g.oz_factor_slices) {
        g.oz_factor_slices = factor_digits;
        sys_pool_clear();   // pooled workspaces carry digit planes sized for the old setting
    }
    return EXQ_OK;
}
int64_t exq_launch_count(void) { return g.launches; }

int exq_set_guard(double widen, double fallback) {
    REQUIRE_READY();
    if (!(widen >= 1.0) || !(fallback >= widen)) return fail(EXQ_ERR_ARG, "exq_set_guard: need 1 <= widen <= fallback");
    g.guard_widen = widen;
    g.guard_fallback = fallback;
    return EXQ_OK;
}
int exq_set_fused(int max_rows, int cluster_ctas) {
    REQUIRE_READY();
    if (!(max_rows == 0 || (max_rows >= 256 && max_rows <= 1024 && max_rows % 128 == 0)))
        return fail(EXQ_ERR_ARG, "exq_set_fused: max_rows must be 0 or a multiple of 128 in [256, 1024]");
    if (!(cluster_ctas == 0 || cluster_ctas == 1 || cluster_ctas == 2 || cluster_ctas == 4 || cluster_ctas == 8 ||
          cluster_ctas == 16))
        return fail(EXQ_ERR_ARG, "exq_set_fused: cluster_ctas must be 0 (automatic), 1, 2, 4, 8 or 16");
    g.fuse_n = max_rows;
    g.fuse_g = cluster_ctas;
    return EXQ_OK;
}
int exq_set_pei_fused(int on) {
    REQUIRE_READY();
    g.pei_fused = on != 0;
    return EXQ_OK;
}
int exq_set_grad_fused(int on) {
    REQUIRE_READY();
    g.grad_fused = on != 0;
    return EXQ_OK;
}
int exq_set_refine_steps(int steps) {
    REQUIRE_READY();
    if (steps < 0 || steps > 4) return fail(EXQ_ERR_ARG, "exq_set_refine_steps: 0..4");
    g.refine_steps = steps;
    return EXQ_OK;
}
int exq_guard_counts(int64_t* widened, int64_t* fallbacks) {
    API_LOCK();
    if (widened) *widened = g.guard_widened;
    if (fallbacks) *fallbacks = g.guard_fallbacks;
    return EXQ_OK;
}
int exq_gp_kappa(exq_gp_t gp, double* kappa, int* on_fp64) {
    REQUIRE_READY();
    if (!gp) return fail(EXQ_ERR_ARG, "exq_gp_kappa: bad argument");
    if (!gp->fitted) return fail(EXQ_ERR_STATE, "exq_gp_kappa: GP has not been fitted");
    if (kappa) *kappa = gp->kappa;
    if (on_fp64) *on_fp64 = gp->fit_on_dmma ? 1 : 0;
    return EXQ_OK;
}
double exq_int8_gemm_ops(int M, int N, int K, int B, int krange, int lower_only, int digits) {
    if (M < 128 || N < 64 || K < 64 || (M % 128) || (N % 64) || (K % 64) || B < 1 || digits < 1 || krange < 0 || krange > 4)
        return -1.0;
    return oz::gemm_issued_ops(M, N, K, B, krange, lower_only, digits);
}

int64_t exq_int8_gemm_schedule(int M, int N, int K, int B, int krange, int lower_only, int grid, uint32_t* out, int64_t cap) {
    if (M < 128 || N < 64 || K < 128 || (M % 128) || (N % 64) || (K % 128) || B < 1 || B > 1023 || grid < 1 || krange < 0 || krange > 4 ||
        M / 128 > 1023 || N / 64 > 1023)
        return -1;
    std::vector<uint32_t> list;
    oz::make_schedule(M, N, K, B, krange, lower_only, grid, list);
    if (out)
        for (int64_t i = 0; i < (int64_t)list.size() && i < cap; i++) out[i] = list[i];
    return (int64_t)list.size();
}

int exq_timer_start(void) {
    REQUIRE_READY();
    CUDA_TRY(cudaEventRecord(g.t0, g.stream));
    return EXQ_OK;
}
int exq_timer_stop(double* ms) {
    REQUIRE_READY();
    CUDA_TRY(cudaEventRecord(g.t1, g.stream));
    CUDA_TRY(cudaEventSynchronize(g.t1));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, g.t0, g.t1));
    if (ms) *ms = f;
    return EXQ_OK;
}

int exq_prof_enable(int mask) {
    REQUIRE_READY();
    if (mask && g.pool.empty()) {
        g.pool.resize(131072);
        for (auto& ev : g.pool) CUDA_TRY(cudaEventCreate(&ev));
    }
    g.prof_mask = (unsigned)mask;
    return EXQ_OK;
}
int exq_prof_reset(void) {
    REQUIRE_READY();
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    g.pool_used = 0;
    for (auto& p : g.prof) p = ProfClass();
    return EXQ_OK;
}
int exq_prof_get(int cls, double* total_ms, int64_t* launches, double* flops, double* bytes) {
    REQUIRE_READY();
    if (cls < 0 || cls >= EXQ_PROF_NCLASS) return fail(EXQ_ERR_ARG, "bad profile class");
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    double tot = 0.0;
    for (auto& pr : g.prof[cls].pairs) {
        float f = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&f, g.pool[pr.first], g.pool[pr.second]));
        tot += f;
    }
    // scale up if the event pool ran out before every launch could be timed
    size_t timed = g.prof[cls].pairs.size();
    if (timed > 0 && (int64_t)timed < g.prof[cls].launches) tot *= (double)g.prof[cls].launches / (double)timed;
    if (total_ms) *total_ms = tot;
    if (launches) *launches = g.prof[cls].launches;
    if (flops) *flops = g.prof[cls].flops;
    if (bytes) *bytes = g.prof[cls].bytes;
    return EXQ_OK;
}

// ---- correlation ------------------------------------------------------------------------
int exq_corr(int kernel_id, int d, const double* corr_raw, const double* A, int64_t n1, const double* B,
             int64_t n2, double* out) {
    REQUIRE_READY();
    if (d < 1 || d > EXQ_MAX_D || n1 < 1 || n2 < 1 || !corr_raw || !A || !B || !out)
        return fail(EXQ_ERR_ARG, "exq_corr: bad argument");
    if (kernel_id < 0 || kernel_id > 2) return fail(EXQ_ERR_ARG, "exq_corr: bad kernel id");
    const int p1 = pad128(n1), p2 = pad128(n2);
    double *dA = nullptr, *dB = nullptr, *dT = nullptr, *dO = nullptr;
    std::vector<double> th(d + 2);
    for (int k = 0; k < d; k++) th[k] = exp(corr_raw[k]);
    th[d] = 1.0; th[d + 1] = 0.0;
    int rc = EXQ_OK;
    auto cleanup = [&]() { cudaFree(dA); cudaFree(dB); cudaFree(dT); cudaFree(dO); };
#define CORR_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            cleanup();                                                                   \
            return fail(EXQ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
        }                                                                                \
    } while (0)
    CORR_TRY(cudaMalloc(&dA, (size_t)p1 * d * sizeof(double)));
    CORR_TRY(cudaMalloc(&dB, (size_t)p2 * d * sizeof(double)));
    CORR_TRY(cudaMalloc(&dT, (d + 2) * sizeof(double)));
    CORR_TRY(cudaMalloc(&dO, (size_t)p1 * p2 * sizeof(double)));
    CORR_TRY(cudaMemsetAsync(dA, 0, (size_t)p1 * d * sizeof(double), g.stream));
    CORR_TRY(cudaMemsetAsync(dB, 0, (size_t)p2 * d * sizeof(double), g.stream));
    CORR_TRY(cudaMemcpyAsync(dA, A, (size_t)n1 * d * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    CORR_TRY(cudaMemcpyAsync(dB, B, (size_t)n2 * d * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    CORR_TRY(cudaMemcpyAsync(dT, th.data(), (d + 2) * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    AsmArgs a{};
    a.Xi = dA; a.Xj = dB; a.theta = dT; a.out = dO; a.ld = p2;
    a.n_i = (int)n1; a.n_j = (int)n2; a.d = d; a.kernel_id = kernel_id; a.sym = 0; a.scale_sigma2 = 0;
    rc = launch_assemble(a, p1, p2, 1);
    if (rc) { cleanup(); return rc; }
    CORR_TRY(cudaMemcpy2DAsync(out, (size_t)n2 * sizeof(double), dO, (size_t)p2 * sizeof(double),
                               (size_t)n2 * sizeof(double), (size_t)n1, cudaMemcpyDeviceToHost, g.stream));
    CORR_TRY(cudaStreamSynchronize(g.stream));
    cleanup();
#undef CORR_TRY
    return EXQ_OK;
}

// ---- GP handle ----------------------------------------------------------------------------
int exq_gp_create(int64_t N, int d, const double* X, const double* y, int kernel_id, exq_gp_t* out) {
    REQUIRE_READY();
    if (N < 1 || d < 1 || d > EXQ_MAX_D || !X || !y || !out) return fail(EXQ_ERR_ARG, "exq_gp_create: bad argument");
    if (kernel_id < 0 || kernel_id > 2) return fail(EXQ_ERR_ARG, "exq_gp_create: bad kernel id");
    if (N > 65536) return fail(EXQ_ERR_ARG, "exq_gp_create: N too large");
    exq_gp_t gp = new exq_gp_s();
    gp->N = N; gp->d = d; gp->kernel_id = kernel_id; gp->Np = pad128(N);
    cudaError_t e;
    if ((e = cudaMalloc(&gp->X, (size_t)gp->Np * d * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&gp->y, (size_t)gp->Np * sizeof(double))) != cudaSuccess) {
        exq_gp_destroy(gp);
        return fail(EXQ_ERR_CUDA, std::string("exq_gp_create: ") + cudaGetErrorString(e));
    }
    cudaMemsetAsync(gp->X, 0, (size_t)gp->Np * d * sizeof(double), g.stream);
    cudaMemsetAsync(gp->y, 0, (size_t)gp->Np * sizeof(double), g.stream);
    cudaMemcpyAsync(gp->X, X, (size_t)N * d * sizeof(double), cudaMemcpyHostToDevice, g.stream);
    cudaMemcpyAsync(gp->y, y, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, g.stream);
    e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) {
        exq_gp_destroy(gp);
        return fail(EXQ_ERR_CUDA, std::string("exq_gp_create: ") + cudaGetErrorString(e));
    }
    *out = gp;
    return EXQ_OK;
}

int exq_gp_destroy(exq_gp_t gp) {
    API_LOCK();
    if (!gp) return EXQ_OK;
    if (g.ready) cudaStreamSynchronize(g.stream);
    cudaFree(gp->X); cudaFree(gp->y); cudaFree(gp->rep); cudaFree(gp->zsmall); cudaFree(gp->small_x); cudaFree(gp->small_out);
    planes_park(gp->linv_planes, gp->linv_planes_bytes);
    cudaFree(gp->linv_exp);
    sys_release(gp->fit);
    sys_release(gp->batch);
    delete gp;
    return EXQ_OK;
}

int exq_gp_fit(exq_gp_t gp, const double* corr_raw, double cov, double nugget, int nugget_mode,
               double* nugget_used, double* logdet, double* quad) {
    REQUIRE_READY();
    if (!gp || !corr_raw || !(cov > 0.0) || nugget < 0.0) return fail(EXQ_ERR_ARG, "exq_gp_fit: bad argument");
    const int d = gp->d;
    gp->fitted = false;   // before anything can fail: a failed refit must not leave the old state looking valid
    gp->planes_S = 0;
    int rc = sys_alloc(gp->fit, 1, (int)gp->N, d, false, false);
    if (rc) return rc;
    std::vector<double> th(d + 2);
    for (int k = 0; k < d; k++) {
        th[k] = exp(corr_raw[k]);
        if (!isfinite(th[k])) return fail(EXQ_ERR_ARG, "exq_gp_fit: overflow in exp(corr_raw)");
    }
    th[d] = cov;
    double nug = (nugget_mode == EXQ_NUGGET_ADAPTIVE) ? 0.0 : nugget;
    std::vector<int> info;
    const int max_tries = (nugget_mode == EXQ_NUGGET_ADAPTIVE) ? 6 : 1;
    double ldq[2], dst[2];
    gp->fit_on_dmma = false;
    for (int pass = 0; pass < 2; pass++) {
        // pass 1 only when the conditioning guard rejects an int8 factorisation: same system on FP64 DMMA
        g.oz_off = (pass == 1);
        nug = (nugget_mode == EXQ_NUGGET_ADAPTIVE) ? 0.0 : nugget;
        bool ok = false;
        for (int attempt = 0; attempt < max_tries; attempt++) {
            th[d + 1] = nug;
            if ((rc = upload_theta(gp->fit, th))) break;
            if ((rc = assemble_and_factor(gp->fit, gp->X, 0, (int)gp->N, gp->kernel_id, info))) break;
            if (info[0] == 0) { ok = true; break; }
            nug = (attempt == 0) ? cov * 1e-6 : nug * 10.0;   // mogp jit_cholesky
        }
        if (!rc && !ok) rc = fail(EXQ_NOT_PD, "covariance matrix is not positive definite");
        if (!rc) rc = solve_vectors(gp->fit, gp->y, 0, (int)gp->N);
        // alpha feeds every predictive mean: one refinement step whenever the inverse factor came from the int8 kernels
        if (!rc && g.refine_steps > 0 && factor_uses_int8(gp->fit))
            rc = refine_alpha(gp->fit, gp->X, 0, gp->y, 0, (int)gp->N, gp->kernel_id, g.refine_steps);
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(ldq, gp->fit.ldq, sizeof(ldq), cudaMemcpyDeviceToHost, g.stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dst, gp->fit.dstat, sizeof(dst), cudaMemcpyDeviceToHost, g.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
            if (e != cudaSuccess) rc = fail(EXQ_ERR_CUDA, std::string("exq_gp_fit: ") + cudaGetErrorString(e));
        }
        const bool was_int8 = factor_uses_int8(gp->fit);
        g.oz_off = false;
        if (rc) return rc;
        gp->kappa = (dst[0] > 0.0) ? (dst[1] / dst[0]) * (dst[1] / dst[0]) : INFINITY;
        if (pass == 0 && was_int8 && !(gp->kappa < g.guard_fallback)) {
            g.guard_fallbacks++;
            gp->fit_on_dmma = true;
            continue;
        }
        break;
    }
    gp->fitted = true;
    gp->cov = cov;
    gp->nugget = nug;
    gp->corr_raw.assign(corr_raw, corr_raw + d);
    if (nugget_used) *nugget_used = nug;
    if (logdet) *logdet = ldq[0];
    if (quad) *quad = ldq[1];
    return EXQ_OK;
}

// -------------------------------------------------------------------------------------
// nugget = "pivot" (pivot.cuh)
// -------------------------------------------------------------------------------------
namespace {
struct PivScratch {
    int *piv = nullptr, *rank = nullptr, *cand_i = nullptr;
    double* cand_v = nullptr;
    int blocks = 0;
    ~PivScratch() { cudaFree(piv); cudaFree(rank); cudaFree(cand_i); cudaFree(cand_v); }
};

// K (nugget-free, lower) of the handle's data into gp->fit.A, then the pivoted (pivoting = 1) or plain left-looking
// (pivoting = 0, first stop_at columns) Cholesky kernel; rank and, if wanted, the pivot order come back to the host
int pivot_factor(exq_gp_t gp, const double* corr_raw, double cov, int pivoting, int stop_at, PivScratch& sc,
                 int* rank_host, int* piv_host) {
    const int N = (int)gp->N, d = gp->d;
    if ((size_t)N * sizeof(double) > 200 * 1024) return fail(EXQ_ERR_ARG, "nugget = 'pivot': N <= 25600 on this backend");
    int rc = sys_alloc(gp->fit, 1, N, d, false, false);
    if (rc) return rc;
    Sys& s = gp->fit;
    std::vector<double> th(d + 2);
    for (int k = 0; k < d; k++) {
        th[k] = exp(corr_raw[k]);
        if (!isfinite(th[k])) return fail(EXQ_ERR_ARG, "pivot: overflow in exp(corr_raw)");
    }
    th[d] = cov; th[d + 1] = 0.0;
    if ((rc = upload_theta(s, th))) return rc;
    if ((rc = assemble_sym(s, gp->X, 0, N, gp->kernel_id))) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(pivot_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(trtri_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    const size_t smem = (size_t)N * sizeof(double);
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pivot_chol_kernel, PIV_THREADS, smem));
    if (per_sm < 1) return fail(EXQ_ERR_CUDA, "pivot_chol_kernel does not fit an SM");
    sc.blocks = g.sm_count * std::min(per_sm, 2);
    CUDA_TRY(cudaMalloc(&sc.piv, (size_t)N * sizeof(int)));
    CUDA_TRY(cudaMalloc(&sc.rank, sizeof(int)));
    CUDA_TRY(cudaMalloc(&sc.cand_i, (size_t)sc.blocks * sizeof(int)));
    CUDA_TRY(cudaMalloc(&sc.cand_v, (size_t)sc.blocks * sizeof(double)));
    PivArgs a{};
    a.A = s.A; a.ld = s.Np; a.n = N; a.pivoting = pivoting; a.stop_at = stop_at;
    a.dots = s.tvec; a.work = s.kvec; a.piv = sc.piv; a.rank = sc.rank; a.cand_v = sc.cand_v; a.cand_i = sc.cand_i;
    void* args[] = {&a};
    ProfMark m = prof_begin(EXQ_PROF_LEAF, (double)N * N * N / 3.0, 0.0);
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)pivot_chol_kernel, dim3(sc.blocks), dim3(PIV_THREADS), args, smem, g.stream));
    prof_end(m);
    LAUNCH_CHECK("pivot_chol_kernel");
    CUDA_TRY(cudaMemcpyAsync(rank_host, sc.rank, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    if (piv_host) CUDA_TRY(cudaMemcpyAsync(piv_host, sc.piv, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return EXQ_OK;
}
}  // namespace

int exq_gp_pivot_order(exq_gp_t gp, const double* corr_raw, double cov, int* piv, int* rank) {
    REQUIRE_READY();
    if (!gp || !corr_raw || !(cov > 0.0) || !piv || !rank) return fail(EXQ_ERR_ARG, "exq_gp_pivot_order: bad argument");
    gp->fitted = false;      // the workspace of the fit is used as scratch
    gp->planes_S = 0;
    PivScratch sc;
    return pivot_factor(gp, corr_raw, cov, 1, 0, sc, rank, piv);
}

int exq_gp_fit_pivoted(exq_gp_t gp, const double* corr_raw, double cov, int rank, double* logdet, double* quad) {
    REQUIRE_READY();
    if (!gp || !corr_raw || !(cov > 0.0) || rank < 1 || rank > gp->N)
        return fail(EXQ_ERR_ARG, "exq_gp_fit_pivoted: bad argument");
    const int N = (int)gp->N;
    gp->fitted = false;
    gp->planes_S = 0;
    PivScratch sc;
    int got = 0;
    int rc = pivot_factor(gp, corr_raw, cov, 0, rank, sc, &got, nullptr);
    if (rc) return rc;
    if (got != rank) return fail(EXQ_NOT_PD, "exq_gp_fit_pivoted: non-positive pivot inside the stated rank");
    Sys& s = gp->fit;
    pivot_fill_kernel<<<N, 128, 0, g.stream>>>(s.A, s.Np, N, sc.rank);
    LAUNCH_CHECK("pivot_fill_kernel");
    CUDA_TRY(cudaMemsetAsync(s.Linv, 0, (size_t)s.sq() * sizeof(double), g.stream));
    ProfMark m = prof_begin(EXQ_PROF_LEAF, (double)N * N * N / 3.0, 0.0);
    trtri_lower_kernel<<<s.Np, 128, (size_t)N * sizeof(double), g.stream>>>(s.A, s.Linv, s.Np, N, s.Np);
    prof_end(m);
    LAUNCH_CHECK("trtri_lower_kernel");
    CUDA_TRY(cudaMemsetAsync(s.info, 0, sizeof(int), g.stream));
    if ((rc = solve_vectors(s, gp->y, 0, N))) return rc;
    double ldq[2], dst[2];
    CUDA_TRY(cudaMemcpyAsync(ldq, s.ldq, sizeof(ldq), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaMemcpyAsync(dst, s.dstat, sizeof(dst), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    gp->kappa = (dst[0] > 0.0) ? (dst[1] / dst[0]) * (dst[1] / dst[0]) : INFINITY;
    gp->fit_on_dmma = true;
    gp->fitted = true;
    gp->cov = cov;
    gp->nugget = 0.0;
    gp->corr_raw.assign(corr_raw, corr_raw + gp->d);
    if (logdet) *logdet = ldq[0];
    if (quad) *quad = ldq[1];
    return EXQ_OK;
}

int exq_gp_predict(exq_gp_t gp, const double* Xc, int64_t M, double* mean, double* var) {
    REQUIRE_READY();
    if (!gp || !Xc || M < 1) return fail(EXQ_ERR_ARG, "exq_gp_predict: bad argument");
    if (!gp->fitted) return fail(EXQ_ERR_STATE, "exq_gp_predict: GP has not been fitted");
    if (M <= 128) {
        // scalar / few-point calls (PEICalculator.compute through the unchanged maximise: one predict per objective
        // evaluation, exauq/core/designers.py:522-706): no allocation, one upload, one download
        int rc = ensure_small_scratch(gp);
        if (rc) return rc;
        const int d = gp->d;
        CUDA_TRY(cudaMemsetAsync(gp->small_x, 0, (size_t)128 * d * sizeof(double), g.stream));
        CUDA_TRY(cudaMemcpyAsync(gp->small_x, Xc, (size_t)M * d * sizeof(double), cudaMemcpyHostToDevice, g.stream));
        rc = predict_device(gp, gp->small_x, M, 128, nullptr, 0, 0.0, 0, gp->small_out, gp->small_out + 128, nullptr);
        if (rc) return rc;
        if (mean) CUDA_TRY(cudaMemcpyAsync(mean, gp->small_out, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
        if (var) CUDA_TRY(cudaMemcpyAsync(var, gp->small_out + 128, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        return EXQ_OK;
    }
    exq_cand_t c = nullptr;
    int rc = exq_cand_create(M, gp->d, Xc, &c);
    if (rc) return rc;
    rc = predict_device(gp, c->Xc, M, c->Mpad, nullptr, 0, 0.0, 0, c->mean, c->var, nullptr);
    if (!rc && mean) rc = exq_cand_download(c, 0, mean);
    if (!rc && var) rc = exq_cand_download(c, 1, var);
    exq_cand_destroy(c);
    return rc;
}

int exq_gp_cross_cov(exq_gp_t gp, const double* Xq, int64_t n, double* out) {
    REQUIRE_READY();
    if (!gp || !Xq || !out || n < 1) return fail(EXQ_ERR_ARG, "exq_gp_cross_cov: bad argument");
    if (!gp->fitted) return fail(EXQ_ERR_STATE, "exq_gp_cross_cov: GP has not been fitted");
    const int d = gp->d, Np = gp->Np;
    const int64_t npad = ((n + 127) / 128) * 128;
    int rc = ensure_pred_scratch(gp, std::max<int64_t>(npad, 128));
    if (rc) return rc;
    double* xq = nullptr;
    const bool small = npad <= 128;
    if (small) {
        if ((rc = ensure_small_scratch(gp))) return rc;
        xq = gp->small_x;
    } else {
        CUDA_TRY(cudaMalloc(&xq, (size_t)npad * d * sizeof(double)));
    }
    cudaError_t e = cudaMemsetAsync(xq, 0, (size_t)npad * d * sizeof(double), g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(xq, Xq, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, g.stream);
    if (e == cudaSuccess) {
        AsmArgs a{};
        a.Xi = xq; a.Xj = gp->X; a.theta = gp->fit.theta; a.out = gp->KcT;
        a.ld = Np; a.n_i = (int)n; a.n_j = (int)gp->N; a.d = d; a.kernel_id = gp->kernel_id;
        a.sym = 0; a.scale_sigma2 = 1;
        rc = launch_assemble(a, (int)npad, Np, 1);
    }
    if (e == cudaSuccess && !rc)
        e = cudaMemcpy2DAsync(out, (size_t)gp->N * sizeof(double), gp->KcT, (size_t)Np * sizeof(double),
                              (size_t)gp->N * sizeof(double), (size_t)n, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (!small) cudaFree(xq);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(EXQ_ERR_CUDA, std::string("exq_gp_cross_cov: ") + cudaGetErrorString(e));
    return EXQ_OK;
}

int exq_gp_loo(exq_gp_t gp, double* mean, double* var, double* nes) {
    REQUIRE_READY();
    if (!gp) return fail(EXQ_ERR_ARG, "exq_gp_loo: bad argument");
    if (!gp->fitted) return fail(EXQ_ERR_STATE, "exq_gp_loo: GP has not been fitted");
    const int N = (int)gp->N;
    double* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, (size_t)3 * N * sizeof(double)));
    ProfMark m = prof_begin(EXQ_PROF_STREAM, 20.0 * N, 48.0 * N);
    loo_kernel<<<(N + 255) / 256, 256, 0, g.stream>>>(gp->y, gp->fit.alpha, gp->fit.dinv, N, out, out + N, out + 2 * N);
    prof_end(m);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && mean) e = cudaMemcpyAsync(mean, out, N * sizeof(double), cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess && var) e = cudaMemcpyAsync(var, out + N, N * sizeof(double), cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess && nes) e = cudaMemcpyAsync(nes, out + 2 * N, N * sizeof(double), cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    cudaFree(out);
    if (e != cudaSuccess) return fail(EXQ_ERR_CUDA, std::string("exq_gp_loo: ") + cudaGetErrorString(e));
    return EXQ_OK;
}

int exq_gp_loo_refit(exq_gp_t gp, const int* idx, int cnt, double* mean, double* var, double* nes) {
    REQUIRE_READY();
    if (!gp || !idx || cnt < 1) return fail(EXQ_ERR_ARG, "exq_gp_loo_refit: bad argument");
    if (!gp->fitted) return fail(EXQ_ERR_STATE, "exq_gp_loo_refit: GP has not been fitted");
    if (gp->N < 2) return fail(EXQ_ERR_ARG, "exq_gp_loo_refit: need at least 2 training points");
    for (int i = 0; i < cnt; i++)
        if (idx[i] < 0 || idx[i] >= gp->N) return fail(EXQ_ERR_ARG, "exq_gp_loo_refit: index out of range");
    const int Nm = (int)gp->N - 1, d = gp->d;
    const int Npm = pad128(Nm);
    // systems per pass: bounded by ~24 GiB of matrices
    int64_t per = (int64_t)2 * Npm * Npm * 8;
    int Bmax = (int)std::max<int64_t>(1, std::min<int64_t>(64, ((int64_t)24 << 30) / per));
    Bmax = std::min(Bmax, cnt);
    int rc = sys_alloc(gp->batch, Bmax, Nm, d, false, true);
    if (rc) return rc;
    Sys& s = gp->batch;
    std::vector<double> th(d + 2);
    for (int k = 0; k < d; k++) th[k] = exp(gp->corr_raw[k]);
    th[d] = gp->cov; th[d + 1] = gp->nugget;
    double* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, (size_t)3 * Bmax * sizeof(double)));
    for (int b0 = 0; b0 < cnt; b0 += Bmax) {
        const int B = std::min(Bmax, cnt - b0);
        const int Bsave = s.B;
        s.B = B;
        std::vector<double> thb((size_t)B * (d + 2));
        for (int b = 0; b < B; b++) std::copy(th.begin(), th.end(), thb.begin() + (size_t)b * (d + 2));
        cudaMemcpyAsync(s.theta, thb.data(), thb.size() * sizeof(double), cudaMemcpyHostToDevice, g.stream);
        cudaMemcpyAsync(s.skip, idx + b0, B * sizeof(int), cudaMemcpyHostToDevice, g.stream);
        ProfMark m = prof_begin(EXQ_PROF_STREAM, 0.0, (double)B * Npm * (d + 1) * 16.0);
        gather_skip_kernel<<<dim3((Npm + 255) / 256, B), 256, 0, g.stream>>>(gp->X, gp->y, (int)gp->N, d, s.skip, s.Xb,
                                                                          (int64_t)s.Np * d, s.yb, s.Np, s.Np);
        prof_end(m);
        std::vector<int> info;
        rc = check_launch("gather_skip_kernel");
        if (!rc) rc = assemble_and_factor(s, s.Xb, (int64_t)s.Np * d, Nm, gp->kernel_id, info);
        if (!rc)
            for (int b = 0; b < B; b++)
                if (info[b] != 0) rc = fail(EXQ_NOT_PD, "LOO system not positive definite");
        if (!rc) rc = solve_vectors(s, s.yb, s.Np, Nm);
        if (!rc && g.refine_steps > 0 && factor_uses_int8(s))
            rc = refine_alpha(s, s.Xb, (int64_t)s.Np * d, s.yb, s.Np, Nm, gp->kernel_id, g.refine_steps);
        if (!rc) {
            m = prof_begin(EXQ_PROF_STREAM, (double)B * Nm * (3.0 * d + 30.0), (double)B * Nm * d * 8.0);
            loo_cross_vec_kernel<<<dim3((Npm + 255) / 256, B), 256, 0, g.stream>>>(
                gp->X, s.Xb, (int64_t)s.Np * d, s.skip, s.theta, d + 2, (int)gp->N, d, gp->kernel_id, s.kvec, s.Np, s.Np);
            prof_end(m);
            rc = check_launch("loo_cross_vec_kernel");
        }
        if (!rc) {
            m = prof_begin(EXQ_PROF_STREAM, (double)B * s.Np * s.Np, (double)B * s.Np * s.Np * 4.0);
            trmv_lower_kernel<<<dim3(s.Np / 8, 1, B), 256, 0, g.stream>>>(s.Linv, s.Np, s.sq(), s.kvec, s.Np, s.tvec, s.Np,
                                                                       s.Np);
            prof_end(m);
            rc = check_launch("trmv_lower_kernel");
        }
        if (!rc) {
            m = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
            loo_refit_finish_kernel<<<B, 256, 0, g.stream>>>(s.kvec, s.tvec, s.alpha, s.Np, s.theta, d + 2, d, gp->y, s.skip,
                                                           s.Np, out, out + Bmax, out + 2 * Bmax);
            prof_end(m);
            rc = check_launch("loo_refit_finish_kernel");
        }
        cudaError_t e = cudaSuccess;
        if (!rc && mean) e = cudaMemcpyAsync(mean + b0, out, B * sizeof(double), cudaMemcpyDeviceToHost, g.stream);
        if (!rc && e == cudaSuccess && var) e = cudaMemcpyAsync(var + b0, out + Bmax, B * sizeof(double), cudaMemcpyDeviceToHost, g.stream);
        if (!rc && e == cudaSuccess && nes) e = cudaMemcpyAsync(nes + b0, out + 2 * Bmax, B * sizeof(double), cudaMemcpyDeviceToHost, g.stream);
        if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
        s.B = Bsave;
        if (!rc && e != cudaSuccess) rc = fail(EXQ_ERR_CUDA, std::string("exq_gp_loo_refit: ") + cudaGetErrorString(e));
        if (rc) break;
    }
    cudaFree(out);
    return rc;
}

int exq_gp_kinv(exq_gp_t gp, double* out) {
    REQUIRE_READY();
    if (!gp || !out) return fail(EXQ_ERR_ARG, "exq_gp_kinv: bad argument");
    if (!gp->fitted) return fail(EXQ_ERR_STATE, "exq_gp_kinv: GP has not been fitted");
    const int N = (int)gp->N, d = gp->d;
    Sys tmp;
    int rc = sys_alloc(tmp, 1, N, d, true, false);
    if (rc) { sys_free(tmp); return rc; }
    const int Np = tmp.Np;
    std::vector<double> th(d + 2);
    for (int k = 0; k < d; k++) th[k] = exp(gp->corr_raw[k]);
    th[d] = gp->cov; th[d + 1] = 0.0;
    std::vector<int> info;
    const char* singular = "Covariance Matrix is, or too close to singular.";
    double kappa = 1.0;
    bool on_dmma = false;
    for (int pass = 0; pass < 2 && !rc; pass++) {
        g.oz_off = (pass == 1);                  // pass 1: the conditioning guard sends the factorisation to FP64 DMMA
        rc = upload_theta(tmp, th);
        if (!rc) rc = assemble_and_factor(tmp, gp->X, 0, N, gp->kernel_id, info);
        const bool was_int8 = factor_uses_int8(tmp);
        g.oz_off = false;
        if (!rc && info[0] != 0) rc = fail(EXQ_NOT_PD, singular);
        if (rc) break;
        // cond_2(K) >= (max L_ii / min L_ii)^2: when even this lower bound reaches 1/eps the reference's check
        // (np.linalg.cond(k) >= 1/eps, exauq/core/modelling.py:1082-1093) refuses the matrix
        std::vector<double> diag(N);
        cudaError_t e = cudaMemcpy2DAsync(diag.data(), sizeof(double), tmp.A, (size_t)(Np + 1) * sizeof(double),
                                          sizeof(double), (size_t)N, cudaMemcpyDeviceToHost, g.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
        if (e != cudaSuccess) { rc = fail(EXQ_ERR_CUDA, std::string("exq_gp_kinv: ") + cudaGetErrorString(e)); break; }
        double lo = diag[0], hi = diag[0];
        for (int i = 1; i < N; i++) { lo = std::min(lo, diag[i]); hi = std::max(hi, diag[i]); }
        kappa = (lo > 0.0) ? (hi / lo) * (hi / lo) : INFINITY;
        if (!(kappa < 1.0 / 2.220446049250313e-16)) { rc = fail(EXQ_NOT_PD, singular); break; }
        if (pass == 0 && was_int8 && kappa >= g.guard_fallback) {
            g.guard_fallbacks++;
            on_dmma = true;
            continue;
        }
        break;
    }
    if (!rc) {
        // 2-norm condition estimate, the quantity the reference thresholds: lambda_max(K) by power iteration on a
        // freshly assembled K (in W, free until LAUUM), lambda_max(K^-1) = |Linv|_2^2 by power iteration on
        // Linv^T Linv.  Both converge from below, so cond is never over-estimated (no false refusals).
        AsmArgs a{};
        a.Xi = gp->X; a.Xj = gp->X; a.theta = tmp.theta; a.out = tmp.W;
        a.ld = Np; a.n_i = N; a.n_j = N; a.d = d; a.kernel_id = gp->kernel_id; a.sym = 0; a.scale_sigma2 = 1;
        rc = launch_assemble(a, Np, Np, 1);
        double lmax_k = 0.0, lmax_inv = 0.0;
        const int iters = 40;
        if (!rc)
            rc = power_iteration(
                [&](double* x, double* y) {
                    gemv_kernel<<<Np / 8, 256, 0, g.stream>>>(tmp.W, Np, x, y, Np);
                    return EXQ_OK;
                },
                tmp.tvec, tmp.kvec, tmp.ldq, Np, N, iters, &lmax_k);
        if (!rc)
            rc = power_iteration(
                [&](double* x, double* y) {
                    trmv_lower_kernel<<<dim3(Np / 8, 1, 1), 256, 0, g.stream>>>(tmp.Linv, Np, 0, x, 0, tmp.alpha, 0, Np);
                    trmv_lower_t_kernel<<<dim3(Np / 32, TRMV_T_SPLITS, 1), 256, 0, g.stream>>>(tmp.Linv, Np, 0, tmp.alpha, 0,
                                                                                              tmp.tpart, Np);
                    trmv_t_reduce_kernel<<<dim3((Np + 255) / 256, 1, 1), 256, 0, g.stream>>>(tmp.tpart, y, tmp.dinv, 0, Np);
                    g.launches += 2;
                    return EXQ_OK;
                },
                tmp.tvec, tmp.kvec, tmp.ldq, Np, N, iters, &lmax_inv);
        // (padding: K assembled with sym = 0 is zero there and Linv is the identity there with no coupling to the real
        // block, so a start vector that vanishes on the padding stays in the real subspace)
        if (!rc) {
            const double cond2 = std::max(lmax_k * lmax_inv, kappa);
            if (!(cond2 < 1.0 / 2.220446049250313e-16)) rc = fail(EXQ_NOT_PD, singular);
        }
    }
    if (!rc) {
        // kinv is multiplied into 1e-8-tolerance quantities by the unchanged multi-level path
        // (exauq/core/designers.py:1029): always 7 digits on the int8 path, FP64 DMMA when the guard says so
        g.oz_off = on_dmma;
        rc = lauum(tmp, 7);
        g.oz_off = false;
    }
    if (!rc) {
        cudaError_t e = cudaMemcpy2DAsync(out, (size_t)N * sizeof(double), tmp.W, (size_t)tmp.Np * sizeof(double),
                                          (size_t)N * sizeof(double), (size_t)N, cudaMemcpyDeviceToHost, g.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
        if (e != cudaSuccess) rc = fail(EXQ_ERR_CUDA, std::string("exq_gp_kinv: ") + cudaGetErrorString(e));
    }
    sys_release(tmp);
    if (rc) return rc;
    for (int i = 0; i < N; i++)   // mirror the lower triangle
        for (int j = i + 1; j < N; j++) out[(size_t)i * N + j] = out[(size_t)j * N + i];
    return EXQ_OK;
}

int exq_gp_logpost_batch(exq_gp_t gp, int R, const double* theta_raw, double nugget, int nugget_mode,
                         double* val, double* grad, int* info_out, double* nugget_used) {
    REQUIRE_READY();
    if (!gp || R < 1 || !theta_raw || !val) return fail(EXQ_ERR_ARG, "exq_gp_logpost_batch: bad argument");
    const int N = (int)gp->N, d = gp->d, Np = gp->Np;
    const bool fit_nug = (nugget_mode == EXQ_NUGGET_FIT);
    const int np_raw = d + 1 + (fit_nug ? 1 : 0);     // raw parameters per restart
    int64_t per = (int64_t)3 * Np * Np * 8;
    int Bmax = (int)std::max<int64_t>(1, std::min<int64_t>(128, ((int64_t)96 << 30) / per));
    Bmax = std::min(Bmax, R);
    int rc = sys_alloc(gp->batch, Bmax, N, d, true, false);
    if (rc) return rc;
    struct InLogpost {
        InLogpost() { g.in_logpost = true; }
        ~InLogpost() { g.in_logpost = false; }
    } in_logpost_scope;
    Sys& s = gp->batch;
    const int dmax = d <= 8 ? 8 : d <= 16 ? 16 : d <= 32 ? 32 : 64;
    const double ln2pi = 1.8378770664093453;
    for (int b0 = 0; b0 < R; b0 += Bmax) {
        const int B = std::min(Bmax, R - b0);
        const int Bsave = s.B;
        s.B = B;
        std::vector<double> thb((size_t)B * (d + 2));
        std::vector<double> nug(B, nugget_mode == EXQ_NUGGET_ADAPTIVE ? 0.0 : nugget);
        std::vector<int> bad(B, 0);
        for (int b = 0; b < B; b++) {
            const double* raw = theta_raw + (size_t)(b0 + b) * np_raw;
            if (fit_nug) {
                nug[b] = exp(raw[d + 1]);
                if (!isfinite(nug[b])) bad[b] = 1, nug[b] = 0.0;
            }
            for (int k = 0; k < d; k++) {
                thb[(size_t)b * (d + 2) + k] = exp(raw[k]);
                if (!isfinite(thb[(size_t)b * (d + 2) + k])) bad[b] = 1;
            }
            double cv = exp(raw[d]);
            if (!isfinite(cv) || !(cv > 0.0)) bad[b] = 1;
            if (bad[b]) {   // keep the kernels well defined; the result is flagged
                for (int k = 0; k < d; k++) thb[(size_t)b * (d + 2) + k] = 1.0;
                cv = 1.0;
            }
            thb[(size_t)b * (d + 2) + d] = cv;
            thb[(size_t)b * (d + 2) + d + 1] = nug[b];
        }
        // Phase A (all systems of the pass in every launch): assembly, potrf + trtri, alpha / logdet / quad.  The
        // host then reads info[] and the diag(L) range: failed factorisations are retried system by system with the
        // next jitter (adaptive nugget); systems whose conditioning proxy reaches guard_fallback are re-factored on
        // the FP64 DMMA path.  Phase B (LAUUM + gradient contraction) runs for the systems that are positive
        // definite, with 7 instead of 6 digits when any of them reaches guard_widen.  The one extra host round trip
        // between the phases costs ~30 us against >= 5 ms of device work per round.
        std::vector<int> info(B, 0);
        std::vector<double> dst((size_t)B * 2, 1.0);
        std::vector<char> on_dmma(B, 0);
        const int max_tries = (nugget_mode == EXQ_NUGGET_ADAPTIVE) ? 6 : 1;
        const bool int8_factor = factor_uses_int8(s);
        auto read_state = [&]() -> int {
            cudaError_t e = cudaMemcpyAsync(info.data(), s.info, B * sizeof(int), cudaMemcpyDeviceToHost, g.stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dst.data(), s.dstat, dst.size() * sizeof(double), cudaMemcpyDeviceToHost, g.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
            if (e != cudaSuccess) return fail(EXQ_ERR_CUDA, std::string("exq_gp_logpost_batch: ") + cudaGetErrorString(e));
            return EXQ_OK;
        };
        auto kappa_of = [&](int b) {
            const double lo = dst[2 * b], hi = dst[2 * b + 1];
            return (lo > 0.0) ? (hi / lo) * (hi / lo) : INFINITY;
        };
        for (int attempt = 0; attempt < max_tries && !rc; attempt++) {
            std::vector<std::pair<int, int>> groups;     // (first system, count) still to be evaluated
            if (attempt == 0) {
                groups.emplace_back(0, B);
            } else {
                for (int b = 0; b < B; b++)
                    if (info[b] != 0) groups.emplace_back(b, 1);
            }
            for (size_t k = 0; k < groups.size() && !rc; k++) {
                Sys v = sys_view(s, groups[k].first, groups[k].second, dmax);
                rc = enqueue_factor(gp, v, thb.data() + (size_t)groups[k].first * (d + 2));
            }
            if (!rc) rc = read_state();
            if (rc) break;
            bool any = false;
            for (int b = 0; b < B; b++) {
                if (info[b] != 0 && attempt + 1 < max_tries) {
                    double cv = thb[(size_t)b * (d + 2) + d];
                    nug[b] = (nug[b] == 0.0) ? cv * 1e-6 : nug[b] * 10.0;
                    thb[(size_t)b * (d + 2) + d + 1] = nug[b];
                    any = true;
                }
            }
            if (!any) break;
        }
        // conditioning guard: badly conditioned systems leave the int8 path (kappa >= guard_fallback) or keep it on the top
        // levels only (kappa >= guard_widen)
        if (!rc && int8_factor) {
            bool redo = false;
            const bool deep_rules = g.oz_min_k < 1024 || (g.oz_min_tiles >= 0 && g.oz_min_tiles < g.sm_count);
            for (int b = 0; b < B && !rc; b++) {
                if (bad[b] || info[b] != 0) continue;
                const double kap = kappa_of(b);
                if (kap >= g.guard_fallback) {
                    g.guard_fallbacks++;
                    on_dmma[b] = 1;
                    g.oz_off = true;
                } else if (kap >= g.guard_widen && deep_rules) {
                    g.guard_widened++;
                    g.oz_shallow = true;
                } else {
                    continue;
                }
                Sys v = sys_view(s, b, 1, dmax);
                rc = enqueue_factor(gp, v, thb.data() + (size_t)b * (d + 2));
                g.oz_off = false;
                g.oz_shallow = false;
                redo = true;
            }
            if (!rc && redo) rc = read_state();
        }
        if (!rc && grad) {
            int lauum_digits = 0;
            for (int b = 0; b < B; b++)
                if (!bad[b] && info[b] == 0 && !on_dmma[b] && kappa_of(b) >= g.guard_widen) lauum_digits = 7;
            if (lauum_digits == 7 && int8_factor && g.oz_lauum_slices < 7) g.guard_widened++;
            // maximal runs of systems that share a path: int8 (positive definite, not guarded) or FP64 DMMA
            for (int b = 0; b < B && !rc;) {
                if (bad[b] || info[b] != 0) { b++; continue; }
                int e = b + 1;
                while (e < B && !bad[e] && info[e] == 0 && on_dmma[e] == on_dmma[b]) e++;
                Sys v = sys_view(s, b, e - b, dmax);
                g.oz_off = on_dmma[b] != 0;
                rc = enqueue_grad(gp, v, dmax, lauum_digits);
                g.oz_off = false;
                b = e;
            }
        }
        std::vector<double> ldq((size_t)B * 2), gh((size_t)B * (d + 2));
        if (!rc) {
            cudaMemcpyAsync(ldq.data(), s.ldq, ldq.size() * sizeof(double), cudaMemcpyDeviceToHost, g.stream);
            if (grad) cudaMemcpyAsync(gh.data(), s.grad, gh.size() * sizeof(double), cudaMemcpyDeviceToHost, g.stream);
            cudaError_t e = cudaStreamSynchronize(g.stream);
            if (e != cudaSuccess) rc = fail(EXQ_ERR_CUDA, std::string("exq_gp_logpost_batch: ") + cudaGetErrorString(e));
        }
        s.B = Bsave;
        if (rc) return rc;
        for (int b = 0; b < B; b++) {
            const bool failed = bad[b] || info[b] != 0;
            val[b0 + b] = failed ? NAN : 0.5 * (ldq[2 * b] + ldq[2 * b + 1] + N * ln2pi);
            if (grad)
                for (int k = 0; k < np_raw; k++) grad[(size_t)(b0 + b) * np_raw + k] = failed ? NAN : gh[(size_t)b * (d + 2) + k];
            if (info_out) info_out[b0 + b] = failed ? EXQ_NOT_PD : 0;
            if (nugget_used) nugget_used[b0 + b] = nug[b];
        }
    }
    return EXQ_OK;
}

// ---- candidates ---------------------------------------------------------------------------
int exq_cand_create(int64_t M, int d, const double* Xc, exq_cand_t* out) {
    REQUIRE_READY();
    if (M < 1 || d < 1 || d > EXQ_MAX_D || !Xc || !out) return fail(EXQ_ERR_ARG, "exq_cand_create: bad argument");
    const int64_t Mpad = ((M + 127) / 128) * 128;
    if (g_cand_parked && g_cand_parked->Mpad == Mpad && g_cand_parked->d == d) {
        exq_cand_t c = g_cand_parked;
        g_cand_parked = nullptr;
        c->M = M;
        c->nblocks = (int)std::min<int64_t>((M + 255) / 256, 4 * std::max(1, g.sm_count));
        c->evaluated = false;
        cudaError_t e2 = cudaMemsetAsync(c->Xc, 0, (size_t)c->Mpad * d * sizeof(double), g.stream);
        if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(c->Xc, Xc, (size_t)M * d * sizeof(double), cudaMemcpyHostToDevice, g.stream);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(g.stream);
        if (e2 != cudaSuccess) {
            cand_free_buffers(c);
            delete c;
            return fail(EXQ_ERR_CUDA, std::string("exq_cand_create: ") + cudaGetErrorString(e2));
        }
        *out = c;
        return EXQ_OK;
    }
    exq_cand_t c = new exq_cand_s();
    c->M = M; c->d = d; c->Mpad = Mpad;
    c->nblocks = (int)std::min<int64_t>((M + 255) / 256, 4 * std::max(1, g.sm_count));
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void**)&c->Xc, (size_t)c->Mpad * d * sizeof(double));
    A((void**)&c->mean, (size_t)c->Mpad * sizeof(double));
    A((void**)&c->var, (size_t)c->Mpad * sizeof(double));
    A((void**)&c->pei, (size_t)c->Mpad * sizeof(double));
    A((void**)&c->partial, (size_t)c->nblocks * sizeof(ArgPair));
    A((void**)&c->best, sizeof(ArgPair));
    A((void**)&c->xnew, (size_t)d * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(c->Xc, 0, (size_t)c->Mpad * d * sizeof(double), g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->Xc, Xc, (size_t)M * d * sizeof(double), cudaMemcpyHostToDevice, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) {
        exq_cand_destroy(c);
        return fail(EXQ_ERR_CUDA, std::string("exq_cand_create: ") + cudaGetErrorString(e));
    }
    *out = c;
    return EXQ_OK;
}

int exq_cand_set_points(exq_cand_t c, const double* Xc) {
    REQUIRE_READY();
    if (!c || !Xc) return fail(EXQ_ERR_ARG, "exq_cand_set_points: bad argument");
    CUDA_TRY(cudaMemcpyAsync(c->Xc, Xc, (size_t)c->M * c->d * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    c->evaluated = false;
    return EXQ_OK;
}

int exq_cand_destroy(exq_cand_t c) {
    API_LOCK();
    if (!c) return EXQ_OK;
    if (g.ready) cudaStreamSynchronize(g.stream);
    if (g.ready && !g_cand_parked && c->Mpad >= 65536) {   // park large sets for reuse
        g_cand_parked = c;
        return EXQ_OK;
    }
    cand_free_buffers(c);
    delete c;
    return EXQ_OK;
}

int exq_pei_eval(exq_gp_t gp, exq_cand_t c, const double* rep, int n_rep, double tmax) {
    REQUIRE_READY();
    if (!gp || !c || n_rep < 0 || (n_rep > 0 && !rep)) return fail(EXQ_ERR_ARG, "exq_pei_eval: bad argument");
    if (!gp->fitted) return fail(EXQ_ERR_STATE, "exq_pei_eval: GP has not been fitted");
    if (c->d != gp->d) return fail(EXQ_ERR_ARG, "exq_pei_eval: dimension mismatch");
    if (n_rep > gp->rep_cap) {
        cudaFree(gp->rep);
        gp->rep = nullptr;
        gp->rep_cap = 0;
        CUDA_TRY(cudaMalloc(&gp->rep, (size_t)n_rep * gp->d * sizeof(double)));
        gp->rep_cap = n_rep;
    }
    if (n_rep > 0)
        CUDA_TRY(cudaMemcpyAsync(gp->rep, rep, (size_t)n_rep * gp->d * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    int rc = predict_device(gp, c->Xc, c->M, c->Mpad, gp->rep, n_rep, tmax, 1, c->mean, c->var, c->pei);
    if (rc) return rc;
    rc = pei_argmax_launch(gp, c, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    c->evaluated = true;
    return EXQ_OK;
}

int exq_pei_argmax(exq_cand_t c, int64_t* idx, double* val) {
    REQUIRE_READY();
    if (!c) return fail(EXQ_ERR_ARG, "exq_pei_argmax: bad argument");
    if (!c->evaluated) return fail(EXQ_ERR_STATE, "exq_pei_argmax: call exq_pei_eval first");
    ArgPair h;
    CUDA_TRY(cudaMemcpyAsync(&h, c->best, sizeof(h), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    if (h.i < 0 || h.i >= c->M)
        return fail(EXQ_ERR_STATE, "exq_pei_argmax: no finite pseudo-expected improvement among the candidates");
    if (idx) *idx = h.i;
    if (val) *val = h.v;
    return EXQ_OK;
}

int exq_pei_add_point(exq_gp_t gp, exq_cand_t c, const double* x) {
    REQUIRE_READY();
    if (!gp || !c || !x) return fail(EXQ_ERR_ARG, "exq_pei_add_point: bad argument");
    if (!c->evaluated) return fail(EXQ_ERR_STATE, "exq_pei_add_point: call exq_pei_eval first");
    CUDA_TRY(cudaMemcpyAsync(c->xnew, x, c->d * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    int rc = pei_argmax_launch(gp, c, c->xnew);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return EXQ_OK;
}

int exq_pei_batch(exq_gp_t gp, exq_cand_t c, int batch_size, int64_t* idx, double* val) {
    REQUIRE_READY();
    if (!gp || !c || batch_size < 1) return fail(EXQ_ERR_ARG, "exq_pei_batch: bad argument");
    if (!c->evaluated) return fail(EXQ_ERR_STATE, "exq_pei_batch: call exq_pei_eval first");
    if (c->chosen_cap < batch_size) {
        cudaFree(c->chosen);
        c->chosen = nullptr;
        c->chosen_cap = 0;
        CUDA_TRY(cudaMalloc(&c->chosen, (size_t)batch_size * sizeof(ArgPair)));
        c->chosen_cap = batch_size;
    }
    for (int s = 0; s < batch_size; s++) {
        ProfMark m = prof_begin(EXQ_PROF_STREAM, 0.0, 0.0);
        select_best_kernel<<<1, 64, 0, g.stream>>>(c->best, c->Xc, c->M, c->d, c->xnew, c->chosen, s);
        prof_end(m);
        LAUNCH_CHECK("select_best_kernel");
        int rc = pei_argmax_launch(gp, c, c->xnew);
        if (rc) return rc;
    }
    std::vector<ArgPair> h(batch_size);
    CUDA_TRY(cudaMemcpyAsync(h.data(), c->chosen, batch_size * sizeof(ArgPair), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    for (int s = 0; s < batch_size; s++) {
        if (h[s].i < 0 || h[s].i >= c->M)
            return fail(EXQ_ERR_STATE, "exq_pei_batch: no finite pseudo-expected improvement among the candidates");
        if (idx) idx[s] = h[s].i;
        if (val) val[s] = h[s].v;
    }
    return EXQ_OK;
}

int exq_cand_download(exq_cand_t c, int which, double* out) {
    REQUIRE_READY();
    if (!c || !out || which < 0 || which > 2) return fail(EXQ_ERR_ARG, "exq_cand_download: bad argument");
    const double* src = which == 0 ? c->mean : which == 1 ? c->var : c->pei;
    CUDA_TRY(cudaMemcpyAsync(out, src, (size_t)c->M * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return EXQ_OK;
}

void* exq_cand_best_devptr(exq_cand_t c) { return c ? (void*)c->best : nullptr; }

}  // extern "C"
