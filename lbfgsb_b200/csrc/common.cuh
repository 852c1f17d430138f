// Shared device-side definitions for the B200 L-BFGS-B engine.
//
// State block (DevState) = the reference's mainlb locals + the 2m x 2m matrices
// (src/lbfgsb.f90:416-424, :390-412) kept resident in HBM so that the whole
// iteration runs without host round trips; the host mirrors the scalar header
// once per setulb return to fill isave/dsave/lsave/task.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <math.h>

#include "../../include/lbfgsb_b200_shape.h"

typedef long long i64;

#define LB_MMAX 20               // largest history size m supported by the kernels
#define LB_KMAX (6 * LB_MMAX + 8)   // real-valued reduction slots per site (the merged freev+formk+wv record: 6m)
#define LB_IMAX 8                // integer reduction slots per site

// task codes (mainlb's character(60) Task, decoded on the host)
enum {
    TK_START = 0, TK_FG_START = 1, TK_FG_LNSRCH = 2, TK_NEW_X = 3, TK_CONV_PG = 4, TK_CONV_F = 5,
    TK_ABNORMAL = 6, TK_RESTART = 7, TK_STOP = 8,
    TK_ERR_N = 10, TK_ERR_M = 11, TK_ERR_FACTR = 12, TK_ERR_NBD = 13, TK_ERR_INFEAS = 14
};
// resume points of the general pipeline after a pause of the fast one (Engine::enqueue_body)
enum { PAUSE_NONE = 0, PAUSE_CLASSIFY = 1, PAUSE_WALK = 2, PAUSE_GCP_FREEV = 3, PAUSE_DELTA = 4, PAUSE_BACKTRACK = 5, PAUSE_LSINIT = 6 };
// dcsrch task codes (Csave)
enum {
    CS_START = 0, CS_FG = 1, CS_CONV = 2, CS_WARN_ROUND = 3, CS_WARN_XTOL = 4, CS_WARN_STPMAX = 5,
    CS_WARN_STPMIN = 6, CS_ERR_STP_LT_MIN = 10, CS_ERR_STP_GT_MAX = 11, CS_ERR_G_GE_0 = 12,
    CS_ERR_FTOL = 13, CS_ERR_GTOL = 14, CS_ERR_XTOL = 15, CS_ERR_STPMIN = 16, CS_ERR_STPMAX = 17,
    CS_BLANK = 99
};
__host__ __device__ inline bool cs_is_warn(int c) { return c >= CS_WARN_ROUND && c <= CS_WARN_STPMIN; }
__host__ __device__ inline bool cs_is_err(int c) { return c >= CS_ERR_STP_LT_MIN && c <= CS_ERR_STPMAX; }

// messages of mainlb / cauchy / subsm / lnsrlb that the host prints (iprint >= 0), logged in order by the
// scalar kernels during one setulb call (host_print.h)
enum {
    EV_ITER_BEGIN = 1,      // :603   'ITERATION n' (iprint >= 99); a = n
    EV_SUBGNORM0 = 2,       // :1246  'Subgnorm = 0.  GCP = X.'
    EV_CAUCHY_SINGULAR = 3, // :622   singular triangular system in cauchy's bmv
    EV_FORMK_FAIL = 4,      // :669
    EV_SUBSM_SINGULAR = 5,  // :697   cmprlb / subsm
    EV_BACKTRACK = 6,       // :2831-2832
    EV_ASCENT = 7,          // :2250  a = gd
    EV_LNSRCH_RESTART = 8,  // :754
    EV_SKIP = 9,            // :830   a = dr, b = ddum
    EV_FORMT_FAIL = 10      // :854
};
#define LB_EVMAX 24

template <typename T> struct Real;
template <> struct Real<double> {
    static constexpr int VEC = LBFGSB_VEC_F64;
    static constexpr int UNROLL = LBFGSB_UNROLL_F64;
    typedef double2 vec_t;
    typedef int2 ivec_t;
    typedef unsigned long long key_t;
    __host__ __device__ static double eps() { return DBL_EPSILON; }
};
template <> struct Real<float> {
    static constexpr int VEC = LBFGSB_VEC_F32;
    static constexpr int UNROLL = LBFGSB_UNROLL_F32;
    typedef float2 vec_t;
    typedef int2 ivec_t;
    typedef unsigned int key_t;
    __host__ __device__ static float eps() { return FLT_EPSILON; }
};

// ---------------------------------------------------------------------------
// Device-resident state.  Scalars first (mirrored to the host per return),
// small matrices after.
// ---------------------------------------------------------------------------
template <typename T>
struct DevState {
    // ---- pipeline control (device-side predication of the enqueued kernels) ----
    int go;            // 1: keep executing the enqueued pipeline; 0: a return point was reached
    int pause;         // fast pipeline (engine.cu "fast path"): a merged scalar kernel met a branch that the common-path
                       // sequence does not cover; everything enqueued after it returns at once and the host resumes the
                       // general pipeline at this stage (PAUSE_* below)
    int in_body;       // 1: the "prelims + first lnsrlb" kernels are enabled
    int restart;       // 1: L-BFGS memory was reset; host must enqueue the body again
    int need_walk;     // cauchy: the breakpoint walk (sort + scans) is required
    int cauchy_mode;   // 0: full classify pass; 1: xcp = x only (:607-611 or :1245-1249)
    int do_subspace, do_formk, do_delta, do_backtrack, do_update, do_step, do_restore;
    int fuse_uc;       // NEW_X entry: the S/Y update and cauchy's per-variable pass run as one fused kernel
    int classify_done; // cauchy's per-variable pass of this body was already done by that fused kernel
    int lsinit_done;   // d = z - x and lnsrlb's first-entry sums were already formed by k_subsm_lsinit
    int spec_step;     // k_subsm_lsinit wrote the stp = 1 trial point straight into x (z, xp, r left untouched)
    int step_done;     // ... and lnsrlb's first trial is that point: k_ls_step has nothing to do
    int do_unstep;     // ... but the line search failed at its first entry: k_ls_step puts x = t back
    int lazy_gcp;      // cauchy's d and xcp are not materialised: xcp = x + tsum*d, d = -g or 0 by iwhere
    int fuse_gf;       // the cauchy tail (:1515) and freev (:1980-2059) run inside k_formk_cmprlb
    int lazy_z;        // ... and xcp was not stored: xcp = x + tsum*d with d = -g where state bit 2 is set, else 0
    int p2p_timeout;   // a peer's record did not arrive (sharded runs over peer memory): the call ends in an error
    int z_in_x;        // line search: the Newton point z exists only as the current x (speculative step); persists over calls
    int save_z;        // this call's k_ls_step moves x away from it: copy it into z first (a later trial may ask for stp = 1 again)
    int task, csave, info;
    // ---- mainlb locals (:416-424) ----
    int col, head, itail, iupdat, iter, nfgv, nskip, ifun, iback, iword;
    int updatd, prjctd, cnstnd, boxed, wrk, bnded;
    int brackt, stage;            // dcsrch isave(1:2)
    int m;
    int gstage;                   // device-resident iteration (Engine::minimize_graph): which call ran last, 1 FG_LNSRCH entry, 2 NEW_X entry
    i64 n, nintol, nseg, nfree, nact, nenter, nleave, nbreak, nfreec, nbdd, errk;
    i64 ibkmin;                   // variable index (0-based) of the smallest breakpoint
    i64 ibd;                      // subsm backtrack: variable index of the binding bound
    i64 walk_J;                   // exit position in the sorted breakpoint list of the current round (-1: none yet)
    i64 walk_done;                // number of sorted breakpoints of the current round already processed
    // the walk proceeds in rounds over increasing ranges of t (cauchy_walk.cuh "rounds")
    i64 walk_base;                // breakpoints passed in the earlier rounds (all ranks)
    i64 walk_rcount;              // breakpoints of the current round (all ranks)
    i64 walk_lcount;              // ... of which on this rank (rank-local field)
    i64 walk_rem;                 // breakpoints not yet passed when the current round began (all ranks)
    i64 walk_cstart;              // start of the chunk that holds walk_J
    i64 walk_fixn;                // how many entries of this rank's current sorted list are fixed at the end of the round
    int walk_closed, pad1;        // the search is finished (exit found or every breakpoint passed)
    // equal breakpoints at the exit (cauchy_walk.cuh "heap replay"): the reference pops them in hpsolb's order
    int tie_redo;                 // the round found its exit inside a group of equal breakpoints: redo it in heap order
    int tie_round;                // the current round is such a group, listed in heap order
    i64 tie_limit;                // replay the heap only for nbreak <= tie_limit (0: never; ties then stay in variable order)
    i64 tie_events;               // how many exits fell inside a group of equal breakpoints but were not replayed
    unsigned long long tie_key;   // bit pattern of that breakpoint value
    i64 snap_base;                // walk_base at the start of the current round
    // message log of this setulb call (printed by the host in order)
    int ev_n, ev_code[LB_EVMAX];
    T ev_a[LB_EVMAX], ev_b[LB_EVMAX];
    i64 n_el;                     // entering + leaving rows compacted for formk
    T theta, fold, tol, dnorm, epsmch, gd, gdold, stp, stpmx, sbgnrm, dtd, xstep, f, rr, dr, ddum;
    T pgtol, factr, sbg_spec;
    T ls[13];                     // dcsrch dsave(1:13)
    T f1, f2, f2_org, dtm, tsum, bkmin;
    T dd_p, alpha;
    T walk_f1, walk_f2, walk_tlast, walk_tprev2;   // carries across walk chunks (and ranks)
    // ---- small matrices (column-major, leading dimension m or 2m as in the reference) ----
    T sy[LB_MMAX * LB_MMAX], ss[LB_MMAX * LB_MMAX], wt[LB_MMAX * LB_MMAX];
    T wn[4 * LB_MMAX * LB_MMAX], wn1[4 * LB_MMAX * LB_MMAX];
    T p[2 * LB_MMAX], c[2 * LB_MMAX], v[2 * LB_MMAX], wv[2 * LB_MMAX], a[2 * LB_MMAX];
    T p0[2 * LB_MMAX];            // p at the start of the walk
    T walkA[2 * LB_MMAX], walkB[2 * LB_MMAX];  // carries of the 2col-vector prefix sums
    T snapA[2 * LB_MMAX], snapB[2 * LB_MMAX];  // the carries as they were at the start of the current round
    T snap_f1, snap_f2, snap_tlast, snap_tprev2;
};

template <typename T> __device__ inline void ev_push(DevState<T>* s, int code, T a = (T)0, T b = (T)0) {
    const int k = s->ev_n;
    if (k < LB_EVMAX) { s->ev_code[k] = code; s->ev_a[k] = a; s->ev_b[k] = b; s->ev_n = k + 1; }
}

// ---------------------------------------------------------------------------
// Fixed-shape reductions (include/lbfgsb_b200_shape.h).
// ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int off);
template <> __device__ __forceinline__ double shfl_xor_t<double>(double v, int off) {
    return __shfl_xor_sync(0xffffffffu, v, off);
}
template <> __device__ __forceinline__ float shfl_xor_t<float>(float v, int off) {
    return __shfl_xor_sync(0xffffffffu, v, off);
}
template <> __device__ __forceinline__ i64 shfl_xor_t<i64>(i64 v, int off) {
    return __shfl_xor_sync(0xffffffffu, v, off);
}
template <> __device__ __forceinline__ int shfl_xor_t<int>(int v, int off) {
    return __shfl_xor_sync(0xffffffffu, v, off);
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + shfl_xor_t<T>(v, off);
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { T o = shfl_xor_t<T>(v, off); v = o > v ? o : v; }
    return v;
}
template <typename T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { T o = shfl_xor_t<T>(v, off); v = o < v ? o : v; }
    return v;
}

// Block-level sum of K per-thread accumulators -> part[k*GRID + blockIdx.x].
// smem must hold K * (BLOCK/32) values.  All threads must call.  Only the first LBFGSB_BLOCK
// threads contribute (the TMA-staged kernels carry an extra producer warp).
template <typename T, int K>
__device__ __forceinline__ void block_sum_store(const T (&acc)[K], int kcount, T* smem, T* part) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = LBFGSB_BLOCK / 32;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (k < kcount) {
            T v = warp_sum<T>(acc[k]);
            if (lane == 0 && w < NW) smem[k * NW + w] = v;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kcount; k += LBFGSB_BLOCK) {
        T s = smem[k * NW];
#pragma unroll
        for (int q = 1; q < NW; ++q) s = s + smem[k * NW + q];
        part[(i64)k * LBFGSB_GRID + blockIdx.x] = s;
    }
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ T block_max(T v, T* smem) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = LBFGSB_BLOCK / 32;
    v = warp_max<T>(v);
    if (lane == 0 && w < NW) smem[w] = v;
    __syncthreads();
    T s = smem[0];
#pragma unroll
    for (int q = 1; q < NW; ++q) s = smem[q] > s ? smem[q] : s;
    __syncthreads();
    return s;
}
template <typename T>
__device__ __forceinline__ T block_min(T v, T* smem) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = LBFGSB_BLOCK / 32;
    v = warp_min<T>(v);
    if (lane == 0 && w < NW) smem[w] = v;
    __syncthreads();
    T s = smem[0];
#pragma unroll
    for (int q = 1; q < NW; ++q) s = smem[q] < s ? smem[q] : s;
    __syncthreads();
    return s;
}
__device__ __forceinline__ i64 block_isum(i64 v, i64* smem) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = LBFGSB_BLOCK / 32;
    v = warp_sum<i64>(v);
    if (lane == 0 && w < NW) smem[w] = v;
    __syncthreads();
    i64 s = 0;
#pragma unroll
    for (int q = 0; q < NW; ++q) s += smem[q];
    __syncthreads();
    return s;
}
// lexicographic (value, index) minimum over the block
template <typename T>
__device__ __forceinline__ void block_argmin(T& v, i64& idx, T* smv, i64* smi) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = LBFGSB_BLOCK / 32;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        T ov = shfl_xor_t<T>(v, off);
        i64 oi = shfl_xor_t<i64>(idx, off);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0 && w < NW) { smv[w] = v; smi[w] = idx; }
    __syncthreads();
    v = smv[0]; idx = smi[0];
#pragma unroll
    for (int q = 1; q < NW; ++q)
        if (smv[q] < v || (smv[q] == v && smi[q] < idx)) { v = smv[q]; idx = smi[q]; }
    __syncthreads();
}

// ---- final stage: one warp finishes one slot over the GRID block partials ----
// Lane t combines partials t, t+32, ... serially, then the butterfly.  The partials are first fetched into registers
// in one batch (independent loads in flight together), then combined in the fixed order: same bits, a fraction of
// the latency of a load-add chain.
#define LB_FINAL_PER_LANE ((LBFGSB_GRID + 31) / 32)
template <typename V> __device__ __forceinline__ void final_fetch(const V* part, V (&v)[LB_FINAL_PER_LANE]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < LB_FINAL_PER_LANE; ++q) {
        const int b = lane + 32 * q;
        v[q] = part[b < LBFGSB_GRID ? b : lane];
    }
}
template <typename T> __device__ __forceinline__ T final_sum_warp(const T* part) {
    const int lane = threadIdx.x & 31;
    T v[LB_FINAL_PER_LANE];
    final_fetch<T>(part, v);
    T acc = (T)0;
#pragma unroll
    for (int q = 0; q < LB_FINAL_PER_LANE; ++q) if (lane + 32 * q < LBFGSB_GRID) acc = acc + v[q];
    return warp_sum<T>(acc);
}
template <typename T> __device__ __forceinline__ T final_max_warp(const T* part) {
    const int lane = threadIdx.x & 31;
    T v[LB_FINAL_PER_LANE];
    final_fetch<T>(part, v);
    T acc = v[0];
#pragma unroll
    for (int q = 0; q < LB_FINAL_PER_LANE; ++q) if (lane + 32 * q < LBFGSB_GRID) acc = v[q] > acc ? v[q] : acc;
    return warp_max<T>(acc);
}
template <typename T> __device__ __forceinline__ T final_min_warp(const T* part) {
    const int lane = threadIdx.x & 31;
    T v[LB_FINAL_PER_LANE];
    final_fetch<T>(part, v);
    T acc = v[0];
#pragma unroll
    for (int q = 0; q < LB_FINAL_PER_LANE; ++q) if (lane + 32 * q < LBFGSB_GRID) acc = v[q] < acc ? v[q] : acc;
    return warp_min<T>(acc);
}
__device__ __forceinline__ i64 final_isum_warp(const i64* part) {
    const int lane = threadIdx.x & 31;
    i64 v[LB_FINAL_PER_LANE];
    final_fetch<i64>(part, v);
    i64 acc = 0;
#pragma unroll
    for (int q = 0; q < LB_FINAL_PER_LANE; ++q) if (lane + 32 * q < LBFGSB_GRID) acc += v[q];
    return warp_sum<i64>(acc);
}
__device__ __forceinline__ i64 final_imin_warp(const i64* part) {
    const int lane = threadIdx.x & 31;
    i64 v[LB_FINAL_PER_LANE];
    final_fetch<i64>(part, v);
    i64 acc = 0x7fffffffffffffffLL;
#pragma unroll
    for (int q = 0; q < LB_FINAL_PER_LANE; ++q) if (lane + 32 * q < LBFGSB_GRID) acc = v[q] < acc ? v[q] : acc;
    return warp_min<i64>(acc);
}
__device__ __forceinline__ i64 final_imax_warp(const i64* part) {
    const int lane = threadIdx.x & 31;
    i64 v[LB_FINAL_PER_LANE];
    final_fetch<i64>(part, v);
    i64 acc = -0x7fffffffffffffffLL;
#pragma unroll
    for (int q = 0; q < LB_FINAL_PER_LANE; ++q) if (lane + 32 * q < LBFGSB_GRID) acc = v[q] > acc ? v[q] : acc;
    return warp_max<i64>(acc);
}
template <typename T>
__device__ __forceinline__ void final_argmin_warp(const T* pv, const i64* pi, T& v, i64& idx) {
    const int lane = threadIdx.x & 31;
    v = pv[lane < LBFGSB_GRID ? lane : 0];
    idx = pi[lane < LBFGSB_GRID ? lane : 0];
    for (int b = lane; b < LBFGSB_GRID; b += 32) {
        T ov = pv[b]; i64 oi = pi[b];
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        T ov = shfl_xor_t<T>(v, off);
        i64 oi = shfl_xor_t<i64>(idx, off);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

// Finish `count` sum slots with the warps of the calling block; result in out[k] (shared).
template <typename T>
__device__ __forceinline__ void final_sums(const T* part, int count, T* out) {
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int k = w; k < count; k += nw) {
        T s = final_sum_warp<T>(part + (i64)k * LBFGSB_GRID);
        if (lane == 0) out[k] = s;
    }
}

// ---------------------------------------------------------------------------
// Vector loads / stores with a ragged tail (user buffers are not padded).
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void ldv(const T* __restrict__ p, i64 base, i64 n, T (&out)[Real<T>::VEC]) {
    constexpr int VEC = Real<T>::VEC;
    if (base + VEC <= n) {
        typename Real<T>::vec_t q = *reinterpret_cast<const typename Real<T>::vec_t*>(p + base);
        const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = e[v];
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = (base + v < n) ? p[base + v] : (T)0;
    }
}
template <typename T>
__device__ __forceinline__ void ldvi(const int* __restrict__ p, i64 base, i64 n, int (&out)[Real<T>::VEC]) {
    constexpr int VEC = Real<T>::VEC;
    if (base + VEC <= n) {
        typename Real<T>::ivec_t q = *reinterpret_cast<const typename Real<T>::ivec_t*>(p + base);
        const int* e = reinterpret_cast<const int*>(&q);
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = e[v];
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = (base + v < n) ? p[base + v] : 0;
    }
}
template <typename T>
__device__ __forceinline__ void ldvb(const unsigned char* __restrict__ p, i64 base, i64 n,
                                     int (&out)[Real<T>::VEC]) {
    constexpr int VEC = Real<T>::VEC;
    if (base + VEC <= n) {
        if (VEC == 2) {
            unsigned short q = *reinterpret_cast<const unsigned short*>(p + base);
            out[0] = q & 0xff; out[1] = q >> 8;
        } else {
            unsigned int q = *reinterpret_cast<const unsigned int*>(p + base);
#pragma unroll
            for (int v = 0; v < VEC; ++v) out[v] = (q >> (8 * v)) & 0xff;
        }
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = (base + v < n) ? p[base + v] : 0;
    }
}
template <typename T>
__device__ __forceinline__ void stv(T* __restrict__ p, i64 base, i64 n, const T (&in)[Real<T>::VEC]) {
    constexpr int VEC = Real<T>::VEC;
    if (base + VEC <= n) {
        typename Real<T>::vec_t q;
        T* e = reinterpret_cast<T*>(&q);
#pragma unroll
        for (int v = 0; v < VEC; ++v) e[v] = in[v];
        *reinterpret_cast<typename Real<T>::vec_t*>(p + base) = q;
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) if (base + v < n) p[base + v] = in[v];
    }
}
template <typename T>
__device__ __forceinline__ void stvi(int* __restrict__ p, i64 base, i64 n, const int (&in)[Real<T>::VEC]) {
    constexpr int VEC = Real<T>::VEC;
    if (base + VEC <= n) {
        typename Real<T>::ivec_t q;
        int* e = reinterpret_cast<int*>(&q);
#pragma unroll
        for (int v = 0; v < VEC; ++v) e[v] = in[v];
        *reinterpret_cast<typename Real<T>::ivec_t*>(p + base) = q;
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) if (base + v < n) p[base + v] = in[v];
    }
}
template <typename T>
__device__ __forceinline__ void stvb(unsigned char* __restrict__ p, i64 base, i64 n,
                                     const int (&in)[Real<T>::VEC]) {
    constexpr int VEC = Real<T>::VEC;
    if (base + VEC <= n) {
        if (VEC == 2) {
            *reinterpret_cast<unsigned short*>(p + base) = (unsigned short)((in[0] & 0xff) | ((in[1] & 0xff) << 8));
        } else {
            unsigned int q = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) q |= (unsigned int)(in[v] & 0xff) << (8 * v);
            *reinterpret_cast<unsigned int*>(p + base) = q;
        }
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) if (base + v < n) p[base + v] = (unsigned char)in[v];
    }
}

// Tile walk of the fixed shape: body(base) sees VEC consecutive variables at `base`.
#define LB_FOR_TILES(T, n, base)                                                               \
    for (i64 _tl = blockIdx.x, _nt = ((n) + (i64)LBFGSB_BLOCK * Real<T>::VEC * Real<T>::UNROLL - 1) / \
                                     ((i64)LBFGSB_BLOCK * Real<T>::VEC * Real<T>::UNROLL);       \
         _tl < _nt; _tl += LBFGSB_GRID)                                                        \
        _Pragma("unroll") for (int _k = 0; _k < Real<T>::UNROLL; ++_k)                           \
            for (i64 base = _tl * ((i64)LBFGSB_BLOCK * Real<T>::VEC * Real<T>::UNROLL) +         \
                            (i64)_k * (LBFGSB_BLOCK * Real<T>::VEC) + (i64)threadIdx.x * Real<T>::VEC, \
                     _once = 1;                                                                \
                 _once && base < (n); _once = 0)

// Same walk without unrolling the k loop (kernels that stream 2*col S/Y columns already have
// enough loads in flight per step; unrolling would only multiply register pressure).
#define LB_FOR_TILES_NU(T, n, base)                                                            \
    for (i64 _tl = blockIdx.x, _nt = ((n) + (i64)LBFGSB_BLOCK * Real<T>::VEC * Real<T>::UNROLL - 1) / \
                                     ((i64)LBFGSB_BLOCK * Real<T>::VEC * Real<T>::UNROLL);       \
         _tl < _nt; _tl += LBFGSB_GRID)                                                        \
        _Pragma("unroll 1") for (int _k = 0; _k < Real<T>::UNROLL; ++_k)                         \
            for (i64 base = _tl * ((i64)LBFGSB_BLOCK * Real<T>::VEC * Real<T>::UNROLL) +         \
                            (i64)_k * (LBFGSB_BLOCK * Real<T>::VEC) + (i64)threadIdx.x * Real<T>::VEC, \
                     _once = 1;                                                                \
                 _once && base < (n); _once = 0)

// ---------------------------------------------------------------------------
// Small dense routines (single thread).  Same operation order as the reference
// so that identical inputs give identical outputs (the engine is compiled with
// -fmad=false).
// ---------------------------------------------------------------------------
namespace dense {

// ddot, strict left-to-right (lbfgsb_blas_module.F90:187-202)
template <typename T> __device__ inline T ddot(int n, const T* dx, const T* dy) {
    T s = (T)0;
    for (int i = 0; i < n; ++i) s = s + dx[i] * dy[i];
    return s;
}
template <typename T> __device__ inline void daxpy(int n, T da, const T* dx, T* dy) {
    if (n <= 0 || da == (T)0) return;  // :49-50
    for (int i = 0; i < n; ++i) dy[i] = dy[i] + da * dx[i];
}

// dpofa (lbfgsb_linpack_module.f90:30-67); a(lda,*) column-major, 0-based storage
template <typename T> __device__ inline int dpofa(T* a, int lda, int n) {
    for (int j = 0; j < n; ++j) {
        T s = (T)0;
        for (int k = 0; k < j; ++k) {
            T t = a[k + j * lda] - ddot<T>(k, a + k * lda, a + j * lda);
            t = t / a[k + k * lda];
            a[k + j * lda] = t;
            s = s + t * t;
        }
        s = a[j + j * lda] - s;
        if (s <= (T)0) return j + 1;
        a[j + j * lda] = sqrt(s);
    }
    return 0;
}

// dtrsl (lbfgsb_linpack_module.f90:87-165); only the upper-triangular jobs the
// reference uses: job 01 (t*x=b) and job 11 (trans(t)*x=b).
template <typename T> __device__ inline int dtrsl(const T* t, int ldt, int n, T* b, int job) {
    for (int k = 0; k < n; ++k)
        if (t[k + k * ldt] == (T)0) return k + 1;
    if (job == 1) {  // case 2: t*x=b, t upper (:136-146)
        b[n - 1] = b[n - 1] / t[(n - 1) + (n - 1) * ldt];
        for (int jj = 2; jj <= n; ++jj) {
            int j = n - jj;  // 0-based
            T temp = -b[j + 1];
            daxpy<T>(j + 1, temp, t + (j + 1) * ldt, b);
            b[j] = b[j] / t[j + j * ldt];
        }
    } else {  // job 11, case 4: trans(t)*x=b, t upper (:159-166)
        b[0] = b[0] / t[0];
        for (int j = 1; j < n; ++j) {
            b[j] = b[j] - ddot<T>(j, t + j * ldt, b);
            b[j] = b[j] / t[j + j * ldt];
        }
    }
    return 0;
}

// bmv (src/lbfgsb.f90:1057-1123)
template <typename T>
__device__ inline int bmv(int m, const T* sy, const T* wt, int col, const T* v, T* p) {
    if (col == 0) return 0;
    p[col] = v[col];
    for (int i = 2; i <= col; ++i) {
        int i2 = col + i;
        T sum = (T)0;
        for (int k = 1; k <= i - 1; ++k)
            sum = sum + sy[(i - 1) + (k - 1) * m] * v[k - 1] / sy[(k - 1) + (k - 1) * m];
        p[i2 - 1] = v[i2 - 1] + sum;
    }
    int info = dtrsl<T>(wt, m, col, p + col, 11);
    if (info != 0) return info;
    for (int i = 0; i < col; ++i) p[i] = v[i] / sqrt(sy[i + i * m]);
    info = dtrsl<T>(wt, m, col, p + col, 1);
    if (info != 0) return info;
    for (int i = 0; i < col; ++i) p[i] = -p[i] / sqrt(sy[i + i * m]);
    for (int i = 1; i <= col; ++i) {
        T sum = (T)0;
        for (int k = i + 1; k <= col; ++k)
            sum = sum + sy[(k - 1) + (i - 1) * m] * p[col + k - 1] / sy[(i - 1) + (i - 1) * m];
        p[i - 1] = p[i - 1] + sum;
    }
    return 0;
}

// formt (src/lbfgsb.f90:1926-1963)
template <typename T>
__device__ inline int formt(int m, T* wt, const T* sy, const T* ss, int col, T theta) {
    for (int j = 1; j <= col; ++j) wt[0 + (j - 1) * m] = theta * ss[0 + (j - 1) * m];
    for (int i = 2; i <= col; ++i)
        for (int j = i; j <= col; ++j) {
            int k1 = (i < j ? i : j) - 1;
            T ddum = (T)0;
            for (int k = 1; k <= k1; ++k)
                ddum = ddum + sy[(i - 1) + (k - 1) * m] * sy[(j - 1) + (k - 1) * m] / sy[(k - 1) + (k - 1) * m];
            wt[(i - 1) + (j - 1) * m] = ddum + theta * ss[(i - 1) + (j - 1) * m];
        }
    int info = dpofa<T>(wt, m, col);
    return info != 0 ? -3 : 0;
}

template <typename T> __device__ inline T tmax(T a, T b) { return a > b ? a : b; }  // Fortran max
template <typename T> __device__ inline T tmin(T a, T b) { return a < b ? a : b; }

// dcstep (src/lbfgsb.f90:3227-3415)
template <typename T>
__device__ inline void dcstep(T& stx, T& fx, T& dx, T& sty, T& fy, T& dy, T& stp, T fp, T dp,
                              bool& brackt, T stpmin, T stpmax) {
    const T zero = (T)0, two = (T)2, three = (T)3, p66 = (T)0.66;
    T gamma, p, q, r, s, sgnd, stpc, stpf, stpq, theta;
    sgnd = dp * (dx / fabs(dx));
    if (fp > fx) {
        theta = three * (fx - fp) / (stp - stx) + dx + dp;
        s = tmax(tmax(fabs(theta), fabs(dx)), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / two) * (stp - stx);
        if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / two;
        brackt = true;
    } else if (sgnd < zero) {
        theta = three * (fx - fp) / (stp - stx) + dx + dp;
        s = tmax(tmax(fabs(theta), fabs(dx)), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
        else stpf = stpq;
        brackt = true;
    } else if (fabs(dp) < fabs(dx)) {
        theta = three * (fx - fp) / (stp - stx) + dx + dp;
        s = tmax(tmax(fabs(theta), fabs(dx)), fabs(dp));
        gamma = s * sqrt(tmax(zero, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < zero && gamma != zero) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            if (stp > stx) stpf = tmin(stp + p66 * (sty - stp), stpf);
            else stpf = tmax(stp + p66 * (sty - stp), stpf);
        } else {
            if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            stpf = tmin(stpmax, stpf);
            stpf = tmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            theta = three * (fp - fy) / (sty - stp) + dy + dp;
            s = tmax(tmax(fabs(theta), fabs(dy)), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) stpf = stpmax;
        else stpf = stpmin;
    }
    if (fp > fx) {
        sty = stp; fy = fp; dy = dp;
    } else {
        if (sgnd < zero) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

// dcsrch (src/lbfgsb.f90:2942-3198); task is an integer code, state in (brackt, stage, ls[13]).
template <typename T>
__device__ inline void dcsrch(T f, T g, T& stp, T ftol, T gtol, T xtol, T stpmin, T stpmax,
                              int& task, int& ibrackt, int& stage, T* ds) {
    const T zero = (T)0, p5 = (T)0.5, p66 = (T)0.66, xtrapl = (T)1.1, xtrapu = (T)4.0;
    bool brackt;
    T finit, ftest, fm, fx, fxm, fy, fym, ginit, gtest, gm, gx, gxm, gy, gym, stx, sty, stmin, stmax,
        width, width1;
    if (task == CS_START) {
        if (stp < stpmin) task = CS_ERR_STP_LT_MIN;
        if (stp > stpmax) task = CS_ERR_STP_GT_MAX;
        if (g >= zero) task = CS_ERR_G_GE_0;
        if (ftol < zero) task = CS_ERR_FTOL;
        if (gtol < zero) task = CS_ERR_GTOL;
        if (xtol < zero) task = CS_ERR_XTOL;
        if (stpmin < zero) task = CS_ERR_STPMIN;
        if (stpmax < stpmin) task = CS_ERR_STPMAX;
        if (cs_is_err(task)) return;
        brackt = false; stage = 1; finit = f; ginit = g; gtest = ftol * ginit;
        width = stpmax - stpmin; width1 = width / p5;
        stx = zero; fx = finit; gx = ginit; sty = zero; fy = finit; gy = ginit;
        stmin = zero; stmax = stp + xtrapu * stp;
        task = CS_FG;
        goto save;
    } else {
        brackt = (ibrackt == 1);
        ginit = ds[0]; gtest = ds[1]; gx = ds[2]; gy = ds[3]; finit = ds[4]; fx = ds[5]; fy = ds[6];
        stx = ds[7]; sty = ds[8]; stmin = ds[9]; stmax = ds[10]; width = ds[11]; width1 = ds[12];
    }
    ftest = finit + stp * gtest;
    if (stage == 1 && f <= ftest && g >= zero) stage = 2;
    if (brackt && (stp <= stmin || stp >= stmax)) task = CS_WARN_ROUND;
    if (brackt && stmax - stmin <= xtol * stmax) task = CS_WARN_XTOL;
    if (stp == stpmax && f <= ftest && g <= gtest) task = CS_WARN_STPMAX;
    if (stp == stpmin && (f > ftest || g >= gtest)) task = CS_WARN_STPMIN;
    if (f <= ftest && fabs(g) <= gtol * (-ginit)) task = CS_CONV;
    if (cs_is_warn(task) || task == CS_CONV) goto save;
    if (stage == 1 && f <= fx && f > ftest) {
        fm = f - stp * gtest; fxm = fx - stx * gtest; fym = fy - sty * gtest;
        gm = g - gtest; gxm = gx - gtest; gym = gy - gtest;
        dcstep<T>(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
        fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
    } else {
        dcstep<T>(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax);
    }
    if (brackt) {
        if (fabs(sty - stx) >= p66 * width1) stp = stx + p5 * (sty - stx);
        width1 = width;
        width = fabs(sty - stx);
    }
    if (brackt) { stmin = tmin(stx, sty); stmax = tmax(stx, sty); }
    else { stmin = stp + xtrapl * (stp - stx); stmax = stp + xtrapu * (stp - stx); }
    stp = tmax(stp, stpmin);
    stp = tmin(stp, stpmax);
    if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
    task = CS_FG;
save:
    ibrackt = brackt ? 1 : 0;
    ds[0] = ginit; ds[1] = gtest; ds[2] = gx; ds[3] = gy; ds[4] = finit; ds[5] = fx; ds[6] = fy;
    ds[7] = stx; ds[8] = sty; ds[9] = stmin; ds[10] = stmax; ds[11] = width; ds[12] = width1;
}

}  // namespace dense

// ---------------------------------------------------------------------------
// The same small dense routines run by ONE WARP (all 32 lanes call; `lane` = threadIdx.x & 31).  Every matrix
// or vector entry is formed by exactly the operations of the single-thread versions above, in the same order
// -- only entries that do not depend on one another are formed by different lanes at the same time -- so the
// results are bit-identical (tests/test_gpu_kernels.py compares both with the oracle).  What this buys is
// latency: the chains of dependent FP divisions and the shared-memory round trips of the 2m x 2m algebra are
// what the scalar kernels spend their time on.
// ---------------------------------------------------------------------------
namespace wdense {

#define LB_FULL 0xffffffffu

// dpofa, row by row: at step k the diagonal entry of column k (its sum of squares runs over rows 0..k-1 in order,
// as LINPACK accumulates it), then row k of every column j > k, one lane per column.
template <typename T> __device__ inline int dpofa(T* a, int lda, int n) {
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < n; ++k) {
        int bad = 0;
        if (lane == 0) {
            T s = (T)0;
            for (int q = 0; q < k; ++q) { const T t = a[q + k * lda]; s = s + t * t; }
            s = a[k + k * lda] - s;
            if (s <= (T)0) bad = 1;
            else a[k + k * lda] = sqrt(s);
        }
        bad = __shfl_sync(LB_FULL, bad, 0);
        if (bad) return k + 1;
        __syncwarp();
        const T dk = a[k + k * lda];
        for (int j = k + 1 + lane; j < n; j += 32) {
            T t = a[k + j * lda] - dense::ddot<T>(k, a + k * lda, a + j * lda);
            t = t / dk;
            a[k + j * lda] = t;
        }
        __syncwarp();
    }
    return 0;
}

// dtrsl, jobs 01 and 11, n <= 64: lane q owns entries q and q + 32 of b.
template <typename T> __device__ inline int dtrsl(const T* t, int ldt, int n, T* b, int job) {
    const int lane = threadIdx.x & 31;
    unsigned z0 = __ballot_sync(LB_FULL, lane < n && t[lane + lane * ldt] == (T)0);
    unsigned z1 = __ballot_sync(LB_FULL, lane + 32 < n && t[(lane + 32) + (lane + 32) * ldt] == (T)0);
    if (z0) return __ffs(z0);
    if (z1) return 32 + __ffs(z1);
    T b0 = lane < n ? b[lane] : (T)0, b1 = lane + 32 < n ? b[lane + 32] : (T)0;
    if (job == 1) {   // t*x = b, t upper: b(j) final from the last row up; then b(i) += (-b(j)) t(i,j) for i < j (:136-146)
        for (int j = n - 1; j >= 0; --j) {
            const int own = j & 31;
            T bj = (j < 32) ? b0 : b1;
            if (lane == own) { bj = bj / t[j + j * ldt]; if (j < 32) b0 = bj; else b1 = bj; }
            bj = __shfl_sync(LB_FULL, bj, own);
            const T temp = -bj;
            if (temp != (T)0) {   // daxpy's early-out (:49-50)
                if (lane < j) b0 = b0 + temp * t[lane + j * ldt];
                if (lane + 32 < j) b1 = b1 + temp * t[(lane + 32) + j * ldt];
            }
        }
    } else {          // trans(t)*x = b, t upper: b(j) = (b(j) - sum_{i<j} t(i,j) b(i)) / t(j,j), the sum in i order (:159-166)
        T s0 = (T)0, s1 = (T)0;
        for (int i = 0; i < n; ++i) {
            const int own = i & 31;
            T bi = (i < 32) ? b0 : b1;
            if (lane == own) {
                if (i > 0) bi = bi - ((i < 32) ? s0 : s1);
                bi = bi / t[i + i * ldt];
                if (i < 32) b0 = bi; else b1 = bi;
            }
            bi = __shfl_sync(LB_FULL, bi, own);
            if (lane > i && lane < n) s0 = s0 + t[i + lane * ldt] * bi;
            if (lane + 32 > i && lane + 32 < n) s1 = s1 + t[i + (lane + 32) * ldt] * bi;
        }
    }
    if (lane < n) b[lane] = b0;
    if (lane + 32 < n) b[lane + 32] = b1;
    __syncwarp();
    return 0;
}

// bmv (src/lbfgsb.f90:1057-1123); v and p must not overlap
template <typename T>
__device__ inline int bmv(int m, const T* sy, const T* wt, int col, const T* v, T* p) {
    const int lane = threadIdx.x & 31;
    if (col == 0) return 0;
    for (int i = 1 + lane; i <= col; i += 32) {
        if (i == 1) { p[col] = v[col]; continue; }
        T sum = (T)0;
        for (int k = 1; k <= i - 1; ++k)
            sum = sum + sy[(i - 1) + (k - 1) * m] * v[k - 1] / sy[(k - 1) + (k - 1) * m];
        p[col + i - 1] = v[col + i - 1] + sum;
    }
    __syncwarp();
    int info = dtrsl<T>(wt, m, col, p + col, 11);
    if (info != 0) return info;
    for (int i = lane; i < col; i += 32) p[i] = v[i] / sqrt(sy[i + i * m]);
    __syncwarp();
    info = dtrsl<T>(wt, m, col, p + col, 1);
    if (info != 0) return info;
    for (int i = 1 + lane; i <= col; i += 32) {
        T pi = -p[i - 1] / sqrt(sy[(i - 1) + (i - 1) * m]);
        T sum = (T)0;
        for (int k = i + 1; k <= col; ++k)
            sum = sum + sy[(k - 1) + (i - 1) * m] * p[col + k - 1] / sy[(i - 1) + (i - 1) * m];
        p[i - 1] = pi + sum;
    }
    __syncwarp();
    return 0;
}

// formt (src/lbfgsb.f90:1926-1963): the entries of T one per lane, then dpofa
template <typename T>
__device__ inline int formt(int m, T* wt, const T* sy, const T* ss, int col, T theta) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < col * col; e += 32) {
        const int i = e % col + 1, j = e / col + 1;
        if (i > j) continue;
        if (i == 1) { wt[0 + (j - 1) * m] = theta * ss[0 + (j - 1) * m]; continue; }
        const int k1 = i - 1;   // min(i, j) - 1 with i <= j
        T ddum = (T)0;
        for (int k = 1; k <= k1; ++k)
            ddum = ddum + sy[(i - 1) + (k - 1) * m] * sy[(j - 1) + (k - 1) * m] / sy[(k - 1) + (k - 1) * m];
        wt[(i - 1) + (j - 1) * m] = ddum + theta * ss[(i - 1) + (j - 1) * m];
    }
    __syncwarp();
    int info = dpofa<T>(wt, m, col);
    return info != 0 ? -3 : 0;
}

}  // namespace wdense
