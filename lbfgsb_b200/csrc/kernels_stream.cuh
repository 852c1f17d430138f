// Streaming (HBM-bound) kernels of one L-BFGS-B iteration.  One pass over the
// variables each, 128-bit loads, fixed-shape block reductions (common.cuh), no
// floating-point atomics.  Every kernel is launched with LBFGSB_GRID blocks of
// LBFGSB_BLOCK threads and predicates itself on the device state block, so the
// host can enqueue a whole setulb call without reading anything back.
//
// Each kernel names the reference loop(s) it replaces (src/lbfgsb.f90).
#pragma once
#include "common.cuh"

template <typename T>
struct Wk {               // workspace + user vectors of one problem (device pointers)
    i64 n, ldw;           // variables; leading dimension of ws/wy (multiple of 32)
    i64 off;              // global index of this shard's first variable (0 on a single GPU)
    int m;
    T *ws, *wy;           // S, Y histories, column-major [m][ldw]        (:390-391)
    T *z, *r, *d, *t, *xp;  // n-vectors of mainlb                         (:382-388)
    T* gold;              // previous gradient (the reference keeps it in r between lnsrlb :2236 and :814; here r stays
                          // the reduced gradient / Newton direction of subsm so that the two can live in one pass)
    int* iwhere;          // (:348-355)
    unsigned char* state; // bit0: free at the GCP (freev :2047); bit1: free at the previous freev
    T* part;              // [LB_KMAX][GRID] block partials
    i64* ipart;           // [LB_IMAX][GRID]
    T* part2;             // second set of partials: a fused pass feeds two reduction sites
    i64* ipart2;          //   (cauchy's site and the W'Zr site always live here)
    DevState<T>* s;
    T* x; const T* l; const T* u; const int* nbd; T* g;   // caller's vectors
    int bp_hint;          // cauchy's per-variable pass also stores every breakpoint (in r) and xcp = x (in z): the previous
                          // iteration needed a breakpoint walk, so this one probably does (cauchy_walk.cuh)
};

#define LB_SLOT(part, k) ((part) + (i64)(k) * LBFGSB_GRID)
#define LB_INF(T) ((T)INFINITY)
#define LB_I64MAX 0x7fffffffffffffffLL

// ---------------------------------------------------------------------------
// errclb (:1601-1643): last offending index of each error kind.
// ipart slot 0: max index with invalid nbd (-1 none); slot 1: max index with l>u.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_errclb(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    i64 k6 = -1, k7 = -1;
    LB_FOR_TILES(T, w.n, base) {
        T l[VEC], u[VEC]; int nb[VEC];
        ldv<T>(w.l, base, w.n, l); ldv<T>(w.u, base, w.n, u); ldvi<T>(w.nbd, base, w.n, nb);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            i64 i = base + v;
            if (i < w.n) {
                if (nb[v] < 0 || nb[v] > 3) k6 = i;
                if (nb[v] == 2 && l[v] > u[v]) k7 = i;
            }
        }
    }
    i64 m6 = k6, m7 = k7;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        i64 o6 = shfl_xor_t<i64>(m6, off), o7 = shfl_xor_t<i64>(m7, off);
        m6 = o6 > m6 ? o6 : m6; m7 = o7 > m7 ? o7 : m7;
    }
    __shared__ i64 s6[LBFGSB_BLOCK / 32], s7[LBFGSB_BLOCK / 32];
    if ((threadIdx.x & 31) == 0) { s6[threadIdx.x >> 5] = m6; s7[threadIdx.x >> 5] = m7; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < LBFGSB_BLOCK / 32; ++q) { m6 = s6[q] > m6 ? s6[q] : m6; m7 = s7[q] > m7 ? s7[q] : m7; }
        LB_SLOT(w.ipart, 0)[blockIdx.x] = m6;
        LB_SLOT(w.ipart, 1)[blockIdx.x] = m7;
    }
}

// ---------------------------------------------------------------------------
// active (:965-1040): project x into the box, initialise iwhere, flags.
// ipart: 0 nbdd (sum), 1 prjctd (sum>0), 2 cnstnd (sum>0), 3 not-boxed (sum>0)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_active(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    if (!w.s->go) return;
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    i64 nbdd = 0, prj = 0, cns = 0, nbx = 0;
    LB_FOR_TILES(T, w.n, base) {
        T x[VEC], l[VEC], u[VEC]; int nb[VEC], iw[VEC], st[VEC];
        ldv<T>(w.x, base, w.n, x); ldv<T>(w.l, base, w.n, l); ldv<T>(w.u, base, w.n, u);
        ldvi<T>(w.nbd, base, w.n, nb);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            iw[v] = 0; st[v] = 3;
            if (base + v < w.n) {
                if (nb[v] > 0) {
                    if (nb[v] <= 2 && x[v] <= l[v]) {
                        if (x[v] < l[v]) { prj = 1; x[v] = l[v]; }
                        nbdd++;
                    } else if (nb[v] >= 2 && x[v] >= u[v]) {
                        if (x[v] > u[v]) { prj = 1; x[v] = u[v]; }
                        nbdd++;
                    }
                }
                if (nb[v] != 2) nbx = 1;
                if (nb[v] == 0) iw[v] = -1;
                else {
                    cns = 1;
                    iw[v] = (nb[v] == 2 && u[v] - l[v] <= (T)0) ? 3 : 0;
                }
            }
        }
        stv<T>(w.x, base, w.n, x);
        stvi<T>(w.iwhere, base, w.n, iw);
        stvb<T>(w.state, base, w.n, st);
    }
    i64 r0 = block_isum(nbdd, smi), r1 = block_isum(prj, smi), r2 = block_isum(cns, smi), r3 = block_isum(nbx, smi);
    if (threadIdx.x == 0) {
        LB_SLOT(w.ipart, 0)[blockIdx.x] = r0; LB_SLOT(w.ipart, 1)[blockIdx.x] = r1;
        LB_SLOT(w.ipart, 2)[blockIdx.x] = r2; LB_SLOT(w.ipart, 3)[blockIdx.x] = r3;
    }
}

// projected gradient of one variable (projgr :2611-2619)
template <typename T>
__device__ __forceinline__ T projg_one(T x, T g, T l, T u, int nb) {
    T gi = g;
    if (nb != 0) {
        if (gi < (T)0) { if (nb >= 2) gi = dense::tmax(x - u, gi); }
        else           { if (nb <= 2) gi = dense::tmin(x - l, gi); }
    }
    return fabs(gi);
}

// ---------------------------------------------------------------------------
// projgr (:2594-2622).  part slot 0: max |proj g|.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_projgr(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    if (!w.s->go) return;
    __shared__ T sm[LBFGSB_BLOCK / 32];
    T acc = (T)0;
    LB_FOR_TILES(T, w.n, base) {
        T x[VEC], l[VEC], u[VEC], g[VEC]; int nb[VEC];
        ldv<T>(w.x, base, w.n, x); ldv<T>(w.g, base, w.n, g); ldv<T>(w.l, base, w.n, l);
        ldv<T>(w.u, base, w.n, u); ldvi<T>(w.nbd, base, w.n, nb);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (base + v < w.n) acc = dense::tmax(acc, projg_one<T>(x[v], g[v], l[v], u[v], nb[v]));
    }
    T r = block_max<T>(acc, sm);
    if (threadIdx.x == 0) LB_SLOT(w.part, 0)[blockIdx.x] = r;
}

// ---------------------------------------------------------------------------
// cauchy tail (:1515) fused with freev (:1980-2059): xcp += tsum*d, then count
// free / entering / leaving variables and refresh the per-variable state byte.
// The reference's index lists are not materialised: every later "over the free
// set" loop is a masked pass over the variables.
// ipart: 0 nfree ; 1 nenter ; 2 nleave
// ---------------------------------------------------------------------------
// Cauchy direction of one variable from its class (cauchy :1294-1298): -g if the variable moves, else 0.
template <typename T>
__device__ __forceinline__ T cauchy_dir(int iw, T g) { return (iw == 0 || iw == -1) ? -g : (T)0; }

// (cauchy's d and xcp = x are written out only in front of a breakpoint walk: k_bp_count<T, true>, cauchy_walk.cuh)
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_gcp_freev(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || s->cauchy_mode == 1 || s->fuse_gf) return;
    const i64 n = w.n;
    const T tsum = s->tsum;
    const bool axpy = (s->cauchy_mode == 0) && (tsum != (T)0);   // daxpy early-out (:49-50)
    const bool lazy = (s->cauchy_mode == 0) && s->lazy_gcp;      // d, xcp not materialised: xcp = x + tsum*d
    const bool cnt = (s->iter > 0 && s->cnstnd);
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    i64 nfr = 0, nen = 0, nle = 0;
    LB_FOR_TILES(T, n, base) {
        int iw[VEC], st[VEC];
        ldvi<T>(w.iwhere, base, n, iw); ldvb<T>(w.state, base, n, st);
        if (lazy) {
            T z[VEC];
            ldv<T>(w.x, base, n, z);
            if (axpy) {
                T g[VEC];
                ldv<T>(w.g, base, n, g);
#pragma unroll
                for (int v = 0; v < VEC; ++v) z[v] = z[v] + tsum * cauchy_dir<T>(iw[v], g[v]);
            }
            stv<T>(w.z, base, n, z);
        } else if (axpy) {   // after a breakpoint walk: xcp holds the bounds of the fixed variables, whose iwhere is 1 or 2 (d = 0)
            T g[VEC], z[VEC];
            ldv<T>(w.g, base, n, g); ldv<T>(w.z, base, n, z);
#pragma unroll
            for (int v = 0; v < VEC; ++v) z[v] = z[v] + tsum * cauchy_dir<T>(iw[v], g[v]);
            stv<T>(w.z, base, n, z);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (base + v < n) {
                const int fr = iw[v] <= 0 ? 1 : 0;
                const int old = st[v] & 1;
                nfr += fr;
                if (cnt) { nen += (fr && !old); nle += (!fr && old); }
                st[v] = fr | ((cnt ? old : fr) << 1);
            }
        }
        stvb<T>(w.state, base, n, st);
    }
    i64 r0 = block_isum(nfr, smi), r1 = block_isum(nen, smi), r2 = block_isum(nle, smi);
    if (threadIdx.x == 0) {
        LB_SLOT(w.ipart, 0)[blockIdx.x] = r0; LB_SLOT(w.ipart, 1)[blockIdx.x] = r1;
        LB_SLOT(w.ipart, 2)[blockIdx.x] = r2;
    }
}

// ---------------------------------------------------------------------------
// subsm backtrack (:2836-2863): largest feasible step along the Newton direction
// from xp, first binding variable.  part 0: alpha candidate; ipart 0: its variable.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_bt_alpha(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->do_backtrack) return;
    const i64 n = w.n;
    __shared__ T smv[LBFGSB_BLOCK / 32];
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    T best = LB_INF(T); i64 ib = LB_I64MAX;
    const T* xcp = s->spec_step ? w.z : w.xp;   // the speculative subspace pass leaves xcp in z
    LB_FOR_TILES(T, n, base) {
        int st[VEC], nb[VEC]; T dk[VEC], xk[VEC], l[VEC], u[VEC];
        ldvb<T>(w.state, base, n, st); ldvi<T>(w.nbd, base, n, nb);
        ldv<T>(w.r, base, n, dk); ldv<T>(xcp, base, n, xk); ldv<T>(w.l, base, n, l); ldv<T>(w.u, base, n, u);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (base + v < n && (st[v] & 1) && nb[v] != 0) {
                T cand = LB_INF(T);
                if (dk[v] < (T)0 && nb[v] <= 2) {
                    T t2 = l[v] - xk[v];
                    cand = (t2 >= (T)0) ? (T)0 : t2 / dk[v];
                } else if (dk[v] > (T)0 && nb[v] >= 2) {
                    T t2 = u[v] - xk[v];
                    cand = (t2 <= (T)0) ? (T)0 : t2 / dk[v];
                }
                if (cand < best) { best = cand; ib = base + v + w.off; }   // global index: ties resolve identically on every rank
            }
        }
    }
    block_argmin<T>(best, ib, smv, smi);
    if (threadIdx.x == 0) { LB_SLOT(w.part, 0)[blockIdx.x] = best; LB_SLOT(w.ipart, 0)[blockIdx.x] = ib; }
}

// subsm backtrack apply (:2830, :2865-2879): x = xp + alpha d on the free set,
// the binding variable pinned to its bound.
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_bt_apply(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->do_backtrack) return;
    const i64 n = w.n;
    const T alpha = s->alpha;
    const i64 ibd = s->ibd;
    const T* xcp = s->spec_step ? w.z : w.xp;
    LB_FOR_TILES(T, n, base) {
        int st[VEC]; T dk[VEC], xk[VEC], l[VEC], u[VEC];
        ldvb<T>(w.state, base, n, st);
        ldv<T>(w.r, base, n, dk); ldv<T>(xcp, base, n, xk); ldv<T>(w.l, base, n, l); ldv<T>(w.u, base, n, u);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (base + v < n) {
                if (st[v] & 1) {
                    if (alpha < (T)1 && base + v + w.off == ibd) {
                        if (dk[v] > (T)0) { xk[v] = u[v]; dk[v] = (T)0; }
                        else if (dk[v] < (T)0) { xk[v] = l[v]; dk[v] = (T)0; }
                    }
                    xk[v] = xk[v] + alpha * dk[v];
                }
            }
        }
        stv<T>(w.z, base, n, xk);
    }
}

// ---------------------------------------------------------------------------
// d = z - x (:720-722) fused with the first entry of lnsrlb (:2196-2244):
// dtd, stpmx candidates, t = x, gold = g, gd = g.d
// part2: 0 dtd ; 1 gd ; 2 stpmx candidate (min)
// Skipped when k_subsm_lsinit already did this work in the subspace pass (lsinit_done).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_ls_init(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || s->lsinit_done) return;
    const i64 n = w.n;
    const bool bounds = (s->cnstnd && s->iter != 0);
    __shared__ T sm[2 * (LBFGSB_BLOCK / 32)];
    __shared__ T smm[LBFGSB_BLOCK / 32];
    T acc[2]; acc[0] = (T)0; acc[1] = (T)0;
    T smx = LB_INF(T);
    const T* xs = s->spec_step ? w.t : w.x;   // after a speculative step the iterate is in t
    const bool lz = s->lazy_z != 0;           // xcp not stored (fuse_gf without a subspace step): form and store it here
    const T tsum = s->tsum;
    const bool axpy = tsum != (T)0;
    LB_FOR_TILES(T, n, base) {
        T z[VEC], x[VEC], g[VEC], d[VEC];
        ldv<T>(xs, base, n, x); ldv<T>(w.g, base, n, g);
        if (lz) {
            int st[VEC];
            ldvb<T>(w.state, base, n, st);
#pragma unroll
            for (int v = 0; v < VEC; ++v) z[v] = axpy ? (x[v] + tsum * ((st[v] & 4) ? -g[v] : (T)0)) : x[v];
            stv<T>(w.z, base, n, z);
        } else ldv<T>(w.z, base, n, z);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            d[v] = z[v] - x[v];
            if (base + v < n) { acc[0] = acc[0] + d[v] * d[v]; acc[1] = acc[1] + g[v] * d[v]; }
        }
        stv<T>(w.d, base, n, d); stv<T>(w.t, base, n, x); stv<T>(w.gold, base, n, g);
        if (bounds) {
            T l[VEC], u[VEC]; int nb[VEC];
            ldv<T>(w.l, base, n, l); ldv<T>(w.u, base, n, u); ldvi<T>(w.nbd, base, n, nb);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (base + v < n && nb[v] != 0) {
                    const T a1 = d[v];
                    if (a1 < (T)0 && nb[v] <= 2) {
                        const T a2 = l[v] - x[v];
                        const T cand = (a2 >= (T)0) ? (T)0 : a2 / a1;
                        smx = dense::tmin(smx, cand);
                    } else if (a1 > (T)0 && nb[v] >= 2) {
                        const T a2 = u[v] - x[v];
                        const T cand = (a2 <= (T)0) ? (T)0 : a2 / a1;
                        smx = dense::tmin(smx, cand);
                    }
                }
            }
        }
    }
    block_sum_store<T, 2>(acc, 2, sm, w.part2);
    T r = block_min<T>(smx, smm);
    if (threadIdx.x == 0) LB_SLOT(w.part2, 2)[blockIdx.x] = r;
}

// ---------------------------------------------------------------------------
// lnsrlb trial point (:2264-2270): x = z (stp == 1) or x = stp*d + t.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_ls_step(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    const DevState<T>* s = w.s;
    const i64 n = w.n;
    if (s->do_unstep) {   // the speculative step is withdrawn (line search failed at its first entry)
        LB_FOR_TILES(T, n, base) { T t[VEC]; ldv<T>(w.t, base, n, t); stv<T>(w.x, base, n, t); }
        return;
    }
    if (!s->do_step || s->step_done) return;
    const T stp = s->stp;
    if (stp == (T)1) {
        LB_FOR_TILES(T, n, base) { T z[VEC]; ldv<T>(w.z, base, n, z); stv<T>(w.x, base, n, z); }
    } else {
        const bool keep = s->save_z != 0;   // x is the only copy of the Newton point z: store it before x moves on
        LB_FOR_TILES(T, n, base) {
            T d[VEC], t[VEC], x[VEC];
            if (keep) { ldv<T>(w.x, base, n, x); stv<T>(w.z, base, n, x); }
            ldv<T>(w.d, base, n, d); ldv<T>(w.t, base, n, t);
#pragma unroll
            for (int v = 0; v < VEC; ++v) x[v] = stp * d[v] + t[v];
            stv<T>(w.x, base, n, x);
        }
    }
}

// restore the previous iterate (:736-738, :568-570): x = t, g = r
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_restore(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    if (!w.s->do_restore) return;
    const i64 n = w.n;
    LB_FOR_TILES(T, n, base) {
        T t[VEC], r[VEC];
        ldv<T>(w.t, base, n, t); ldv<T>(w.gold, base, n, r);
        stv<T>(w.x, base, n, t); stv<T>(w.g, base, n, r);
    }
}

// ---------------------------------------------------------------------------
// lnsrlb re-entry (:2244): gd = g.d at the trial point, fused with a
// speculative projgr (:781) that is used if the line search accepts the point.
// part: 0 gd ; 1 max |proj g|
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_ls_trial(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    // (launched only on the FG_LNSRCH entry, in front of the scalar kernel that opens the call: no predicate)
    const i64 n = w.n;
    __shared__ T sm[LBFGSB_BLOCK / 32];
    __shared__ T smm[LBFGSB_BLOCK / 32];
    T acc[1]; acc[0] = (T)0;
    T pg = (T)0;
    LB_FOR_TILES(T, n, base) {
        T x[VEC], l[VEC], u[VEC], g[VEC], d[VEC]; int nb[VEC];
        ldv<T>(w.g, base, n, g); ldv<T>(w.d, base, n, d); ldv<T>(w.x, base, n, x);
        ldv<T>(w.l, base, n, l); ldv<T>(w.u, base, n, u); ldvi<T>(w.nbd, base, n, nb);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (base + v < n) {
                acc[0] = acc[0] + g[v] * d[v];
                pg = dense::tmax(pg, projg_one<T>(x[v], g[v], l[v], u[v], nb[v]));
            }
    }
    block_sum_store<T, 1>(acc, 1, sm, w.part);
    T r = block_max<T>(pg, smm);
    if (threadIdx.x == 0) LB_SLOT(w.part, 1)[blockIdx.x] = r;
}

// ---------------------------------------------------------------------------
// Identity of the active set {i : iwhere(i) > 0}: sum of splitmix64(i) mod 2^64
// (order independent, integer).  ipart 0: hash, 1: count.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long lb_splitmix64(unsigned long long v) {
    v += 0x9E3779B97F4A7C15ULL;
    v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ULL;
    v = (v ^ (v >> 27)) * 0x94D049BB133111EBULL;
    return v ^ (v >> 31);
}
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_active_hash(Wk<T> w, i64 index_offset) {
    constexpr int VEC = Real<T>::VEC;
    const i64 n = w.n;
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    i64 h = 0, c = 0;
    LB_FOR_TILES(T, n, base) {
        int iw[VEC];
        ldvi<T>(w.iwhere, base, n, iw);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (base + v < n && iw[v] > 0) { h += (i64)lb_splitmix64((unsigned long long)(base + v + index_offset)); c++; }
    }
    i64 r0 = block_isum(h, smi), r1 = block_isum(c, smi);
    if (threadIdx.x == 0) { LB_SLOT(w.ipart, 0)[blockIdx.x] = r0; LB_SLOT(w.ipart, 1)[blockIdx.x] = r1; }
}
