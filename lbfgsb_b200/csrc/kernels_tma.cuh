// The five passes of one L-BFGS-B iteration that stream the 2*col columns of the S/Y
// history, as TMA-staged kernels (tma_pipe.cuh): bulk copies land every input stream of a
// sub-tile in a two-stage shared-memory ring, the CTA computes from shared memory, outputs go
// straight to global memory with 128-bit stores.  One CTA per SM (the ring takes ~200 KB);
// LBFGSB_GRID = 4 x 148 CTAs run as four equal waves.  A CTA is 8 consumer warps (the 256
// threads of the fixed reduction shape) plus one producer warp (tma_pipe.cuh).
//
// Same reductions, same thread -> element mapping and same accumulation order as the plain
// streaming kernels (include/lbfgsb_b200_shape.h), so the oracle's device-order mode replays
// them.  Each kernel names the reference loops it replaces (src/lbfgsb.f90).
#pragma once
#include "kernels_stream.cuh"
#include "tma_pipe.cuh"

// Threads per staged sub-tile: all 256 whenever the 2*MT + 8 streams of a sub-tile (256 * VEC reals each) fit twice in
// shared memory -- REAL64 with m <= 10 and REAL32 with m <= 20 (VEC = 2 for both kinds, lbfgsb_b200_shape.h); REAL64 with
// m > 10 stages half sub-tiles and only the 128 owning threads work on a stage.
template <typename T, int MT> struct SubT { static constexpr int v = (sizeof(T) * MT <= 80) ? 256 : 128; };

// REAL32 passes are bound by instruction issue and shared-memory latency rather than by DRAM (two resident warps per
// scheduler, one staged load feeding a handful of single-precision operations): for them the steady state of the
// iteration -- history full, col == MT -- gets a second instantiation of the pass body in which the ring length is a
// compile-time constant, so that the 2*MT column steps are one branch-free block and the compiler overlaps the staged
// loads of later columns with the arithmetic of earlier ones.  Same operations in the same order on every accumulator:
// the bits do not change.  (Measured and rejected for REAL64, whose passes sit at the DRAM roofline: larger code, no
// fewer stalls.)
template <typename T> struct SpecializeFullHistory { static constexpr bool value = sizeof(T) == 4; };
template <bool B> struct BoolC { static constexpr bool value = B; };

// The stage ring starts at the (128-byte aligned) base of the dynamic shared memory.  It must stay a pointer
// derived from `dyn` (no integer round trip): the compiler then knows the address space and reads the
// stages with LDS instead of generic loads.
#define LB_DYN_STAGES(dyn) (dyn)

// add the ring columns head, head+1, ... (count pairs) as (Wy_j, Ws_j) stream pairs
template <typename T>
__device__ __forceinline__ void pipe_add_w(PipeSrc* ps, const Wk<T>& w, int head0, int count, unsigned slot) {
    int pj = head0;
    for (int j = 0; j < count; ++j) {
        pipe_add(ps, w.wy + (i64)pj * w.ldw, (int)sizeof(T), slot);
        pipe_add(ps, w.ws + (i64)pj * w.ldw, (int)sizeof(T), slot);
        pj = (pj + 1 == w.m) ? 0 : pj + 1;
    }
}

// ---------------------------------------------------------------------------
// y/s preparation (:813-824) fused with matupd (:2313-2338): y = g - r,
// s = stp*d written straight into the ring columns, rr = y.y, and the new row of
// S'Y / column of S'S against the col-1 older pairs -- one pass.
// part: 0 rr ; [1,1+MT) sum s*Wy(:,j) ; [1+MT,1+2MT) sum Ws(:,j)*s   (j = ring position, < col-1)
// ---------------------------------------------------------------------------
template <typename T, int MT>
__global__ void __launch_bounds__(LB_TMA_THREADS, 1) k_update(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    constexpr int SUBT = SubT<T, MT>::v;
    typedef PipeGeom<T, SUBT> G;
    extern __shared__ __align__(128) char dyn[];
    __shared__ unsigned long long full[2 * LB_PIPE_STAGES];
    __shared__ PipeSrc ps;
    __shared__ T sm[(2 * MT + 1) * (LBFGSB_BLOCK / 32)];
    const DevState<T>* s = w.s;
    if (!s->go || !s->do_update || s->fuse_uc) return;
    const i64 n = w.n;
    const int col = s->col, m = s->m, head0 = s->head - 1, itail0 = s->itail - 1;
    const T stp = s->stp;
    if (threadIdx.x == 0) {
        pipe_begin(&ps);
        pipe_add(&ps, w.g, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.gold, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.d, sizeof(T), G::REAL_SLOT);
        pipe_add_w<T>(&ps, w, head0, col - 1, G::REAL_SLOT);
        pipe_end(&ps);
    }
    T acc[2 * MT + 1];
#pragma unroll
    for (int k = 0; k < 2 * MT + 1; ++k) acc[k] = (T)0;
    T* wsn = w.ws + (i64)itail0 * w.ldw;
    T* wyn = w.wy + (i64)itail0 * w.ldw;
    (void)m;
    tma_pass<T, SUBT>(n, &ps, LB_DYN_STAGES(dyn), full, [&](i64 base, const char* sb, int lt) {
        T g[VEC], r[VEC], d[VEC];
        lds_real<T>(sb, 0 * G::REAL_SLOT, lt, g);
        lds_real<T>(sb, 1 * G::REAL_SLOT, lt, r);
        lds_real<T>(sb, 2 * G::REAL_SLOT, lt, d);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            r[v] = g[v] - r[v];
            if (stp != (T)1) d[v] = stp * d[v];
            if (base + v < n) acc[0] = acc[0] + r[v] * r[v];
        }
#pragma unroll
        for (int j = 0; j < MT; ++j) {
            if (j < col - 1) {
                T wy[VEC], wsv[VEC];
                lds_real<T>(sb, (3 + 2 * j) * G::REAL_SLOT, lt, wy);
                lds_real<T>(sb, (4 + 2 * j) * G::REAL_SLOT, lt, wsv);
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (base + v < n) {
                        acc[1 + j] = acc[1 + j] + d[v] * wy[v];
                        acc[1 + MT + j] = acc[1 + MT + j] + wsv[v] * d[v];
                    }
            }
        }
        stv<T>(wsn, base, n, d);
        stv<T>(wyn, base, n, r);
    });
    block_sum_store<T, 2 * MT + 1>(acc, 2 * MT + 1, sm, w.part);
}
template <typename T, int MT> constexpr unsigned smem_update() { return pipe_smem_bytes<T, SubT<T, MT>::v>(3 + 2 * (MT - 1), 0, 0); }

// ---------------------------------------------------------------------------
// cauchy, one variable of the per-variable pass (:1270-1327): new iwhere, Cauchy direction,
// the term of f1, breakpoint bookkeeping.  Shared by the standalone pass and by the pass fused
// with the S/Y update, so that both produce the same bits.
// ---------------------------------------------------------------------------
template <typename T>
struct CauchyScan {
    T bk; i64 ibk, nbr, nfc, bnd;
    __device__ __forceinline__ void init() { bk = LB_INF(T); ibk = LB_I64MAX; nbr = 0; nfc = 0; bnd = 1; }
};
template <typename T>
__device__ __forceinline__ void cauchy_classify_one(T x, T l, T u, T g, int nb, int& iw, T& d, bool& mv, T& d2acc,
                                                    CauchyScan<T>& cs, i64 gidx, T& tbp) {
    tbp = (T)-1;   // the variable's breakpoint, -1: none
    const T neggi = -g;
    T tl = (T)0, tu = (T)0;
    if (iw != 3 && iw != -1) {
        if (nb <= 2) tl = x - l;
        if (nb >= 2) tu = u - x;
        const bool xlower = nb <= 2 && tl <= (T)0;
        const bool xupper = nb >= 2 && tu <= (T)0;
        iw = 0;
        if (xlower) { if (neggi <= (T)0) iw = 1; }
        else if (xupper) { if (neggi >= (T)0) iw = 2; }
        else { if (fabs(neggi) <= (T)0) iw = -3; }
    }
    if (iw == 0 || iw == -1) {
        mv = true;
        d = neggi;
        d2acc = d2acc + neggi * neggi;
        T tb; bool hasb = false;
        if (nb <= 2 && nb != 0 && neggi < (T)0) { tb = tl / (-neggi); hasb = true; }
        else if (nb >= 2 && neggi > (T)0) { tb = tu / neggi; hasb = true; }
        if (hasb) {
            tbp = tb;
            cs.nbr++;
            if (tb < cs.bk) { cs.bk = tb; cs.ibk = gidx; }   // strict <: lowest index among ties (:1310)
        } else {
            cs.nfc++;
            if (fabs(neggi) > (T)0) cs.bnd = 0;
        }
    }
}
// block partials of the scan quantities -> part2 slot 2MT+1 (bkmin), ipart2 0..3
template <typename T, int MT>
__device__ __forceinline__ void cauchy_scan_store(const Wk<T>& w, CauchyScan<T>& cs, T* smv, i64* smi) {
    block_argmin<T>(cs.bk, cs.ibk, smv, smi);
    i64 r1 = block_isum(cs.nbr, smi), r2 = block_isum(cs.nfc, smi);
    i64 r3 = -block_isum(cs.bnd ? 0 : 1, smi);  // <0 if any thread saw an unbounded moving variable
    if (threadIdx.x == 0) {
        LB_SLOT(w.part2, 2 * MT + 1)[blockIdx.x] = cs.bk;
        LB_SLOT(w.ipart2, 0)[blockIdx.x] = cs.ibk;
        LB_SLOT(w.ipart2, 1)[blockIdx.x] = r1;
        LB_SLOT(w.ipart2, 2)[blockIdx.x] = r2;
        LB_SLOT(w.ipart2, 3)[blockIdx.x] = (r3 < 0) ? 0 : 1;
    }
}

// ---------------------------------------------------------------------------
// cauchy, per-variable pass (:1270-1341): classify iwhere, Cauchy direction d,
// f1 = -sum d^2, p = W'd, smallest breakpoint.  (xcp = x and d are implied, see lazy_gcp.)
// part2: [0,MT) sum Wy(:,j) d ; [MT,2MT) sum Ws(:,j) d ; 2MT: sum d^2 ; 2MT+1: bkmin
// ipart2: 0 argmin variable ; 1 nbreak ; 2 count of moving variables without breakpoint ;
//        3 bnded (min over blocks)
// ---------------------------------------------------------------------------
template <typename T, int MT>
__global__ void __launch_bounds__(LB_TMA_THREADS, 1) k_cauchy_classify(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    constexpr int SUBT = SubT<T, MT>::v;
    typedef PipeGeom<T, SUBT> G;
    extern __shared__ __align__(128) char dyn[];
    __shared__ unsigned long long full[2 * LB_PIPE_STAGES];
    __shared__ PipeSrc ps;
    __shared__ T sm[(2 * MT + 1) * (LBFGSB_BLOCK / 32)];
    __shared__ T smv[LBFGSB_BLOCK / 32];
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || s->classify_done) return;
    const int mode = s->cauchy_mode;
    const i64 n = w.n;
    if (mode != 0) {  // xcp = x only (:609 or :1247)
        if (threadIdx.x >= LBFGSB_BLOCK) return;
        LB_FOR_TILES(T, n, base) {
            T x[VEC];
            ldv<T>(w.x, base, n, x);
            stv<T>(w.z, base, n, x);
        }
        return;
    }
    const int col = s->col, head0 = s->head - 1;
    constexpr unsigned OX = 0, OG = G::REAL_SLOT, OL = 2 * G::REAL_SLOT, OU = 3 * G::REAL_SLOT, ONB = 4 * G::REAL_SLOT,
                       OIW = ONB + G::INT_SLOT, OW = OIW + G::INT_SLOT;
    if (threadIdx.x == 0) {
        pipe_begin(&ps);
        pipe_add(&ps, w.x, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.g, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.l, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.u, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.nbd, 4, G::INT_SLOT);
        pipe_add(&ps, w.iwhere, 4, G::INT_SLOT);
        pipe_add_w<T>(&ps, w, head0, col, G::REAL_SLOT);
        pipe_end(&ps);
    }
    T acc[2 * MT + 1];
#pragma unroll
    for (int k = 0; k < 2 * MT + 1; ++k) acc[k] = (T)0;
    CauchyScan<T> cs; cs.init();
    tma_pass<T, SUBT>(n, &ps, LB_DYN_STAGES(dyn), full, [&](i64 base, const char* sb, int lt) {
        T x[VEC], l[VEC], u[VEC], g[VEC], d[VEC];
        int nb[VEC], iw[VEC];
        lds_real<T>(sb, OX, lt, x); lds_real<T>(sb, OG, lt, g); lds_real<T>(sb, OL, lt, l); lds_real<T>(sb, OU, lt, u);
        lds_int<T>(sb, ONB, lt, nb); lds_int<T>(sb, OIW, lt, iw);
        bool mv[VEC]; T tk[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            mv[v] = false; d[v] = (T)0; tk[v] = (T)-1;
            if (base + v < n) cauchy_classify_one<T>(x[v], l[v], u[v], g[v], nb[v], iw[v], d[v], mv[v], acc[2 * MT], cs, base + v + w.off, tk[v]);
        }
        if (w.bp_hint) { stv<T>(w.r, base, n, tk); stv<T>(w.z, base, n, x); }   // in front of a probable walk
        // d and xcp = x are not written: they follow from iwhere, g and x (lazy_gcp; k_bp_count<T, true> writes them
        // out in front of a breakpoint walk)
        stvi<T>(w.iwhere, base, n, iw);
#pragma unroll
        for (int j = 0; j < MT; ++j) {
            if (j < col) {
                T wy[VEC], wsv[VEC];
                lds_real<T>(sb, OW + (2 * j) * G::REAL_SLOT, lt, wy);
                lds_real<T>(sb, OW + (2 * j + 1) * G::REAL_SLOT, lt, wsv);
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (mv[v]) {
                        acc[j] = acc[j] + wy[v] * d[v];
                        acc[MT + j] = acc[MT + j] + wsv[v] * d[v];
                    }
            }
        }
    });
    block_sum_store<T, 2 * MT + 1>(acc, 2 * MT + 1, sm, w.part2);
    cauchy_scan_store<T, MT>(w, cs, smv, smi);
}
template <typename T, int MT> constexpr unsigned smem_classify() { return pipe_smem_bytes<T, SubT<T, MT>::v>(4 + 2 * MT, 2, 0); }

// ---------------------------------------------------------------------------
// formk, new row/column of WN1 (:1756-1793): one pass over all rows of S,Y.
// part: [0,MT) A_j = sum_free Wy_last*Wy_j ; [MT,2MT) B_j = sum_act Ws_last*Ws_j ;
//       [2MT,3MT) C_j = sum_act Ws_last*Wy_j ; [3MT,4MT) D_j = sum_free Ws_j*Wy_last
// ---------------------------------------------------------------------------
template <typename T, int MT>
__global__ void __launch_bounds__(LB_TMA_THREADS, 1) k_formk_gram(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    constexpr int SUBT = SubT<T, MT>::v;
    typedef PipeGeom<T, SUBT> G;
    extern __shared__ __align__(128) char dyn[];
    __shared__ unsigned long long full[2 * LB_PIPE_STAGES];
    __shared__ PipeSrc ps;
    __shared__ T sm[4 * MT * (LBFGSB_BLOCK / 32)];
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->do_formk || !s->updatd) return;
    const i64 n = w.n;
    const int col = s->col, head0 = s->head - 1;
    if (threadIdx.x == 0) {
        pipe_begin(&ps);
        pipe_add_w<T>(&ps, w, head0, col, G::REAL_SLOT);
        pipe_add(&ps, w.state, 1, G::BYTE_SLOT);
        pipe_end(&ps);
    }
    const unsigned ost = 2u * (unsigned)col * G::REAL_SLOT;
    const unsigned olast = 2u * (unsigned)(col - 1) * G::REAL_SLOT;   // the newest pair sits at ring position col-1
    T acc[4 * MT];
#pragma unroll
    for (int k = 0; k < 4 * MT; ++k) acc[k] = (T)0;
    tma_pass<T, SUBT>(n, &ps, LB_DYN_STAGES(dyn), full, [&](i64 base, const char* sb, int lt) {
        int st[VEC];
        lds_byte<T>(sb, ost, lt, st);
        T wyl[VEC], wsl[VEC];
        lds_real<T>(sb, olast, lt, wyl);
        lds_real<T>(sb, olast + G::REAL_SLOT, lt, wsl);
#pragma unroll
        for (int j = 0; j < MT; ++j) {
            if (j < col) {
                T wy[VEC], wsv[VEC];
                lds_real<T>(sb, (2 * j) * G::REAL_SLOT, lt, wy);
                lds_real<T>(sb, (2 * j + 1) * G::REAL_SLOT, lt, wsv);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if (base + v < n) {
                        if (st[v] & 1) {
                            acc[j] = acc[j] + wyl[v] * wy[v];
                            acc[3 * MT + j] = acc[3 * MT + j] + wsv[v] * wyl[v];
                        } else {
                            acc[MT + j] = acc[MT + j] + wsl[v] * wsv[v];
                            acc[2 * MT + j] = acc[2 * MT + j] + wsl[v] * wy[v];
                        }
                    }
                }
            }
        }
    });
    block_sum_store<T, 4 * MT>(acc, 4 * MT, sm, w.part);
}
template <typename T, int MT> constexpr unsigned smem_formk() { return pipe_smem_bytes<T, SubT<T, MT>::v>(2 * MT, 0, 1); }

// ---------------------------------------------------------------------------
// cmprlb (:1565-1583) fused with the first half of subsm (:2742-2754):
//   r_k = -theta (z_k - x_k) - g_k + sum_j Wy(k,j) a1_j + Ws(k,j) a2_j   (free k)
//   wv  = W' Z r
// r is kept by variable (the reference keeps it compact over the free list).
// part2: [0,MT) sum Wy(:,j) r ; [MT,2MT) sum Ws(:,j) r
// ---------------------------------------------------------------------------
template <typename T, int MT>
__global__ void __launch_bounds__(LB_TMA_THREADS, 1) k_cmprlb_wv(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    constexpr int SUBT = SubT<T, MT>::v;
    typedef PipeGeom<T, SUBT> G;
    extern __shared__ __align__(128) char dyn[];
    __shared__ unsigned long long full[2 * LB_PIPE_STAGES];
    __shared__ PipeSrc ps;
    __shared__ T sm[2 * MT * (LBFGSB_BLOCK / 32)];
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->do_subspace) return;
    const i64 n = w.n;
    const int col = s->col, head0 = s->head - 1;
    const T theta = s->theta;
    const bool uc = (!s->cnstnd && col > 0);   // :1560-1563
    constexpr unsigned OG = 0, OZ = G::REAL_SLOT, OX = 2 * G::REAL_SLOT, OW = 3 * G::REAL_SLOT;
    if (threadIdx.x == 0) {
        pipe_begin(&ps);
        pipe_add(&ps, w.g, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.z, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.x, sizeof(T), G::REAL_SLOT);
        pipe_add_w<T>(&ps, w, head0, col, G::REAL_SLOT);
        pipe_add(&ps, w.state, 1, G::BYTE_SLOT);
        pipe_end(&ps);
    }
    const unsigned ost = OW + 2u * (unsigned)col * G::REAL_SLOT;
    T a1[MT], a2[MT], acc[2 * MT];
#pragma unroll
    for (int j = 0; j < MT; ++j) {
        a1[j] = (j < col) ? s->a[j] : (T)0;
        a2[j] = (j < col) ? theta * s->a[col + j] : (T)0;
        acc[j] = (T)0; acc[MT + j] = (T)0;
    }
    tma_pass<T, SUBT>(n, &ps, LB_DYN_STAGES(dyn), full, [&](i64 base, const char* sb, int lt) {
        int st[VEC];
        lds_byte<T>(sb, ost, lt, st);
        bool fr[VEC]; bool any = false;
#pragma unroll
        for (int v = 0; v < VEC; ++v) { fr[v] = (base + v < n) && (st[v] & 1); any |= fr[v]; }
        if (!any) return;
        T z[VEC], x[VEC], g[VEC], r[VEC];
        lds_real<T>(sb, OG, lt, g); lds_real<T>(sb, OZ, lt, z); lds_real<T>(sb, OX, lt, x);
#pragma unroll
        for (int v = 0; v < VEC; ++v) r[v] = uc ? -g[v] : (-theta * (z[v] - x[v]) - g[v]);
        if (!uc) {
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                if (j < col) {
                    T wy[VEC], wsv[VEC];
                    lds_real<T>(sb, OW + (2 * j) * G::REAL_SLOT, lt, wy);
                    lds_real<T>(sb, OW + (2 * j + 1) * G::REAL_SLOT, lt, wsv);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) r[v] = r[v] + wy[v] * a1[j] + wsv[v] * a2[j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < MT; ++j) {
            if (j < col) {
                T wy[VEC], wsv[VEC];
                lds_real<T>(sb, OW + (2 * j) * G::REAL_SLOT, lt, wy);
                lds_real<T>(sb, OW + (2 * j + 1) * G::REAL_SLOT, lt, wsv);
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (fr[v]) {
                        acc[j] = acc[j] + wy[v] * r[v];
                        acc[MT + j] = acc[MT + j] + wsv[v] * r[v];
                    }
            }
        }
        // r of a non-free variable is never read (every later pass masks on the free set),
        // so the whole vector is stored: no partial-sector writes.
        stv<T>(w.r, base, n, r);
    });
    block_sum_store<T, 2 * MT>(acc, 2 * MT, sm, w.part2);
}
template <typename T, int MT> constexpr unsigned smem_cmprlb() { return pipe_smem_bytes<T, SubT<T, MT>::v>(3 + 2 * MT, 0, 1); }

// ---------------------------------------------------------------------------
// subsm second half (:2770-2827): Newton direction on the free set, projected
// step, xp = xcp backup, and the directional derivative dd_p over all variables.
// part: 0 dd_p.  ipart: 0 iword (sum>0)
// ---------------------------------------------------------------------------
template <typename T, int MT>
__global__ void __launch_bounds__(LB_TMA_THREADS, 1) k_subsm_step(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    constexpr int SUBT = SubT<T, MT>::v;
    typedef PipeGeom<T, SUBT> G;
    extern __shared__ __align__(128) char dyn[];
    __shared__ unsigned long long full[2 * LB_PIPE_STAGES];
    __shared__ PipeSrc ps;
    __shared__ T sm[LBFGSB_BLOCK / 32];
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->do_subspace) return;
    const i64 n = w.n;
    const int col = s->col, head0 = s->head - 1;
    const T theta = s->theta, rtheta = (T)1 / theta;
    constexpr unsigned OZ = 0, OX = G::REAL_SLOT, OG = 2 * G::REAL_SLOT, OR = 3 * G::REAL_SLOT, OL = 4 * G::REAL_SLOT,
                       OU = 5 * G::REAL_SLOT, OW = 6 * G::REAL_SLOT;
    if (threadIdx.x == 0) {
        pipe_begin(&ps);
        pipe_add(&ps, w.z, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.x, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.g, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.r, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.l, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.u, sizeof(T), G::REAL_SLOT);
        pipe_add_w<T>(&ps, w, head0, col, G::REAL_SLOT);
        pipe_add(&ps, w.nbd, 4, G::INT_SLOT);
        pipe_add(&ps, w.state, 1, G::BYTE_SLOT);
        pipe_end(&ps);
    }
    const unsigned onb = OW + 2u * (unsigned)col * G::REAL_SLOT, ost = onb + G::INT_SLOT;
    T wv1[MT], wv2[MT];
#pragma unroll
    for (int j = 0; j < MT; ++j) { wv1[j] = (j < col) ? s->wv[j] : (T)0; wv2[j] = (j < col) ? s->wv[col + j] : (T)0; }
    T acc[1]; acc[0] = (T)0;
    i64 iwd = 0;
    tma_pass<T, SUBT>(n, &ps, LB_DYN_STAGES(dyn), full, [&](i64 base, const char* sb, int lt) {
        int st[VEC];
        lds_byte<T>(sb, ost, lt, st);
        bool fr[VEC]; bool any = false;
#pragma unroll
        for (int v = 0; v < VEC; ++v) { fr[v] = (base + v < n) && (st[v] & 1); any |= fr[v]; }
        T z[VEC], x[VEC], g[VEC];
        lds_real<T>(sb, OZ, lt, z); lds_real<T>(sb, OX, lt, x); lds_real<T>(sb, OG, lt, g);
        stv<T>(w.xp, base, n, z);   // :2787
        if (any) {
            T dk[VEC], l[VEC], u[VEC]; int nb[VEC];
            lds_real<T>(sb, OR, lt, dk); lds_real<T>(sb, OL, lt, l); lds_real<T>(sb, OU, lt, u);
            lds_int<T>(sb, onb, lt, nb);
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                if (j < col) {
                    T wy[VEC], wsv[VEC];
                    lds_real<T>(sb, OW + (2 * j) * G::REAL_SLOT, lt, wy);
                    lds_real<T>(sb, OW + (2 * j + 1) * G::REAL_SLOT, lt, wsv);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) dk[v] = dk[v] + wy[v] * wv1[j] / theta + wsv[v] * wv2[j];
                }
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                dk[v] = rtheta * dk[v];   // dscal(nsub, one/theta, d) :2780
                if (fr[v]) {
                    T xk = z[v];
                    if (nb[v] != 0) {
                        if (nb[v] == 1) { z[v] = dense::tmax(l[v], xk + dk[v]); if (z[v] == l[v]) iwd = 1; }
                        else if (nb[v] == 2) {
                            xk = dense::tmax(l[v], xk + dk[v]);
                            z[v] = dense::tmin(u[v], xk);
                            if (z[v] == l[v] || z[v] == u[v]) iwd = 1;
                        } else if (nb[v] == 3) { z[v] = dense::tmin(u[v], xk + dk[v]); if (z[v] == u[v]) iwd = 1; }
                    } else z[v] = xk + dk[v];
                }
            }
            // direction (don't-care on non-free variables) and new point (unchanged there)
            stv<T>(w.r, base, n, dk);
            stv<T>(w.z, base, n, z);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (base + v < n) acc[0] = acc[0] + (z[v] - x[v]) * g[v];   // :2825-2827
    });
    block_sum_store<T, 1>(acc, 1, sm, w.part);
    i64 r0 = block_isum(iwd, smi);
    if (threadIdx.x == 0) LB_SLOT(w.ipart, 0)[blockIdx.x] = r0;
}
template <typename T, int MT> constexpr unsigned smem_subsm() { return pipe_smem_bytes<T, SubT<T, MT>::v>(6 + 2 * MT, 1, 1); }

// ===========================================================================
// Cross-routine fusions.  The iteration streams the 2*col S/Y columns once per routine in the
// reference (matupd, cauchy, formk, cmprlb, subsm); two pairs of those passes have no global
// reduction between them and are merged here, so the history is read three times per iteration
// instead of five.  Thread -> element mapping, accumulation order and every arithmetic expression
// are those of the separate kernels above: the results are bit-identical.
// Used when the accumulators fit in registers (fused_passes_ok): double m <= 10, float m <= 20.
// ===========================================================================
template <typename T, int MT> constexpr bool fused_passes_ok() { return sizeof(T) * MT <= 80; }

// ---------------------------------------------------------------------------
// k_update (y/s preparation + matupd) fused with the per-variable pass of the NEXT cauchy
// (k_cauchy_classify).  Both run inside the same setulb call (NEW_X entry, :813-839 then :617)
// with only formt's 2m x 2m algebra between them, and everything cauchy's pass needs (x, g, l, u,
// nbd, iwhere, the ring after the update) is known here: the newest pair is taken from registers.
// Runs when s_newx_tests set fuse_uc (update not skipped, bounds present, sbgnrm > 0).
// part : as k_update            part2 / ipart2 : as k_cauchy_classify
// ---------------------------------------------------------------------------
template <typename T, int MT>
__global__ void __launch_bounds__(LB_TMA_THREADS, 1) k_update_classify(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    constexpr int SUBT = SubT<T, MT>::v;
    typedef PipeGeom<T, SUBT> G;
    extern __shared__ __align__(128) char dyn[];
    __shared__ unsigned long long full[2 * LB_PIPE_STAGES];
    __shared__ PipeSrc ps;
    __shared__ T sm[(2 * MT + 1) * (LBFGSB_BLOCK / 32)];
    __shared__ T smv[LBFGSB_BLOCK / 32];
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    const DevState<T>* s = w.s;
    if (!s->go || !s->do_update || !s->fuse_uc) return;
    const i64 n = w.n;
    const int col = s->col, head0 = s->head - 1, itail0 = s->itail - 1;
    const T stp = s->stp;
    constexpr unsigned OG = 0, OR = G::REAL_SLOT, OD = 2 * G::REAL_SLOT, OX = 3 * G::REAL_SLOT, OL = 4 * G::REAL_SLOT,
                       OU = 5 * G::REAL_SLOT, ONB = 6 * G::REAL_SLOT, OIW = ONB + G::INT_SLOT, OW = OIW + G::INT_SLOT;
    if (threadIdx.x == 0) {
        pipe_begin(&ps);
        pipe_add(&ps, w.g, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.gold, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.d, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.x, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.l, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.u, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.nbd, 4, G::INT_SLOT);
        pipe_add(&ps, w.iwhere, 4, G::INT_SLOT);
        pipe_add_w<T>(&ps, w, head0, col - 1, G::REAL_SLOT);
        pipe_end(&ps);
    }
    T au[2 * MT + 1], ac[2 * MT + 1];
#pragma unroll
    for (int k = 0; k < 2 * MT + 1; ++k) { au[k] = (T)0; ac[k] = (T)0; }
    CauchyScan<T> cs; cs.init();
    T* wsn = w.ws + (i64)itail0 * w.ldw;
    T* wyn = w.wy + (i64)itail0 * w.ldw;
    auto pass = [&](auto full_c) {
    constexpr bool FULLC = decltype(full_c)::value;   // col == MT at compile time
    tma_pass<T, SUBT>(n, &ps, LB_DYN_STAGES(dyn), full, [&](i64 base, const char* sb, int lt) {
        T g[VEC], y[VEC], sn[VEC];
        lds_real<T>(sb, OG, lt, g);
        lds_real<T>(sb, OR, lt, y);
        lds_real<T>(sb, OD, lt, sn);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            y[v] = g[v] - y[v];
            if (stp != (T)1) sn[v] = stp * sn[v];
            if (base + v < n) au[0] = au[0] + y[v] * y[v];
        }
        stv<T>(wsn, base, n, sn);
        stv<T>(wyn, base, n, y);
        T x[VEC], l[VEC], u[VEC], dc[VEC];
        int nb[VEC], iw[VEC];
        lds_real<T>(sb, OX, lt, x); lds_real<T>(sb, OL, lt, l); lds_real<T>(sb, OU, lt, u);
        lds_int<T>(sb, ONB, lt, nb); lds_int<T>(sb, OIW, lt, iw);
        bool mv[VEC]; T tk[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            mv[v] = false; dc[v] = (T)0; tk[v] = (T)-1;
            if (base + v < n) cauchy_classify_one<T>(x[v], l[v], u[v], g[v], nb[v], iw[v], dc[v], mv[v], ac[2 * MT], cs, base + v + w.off, tk[v]);
        }
        if (w.bp_hint) { stv<T>(w.r, base, n, tk); stv<T>(w.z, base, n, x); }   // in front of a probable walk
        stvi<T>(w.iwhere, base, n, iw);   // d, xcp: see k_cauchy_classify
#pragma unroll
        for (int j = 0; j < MT; ++j) {
            if (FULLC ? (j < MT - 1) : (j < col - 1)) {
                T wy[VEC], wsv[VEC];
                lds_real<T>(sb, OW + (2 * j) * G::REAL_SLOT, lt, wy);
                lds_real<T>(sb, OW + (2 * j + 1) * G::REAL_SLOT, lt, wsv);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if (base + v < n) {
                        au[1 + j] = au[1 + j] + sn[v] * wy[v];
                        au[1 + MT + j] = au[1 + MT + j] + wsv[v] * sn[v];
                    }
                    if (mv[v]) {
                        ac[j] = ac[j] + wy[v] * dc[v];
                        ac[MT + j] = ac[MT + j] + wsv[v] * dc[v];
                    }
                }
            } else if (FULLC ? (j == MT - 1) : (j == col - 1)) {   // the pair being written: ring position col-1
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (mv[v]) {
                        ac[j] = ac[j] + y[v] * dc[v];
                        ac[MT + j] = ac[MT + j] + sn[v] * dc[v];
                    }
            }
        }
    });
    };
    if constexpr (SpecializeFullHistory<T>::value) { if (col == MT) pass(BoolC<true>{}); else pass(BoolC<false>{}); }
    else pass(BoolC<false>{});
    block_sum_store<T, 2 * MT + 1>(au, 2 * MT + 1, sm, w.part);
    block_sum_store<T, 2 * MT + 1>(ac, 2 * MT + 1, sm, w.part2);
    cauchy_scan_store<T, MT>(w, cs, smv, smi);
}
template <typename T, int MT> constexpr unsigned smem_update_classify() { return pipe_smem_bytes<T, SubT<T, MT>::v>(6 + 2 * (MT - 1), 2, 0); }

// ---------------------------------------------------------------------------
// k_formk_gram fused with k_cmprlb_wv: the new row/column of WN1 over all rows of S,Y, the reduced
// gradient r on the free set and wv = W'Zr in one pass.  a = M c (cmprlb's bmv, :1569) depends only
// on sy, wt and c and is therefore computed before this pass (s_freev).
// part : as k_formk_gram (written when the Gram row is needed)     part2 : as k_cmprlb_wv
// ---------------------------------------------------------------------------
template <typename T, int MT, int GF>
__global__ void __launch_bounds__(LB_TMA_THREADS, 1) k_formk_cmprlb(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    constexpr int SUBT = SubT<T, MT>::v;
    typedef PipeGeom<T, SUBT> G;
    extern __shared__ __align__(128) char dyn[];
    __shared__ unsigned long long full[2 * LB_PIPE_STAGES];
    __shared__ PipeSrc ps;
    __shared__ T sm[4 * MT * (LBFGSB_BLOCK / 32)];
    __shared__ T coef[2 * MT];
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    const DevState<T>* s = w.s;
    if (!s->go || s->pause || !s->in_body || !s->do_subspace) return;
    // GF != 0: the tail of cauchy (xcp = xcp + tsum*d, :1515) and freev (:1980-2059) are part of this pass; the flags
    // do_subspace / do_formk are then s_freev's tentative values (the counts are only known after this pass).
    // GF = 1: no breakpoint walk ran, xcp = x is implied and the Cauchy point is not stored either (lazy_z);
    // GF = 2: after a walk -- xcp holds the bounds of the variables the walk fixed, and the Cauchy point is stored.
    // The host launches the instantiation(s) that can match the device flag fuse_gf.
    constexpr bool gf = GF != 0;
    if (s->fuse_gf != GF) return;
    const bool gram = s->do_formk && s->updatd;
    const i64 n = w.n;
    const int col = s->col, head0 = s->head - 1;
    const T theta = s->theta;
    const bool uc = (!s->cnstnd && col > 0);   // :1560-1563
    const T tsum = s->tsum;
    const bool axpy = tsum != (T)0;            // daxpy early-out (:49-50)
    const bool cnt = (s->iter > 0 && s->cnstnd);
    constexpr unsigned OG = 0, OX = G::REAL_SLOT, OZ = 2 * G::REAL_SLOT;
    constexpr unsigned OW = (GF == 1 ? 2u : 3u) * G::REAL_SLOT;
    if (threadIdx.x == 0) {
        pipe_begin(&ps);
        pipe_add(&ps, w.g, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.x, sizeof(T), G::REAL_SLOT);
        if (GF != 1) pipe_add(&ps, w.z, sizeof(T), G::REAL_SLOT);
        pipe_add_w<T>(&ps, w, head0, col, G::REAL_SLOT);
        pipe_add(&ps, w.state, 1, G::BYTE_SLOT);
        if (gf) pipe_add(&ps, w.iwhere, 4, G::INT_SLOT);
        pipe_end(&ps);
    }
    if (threadIdx.x < 2 * MT) {
        const int j = threadIdx.x % MT;
        coef[threadIdx.x] = (j < col) ? ((threadIdx.x < MT) ? s->a[j] : theta * s->a[col + j]) : (T)0;
    }
    T af[4 * MT], aw[2 * MT];
#pragma unroll
    for (int k = 0; k < 4 * MT; ++k) af[k] = (T)0;
#pragma unroll
    for (int k = 0; k < 2 * MT; ++k) aw[k] = (T)0;
    int nfr = 0, nen = 0, nle = 0;   // per thread: at most n / (GRID*BLOCK) * VEC, far below 2^31
    // (tma_pass starts with a __syncthreads: coef is visible to every consumer)
    // Elements at or beyond n read as zero from the stage (state 0, W = 0): their terms are +0 and leave
    // every accumulator unchanged, so the loops below carry no range checks.  The conditionals are kept
    // to a couple of instructions so that they compile to predication, not branches.
    auto pass = [&](auto full_c) {
    // FULLC: the steady state (history full, Gram row needed, bounds present) with col, gram, uc as compile-time constants
    constexpr bool FULLC = decltype(full_c)::value;
    const int colv = FULLC ? MT : col;
    const bool gramv = FULLC ? true : gram, ucv = FULLC ? false : uc;
    const unsigned ost = OW + 2u * (unsigned)colv * G::REAL_SLOT;
    const unsigned oiw = ost + G::BYTE_SLOT;
    const unsigned olast = OW + 2u * (unsigned)(colv - 1) * G::REAL_SLOT;   // the newest pair sits at ring position col-1
    tma_pass<T, SUBT>(n, &ps, LB_DYN_STAGES(dyn), full, [&](i64 base, const char* sb, int lt) {
        int st[VEC];
        lds_byte<T>(sb, ost, lt, st);
        bool fr[VEC]; bool any = false;
        T r[VEC];
        if (gf) {
            int iw[VEC];
            lds_int<T>(sb, oiw, lt, iw);
            T z[VEC], x[VEC], g[VEC];
            lds_real<T>(sb, OG, lt, g); lds_real<T>(sb, OX, lt, x);
            if (GF == 2) lds_real<T>(sb, OZ, lt, z);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const bool inr = base + v < n;
                const bool f = inr && iw[v] <= 0;
                const int old = st[v] & 1;
                nfr += f ? 1 : 0;
                if (cnt) { nen += (f && !old) ? 1 : 0; nle += (inr && !f && old) ? 1 : 0; }
                // bit 2: the variable moves along the Cauchy direction (d = -g): with it xcp can be formed again
                // from x and g by the subspace pass, and is not stored here (lazy_z)
                st[v] = (f ? 1 : 0) | ((cnt ? old : (f ? 1 : 0)) << 1) | ((iw[v] == 0 || iw[v] == -1) ? 4 : 0);
                fr[v] = f; any |= f;
                const T z0 = (GF == 2) ? z[v] : x[v];
                z[v] = axpy ? (z0 + tsum * cauchy_dir<T>(iw[v], g[v])) : z0;
                r[v] = ucv ? -g[v] : (-theta * (z[v] - x[v]) - g[v]);
            }
            stvb<T>(w.state, base, n, st);
            if (GF == 2 && axpy) stv<T>(w.z, base, n, z);   // the Cauchy point (:1515), read by the subspace pass
            if (!any && !gramv) return;
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { fr[v] = (st[v] & 1) != 0; any |= fr[v]; }
            if (!any && !gramv) return;
            T z[VEC], x[VEC], g[VEC];
            lds_real<T>(sb, OG, lt, g); lds_real<T>(sb, OZ, lt, z); lds_real<T>(sb, OX, lt, x);
#pragma unroll
            for (int v = 0; v < VEC; ++v) r[v] = ucv ? -g[v] : (-theta * (z[v] - x[v]) - g[v]);
        }
        T wl[VEC];   // the newest pair's Wy on a free row, its Ws on an active row
        if (gramv) {
            T wyl[VEC], wsl[VEC];
            lds_real<T>(sb, olast, lt, wyl);
            lds_real<T>(sb, olast + G::REAL_SLOT, lt, wsl);
#pragma unroll
            for (int v = 0; v < VEC; ++v) wl[v] = fr[v] ? wyl[v] : wsl[v];
        }
        // PREF (REAL32): the staged values and coefficients of column j + 1 are requested before the arithmetic of
        // column j (volatile, so that the loads stay where they are written); same operations, same order
        constexpr bool PREF = SpecializeFullHistory<T>::value;
        if (gramv || !ucv) {
            T wyn[VEC], wsn[VEC]; T a1n = (T)0, a2n = (T)0;
            if (PREF) {
                lds_real_v<T>(sb, OW, lt, wyn); lds_real_v<T>(sb, OW + G::REAL_SLOT, lt, wsn);
                a1n = *(volatile const T*)&coef[0]; a2n = *(volatile const T*)&coef[MT];
            }
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                if (j < col) {
                    T wy[VEC], wsv[VEC]; T a1, a2;
                    if (PREF) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) { wy[v] = wyn[v]; wsv[v] = wsn[v]; }
                        a1 = a1n; a2 = a2n;
                        if (j + 1 < MT) {
                            lds_real_v<T>(sb, OW + (2 * j + 2) * G::REAL_SLOT, lt, wyn);
                            lds_real_v<T>(sb, OW + (2 * j + 3) * G::REAL_SLOT, lt, wsn);
                            a1n = *(volatile const T*)&coef[j + 1]; a2n = *(volatile const T*)&coef[MT + j + 1];
                        }
                    } else {
                        lds_real<T>(sb, OW + (2 * j) * G::REAL_SLOT, lt, wy);
                        lds_real<T>(sb, OW + (2 * j + 1) * G::REAL_SLOT, lt, wsv);
                        a1 = coef[j]; a2 = coef[MT + j];
                    }
                    if (gramv) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            const T py = wl[v] * wy[v], psv = wl[v] * wsv[v];
                            if (fr[v]) { af[j] = af[j] + py; af[3 * MT + j] = af[3 * MT + j] + psv; }
                            else { af[2 * MT + j] = af[2 * MT + j] + py; af[MT + j] = af[MT + j] + psv; }
                        }
                    }
                    if (!ucv) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) r[v] = r[v] + wy[v] * a1 + wsv[v] * a2;
                    }
                }
            }
        }
        if (!any) return;
        {
            T wyn[VEC], wsn[VEC];
            if (PREF) { lds_real_v<T>(sb, OW, lt, wyn); lds_real_v<T>(sb, OW + G::REAL_SLOT, lt, wsn); }
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                if (j < col) {
                    T wy[VEC], wsv[VEC];
                    if (PREF) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) { wy[v] = wyn[v]; wsv[v] = wsn[v]; }
                        if (j + 1 < MT) {
                            lds_real_v<T>(sb, OW + (2 * j + 2) * G::REAL_SLOT, lt, wyn);
                            lds_real_v<T>(sb, OW + (2 * j + 3) * G::REAL_SLOT, lt, wsn);
                        }
                    } else {
                        lds_real<T>(sb, OW + (2 * j) * G::REAL_SLOT, lt, wy);
                        lds_real<T>(sb, OW + (2 * j + 1) * G::REAL_SLOT, lt, wsv);
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const T py = wy[v] * r[v], psv = wsv[v] * r[v];
                        if (fr[v]) { aw[j] = aw[j] + py; aw[MT + j] = aw[MT + j] + psv; }
                    }
                }
            }
        }
        stv<T>(w.r, base, n, r);
    });
    };
    pass(BoolC<false>{});   // (the col == MT instantiation of this pass runs out of registers: 120 accumulators)
    if (gram) block_sum_store<T, 4 * MT>(af, 4 * MT, sm, w.part);
    block_sum_store<T, 2 * MT>(aw, 2 * MT, sm, w.part2);
    if (gf) {   // site freev
        i64 r0 = block_isum((i64)nfr, smi), r1 = block_isum((i64)nen, smi), r2 = block_isum((i64)nle, smi);
        if (threadIdx.x == 0) {
            LB_SLOT(w.ipart, 0)[blockIdx.x] = r0; LB_SLOT(w.ipart, 1)[blockIdx.x] = r1;
            LB_SLOT(w.ipart, 2)[blockIdx.x] = r2;
        }
    }
}
template <typename T, int MT> constexpr unsigned smem_formk_cmprlb() { return pipe_smem_bytes<T, SubT<T, MT>::v>(3 + 2 * MT, 1, 1); }

// ---------------------------------------------------------------------------
// k_subsm_step fused with k_ls_init: the subspace pass has z (the Newton point), x, g, l, u, nbd of every
// variable in hand, which is all that d = z - x (:720-722) and the first entry of lnsrlb (:2196-2244) need.
// gd = g.d is the same sum as subsm's dd_p (same products, same order).  The results stand unless the
// backtrack (:2830-2879, rare) moves z afterwards; then k_ls_init runs as a separate pass (lsinit_done = 0).
// part: 0 dd_p ; ipart: 0 iword        part2: 0 dtd ; 1 gd ; 2 stpmx candidate (min)
//
// Speculative step (spec_step, set by s_subsm_dense when lnsrlb's first trial will be stp = 1, i.e. iter > 0
// or a boxed problem): that trial point is the Newton point itself (x = z, :2265), so it is written straight
// into x and neither z, nor the backup xp (:2787), nor the direction r (read only by the backtrack) are
// stored.  If the backtrack is needed after all, xcp is still intact in z, the iterate is in t, and PASS 1
// of this kernel (same arithmetic, no other effect) writes the direction into r.
//
// With lazy_z, xcp is not read from memory either: it is x + tsum*d with d = -g on the variables whose state
// bit 2 is set (one multiply-add from streams that pass through this kernel anyway); PASS 1 then also stores
// xcp for the backtrack.  (Forming cmprlb's reduced gradient r again here in the same way was measured and
// rejected: its 2*col dependent additions per variable double the dependency chain of this pass, which is
// latency-bound at 8 warps per SM, and cost as much time as the 16 bytes per variable saved.)
// ---------------------------------------------------------------------------
template <typename T, int MT, int PASS>
__global__ void __launch_bounds__(LB_TMA_THREADS, 1) k_subsm_lsinit(Wk<T> w) {
    constexpr int VEC = Real<T>::VEC;
    constexpr int SUBT = SubT<T, MT>::v;
    typedef PipeGeom<T, SUBT> G;
    extern __shared__ __align__(128) char dyn[];
    __shared__ unsigned long long full[2 * LB_PIPE_STAGES];
    __shared__ PipeSrc ps;
    __shared__ T sm[2 * (LBFGSB_BLOCK / 32)];
    __shared__ T smm[LBFGSB_BLOCK / 32];
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    const DevState<T>* s = w.s;
    if (!s->go || s->pause || !s->in_body || !s->do_subspace) return;
    const bool spec = s->spec_step != 0;
    if (PASS == 1 && !(spec && s->do_backtrack)) return;
    const i64 n = w.n;
    const int col = s->col, head0 = s->head - 1;
    const T theta = s->theta, rtheta = (T)1 / theta;
    const bool bounds = (s->cnstnd && s->iter != 0);
    // lz: xcp was not stored by k_formk_cmprlb (fuse_gf); it is x + tsum*d with d = -g where state bit 2 is set
    const bool lz = s->lazy_z != 0;
    const T tsum = s->tsum;
    const bool axpy = tsum != (T)0;
    constexpr unsigned OX = 0, OG = G::REAL_SLOT, OL = 2 * G::REAL_SLOT, OU = 3 * G::REAL_SLOT, OR = 4 * G::REAL_SLOT,
                       OZ = 5 * G::REAL_SLOT;
    const unsigned OW = (lz ? 5u : 6u) * G::REAL_SLOT;
    if (threadIdx.x == 0) {
        pipe_begin(&ps);
        // PASS 1 runs after PASS 0 stepped speculatively (x = z): the iterate is then in t
        pipe_add(&ps, PASS == 1 ? w.t : w.x, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.g, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.l, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.u, sizeof(T), G::REAL_SLOT);
        pipe_add(&ps, w.r, sizeof(T), G::REAL_SLOT);
        if (!lz) pipe_add(&ps, w.z, sizeof(T), G::REAL_SLOT);
        pipe_add_w<T>(&ps, w, head0, col, G::REAL_SLOT);
        pipe_add(&ps, w.nbd, 4, G::INT_SLOT);
        pipe_add(&ps, w.state, 1, G::BYTE_SLOT);
        pipe_end(&ps);
    }
    T wv1[MT], wv2[MT];
#pragma unroll
    for (int j = 0; j < MT; ++j) { wv1[j] = (j < col) ? s->wv[j] : (T)0; wv2[j] = (j < col) ? s->wv[col + j] : (T)0; }
    T acc[2]; acc[0] = (T)0; acc[1] = (T)0;   // dtd, dd_p (= gd)
    T smx = LB_INF(T);
    i64 iwd = 0;
    auto pass = [&](auto full_c) {
    constexpr bool FULLC = decltype(full_c)::value;   // col == MT at compile time
    const unsigned onb = OW + 2u * (unsigned)(FULLC ? MT : col) * G::REAL_SLOT, ost = onb + G::INT_SLOT;
    tma_pass<T, SUBT>(n, &ps, LB_DYN_STAGES(dyn), full, [&](i64 base, const char* sb, int lt) {
        int st[VEC];
        lds_byte<T>(sb, ost, lt, st);
        bool fr[VEC]; bool any = false;
#pragma unroll
        for (int v = 0; v < VEC; ++v) { fr[v] = (base + v < n) && (st[v] & 1); any |= fr[v]; }
        if (PASS == 1 && !any && !lz) return;
        T z[VEC], x[VEC], g[VEC];
        lds_real<T>(sb, OX, lt, x); lds_real<T>(sb, OG, lt, g);
        if (lz) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) z[v] = axpy ? (x[v] + tsum * ((st[v] & 4) ? -g[v] : (T)0)) : x[v];
        } else lds_real<T>(sb, OZ, lt, z);
        // the Newton direction of subsm on this sub-tile (free variables), from cmprlb's reduced gradient
        T dk[VEC];
        if (any) {
            lds_real<T>(sb, OR, lt, dk);
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                if (FULLC || j < col) {
                    T wy[VEC], wsv[VEC];
                    lds_real<T>(sb, OW + (2 * j) * G::REAL_SLOT, lt, wy);
                    lds_real<T>(sb, OW + (2 * j + 1) * G::REAL_SLOT, lt, wsv);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) dk[v] = dk[v] + wy[v] * wv1[j] / theta + wsv[v] * wv2[j];
                }
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) dk[v] = rtheta * dk[v];   // dscal(nsub, one/theta, d) :2780
        }
        if (PASS == 1) {   // what the backtrack reads after a speculative step: the direction, and xcp in z
            if (any) stv<T>(w.r, base, n, dk);
            if (lz) stv<T>(w.z, base, n, z);
            return;
        }
        T l[VEC], u[VEC]; int nb[VEC];
        lds_real<T>(sb, OL, lt, l); lds_real<T>(sb, OU, lt, u);
        lds_int<T>(sb, onb, lt, nb);
        if (!spec) stv<T>(w.xp, base, n, z);   // :2787
        if (any) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (fr[v]) {
                    T xk = z[v];
                    if (nb[v] != 0) {
                        if (nb[v] == 1) { z[v] = dense::tmax(l[v], xk + dk[v]); if (z[v] == l[v]) iwd = 1; }
                        else if (nb[v] == 2) {
                            xk = dense::tmax(l[v], xk + dk[v]);
                            z[v] = dense::tmin(u[v], xk);
                            if (z[v] == l[v] || z[v] == u[v]) iwd = 1;
                        } else if (nb[v] == 3) { z[v] = dense::tmin(u[v], xk + dk[v]); if (z[v] == u[v]) iwd = 1; }
                    } else z[v] = xk + dk[v];
                }
            }
            if (!spec) {
                // direction (don't-care on non-free variables) and new point (unchanged there)
                stv<T>(w.r, base, n, dk);
                stv<T>(w.z, base, n, z);
            }
        } else if (!spec && lz) stv<T>(w.z, base, n, z);
        // d = z - x, dtd, gd (= dd_p :2825-2827), stpmx candidates (:2201-2227), t = x, gold = g
        T d[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            d[v] = z[v] - x[v];
            if (base + v < n) { acc[0] = acc[0] + d[v] * d[v]; acc[1] = acc[1] + d[v] * g[v]; }
        }
        stv<T>(w.d, base, n, d); stv<T>(w.t, base, n, x); stv<T>(w.gold, base, n, g);
        if (spec) stv<T>(w.x, base, n, z);   // the stp = 1 trial point (:2265)
        if (bounds) {
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
    });
    };
    if constexpr (SpecializeFullHistory<T>::value && PASS == 0) { if (col == MT) pass(BoolC<true>{}); else pass(BoolC<false>{}); }
    else pass(BoolC<false>{});
    if (PASS == 1) return;
    // site subsm: dd_p in part slot 0, iword in ipart slot 0
    T ddp[1]; ddp[0] = acc[1];
    block_sum_store<T, 1>(ddp, 1, sm, w.part);
    i64 r0 = block_isum(iwd, smi);
    if (threadIdx.x == 0) LB_SLOT(w.ipart, 0)[blockIdx.x] = r0;
    // site lsinit: dtd, gd, stpmx in part2
    block_sum_store<T, 2>(acc, 2, sm, w.part2);
    T rm = block_min<T>(smx, smm);
    if (threadIdx.x == 0) LB_SLOT(w.part2, 2)[blockIdx.x] = rm;
}
