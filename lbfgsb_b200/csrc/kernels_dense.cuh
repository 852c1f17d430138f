// Single-CTA kernels: finish the block partials of the preceding streaming pass,
// run the 2m x 2m dense algebra (bmv, formt, formk tail, subsm solves, dcsrch)
// and take every data-dependent branch of mainlb (src/lbfgsb.f90:599-872) on the
// device, writing the flags that predicate the kernels enqueued after them.
//
// In a sharded run (R ranks) the partials of every rank are finished locally
// (k_rank_finish), all-gathered, and combined here in rank order, so every rank
// takes bit-identical decisions.
#pragma once
#include "kernels_stream.cuh"

#define LB_SCALAR_THREADS 256

// Reduced values of one site.  rv[] real slots, iv[] integer slots.
template <typename T>
struct Red {
    T rv[LB_KMAX];
    i64 iv[LB_IMAX];
};

// op codes for the slots of a site
enum { OP_SUM = 0, OP_MAX = 1, OP_MIN = 2, OP_ARGMIN = 3 /* real slot paired with iv[0] */ };

// Per-site description: ops of the real slots come as (first, count, op) runs.
struct SiteSpec {
    int nreal;            // real slots in use
    int nint;             // integer slots in use
    int argmin_slot;      // real slot that carries the ARGMIN value (-1: none); its index is iv[0]
    int max_from, max_to; // real slots [max_from, max_to) are OP_MAX
    int min_from, min_to; // real slots [min_from, min_to) are OP_MIN (excluding argmin_slot)
    int imin_from, imin_to;  // integer slots that are minima; others are sums; iv[0] is the argmin index if argmin_slot>=0
    int imax_from, imax_to;  // integer slots that are maxima
    int buf;              // 0: partials in w.part / w.ipart ; 1: in w.part2 / w.ipart2
};

// Finish the local block partials of a site into red (shared memory), using all warps.  The site's real slots land
// at red->rv[roff + k], its integer slots at red->iv[ioff + k] (a merged record holds several sites back to back).
template <typename T>
__device__ void site_finish_local(const Wk<T>& w, const SiteSpec sp, Red<T>* red, int roff = 0, int ioff = 0) {
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
    const T* part = sp.buf ? w.part2 : w.part;
    const i64* ipart = sp.buf ? w.ipart2 : w.ipart;
    for (int k = wid; k < sp.nreal; k += nw) {
        const T* p = LB_SLOT(part, k);
        T r;
        if (k == sp.argmin_slot) {
            i64 idx;
            final_argmin_warp<T>(p, LB_SLOT(ipart, 0), r, idx);
            if (lane == 0) red->iv[ioff] = idx;
        } else if (k >= sp.max_from && k < sp.max_to) r = final_max_warp<T>(p);
        else if (k >= sp.min_from && k < sp.min_to) r = final_min_warp<T>(p);
        else r = final_sum_warp<T>(p);
        if (lane == 0) red->rv[roff + k] = r;
    }
    for (int k = wid; k < sp.nint; k += nw) {
        if (k == 0 && sp.argmin_slot >= 0) continue;
        const i64* p = LB_SLOT(ipart, k);
        i64 r;
        if (k >= sp.imin_from && k < sp.imin_to) r = final_imin_warp(p);
        else if (k >= sp.imax_from && k < sp.imax_to) r = final_imax_warp(p);
        else r = final_isum_warp(p);
        if (lane == 0) red->iv[ioff + k] = r;
    }
    __syncthreads();
}

// Combine the per-rank records (rank order) into red; the site sits at (roff, ioff) inside each record.
template <typename T>
__device__ void site_combine_ranks(const Red<T>* all, int R, const SiteSpec sp, Red<T>* red, int roff = 0, int ioff = 0) {
    for (int k = threadIdx.x; k < sp.nreal; k += blockDim.x) {
        T r = all[0].rv[roff + k];
        if (k == sp.argmin_slot) {
            i64 idx = all[0].iv[ioff];
            for (int q = 1; q < R; ++q) {
                T ov = all[q].rv[roff + k]; i64 oi = all[q].iv[ioff];
                if (ov < r || (ov == r && oi < idx)) { r = ov; idx = oi; }
            }
            red->iv[0] = idx;
        } else if (k >= sp.max_from && k < sp.max_to) {
            for (int q = 1; q < R; ++q) r = all[q].rv[roff + k] > r ? all[q].rv[roff + k] : r;
        } else if (k >= sp.min_from && k < sp.min_to) {
            for (int q = 1; q < R; ++q) r = all[q].rv[roff + k] < r ? all[q].rv[roff + k] : r;
        } else {
            for (int q = 1; q < R; ++q) r = r + all[q].rv[roff + k];
        }
        red->rv[k] = r;
    }
    for (int k = threadIdx.x; k < sp.nint; k += blockDim.x) {
        if (k == 0 && sp.argmin_slot >= 0) continue;
        i64 r = all[0].iv[ioff + k];
        if (k >= sp.imin_from && k < sp.imin_to) { for (int q = 1; q < R; ++q) r = all[q].iv[ioff + k] < r ? all[q].iv[ioff + k] : r; }
        else if (k >= sp.imax_from && k < sp.imax_to) { for (int q = 1; q < R; ++q) r = all[q].iv[ioff + k] > r ? all[q].iv[ioff + k] : r; }
        else { for (int q = 1; q < R; ++q) r += all[q].iv[ioff + k]; }
        red->iv[k] = r;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// Exchange of the per-rank records over peer memory (NVLink / NVSwitch).  Every rank owns one P2PBuf; the
// finish kernel of a reduction site stores its record straight into slot [rank] of every peer's buffer
// (remote stores), fences, and then publishes the site's sequence number in the peer's flag word; the scalar
// kernel that consumes the site spins on its own R flag words.  No collective launch sits between a streaming
// pass and the 2m x 2m algebra that needs its sums.  Two slots alternate, and the protocol is safe by construction:
//   * every launched producer has exactly one consumer kernel, launched right after it on every rank (the host
//     sequence is identical on all ranks because every decision is taken from bit-identical combined records);
//   * that consumer ALWAYS waits for all R flags of its site and copies the records out BEFORE it tests any
//     predicate that could make it skip its work (site_fetch at the top of every scalar kernel).
// Hence a rank produces site j+2 (same slot as j) only after its consumer of j+1 saw every rank's record of j+1,
// which each rank stored after its own consumer of j had finished reading slot j (stream order).
// ---------------------------------------------------------------------------
#ifndef LB_MAXR
#define LB_MAXR 16
#endif
#define LB_P2P_SPIN_LIMIT (20000000000LL)   // ~10 s of SM clocks: a lost peer ends in an error, not in a hung GPU
template <typename T>
struct P2PBuf {
    Red<T> rec[2][LB_MAXR];
    unsigned long long flag[2][LB_MAXR];
    unsigned long long dflag[LB_MAXR];     // formk's entering/leaving corrections (delta_all) of rank q are complete
    // chain-coupled sample objectives on a shard (lbfgsb_problem_sharded_f64): the neighbours' boundary values and the
    // per-rank parts of f, two alternating slots like the records
    T halo[2][2];                          // [slot][0]: last x of the left neighbour, [slot][1]: first x of the right one
    unsigned long long hflag[2][2];
    T fpart[2][LB_MAXR];
    unsigned long long fflag[2][LB_MAXR];
};
struct Peers { void* p[LB_MAXR]; };

// Distribution context passed to every scalar kernel.
template <typename T>
struct Dist {
    int R;                 // ranks (1: single GPU)
    Red<T>* all;           // [R] gathered records (R > 1)
    P2PBuf<T>* p2p;        // this rank's peer-memory buffer (nullptr: the records came through ncclAllGather)
    int slot;              // which of the two record slots the current site uses
    unsigned long long seq;   // sequence number of the current site
    unsigned long long dseq;  // sequence number of the current delta exchange
    int wait;              // a producer of this kernel's record was launched right before it (site_fetch waits iff set)
};

// wait until every rank's word equals seq (threads 0..R-1 spin), then make the peers' stores visible
template <typename T>
__device__ __forceinline__ void p2p_wait(const Wk<T>& w, volatile unsigned long long* flags, int R, unsigned long long seq) {
    if ((int)threadIdx.x < R) {
        const long long t0 = clock64();
        while (flags[threadIdx.x] != seq) {
            if (clock64() - t0 > LB_P2P_SPIN_LIMIT) { w.s->p2p_timeout = 1; break; }
        }
    }
    __threadfence_system();
    __syncthreads();
}

// First statement of every scalar kernel on a sharded run: wait for the R records of the site produced right before
// this kernel and copy them out of the peer-written buffer -- unconditionally (see the protocol above).
template <typename T>
__device__ __forceinline__ void site_fetch(const Wk<T>& w, const Dist<T>& dist) {
    if (dist.R <= 1 || !dist.p2p || !dist.wait) return;
    p2p_wait<T>(w, dist.p2p->flag[dist.slot], dist.R, dist.seq);
    // volatile: never from a stale L1 line
    const int words = (int)(sizeof(Red<T>) / 8);
    for (int q = 0; q < dist.R; ++q) {
        const volatile unsigned long long* src = (const volatile unsigned long long*)&dist.p2p->rec[dist.slot][q];
        unsigned long long* dst = (unsigned long long*)&dist.all[q];
        for (int k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
}
// The reduced values of one site: local block partials on a single GPU, the fetched (or all-gathered) per-rank
// records combined in rank order on a sharded run.
template <typename T>
__device__ __forceinline__ void site_reduce(const Wk<T>& w, const Dist<T>& dist, const SiteSpec sp, Red<T>* red, int roff = 0,
                                            int ioff = 0) {
    if (dist.R <= 1) { site_finish_local<T>(w, sp, red); return; }
    site_combine_ranks<T>(dist.all, dist.R, sp, red, roff, ioff);
}

// A record that carries up to three sites back to back (the fast pipeline's merged scalar kernels).
struct MSite { int nsub; SiteSpec sp[3]; int roff[3]; int ioff[3]; };

// Rank-local finish of a site, stored into every peer's buffer (see P2PBuf).
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) k_rank_finish_p2p(Wk<T> w, SiteSpec sp, Peers peers, int R, int rank, int slot,
                                                                      unsigned long long seq) {
    __shared__ Red<T> red;
    site_finish_local<T>(w, sp, &red);
    for (int q = 0; q < R; ++q) {
        Red<T>* dst = &((P2PBuf<T>*)peers.p[q])->rec[slot][rank];
        for (int k = threadIdx.x; k < LB_KMAX; k += blockDim.x) dst->rv[k] = (k < sp.nreal) ? red.rv[k] : (T)0;
        for (int k = threadIdx.x; k < LB_IMAX; k += blockDim.x) dst->iv[k] = (k < sp.nint) ? red.iv[k] : 0;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < R) {
        __threadfence_system();
        *(volatile unsigned long long*)&((P2PBuf<T>*)peers.p[threadIdx.x])->flag[slot][rank] = seq;
    }
}
// The same for a merged record; p2p = 0: into `out` for the all-gather that follows.
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) k_rank_finish_m(Wk<T> w, MSite ms, Red<T>* out, Peers peers, int R, int rank,
                                                                    int slot, unsigned long long seq, int p2p) {
    __shared__ Red<T> red;
    for (int k = threadIdx.x; k < LB_KMAX; k += blockDim.x) red.rv[k] = (T)0;
    for (int k = threadIdx.x; k < LB_IMAX; k += blockDim.x) red.iv[k] = 0;
    __syncthreads();
    for (int q = 0; q < ms.nsub; ++q) site_finish_local<T>(w, ms.sp[q], &red, ms.roff[q], ms.ioff[q]);
    if (!p2p) {
        for (int k = threadIdx.x; k < LB_KMAX; k += blockDim.x) out->rv[k] = red.rv[k];
        for (int k = threadIdx.x; k < LB_IMAX; k += blockDim.x) out->iv[k] = red.iv[k];
        return;
    }
    for (int q = 0; q < R; ++q) {
        Red<T>* dst = &((P2PBuf<T>*)peers.p[q])->rec[slot][rank];
        for (int k = threadIdx.x; k < LB_KMAX; k += blockDim.x) dst->rv[k] = red.rv[k];
        for (int k = threadIdx.x; k < LB_IMAX; k += blockDim.x) dst->iv[k] = red.iv[k];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < R) {
        __threadfence_system();
        *(volatile unsigned long long*)&((P2PBuf<T>*)peers.p[threadIdx.x])->flag[slot][rank] = seq;
    }
}
// formk's entering/leaving corrections of this rank into every peer's delta_all[rank] (only when they are needed:
// do_delta is the same on every rank)
template <typename T>
__global__ void __launch_bounds__(256) k_delta_push(Wk<T> w, const T* delta, Peers peers_delta, Peers peers, int R, int rank,
                                                    unsigned long long dseq) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->do_delta) return;
    const int ne = 6 * LB_MMAX * LB_MMAX;
    for (int q = 0; q < R; ++q) {
        T* dst = (T*)peers_delta.p[q] + (i64)rank * ne;
        for (int e = threadIdx.x; e < ne; e += blockDim.x) dst[e] = delta[e];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < R) {
        __threadfence_system();
        *(volatile unsigned long long*)&((P2PBuf<T>*)peers.p[threadIdx.x])->dflag[rank] = dseq;
    }
}

// Rank-local finish of a site into a global record (sharded runs; followed by an all-gather).
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) k_rank_finish(Wk<T> w, SiteSpec sp, Red<T>* out) {
    __shared__ Red<T> red;
    site_finish_local<T>(w, sp, &red);
    for (int k = threadIdx.x; k < LB_KMAX; k += blockDim.x) out->rv[k] = (k < sp.nreal) ? red.rv[k] : (T)0;
    for (int k = threadIdx.x; k < LB_IMAX; k += blockDim.x) out->iv[k] = (k < sp.nint) ? red.iv[k] : 0;
}

__host__ __device__ inline SiteSpec make_site(int nreal, int nint) {
    SiteSpec sp;
    sp.nreal = nreal; sp.nint = nint; sp.argmin_slot = -1;
    sp.max_from = sp.max_to = sp.min_from = sp.min_to = 0;
    sp.imin_from = sp.imin_to = sp.imax_from = sp.imax_to = 0;
    sp.buf = 0;
    return sp;
}
// the sites
__host__ __device__ inline SiteSpec site_errclb() { SiteSpec s = make_site(0, 2); s.imax_from = 0; s.imax_to = 2; return s; }
__host__ __device__ inline SiteSpec site_active() { return make_site(0, 4); }
__host__ __device__ inline SiteSpec site_projgr() { SiteSpec s = make_site(1, 0); s.max_from = 0; s.max_to = 1; return s; }
__host__ __device__ inline SiteSpec site_cauchy(int mt) {
    SiteSpec s = make_site(2 * mt + 2, 4); s.argmin_slot = 2 * mt + 1; s.imin_from = 3; s.imin_to = 4; s.buf = 1; return s;
}
__host__ __device__ inline SiteSpec site_freev() { return make_site(0, 3); }
__host__ __device__ inline SiteSpec site_formk(int mt) { return make_site(4 * mt, 0); }
__host__ __device__ inline SiteSpec site_wv(int mt) { SiteSpec s = make_site(2 * mt, 0); s.buf = 1; return s; }
__host__ __device__ inline SiteSpec site_subsm() { return make_site(1, 1); }
__host__ __device__ inline SiteSpec site_bt() { SiteSpec s = make_site(1, 1); s.argmin_slot = 0; return s; }
__host__ __device__ inline SiteSpec site_lsinit() { SiteSpec s = make_site(3, 0); s.min_from = 2; s.min_to = 3; s.buf = 1; return s; }
__host__ __device__ inline SiteSpec site_lstrial() { SiteSpec s = make_site(2, 0); s.max_from = 1; s.max_to = 2; return s; }
__host__ __device__ inline SiteSpec site_update(int mt) { return make_site(2 * mt + 1, 0); }
__host__ __device__ inline SiteSpec site_hash() { return make_site(0, 2); }

// The 2m x 2m algebra runs in one thread (reference operation order), and that thread would meet every element of
// the small matrices as a dependent load from global memory.  The block therefore copies them into shared memory
// first (and back afterwards where they were changed): same arithmetic, a fraction of the latency.
template <typename T>
__device__ __forceinline__ void stage_in(T* dst, const T* src, int count) {
    for (int q = threadIdx.x; q < count; q += blockDim.x) dst[q] = src[q];
}
template <typename T>
__device__ __forceinline__ void stage_out(T* dst, const T* src, int count) {
    for (int q = threadIdx.x; q < count; q += blockDim.x) dst[q] = src[q];
}

// "refresh the lbfgs memory" (:625-630 and the four other sites)
template <typename T> __device__ inline void reset_memory(DevState<T>* s) {
    s->info = 0; s->col = 0; s->head = 1; s->theta = (T)1; s->iupdat = 0; s->updatd = 0;
}
// prepare the flags for the prelims block (:601-612)
template <typename T> __device__ inline void begin_body(DevState<T>* s) {
    s->in_body = 1; s->restart = 0; s->need_walk = 0; s->lsinit_done = 0; s->z_in_x = 0; s->save_z = 0;
    s->spec_step = 0; s->step_done = 0; s->do_unstep = 0; s->lazy_gcp = 0; s->fuse_gf = 0; s->lazy_z = 0;
    s->do_subspace = 0; s->do_formk = 0; s->do_delta = 0; s->do_backtrack = 0; s->do_step = 0;
    s->iword = -1;
    ev_push<T>(s, EV_ITER_BEGIN, (T)(s->iter + 1));
    if (!s->cnstnd && s->col > 0) {  // :607-611
        s->cauchy_mode = 1; s->wrk = s->updatd; s->nseg = 0;
    } else if (s->sbgnrm <= (T)0) {  // :1245-1249
        s->cauchy_mode = 2;
        ev_push<T>(s, EV_SUBGNORM0);
    } else s->cauchy_mode = 0;
    s->tsum = (T)0;
}

// ===========================================================================
// The scalar steps of mainlb as single-thread device functions (t0_*: thread 0 of a scalar kernel), shared by the
// general pipeline's kernels (s_*) and the fast pipeline's merged kernels (f_*), so that both produce the same bits.
// ===========================================================================

// clear the per-call flags at the start of every setulb call
template <typename T> __device__ inline void t0_call_begin(DevState<T>* s, T f) {
    s->go = 1; s->pause = 0; s->in_body = 0; s->restart = 0; s->need_walk = 0;
    s->do_step = 0; s->do_restore = 0; s->do_update = 0; s->save_z = 0;
    s->do_subspace = 0; s->do_formk = 0; s->do_delta = 0; s->do_backtrack = 0;
    s->fuse_uc = 0; s->classify_done = 0; s->lsinit_done = 0;
    s->spec_step = 0; s->step_done = 0; s->do_unstep = 0; s->lazy_gcp = 0; s->fuse_gf = 0; s->lazy_z = 0;
    s->ev_n = 0;
    s->f = f;
}

// NEW_X entry, part 1 (:795-834, :838-839, matupd :2303-2309): termination tests, skip rule, ring pointers.
template <typename T> __device__ inline void t0_newx_tests(DevState<T>* s, int fused_supported) {
    const T one = (T)1;
    s->fuse_uc = 0; s->classify_done = 0;
    if (s->sbgnrm <= s->pgtol) { s->task = TK_CONV_PG; s->go = 0; return; }
    T ddum = dense::tmax(dense::tmax(fabs(s->fold), fabs(s->f)), one);
    if ((s->fold - s->f) <= s->tol * ddum) {
        s->task = TK_CONV_F;
        if (s->iback >= 10) s->info = -5;
        s->go = 0;
        return;
    }
    if (s->stp == one) { s->dr = s->gd - s->gdold; s->ddum = -s->gdold; }
    else { s->dr = (s->gd - s->gdold) * s->stp; s->ddum = -s->gdold * s->stp; }
    if (s->dr <= s->epsmch * s->ddum) {
        s->nskip = s->nskip + 1; s->updatd = 0; s->do_update = 0;
        ev_push<T>(s, EV_SKIP, s->dr, s->ddum);
    } else {
        s->updatd = 1; s->iupdat = s->iupdat + 1; s->do_update = 1;
        const int m = s->m;
        if (s->iupdat <= m) { s->col = s->iupdat; s->itail = (s->head + s->iupdat - 2) % m + 1; }
        else { s->itail = s->itail % m + 1; s->head = s->head % m + 1; }
        // begin_body will choose the full per-variable pass of cauchy (cauchy_mode 0) exactly when
        // bounds are present and sbgnrm > 0: that pass is then fused with the update (k_update_classify)
        if (fused_supported && s->cnstnd && s->sbgnrm > (T)0) { s->fuse_uc = 1; s->classify_done = 1; }
    }
}

// NEW_X entry, part 2: matupd's small matrices (:2318-2344) + formt (:1926-1963).  sy, ss, wt: staged copies (the
// shifted entries are read from the originals in the state block).  Called by warp 0.
template <typename T> __device__ inline void w_update_dense(DevState<T>* s, const Red<T>& red, int mt, T* sy, T* ss, T* swt) {
    const int lane = threadIdx.x & 31;
    const int m = s->m, col = s->col;
    if (lane == 0) {
        s->rr = red.rv[0];
        s->theta = s->rr / s->dr;
    }
    if (s->iupdat > m) {  // :2324-2330: every entry moves one place up and to the left
        for (int e = lane; e < m * m; e += 32) {
            const int j = e / m + 1, q = e % m;   // reference loop j = 1..col-1
            if (j > col - 1) continue;
            if (q < j) ss[q + (j - 1) * m] = s->ss[(1 + q) + j * m];                         // dcopy(j,Ss(2,j+1),Ss(1,j))
            if (q < col - j) sy[(j - 1 + q) + (j - 1) * m] = s->sy[(j + q) + j * m];         // dcopy(col-j,Sy(j+1,j+1),Sy(j,j))
        }
    }
    __syncwarp();
    for (int j = 1 + lane; j <= col - 1; j += 32) {
        sy[(col - 1) + (j - 1) * m] = red.rv[1 + (j - 1)];
        ss[(j - 1) + (col - 1) * m] = red.rv[1 + mt + (j - 1)];
    }
    if (lane == 0) {
        ss[(col - 1) + (col - 1) * m] = (s->stp == (T)1) ? s->dtd : s->stp * s->stp * s->dtd;
        sy[(col - 1) + (col - 1) * m] = s->dr;
    }
    __syncwarp();
    int info = wdense::formt<T>(m, swt, sy, ss, col, s->theta);
    if (info != 0 && lane == 0) { ev_push<T>(s, EV_FORMT_FAIL); reset_memory<T>(s); }   // :851-863
    __syncwarp();
}

// cauchy after the per-variable pass (:1337-1366) and the first-segment exit test (:1384-1416 with iter == 1).
// ssy, swt: staged copies of sy, wt.  Called by warp 0.
template <typename T> __device__ inline void w_cauchy(DevState<T>* s, const Red<T>& red, int mt, const T* ssy, const T* swt,
                                                      int fused_supported) {
    const int lane = threadIdx.x & 31;
    const int col = s->col, col2 = 2 * col, m = s->m;
    const T zero = (T)0, one = (T)1;
    int stop = 0;
    if (lane == 0) {
        s->classify_done = 0;
        for (int j = 0; j < col; ++j) { s->p[j] = red.rv[j]; s->p[col + j] = red.rv[mt + j]; }
        s->f1 = -red.rv[2 * mt];
        s->bkmin = red.rv[2 * mt + 1];
        s->ibkmin = red.iv[0];
        s->nbreak = red.iv[1];
        s->nfreec = red.iv[2];
        s->bnded = red.iv[3] > 0;
        if (s->theta != one) for (int j = 0; j < col; ++j) s->p[col + j] = s->theta * s->p[col + j];  // :1337
        s->tsum = zero;
        s->lazy_gcp = 1;   // the per-variable pass did not write d and xcp = x (k_gcp_freev / k_formk_cmprlb form xcp)
        if (s->nbreak == 0 && s->nfreec == 0) stop = 1;   // d is the zero vector (:1343-1347); nseg untouched
        else {
            for (int j = 0; j < col2; ++j) s->c[j] = zero;
            s->f2 = -s->theta * s->f1;
            s->f2_org = s->f2;
        }
    }
    stop = __shfl_sync(LB_FULL, stop, 0);
    if (stop) return;
    __syncwarp();
    if (col > 0) {
        int info = wdense::bmv<T>(m, ssy, swt, col, s->p, s->v);
        if (info != 0) {   // :620-635
            if (lane == 0) {
                ev_push<T>(s, EV_CAUCHY_SINGULAR);
                reset_memory<T>(s);
                s->restart = 1; s->in_body = 0;
            }
            __syncwarp();
            return;
        }
    }
    if (lane == 0) {
        if (col > 0) s->f2 = s->f2 - dense::ddot<T>(col2, s->v, s->p);
        s->dtm = -s->f1 / s->f2;
        s->nseg = 1;
        if (s->nbreak != 0 && !(s->dtm < s->bkmin)) {
            // the first breakpoint is reached: the sorted walk takes over (its first count pass writes d, xcp = x first)
            s->need_walk = 1; s->lazy_gcp = 0;
            for (int j = 0; j < col2; ++j) { s->p0[j] = s->p[j]; s->walkA[j] = zero; s->walkB[j] = zero; }
            s->walk_f1 = s->f1; s->walk_f2 = s->f2; s->walk_tlast = zero; s->walk_tprev2 = zero;
            s->walk_J = -1; s->walk_done = 0; s->walk_base = 0; s->walk_rcount = 0; s->walk_lcount = 0; s->walk_rem = 0;
            s->walk_cstart = 0; s->walk_fixn = 0; s->walk_closed = 0; s->tie_redo = 0; s->tie_round = 0;
            stop = 1;
        } else {
            if (s->dtm <= zero) s->dtm = zero;   // :1509
            s->tsum = s->tsum + s->dtm;
            if (col > 0) dense::daxpy<T>(col2, s->dtm, s->p, s->c);  // :1526
        }
    }
    stop = __shfl_sync(LB_FULL, stop, 0);
    if (stop) return;
    __syncwarp();
    // The generalized Cauchy point is known without a walk and c is final: cmprlb's a = M c (:1569) can be
    // formed here, and then the tail of cauchy (:1515) and freev run inside k_formk_cmprlb (fuse_gf).  The
    // counts of freev, and with them nfree == 0 / wrk, are evaluated after that pass (s_freev phase 1).  If
    // the product fails, the separate passes run and s_freev reports the failure in the reference's order.
    if (fused_supported && col > 0 && s->cnstnd) {
        const int info = wdense::bmv<T>(m, ssy, swt, col, s->c, s->a);
        if (info == 0 && lane == 0) s->fuse_gf = 1;
    }
    __syncwarp();
}

// after freev (:638-648): counters, wrk, what of the subspace phase runs.  `have` = the counts exist (cauchy_mode != 1).
// Called by warp 0.
template <typename T> __device__ inline void w_freev(DevState<T>* s, const Red<T>& red, i64 n_global, bool gf, bool have) {
    const int lane = threadIdx.x & 31;
    int need_a = 0;
    if (lane == 0) {
        if (s->fuse_gf == 1) s->lazy_z = 1;   // k_formk_cmprlb did not store xcp (state bit 2 tells where d = -g)
        if (have) {
            s->nintol = s->nintol + s->nseg;
            s->nfree = red.iv[0];
            s->nenter = red.iv[1];
            s->nleave = red.iv[2];
            s->wrk = (s->nleave > 0) || (s->nenter > 0) || s->updatd;
            s->nact = n_global - s->nfree;
        }
        s->do_subspace = !(s->nfree == 0 || s->col == 0);
        s->do_formk = s->do_subspace && s->wrk;
        s->do_delta = s->do_formk && (s->nenter + s->nleave > 0);
        // cmprlb's a = M c (:1569; not needed on the unconstrained shortcut :1560-1563).  It depends only on
        // sy, wt and c, which are final here, and is hoisted in front of formk so that formk's Gram pass and
        // cmprlb's pass over S/Y can run as one kernel.  A failure of either ends in the same memory reset.
        need_a = (!gf && s->do_subspace && !(!s->cnstnd && s->col > 0)) ? 1 : 0;   // (with fuse_gf s_cauchy formed a)
    }
    need_a = __shfl_sync(LB_FULL, need_a, 0);
    __syncwarp();
    if (need_a) {
        int info = wdense::bmv<T>(s->m, s->sy, s->wt, s->col, s->c, s->a);
        if (info != 0 && lane == 0) {   // info = -8 -> :694-710
            ev_push<T>(s, EV_SUBSM_SINGULAR);
            reset_memory<T>(s);
            s->restart = 1; s->in_body = 0; s->do_subspace = 0; s->do_formk = 0; s->do_delta = 0;
        }
    }
    __syncwarp();
}

// formk, everything after the long sums (:1735-1744 shift, :1772-1792 new row/column, :1821-1847 corrections,
// :1853-1906 assembly and factorisation).  wn, wn1: staged copies (the shifted entries are read from the original
// wn1 in the state block).  Called by warp 0.
// delta: [3][2m][2m] enter-minus-leave sums of Wy.Wy, Ws.Ws, Ws.Wy over ring positions
//        (dE - dL kept separately: delta[0..2] enter, delta[3..5] leave)
template <typename T> __device__ inline void w_formk_dense(DevState<T>* s, const Red<T>& red, int mt, const T* delta, T* wn, T* wn1) {
    const int lane = threadIdx.x & 31;
    const int m = s->m, col = s->col, m2 = 2 * m;
#define WN(i, j) wn[((i)-1) + ((j)-1) * m2]
#define WN1(i, j) wn1[((i)-1) + ((j)-1) * m2]
#define WN1O(i, j) s->wn1[((i)-1) + ((j)-1) * m2]
    if (s->do_formk) {
        int upcl;
        if (s->updatd) {
            if (s->iupdat > m) {   // :1736-1744: the three blocks move one place up and to the left
                for (int e = lane; e < (m - 1) * m; e += 32) {
                    const int jy = e / m + 1, q = e % m, js = m + jy;
                    if (q < m - jy) { WN1(jy + q, jy) = WN1O(jy + 1 + q, jy + 1); WN1(js + q, js) = WN1O(js + 1 + q, js + 1); }
                    if (q < m - 1) WN1(m + 1 + q, jy) = WN1O(m + 2 + q, jy + 1);
                }
                __syncwarp();
            }
            const int iy = col, is = m + col;
            for (int jy = 1 + lane; jy <= col; jy += 32) {   // :1772-1774
                const int js = m + jy;
                WN1(iy, jy) = red.rv[jy - 1];
                WN1(is, js) = red.rv[mt + jy - 1];
                WN1(is, jy) = red.rv[2 * mt + jy - 1];
            }
            __syncwarp();
            for (int i = 1 + lane; i <= col; i += 32) WN1(m + i, col) = red.rv[3 * mt + i - 1];   // :1792
            __syncwarp();
            upcl = col - 1;
        } else upcl = col;
        if (s->do_delta) {
            // delta layout: block b in 0..5, entry (i,j) ring positions, leading dimension LB_MMAX
            const T* eYY = delta + 0 * LB_MMAX * LB_MMAX; const T* eSS = delta + 1 * LB_MMAX * LB_MMAX;
            const T* eSY = delta + 2 * LB_MMAX * LB_MMAX; const T* lYY = delta + 3 * LB_MMAX * LB_MMAX;
            const T* lSS = delta + 4 * LB_MMAX * LB_MMAX; const T* lSY = delta + 5 * LB_MMAX * LB_MMAX;
            for (int e = lane; e < upcl * upcl; e += 32) {
                const int a = e / upcl + 1, b = e % upcl + 1;
                if (b <= a) {   // :1802-1826 with (iy, jy) = (a, b)
                    const int iy = a, jy = b, is = m + iy, js = m + jy;
                    const int d = (iy - 1) + (jy - 1) * LB_MMAX;
                    WN1(iy, jy) = WN1(iy, jy) + eYY[d] - lYY[d];
                    WN1(is, js) = WN1(is, js) - eSS[d] + lSS[d];
                }
                {               // :1830-1851 with (is, jy) = (m + a, b)
                    const int is = m + a, jy = b;
                    const int d = (is - m - 1) + (jy - 1) * LB_MMAX;
                    if (is <= jy + m) WN1(is, jy) = WN1(is, jy) + eSY[d] - lSY[d];
                    else WN1(is, jy) = WN1(is, jy) - eSY[d] + lSY[d];
                }
            }
            __syncwarp();
        }
        const T theta = s->theta;
        for (int e = lane; e < col * col; e += 32) {   // :1857-1873, one (iy, jy) pair per lane
            const int iy = e / col + 1, jy = e % col + 1;
            const int is = col + iy, is1 = m + iy, js = col + jy, js1 = m + jy;
            if (jy <= iy) {
                T v = WN1(iy, jy) / theta;
                if (jy == iy) v = v + s->sy[(iy - 1) + (iy - 1) * m];
                WN(jy, iy) = v;
                WN(js, is) = WN1(is1, js1) * theta;
            }
            if (jy <= iy - 1) WN(jy, is) = -WN1(is1, jy);
            else WN(jy, is) = WN1(is1, jy);
        }
        __syncwarp();
        int info = wdense::dpofa<T>(wn, m2, col);
        if (info != 0) info = -1;
        else {
            const int col2 = 2 * col;
            for (int js = col + 1 + lane; js <= col2; js += 32) dense::dtrsl<T>(wn, m2, col, &WN(1, js), 11);
            __syncwarp();
            for (int e = lane; e < col * col; e += 32) {
                const int is = col + 1 + e / col, js = col + 1 + e % col;
                if (js >= is) WN(is, js) = WN(is, js) + dense::ddot<T>(col, &WN(1, is), &WN(1, js));
            }
            __syncwarp();
            info = wdense::dpofa<T>(&WN(col + 1, col + 1), m2, col);
            if (info != 0) info = -2;
        }
        if (info != 0 && lane == 0) {   // :666-682
            ev_push<T>(s, EV_FORMK_FAIL);
            reset_memory<T>(s);
            s->restart = 1; s->in_body = 0;
        }
        __syncwarp();
    }
#undef WN
#undef WN1
#undef WN1O
}

// subsm: wv = K^{-1} wv (:2751-2766).  swn: staged copy of wn.  Called by warp 0.
template <typename T> __device__ inline void w_subsm_dense(DevState<T>* s, const Red<T>& red, int mt, const T* swn, int fused_lsinit) {
    const int lane = threadIdx.x & 31;
    const int m = s->m, col = s->col, m2 = 2 * m, col2 = 2 * col;
    for (int i = lane; i < col; i += 32) { s->wv[i] = red.rv[i]; s->wv[col + i] = s->theta * red.rv[mt + i]; }
    __syncwarp();
    int info = wdense::dtrsl<T>(swn, m2, col2, s->wv, 11);
    if (info == 0) {
        for (int i = lane; i < col; i += 32) s->wv[i] = -s->wv[i];
        __syncwarp();
        info = wdense::dtrsl<T>(swn, m2, col2, s->wv, 1);
    }
    if (lane == 0) {
        if (info != 0) {   // :694-710
            ev_push<T>(s, EV_SUBSM_SINGULAR);
            reset_memory<T>(s);
            s->restart = 1; s->in_body = 0; s->do_subspace = 0;
        } else {
            // lnsrlb starts at stp = 1 unless this is the first iteration of a problem that is not boxed (:2229-2233):
            // the fused subspace pass then writes that trial point (x = z, :2265) itself
            s->spec_step = (fused_lsinit && (s->iter != 0 || s->boxed)) ? 1 : 0;
        }
    }
    __syncwarp();
}

// subsm: projection outcome (:2820-2828)
template <typename T> __device__ inline void t0_subsm_post(DevState<T>* s, const Red<T>& red, int fused_lsinit) {
    s->iword = red.iv[0] > 0 ? 1 : 0;
    s->dd_p = red.rv[0];
    s->do_backtrack = (s->iword == 1 && s->dd_p > (T)0);
    if (s->do_backtrack) ev_push<T>(s, EV_BACKTRACK);
    // the subspace pass also formed d = z - x and lnsrlb's sums; they stand unless the backtrack moves z
    s->lsinit_done = (fused_lsinit && !s->do_backtrack) ? 1 : 0;
}

// lnsrlb first entry (:2196-2273) and mainlb's handling of its outcome (:734-773)
template <typename T>
__device__ inline void ls_failure(DevState<T>* s, bool at_first_entry) {
    // :734-769 ; x = t, g = r, f = fold
    if (!at_first_entry) s->do_restore = 1;   // at the first entry x, g are still the old iterate ...
    else if (s->spec_step) s->do_unstep = 1;  // ... unless the subspace pass already stepped: x = t again
    s->do_step = 0;
    s->f = s->fold;
    if (s->col == 0) {
        if (s->info == 0) { s->info = -9; s->nfgv -= 1; s->ifun -= 1; s->iback -= 1; }
        s->task = TK_ABNORMAL;
        s->iter += 1;
        s->go = 0; s->in_body = 0;
    } else {
        ev_push<T>(s, EV_LNSRCH_RESTART);
        if (s->info == 0) s->nfgv -= 1;
        reset_memory<T>(s);
        s->task = TK_RESTART;
        s->restart = 1; s->in_body = 0;
    }
}
template <typename T> __device__ inline void t0_ls_init(DevState<T>* s, const Red<T>& red) {
    const T zero = (T)0, one = (T)1, big = (T)1.0e+10, ftol = (T)1.0e-3, gtol = (T)0.9, xtol = (T)0.1;
    s->dtd = red.rv[0];
    s->dnorm = sqrt(s->dtd);
    s->stpmx = big;
    if (s->cnstnd) {
        if (s->iter == 0) s->stpmx = one;
        else s->stpmx = dense::tmin(big, red.rv[2]);
    }
    if (s->iter == 0 && !s->boxed) s->stp = dense::tmin(one / s->dnorm, s->stpmx);
    else s->stp = one;
    s->fold = s->f;
    s->ifun = 0; s->iback = 0; s->csave = CS_START;
    s->gd = red.rv[1];
    s->gdold = s->gd;
    if (s->gd >= zero) {   // :2247-2253
        ev_push<T>(s, EV_ASCENT, s->gd);
        s->info = -4;
        ls_failure<T>(s, true);
        return;
    }
    dense::dcsrch<T>(s->f, s->gd, s->stp, ftol, gtol, xtol, zero, s->stpmx, s->csave, s->brackt, s->stage, s->ls);
    s->xstep = s->stp * s->dnorm;
    if (s->csave != CS_CONV && !cs_is_warn(s->csave)) {
        s->task = TK_FG_LNSRCH;
        s->ifun += 1; s->nfgv += 1; s->iback = s->ifun - 1;
        s->do_step = 1;
        s->step_done = (s->spec_step && s->lsinit_done && s->stp == one) ? 1 : 0;   // x already holds z
        s->z_in_x = s->step_done;   // ... and z itself was not stored
        s->go = 0; s->in_body = 0;   // return to the caller for f and g
    } else {
        // cannot happen on the first entry (dcsrch 'START' returns 'FG' or 'ERROR')
        s->task = TK_NEW_X; s->iter += 1; s->go = 0; s->in_body = 0;
    }
}

// lnsrlb re-entry (:2244-2273) and the outcome (:734-788)
template <typename T> __device__ inline void t0_ls_trial(DevState<T>* s, const Red<T>& red) {
    const T zero = (T)0, ftol = (T)1.0e-3, gtol = (T)0.9, xtol = (T)0.1;
    s->gd = red.rv[0];
    if (s->ifun == 0) {   // only after a dcsrch input error on the first entry; kept for fidelity
        s->gdold = s->gd;
        if (s->gd >= zero) { ev_push<T>(s, EV_ASCENT, s->gd); s->info = -4; ls_failure<T>(s, false); return; }
    }
    dense::dcsrch<T>(s->f, s->gd, s->stp, ftol, gtol, xtol, zero, s->stpmx, s->csave, s->brackt, s->stage, s->ls);
    s->xstep = s->stp * s->dnorm;
    bool more = (s->csave != CS_CONV && !cs_is_warn(s->csave));
    if (more) {
        s->task = TK_FG_LNSRCH;
        s->ifun += 1; s->nfgv += 1; s->iback = s->ifun - 1;
    } else s->task = TK_NEW_X;
    if (s->info != 0 || s->iback >= 20) {
        ls_failure<T>(s, false);
        return;
    }
    if (more) {
        s->do_step = 1; s->go = 0;
        if (s->z_in_x) {   // the Newton point lives only in x (speculative step)
            if (s->stp == (T)1) s->step_done = 1;            // asked for again (:2265): x is already there
            else { s->save_z = 1; s->z_in_x = 0; }           // x moves on: k_ls_step keeps a copy in z first
        }
        return;
    }
    // accepted: NEW_X (:775-787)
    s->iter += 1;
    s->sbgnrm = red.rv[1];
    s->go = 0;
}

// ===========================================================================
// General pipeline: one scalar kernel per reduction site.
// ===========================================================================

// START, part 1: errclb (:1601-1643) result.
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_errclb(Wk<T> w, Dist<T> dist, i64 index_offset) {
    __shared__ Red<T> red;
    site_fetch<T>(w, dist);
    site_reduce<T>(w, dist, site_errclb(), &red);
    if (threadIdx.x != 0) return;
    DevState<T>* s = w.s;
    i64 k6 = red.iv[0], k7 = red.iv[1];
    // Each offending i overwrites (task, info, k); the last one wins.  On a shard the
    // indices are local, so they are globalised by the caller through index_offset.
    if (dist.R <= 1) { if (k6 >= 0) k6 += index_offset; if (k7 >= 0) k7 += index_offset; }
    if (k6 >= 0 || k7 >= 0) {
        if (k6 > k7) { s->task = TK_ERR_NBD; s->info = -6; s->errk = k6 + 1; }
        else { s->task = TK_ERR_INFEAS; s->info = -7; s->errk = k7 + 1; }
        s->go = 0;
    } else if (s->task >= TK_ERR_N) s->go = 0;   // factr < 0 stands when no array error overwrites it
}

// START, part 2: active (:965-1040) flags; then start() (:884-890).
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_active(Wk<T> w, Dist<T> dist) {
    __shared__ Red<T> red;
    site_fetch<T>(w, dist);
    if (!w.s->go) return;
    site_reduce<T>(w, dist, site_active(), &red);
    if (threadIdx.x != 0) return;
    DevState<T>* s = w.s;
    s->nbdd = red.iv[0];
    s->prjctd = red.iv[1] > 0; s->cnstnd = red.iv[2] > 0; s->boxed = !(red.iv[3] > 0);
    s->task = TK_FG_START;
    s->go = 0;
}

// FG_START entry (:579-596): sbgnrm, first termination test, open the body.
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_fg_start(Wk<T> w, Dist<T> dist) {
    __shared__ Red<T> red;
    site_fetch<T>(w, dist);
    site_reduce<T>(w, dist, site_projgr(), &red);
    if (threadIdx.x != 0) return;
    DevState<T>* s = w.s;
    s->nfgv = 1;
    s->sbgnrm = red.rv[0];
    if (s->sbgnrm <= s->pgtol) { s->task = TK_CONV_PG; s->go = 0; return; }
    begin_body<T>(s);
}

template <typename T>
__global__ void s_newx_tests(Wk<T> w, int fused_supported) {
    if (threadIdx.x != 0) return;
    t0_newx_tests<T>(w.s, fused_supported);
}

// NEW_X entry, part 2: matupd's small matrices + formt, then open the body.
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_update_dense(Wk<T> w, Dist<T> dist, int mt) {
    __shared__ Red<T> red;
    __shared__ T ssy[LB_MMAX * LB_MMAX], sss[LB_MMAX * LB_MMAX], swt[LB_MMAX * LB_MMAX];
    DevState<T>* s = w.s;
    site_fetch<T>(w, dist);
    if (!s->go) return;
    const bool upd = s->do_update;
    const int mm = s->m * s->m;
    if (upd) {
        site_reduce<T>(w, dist, site_update(mt), &red);
        stage_in<T>(ssy, s->sy, mm); stage_in<T>(sss, s->ss, mm); stage_in<T>(swt, s->wt, mm);
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        if (upd) w_update_dense<T>(s, red, mt, ssy, sss, swt);
        if (threadIdx.x == 0) begin_body<T>(s);
    }
    if (upd) {
        __syncthreads();
        stage_out<T>(s->sy, ssy, mm); stage_out<T>(s->ss, sss, mm); stage_out<T>(s->wt, swt, mm);
    }
}

// Host asked for another pass of the body after a memory reset ("cycle main_loop").
template <typename T>
__global__ void s_restart_body(Wk<T> w) {
    if (threadIdx.x != 0) return;
    DevState<T>* s = w.s;
    s->go = 1; s->pause = 0;
    s->classify_done = 0;
    begin_body<T>(s);
}
// The host resumes the general pipeline after a pause of the fast one.
template <typename T>
__global__ void s_resume(Wk<T> w) {
    if (threadIdx.x == 0) w.s->pause = 0;
}

template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_cauchy(Wk<T> w, Dist<T> dist, int mt, int fused_supported) {
    __shared__ Red<T> red;
    __shared__ T ssy[LB_MMAX * LB_MMAX], swt[LB_MMAX * LB_MMAX];
    DevState<T>* s = w.s;
    site_fetch<T>(w, dist);
    if (!s->go || !s->in_body || s->cauchy_mode != 0) return;
    site_reduce<T>(w, dist, site_cauchy(mt), &red);
    stage_in<T>(ssy, s->sy, s->m * s->m); stage_in<T>(swt, s->wt, s->m * s->m);
    __syncthreads();
    if (threadIdx.x >= 32) return;
    w_cauchy<T>(s, red, mt, ssy, swt, fused_supported);
}

// After a breakpoint walk has closed: c is final, so cmprlb's a = M c (:1569) can be formed here as s_cauchy forms it when
// no walk is needed, and the tail of cauchy and freev then run inside k_formk_cmprlb as well (fuse_gf = 2: the Cauchy point
// starts from the stored xcp, which holds the bounds of the variables the walk fixed).
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_walk_gf(Wk<T> w, int fused_supported) {
    __shared__ T ssy[LB_MMAX * LB_MMAX], swt[LB_MMAX * LB_MMAX];
    DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || !s->walk_closed || s->cauchy_mode != 0) return;
    if (!fused_supported || s->col <= 0 || !s->cnstnd || s->fuse_gf) return;
    stage_in<T>(ssy, s->sy, s->m * s->m); stage_in<T>(swt, s->wt, s->m * s->m);
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int info = wdense::bmv<T>(s->m, ssy, swt, s->col, s->c, s->a);
    if (info == 0 && (threadIdx.x & 31) == 0) s->fuse_gf = 2;
}

template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_freev(Wk<T> w, Dist<T> dist, i64 n_global, int phase) {
    __shared__ Red<T> red;
    DevState<T>* s = w.s;
    site_fetch<T>(w, dist);
    if (!s->go || !s->in_body) return;
    // phase 0: after k_gcp_freev (the separate pass).  With fuse_gf the counts do not exist yet: the subspace
    // flags are set tentatively (col > 0 holds; the Gram row is wanted iff updatd) and phase 1, after
    // k_formk_cmprlb, evaluates them.  phase 1 does nothing without fuse_gf.
    const bool gf = s->fuse_gf != 0;
    if (phase == 1 && !gf) return;
    if (phase == 0 && gf) {
        if (threadIdx.x == 0) { s->do_subspace = 1; s->do_formk = s->updatd; s->do_delta = 0; }
        return;
    }
    const int mode = s->cauchy_mode;
    if (mode != 1) site_reduce<T>(w, dist, site_freev(), &red);
    if (threadIdx.x >= 32) return;
    w_freev<T>(s, red, n_global, gf, mode != 1);
}

template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_formk_dense(Wk<T> w, Dist<T> dist, int mt, const T* delta, T* delta_sum) {
    __shared__ Red<T> red;
    DevState<T>* s = w.s;
    site_fetch<T>(w, dist);
    if (!s->go || !s->in_body || !s->do_subspace) return;
    const bool newrow = s->do_formk && s->updatd;
    if (newrow) site_reduce<T>(w, dist, site_formk(mt), &red);
    if (dist.R > 1 && s->do_delta) {
        // sharded: `delta` holds the per-rank corrections [R][6*MMAX*MMAX] (all-gathered, or pushed by the peers);
        // add them in rank order
        if (dist.p2p) p2p_wait<T>(w, dist.p2p->dflag, dist.R, dist.dseq);
        const volatile T* dv = delta;
        for (int e = threadIdx.x; e < 6 * LB_MMAX * LB_MMAX; e += blockDim.x) {
            T acc = dv[e];
            for (int q = 1; q < dist.R; ++q) acc = acc + dv[(i64)q * (6 * LB_MMAX * LB_MMAX) + e];
            delta_sum[e] = acc;
        }
        __syncthreads();
        delta = delta_sum;
    }
    __shared__ T swn[4 * LB_MMAX * LB_MMAX], swn1[4 * LB_MMAX * LB_MMAX];
    const bool stage = s->do_formk != 0;
    if (stage) {
        stage_in<T>(swn, s->wn, 4 * s->m * s->m); stage_in<T>(swn1, s->wn1, 4 * s->m * s->m);
        __syncthreads();
    }
    if (threadIdx.x < 32) w_formk_dense<T>(s, red, mt, delta, swn, swn1);
    if (stage) {
        __syncthreads();
        stage_out<T>(s->wn, swn, 4 * s->m * s->m); stage_out<T>(s->wn1, swn1, 4 * s->m * s->m);
    }
}

template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_subsm_dense(Wk<T> w, Dist<T> dist, int mt, int fused_lsinit) {
    __shared__ Red<T> red;
    DevState<T>* s = w.s;
    site_fetch<T>(w, dist);
    if (!s->go || !s->in_body || !s->do_subspace) return;
    site_reduce<T>(w, dist, site_wv(mt), &red);
    __shared__ T swn[4 * LB_MMAX * LB_MMAX];
    stage_in<T>(swn, s->wn, 4 * s->m * s->m);
    __syncthreads();
    if (threadIdx.x >= 32) return;
    w_subsm_dense<T>(s, red, mt, swn, fused_lsinit);
}

template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_subsm_post(Wk<T> w, Dist<T> dist, int fused_lsinit) {
    __shared__ Red<T> red;
    DevState<T>* s = w.s;
    site_fetch<T>(w, dist);
    if (!s->go || !s->in_body || !s->do_subspace) return;
    site_reduce<T>(w, dist, site_subsm(), &red);
    if (threadIdx.x != 0) return;
    t0_subsm_post<T>(s, red, fused_lsinit);
}

// subsm: backtrack step length (:2836-2863)
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_bt(Wk<T> w, Dist<T> dist, i64 index_offset) {
    __shared__ Red<T> red;
    DevState<T>* s = w.s;
    site_fetch<T>(w, dist);
    if (!s->go || !s->in_body || !s->do_backtrack) return;
    site_reduce<T>(w, dist, site_bt(), &red);
    if (threadIdx.x != 0) return;
    T alpha = (T)1; i64 ibd = -1;
    if (red.rv[0] < alpha) { alpha = red.rv[0]; ibd = red.iv[0]; }
    s->alpha = alpha;
    s->ibd = ibd;   // global variable index (k_bt_alpha adds the shard offset)
    s->lazy_z = 0;  // k_bt_apply stores the backtracked point in z
    (void)index_offset;
}

template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_ls_init(Wk<T> w, Dist<T> dist) {
    __shared__ Red<T> red;
    DevState<T>* s = w.s;
    site_fetch<T>(w, dist);
    if (!s->go || !s->in_body) return;
    site_reduce<T>(w, dist, site_lsinit(), &red);
    if (threadIdx.x != 0) return;
    t0_ls_init<T>(s, red);
}

// lnsrlb re-entry.  call_begin != 0: this kernel opens the setulb call itself (the FG_LNSRCH entry launches nothing
// in front of it that reads the per-call flags).
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) s_ls_trial(Wk<T> w, Dist<T> dist, int call_begin, T f) {
    __shared__ Red<T> red;
    DevState<T>* s = w.s;
    site_fetch<T>(w, dist);
    site_reduce<T>(w, dist, site_lstrial(), &red);
    if (threadIdx.x != 0) return;
    if (call_begin) t0_call_begin<T>(s, f);
    t0_ls_trial<T>(s, red);
}

template <typename T>
__global__ void s_call_begin(Wk<T> w, T f, int entry_task) {
    if (threadIdx.x != 0) return;
    t0_call_begin<T>(w.s, f);
    (void)entry_task;
}

// START: initialise the state block (:436-480)
template <typename T>
__global__ void s_start(Wk<T> w, T factr, T pgtol, int host_err_task) {
    if (threadIdx.x != 0) return;
    DevState<T>* s = w.s;
    const T zero = (T)0;
    s->go = 1; s->pause = 0; s->in_body = 0; s->restart = 0; s->need_walk = 0; s->cauchy_mode = 0;
    s->do_subspace = s->do_formk = s->do_delta = s->do_backtrack = s->do_update = s->do_step = s->do_restore = 0;
    s->z_in_x = 0; s->save_z = 0;
    s->fuse_uc = 0; s->classify_done = 0; s->lsinit_done = 0; s->ev_n = 0;
    s->spec_step = 0; s->step_done = 0; s->do_unstep = 0; s->lazy_gcp = 0; s->fuse_gf = 0; s->lazy_z = 0;
    s->task = TK_START; s->csave = CS_BLANK; s->info = 0;
    s->col = 0; s->head = 1; s->theta = (T)1; s->iupdat = 0; s->updatd = 0;
    s->iback = 0; s->itail = 0; s->iword = 0; s->nact = 0; s->nleave = 0; s->nenter = 0;
    s->fold = zero; s->dnorm = zero; s->gd = zero; s->stpmx = zero; s->sbgnrm = zero; s->stp = zero;
    s->gdold = zero; s->dtd = zero; s->xstep = zero;
    s->iter = 0; s->nfgv = 0; s->nseg = 0; s->nintol = 0; s->nskip = 0; s->nfree = w.n; s->ifun = 0;
    s->epsmch = Real<T>::eps();
    s->tol = factr * s->epsmch;
    s->pgtol = pgtol; s->factr = factr;
    s->brackt = 0; s->stage = 0;
    s->prjctd = s->cnstnd = s->boxed = 0; s->wrk = 0; s->bnded = 0;
    s->errk = 0; s->nbdd = 0; s->n = w.n; s->m = w.m;
    s->nbreak = 0; s->nfreec = 0; s->ibkmin = 0; s->ibd = -1; s->walk_J = -1; s->walk_done = 0; s->n_el = 0;
    s->tie_redo = 0; s->tie_round = 0; s->tie_events = 0;
    for (int q = 0; q < 13; ++q) s->ls[q] = zero;
    s->f = zero; s->rr = zero; s->dr = zero; s->ddum = zero; s->tsum = zero; s->dtm = zero;
    if (host_err_task != 0) { s->task = host_err_task; }   // n<=0 / m<=0 / factr<0 (:1618-1620); array errors overwrite
}

// ===========================================================================
// Fast pipeline (Engine::fast_newx): the common path of one NEW_X entry -- bounds present, S/Y update not skipped,
// the Cauchy search ends in its first segment, no variable enters or leaves the free set, no backtrack -- is
//     f_head -> k_update_classify -> f_ucf -> k_formk_cmprlb<gf> -> f_mid -> k_subsm_lsinit -> f_tail
// three streaming passes and four scalar kernels, each scalar kernel consuming ONE merged record (MSite) on a
// sharded run.  A merged kernel that meets any other branch leaves the state exactly as the general pipeline
// would have it at that point, sets s->pause to the stage at which the host must resume it (PAUSE_*), and
// everything enqueued after it returns at once.
// ===========================================================================
__host__ __device__ inline MSite msite_ucf(int mt) {
    MSite ms; ms.nsub = 2;
    ms.sp[0] = site_update(mt); ms.roff[0] = 0; ms.ioff[0] = 0;
    ms.sp[1] = site_cauchy(mt); ms.roff[1] = 2 * mt + 1; ms.ioff[1] = 0;
    ms.sp[2] = make_site(0, 0); ms.roff[2] = 0; ms.ioff[2] = 0;
    return ms;
}
__host__ __device__ inline MSite msite_mid(int mt) {
    MSite ms; ms.nsub = 3;
    ms.sp[0] = site_freev(); ms.roff[0] = 0; ms.ioff[0] = 0;
    ms.sp[1] = site_formk(mt); ms.roff[1] = 0; ms.ioff[1] = 3;
    ms.sp[2] = site_wv(mt); ms.roff[2] = 4 * mt; ms.ioff[2] = 3;
    return ms;
}
__host__ __device__ inline MSite msite_tail() {
    MSite ms; ms.nsub = 2;
    ms.sp[0] = site_subsm(); ms.roff[0] = 0; ms.ioff[0] = 0;
    ms.sp[1] = site_lsinit(); ms.roff[1] = 1; ms.ioff[1] = 1;
    ms.sp[2] = make_site(0, 0); ms.roff[2] = 0; ms.ioff[2] = 0;
    return ms;
}

// NEW_X entry: s_call_begin + s_newx_tests.
template <typename T>
__global__ void f_head(Wk<T> w, T f, int fused_supported) {
    if (threadIdx.x != 0) return;
    t0_call_begin<T>(w.s, f);
    t0_newx_tests<T>(w.s, fused_supported);
}

// ===========================================================================
// Device-resident iteration (Engine::minimize_graph): one CUDA graph holds the caller's objective kernels, the FG_LNSRCH
// entry and the NEW_X entry of the fast pipeline.  The two entry kernels below open their setulb call only if the state
// block asks for exactly that call -- what the caller's loop (test/driver1.f90:263-292) decides from `task` on the host --
// and take f from device memory; otherwise everything enqueued behind them returns at once.
// ===========================================================================
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) g_ls_trial(Wk<T> w, Dist<T> dist, const T* f_dev) {
    __shared__ Red<T> red;
    DevState<T>* s = w.s;
    if (s->task != TK_FG_LNSRCH) return;
    site_reduce<T>(w, dist, site_lstrial(), &red);
    if (threadIdx.x != 0) return;
    t0_call_begin<T>(s, *f_dev);
    s->gstage = 1;
    t0_ls_trial<T>(s, red);
}
// NEW_X entry, unless the caller's limits say STOP (test/driver2.f90:174-181: the host then issues it through setulb)
template <typename T>
__global__ void g_head(Wk<T> w, int fused_supported, int max_iter, int max_fg) {
    if (threadIdx.x != 0) return;
    DevState<T>* s = w.s;
    const bool limit = (max_iter > 0 && s->iter >= max_iter) || (max_fg > 0 && s->nfgv >= max_fg);
    if (s->task != TK_NEW_X || limit) { s->go = 0; return; }   // (a restart of the line search left go = 1: the host resumes it)
    t0_call_begin<T>(s, s->f);
    s->gstage = 2;
    t0_newx_tests<T>(s, fused_supported);
}

// s_update_dense + s_cauchy + s_freev(phase 0 with fuse_gf).
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) f_ucf(Wk<T> w, Dist<T> dist, int mt) {
    __shared__ Red<T> red;
    __shared__ T ssy[LB_MMAX * LB_MMAX], sss[LB_MMAX * LB_MMAX], swt[LB_MMAX * LB_MMAX];
    __shared__ int cont;
    DevState<T>* s = w.s;
    const MSite ms = msite_ucf(mt);
    site_fetch<T>(w, dist);
    if (!s->go) return;
    const bool upd = s->do_update;
    const int mm = s->m * s->m;
    stage_in<T>(ssy, s->sy, mm); stage_in<T>(swt, s->wt, mm);
    if (upd) {
        site_reduce<T>(w, dist, ms.sp[0], &red, ms.roff[0], ms.ioff[0]);
        stage_in<T>(sss, s->ss, mm);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        if (upd) w_update_dense<T>(s, red, mt, ssy, sss, swt);
        if (threadIdx.x == 0) {
            begin_body<T>(s);
            // the fast sequence goes on only into cauchy's full per-variable pass done by k_update_classify
            cont = (s->cauchy_mode == 0 && s->classify_done) ? 1 : 0;
            if (!cont) s->pause = PAUSE_CLASSIFY;
        }
    }
    __syncthreads();
    if (upd) { stage_out<T>(s->sy, ssy, mm); stage_out<T>(s->ss, sss, mm); stage_out<T>(s->wt, swt, mm); }
    if (!cont) return;
    site_reduce<T>(w, dist, ms.sp[1], &red, ms.roff[1], ms.ioff[1]);   // (ends with a barrier: the staged sy, wt are final)
    if (threadIdx.x >= 32) return;
    w_cauchy<T>(s, red, mt, ssy, swt, 1);
    if (threadIdx.x != 0) return;
    if (!s->in_body) return;                                 // singular bmv: memory reset, the host restarts the body
    if (s->need_walk) { s->pause = PAUSE_WALK; return; }
    if (!s->fuse_gf) { s->pause = PAUSE_GCP_FREEV; return; }
    s->do_subspace = 1; s->do_formk = s->updatd; s->do_delta = 0;   // s_freev phase 0 with fuse_gf
}

// s_freev(phase 1) + s_formk_dense + s_subsm_dense.
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) f_mid(Wk<T> w, Dist<T> dist, int mt, i64 n_global) {
    __shared__ Red<T> red;
    __shared__ T swn[4 * LB_MMAX * LB_MMAX], swn1[4 * LB_MMAX * LB_MMAX];
    __shared__ int cont;
    DevState<T>* s = w.s;
    const MSite ms = msite_mid(mt);
    site_fetch<T>(w, dist);
    if (!s->go || s->pause || !s->in_body) return;
    site_reduce<T>(w, dist, ms.sp[0], &red, ms.roff[0], ms.ioff[0]);
    if (threadIdx.x < 32) {
        w_freev<T>(s, red, n_global, true, true);
        if (threadIdx.x == 0) {
            cont = 1;
            if (!s->do_subspace) { s->pause = PAUSE_LSINIT; cont = 0; }        // no free variable: straight to the line search
            else if (s->do_delta) { s->pause = PAUSE_DELTA; cont = 0; }        // variables entered / left: formk's corrections first
        }
    }
    __syncthreads();
    if (!cont) return;
    const bool newrow = s->do_formk && s->updatd;
    if (newrow) site_reduce<T>(w, dist, ms.sp[1], &red, ms.roff[1], ms.ioff[1]);
    const bool stage = s->do_formk != 0;
    const int m4 = 4 * s->m * s->m;
    stage_in<T>(swn, s->wn, m4);
    if (stage) stage_in<T>(swn1, s->wn1, m4);
    __syncthreads();
    if (threadIdx.x < 32) w_formk_dense<T>(s, red, mt, (const T*)nullptr, swn, swn1);
    __syncthreads();
    if (stage) { stage_out<T>(s->wn, swn, m4); stage_out<T>(s->wn1, swn1, m4); }
    if (!s->in_body) return;                                 // formk failed: memory reset, the host restarts the body
    site_reduce<T>(w, dist, ms.sp[2], &red, ms.roff[2], ms.ioff[2]);
    if (threadIdx.x >= 32) return;
    w_subsm_dense<T>(s, red, mt, swn, 1);
}

// s_subsm_post + s_ls_init.
template <typename T>
__global__ void __launch_bounds__(LB_SCALAR_THREADS) f_tail(Wk<T> w, Dist<T> dist) {
    __shared__ Red<T> red;
    __shared__ int cont;
    DevState<T>* s = w.s;
    const MSite ms = msite_tail();
    site_fetch<T>(w, dist);
    if (!s->go || s->pause || !s->in_body || !s->do_subspace) return;
    site_reduce<T>(w, dist, ms.sp[0], &red, ms.roff[0], ms.ioff[0]);
    if (threadIdx.x == 0) {
        t0_subsm_post<T>(s, red, 1);
        cont = 1;
        if (s->do_backtrack) { s->pause = PAUSE_BACKTRACK; cont = 0; }
    }
    __syncthreads();
    if (!cont) return;
    site_reduce<T>(w, dist, ms.sp[1], &red, ms.roff[1], ms.ioff[1]);
    if (threadIdx.x != 0) return;
    t0_ls_init<T>(s, red);
}
