// The generalized-Cauchy-point search when the first breakpoint is reached
// (src/lbfgsb.f90:1378-1497).  The reference pops breakpoints one at a time from a
// heap (hpsolb :2079) and updates f1, f2, p, c sequentially.  Here:
//
//   1. ordered compaction of the breakpoints (t_i, i)                (k_flag_count/k_tile_scan/k_bp_write)
//   2. stable LSD radix sort on the bit pattern of t (t > 0)          (k_rs_*)  -> ties stay in index order
//   3. the loop-carried recurrences as prefix scans over the sorted list, in chunks:
//        p_j   = p_0 - sum_{i<j} d_i w_i                      (2col-vector prefix sum A)
//        c_j+1 = t_j p_j + sum_{i<j} z_i w_i                  (2col-vector prefix sum B, Abel summation)
//        f2_j+1 = max(eps f2_org, f2_j + g2_j)                ((max,+) scan)
//        f1_j+1 = f1_j + dt_j f2_j + g1_j                     (prefix sum)
//      exit at the first j with -f1_j/f2_j < dt_j  (:1416)
//   4. scatter of the fixed variables (xcp = bound, iwhere = 1/2, d = 0)  (k_walk_fix)
//
// Rounds.  The search usually ends after a small fraction of the breakpoints (the reference pops them one
// by one for the same reason), so steps 1-4 run in rounds over increasing ranges of t -- (0, 2 dtm0],
// (2 dtm0, 16 dtm0], (16 dtm0, inf) with dtm0 the minimiser of the first segment -- and stop as soon as
// the exit is known.  The count pass of a round also returns the smallest breakpoint not yet passed,
// which is all the exit test of the reference needs ("pop the next t, compare dt with dtm"), so a round
// whose predecessors already contain the exit is never compacted or sorted.  The running state
// (prefix sums, f1, f2, the last two t) is carried from round to round in the state block, exactly as
// it is carried from chunk to chunk and, on a sharded problem, from rank to rank.
//
// Decisions are the reference's decisions up to the rounding of the re-associated
// sums; see DESIGN.md "Cauchy walk" for the tie-order caveat.
#pragma once
#include "kernels_dense.cuh"

#define LB_WB 256                 // threads (= breakpoints) per walk block
#define LB_RS_GRID 592            // blocks of the radix sort (one contiguous chunk each)
#define LB_RS_ITEMS 16            // keys per thread per sub-tile
#define LB_RS_TILE (256 * LB_RS_ITEMS)

template <typename T> struct KeyBits;
template <> struct KeyBits<double> {
    __device__ static unsigned long long to(double t) { return (unsigned long long)__double_as_longlong(t); }
    __device__ static double from(unsigned long long k) { return __longlong_as_double((long long)k); }
};
template <> struct KeyBits<float> {
    __device__ static unsigned int to(float t) { return __float_as_uint(t); }
    __device__ static float from(unsigned int k) { return __uint_as_float(k); }
};

// small device block shared by the sort / compaction kernels
struct SortCtl {
    i64 count;        // number of items
    int cur;          // which of the two (key,val) buffers holds the current order
    int skip;         // current pass is a no-op (all keys share the digit)
};

// ---- block scans -----------------------------------------------------------
// exclusive prefix of v over the block (thread order); total returned to all threads.
// smem: 33 elements.
template <typename V>
__device__ __forceinline__ V block_excl_scan(V v, V* smem, V& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    V inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { V o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc = inc + o; }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        V run = (V)0;
        for (int q = 0; q < nw; ++q) { V t = smem[q]; smem[q] = run; run = run + t; }
        smem[32] = run;
    }
    __syncthreads();
    V excl = smem[wid] + (inc - v);
    total = smem[32];
    __syncthreads();
    return excl;
}

// breakpoint of variable i, recomputed from the classify pass outputs (:1305-1322)
template <typename T>
__device__ __forceinline__ bool bp_of(T d, T x, T l, T u, int nb, T& t) {
    if (nb <= 2 && nb != 0 && d < (T)0) { t = (x - l) / (-d); return true; }
    if (nb >= 2 && d > (T)0) { t = (u - x) / d; return true; }
    return false;
}
__device__ __forceinline__ bool el_of(int st) { return ((st & 1) != 0) != ((st & 2) != 0); }
// The breakpoint a variable had when cauchy's per-variable pass ran, also if the walk has fixed the variable
// since (d = 0, iwhere = 1/2, xcp = bound): such a variable is recognised by sitting off the bound it was
// sent to (the per-variable pass itself marks only variables that are ON the bound), and its d was -g.
template <typename T>
__device__ __forceinline__ bool bp_orig(T d, T g, T x, T l, T u, int nb, int iw, T& t) {
    if (d == (T)0) {
        if (iw == 1 && nb != 0 && nb <= 2 && x > l) d = -g;
        else if (iw == 2 && nb >= 2 && x < u) d = -g;
    }
    return bp_of<T>(d, x, l, u, nb, t);
}

// ---------------------------------------------------------------------------
// Ordered compaction.  MODE 1: entering / leaving variables -> val = variable (state bits tell which);
// MODE 2: every breakpoint of this cauchy call, passed ones included -> (key = bits of t, val = variable).
// ---------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_flag_count(Wk<T> w, int* tile_counts) {
    constexpr int VEC = Real<T>::VEC;
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body) return;
    if (MODE == 0 && !s->need_walk) return;
    if (MODE == 1 && !s->do_delta) return;
    if (MODE == 2 && !s->need_walk) return;
    const i64 n = w.n;
    const i64 tile = (i64)LBFGSB_BLOCK * VEC * Real<T>::UNROLL;
    const i64 ntiles = (n + tile - 1) / tile;
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    for (i64 tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        i64 c = 0;
#pragma unroll
        for (int k = 0; k < Real<T>::UNROLL; ++k) {
            const i64 base = tl * tile + (i64)k * (LBFGSB_BLOCK * VEC) + (i64)threadIdx.x * VEC;
            if (base >= n) continue;
            if (MODE == 2) {   // every breakpoint of this cauchy call, passed ones included (heap replay)
                T d[VEC], g[VEC], x[VEC], l[VEC], u[VEC]; int nb[VEC], iw[VEC];
                ldv<T>(w.g, base, n, g); ldv<T>(w.x, base, n, x); ldv<T>(w.l, base, n, l);
                ldv<T>(w.u, base, n, u); ldvi<T>(w.nbd, base, n, nb); ldvi<T>(w.iwhere, base, n, iw);
#pragma unroll
                for (int v = 0; v < VEC; ++v) { T t; d[v] = cauchy_dir<T>(iw[v], g[v]); if (base + v < n && bp_orig<T>(d[v], g[v], x[v], l[v], u[v], nb[v], iw[v], t)) c++; }
            } else {
                int st[VEC];
                ldvb<T>(w.state, base, n, st);
#pragma unroll
                for (int v = 0; v < VEC; ++v) if (base + v < n && el_of(st[v])) c++;
            }
        }
        i64 r = block_isum(c, smi);
        if (threadIdx.x == 0) tile_counts[tl] = (int)r;
    }
}

// MODE 1 count on its own: the state bytes of a tile are 2 or 4 KB -- one warp per tile, 16 bytes per load, no block-wide
// synchronisation (the list is short or empty in most iterations, and this pass is then all that the corrections cost).
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_el_count(Wk<T> w, int* tile_counts) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->do_delta) return;
    const i64 n = w.n;
    const i64 tile = (i64)LBFGSB_BLOCK * Real<T>::VEC * Real<T>::UNROLL;
    const i64 ntiles = (n + tile - 1) / tile;
    const int lane = threadIdx.x & 31;
    const i64 wid = (i64)blockIdx.x * (LBFGSB_BLOCK / 32) + (threadIdx.x >> 5), nw = (i64)gridDim.x * (LBFGSB_BLOCK / 32);
    for (i64 tl = wid; tl < ntiles; tl += nw) {
        const i64 t0 = tl * tile;
        int c = 0;
        for (i64 o = (i64)lane * 16; o < tile; o += 32 * 16) {
            const i64 i = t0 + o;
            if (i + 16 <= n) {   // state is allocated with a multiple of 32 entries and 256-byte aligned: 16-byte loads are aligned
                const uint4 q = *reinterpret_cast<const uint4*>(w.state + i);
                c += __popc((q.x ^ (q.x >> 1)) & 0x01010101u) + __popc((q.y ^ (q.y >> 1)) & 0x01010101u) +
                     __popc((q.z ^ (q.z >> 1)) & 0x01010101u) + __popc((q.w ^ (q.w >> 1)) & 0x01010101u);
            } else {
                for (i64 k = i; k < n && k < i + 16; ++k) c += el_of(w.state[k]) ? 1 : 0;
            }
        }
        c = warp_sum<int>(c);
        if (lane == 0) tile_counts[tl] = c;
    }
}

// Exclusive scan of counts[0..ntiles) into offsets by ONE block: thread t owns the contiguous run of ceil(ntiles / B)
// entries, sums it, the block scans the B run totals once (three barriers in all, whatever ntiles), and every thread
// writes its run's offsets.  Returns the total to every thread.  smem: 33 entries.
__device__ __forceinline__ i64 block_scan_runs(const int* counts, i64* offsets, i64 ntiles, i64* smem) {
    const i64 per = (ntiles + blockDim.x - 1) / blockDim.x;
    const i64 a = (i64)threadIdx.x * per;
    const i64 b = (a + per < ntiles) ? a + per : ntiles;
    i64 run = 0;
    for (i64 i = a; i < b; ++i) run += (i64)counts[i];
    i64 tot;
    i64 ex = block_excl_scan<i64>(run, smem, tot);
    for (i64 i = a; i < b; ++i) { offsets[i] = ex; ex += (i64)counts[i]; }
    return tot;
}

// exclusive scan of tile_counts -> offsets; total -> ctl->count.  One block of 1024.
template <typename T>
__global__ void __launch_bounds__(1024) k_tile_scan(Wk<T> w, int mode, const int* counts, i64* offsets, i64 ntiles,
                                                   SortCtl* ctl) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body) return;
    if ((mode == 0 || mode == 2) && !s->need_walk) return;
    if (mode == 1 && !s->do_delta) return;
    __shared__ i64 sm[33];
    const i64 carry = block_scan_runs(counts, offsets, ntiles, sm);
    if (threadIdx.x == 0) { ctl->count = carry; ctl->cur = 0; ctl->skip = 0; }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_flag_write(Wk<T> w, const int* tile_counts, const i64* tile_offsets,
                                                            typename Real<T>::key_t* keys, int* vals) {
    constexpr int VEC = Real<T>::VEC;
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body) return;
    if (MODE == 0 && !s->need_walk) return;
    if (MODE == 1 && !s->do_delta) return;
    if (MODE == 2 && !s->need_walk) return;
    const i64 n = w.n;
    const i64 tile = (i64)LBFGSB_BLOCK * VEC * Real<T>::UNROLL;
    const i64 ntiles = (n + tile - 1) / tile;
    __shared__ i64 sm[33];
    for (i64 tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        if (tile_counts[tl] == 0) continue;   // nothing to write from this tile (the common case: short lists)
        i64 run = tile_offsets[tl];
#pragma unroll 1
        for (int k = 0; k < Real<T>::UNROLL; ++k) {
            const i64 base = tl * tile + (i64)k * (LBFGSB_BLOCK * VEC) + (i64)threadIdx.x * VEC;
            bool fl[VEC]; T tv[VEC];
            i64 c = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) { fl[v] = false; tv[v] = (T)0; }
            if (base < n) {
                if (MODE == 2) {
                    T d[VEC], g[VEC], x[VEC], l[VEC], u[VEC]; int nb[VEC], iw[VEC];
                    ldv<T>(w.g, base, n, g); ldv<T>(w.x, base, n, x); ldv<T>(w.l, base, n, l);
                    ldv<T>(w.u, base, n, u); ldvi<T>(w.nbd, base, n, nb); ldvi<T>(w.iwhere, base, n, iw);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        d[v] = cauchy_dir<T>(iw[v], g[v]);
                        if (base + v < n && bp_orig<T>(d[v], g[v], x[v], l[v], u[v], nb[v], iw[v], tv[v])) { fl[v] = true; c++; }
                    }
                } else {
                    int st[VEC];
                    ldvb<T>(w.state, base, n, st);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) if (base + v < n && el_of(st[v])) { fl[v] = true; c++; }
                }
            }
            i64 tot;
            i64 pos = run + block_excl_scan<i64>(c, sm, tot);
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (fl[v]) {
                    if (MODE != 1) keys[pos] = KeyBits<T>::to(tv[v]);
                    vals[pos] = (int)(base + v);
                    pos++;
                }
            run += tot;
        }
    }
}

// ---------------------------------------------------------------------------
// Stable LSD radix sort, 8-bit digits.  gridDim.x <= LB_RS_GRID blocks (the host sizes the grid by the
// item count: a short list is sorted by a few blocks and a short scan), each owning one
// contiguous chunk of the input; counts are laid out [digit][block] so that one
// exclusive scan of the flattened array gives every (digit, block) its output base.
// ---------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(256) k_rs_hist(const K* k0, const K* k1, const SortCtl* ctl, int shift, int* counts) {
    const i64 nitems = ctl->count;
    const K* keys = ctl->cur ? k1 : k0;
    const int nblk = (int)gridDim.x;
    const i64 chunk = ((nitems + nblk - 1) / nblk + LB_RS_TILE - 1) / LB_RS_TILE * LB_RS_TILE;
    const i64 beg = (i64)blockIdx.x * chunk;
    i64 end = beg + chunk; if (end > nitems) end = nitems;
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (i64 i = beg + threadIdx.x; i < end; i += 256) atomicAdd(&h[(int)((keys[i] >> shift) & 0xff)], 1);
    __syncthreads();
    counts[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of counts[256*GRID] in place; sets ctl->skip if one digit holds everything
__global__ void __launch_bounds__(1024) k_rs_scan(int* counts, SortCtl* ctl, int nblk) {
    __shared__ i64 sm[33];
    __shared__ int allsame;
    const i64 total = 256 * (i64)nblk;
    if (threadIdx.x == 0) allsame = 0;
    __syncthreads();
    // digit totals: thread d < 256 sums its row
    if (threadIdx.x < 256) {
        i64 t = 0;
        for (int b = 0; b < nblk; ++b) t += counts[threadIdx.x * nblk + b];
        if (t == ctl->count) allsame = 1;
    }
    __syncthreads();
    if (allsame) { if (threadIdx.x == 0) ctl->skip = 1; return; }
    i64 carry = 0;
    for (i64 b0 = 0; b0 < total; b0 += 1024) {
        const i64 i = b0 + threadIdx.x;
        i64 v = (i < total) ? (i64)counts[i] : 0;
        i64 tot;
        i64 ex = block_excl_scan<i64>(v, sm, tot);
        if (i < total) counts[i] = (int)(carry + ex);
        carry += tot;
    }
    if (threadIdx.x == 0) ctl->skip = 0;
}

template <typename K>
__global__ void __launch_bounds__(256) k_rs_scatter(K* k0, K* k1, int* v0, int* v1, SortCtl* ctl, int shift,
                                                   const int* offsets) {
    if (ctl->skip) return;
    const i64 nitems = ctl->count;
    const K* kin = ctl->cur ? k1 : k0; K* kout = ctl->cur ? k0 : k1;
    const int* vin = ctl->cur ? v1 : v0; int* vout = ctl->cur ? v0 : v1;
    const int nblk = (int)gridDim.x;
    const i64 chunk = ((nitems + nblk - 1) / nblk + LB_RS_TILE - 1) / LB_RS_TILE * LB_RS_TILE;
    const i64 beg = (i64)blockIdx.x * chunk;
    i64 end = beg + chunk; if (end > nitems) end = nitems;
    __shared__ int cnt[8][257];
    __shared__ int dig_off[256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    dig_off[threadIdx.x] = offsets[threadIdx.x * nblk + blockIdx.x];
    __syncthreads();
    for (i64 t0 = beg; t0 < end; t0 += LB_RS_TILE) {
        for (int q = threadIdx.x; q < 8 * 257; q += 256) (&cnt[0][0])[q] = 0;
        __syncthreads();
        K key[LB_RS_ITEMS]; int val[LB_RS_ITEMS]; int dg[LB_RS_ITEMS]; int rk[LB_RS_ITEMS];
        const i64 wbase = t0 + (i64)wid * (32 * LB_RS_ITEMS);
#pragma unroll
        for (int k = 0; k < LB_RS_ITEMS; ++k) {
            const i64 i = wbase + k * 32 + lane;
            const bool ok = i < end;
            key[k] = ok ? kin[i] : (K)0; val[k] = ok ? vin[i] : 0;
            dg[k] = ok ? (int)((key[k] >> shift) & 0xff) : 256;
        }
#pragma unroll
        for (int k = 0; k < LB_RS_ITEMS; ++k) {
            const unsigned peers = __match_any_sync(0xffffffffu, dg[k]);
            const int below = __popc(peers & lt);
            const int pre = cnt[wid][dg[k]];
            __syncwarp();
            if (below == 0) cnt[wid][dg[k]] = pre + __popc(peers);
            __syncwarp();
            rk[k] = pre + below;
        }
        __syncthreads();
        // per digit: exclusive offsets over the warps, then the sub-tile total
        {
            const int d = threadIdx.x;
            int run = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { int t = cnt[q][d]; cnt[q][d] = run; run += t; }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < LB_RS_ITEMS; ++k) {
                if (dg[k] < 256) {
                    const i64 pos = (i64)dig_off[dg[k]] + cnt[wid][dg[k]] + rk[k];
                    kout[pos] = key[k]; vout[pos] = val[k];
                }
            }
            __syncthreads();
            dig_off[d] += run;
        }
        __syncthreads();
    }
}

__global__ void k_rs_flip(SortCtl* ctl) { if (threadIdx.x == 0 && !ctl->skip) ctl->cur ^= 1; }

// Short lists (the usual round: a few hundred breakpoints) in ONE launch instead of the 4 x sizeof(key) launches of the LSD
// sort: every item's position is its rank -- the number of items with a smaller key, or an equal key and a smaller
// position in the input -- counted against the keys held in shared memory.  Stable by construction: the same order as
// the radix sort.  In: buffer ctl->cur (0); out: the other buffer, ctl->cur = 1.
#define LB_SMALL_SORT_MAX 4096
template <typename K>
__global__ void __launch_bounds__(1024) k_small_sort(const K* k0, K* k1, const int* v0, int* v1, SortCtl* ctl) {
    __shared__ K sk[LB_SMALL_SORT_MAX];
    const int n = (int)ctl->count;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sk[i] = k0[i];
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const K ke = sk[e];
        int rank = 0;
        for (int f = 0; f < n; ++f) { const K kf = sk[f]; rank += (kf < ke || (kf == ke && f < e)) ? 1 : 0; }
        k1[rank] = ke; v1[rank] = v0[e];
    }
    if (threadIdx.x == 0) { ctl->cur = 1; ctl->skip = 0; }
}

// ---------------------------------------------------------------------------
// Walk scratch (one chunk of sorted breakpoints)
// ---------------------------------------------------------------------------
template <typename T>
struct WalkBuf {
    i64 cap;                 // chunk capacity (breakpoints)
    T* wj;                   // [2*MMAX][cap]  rows of W at the breakpoints (theta applied to the S half)
    T* vj;                   // [2*MMAX][cap]  M * w_j
    T *delta, *zeta, *omega, *g2, *g1, *f1a, *f2a;   // [cap]
    T *blkA, *blkB;          // [nblk][2*MMAX] block totals -> exclusive prefixes
    T *blk_a, *blk_b;        // (max,+) block operators -> f2 at block starts (blk_a reused)
    T *blkH;                 // block sums of h -> f1 at block starts
    typename Real<T>::key_t *k0, *k1; int *v0, *v1;   // sorted breakpoints (two buffers)
    SortCtl* ctl;
};

template <typename T>
__device__ __forceinline__ const typename Real<T>::key_t* cur_keys(const WalkBuf<T>& b) { return b.ctl->cur ? b.k1 : b.k0; }
template <typename T>
__device__ __forceinline__ const int* cur_vals(const WalkBuf<T>& b) { return b.ctl->cur ? b.v1 : b.v0; }
// t of a sorted entry.  In a round that replays a group of equal breakpoints on a sharded workspace (tie_round = 2) the
// sort keys are the members' positions in the reference's heap order, and t is the group's common value.
template <typename T>
__device__ __forceinline__ T walk_t_of(const DevState<T>* s, typename Real<T>::key_t key) {
    return (s->tie_round == 2) ? KeyBits<T>::from((typename Real<T>::key_t)s->tie_key) : KeyBits<T>::from(key);
}

// (a) gather: w_j, v_j = M w_j, omega_j, delta_j, zeta_j; block totals of delta*w and zeta*w
template <typename T>
__global__ void __launch_bounds__(LB_WB) k_walk_gather(Wk<T> w, WalkBuf<T> b, i64 start, i64 len) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_J >= 0) return;
    const int col = s->col, col2 = 2 * col, m = s->m, head0 = s->head - 1;
    const T theta = s->theta;
    __shared__ T ssy[LB_MMAX * LB_MMAX], swt[LB_MMAX * LB_MMAX];
    __shared__ T red[LB_WB / 32];
    for (int q = threadIdx.x; q < m * m; q += LB_WB) { ssy[q] = s->sy[q]; swt[q] = s->wt[q]; }
    __syncthreads();
    const i64 j = (i64)blockIdx.x * LB_WB + threadIdx.x;
    const bool ok = j < len;
    T dl = (T)0, ze = (T)0;
    T wl[2 * LB_MMAX], vl[2 * LB_MMAX];
    if (ok) {
        const int var = cur_vals<T>(b)[start + j];
        dl = cauchy_dir<T>(w.iwhere[var], w.g[var]);   // cauchy's d (:1294-1298), never stored
        ze = (dl > (T)0) ? (w.u[var] - w.x[var]) : (w.l[var] - w.x[var]);   // zibp (:1427,:1431)
        b.delta[j] = dl; b.zeta[j] = ze;
        if (col > 0) {
            int pj = head0;
            for (int c = 0; c < col; ++c) {
                wl[c] = w.wy[(i64)pj * w.ldw + var];
                wl[col + c] = theta * w.ws[(i64)pj * w.ldw + var];   // :1463-1464
                pj = (pj + 1 == m) ? 0 : pj + 1;
            }
            dense::bmv<T>(m, ssy, swt, col, wl, vl);   // singular T was already excluded by the first bmv (:1360)
            T om = (T)0;
            for (int c = 0; c < col2; ++c) { om = om + wl[c] * vl[c]; b.wj[(i64)c * b.cap + j] = wl[c]; b.vj[(i64)c * b.cap + j] = vl[c]; }
            b.omega[j] = om;
        } else b.omega[j] = (T)0;
    }
    // block totals (fixed order: butterfly + serial over warps)
    for (int c = 0; c < col2; ++c) {
        for (int which = 0; which < 2; ++which) {
            T v = ok ? ((which == 0 ? dl : ze) * wl[c]) : (T)0;
            v = warp_sum<T>(v);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
            __syncthreads();
            if (threadIdx.x == 0) {
                T t = red[0];
                for (int q = 1; q < LB_WB / 32; ++q) t = t + red[q];
                (which == 0 ? b.blkA : b.blkB)[(i64)blockIdx.x * (2 * LB_MMAX) + c] = t;
            }
            __syncthreads();
        }
    }
}

// (b) exclusive scan of the block totals over the blocks, with the carry from earlier chunks.
//     One warp per sequence.  Leaves the chunk totals (incl. carry) in s->walkA/B *candidates*
//     (tmpA/tmpB), committed by s_walk_chunk_end if no exit was found in this chunk.
template <typename T>
__global__ void __launch_bounds__(1024) k_walk_scan_vec(Wk<T> w, WalkBuf<T> b, i64 nblk, T* tmpAB) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_J >= 0) return;
    const int col2 = 2 * s->col;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int seq = wid; seq < 2 * col2; seq += nw) {
        const int which = seq / col2, c = seq % col2;
        T* arr = which == 0 ? b.blkA : b.blkB;
        T carry = which == 0 ? s->walkA[c] : s->walkB[c];
        for (i64 b0 = 0; b0 < nblk; b0 += 32) {
            const i64 i = b0 + lane;
            T v = (i < nblk) ? arr[i * (2 * LB_MMAX) + c] : (T)0;
            T inc = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { T o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc = inc + o; }
            if (i < nblk) arr[i * (2 * LB_MMAX) + c] = carry + (inc - v);
            carry = carry + __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) tmpAB[which * (2 * LB_MMAX) + c] = carry;
    }
}

// (max,+) operator: x -> max(a, x + b); compose(first, then)
template <typename T> struct MaxPlus { T a, b; };
template <typename T>
__device__ __forceinline__ MaxPlus<T> mp_compose(MaxPlus<T> f, MaxPlus<T> g) {   // g after f
    MaxPlus<T> r;
    r.a = dense::tmax(g.a, f.a + g.b);
    r.b = f.b + g.b;
    return r;
}
template <typename T>
__device__ __forceinline__ T mp_apply(MaxPlus<T> f, T x) { return dense::tmax(f.a, x + f.b); }

// inclusive scan of (max,+) operators over the block; returns the exclusive operator of
// this thread and the block aggregate.  smem: 2*(33) T.
template <typename T>
__device__ __forceinline__ MaxPlus<T> block_excl_scan_mp(MaxPlus<T> v, T* sma, T* smb, MaxPlus<T>& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const T NEG = -LB_INF(T);
    MaxPlus<T> inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        MaxPlus<T> o; o.a = __shfl_up_sync(0xffffffffu, inc.a, off); o.b = __shfl_up_sync(0xffffffffu, inc.b, off);
        if (lane >= off) inc = mp_compose<T>(o, inc);
    }
    // exclusive within warp
    MaxPlus<T> ex; ex.a = __shfl_up_sync(0xffffffffu, inc.a, 1); ex.b = __shfl_up_sync(0xffffffffu, inc.b, 1);
    if (lane == 0) { ex.a = NEG; ex.b = (T)0; }
    if (lane == 31) { sma[wid] = inc.a; smb[wid] = inc.b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        MaxPlus<T> run; run.a = NEG; run.b = (T)0;
        for (int q = 0; q < nw; ++q) {
            MaxPlus<T> t; t.a = sma[q]; t.b = smb[q];
            sma[q] = run.a; smb[q] = run.b;
            run = mp_compose<T>(run, t);
        }
        sma[32] = run.a; smb[32] = run.b;
    }
    __syncthreads();
    MaxPlus<T> pre; pre.a = sma[wid]; pre.b = smb[wid];
    MaxPlus<T> r = mp_compose<T>(pre, ex);
    total.a = sma[32]; total.b = smb[32];
    __syncthreads();
    return r;
}

// (c) wmp_j, wmc_j via the prefix sums; g2_j, g1_j; block (max,+) aggregates
template <typename T>
__global__ void __launch_bounds__(LB_WB) k_walk_dots(Wk<T> w, WalkBuf<T> b, i64 start, i64 len) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_J >= 0) return;
    const int col = s->col, col2 = 2 * col;
    const T theta = s->theta, two = (T)2;
    __shared__ T sm[33], sma[33], smb[33];
    const i64 j = (i64)blockIdx.x * LB_WB + threadIdx.x;
    const bool ok = j < len;
    const T dl = ok ? b.delta[j] : (T)0, ze = ok ? b.zeta[j] : (T)0;
    const T tj = ok ? walk_t_of<T>(s, cur_keys<T>(b)[start + j]) : (T)0;
    T wmp = (T)0, wmc = (T)0;
    for (int c = 0; c < col2; ++c) {
        const T wv_ = ok ? b.wj[(i64)c * b.cap + j] : (T)0;
        const T vv = ok ? b.vj[(i64)c * b.cap + j] : (T)0;
        T tot;
        const T exA = block_excl_scan<T>(dl * wv_, sm, tot);
        const T exB = block_excl_scan<T>(ze * wv_, sm, tot);
        const T Aj = b.blkA[(i64)blockIdx.x * (2 * LB_MMAX) + c] + exA;
        const T Bj = b.blkB[(i64)blockIdx.x * (2 * LB_MMAX) + c] + exB;
        const T pj = s->p0[c] - Aj;            // p at the start of segment j
        const T cj = tj * pj + Bj;             // c after "c = c + dt*p" (:1457)
        wmp = wmp + pj * vv;                   // :1472
        wmc = wmc + cj * vv;                   // :1471
    }
    const T d2 = dl * dl;
    T g2 = -theta * d2, g1 = d2 - theta * dl * ze;   // :1452-1453
    if (col > 0 && ok) {
        g1 = g1 + dl * wmc;                    // :1479
        g2 = g2 + two * dl * wmp - d2 * b.omega[j];   // :1480
    }
    if (ok) { b.g2[j] = g2; b.g1[j] = g1; }
    MaxPlus<T> op; op.a = ok ? s->epsmch * s->f2_org : -LB_INF(T); op.b = ok ? g2 : (T)0;   // :1483
    MaxPlus<T> total;
    block_excl_scan_mp<T>(op, sma, smb, total);
    if (threadIdx.x == 0) { b.blk_a[blockIdx.x] = total.a; b.blk_b[blockIdx.x] = total.b; }
}

// (d) f2 at the block starts: serial composition over the blocks by one warp
template <typename T>
__global__ void __launch_bounds__(32) k_walk_scan_f2(Wk<T> w, WalkBuf<T> b, i64 nblk, T* tmpF) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_J >= 0) return;
    const int lane = threadIdx.x;
    const T NEG = -LB_INF(T);
    T f2 = s->walk_f2;
    for (i64 b0 = 0; b0 < nblk; b0 += 32) {
        const i64 i = b0 + lane;
        MaxPlus<T> v; v.a = (i < nblk) ? b.blk_a[i] : NEG; v.b = (i < nblk) ? b.blk_b[i] : (T)0;
        MaxPlus<T> inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            MaxPlus<T> o; o.a = __shfl_up_sync(0xffffffffu, inc.a, off); o.b = __shfl_up_sync(0xffffffffu, inc.b, off);
            if (lane >= off) inc = mp_compose<T>(o, inc);
        }
        MaxPlus<T> ex; ex.a = __shfl_up_sync(0xffffffffu, inc.a, 1); ex.b = __shfl_up_sync(0xffffffffu, inc.b, 1);
        if (lane == 0) { ex.a = NEG; ex.b = (T)0; }
        if (i < nblk) b.blk_a[i] = mp_apply<T>(ex, f2);   // f2 at the start of block i
        MaxPlus<T> last; last.a = __shfl_sync(0xffffffffu, inc.a, 31); last.b = __shfl_sync(0xffffffffu, inc.b, 31);
        f2 = mp_apply<T>(last, f2);
    }
    if (lane == 0) tmpF[1] = f2;   // f2 after the whole chunk
}

// (e) f2_j, h_j = dt_j f2_j + g1_j, block sums of h
template <typename T>
__global__ void __launch_bounds__(LB_WB) k_walk_f2(Wk<T> w, WalkBuf<T> b, i64 start, i64 len) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_J >= 0) return;
    __shared__ T sma[33], smb[33], red[LB_WB / 32];
    const i64 j = (i64)blockIdx.x * LB_WB + threadIdx.x;
    const bool ok = j < len;
    MaxPlus<T> op; op.a = ok ? s->epsmch * s->f2_org : -LB_INF(T); op.b = ok ? b.g2[j] : (T)0;
    MaxPlus<T> total;
    MaxPlus<T> ex = block_excl_scan_mp<T>(op, sma, smb, total);
    const T f2j = mp_apply<T>(ex, b.blk_a[blockIdx.x]);
    T h = (T)0;
    if (ok) {
        const typename Real<T>::key_t* keys = cur_keys<T>(b);
        const T tj = walk_t_of<T>(s, keys[start + j]);
        const T tp = (start + j > 0) ? walk_t_of<T>(s, keys[start + j - 1]) : s->walk_tlast;   // 0, or the last key of the lower ranks
        const T dt = tj - tp;
        h = dt * f2j + b.g1[j];
        b.f2a[j] = f2j;
        b.g2[j] = dt;       // g2 no longer needed: keep dt_j for the test
        b.g1[j] = h;
    }
    T v = warp_sum<T>(h);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { T t = red[0]; for (int q = 1; q < LB_WB / 32; ++q) t = t + red[q]; b.blkH[blockIdx.x] = t; }
}

// (f) f1 at the block starts
template <typename T>
__global__ void __launch_bounds__(32) k_walk_scan_f1(Wk<T> w, WalkBuf<T> b, i64 nblk, T* tmpF) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_J >= 0) return;
    const int lane = threadIdx.x;
    T carry = s->walk_f1;
    for (i64 b0 = 0; b0 < nblk; b0 += 32) {
        const i64 i = b0 + lane;
        T v = (i < nblk) ? b.blkH[i] : (T)0;
        T inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { T o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc = inc + o; }
        if (i < nblk) b.blkH[i] = carry + (inc - v);
        carry = carry + __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) tmpF[0] = carry;   // f1 after the whole chunk
}

// (g) f1_j and the exit test (:1416)
template <typename T>
__global__ void __launch_bounds__(LB_WB) k_walk_test(Wk<T> w, WalkBuf<T> b, i64 start, i64 len, unsigned long long* jmin) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_J >= 0) return;
    __shared__ T sm[33];
    const i64 j = (i64)blockIdx.x * LB_WB + threadIdx.x;
    const bool ok = j < len;
    T tot;
    const T ex = block_excl_scan<T>(ok ? b.g1[j] : (T)0, sm, tot);
    if (ok) {
        const T f1j = b.blkH[blockIdx.x] + ex;
        const T f2j = b.f2a[j];
        b.f1a[j] = f1j;
        const T dtm = -f1j / f2j;
        if (dtm < b.g2[j]) atomicMin(jmin, (unsigned long long)(start + j));
    }
}

// end of a chunk: commit the exit position or the carries
template <typename T>
__global__ void k_walk_chunk_end(Wk<T> w, WalkBuf<T> b, i64 start, i64 len, const T* tmpAB, const T* tmpF,
                                 const unsigned long long* jmin) {
    DevState<T>* s = w.s;
    if (threadIdx.x != 0) return;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_J >= 0) return;
    if (*jmin != 0xffffffffffffffffULL) { s->walk_J = (i64)*jmin; s->walk_cstart = start; return; }
    const int col2 = 2 * s->col;
    for (int c = 0; c < col2; ++c) { s->walkA[c] = tmpAB[c]; s->walkB[c] = tmpAB[2 * LB_MMAX + c]; }
    s->walk_f1 = tmpF[0]; s->walk_f2 = tmpF[1];
    s->walk_done = start + len;
}

// ---------------------------------------------------------------------------
// Rounds: range-filtered compaction, the "next breakpoint" peek, and the closing of the search.
// ---------------------------------------------------------------------------
struct BpRange { unsigned long long lo, hi; int lo_valid; };   // keys in (lo, hi]; no lower limit when !lo_valid
__device__ __forceinline__ bool bp_rem(const BpRange& r, unsigned long long k) { return !r.lo_valid || k > r.lo; }
__device__ __forceinline__ bool bp_in(const BpRange& r, unsigned long long k) { return bp_rem(r, k) && k <= r.hi; }
struct RoundRec { i64 count, rem; unsigned long long kmin; i64 pad; };   // of one rank

// The breakpoint of every variable is stored once per cauchy call, in front of the first round (k_bp_count<T, true>):
// w.r -- dead between subsm of one iteration and cmprlb of the next -- holds t_i, or -1 where the variable has none
// (or has been fixed by an earlier round, k_walk_fix).  The passes of a round then read 8 bytes per variable
// instead of the 36 of d, x, l, u, nbd.
//
// count pass of a round: per tile the breakpoints in range; per block the number of breakpoints not yet
// passed (key > lo) and the smallest of them.  ipart slot 0: rem, slot 1: kmin (keys of t >= 0 fit in i64).
// FIRST: the first round of a call whose per-variable pass did not expect a walk (Wk::bp_hint off) -- also writes
// xcp = x (:1341; the walk scatters the bounds of the fixed variables into it) and the breakpoints themselves.
// cauchy's d (:1294-1298) is never stored: it is -g where iwhere is 0 or -1 and 0 elsewhere (cauchy_dir), also after
// the walk has fixed a variable (iwhere = 1 / 2).
template <typename T, bool FIRST>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_bp_count(Wk<T> w, BpRange rg, int* tile_counts) {
    constexpr int VEC = Real<T>::VEC;
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_closed) return;
    const i64 n = w.n;
    const i64 tile = (i64)LBFGSB_BLOCK * VEC * Real<T>::UNROLL;
    const i64 ntiles = (n + tile - 1) / tile;
    __shared__ i64 smi[LBFGSB_BLOCK / 32];
    i64 rem = 0, kmin = LB_I64MAX;
    for (i64 tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        i64 c = 0;
#pragma unroll
        for (int k = 0; k < Real<T>::UNROLL; ++k) {
            const i64 base = tl * tile + (i64)k * (LBFGSB_BLOCK * VEC) + (i64)threadIdx.x * VEC;
            if (base >= n) continue;
            T tk[VEC];
            if (FIRST) {
                int iw[VEC], nb[VEC]; T g[VEC], x[VEC], l[VEC], u[VEC], d[VEC];
                ldvi<T>(w.iwhere, base, n, iw); ldv<T>(w.g, base, n, g); ldv<T>(w.x, base, n, x);
                ldv<T>(w.l, base, n, l); ldv<T>(w.u, base, n, u); ldvi<T>(w.nbd, base, n, nb);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    d[v] = cauchy_dir<T>(iw[v], g[v]);
                    T t;
                    tk[v] = (base + v < n && bp_of<T>(d[v], x[v], l[v], u[v], nb[v], t)) ? t : (T)-1;
                }
                stv<T>(w.z, base, n, x); stv<T>(w.r, base, n, tk);
            } else ldv<T>(w.r, base, n, tk);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (base + v < n && tk[v] >= (T)0) {
                    const unsigned long long key = (unsigned long long)KeyBits<T>::to(tk[v]);
                    if (bp_rem(rg, key)) { rem++; if ((i64)key < kmin) kmin = (i64)key; }
                    if (bp_in(rg, key)) c++;
                }
            }
        }
        i64 r = block_isum(c, smi);
        if (threadIdx.x == 0) tile_counts[tl] = (int)r;
    }
    const i64 r0 = block_isum(rem, smi);
    i64 km = warp_min<i64>(kmin);
    if ((threadIdx.x & 31) == 0) smi[threadIdx.x >> 5] = km;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < LBFGSB_BLOCK / 32; ++q) km = smi[q] < km ? smi[q] : km;
        LB_SLOT(w.ipart, 0)[blockIdx.x] = r0;
        LB_SLOT(w.ipart, 1)[blockIdx.x] = km;
    }
}

// ordered compaction of the breakpoints in range -> (key, local variable)
template <typename T>
__global__ void __launch_bounds__(LBFGSB_BLOCK) k_bp_write(Wk<T> w, BpRange rg, const int* tile_counts, const i64* tile_offsets,
                                                          typename Real<T>::key_t* keys, int* vals) {
    constexpr int VEC = Real<T>::VEC;
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_closed) return;
    const i64 n = w.n;
    const i64 tile = (i64)LBFGSB_BLOCK * VEC * Real<T>::UNROLL;
    const i64 ntiles = (n + tile - 1) / tile;
    __shared__ i64 sm[33];
    for (i64 tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        if (tile_counts[tl] == 0) continue;   // nothing to write from this tile (the common case: short lists)
        i64 run = tile_offsets[tl];
#pragma unroll 1
        for (int k = 0; k < Real<T>::UNROLL; ++k) {
            const i64 base = tl * tile + (i64)k * (LBFGSB_BLOCK * VEC) + (i64)threadIdx.x * VEC;
            bool fl[VEC]; T tv[VEC];
            i64 c = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) { fl[v] = false; tv[v] = (T)0; }
            if (base < n) {
                ldv<T>(w.r, base, n, tv);   // the stored breakpoints (k_bp_count<T, true>)
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (base + v < n && tv[v] >= (T)0 && bp_in(rg, (unsigned long long)KeyBits<T>::to(tv[v]))) { fl[v] = true; c++; }
            }
            i64 tot;
            i64 pos = run + block_excl_scan<i64>(c, sm, tot);
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (fl[v]) { keys[pos] = KeyBits<T>::to(tv[v]); vals[pos] = (int)(base + v); pos++; }
            run += tot;
        }
    }
}

// this rank's part of a round: tile offsets, ctl->count, and the record (count, rem, kmin).  One block of 1024.
template <typename T>
__global__ void __launch_bounds__(1024) k_walk_round_local(Wk<T> w, const int* counts, i64* offsets, i64 ntiles, SortCtl* ctl,
                                                          RoundRec* out) {
    DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_closed) return;
    __shared__ i64 sm[33];
    const i64 carry = block_scan_runs(counts, offsets, ntiles, sm);
    if (threadIdx.x < 32) {
        const i64 rem = final_isum_warp(LB_SLOT(w.ipart, 0));
        const i64 km = final_imin_warp(LB_SLOT(w.ipart, 1));
        if (threadIdx.x == 0) {
            ctl->count = carry; ctl->cur = 0; ctl->skip = 0;
            s->walk_lcount = carry;
            out->count = carry; out->rem = rem; out->kmin = (unsigned long long)km; out->pad = 0;
        }
    }
}

// every breakpoint was passed (:1436-1442, :1484-1495): close from the carried state
template <typename T>
__device__ inline void walk_close_all_passed(DevState<T>* s, i64 n_global) {
    const int col2 = 2 * s->col;
    const i64 nb = s->nbreak;
    const T tlast = s->walk_tlast;
    T dtm;
    if (nb == n_global) {   // all n variables fixed (:1436-1442)
        dtm = tlast - ((nb > 1) ? s->walk_tprev2 : (T)0);
        s->nseg = nb;
        s->tsum = tlast;
        for (int c = 0; c < col2; ++c) { const T pJ = s->p0[c] - s->walkA[c]; s->p[c] = pJ; s->c[c] = tlast * pJ + s->walkB[c]; }
        s->dtm = dtm;
    } else {
        s->nseg = nb + 1;
        T f1 = s->walk_f1, f2 = s->walk_f2;
        if (s->bnded) { f1 = (T)0; f2 = (T)0; dtm = (T)0; }
        else dtm = -f1 / f2;
        if (dtm <= (T)0) dtm = (T)0;
        s->f1 = f1; s->f2 = f2; s->dtm = dtm;
        s->tsum = tlast + dtm;
        for (int c = 0; c < col2; ++c) {
            const T pJ = s->p0[c] - s->walkA[c];
            s->p[c] = pJ;
            s->c[c] = (tlast * pJ + s->walkB[c]) + dtm * pJ;
        }
    }
    s->walk_closed = 1;
}
// the exit lies in the segment that starts at tprev, with prefix sums AJ, BJ (:1416, :1509-1526)
template <typename T>
__device__ inline void walk_close_at(DevState<T>* s, T f1, T f2, T tprev, const T* AJ, const T* BJ, i64 nseg) {
    const int col2 = 2 * s->col;
    T dtm = -f1 / f2;
    if (dtm <= (T)0) dtm = (T)0;
    s->f1 = f1; s->f2 = f2; s->dtm = dtm;
    s->tsum = tprev + dtm;
    s->nseg = nseg;
    for (int c = 0; c < col2; ++c) {
        const T pJ = s->p0[c] - AJ[c];
        s->p[c] = pJ;
        s->c[c] = (tprev * pJ + BJ[c]) + dtm * pJ;
    }
    s->walk_closed = 1;
}

// Start of a round (every rank, identical inputs): totals over the ranks, and the test the reference makes
// when it pops the next breakpoint (:1416) against the smallest breakpoint not yet passed -- exactly what
// k_walk_test would compute for the first entry of this or a later round.
template <typename T>
__global__ void k_walk_round_begin(Wk<T> w, const RoundRec* all, int R, i64 n_global) {
    if (threadIdx.x != 0) return;
    DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_closed) return;
    i64 cnt = 0, rem = 0; unsigned long long kmin = 0xffffffffffffffffULL;
    for (int q = 0; q < R; ++q) { cnt += all[q].count; rem += all[q].rem; if (all[q].rem > 0 && all[q].kmin < kmin) kmin = all[q].kmin; }
    s->walk_rcount = cnt; s->walk_rem = rem; s->walk_J = -1; s->walk_done = 0; s->walk_fixn = 0;
    s->tie_round = 0; s->tie_redo = 0;
    {   // the carried state at the start of the round (restored if the round is redone up to a group of ties)
        const int col2 = 2 * s->col;
        for (int c = 0; c < col2; ++c) { s->snapA[c] = s->walkA[c]; s->snapB[c] = s->walkB[c]; }
        s->snap_f1 = s->walk_f1; s->snap_f2 = s->walk_f2; s->snap_tlast = s->walk_tlast; s->snap_tprev2 = s->walk_tprev2;
        s->snap_base = s->walk_base;
    }
    if (rem == 0) { walk_close_all_passed<T>(s, n_global); return; }
    if (s->walk_base > 0) {
        const T tnext = KeyBits<T>::from((typename Real<T>::key_t)kmin);
        const T dt = tnext - s->walk_tlast;
        const T dtm = -s->walk_f1 / s->walk_f2;
        if (dtm < dt) walk_close_at<T>(s, s->walk_f1, s->walk_f2, s->walk_tlast, s->walkA, s->walkB, 1 + s->walk_base);
    }
}

// End of a round on a single GPU, after its chunks.  One block of LB_WB threads.
template <typename T>
__global__ void __launch_bounds__(LB_WB) k_walk_round_end(Wk<T> w, WalkBuf<T> b, i64 n_global, int single_gpu) {
    DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_closed) return;
    __shared__ T AJ[2 * LB_MMAX], BJ[2 * LB_MMAX];
    __shared__ T sm[33];
    const int col2 = 2 * s->col;
    const i64 cnt = b.ctl->count;
    const i64 J = s->walk_J;
    const typename Real<T>::key_t* keys = cur_keys<T>(b);
    if (J >= 0) {
        // prefix sums at J: block prefix + in-block partial sums of the chunk that holds J
        const i64 start = s->walk_cstart;
        const i64 jl = J - start;
        const i64 blk = jl / LB_WB;
        const i64 j = blk * LB_WB + threadIdx.x;
        const bool in = j < jl;
        for (int c = 0; c < col2; ++c) {
            const T wv_ = in ? b.wj[(i64)c * b.cap + j] : (T)0;
            T tot;
            block_excl_scan<T>(in ? b.delta[j] * wv_ : (T)0, sm, tot);
            if (threadIdx.x == 0) AJ[c] = b.blkA[blk * (2 * LB_MMAX) + c] + tot;
            block_excl_scan<T>(in ? b.zeta[j] * wv_ : (T)0, sm, tot);
            if (threadIdx.x == 0) BJ[c] = b.blkB[blk * (2 * LB_MMAX) + c] + tot;
        }
        __syncthreads();
        if (threadIdx.x != 0) return;
        if (J > 0 && keys[J - 1] == keys[J] && !s->tie_round) {
            // The exit lies inside a group of equal breakpoints: which members of the group were passed depends
            // on the order in which they are taken.  The reference takes them in hpsolb's heap order; redo the
            // round up to the group and then the group in that order (host: tie_replay), if the heap is small
            // enough to be replayed.  Otherwise the variable order stands and the event is counted.
            if (single_gpu && s->tie_limit > 0 && s->nbreak <= s->tie_limit) {
                s->tie_redo = 1; s->tie_key = (unsigned long long)keys[J]; s->walk_J = -1; s->walk_fixn = 0;
                return;
            }
            s->tie_events += 1;
        }
        const T tprev = (J > 0) ? walk_t_of<T>(s, keys[J - 1]) : s->walk_tlast;
        walk_close_at<T>(s, b.f1a[jl], b.f2a[jl], tprev, AJ, BJ, 1 + s->walk_base + J);
        s->walk_fixn = J;
        return;
    }
    if (threadIdx.x != 0) return;
    // every breakpoint of the round was passed
    if (cnt >= 2) { s->walk_tprev2 = walk_t_of<T>(s, keys[cnt - 2]); s->walk_tlast = walk_t_of<T>(s, keys[cnt - 1]); }
    else if (cnt == 1) { s->walk_tprev2 = s->walk_tlast; s->walk_tlast = walk_t_of<T>(s, keys[0]); }
    s->walk_base += cnt;
    s->walk_fixn = cnt;
    if (cnt == s->walk_rem) walk_close_all_passed<T>(s, n_global);
}

// fix the variables whose breakpoints were passed in this round (:1424-1434)
template <typename T>
__global__ void __launch_bounds__(256) k_walk_fix(Wk<T> w, WalkBuf<T> b) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk) return;
    const i64 J = s->walk_fixn;
    const int* vals = cur_vals<T>(b);
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < J; j += (i64)gridDim.x * blockDim.x) {
        const int var = vals[j];
        const T dl = cauchy_dir<T>(w.iwhere[var], w.g[var]);
        // d = 0 from here on is implied by iwhere = 1 / 2 (cauchy_dir)
        if (dl > (T)0) { w.z[var] = w.u[var]; w.iwhere[var] = 2; }
        else { w.z[var] = w.l[var]; w.iwhere[var] = 1; }
        w.r[var] = (T)-1;   // no breakpoint any more (what bp_of gives for d = 0)
    }
}

// ---------------------------------------------------------------------------
// Heap replay.  When the search ends inside a group of equal breakpoints, the members of the group that
// were passed (fixed at their bounds, iwhere = 1/2) are the first ones in the order in which the reference
// pops them from its heap (hpsolb :2079-2157) -- an order that depends on the whole history of the heap.
// The replay rebuilds that history: the breakpoint array in variable order as cauchy's per-variable pass
// leaves it (:1305-1322), the first minimum replaced by the last entry (:1391-1397), the heap built by
// successive insertion and popped until the group is exhausted.  One thread does the heap operations (they
// are a chain of dependent compares), which is why the replay is limited to tie_limit breakpoints.
// The popped members of the group, in order, become the sorted list of the next round.
// ---------------------------------------------------------------------------
template <typename K>
__host__ __device__ inline void heap_build(K* t, int* io, i64 n) {   // t, io: 1-based views; hpsolb :2096-2119
    for (i64 k = 2; k <= n; ++k) {
        const K ddum = t[k];
        const int indxin = io[k];
        i64 i = k;
        while (i > 1) {
            const i64 j = i / 2;
            const K tj = t[j];
            if (ddum < tj) { t[i] = tj; io[i] = io[j]; i = j; } else break;
        }
        t[i] = ddum; io[i] = indxin;
    }
}
// least member out, the rest re-heaped as t(1..n-1); the reference leaves the least member in t(n)  (:2125-2155)
template <typename K>
__host__ __device__ inline void heap_pop(K* t, int* io, i64 n, K& out, int& indxou) {
    out = t[1]; indxou = io[1];
    if (n > 1) {
        i64 i = 1;
        const K ddum = t[n];
        const int indxin = io[n];
        for (;;) {
            i64 j = i + i;
            if (j <= n - 1) {
                K tj = t[j];
                const K tj1 = t[j + 1];   // j + 1 <= n: the last slot still holds ddum, as in the reference
                if (tj1 < tj) { j = j + 1; tj = tj1; }
                if (tj < ddum) { t[i] = tj; io[i] = io[j]; i = j; continue; }
            }
            break;
        }
        t[i] = ddum; io[i] = indxin;
        t[n] = out; io[n] = indxou;
    }
}

// carried state back to the start of the round
template <typename T>
__global__ void k_tie_restore(Wk<T> w) {
    if (threadIdx.x != 0) return;
    DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || !s->tie_redo) return;
    const int col2 = 2 * s->col;
    for (int c = 0; c < col2; ++c) { s->walkA[c] = s->snapA[c]; s->walkB[c] = s->snapB[c]; }
    s->walk_f1 = s->snap_f1; s->walk_f2 = s->snap_f2; s->walk_tlast = s->snap_tlast; s->walk_tprev2 = s->snap_tprev2;
    s->walk_base = s->snap_base;
    s->walk_J = -1; s->walk_done = 0; s->walk_fixn = 0; s->walk_closed = 0;
    s->tie_redo = 0;
}

// b.k0 / b.v0: every breakpoint of this cauchy call in variable order (k_flag_write<T, 2>), b.ctl->count of them.
// Out: the group with key == tie_key in heap order in b.k1 / b.v1, b.ctl->count = its size, b.ctl->cur = 1, and the
// fields a round needs (what k_walk_round_begin sets).
template <typename T>
__global__ void __launch_bounds__(1024) k_heap_replay(Wk<T> w, WalkBuf<T> b) {
    typedef typename Real<T>::key_t K;
    DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_closed) return;
    __shared__ i64 spos;
    const i64 nb = b.ctl->count;
    const int vmin = (int)(s->ibkmin - w.off);
    if (threadIdx.x == 0) spos = -1;
    __syncthreads();
    for (i64 i = threadIdx.x; i < nb; i += blockDim.x) if (b.v0[i] == vmin) spos = i;
    __syncthreads();
    if (threadIdx.x != 0) return;
    K* t = b.k0 - 1; int* io = b.v0 - 1;      // 1-based
    const K tk = (K)s->tie_key;
    const K kfirst = KeyBits<T>::to(s->bkmin);
    i64 cnt = 0;
    if (kfirst == tk) { b.k1[cnt] = tk; b.v1[cnt] = vmin; cnt++; }   // the first breakpoint is taken before the heap exists (:1384-1389)
    const i64 ib = spos + 1;
    if (ib >= 1 && ib != nb) { t[ib] = t[nb]; io[ib] = io[nb]; }   // :1394-1397
    i64 nleft = nb - 1;
    heap_build<K>(t, io, nleft);
    while (nleft > 0) {
        K out; int var;
        heap_pop<K>(t, io, nleft, out, var);
        nleft--;
        if (out > tk) break;
        if (out == tk) { b.k1[cnt] = out; b.v1[cnt] = var; cnt++; }
    }
    b.ctl->count = cnt; b.ctl->cur = 1; b.ctl->skip = 0;
    s->walk_lcount = cnt; s->walk_rcount = cnt; s->walk_rem = s->nbreak - s->walk_base;
    s->walk_J = -1; s->walk_done = 0; s->walk_fixn = 0;
    s->tie_round = 1;
}

// The replay done on the host (long breakpoint lists: a single device thread would take a second per million heap
// insertions, a host core takes a few tens of nanoseconds each): the host has popped the heap from a copy of b.k0 / b.v0
// and left the group in heap order in b.k1 / b.v1; this sets what k_heap_replay sets at its end.
// Sharded workspaces: `cnt` = this rank's members (b.k1 = their positions in the heap order, b.v1 = local variables),
// `cnt_all` = the size of the group over all ranks, kind = 2 (positional keys, walk_t_of).
template <typename T>
__global__ void k_heap_replay_commit(Wk<T> w, WalkBuf<T> b, i64 cnt, i64 cnt_all, int kind) {
    if (threadIdx.x != 0) return;
    DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->need_walk || s->walk_closed) return;
    b.ctl->count = cnt; b.ctl->cur = 1; b.ctl->skip = 0;
    s->walk_lcount = cnt; s->walk_rcount = cnt_all; s->walk_rem = s->nbreak - s->walk_base;
    s->walk_J = -1; s->walk_done = 0; s->walk_fixn = 0;
    s->tie_round = kind;
}

// ---------------------------------------------------------------------------
// formk: corrections of the old blocks of WN1 for the variables that entered or
// left the free set (:1801-1851).  The reference re-gathers the lists for every
// (iy, jy) pair; here every listed row is read once into shared memory and all
// 3 col^2 products are accumulated from it (enter and leave sums kept apart).
// out: [gridDim][6][MMAX*MMAX] partials, finished by k_formk_delta_final.
// ---------------------------------------------------------------------------
#define LB_FD_GRID 592      // capacity of the partial buffer; the launch uses fd_grid<T>() blocks
// resident blocks per SM that the register count of k_formk_delta allows (REAL32: 3, REAL64: 2) x 148 SMs: one wave
template <typename T> constexpr int fd_blocks_per_sm() { return sizeof(T) == 4 ? 3 : 2; }
template <typename T> constexpr int fd_grid() { return 148 * fd_blocks_per_sm<T>(); }
#define LB_FD_ROWS 32
#define LB_FD_TB 4          // a thread owns a 4 x 4 block of (i, j) pairs of one of the three products
#define LB_FD_LD (2 * LB_MMAX + 4)   // row stride of the shared tile: 16-byte aligned rows
// Work decomposition: the (i, j) space of each product is cut into 4 x 4 blocks -- Wy_i Wy_j and Ws_i Ws_j only on and
// below the diagonal (formk uses jy <= iy, :1802-1826), Ws_i Wy_j in full (:1830-1851) -- and the 256 threads are
// NSUB copies of that block list, copy s taking every NSUB-th row of a 32-row tile.  Per row a thread reads its 4 + 4
// values with two 128-bit shared loads (the Wy and Ws halves of a row start at multiples of four columns) and does 16
// multiply-adds.  The rows of a tile are stored entering rows first, leaving rows after them (each in list order), and
// processed in two loops, so that no warp diverges on the row type.  The tile is double-buffered: the gather of the
// next 32 rows (lane = listed row, the 8 warps share the 2*col columns, all of a thread's loads in flight together) is
// issued before the multiply-adds of the current tile and lands in the other buffer after them -- one barrier per tile.
// The NSUB partial sums are added in copy order at the end.
__host__ __device__ inline int fd_tiles_1d(int col) { return (col + LB_FD_TB - 1) / LB_FD_TB; }
__host__ __device__ inline int fd_ntiles(int col) { const int t = fd_tiles_1d(col); return t * (t + 1) + t * t; }
#define LB_FD_MAXC ((2 * LB_MMAX + 7) / 8)   // columns per warp in the gather
template <typename T>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 ? 3 : 2)) k_formk_delta(Wk<T> w, const int* list, const SortCtl* ctl, T* out) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->do_delta) return;
    const int col = s->col, m = s->m, head0 = s->head - 1, col2 = 2 * col;
    const int cpad = fd_tiles_1d(col) * LB_FD_TB;   // the Ws half of a row starts here
    const i64 nel = ctl->count;
    const i64 chunk = (nel + gridDim.x - 1) / gridDim.x;
    const i64 beg = (i64)blockIdx.x * chunk;
    i64 end = beg + chunk; if (end > nel) end = nel;
    __shared__ __align__(16) T tile[2][LB_FD_ROWS][LB_FD_LD];
    __shared__ int nent[2];
    __shared__ T accs[6 * LB_MMAX * LB_MMAX];
    // this thread's block of pairs
    const int t1 = fd_tiles_1d(col), ntl = fd_ntiles(col), ntri = t1 * (t1 + 1) / 2;
    const int nsub = 256 / ntl;
    const int sub = threadIdx.x / ntl, tl = threadIdx.x % ntl;
    const bool active = sub < nsub;
    int blk, ti, tj;   // product (0: Wy.Wy, 1: Ws.Ws, 2: Ws.Wy), block row, block column
    if (tl < 2 * ntri) {
        blk = tl / ntri;
        int k = tl % ntri; ti = 0;
        while (k >= ti + 1) { k -= ti + 1; ++ti; }
        tj = k;
    } else { blk = 2; const int k = tl - 2 * ntri; ti = k / t1; tj = k % t1; }
    const int ca0 = ((blk == 0) ? 0 : cpad) + ti * LB_FD_TB;    // tile columns of the i factors
    const int cb0 = ((blk == 1) ? cpad : 0) + tj * LB_FD_TB;    // ... of the j factors
    T accE[LB_FD_TB][LB_FD_TB], accL[LB_FD_TB][LB_FD_TB];
#pragma unroll
    for (int a = 0; a < LB_FD_TB; ++a)
#pragma unroll
        for (int b = 0; b < LB_FD_TB; ++b) { accE[a][b] = (T)0; accL[a][b] = (T)0; }
    for (int e = threadIdx.x; e < 6 * LB_MMAX * LB_MMAX; e += 256) accs[e] = (T)0;
    // the columns a 4-wide block reads beyond col in either half (col not a multiple of 4), and the rows beyond a short
    // last tile, must be finite: zero both buffers once
    for (int e = threadIdx.x; e < 2 * LB_FD_ROWS * LB_FD_LD; e += 256) (&tile[0][0][0])[e] = (T)0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    // column c of the gather (c < col: Wy, else Ws) -> address offset of its ring column and its place in a tile row
    i64 goff[LB_FD_MAXC]; int tcol[LB_FD_MAXC];
#pragma unroll
    for (int q = 0; q < LB_FD_MAXC; ++q) {
        const int c = wrp + 8 * q;
        const int ring = c < col ? c : c - col;
        int pj = head0 + ring; if (pj >= m) pj -= m;
        goff[q] = (i64)pj * w.ldw;
        tcol[q] = c < col ? c : cpad + (c - col);
    }
    T pre[LB_FD_MAXC]; int pent = 0;
    // the loads of one tile: this lane's row
    auto fetch = [&](i64 r0) {
        const bool have = r0 + lane < end;
        const i64 var = have ? (i64)list[r0 + lane] : 0;
        pent = have ? (int)(w.state[var] & 1) : -1;   // free now => entering; -1: no row
#pragma unroll
        for (int q = 0; q < LB_FD_MAXC; ++q) {
            const int c = wrp + 8 * q;
            pre[q] = (have && c < col2) ? ((c < col) ? w.wy[goff[q] + var] : w.ws[goff[q] + var]) : (T)0;
        }
    };
    // entering rows first, leaving rows after them, each in list order
    auto place = [&](int buf) {
        const unsigned me = __ballot_sync(0xffffffffu, pent == 1), ml = __ballot_sync(0xffffffffu, pent == 0);
        const unsigned below = (1u << lane) - 1u;
        const int ne = __popc(me);
        const int row = pent == 1 ? __popc(me & below) : ne + __popc(ml & below);
        if (pent >= 0) {
#pragma unroll
            for (int q = 0; q < LB_FD_MAXC; ++q) if (wrp + 8 * q < col2) tile[buf][row][tcol[q]] = pre[q];
        }
        if (threadIdx.x == 0) nent[buf] = ne;
    };
    if (beg < end) { fetch(beg); place(0); }
    __syncthreads();
    int buf = 0;
    for (i64 r0 = beg; r0 < end; r0 += LB_FD_ROWS, buf ^= 1) {
        const int nr = (int)((end - r0 < LB_FD_ROWS) ? (end - r0) : LB_FD_ROWS);
        const bool more = r0 + LB_FD_ROWS < end;
        if (more) fetch(r0 + LB_FD_ROWS);
        if (active) {
            const int ne = nent[buf];
            for (int r = sub; r < ne; r += nsub) {
                T av[LB_FD_TB], bv[LB_FD_TB];
                const T* row = &tile[buf][r][0];
                if (sizeof(T) == 4) {
                    const float4 qa = *reinterpret_cast<const float4*>(row + ca0), qb = *reinterpret_cast<const float4*>(row + cb0);
                    av[0] = (T)qa.x; av[1] = (T)qa.y; av[2] = (T)qa.z; av[3] = (T)qa.w;
                    bv[0] = (T)qb.x; bv[1] = (T)qb.y; bv[2] = (T)qb.z; bv[3] = (T)qb.w;
                } else {
                    const double2 qa0 = *reinterpret_cast<const double2*>(row + ca0), qa1 = *reinterpret_cast<const double2*>(row + ca0 + 2);
                    const double2 qb0 = *reinterpret_cast<const double2*>(row + cb0), qb1 = *reinterpret_cast<const double2*>(row + cb0 + 2);
                    av[0] = (T)qa0.x; av[1] = (T)qa0.y; av[2] = (T)qa1.x; av[3] = (T)qa1.y;
                    bv[0] = (T)qb0.x; bv[1] = (T)qb0.y; bv[2] = (T)qb1.x; bv[3] = (T)qb1.y;
                }
#pragma unroll
                for (int a = 0; a < LB_FD_TB; ++a)
#pragma unroll
                    for (int b = 0; b < LB_FD_TB; ++b) accE[a][b] = accE[a][b] + av[a] * bv[b];
            }
            for (int r = ne + sub; r < nr; r += nsub) {
                T av[LB_FD_TB], bv[LB_FD_TB];
                const T* row = &tile[buf][r][0];
                if (sizeof(T) == 4) {
                    const float4 qa = *reinterpret_cast<const float4*>(row + ca0), qb = *reinterpret_cast<const float4*>(row + cb0);
                    av[0] = (T)qa.x; av[1] = (T)qa.y; av[2] = (T)qa.z; av[3] = (T)qa.w;
                    bv[0] = (T)qb.x; bv[1] = (T)qb.y; bv[2] = (T)qb.z; bv[3] = (T)qb.w;
                } else {
                    const double2 qa0 = *reinterpret_cast<const double2*>(row + ca0), qa1 = *reinterpret_cast<const double2*>(row + ca0 + 2);
                    const double2 qb0 = *reinterpret_cast<const double2*>(row + cb0), qb1 = *reinterpret_cast<const double2*>(row + cb0 + 2);
                    av[0] = (T)qa0.x; av[1] = (T)qa0.y; av[2] = (T)qa1.x; av[3] = (T)qa1.y;
                    bv[0] = (T)qb0.x; bv[1] = (T)qb0.y; bv[2] = (T)qb1.x; bv[3] = (T)qb1.y;
                }
#pragma unroll
                for (int a = 0; a < LB_FD_TB; ++a)
#pragma unroll
                    for (int b = 0; b < LB_FD_TB; ++b) accL[a][b] = accL[a][b] + av[a] * bv[b];
            }
        }
        if (more) place(buf ^ 1);
        __syncthreads();
    }
    // add the NSUB copies in copy order (fixed), then one partial per block of the grid
    for (int sb = 0; sb < nsub; ++sb) {
        if (active && sub == sb) {
#pragma unroll
            for (int a = 0; a < LB_FD_TB; ++a)
#pragma unroll
                for (int b = 0; b < LB_FD_TB; ++b) {
                    const int i = ti * LB_FD_TB + a, j = tj * LB_FD_TB + b;
                    if (i < col && j < col) {
                        const int e = i + j * LB_MMAX;
                        accs[blk * LB_MMAX * LB_MMAX + e] = accs[blk * LB_MMAX * LB_MMAX + e] + accE[a][b];
                        accs[(3 + blk) * LB_MMAX * LB_MMAX + e] = accs[(3 + blk) * LB_MMAX * LB_MMAX + e] + accL[a][b];
                    }
                }
        }
        __syncthreads();
    }
    T* o = out + (i64)blockIdx.x * (6 * LB_MMAX * LB_MMAX);
    for (int e = threadIdx.x; e < 6 * LB_MMAX * LB_MMAX; e += 256) o[e] = accs[e];
}

// sum the per-block partials in block order (one thread per entry, serial: fixed order)
template <typename T>
__global__ void __launch_bounds__(256) k_formk_delta_final(Wk<T> w, const T* parts, int nparts, T* delta) {
    const DevState<T>* s = w.s;
    if (!s->go || !s->in_body || !s->do_delta) return;
    const int col = s->col;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 6 * LB_MMAX * LB_MMAX) return;
    const int ij = e % (LB_MMAX * LB_MMAX);
    const int i = ij % LB_MMAX, j = ij / LB_MMAX;
    if (i >= col || j >= col) return;
    T acc = (T)0;
    for (int p = 0; p < nparts; ++p) acc = acc + parts[(i64)p * (6 * LB_MMAX * LB_MMAX) + e];
    delta[e] = acc;
}
