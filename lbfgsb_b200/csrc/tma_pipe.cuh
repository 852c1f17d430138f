// TMA-staged streaming pass for the kernels that read the 2*col S/Y columns.
//
// A pass over the variables touches up to 2*col + 8 input streams (the ring columns of
// Wy/Ws are n apart in memory).  Issuing those as per-thread global loads costs two registers
// per stream per thread and serialises the load latency behind every use.  Here a producer
// warp issues 1-D bulk copies (cp.async.bulk, SASS UBLKCP) of one sub-tile of every stream
// into a two-stage shared-memory ring while eight consumer warps compute from the other
// stage; a full[] / empty[] mbarrier pair per stage (expect_tx / complete_tx, consumer release)
// is the only synchronisation.  Registers hold only the accumulators, and 100 KB per SM is in
// flight regardless of occupancy.
//
// The thread -> element mapping and the order in which a thread meets its elements are
// exactly those of LB_FOR_TILES (include/lbfgsb_b200_shape.h): a stage is one k-sub-tile
// (SUBT = 256: BLOCK*VEC consecutive variables) or, when 2*col columns of a sub-tile do not fit
// twice in shared memory (m > 10), half of it (SUBT = 128), in which case only the 128 threads
// that own those elements work on the stage.  Each thread reads back only its own VEC elements
// of every stream, so a ragged last stage is filled by guarded per-thread loads into the same
// slots and needs no extra synchronisation.
#pragma once
#include "common.cuh"

namespace tma {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared 1-D bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace tma

#define LB_PIPE_MAXSRC 50
#define LB_PIPE_STAGES 2

// Stream table of one pass (shared memory, filled by thread 0).
struct PipeSrc {
    const char* p[LB_PIPE_MAXSRC];     // base address of the stream
    unsigned off[LB_PIPE_MAXSRC];      // byte offset of its slot inside a stage
    unsigned char esz[LB_PIPE_MAXSRC]; // element size: sizeof(real), 4 (int) or 1 (state byte)
    int nsrc;
    unsigned stage_bytes;              // size of one stage (multiple of 128)
    unsigned stage_tx;                 // bytes one full stage transfers
};

template <typename T, int SUBT>
struct PipeGeom {
    static constexpr int VEC = Real<T>::VEC;
    static constexpr int HALVES = LBFGSB_BLOCK / SUBT;
    static constexpr int IPT = Real<T>::UNROLL * HALVES;            // stages per tile
    static constexpr int ELEMS = SUBT * VEC;                      // variables per stage
    static constexpr unsigned REAL_SLOT = ELEMS * sizeof(T);
    static constexpr unsigned INT_SLOT = ELEMS * 4;
    static constexpr unsigned BYTE_SLOT = ELEMS;
    __host__ __device__ static constexpr unsigned slot(int esz) { return (unsigned)(ELEMS * esz); }
};

// append a stream to the table (thread 0, before the pass)
__device__ __forceinline__ void pipe_add(PipeSrc* ps, const void* p, int esz, unsigned slot_bytes) {
    const int k = ps->nsrc;
    ps->p[k] = (const char*)p;
    ps->esz[k] = (unsigned char)esz;
    ps->off[k] = ps->stage_bytes;
    ps->stage_bytes += slot_bytes;
    ps->stage_tx += slot_bytes;
    ps->nsrc = k + 1;
}
__device__ __forceinline__ void pipe_begin(PipeSrc* ps) { ps->nsrc = 0; ps->stage_bytes = 0; ps->stage_tx = 0; }
__device__ __forceinline__ void pipe_end(PipeSrc* ps) { ps->stage_bytes = (ps->stage_bytes + 127u) & ~127u; }

// shared-memory bytes of a pass with `nreal` real streams, `nint` int streams, `nbyte` byte streams
template <typename T, int SUBT>
__host__ __device__ constexpr unsigned pipe_smem_bytes(int nreal, int nint, int nbyte) {
    return LB_PIPE_STAGES * (((unsigned)nreal * PipeGeom<T, SUBT>::REAL_SLOT + (unsigned)nint * PipeGeom<T, SUBT>::INT_SLOT +
                               (unsigned)nbyte * PipeGeom<T, SUBT>::BYTE_SLOT + 127u) & ~127u) + 128u;
}

// read this thread's VEC elements of a real / int / byte slot
template <typename T>
__device__ __forceinline__ void lds_real(const char* sm, unsigned off, int lt, T (&out)[Real<T>::VEC]) {
    const typename Real<T>::vec_t q = *reinterpret_cast<const typename Real<T>::vec_t*>(sm + off + lt * (int)sizeof(typename Real<T>::vec_t));
    const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int v = 0; v < Real<T>::VEC; ++v) out[v] = e[v];
}
// the same through a volatile pointer: the load is issued where it is written (software prefetch of the next column)
template <typename T>
__device__ __forceinline__ void lds_real_v(const char* sm, unsigned off, int lt, T (&out)[Real<T>::VEC]) {
    const volatile T* e = reinterpret_cast<const volatile T*>(sm + off + lt * (int)sizeof(typename Real<T>::vec_t));
    if (Real<T>::VEC == 2 && sizeof(T) == 4) {
        float a, b;
        asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(tma::smem_u32((const void*)e)));
        out[0] = (T)a; out[1] = (T)b;
    } else {
#pragma unroll
        for (int v = 0; v < Real<T>::VEC; ++v) out[v] = e[v];
    }
}
template <typename T>
__device__ __forceinline__ void lds_int(const char* sm, unsigned off, int lt, int (&out)[Real<T>::VEC]) {
    const typename Real<T>::ivec_t q = *reinterpret_cast<const typename Real<T>::ivec_t*>(sm + off + lt * (Real<T>::VEC * 4));
    const int* e = reinterpret_cast<const int*>(&q);
#pragma unroll
    for (int v = 0; v < Real<T>::VEC; ++v) out[v] = e[v];
}
template <typename T>
__device__ __forceinline__ void lds_byte(const char* sm, unsigned off, int lt, int (&out)[Real<T>::VEC]) {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(sm + off + lt * Real<T>::VEC);
#pragma unroll
    for (int v = 0; v < Real<T>::VEC; ++v) out[v] = b[v];
}

// The pass.  body(base, sm, lt): `base` = first variable of this thread in the stage, `sm` = stage
// buffer, `lt` = thread index inside the stage.  Elements at or beyond n read as zero.
// `stages` must be 128-byte aligned; `bars` are 2*LB_PIPE_STAGES mbarriers in shared memory
// (full[], then empty[]).
//
// The CTA has LB_TMA_THREADS = LBFGSB_BLOCK + 32 threads: warps 0..7 are the consumers (the
// LBFGSB_BLOCK threads of the fixed reduction shape), warp 8 is the producer -- its lanes issue
// the bulk copies of a stage in parallel (one stream per lane) as soon as all consumer warps
// have released the buffer (empty[] barrier).  Consumer warps never wait for one another.
// All threads of the block must call.
#define LB_TMA_THREADS (LBFGSB_BLOCK + 32)

template <typename T, int SUBT, typename Body>
__device__ __forceinline__ void tma_pass(i64 n, const PipeSrc* ps, char* stages, unsigned long long* bars, Body body) {
    typedef PipeGeom<T, SUBT> G;
    constexpr int VEC = G::VEC;
    constexpr int NCW = LBFGSB_BLOCK / 32;   // consumer warps
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    unsigned long long* full = bars;
    unsigned long long* empty = bars + LB_PIPE_STAGES;
    const i64 tile = (i64)LBFGSB_BLOCK * VEC * Real<T>::UNROLL;
    const i64 ntiles = (n + tile - 1) / tile;
    const i64 b = blockIdx.x;
    i64 nitems = 0;
    if (ntiles > b) {
        const i64 ntl = (ntiles - b + LBFGSB_GRID - 1) / LBFGSB_GRID;
        const i64 tl_last = b + (ntl - 1) * LBFGSB_GRID;
        const i64 rem = n - tl_last * tile;
        i64 last = (rem + G::ELEMS - 1) / G::ELEMS;
        if (last > G::IPT) last = G::IPT;
        nitems = (ntl - 1) * G::IPT + last;
    }
    if (tid == 0) {
        for (int s = 0; s < LB_PIPE_STAGES; ++s) { tma::mbar_init(&full[s], 1); tma::mbar_init(&empty[s], NCW); }
        tma::fence_barrier_init();
    }
    __syncthreads();
    const unsigned stage_bytes = ps->stage_bytes;
    const int nsrc = ps->nsrc;
    auto stage_base = [&](i64 q) -> i64 {
        const i64 tl = b + (q / G::IPT) * LBFGSB_GRID;
        const int r = (int)(q % G::IPT);
        return tl * tile + (i64)r * G::ELEMS;        // k*(BLOCK*VEC) + h*(SUBT*VEC) == r*ELEMS
    };
    if (wid == NCW) {
        // ---- producer warp ----
        const unsigned stage_tx = ps->stage_tx;
        for (i64 q = 0; q < nitems; ++q) {
            const int st = (int)(q % LB_PIPE_STAGES);
            if (q >= LB_PIPE_STAGES) tma::mbar_wait(&empty[st], (unsigned)(((q / LB_PIPE_STAGES) - 1) & 1));
            const i64 sb = stage_base(q);
            if (sb + G::ELEMS <= n) {
                if (lane == 0) tma::mbar_expect_tx(&full[st], stage_tx);
                __syncwarp();
                char* dst = stages + (size_t)st * stage_bytes;
                for (int s = lane; s < nsrc; s += 32) {
                    const unsigned e = ps->esz[s];
                    tma::bulk_g2s(dst + ps->off[s], ps->p[s] + sb * e, (unsigned)G::ELEMS * e, &full[st]);
                }
            } else if (lane == 0) {
                tma::mbar_arrive(&full[st]);   // ragged stage: filled by the consumers themselves
            }
            __syncwarp();
        }
        return;
    }
    // ---- consumer warps ----
    for (i64 q = 0; q < nitems; ++q) {
        const int st = (int)(q % LB_PIPE_STAGES);
        tma::mbar_wait(&full[st], (unsigned)((q / LB_PIPE_STAGES) & 1));
        const i64 sb = stage_base(q);
        const int h = (int)(q % G::HALVES);
        if (G::HALVES == 1 || tid / SUBT == h) {
            const int lt = tid % SUBT;
            char* sm = stages + (size_t)st * stage_bytes;
            const i64 base = sb + (i64)lt * VEC;
            if (sb + G::ELEMS > n) {
                // ragged stage: guarded loads of this thread's own elements into its own slots
                for (int s = 0; s < nsrc; ++s) {
                    const unsigned e = ps->esz[s];
                    char* d = sm + ps->off[s] + (unsigned)lt * VEC * e;
                    const char* g = ps->p[s] + base * e;
                    for (unsigned v = 0; v < (unsigned)VEC; ++v)
                        for (unsigned c = 0; c < e; ++c) d[v * e + c] = (base + v < n) ? g[v * e + c] : (char)0;
                }
            }
            if (base < n) body(base, (const char*)sm, lt);
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&empty[st]);
    }
}
