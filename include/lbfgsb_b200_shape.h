/* Fixed reduction shape shared by the CUDA kernels and by the oracle's
 * "device order" summation mode (oracle/lbfgsb_oracle.cpp: device_order_sum).
 *
 * Every O(n) sum in the engine is formed the same way, independent of n and of
 * the GPU it runs on, so that results are bit-reproducible run to run and can
 * be replayed on a CPU:
 *
 *   - LBFGSB_GRID blocks of LBFGSB_BLOCK threads; block b walks tiles
 *     b, b+GRID, b+2*GRID, ...; a tile is BLOCK*VEC*UNROLL consecutive
 *     variables.  VEC = 2 for both real kinds: one 128-bit load per thread per
 *     stream in REAL64, one 64-bit load in REAL32, where UNROLL is doubled
 *     instead -- a sub-tile of BLOCK*VEC variables is then 4 KB (REAL64) or
 *     2 KB (REAL32) per stream, so that the 2*m + 8 streams of one sub-tile fit
 *     twice in shared memory for every (kind, m) the fused passes support and
 *     ALL threads of the block work on every staged sub-tile (tma_pipe.cuh);
 *   - in a tile, thread t owns the VEC variables at t*VEC + k*(BLOCK*VEC),
 *     k = 0..UNROLL-1, and adds their terms serially into one accumulator that
 *     lives across all of the block's tiles;
 *   - lanes combine by xor-butterfly shuffles (offsets 16,8,4,2,1), the warp
 *     sums are added serially in warp order by one thread -> block partial;
 *   - the GRID block partials of one sum are combined by one warp
 *     (LBFGSB_FINAL_BLOCK = 32 threads): lane t adds partials t, t+32, ...
 *     serially, then the same butterfly.
 *
 * No floating-point atomics anywhere.
 */
#ifndef LBFGSB_B200_SHAPE_H
#define LBFGSB_B200_SHAPE_H

#define LBFGSB_BLOCK 256        /* threads per streaming block                  */
#define LBFGSB_VEC_F64 2        /* variables per thread per load, REAL64 (128-bit) */
#define LBFGSB_VEC_F32 2        /* variables per thread per load, REAL32 (64-bit)  */
#define LBFGSB_UNROLL_F64 4     /* loads in flight per thread per stream, REAL64   */
#define LBFGSB_UNROLL_F32 8     /* the same, REAL32: a tile is 4096 variables      */
#define LBFGSB_VEC(real_bytes) ((real_bytes) == 8 ? LBFGSB_VEC_F64 : LBFGSB_VEC_F32)
#define LBFGSB_UNROLL(real_bytes) ((real_bytes) == 8 ? LBFGSB_UNROLL_F64 : LBFGSB_UNROLL_F32)
#define LBFGSB_GRID 592         /* 148 SMs x 4 blocks                            */
#define LBFGSB_FINAL_BLOCK 32   /* one warp finishes one reduction slot          */

#endif
