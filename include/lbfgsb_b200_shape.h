/* Fixed reduction shape shared by the CUDA kernels and by the oracle's
 * "device order" summation mode (oracle/lbfgsb_oracle.cpp: device_order_sum).
 *
 * Every O(n) sum in the engine is formed the same way, independent of n and of
 * the GPU it runs on, so that results are bit-reproducible run to run and can
 * be replayed on a CPU:
 *
 *   - LBFGSB_GRID blocks of LBFGSB_BLOCK threads; block b walks tiles
 *     b, b+GRID, b+2*GRID, ...; a tile is BLOCK*VEC*UNROLL consecutive
 *     variables with VEC = 16/sizeof(real) (one 128-bit load per thread);
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
#define LBFGSB_UNROLL 4         /* 128-bit loads in flight per thread per stream */
#define LBFGSB_GRID 592         /* 148 SMs x 4 blocks                            */
#define LBFGSB_FINAL_BLOCK 32   /* one warp finishes one reduction slot          */

#endif
