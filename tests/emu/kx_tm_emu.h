// TEST INFRASTRUCTURE -- host model of csrc/kx_tm.cuh: tensor memory as 128 lanes x 512 32-bit columns; the 32x32b
// access of lane i of warp w goes to TMEM lane 32 (w mod 4) + i at the (warp-uniform) column of the address.
// The harness poisons a thread's lane with NaN patterns before the thread runs, so a slot that is read before it
// was written (or written by another warp group's column block) shows up as NaN in the results.
#pragma once
#include "cuda_emu.h"

static uint32_t emu_tmem[128][512];
static unsigned emu_tm_max_col = 0;

static inline void emu_tm_decode(unsigned taddr, unsigned& lane, unsigned& col)
{
  lane = (taddr >> 16) + (threadIdx.x & 31u);
  col = taddr & 0xffffu;
  if (lane >= 128 || col + 1 >= 512) {
    fprintf(stderr, "emu: TMEM address out of range (lane %u, column %u)\n", lane, col);
    abort();
  }
  if (lane != (((threadIdx.x >> 5) & 3u) << 5) + (threadIdx.x & 31u)) {
    fprintf(stderr, "emu: thread %u touches TMEM lane %u outside its warp's quadrant\n", threadIdx.x, lane);
    abort();
  }
  emu_tm_max_col = max(emu_tm_max_col, col + 1);
}
static inline void kx_tm_wait_ld() {}
static inline void kx_tm_wait_st() {}
static inline void kx_tm_ld_d(unsigned taddr, unsigned& lo, unsigned& hi)
{
  unsigned lane, col;
  emu_tm_decode(taddr, lane, col);
  lo = emu_tmem[lane][col];
  hi = emu_tmem[lane][col + 1];
}
static inline void kx_tm_st_d(unsigned taddr, double v)
{
  unsigned lane, col;
  emu_tm_decode(taddr, lane, col);
  emu_tmem[lane][col] = (uint32_t)__double2loint(v);
  emu_tmem[lane][col + 1] = (uint32_t)__double2hiint(v);
}
static inline double kx_tm_pin(unsigned lo, unsigned hi) { return __hiloint2double((int)hi, (int)lo); }

// the allocator hands out column 0 (or, for half allocations, alternating halves like two co-resident CTAs)
template <int COLS = 512>
static inline unsigned kx_tm_alloc_all(unsigned* slot)
{
  *slot = COLS == 512 ? 0u : (blockIdx.x & 1u) * (unsigned)COLS;
  return *slot;
}
template <int COLS = 512>
static inline void kx_tm_free_all(unsigned)
{
}
