// Tensor memory (TMEM) as a per-thread scratchpad -- helpers shared by the BK1 emitter (large mechanisms:
// exp(+-g_k) and third-body sums that do not fit shared memory) and the BK2 tensor-memory kernel.
//
// Blackwell's 256 KB of tensor memory per SM is addressed as 128 lanes x 512 32-bit columns.  With the 32x32b
// shape, lane i of warp w reads / writes consecutive columns of TMEM lane 32 (w mod 4) + i, i.e. every thread owns
// one TMEM lane (shared with the thread of warp w + 4 at the same lane index, which takes other columns).  No
// tensor-core instruction is involved: tcgen05.ld / tcgen05.st (SASS LDTM / STTM) move registers <-> TMEM on a
// datapath of their own, beside the shared-memory (LSU) pipe.  All column addresses are warp-uniform and the
// instructions are .sync.aligned: every lane of the warp must execute them (no divergent control flow around).
#pragma once
#include "kx_math.cuh"

// loads are asynchronous: destination registers are valid after kx_tm_wait_ld(); stores are complete (visible
// to later loads of the same thread) after kx_tm_wait_st()
KX_DEVICE void kx_tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
KX_DEVICE void kx_tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one double = 2 columns
KX_DEVICE void kx_tm_ld_d(unsigned taddr, unsigned& lo, unsigned& hi)
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr));
}
KX_DEVICE void kx_tm_st_d(unsigned taddr, double v)
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(__double2loint(v)),
               "r"(__double2hiint(v))
               : "memory");
}
// after kx_tm_wait_ld(): the empty asm keeps every consumer of the value behind the wait for the compiler
KX_DEVICE double kx_tm_pin(unsigned lo, unsigned hi)
{
  asm volatile("" : "+r"(lo), "+r"(hi));
  return __hiloint2double((int)hi, (int)lo);
}

// COLS (a power of two, 32..512) columns for this CTA; the CTAs resident on one SM must not ask for more than 512
// together (the callers' shared-memory / register footprints see to that; a CTA too many would wait in
// tcgen05.alloc until another has released).  `slot` = a 32-bit word in shared memory that receives the TMEM base
// address.  Call from every thread; contains a CTA-wide barrier.
template <int COLS = 512>
KX_DEVICE unsigned kx_tm_alloc_all(unsigned* slot)
{
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(slot)), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *reinterpret_cast<volatile unsigned*>(slot);
}
template <int COLS = 512>
KX_DEVICE void kx_tm_free_all(unsigned base)
{
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
  }
}
