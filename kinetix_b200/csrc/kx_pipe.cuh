// kx_pipe.cuh -- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers shared by the BK2 kernels that
// stream their coefficient tables through a ring of shared-memory stages.
#pragma once
#include <cstdint>
#include "kx_math.cuh"

// ---- TMA bulk copy + mbarrier plumbing -----------------------------------------------------------
KX_DEVICE unsigned kx_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
KX_DEVICE void kx_mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kx_smem_addr(bar)), "r"(count));
}
KX_DEVICE void kx_mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(kx_smem_addr(bar)) : "memory");
}
KX_DEVICE void kx_bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar)
{
  const unsigned b = kx_smem_addr(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(kx_smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}
KX_DEVICE void kx_mbar_wait(uint64_t* bar, unsigned parity)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "KX_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra KX_DONE_%=;\n\t"
      "bra KX_WAIT_%=;\n\t"
      "KX_DONE_%=:\n\t}" ::"r"(kx_smem_addr(bar)), "r"(parity)
      : "memory");
}

