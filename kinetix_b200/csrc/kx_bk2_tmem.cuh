// kx_bk2_tmem.cuh -- BK2 (mixture-averaged conductivity, viscosity, rho*D_km): KX_P states per thread, the
// per-state sums S_k in TENSOR MEMORY, the mole fractions X_k in shared memory.  FP64 only.
//
// Same arithmetic as csrc/kx_bk2.cuh (reference benchmark/okl/transportProps.okl:11-49 around
// kinetix/core/mix_transport.py:474-626).  Design, and why (profiles/ncu_r01_bk2_*.txt, ncu_r02_bk2_*.txt):
//   * The one-state-per-thread kernel is bound by the shared-memory data pipe, not by the FP64 pipe: every
//     species pair costs 3 LDS.128 (the quartic's coefficients, a broadcast) = 6 LSU wavefronts per warp against
//     9 DFMA.  So every coefficient fetched from shared memory is used for KX_P = 2 states here.
//   * Blackwell's tensor memory is 256 KB per SM that this kernel would leave idle.  It is addressed as
//     128 lanes x 512 32-bit columns, and warp w may touch lanes 32 (w % 4) .. +31: exactly a per-thread
//     scratchpad.  The running sums S_k live there (thread = lane, 2 * KX_NS columns per state,
//     tcgen05.st / tcgen05.ld 32x32b, SASS STTM / LDTM), the X_k stay in shared memory as [k][state].
//   * The Wilke sum uses a rank-KX_WR factorisation of its mass-factor matrix: 6 N r instead of 3 N^2 multiply-adds.
//   * All tables (species quartics, Wilke factors, diffusion tiles) are ONE stream of chunks that goes through a
//     KX_STAGES-deep ring of TMA bulk copies with full/empty mbarriers.
//   * Persistent CTAs (one per SM) loop over batches of KX_BK2_BLOCK * KX_P states.
//   * Round 2: the row blocks of the pair loop run in DESCENDING order.  Block kb then receives its column
//     contributions (from the blocks above it) BEFORE its own row pass, so its S_k are final when that pass ends:
//     rho*D_km of the block is formed and stored right there, while the FP64 pipe works on the next block (the
//     former epilogue -- reciprocals + N row stores at < 35 % pipe utilisation -- is gone), the top block never
//     goes to tensor memory at all (KX_NS = KX_NP - KX_TB: a 129-species mechanism fits two warps per lane
//     quadrant), and the X rows of a finished block are dead, so the NEXT batch's mass fractions are fetched into
//     them with cp.async (LDGSTS) while this batch still computes: the next batch starts with a pass over shared
//     memory instead of 2 x N dependent DRAM round trips (the former prologue).
//
// The including translation unit defines KX_N, KX_NP (multiple of KX_TB), KX_TB, KX_P, KX_BK2_BLOCK (threads, a
// multiple of 32: warp w uses TMEM lane quadrant w % 4, column slot w / 4), KX_NS (doubles per state in TMEM,
// = KX_NP - KX_TB), KX_STAGES, KX_CHUNK_MAX (reals per stage), KX_WR (even: rank of the Wilke factorisation), the
// chunk layout KX_N_CHUNKS / KX_NVC / KX_NUC / KX_VROWS / KX_UROWS (below), KX_RCP_DIFF and the tables
//   __constant__ double kx_rcpM[KX_N], kx_M[KX_N]          1/M_k, M_k
//   __constant__ int    kx_chunk_off[KX_N_CHUNKS + 1]      chunk boundaries in kx_bk2_stream (reals)
//   __device__   double kx_bk2_stream[]                    the concatenated chunks, 16-byte aligned each
#pragma once
#include <cstdint>
#include "kx_math.cuh"
#include "kx_pipe.cuh"
#include "kx_tm.cuh"

#define KX_NB (KX_NP / KX_TB)
// 1/D_kj in the pair loops: MUFU.RCP64H seed + ONE quadratic Newton step (2 FP64 instructions, relative error
// <= 1e-12: the seed is good to 9.9e-7).  The terms X_j/D_kj are all positive, so S_k and rho*D_km inherit that
// bound, 100x inside the 1e-10 parity contract (measured 2.6e-13 against the reference, 4.7e-14 with the cubic
// step of kx_rcp; +3.7 % throughput).  -DKX_BK2_FULL_RCP restores the cubic step.
#ifdef KX_BK2_FULL_RCP
#define KX_PAIR_RCP kx_rcp
#else
#define KX_PAIR_RCP kx_rcp_fast
#endif
#define KX_N_DTILES (KX_NB * (KX_NB + 1) / 2)
#ifndef KX_COL_UNROLL
#define KX_COL_UNROLL 1     // columns per iteration of the pair loop's body (1: ~3.5 KB of code, L0-resident)
#endif
static_assert(sizeof(real) == 8, "the tensor-memory BK2 kernel is FP64 only");

// species quartic in ln T, Estrin form (dependency depth 3 instead of Horner's 4; these sit in the latency-bound
// prologue and in the per-species epilogue of the Wilke pass)
KX_DEVICE real kx_quartic(const real* __restrict__ c, real l, real l2, real l4)
{
  return fma(c[4], l4, fma(fma(c[3], l, c[2]), l2, fma(c[1], l, c[0])));
}
// The coefficient stream of one batch: KX_N_CHUNKS chunks of the concatenated table kx_bk2_stream, chunk c =
// reals [kx_chunk_off[c], kx_chunk_off[c + 1]):
//   KX_NVC chunks of species rows, first pass (KX_VROWS rows x (12 + KX_WR): conductivity quartic, viscosity
//          quartic, M^-1/4, -, row of the Wilke factor V)
//   KX_NUC chunks of the Wilke factor U (KX_UROWS rows x (KX_WR + 6): U row, viscosity quartic, M^-1/4)
//   the lower-triangular diffusion tiles (KX_TB columns of KX_COL reals: the column's KX_TB quartics, rows paired and
//          interleaved for 16-byte loads): kb = KX_NB-1 .. 0, jb = 0 .. kb
// Rows per chunk are multiples of KX_TB; padded species rows hold quartics = 1, M^-1/4 = 1, U = V = 0.
KX_DEVICE const real* kx_chunk_src(int c) { return kx_bk2_stream + kx_chunk_off[c]; }
KX_DEVICE unsigned kx_chunk_bytes(int c) { return (unsigned)((kx_chunk_off[c + 1] - kx_chunk_off[c]) * sizeof(real)); }

// ---- tensor memory as a per-thread scratchpad ------------------------------------------------------
// 32x32b shape: lane i of the warp reads / writes N consecutive 32-bit columns of TMEM lane (quadrant base + i).
// All accesses are warp-uniform in their column address.  Loads are asynchronous: the destination registers
// are valid after kx_tm_wait_ld(); kx_tm_unpack() pins the consumers behind that wait for the compiler.
#define KX_TM_LD(N, REGS, ...)                                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #N ".b32 {" REGS "}, [%" #N "];" : __VA_ARGS__ : "r"(taddr))
KX_DEVICE void kx_tm_ld2(unsigned taddr, unsigned* r)
{
  KX_TM_LD(2, "%0,%1", "=r"(r[0]), "=r"(r[1]));
}
KX_DEVICE void kx_tm_ld4(unsigned taddr, unsigned* r)
{
  KX_TM_LD(4, "%0,%1,%2,%3", "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]));
}
KX_DEVICE void kx_tm_ld8(unsigned taddr, unsigned* r)
{
  KX_TM_LD(8, "%0,%1,%2,%3,%4,%5,%6,%7", "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
           "=r"(r[6]), "=r"(r[7]));
}
KX_DEVICE void kx_tm_ld16(unsigned taddr, unsigned* r)
{
  KX_TM_LD(16, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15", "=r"(r[0]), "=r"(r[1]), "=r"(r[2]),
           "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
           "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]));
}
KX_DEVICE void kx_tm_st2(unsigned taddr, const unsigned* r)
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
KX_DEVICE void kx_tm_st4(unsigned taddr, const unsigned* r)
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
KX_DEVICE void kx_tm_st8(unsigned taddr, const unsigned* r)
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
KX_DEVICE void kx_tm_st16(unsigned taddr, const unsigned* r)
{
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// kx_tm_wait_ld() / kx_tm_wait_st(): kx_tm.cuh

// issue the loads of N consecutive doubles (column `taddr`, 2 columns per double) in pieces of 8 / 4 / 2 / 1
template <int N, int OFF = 0>
KX_DEVICE void kx_tm_load(unsigned taddr, unsigned* r)
{
  if constexpr (N >= 8) {
    kx_tm_ld16(taddr + 2 * OFF, r + 2 * OFF);
    kx_tm_load<N - 8, OFF + 8>(taddr, r);
  } else if constexpr (N >= 4) {
    kx_tm_ld8(taddr + 2 * OFF, r + 2 * OFF);
    kx_tm_load<N - 4, OFF + 4>(taddr, r);
  } else if constexpr (N >= 2) {
    kx_tm_ld4(taddr + 2 * OFF, r + 2 * OFF);
    kx_tm_load<N - 2, OFF + 2>(taddr, r);
  } else if constexpr (N == 1) {
    kx_tm_ld2(taddr + 2 * OFF, r + 2 * OFF);
  }
}
// after kx_tm_wait_ld(): turn the raw registers into doubles (the empty asm keeps every use behind the wait)
template <int N>
KX_DEVICE void kx_tm_unpack(unsigned (&r)[2 * N], real (&d)[N])
{
#pragma unroll
  for (int i = 0; i < N; i++) {
    asm volatile("" : "+r"(r[2 * i]), "+r"(r[2 * i + 1]));
    d[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
  }
}
template <int N, int OFF = 0>
KX_DEVICE void kx_tm_store_raw(unsigned taddr, const unsigned* r)
{
  if constexpr (N >= 8) {
    kx_tm_st16(taddr + 2 * OFF, r + 2 * OFF);
    kx_tm_store_raw<N - 8, OFF + 8>(taddr, r);
  } else if constexpr (N >= 4) {
    kx_tm_st8(taddr + 2 * OFF, r + 2 * OFF);
    kx_tm_store_raw<N - 4, OFF + 4>(taddr, r);
  } else if constexpr (N >= 2) {
    kx_tm_st4(taddr + 2 * OFF, r + 2 * OFF);
    kx_tm_store_raw<N - 2, OFF + 2>(taddr, r);
  } else if constexpr (N == 1) {
    kx_tm_st2(taddr + 2 * OFF, r + 2 * OFF);
  }
}
template <int N>
KX_DEVICE void kx_tm_store(unsigned taddr, const real (&d)[N])
{
  unsigned r[2 * N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    r[2 * i] = (unsigned)__double2loint(d[i]);
    r[2 * i + 1] = (unsigned)__double2hiint(d[i]);
  }
  kx_tm_store_raw<N>(taddr, r);
}

// asynchronous 8-byte global -> shared copy (LDGSTS), no commit: the caller commits one group per row block
KX_DEVICE void kx_cp_async8_nc(unsigned smem_addr, const void* gptr)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
// predicated shared-memory store (no branch in the instruction stream)
KX_DEVICE void kx_sts_if(unsigned addr, double v, bool p)
{
  asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n @q st.shared.f64 [%0], %1;\n}" ::"r"(addr), "d"(v), "r"((unsigned)p) : "memory");
}
template <int V> struct kx_int { static constexpr int value = V; };    // compile-time tag for generic lambdas
KX_DEVICE void kx_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
KX_DEVICE void kx_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ST: storage type of the state / result buffers (reference: dfloat); P: states per thread -- KX_P for full batches,
// 1 for small launches (fewer than one KX_P-state batch per SM: half-size batches put twice as many SMs to work).
// L: warps per state group.  L = 2 (KX_L, mechanisms whose X_k leave room for only 128 states per SM, i.e. one
// one-state warp per scheduler): the state group of warp w is shared with warp w + 4 -- same TMEM lane quadrant, same
// lanes, hence the same tensor-memory columns and the same X_k in shared memory.  The two warps split every loop over
// species (rows of the Wilke passes by parity, columns of a diffusion tile by halves) and add their partial sums
// through spare tensor-memory columns around a CTA barrier: two warps per scheduler for the same on-chip footprint,
// no coefficient is fetched twice.
template <typename ST, int P, int L>
__global__ void __launch_bounds__(KX_BK2_BLOCK, 1)
kx_bk2(const long long n_states, const long long offsetT, const long long offset, const real pressure,
       const ST* __restrict__ state, ST* __restrict__ conductivity, ST* __restrict__ viscosity,
       ST* __restrict__ rhoD, const double Tref)
{
  extern __shared__ __align__(16) unsigned char kx_sm_raw[];
  // PERSISTENT CTA of TT threads = L halves of TS threads; a thread carries P states of the current batch (slots
  // ts, ts + TS, ...): LDT = TS * P states per batch.
  constexpr int TT = KX_BK2_BLOCK, TS = TT / L, LDT = TS * P, TB = KX_TB, NB = KX_NB;
  constexpr int NWT = TT / 32, NWS = NWT / L, STG = KX_STAGES, R = KX_WR, RU = KX_WR + 6, RV = KX_WR + 12;
  constexpr int N_CHUNKS = KX_N_CHUNKS;
  constexpr int C_U = KX_NVC, C_D = C_U + KX_NUC;   // first chunk of U, of the tiles
  constexpr int G = L == 1 ? KX_COL_UNROLL : TB / L;                    // columns per group of the pair loop
  constexpr int NX = L == 1 ? 0 : (3 * R + 2 > TB ? 3 * R + 2 : TB);    // doubles a half hands to the other at a time
  constexpr int TM_SLOTS = (NWS + 3) / 4;                               // state groups per TMEM lane quadrant
  // columns per state: S_k, Mbar and sqrt(T), then one exchange area per half
  constexpr int TM_STATE = 2 * (KX_NS + 2 + L * NX);
  constexpr int TM_COLS = TM_SLOTS * P * TM_STATE;                      // columns in use per TMEM lane
  static_assert((STG & (STG - 1)) == 0 && TT % 32 == 0 && TM_COLS <= 512 && KX_NS >= KX_NP - KX_TB, "shape");
  static_assert(L == 1 || (L == 2 && NWS % 4 == 0 && TB % 2 == 0), "two warps per state group: 128-thread halves, even tile edge");
  static_assert(TB % G == 0, "the column group must divide the tile edge");
  static_assert(C_D + KX_N_DTILES == N_CHUNKS, "chunk table");
  static_assert(sizeof(ST) == 8, "the tensor-memory BK2 kernel serves FP64 buffers");
  const int tt = threadIdx.x, warp = threadIdx.x >> 5;
  const int ts = L == 1 ? tt : tt % TS, half = L == 1 ? 0 : tt / TS;
  uint64_t* const full = reinterpret_cast<uint64_t*>(kx_sm_raw);          // STG full + STG empty barriers
  uint64_t* const empty = full + STG;
  unsigned* const tm_base_slot = reinterpret_cast<unsigned*>(empty + STG);
  real* const buf0 = reinterpret_cast<real*>(kx_sm_raw + 16 * STG + 16);  // STG stages of KX_CHUNK_MAX reals
  // X[k] of state p at X[k * LDT + p * TS]; only the KX_N real species have a row
  real* __restrict__ X = buf0 + STG * KX_CHUNK_MAX + ts;
  const unsigned x_smem = kx_smem_addr(X);
  auto x_row = [&](int k) { return (k < KX_N ? k : KX_N - 1) * LDT; };

  // batches of this CTA: blockIdx.x, + gridDim.x, ...
  // (32-bit counters: the pair loop runs at the 255-register limit, and the 64-bit batch / chunk counters were what
  // ptxas spilled -- every reload on the critical path of a block's final pass, 10 % of the stall samples)
  const int n_batches = (int)((n_states + LDT - 1) / LDT);
  const int b_first = (int)blockIdx.x, b_step = (int)gridDim.x;
  const int my_batches = b_first < n_batches ? (n_batches - b_first + b_step - 1) / b_step : 0;
  const unsigned total_chunks = (unsigned)my_batches * N_CHUNKS;

  if (warp == 0) {
    // all 512 columns: this CTA is alone on its SM (shared memory), nobody else needs tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(kx_smem_addr(tm_base_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STG; s++) {
      kx_mbar_init(&full[s], 1);
      kx_mbar_init(&empty[s], NWT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();   // mbarrier inits + TMEM base address visible
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // this thread's S_k of state p: lane quadrant (warp % 4), column TM_STATE (group slot x P + p) + 2 k
  const unsigned tm0 = *reinterpret_cast<volatile unsigned*>(tm_base_slot) + ((unsigned)(warp & 3) << 21) +
                       (unsigned)(((warp % NWS) >> 2) * P * TM_STATE);
#define KX_TM(p, k) (tm0 + (unsigned)((p) * TM_STATE + 2 * (k)))
  // the halves of a state group meet here: tensor-memory stores of one visible to the loads of the other
  auto meet = [&]() {
    if constexpr (L > 1) {
      kx_tm_wait_st();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      // (a named barrier of the two partner warps only -- everything the halves exchange belongs to their own states --
      // measured 119.5 vs 118.9 M states/s on EtOHKonnov: not worth a second kind of barrier in the kernel)
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  };
  // v[0..N) += the other half's v: through the exchange areas behind the sums (two barriers: publish, then release)
  auto add_other_half = [&](auto& v, int p) {
    if constexpr (L > 1) {
      constexpr int N = sizeof(v) / sizeof(real);
      static_assert(N <= NX, "exchange area");
      real (&vv)[N] = reinterpret_cast<real (&)[N]>(v);
      kx_tm_store<N>(KX_TM(p, KX_NS + 2 + half * NX), vv);
      meet();
      unsigned raw[2 * N];
      real o[N];
      kx_tm_load<N>(KX_TM(p, KX_NS + 2 + (1 - half) * NX), raw);
      kx_tm_wait_ld();
      kx_tm_unpack<N>(raw, o);
#pragma unroll
      for (int i = 0; i < N; i++) vv[i] += o[i];
    }
  };

  if (tt == 0) {
#pragma unroll
    for (int s = 0; s < STG; s++)
      if (s < total_chunks) kx_bulk_load(buf0 + s * KX_CHUNK_MAX, kx_chunk_src(s % N_CHUNKS), kx_chunk_bytes(s % N_CHUNKS), &full[s]);
  }

  unsigned g = 0;   // chunks consumed so far by this CTA (all batches)
  auto acquire = [&]() -> const real* {
    kx_mbar_wait(&full[g & (STG - 1)], (unsigned)(g / STG) & 1);
    return buf0 + (g & (STG - 1)) * KX_CHUNK_MAX;
  };
  // hand the stage back; the first thread refills it with the chunk STG ahead (possibly of the next batch)
  auto release = [&]() {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) kx_mbar_arrive(&empty[g & (STG - 1)]);
    if (tt == 0 && g + STG < total_chunks) {
      const int s2 = (int)(g & (STG - 1)), c2 = (int)((g + STG) % N_CHUNKS);
      kx_mbar_wait(&empty[s2], (unsigned)(g / STG) & 1);
      kx_bulk_load(buf0 + s2 * KX_CHUNK_MAX, kx_chunk_src(c2), kx_chunk_bytes(c2), &full[s2]);
    }
    g++;
  };
  // state index of slot p of a batch (tail slots recompute the last state and store nothing)
  auto state_id = [&](int batch, int p) -> long long {
    const long long gid = (long long)batch * LDT + (p * TS + ts);
    return gid < n_states ? gid : n_states - 1;
  };
  // fetch the mass fractions of species [k0, k1) of `batch` (this half: those of its parity) into the state's X
  // slots (raw Y_k; the same half turns them into Y_k / M_k when that batch starts).
  auto prefetch_rows = [&](int batch, int k0, int k1) {
#pragma unroll
    for (int p = 0; p < P; p++) {
      const ST* src = state + state_id(batch, p) + offsetT;
      for (int k = k0 + ((k0 + half) & (L - 1)); k < k1; k += L)
        kx_cp_async8_nc(x_smem + (unsigned)((k * LDT + p * TS) * sizeof(real)), src + (size_t)k * offset);
    }
    kx_cp_async_commit();
  };

#ifndef KX_BK2_NO_PREFETCH
  if (b_first < n_batches) prefetch_rows(b_first, 0, KX_N);
#endif

#pragma unroll 1
  for (int batch = b_first; batch < n_batches; batch += b_step) {
    const bool has_next = batch + b_step < n_batches;
    real lnT[P], lnT2[P], lnT4[P], Mbar[P];
    // (state indices are recomputed where they are needed instead of being held in registers across the pair loop)
    auto is_live = [&](int p) { return (long long)batch * LDT + (p * TS + ts) < n_states; };

    // ---- mole fractions (transportProps.okl:23-35): the rows were fetched into X while the previous batch was in
    //      its pair loop; one pass over shared memory turns Y_k into Y_k / M_k and sums ----
    {
      real acc[P];
      ST t_raw[P];
#pragma unroll
      for (int p = 0; p < P; p++) {
        acc[p] = 0;
        t_raw[p] = kx_ld_stream(state + state_id(batch, p));   // L2 hit: prefetched during the previous batch
      }
#ifdef KX_BK2_NO_PREFETCH
      // development switch: the round-1 prologue (batches of 32 independent row loads straight from global memory)
      static_assert(L == 1, "development switch");
      constexpr int LB = 32;
#pragma unroll
      for (int k0 = 0; k0 < KX_N; k0 += LB) {
        ST y[P][LB];
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
          for (int i = 0; i < LB; i++)
            if (k0 + i < KX_N) y[p][i] = kx_ld_stream(state + state_id(batch, p) + offsetT + (size_t)(k0 + i) * offset);
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
          for (int i = 0; i < LB; i++)
            if (k0 + i < KX_N) X[(k0 + i) * LDT + p * TS] = (real)y[p][i];
      }
#else
      kx_cp_async_wait_all();
#endif
#pragma unroll 8
      for (int k = half; k < KX_N; k += L) {
#pragma unroll
        for (int p = 0; p < P; p++) {
          const real yi = X[k * LDT + p * TS];
          const real w = (yi > (real)0 ? yi : (real)0) * kx_rcpM[k];
          X[k * LDT + p * TS] = w;
          acc[p] += w;
        }
      }
#pragma unroll
      for (int p = 0; p < P; p++) {
        real a1[1] = {acc[p]};
        add_other_half(a1, p);
        Mbar[p] = kx_rcp(a1[0]);
        const double Td = Tref * (double)t_raw[p];
        lnT[p] = (real)kx_log(Td);
        lnT2[p] = lnT[p] * lnT[p];
        lnT4[p] = lnT2[p] * lnT2[p];
        // Mbar (after the first Wilke pass) and sqrt(T) are only needed where results are stored: parked in this
        // thread's tensor-memory columns so that the loops have their registers (they run at the 255-register limit)
        const real park[2] = {Mbar[p], kx_sqrt((real)Td)};
        if (half == 0) kx_tm_store<2>(KX_TM(p, KX_NS), park);
      }
      meet();                       // the exchange areas are free again; the parked scalars are visible
    }
    // Mbar and sqrt(T) of state p, back from tensor memory
    auto parked = [&](int p, real (&v)[2]) {
      unsigned praw[4];
      kx_tm_wait_st();
      kx_tm_load<2>(KX_TM(p, KX_NS), praw);
      kx_tm_wait_ld();
      kx_tm_unpack<2>(praw, v);
    };

    // ---- viscosity: Wilke, three matrix-vector products with a LOW-RANK mass-factor matrix ----
    //      (C1 + C2 v_k/v_j)^2 = c_kj (1 + w_k b_j)^2,  Phi_k = sum_j c_kj X_j (1 + 2 w_k b_j + w_k^2 b_j^2)
    //      c_kj = (8 (1 + M_k/M_j))^-1/2 is a smooth kernel in ln M_k - ln M_j: its numerical rank is KX_WR
    //      (12 for GRI-3.0, 14 for the 129-species EtOHKonnov), c = U V^T from an SVD at generation time.
    //      So  t_m = V^T (X b^m),  Phi_k = U_k . (t_0 + 2 w_k t_1 + w_k^2 t_2):  6 N r instead of 3 N^2 DFMA.
    {
      real t0[P][R], t1[P][R], t2[P][R], s1[P], s2[P];
#pragma unroll
      for (int p = 0; p < P; p++) {
        s1[p] = s2[p] = 0;
#pragma unroll
        for (int q = 0; q < R; q++) t0[p][q] = t1[p][q] = t2[p][q] = 0;
      }
      // first pass over the species (rows of [conductivity quartic, viscosity quartic, M^-1/4, -, V_j]): mole
      // fraction X_j, the conductivity sums, b_j = 1/w_j and the three projections t_m -- b_j is used on the spot
#pragma unroll 1
      for (int c = 0; c < KX_NVC; c++) {
        const real* __restrict__ cv = acquire();
        const int jb1 = min(NB, (c + 1) * (KX_VROWS / TB));
#pragma unroll 1
        for (int jb = c * (KX_VROWS / TB); jb < jb1; jb++, cv += TB * RV) {
          // one species row per iteration of a ROLLED loop (~3 KB of code, resident in the L0 instruction cache),
          // software-pipelined: the serial part of row j + L (two quartics, two Newton reciprocals: dependent chains
          // with 5-8 cycle stalls between their instructions) is issued among the 6 R independent accumulations of
          // row j.  KX_BK2_NO_SWP restores the unpipelined loop.
          auto head = [&](int jj, real (&x)[P], real (&xb)[P], real (&xbb)[P]) {
            const int j = jb * TB + jj;
            const real* row = cv + jj * RV;
            const real m4 = row[10];
#pragma unroll
            for (int p = 0; p < P; p++) {
#ifdef KX_BK2_NO_SWP
              x[p] = j < KX_N ? X[x_row(j) + p * TS] * Mbar[p] : (real)0;
              if (j < KX_N) X[x_row(j) + p * TS] = x[p];
#else
              // branch-free (a branch here would end the basic block and with it the interleaving): the load takes the
              // clamped row, the store is predicated
              const real xr = X[x_row(j) + p * TS];
              x[p] = j < KX_N ? xr * Mbar[p] : (real)0;
              kx_sts_if(x_smem + (unsigned)((x_row(j) + p * TS) * sizeof(real)), x[p], j < KX_N);
#endif
              const real lam = kx_quartic(row, lnT[p], lnT2[p], lnT4[p]);
              s1[p] = fma(x[p], lam, s1[p]);
              s2[p] = fma(x[p], kx_rcp(lam), s2[p]);
              const real b = kx_rcp(kx_quartic(row + 5, lnT[p], lnT2[p], lnT4[p]) * m4);
              xb[p] = x[p] * b;
              xbb[p] = xb[p] * b;
            }
          };
          auto project = [&](int jj, const real (&x)[P], const real (&xb)[P], const real (&xbb)[P]) {
            const real* row = cv + jj * RV;
#pragma unroll
            for (int q = 0; q < R; q += 2) {
              const real2 vv = *reinterpret_cast<const real2*>(row + 12 + q);
#pragma unroll
              for (int p = 0; p < P; p++) {
                t0[p][q] = fma(vv.x, x[p], t0[p][q]);
                t1[p][q] = fma(vv.x, xb[p], t1[p][q]);
                t2[p][q] = fma(vv.x, xbb[p], t2[p][q]);
                t0[p][q + 1] = fma(vv.y, x[p], t0[p][q + 1]);
                t1[p][q + 1] = fma(vv.y, xb[p], t1[p][q + 1]);
                t2[p][q + 1] = fma(vv.y, xbb[p], t2[p][q + 1]);
              }
            }
          };
          const int jj0 = (jb * TB + half) & (L - 1);          // this half: the species of its parity
#ifdef KX_BK2_NO_SWP
#pragma unroll 1
          for (int jj = jj0; jj < TB; jj += L) {
            real x[P], xb[P], xbb[P];
            head(jj, x, xb, xbb);
            project(jj, x, xb, xbb);
          }
#else
          real x[P], xb[P], xbb[P];
          head(jj0, x, xb, xbb);
#pragma unroll 1
          for (int jj = jj0; jj + L < TB; jj += L) {
            real xn[P], xbn[P], xbbn[P];
            head(jj + L, xn, xbn, xbbn);
            project(jj, x, xb, xbb);
#pragma unroll
            for (int p = 0; p < P; p++) { x[p] = xn[p]; xb[p] = xbn[p]; xbb[p] = xbbn[p]; }
          }
          project(jj0 + (TB - 1 - jj0) / L * L, x, xb, xbb);
#endif
        }
        release();
      }
      if constexpr (L > 1) {
        // both halves need the complete projections for their share of the second pass
#pragma unroll
        for (int p = 0; p < P; p++) {
          real pack[3 * R + 2];
#pragma unroll
          for (int q = 0; q < R; q++) { pack[q] = t0[p][q]; pack[R + q] = t1[p][q]; pack[2 * R + q] = t2[p][q]; }
          pack[3 * R] = s1[p];
          pack[3 * R + 1] = s2[p];
          add_other_half(pack, p);
#pragma unroll
          for (int q = 0; q < R; q++) { t0[p][q] = pack[q]; t1[p][q] = pack[R + q]; t2[p][q] = pack[2 * R + q]; }
          s1[p] = pack[3 * R];
          s2[p] = pack[3 * R + 1];
        }
        meet();
      }
#pragma unroll
      for (int p = 0; p < P; p++) {
        real pk[2];
        parked(p, pk);
        if (half == 0 && is_live(p))
          kx_st_stream(conductivity + state_id(batch, p), (ST)(pk[1] * ((real)0.5 * (s1[p] + kx_rcp(s2[p])))));
      }
#pragma unroll
      for (int p = 0; p < P; p++)
#pragma unroll
        for (int q = 0; q < R; q++) t1[p][q] += t1[p][q];
      real vis[P], num_prev[P], phi_prev[P];       // (num_prev = 0, phi_prev = 1: the pipelined term of "no species")
#pragma unroll
      for (int p = 0; p < P; p++) { vis[p] = 0; num_prev[p] = 0; phi_prev[p] = 1; }
#pragma unroll 1
      for (int c = 0; c < KX_NUC; c++) {
        const real* __restrict__ cu = acquire();
        const int k1 = min(KX_N, (c + 1) * KX_UROWS);
        const int kf = c * KX_UROWS + ((c * KX_UROWS + half) & (L - 1));
        cu += (kf - c * KX_UROWS) * RU;
#pragma unroll 2
        for (int k = kf; k < k1; k += L, cu += L * RU) {
          const real m4 = cu[R + 5];
#ifdef KX_BK2_NO_SWP
          real v[P], w[P], w2[P], ph[P][4];
#pragma unroll
          for (int p = 0; p < P; p++) {
            v[p] = kx_quartic(cu + R, lnT[p], lnT2[p], lnT4[p]);
            w[p] = v[p] * m4;
            w2[p] = w[p] * w[p];
            ph[p][0] = ph[p][1] = ph[p][2] = ph[p][3] = 0;
          }
#pragma unroll
          for (int q = 0; q < R; q += 2) {
            const real2 uu = *reinterpret_cast<const real2*>(cu + q);
#pragma unroll
            for (int p = 0; p < P; p++) {
              ph[p][q & 2] = fma(uu.x, fma(w2[p], t2[p][q], fma(w[p], t1[p][q], t0[p][q])), ph[p][q & 2]);
              ph[p][(q & 2) + 1] = fma(uu.y, fma(w2[p], t2[p][q + 1], fma(w[p], t1[p][q + 1], t0[p][q + 1])), ph[p][(q & 2) + 1]);
            }
          }
#pragma unroll
          for (int p = 0; p < P; p++) {
            const real phi = (ph[p][0] + ph[p][1]) + (ph[p][2] + ph[p][3]);
            vis[p] = fma(X[k * LDT + p * TS] * (v[p] * v[p]), kx_rcp(phi), vis[p]);
          }
#else
          // Phi_k = U_k.t_0 + w_k (U_k.2t_1) + w_k^2 (U_k.t_2): three dot products per state whose chains depend
          // neither on each other nor on the quartic (the nested form fma(u, fma(w2, t2, fma(w, t1, t0)), ph) left ptxas
          // 3-deep dependent triples that it issued 8 cycles apart: 671 stall cycles for 408 pipe cycles per two
          // species); the previous species' Newton reciprocal is finished among this one's FMAs
          real a0[P], a1[P], a2[P];
#pragma unroll
          for (int p = 0; p < P; p++) {
            a0[p] = a1[p] = a2[p] = 0;
            vis[p] = fma(num_prev[p], kx_rcp(phi_prev[p]), vis[p]);
          }
#pragma unroll
          for (int q = 0; q < R; q += 2) {
            const real2 uu = *reinterpret_cast<const real2*>(cu + q);
#pragma unroll
            for (int p = 0; p < P; p++) {
              a0[p] = fma(uu.x, t0[p][q], a0[p]);
              a1[p] = fma(uu.x, t1[p][q], a1[p]);
              a2[p] = fma(uu.x, t2[p][q], a2[p]);
              a0[p] = fma(uu.y, t0[p][q + 1], a0[p]);
              a1[p] = fma(uu.y, t1[p][q + 1], a1[p]);
              a2[p] = fma(uu.y, t2[p][q + 1], a2[p]);
            }
          }
#pragma unroll
          for (int p = 0; p < P; p++) {
            const real v = kx_quartic(cu + R, lnT[p], lnT2[p], lnT4[p]);
            const real w = v * m4;
            phi_prev[p] = fma(w, fma(w, a2[p], a1[p]), a0[p]);
            num_prev[p] = X[k * LDT + p * TS] * (v * v);
          }
#endif
        }
        release();
      }
#pragma unroll
      for (int p = 0; p < P; p++) {
#ifndef KX_BK2_NO_SWP
        vis[p] = fma(num_prev[p], kx_rcp(phi_prev[p]), vis[p]);
#endif
        real v1[1] = {vis[p]};
        add_other_half(v1, p);
        real pk[2];
        parked(p, pk);
        if (half == 0 && is_live(p)) kx_st_stream(viscosity + state_id(batch, p), (ST)(pk[1] * v1[0]));
      }
      meet();       // every X_j is final (written by the half of its parity) before anybody's pair loop reads it
    }

    // ---- mixture-averaged diffusion: S_k = sum_{j != k} X_j / D_kj over tiles of the lower triangle, row blocks
    //      in DESCENDING order; the coefficients of a pair are fetched once for the P states.  When the row pass of
    //      block kb ends its S_k are complete (the blocks above have already added their column contributions in
    //      tensor memory): rho * D_km of the block is formed and stored at once (mix_transport.py:621-622 and
    //      transportProps.okl:43-47; p and Mbar cancel), and the block's X rows are refilled with the next batch ----
    if (has_next && half == 0) {
#pragma unroll
      for (int p = 0; p < P; p++) asm volatile("prefetch.global.L2 [%0];" ::"l"(state + state_id(batch + b_step, p)));
    }
    if (half == 0) {   // the column sums of blocks 0 .. NB-2 start from zero
      const real zero[TB] = {};
#pragma unroll 1
      for (int jb = 0; jb < NB - 1; jb++)
#pragma unroll
        for (int p = 0; p < P; p++) kx_tm_store<TB>(KX_TM(p, jb * TB), zero);
    }
    meet();
#pragma unroll 1
    for (int kb = NB - 1; kb >= 0; kb--) {
      const bool top = kb == NB - 1;          // the top block never goes to tensor memory: its sums start here
      real xk[P][TB], sk[P][TB];
      {
        unsigned raw[P][2 * TB];
        const bool from_tm = !top && half == 0;      // the column contributions collected so far (one half adds them)
        if (from_tm) {
          kx_tm_wait_st();
#pragma unroll
          for (int p = 0; p < P; p++) kx_tm_load<TB>(KX_TM(p, kb * TB), raw[p]);
        }
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
          for (int i = 0; i < TB; i++) {
            const int k = kb * TB + i;
            xk[p][i] = k < KX_N ? X[x_row(k) + p * TS] : (real)0;
          }
        if (from_tm) {
          kx_tm_wait_ld();
#pragma unroll
          for (int p = 0; p < P; p++) kx_tm_unpack<TB>(raw[p], sk[p]);
        } else {
#pragma unroll
          for (int p = 0; p < P; p++)
#pragma unroll
            for (int i = 0; i < TB; i++) sk[p][i] = 0;
        }
      }
#pragma unroll 1
      for (int jb = 0; jb < kb; jb++) {
        // G columns of the tile per iteration of a ROLLED loop (L = 1: KX_COL_UNROLL columns; L = 2: each half takes
        // its half of the tile in one go).  A column's mole fraction and running sum are scalars per state (shared /
        // tensor memory; the sums of a group travel together, fetched one group ahead), only the row block's xk / sk
        // stay in registers.  One column = TB pairs x P states = ~3 KB of code, resident in the 6 KB L0 instruction
        // cache.  (Unrolled over the whole 9 x 9 tile the body is ~30 KB -- at the edge of the 32 KB L1.5 instruction
        // cache: the round-1 body just fitted (hit rate 95 %), a few more instructions per pair and it thrashes: hit
        // rate 77 %, no_instruction stalls x6, 434 instead of 644 M states/s.)
        unsigned cur[P][2 * G], nxt[P][2 * G];
        kx_tm_wait_st();
#pragma unroll
        for (int p = 0; p < P; p++) kx_tm_load<G>(KX_TM(p, jb * TB + half * G), cur[p]);
        const real* __restrict__ tile = acquire();
        // (L = 2: exactly one group per half -- a trip count the compiler can see, no loop test)
#pragma unroll 1
        for (int j0 = half * G, once = 0; L == 1 ? j0 < TB : once < 1; j0 += L * G, once++) {
          real sj[P][G];
          kx_tm_wait_ld();
#pragma unroll
          for (int p = 0; p < P; p++) kx_tm_unpack<G>(cur[p], sj[p]);
          if constexpr (L == 1) {
            // the next group's sums (the last iteration fetches its own group again: harmless, unused)
#pragma unroll
            for (int p = 0; p < P; p++) kx_tm_load<G>(KX_TM(p, jb * TB + min(j0 + G, TB - G)), nxt[p]);
          }
#pragma unroll
          for (int jj = 0; jj < G; jj++) {
            const int j = j0 + jj;
            real xj[P];
#pragma unroll
            for (int p = 0; p < P; p++) xj[p] = X[(jb * TB + j) * LDT + p * TS];   // jb < kb: real species
            real d[P][TB];
            const real* __restrict__ col = tile + j * KX_COL;
#pragma unroll
            for (int i = 0; i + 1 < TB; i += 2) {      // two rows per record: five 16-byte loads
              const real2* cp = reinterpret_cast<const real2*>(col + (i / 2) * 10);
              const real2 c0 = cp[0], c1 = cp[1], c2 = cp[2], c3 = cp[3], c4 = cp[4];
#pragma unroll
              for (int p = 0; p < P; p++) {
                const real qa = fma(c4.x, lnT4[p], fma(fma(c3.x, lnT[p], c2.x), lnT2[p], fma(c1.x, lnT[p], c0.x)));
                const real qb = fma(c4.y, lnT4[p], fma(fma(c3.y, lnT[p], c2.y), lnT2[p], fma(c1.y, lnT[p], c0.y)));
                d[p][i] = KX_RCP_DIFF ? qa : KX_PAIR_RCP(qa);
                d[p][i + 1] = KX_RCP_DIFF ? qb : KX_PAIR_RCP(qb);
              }
            }
            if (TB & 1) {
              const real* cp = col + (TB / 2) * 10;
              const real2 c01 = *reinterpret_cast<const real2*>(cp), c23 = *reinterpret_cast<const real2*>(cp + 2);
              const real c4 = cp[4];
#pragma unroll
              for (int p = 0; p < P; p++) {
                const real q = fma(c4, lnT4[p], fma(fma(c23.y, lnT[p], c23.x), lnT2[p], fma(c01.y, lnT[p], c01.x)));
                d[p][TB - 1] = KX_RCP_DIFF ? q : KX_PAIR_RCP(q);
              }
            }
#pragma unroll
            for (int p = 0; p < P; p++) {
              real se = 0, so = 0;
#pragma unroll
              for (int i = 0; i < TB; i++) {
                if (i & 1) so = fma(xk[p][i], d[p][i], so); else se = fma(xk[p][i], d[p][i], se);
                sk[p][i] = fma(xj[p], d[p][i], sk[p][i]);
              }
              sj[p][jj] += se + so;
            }
          }
#pragma unroll
          for (int p = 0; p < P; p++) {
            kx_tm_store<G>(KX_TM(p, jb * TB + j0), sj[p]);
            if constexpr (L == 1) {
#pragma unroll
              for (int q = 0; q < 2 * G; q++) cur[p][q] = nxt[p][q];
            }
          }
        }
        if constexpr (L == 1) kx_tm_wait_ld();       // the look-ahead load of the last group
        release();
      }
      // diagonal tile: pairs i > j inside the block.  The halves (L = 2) take alternate pairs; WHICH half this warp is
      // is decided once, outside the unrolled pair list (a test per pair is a branch per ten instructions: the phase
      // ran at 21 % pipe utilisation and took 10 % of EtOHKonnov's time for 4 % of its work)
      {
        const real* __restrict__ tile = acquire();
        auto diagonal = [&](auto which) {
          constexpr int H = decltype(which)::value;
#pragma unroll
          for (int i = 1; i < TB; i++) {
#pragma unroll
            for (int j = 0; j < i; j++) {
              if (L > 1 && ((i + j) & (L - 1)) != H) continue;
              // coefficient m of pair (i, j): paired rows interleaved, the odd last row on its own (see KX_COL)
              const real* cp = tile + j * KX_COL + (i / 2) * 10 + ((i | 1) < TB ? (i & 1) : 0);
              const int cs = (i | 1) < TB ? 2 : 1;
              const real c0 = cp[0], c1 = cp[cs], c2 = cp[2 * cs], c3 = cp[3 * cs], c4 = cp[4 * cs];
#pragma unroll
              for (int p = 0; p < P; p++) {
                const real q = fma(c4, lnT4[p], fma(fma(c3, lnT[p], c2), lnT2[p], fma(c1, lnT[p], c0)));
                const real d = KX_RCP_DIFF ? q : KX_PAIR_RCP(q);
                sk[p][i] = fma(xk[p][j], d, sk[p][i]);
                sk[p][j] = fma(xk[p][i], d, sk[p][j]);
              }
            }
          }
        };
        if (L == 1 || half == 0) diagonal(kx_int<0>{}); else diagonal(kx_int<1>{});
        release();
      }
      // S_k of this block are final (L = 2: after the halves have added their partial sums):
      // rho * D_km = sqrt(T) / R * (Mbar - M_k X_k) / S_k, stored while the pipe goes on
      if constexpr (L > 1) {
#pragma unroll
        for (int p = 0; p < P; p++) add_other_half(sk[p], p);
      }
      unsigned praw[P][4];
      kx_tm_wait_st();
#pragma unroll
      for (int p = 0; p < P; p++) kx_tm_load<2>(KX_TM(p, KX_NS), praw[p]);
      kx_tm_wait_ld();
      auto store_rows = [&](auto which) {          // each half stores the rows of its half of the block
        constexpr int H = decltype(which)::value;
#pragma unroll
        for (int p = 0; p < P; p++) {
          real park[2];
          kx_tm_unpack<2>(praw[p], park);
          const real Mb = park[0], f = park[1] * (real)(1.0 / 8.31446261815324);   // rho*T^1.5/(p*Mbar) = sqrt(T)/R
          const bool lv = is_live(p);
          ST* const dst = rhoD + state_id(batch, p) + (size_t)(kb * TB) * offset;
#pragma unroll
          for (int i = 0; i < TB; i++) {
            if (L > 1 && (i / (TB / L)) != H) continue;
            const int k = kb * TB + i;
            if (k < KX_N) {
              const real num = fma(-kx_M[k], xk[p][i], Mb);
              const real v = f * num * kx_rcp(sk[p][i]);
              if (lv) kx_st_stream(dst + (size_t)i * offset, (ST)v);
            }
          }
        }
      };
      if (L == 1 || half == 0) store_rows(kx_int<0>{}); else store_rows(kx_int<1>{});
      meet();       // both halves are done with the block's X rows and with the exchange areas
      // the block's X rows are dead: the next batch's mass fractions move in
#ifndef KX_BK2_NO_PREFETCH
      if (has_next) prefetch_rows(batch + b_step, kb * TB, min(KX_N, kb * TB + TB));
#endif
    }
  }
  kx_cp_async_wait_all();

  // release tensor memory
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*reinterpret_cast<volatile unsigned*>(tm_base_slot))
                 : "memory");
  }
  (void)pressure;
#undef KX_TM
}
