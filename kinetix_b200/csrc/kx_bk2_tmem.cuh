// kx_bk2_tmem.cuh -- BK2 (mixture-averaged conductivity, viscosity, rho*D_km): KX_P states per thread, the
// per-state sums S_k in TENSOR MEMORY, the mole fractions X_k in shared memory.  FP64 only.
//
// Same arithmetic as csrc/kx_bk2.cuh (reference benchmark/okl/transportProps.okl:11-49 around
// kinetix/core/mix_transport.py:474-626).  What changed, and why (profiles/ncu_r01_bk2_*.txt):
//   * The one-state-per-thread kernel is bound by the shared-memory data pipe, not by the FP64 pipe: every
//     species pair costs 3 LDS.128 (the quartic's coefficients, a broadcast) = 6 LSU wavefronts per warp against
//     9 DFMA; in the pair loops the LSU pipe is the busier one (65 % of peak overall, FP64 58 %).  Replacing the
//     coefficient loads by register moves (timing experiment) gave 739 instead of 446 M states/s.
//   * So every coefficient fetched from shared memory is used for KX_P = 2 states here: half the wavefronts per
//     state.  A thread then needs both states' vectors on chip, and shared memory (227 KB) only holds X_k AND
//     S_k for 256 GRI-3.0 states per SM, i.e. 4 warps of two-state threads -- too few to cover latencies.
//   * Blackwell's tensor memory is 256 KB per SM that this kernel would leave idle.  It is addressed as
//     128 lanes x 512 32-bit columns, and warp w may touch lanes 32 (w % 4) .. +31: exactly a per-thread
//     scratchpad.  The S_k of all 512 resident states live there (thread = lane, 2 * KX_NS columns per state,
//     tcgen05.st / tcgen05.ld 32x32b, SASS STTM / LDTM), the X_k stay in shared memory as [k][state]: 8 warps
//     of two-state threads per SM, and the S_k traffic (b_j in the Wilke sums, the read-modify-write of the
//     column sums per tile) moves off the LSU pipe onto the otherwise unused TMEM datapath.
//   * The Wilke sum uses a rank-KX_WR factorisation of its mass-factor matrix (see the Wilke section below):
//     6 N r instead of 3 N^2 multiply-adds.
//   * All tables (species quartics, Wilke factors, diffusion tiles) are ONE stream of chunks that goes through a
//     KX_STAGES-deep ring of TMA bulk copies with full/empty mbarriers (no CTA-wide barrier per chunk, no
//     constant-cache loads in the loops).
//   * Persistent CTAs (one per SM) loop over batches of KX_BK2_BLOCK * KX_P states; tensor-memory allocation,
//     barrier set-up and the ring's prefetch carry across batches.  KX_TEAMS = 2 splits the CTA into two
//     independent half-CTA teams with skewed phases (measured slower: instruction-cache misses; off).
//   * State rows are loaded in fully unrolled batches of 32 independent loads for all KX_P states.
//
// The including translation unit defines KX_N, KX_NP (multiple of KX_TB), KX_TB, KX_P, KX_TEAMS, KX_BK2_BLOCK
// (threads; KX_BK2_BLOCK / KX_TEAMS a multiple of 128 so that a team's warps cover the four TMEM lane quadrants),
// KX_NS (doubles reserved per state in TMEM, >= KX_NP), KX_STAGES, KX_CHUNK_MAX (reals per stage), KX_WR (even:
// rank of the Wilke factorisation), the chunk layout KX_N_CHUNKS / KX_NVC / KX_NUC / KX_VROWS / KX_UROWS (below), KX_RCP_DIFF and the tables
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
//   the lower-triangular diffusion tiles (KX_TB^2 pairs x 5 coefficients), row-major over (kb, jb <= kb)
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

template <typename ST>   // ST: storage type of the state / result buffers (reference: dfloat)
__global__ void __launch_bounds__(KX_BK2_BLOCK, 1)
kx_bk2(const long long n_states, const long long offsetT, const long long offset, const real pressure,
       const ST* __restrict__ state, ST* __restrict__ conductivity, ST* __restrict__ viscosity,
       ST* __restrict__ rhoD, const double Tref)
{
  extern __shared__ __align__(16) unsigned char kx_sm_raw[];
  // PERSISTENT CTA of TEAMS independent teams of TT threads; a thread carries P states of its team's current
  // batch (slots t, t + TT, ...): LDT = TT * P states per batch.  Each team has its own ring of table stages,
  // so with two teams one team's memory-bound prologue / epilogue overlaps the other's FP64 loops (the second
  // team starts half a batch late and the offset persists).
  constexpr int P = KX_P, TEAMS = KX_TEAMS, TT = KX_BK2_BLOCK / TEAMS, LDT = TT * P, TB = KX_TB;
  constexpr int NWT = TT / 32, STG = KX_STAGES, R = KX_WR, RU = KX_WR + 6, RV = KX_WR + 12;
  constexpr int N_CHUNKS = KX_N_CHUNKS;
  constexpr int C_U = KX_NVC, C_D = C_U + KX_NUC;   // first chunk of U, of the tiles
  constexpr int TM_COLS = (KX_BK2_BLOCK / 128) * P * 2 * KX_NS;         // columns in use per TMEM lane
  static_assert((STG & (STG - 1)) == 0 && TT % 128 == 0 && TM_COLS <= 512 && KX_NS >= KX_NP, "shape");
  static_assert(C_D + KX_N_DTILES == N_CHUNKS, "chunk table");
  const int team = threadIdx.x / TT, tt = threadIdx.x % TT, warp = threadIdx.x >> 5;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(kx_sm_raw);          // per team: STG full + STG empty
  uint64_t* const full = bars + team * 2 * STG;
  uint64_t* const empty = full + STG;
  uint64_t* const skew = bars + TEAMS * 2 * STG;                          // one-time start signal for team 1
  unsigned* const tm_base_slot = reinterpret_cast<unsigned*>(skew + 1);
  real* const bufs = reinterpret_cast<real*>(kx_sm_raw + 16 * STG * TEAMS + 16);
  real* const buf0 = bufs + team * STG * KX_CHUNK_MAX;                    // this team's STG stages
  // X[k] of state p at X[k * LDT + p * TT]; only the KX_N real species have a row
  real* __restrict__ X = bufs + TEAMS * STG * KX_CHUNK_MAX + team * (KX_N * LDT) + tt;
  auto x_row = [&](int k) { return (k < KX_N ? k : KX_N - 1) * LDT; };

  // batches of this team: b = blockIdx.x * TEAMS + team, + gridDim.x * TEAMS, ...
  const long long n_batches = (n_states + LDT - 1) / LDT;
  const long long b_first = (long long)blockIdx.x * TEAMS + team, b_step = (long long)gridDim.x * TEAMS;
  const long long my_batches = b_first < n_batches ? (n_batches - b_first + b_step - 1) / b_step : 0;
  const long long total_chunks = my_batches * N_CHUNKS;

  if (warp == 0) {
    // all 512 columns: this CTA is alone on its SM (shared memory), nobody else needs tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(kx_smem_addr(tm_base_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < TEAMS * STG; s++) {
      kx_mbar_init(&bars[(s / STG) * 2 * STG + s % STG], 1);
      kx_mbar_init(&bars[(s / STG) * 2 * STG + STG + s % STG], NWT);
    }
    kx_mbar_init(skew, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();   // mbarrier inits + TMEM base address visible
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // this thread's S_k of state p: lane quadrant (warp % 4), column (2 KX_NS) ((warp / 4) P + p) + 2 k
  const unsigned tm0 = *reinterpret_cast<volatile unsigned*>(tm_base_slot) + ((unsigned)(warp & 3) << 21) +
                       (unsigned)((warp >> 2) * P * 2 * KX_NS);
#define KX_TM(p, k) (tm0 + (unsigned)((p) * 2 * KX_NS + 2 * (k)))

  if (tt == 0) {
#pragma unroll
    for (int s = 0; s < STG; s++)
      if (s < total_chunks) kx_bulk_load(buf0 + s * KX_CHUNK_MAX, kx_chunk_src(s % N_CHUNKS), kx_chunk_bytes(s % N_CHUNKS), &full[s]);
    if (TEAMS > 1 && team == 0 && my_batches == 0) kx_mbar_arrive(skew);
  }
  if (TEAMS > 1 && team == 1 && my_batches > 0) kx_mbar_wait(skew, 0);

  long long g = 0;   // chunks consumed so far by this team (all batches)
  auto acquire = [&]() -> const real* {
    kx_mbar_wait(&full[g & (STG - 1)], (unsigned)(g / STG) & 1);
    return buf0 + (g & (STG - 1)) * KX_CHUNK_MAX;
  };
  // hand the stage back; the team's first thread refills it with the chunk STG ahead (possibly of the next batch)
  auto release = [&]() {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) kx_mbar_arrive(&empty[g & (STG - 1)]);
    if (tt == 0) {
      if (g + STG < total_chunks) {
        const int s2 = (int)(g & (STG - 1)), c2 = (int)((g + STG) % N_CHUNKS);
        kx_mbar_wait(&empty[s2], (unsigned)(g / STG) & 1);
        kx_bulk_load(buf0 + s2 * KX_CHUNK_MAX, kx_chunk_src(c2), kx_chunk_bytes(c2), &full[s2]);
      }
      if (TEAMS > 1 && team == 0 && g == N_CHUNKS / 2) kx_mbar_arrive(skew);   // lets team 1 start, half a batch late
    }
    g++;
  };

#pragma unroll 1
  for (long long batch = b_first; batch < n_batches; batch += b_step) {
    bool live[P];
    long long id[P];
    real lnT[P], lnT2[P], lnT4[P], sqrT[P], Mbar[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
      const long long gid = batch * LDT + p * TT + tt;
      live[p] = gid < n_states;
      id[p] = live[p] ? gid : n_states - 1;   // tail slots recompute the last state, store nothing
    }

    // ---- mole fractions (transportProps.okl:23-35): batches of 32 independent row loads for all P states ----
    {
      ST t_raw[P];
      real acc[P];
#pragma unroll
      for (int p = 0; p < P; p++) { t_raw[p] = kx_ld_stream(state + id[p]); acc[p] = 0; }
      constexpr int LB = 32;
#pragma unroll
      for (int k0 = 0; k0 < KX_N; k0 += LB) {
        ST y[P][LB];
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
          for (int i = 0; i < LB; i++)
            if (k0 + i < KX_N) y[p][i] = kx_ld_stream(state + id[p] + offsetT + (size_t)(k0 + i) * offset);
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
          for (int i = 0; i < LB; i++) {
            if (k0 + i < KX_N) {
              const real yi = (real)y[p][i];
              const real w = (yi > (real)0 ? yi : (real)0) * kx_rcpM[k0 + i];
              X[(k0 + i) * LDT + p * TT] = w;
              acc[p] += w;
            }
          }
      }
#pragma unroll
      for (int p = 0; p < P; p++) {
        Mbar[p] = kx_rcp(acc[p]);
        const double Td = Tref * (double)t_raw[p];
        lnT[p] = (real)kx_log(Td);
        sqrT[p] = kx_sqrt((real)Td);
        lnT2[p] = lnT[p] * lnT[p];
        lnT4[p] = lnT2[p] * lnT2[p];
      }
      // optional: pull the NEXT batch's state rows into L2 while this batch computes.  Measured slower (600 vs
      // 610 M states/s on GRI-3.0: 108 extra LSU instructions per thread and batch), so off by default.
#ifdef KX_L2_PREFETCH
      if (batch + b_step < n_batches) {
#pragma unroll
        for (int p = 0; p < P; p++) {
          const long long nid = min((batch + b_step) * LDT + p * TT + tt, n_states - 1);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(state + nid));
#pragma unroll 4
          for (int k = 0; k < KX_N; k++)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(state + nid + offsetT + (size_t)k * offset));
        }
      }
#endif
    }

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
        const int jb1 = min(KX_NB, (c + 1) * (KX_VROWS / TB));
#pragma unroll 1
        for (int jb = c * (KX_VROWS / TB); jb < jb1; jb++, cv += TB * RV) {
#pragma unroll
          for (int jj = 0; jj < TB; jj++) {
            const int j = jb * TB + jj;
            const real* row = cv + jj * RV;
            const real m4 = row[10];
            real x[P], xb[P], xbb[P];
#pragma unroll
            for (int p = 0; p < P; p++) {
              x[p] = j < KX_N ? X[x_row(j) + p * TT] * Mbar[p] : (real)0;
              if (j < KX_N) X[x_row(j) + p * TT] = x[p];
              const real lam = kx_quartic(row, lnT[p], lnT2[p], lnT4[p]);
              s1[p] = fma(x[p], lam, s1[p]);
              s2[p] = fma(x[p], kx_rcp(lam), s2[p]);
              const real b = kx_rcp(kx_quartic(row + 5, lnT[p], lnT2[p], lnT4[p]) * m4);
              xb[p] = x[p] * b;
              xbb[p] = xb[p] * b;
            }
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
          }
        }
        release();
      }
#pragma unroll
      for (int p = 0; p < P; p++)
        if (live[p]) kx_st_stream(conductivity + id[p], (ST)(sqrT[p] * ((real)0.5 * (s1[p] + kx_rcp(s2[p])))));
#pragma unroll
      for (int p = 0; p < P; p++)
#pragma unroll
        for (int q = 0; q < R; q++) t1[p][q] += t1[p][q];
      real vis[P];
#pragma unroll
      for (int p = 0; p < P; p++) vis[p] = 0;
#pragma unroll 1
      for (int c = 0; c < KX_NUC; c++) {
        const real* __restrict__ cu = acquire();
        const int k1 = min(KX_N, (c + 1) * KX_UROWS);
#pragma unroll 2
        for (int k = c * KX_UROWS; k < k1; k++, cu += RU) {
          const real m4 = cu[R + 5];
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
            vis[p] = fma(X[k * LDT + p * TT] * (v[p] * v[p]), kx_rcp(phi), vis[p]);
          }
        }
        release();
      }
#pragma unroll
      for (int p = 0; p < P; p++)
        if (live[p]) kx_st_stream(viscosity + id[p], (ST)(sqrT[p] * vis[p]));
    }

    // ---- mixture-averaged diffusion: S_k = sum_{j != k} X_j / D_kj over tiles of the lower triangle; the
    //      coefficients of a pair are fetched once for the P states ----
#pragma unroll 1
    for (int kb = 0; kb < KX_NB; kb++) {
      real xk[P][TB], sk[P][TB];
#pragma unroll
      for (int p = 0; p < P; p++)
#pragma unroll
        for (int i = 0; i < TB; i++) {
          const int k = kb * TB + i;
          xk[p][i] = k < KX_N ? X[x_row(k) + p * TT] : (real)0;
          sk[p][i] = 0;
        }
#pragma unroll 1
      for (int jb = 0; jb < kb; jb++) {
        real xj[P][TB], sj[P][TB];
        unsigned raw[P][2 * TB];
        // running sums of the column block: requested from tensor memory now, unpacked when the tile has landed
        kx_tm_wait_st();
#pragma unroll
        for (int p = 0; p < P; p++) kx_tm_load<TB>(KX_TM(p, jb * TB), raw[p]);
#pragma unroll
        for (int p = 0; p < P; p++)
#pragma unroll
          for (int i = 0; i < TB; i++) xj[p][i] = X[(jb * TB + i) * LDT + p * TT];   // jb < kb: real species
        const real* __restrict__ tile = acquire();
        kx_tm_wait_ld();
#pragma unroll
        for (int p = 0; p < P; p++) kx_tm_unpack<TB>(raw[p], sj[p]);
#pragma unroll
        for (int i = 0; i < TB; i++) {
          real d[P][TB];
#pragma unroll
          for (int j = 0; j < TB; j++) {
            const real* cp = tile + (i * TB + j) * 5;
            const real c0 = cp[0], c1 = cp[1], c2 = cp[2], c3 = cp[3], c4 = cp[4];
#pragma unroll
            for (int p = 0; p < P; p++) {
              const real q = fma(c4, lnT4[p], fma(fma(c3, lnT[p], c2), lnT2[p], fma(c1, lnT[p], c0)));
              d[p][j] = KX_RCP_DIFF ? q : KX_PAIR_RCP(q);
            }
          }
#pragma unroll
          for (int p = 0; p < P; p++) {
            real se = 0, so = 0;
#pragma unroll
            for (int j = 0; j < TB; j++) {
              if (j & 1) so = fma(xj[p][j], d[p][j], so); else se = fma(xj[p][j], d[p][j], se);
              sj[p][j] = fma(xk[p][i], d[p][j], sj[p][j]);
            }
            sk[p][i] += se + so;
          }
        }
        release();
#pragma unroll
        for (int p = 0; p < P; p++) kx_tm_store<TB>(KX_TM(p, jb * TB), sj[p]);
      }
      // diagonal tile: pairs i > j inside the block
      {
        const real* __restrict__ tile = acquire();
#pragma unroll
        for (int i = 1; i < TB; i++) {
#pragma unroll
          for (int j = 0; j < i; j++) {
            const real* cp = tile + (i * TB + j) * 5;
            const real c0 = cp[0], c1 = cp[1], c2 = cp[2], c3 = cp[3], c4 = cp[4];
#pragma unroll
            for (int p = 0; p < P; p++) {
              const real q = fma(c4, lnT4[p], fma(fma(c3, lnT[p], c2), lnT2[p], fma(c1, lnT[p], c0)));
              const real d = KX_RCP_DIFF ? q : KX_PAIR_RCP(q);
              sk[p][i] = fma(xk[p][j], d, sk[p][i]);
              sk[p][j] = fma(xk[p][i], d, sk[p][j]);
            }
          }
        }
        release();
      }
      // first touch of this row block's sums (they replace b_k): later row blocks add their column contributions
#pragma unroll
      for (int p = 0; p < P; p++) kx_tm_store<TB>(KX_TM(p, kb * TB), sk[p]);
    }
    kx_tm_wait_st();

    // ---- rho * D_km  (mix_transport.py:621-622 and transportProps.okl:43-47; p and Mbar cancel) ----
#pragma unroll
    for (int kb = 0; kb < KX_NB; kb++) {
      unsigned raw[P][2 * TB];
      real s[P][TB];
#pragma unroll
      for (int p = 0; p < P; p++) kx_tm_load<TB>(KX_TM(p, kb * TB), raw[p]);
      kx_tm_wait_ld();
#pragma unroll
      for (int p = 0; p < P; p++) kx_tm_unpack<TB>(raw[p], s[p]);
#pragma unroll
      for (int p = 0; p < P; p++) {
        const real f = sqrT[p] * (real)(1.0 / 8.31446261815324);      // rho*T^1.5/(p*Mbar) = sqrt(T)/R
#pragma unroll
        for (int i = 0; i < TB; i++) {
          const int k = kb * TB + i;
          if (k < KX_N) {
            const real num = fma(-kx_M[k], X[k * LDT + p * TT], Mbar[p]);
            const real v = f * num * kx_rcp(s[p][i]);
            if (live[p]) kx_st_stream(rhoD + id[p] + (size_t)k * offset, (ST)v);
          }
        }
      }
    }
  }

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
