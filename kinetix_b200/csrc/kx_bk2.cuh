// kx_bk2.cuh -- BK2: mixture-averaged conductivity, viscosity and rho*D_km, one thread per state.
//
// Computes what the reference's `transport` OKL kernel computes around the generated
// kinetix_conductivity / kinetix_viscosity / kinetix_diffusivity routines
// (reference benchmark/okl/transportProps.okl:11-49, kinetix/core/mix_transport.py:474-626).
//
// Re-designed for B200 instead of translating the reference's fully unrolled per-pair code (which, on
// sm_100a, needs 73 KB of spill traffic per state and ~19k MOVs for its FP64 immediates):
//   * per-state vectors X_k and the running sums live in shared memory as [k][thread] (conflict-free,
//     one 8-byte word per lane), NOT in per-thread local memory;
//   * the N(N-1)/2 binary-diffusion polynomials and the N^2 Wilke mass factors are TABLES in global
//     memory (L2 resident).  They are cut into chunks (one k-block of Wilke factors / one TB x TB tile
//     of diffusion quartics) that the CTA streams through a double buffer in shared memory with TMA
//     bulk copies (cp.async.bulk + mbarrier complete_tx): chunk c+2 is in flight while the warps work
//     on chunk c, and every coefficient read in the inner loops is a conflict-free shared-memory
//     broadcast instead of a global load (v1 of this kernel read the tables with __ldg and spent 2.0
//     stall cycles per issue slot on the long scoreboard);
//   * the loop nest over TB x TB species tiles keeps the accumulators of both the row block and the
//     column block in registers (each D_jk is evaluated once and used for S_j and S_k); the code is a
//     few KB and instruction-cache resident;
//   * Wilke's sum is refactored:  (C1 + C2 v_k/v_j)^2 = c_kj (1 + w_k/w_j)^2 with w = v M^(-1/4) and
//     c_kj = 1/sqrt(8 (1 + M_k/M_j)), so
//        Phi_k = sum_j c_kj X_j  + 2 w_k sum_j c_kj X_j/w_j  + w_k^2 sum_j c_kj X_j/w_j^2
//     i.e. three constant-matrix x per-state-vector products: 3 N^2 DFMA instead of 4 N^2 mixed ops.
//     (DMMA was measured to share the FP64 pipe with DFMA on B200 -- profiles/peaks_r01.json -- so the
//     products stay on the vector pipe);
//   * divisions are MUFU.RCP64H + one cubic Newton step (kx_rcp), rho*D_km is simplified algebraically
//     (p and Mbar cancel exactly as in the reference's formula, transportProps.okl:41-46).
//
// The including translation unit (generated per mechanism) defines:
//   KX_N, KX_NP (= KX_N rounded up to a multiple of KX_TB), KX_TB, KX_BK2_BLOCK, KX_BK2_MINB,
//   KX_WCHUNK (doubles per Wilke chunk, even), KX_DCHUNK (= KX_TB*KX_TB*6)
//   __constant__ double kx_rcpM[KX_N], kx_M[KX_N], kx_m4[KX_N]  (1/M_k, M_k, M_k^(-1/4))
//   __constant__ double kx_cond[KX_N][5], kx_visc[KX_N][5]       quartics in ln T
//   __device__   double kx_wilke[KX_NB][KX_WCHUNK]                c_kj as [kb][j][i], k = kb*TB + i
//   __device__   double kx_diff[n_tiles][KX_DCHUNK]               lower-triangular tiles (kb >= jb),
//                                                                 row-major over (kb, jb); 5 coefs + pad
#pragma once
#include <cstdint>
#include "kx_math.cuh"

#define KX_NB (KX_NP / KX_TB)
// KX_CHUNK_MAX (reals per streaming buffer) is defined by the including translation unit
#define KX_N_DTILES (KX_NB * (KX_NB + 1) / 2)

// `real` (double | float) is the arithmetic + table type of the module, `real2` its 2-vector.
KX_DEVICE real kx_quartic(const real* __restrict__ c, real l)
{
  return fma(fma(fma(fma(c[4], l, c[3]), l, c[2]), l, c[1]), l, c[0]);
}

// quartic in ln T of one species pair from 5 coefficients stored as 3 x double2 (shared memory, all lanes
// read the same address: broadcast).  Estrin form: dependency depth 3 instead of Horner's 4, same 4 DFMA.
KX_DEVICE real kx_pair_poly(const real2* __restrict__ c, real l, real l2, real l4)
{
  const real2 c01 = c[0], c23 = c[1], c4 = c[2];
  return fma(c4.x, l4, fma(fma(c23.y, l, c23.x), l2, fma(c01.y, l, c01.x)));
}

// 1/D_kj from the fitted quartic: a reciprocal, unless the mechanism was generated with
// --fit-rcpdiffcoeffs, in which case the fit already IS the reciprocal (reference mix_transport.py:198-206,
// 613-618) and the division disappears.
KX_DEVICE real kx_pair_rcp_d(const real2* __restrict__ c, real l, real l2, real l4)
{
#if KX_RCP_DIFF
  return kx_pair_poly(c, l, l2, l4);
#elif defined(KX_BK2_FAST_RCP)
  return kx_rcp_fast(kx_pair_poly(c, l, l2, l4));
#else
  return kx_rcp(kx_pair_poly(c, l, l2, l4));
#endif
}

// ---- TMA bulk copy + mbarrier plumbing -----------------------------------------------------------
KX_DEVICE void kx_mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
KX_DEVICE void kx_bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar)
{
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}
KX_DEVICE void kx_mbar_wait(uint64_t* bar, unsigned parity)
{
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "KX_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra KX_DONE_%=;\n\t"
      "bra KX_WAIT_%=;\n\t"
      "KX_DONE_%=:\n\t}" ::"r"(b), "r"(parity)
      : "memory");
}

// chunk stream: the KX_NB Wilke k-blocks, then the KX_N_DTILES diffusion tiles.  With KX_SPLIT == 2 every
// table block is streamed as two sub-chunks (Wilke: columns j < KX_J0 / the rest; tiles: rows i < KX_R0 / the
// rest) so that the double buffer is half as large and TWO 128-thread CTAs fit one SM's shared memory: one
// CTA's memory-bound prologue / epilogue then overlaps the other's FP64 phases.
#ifndef KX_WR
#define KX_WR 0
#endif
#if KX_WR > 0
// low-rank Wilke: the first 2 * KX_NWC chunks are row blocks (KX_WROWS species x KX_WR reals) of the factors V, U
#define KX_N_WCHUNKS (2 * KX_NWC)
#else
#define KX_N_WCHUNKS (KX_NB * KX_SPLIT)
#endif
KX_DEVICE const real* kx_chunk_src(int c)
{
#if KX_WR > 0
  if (c < KX_N_WCHUNKS) return (c < KX_NWC ? kx_wilke_v : kx_wilke_u) + (size_t)(c % KX_NWC) * (KX_WROWS * KX_WR);
  c -= KX_N_WCHUNKS;
  const int blk = c / KX_SPLIT, h = c % KX_SPLIT;
  return kx_diff + (size_t)blk * KX_DCHUNK + (h ? KX_R0 * KX_TB * 6 : 0);
#else
  const int blk = c / KX_SPLIT, h = c % KX_SPLIT;
  if (blk < KX_NB) return kx_wilke + (size_t)blk * KX_WCHUNK + (h ? KX_J0 * KX_TB : 0);
  return kx_diff + (size_t)(blk - KX_NB) * KX_DCHUNK + (h ? KX_R0 * KX_TB * 6 : 0);
#endif
}
KX_DEVICE unsigned kx_chunk_bytes(int c)
{
#if KX_WR > 0
  if (c < KX_N_WCHUNKS) {
    const int r0 = (c % KX_NWC) * KX_WROWS, r1 = min(KX_NP, r0 + KX_WROWS);
    return (unsigned)((r1 - r0) * KX_WR * sizeof(real));
  }
  const int h = (c - KX_N_WCHUNKS) % KX_SPLIT;
  const int whole = KX_DCHUNK, first = KX_R0 * KX_TB * 6;
#else
  const int blk = c / KX_SPLIT, h = c % KX_SPLIT;
  const int whole = blk < KX_NB ? KX_WCHUNK : KX_DCHUNK;
  const int first = blk < KX_NB ? KX_J0 * KX_TB : KX_R0 * KX_TB * 6;
#endif
  return (unsigned)(KX_SPLIT == 1 ? whole : (h ? whole - first : first)) * (unsigned)sizeof(real);
}

template <typename ST>   // ST: storage type of the state / result buffers (reference: dfloat)
__global__ void __launch_bounds__(KX_BK2_BLOCK, KX_BK2_MINB)
kx_bk2(const long long n_states, const long long offsetT, const long long offset, const real pressure,
       const ST* __restrict__ state, ST* __restrict__ conductivity, ST* __restrict__ viscosity,
       ST* __restrict__ rhoD, const double Tref)
{
  extern __shared__ __align__(16) unsigned char kx_sm_raw[];
  constexpr int LD = KX_BK2_BLOCK;
  constexpr int N_CHUNKS = KX_N_WCHUNKS + KX_N_DTILES * KX_SPLIT;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(kx_sm_raw);                   // 2 mbarriers (16 B)
  real* const buf0 = reinterpret_cast<real*>(kx_sm_raw + 16);                      // 2 x KX_CHUNK_MAX reals
  real* __restrict__ X = buf0 + 2 * KX_CHUNK_MAX + threadIdx.x;                    // X[k] at X[k * LD]
#if KX_BK2_SCRATCH
  // The per-state vectors b_k (Wilke phase) and S_k (diffusion phase) live in the thread's own rho*D output
  // rows, used as L2-resident scratch (ld/st .cg) until the final values are written: shared memory then
  // holds X only, which buys 50 % more resident warps (GRI-3.0: 12 instead of 8 per SM).  The loads of a
  // column block's S values are issued when its tile starts and consumed when it ends.
  ST* __restrict__ scr = rhoD;
#define KX_S_LOAD(k) ((live && (k) < KX_N) ? (real)__ldcg(scr + id + (size_t)(k) * offset) : (real)1)
#define KX_S_STORE(k, v)                                                     \
  do {                                                                       \
    if (live && (k) < KX_N) __stcg(scr + id + (size_t)(k) * offset, (ST)(v)); \
  } while (0)
#else
  real* __restrict__ S = X + KX_NP * LD;                        // b_k = 1/w_k, later the sums S_k
#define KX_S_LOAD(k) (S[(k) * LD])
#define KX_S_STORE(k, v) (S[(k) * LD] = (v))
#endif

  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < n_states;
  const long long id = live ? gid : n_states - 1;   // tail threads recompute the last state, store nothing

  if (threadIdx.x == 0) {
    kx_mbar_init(&bars[0], 1);
    kx_mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    kx_bulk_load(buf0, kx_chunk_src(0), kx_chunk_bytes(0), &bars[0]);
    if (N_CHUNKS > 1) kx_bulk_load(buf0 + KX_CHUNK_MAX, kx_chunk_src(1), kx_chunk_bytes(1), &bars[1]);
  }

  const double Td = Tref * (double)kx_ld_stream(state + id);
  const real T = (real)Td;
  const real lnT = (real)kx_log(Td);
  const real sqrT = kx_sqrt(T);
  const real lnT2 = lnT * lnT, lnT4 = lnT2 * lnT2;

  // ---- mole fractions (transportProps.okl:23-35) ----
  // The state rows are fetched in batches of up to 32 independent loads (one DRAM round trip per batch):
  // with a single CTA per SM nothing else hides this latency.
  real rcpMbar = 0;
  {
    const ST* sp = state + id + offsetT;
    constexpr int LB = 32;
#pragma unroll 1
    for (int k0 = 0; k0 < KX_N; k0 += LB) {
      ST y[LB];
#pragma unroll
      for (int i = 0; i < LB; i++)
        if (k0 + i < KX_N) y[i] = kx_ld_stream(sp + (size_t)(k0 + i) * offset);
#pragma unroll
      for (int i = 0; i < LB; i++) {
        if (k0 + i < KX_N) {
          const real yi = (real)y[i];
          const real w = (yi > (real)0 ? yi : (real)0) * kx_rcpM[k0 + i];
          X[(k0 + i) * LD] = w;
          rcpMbar += w;
        }
      }
    }
  }
  const real Mbar = kx_rcp(rcpMbar);

  // ---- conductivity, and per-species viscosity factors ----
  {
    real s1 = 0, s2 = 0;
#pragma unroll 8
    for (int k = 0; k < KX_N; k++) {
      const real x = X[k * LD] * Mbar;
      X[k * LD] = x;
      const real lam = kx_quartic(kx_cond[k], lnT);
      s1 = fma(x, lam, s1);
      s2 = fma(x, kx_rcp(lam), s2);
      const real v = kx_quartic(kx_visc[k], lnT);
      KX_S_STORE(k, kx_rcp(v * kx_m4[k]));                    // b_k = 1 / w_k
    }
    for (int k = KX_N; k < KX_NP; k++) { X[k * LD] = 0; KX_S_STORE(k, (real)1); }
    if (live) kx_st_stream(conductivity + id, (ST)(sqrT * ((real)0.5 * (s1 + kx_rcp(s2)))));
  }
  __syncthreads();   // mbarrier inits visible to all threads before the first wait

  int chunk = 0;
  // release the buffer of the chunk just consumed and refill it with chunk+2
  auto advance = [&]() {
    __syncthreads();
    if (threadIdx.x == 0 && chunk + 2 < N_CHUNKS) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      kx_bulk_load(buf0 + (chunk & 1) * KX_CHUNK_MAX, kx_chunk_src(chunk + 2), kx_chunk_bytes(chunk + 2),
                   &bars[chunk & 1]);
    }
    chunk++;
  };

#if KX_WR > 0
  // ---- viscosity: Wilke, three matrix-vector products with a LOW-RANK mass-factor matrix ----
  //      c_kj = (8 (1 + M_k/M_j))^-1/2 is a smooth kernel in ln M_k - ln M_j: numerical rank KX_WR (12 for GRI-3.0,
  //      14 for the 129-species EtOHKonnov), c = U V^T from an SVD at generation time (core/emit_module.py).
  //      t_m = V^T (X b^m),  Phi_k = U_k . (t_0 + 2 w_k t_1 + w_k^2 t_2):  6 N r instead of 3 N^2 multiply-adds.
  {
    constexpr int R = KX_WR;
    real t0[R], t1[R], t2[R];
#pragma unroll
    for (int q = 0; q < R; q++) t0[q] = t1[q] = t2[q] = 0;
#pragma unroll 1
    for (int c = 0; c < KX_NWC; c++) {
      kx_mbar_wait(&bars[chunk & 1], (chunk >> 1) & 1);
      const real* __restrict__ cv = buf0 + (chunk & 1) * KX_CHUNK_MAX;
      const int j1 = min(KX_NP, (c + 1) * KX_WROWS);
#pragma unroll 3
      for (int j = c * KX_WROWS; j < j1; j++, cv += R) {
        const real x = X[j * LD], b = KX_S_LOAD(j);
        const real xb = x * b, xbb = xb * b;
#pragma unroll
        for (int q = 0; q < R; q += 2) {
          const real2 vv = *reinterpret_cast<const real2*>(cv + q);
          t0[q] = fma(vv.x, x, t0[q]);
          t1[q] = fma(vv.x, xb, t1[q]);
          t2[q] = fma(vv.x, xbb, t2[q]);
          t0[q + 1] = fma(vv.y, x, t0[q + 1]);
          t1[q + 1] = fma(vv.y, xb, t1[q + 1]);
          t2[q + 1] = fma(vv.y, xbb, t2[q + 1]);
        }
      }
      advance();
    }
#pragma unroll
    for (int q = 0; q < R; q++) t1[q] += t1[q];
    real vis = 0;
#pragma unroll 1
    for (int c = 0; c < KX_NWC; c++) {
      kx_mbar_wait(&bars[chunk & 1], (chunk >> 1) & 1);
      const real* __restrict__ cu = buf0 + (chunk & 1) * KX_CHUNK_MAX;
      const int k1 = min(KX_N, (c + 1) * KX_WROWS);
#pragma unroll 2
      for (int k = c * KX_WROWS; k < k1; k++, cu += R) {
        const real v = kx_quartic(kx_visc[k], lnT);
        const real w = v * kx_m4[k], w2 = w * w;
        real ph[4] = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < R; q += 2) {
          const real2 uu = *reinterpret_cast<const real2*>(cu + q);
          ph[q & 2] = fma(uu.x, fma(w2, t2[q], fma(w, t1[q], t0[q])), ph[q & 2]);
          ph[(q & 2) + 1] = fma(uu.y, fma(w2, t2[q + 1], fma(w, t1[q + 1], t0[q + 1])), ph[(q & 2) + 1]);
        }
        const real phi = (ph[0] + ph[1]) + (ph[2] + ph[3]);
        vis = fma(X[k * LD] * (v * v), kx_rcp(phi), vis);
      }
      advance();
    }
    if (live) kx_st_stream(viscosity + id, (ST)(sqrT * vis));
  }
#else
  // ---- viscosity: Wilke with the three-matvec refactoring ----
  {
    real vis = 0;
    for (int kb = 0; kb < KX_NB; kb++) {
      real a0[KX_TB], a1[KX_TB], a2[KX_TB];
#pragma unroll
      for (int i = 0; i < KX_TB; i++) a0[i] = a1[i] = a2[i] = 0;
#pragma unroll
      for (int h = 0; h < KX_SPLIT; h++) {
        const int j_lo = h ? KX_J0 : 0, j_hi = (KX_SPLIT == 1 || h) ? KX_N : KX_J0;
        kx_mbar_wait(&bars[chunk & 1], (chunk >> 1) & 1);
        const real* __restrict__ cw = buf0 + (chunk & 1) * KX_CHUNK_MAX - j_lo * KX_TB;
#pragma unroll 2
        for (int j = j_lo; j < j_hi; j++) {
          const real x = X[j * LD], b = KX_S_LOAD(j);
          const real xb = x * b, xbb = xb * b;
#pragma unroll
          for (int i = 0; i < KX_TB; i++) {
            const real c = cw[j * KX_TB + i];
            a0[i] = fma(c, x, a0[i]);
            a1[i] = fma(c, xb, a1[i]);
            a2[i] = fma(c, xbb, a2[i]);
          }
        }
        if (h == KX_SPLIT - 1) {
#pragma unroll
          for (int i = 0; i < KX_TB; i++) {
            const int k = kb * KX_TB + i;
            if (k < KX_N) {
              const real v = kx_quartic(kx_visc[k], lnT);
              const real w = v * kx_m4[k];
              const real phi = fma(w, fma(w, a2[i], a1[i] + a1[i]), a0[i]);
              vis = fma(X[k * LD] * (v * v), kx_rcp(phi), vis);
            }
          }
        }
        advance();
      }
    }
    if (live) kx_st_stream(viscosity + id, (ST)(sqrT * vis));
  }

#endif

  // ---- mixture-averaged diffusion: S_k = sum_{j != k} X_j / D_kj, tiles of the lower triangle ----
  for (int kb = 0; kb < KX_NB; kb++) {
    real xk[KX_TB], sk[KX_TB];
#pragma unroll
    for (int i = 0; i < KX_TB; i++) { xk[i] = X[(kb * KX_TB + i) * LD]; sk[i] = 0; }
    for (int jb = 0; jb < kb; jb++) {
      real xj[KX_TB], sj[KX_TB];
#pragma unroll
      for (int i = 0; i < KX_TB; i++) {   // running sums of the column block: loaded now, needed after the tile
        xj[i] = X[(jb * KX_TB + i) * LD];
        sj[i] = KX_S_LOAD(jb * KX_TB + i);
      }
#pragma unroll
      for (int h = 0; h < KX_SPLIT; h++) {
        const int i_lo = h ? KX_R0 : 0, i_hi = (KX_SPLIT == 1 || h) ? KX_TB : KX_R0;
        kx_mbar_wait(&bars[chunk & 1], (chunk >> 1) & 1);
        const real2* __restrict__ tile =
            reinterpret_cast<const real2*>(buf0 + (chunk & 1) * KX_CHUNK_MAX) - i_lo * KX_TB * 3;
#pragma unroll
        for (int i = 0; i < KX_TB; i++) {
          if (i < i_lo || i >= i_hi) continue;
          // one tile row: KX_TB independent (quartic -> reciprocal) chains, then the two accumulations;
          // the row sum is split in two partial sums to halve its dependency chain
          real d[KX_TB];
#pragma unroll
          for (int j = 0; j < KX_TB; j++) d[j] = kx_pair_rcp_d(tile + (i * KX_TB + j) * 3, lnT, lnT2, lnT4);
          real se = 0, so = 0;
#pragma unroll
          for (int j = 0; j < KX_TB; j++) {
            if (j & 1) so = fma(xj[j], d[j], so); else se = fma(xj[j], d[j], se);
            sj[j] = fma(xk[i], d[j], sj[j]);
          }
          sk[i] += se + so;
        }
        if (h == KX_SPLIT - 1) {
#pragma unroll
          for (int i = 0; i < KX_TB; i++) KX_S_STORE(jb * KX_TB + i, sj[i]);
        }
        advance();
      }
    }
    // diagonal tile: pairs i > j inside the block
#pragma unroll
    for (int h = 0; h < KX_SPLIT; h++) {
      const int i_lo = h ? KX_R0 : 0, i_hi = (KX_SPLIT == 1 || h) ? KX_TB : KX_R0;
      kx_mbar_wait(&bars[chunk & 1], (chunk >> 1) & 1);
      const real2* __restrict__ tile =
          reinterpret_cast<const real2*>(buf0 + (chunk & 1) * KX_CHUNK_MAX) - i_lo * KX_TB * 3;
#pragma unroll
      for (int i = 1; i < KX_TB; i++) {
        if (i < i_lo || i >= i_hi) continue;
        real d[KX_TB];
#pragma unroll
        for (int j = 0; j < i; j++) d[j] = kx_pair_rcp_d(tile + (i * KX_TB + j) * 3, lnT, lnT2, lnT4);
        real se = 0, so = 0;
#pragma unroll
        for (int j = 0; j < i; j++) {
          if (j & 1) so = fma(xk[j], d[j], so); else se = fma(xk[j], d[j], se);
          sk[j] = fma(xk[i], d[j], sk[j]);
        }
        sk[i] += se + so;
      }
      if (h == KX_SPLIT - 1) {
        // first touch of this row block's sums: later row blocks add their column contributions
#pragma unroll
        for (int i = 0; i < KX_TB; i++) KX_S_STORE(kb * KX_TB + i, sk[i]);
      }
      advance();
    }
  }

  // ---- rho * D_km  (mix_transport.py:621-622 and transportProps.okl:43-47; p and Mbar cancel) ----
  if (live) {
    const real f = sqrT * (real)(1.0 / 8.31446261815324);      // rho*T^1.5/(p*Mbar) = sqrt(T)/R
    ST* out = rhoD + id;
#pragma unroll 4
    for (int k = 0; k < KX_N; k++) {
      const real num = fma(-kx_M[k], X[k * LD], Mbar);
      kx_st_stream(out + k * offset, (ST)(f * num * kx_rcp(KX_S_LOAD(k))));
    }
  }
  (void)pressure;
}
