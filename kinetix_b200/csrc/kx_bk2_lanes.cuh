// kx_bk2_lanes.cuh -- BK2 (mixture-averaged conductivity, viscosity, rho*D_km) with KX_L lanes per state.
//
// STATUS: opt-in variant (emit option bk2_lanes = 2 | 4), parity-tested (tests/test_parity_gpu.py) but NOT the
// default: on the B200 it is slower than both the one-state-per-thread kernel and the tensor-memory kernel
// (GRI-3.0: 336 / 310 vs 446 / 610 M states/s; EtOHKonnov 50 / 31 vs 57 / 97) because the pair loops are bound by
// shared-memory wavefronts, not by occupancy: more warps issue the same coefficient loads.
//
// Same arithmetic as csrc/kx_bk2.cuh (reference benchmark/okl/transportProps.okl:11-49 around
// kinetix/core/mix_transport.py:474-626); what changes is the mapping of work to threads.
//
// Why: the per-state vectors X_k and S_k (2 x NP reals) must stay in shared memory, so an SM holds only
// 227 KB / (16 NP) states: 256 for GRI-3.0, 104 for the 129-species EtOHKonnov.  With one thread per state
// that is 8 (3) warps per SM, and the FP64 pipe starves on dependent-issue and shared-memory latency
// (profiles/ncu_r01_bk2_v2.txt: 2 warps/SMSP, FP64 pipe 58 %, `wait` the top stall).  Here KX_L = 2 or 4
// neighbouring lanes of a warp share one state: every lane owns the species k with k mod KX_L == its
// position in the group -- for the state rows it loads, the tile COLUMNS it evaluates in the pair loops, the
// j-terms it adds in the Wilke sums and the output rows it stores -- so the same shared memory carries
// KX_L times as many warps, and each lane holds 1/KX_L of the register tile.
//   * Row sums (over a lane's columns) are partial per lane and combined once per row block with a
//     recursive-halving exchange over __shfl_xor (KX_TB/2 shuffles for 2 lanes), which also hands every
//     species' total to its owner lane; column sums are owned, so S_k read-modify-writes never race.
//   * The table stream (Wilke k-blocks, then diffusion tiles) goes through a KX_STAGES-deep ring of TMA
//     bulk copies with full/empty mbarriers instead of a CTA-wide barrier per chunk: warps drift apart by up
//     to KX_STAGES - 1 chunks, so one warp's barrier wait overlaps the others' FP64 work.
//   * per-species constants come from one 128-byte record per species in global memory (L1 resident);
//     a __constant__ table would serialise because the lanes of a group index different species.
//
// The including translation unit defines KX_N, KX_NP (multiple of KX_TB), KX_TB (multiple of KX_L), KX_L,
// KX_P (states per lane group: each table coefficient fetched from shared memory is used for KX_P states),
// KX_BK2_BLOCK (threads; KX_BK2_BLOCK / KX_L * KX_P states per CTA), KX_STAGES (power of two), KX_CHUNK_MAX,
// KX_WCHUNK (= KX_NP * KX_TB, padded), KX_DCHUNK (= KX_TB * KX_TB * 6, padded), KX_RCP_DIFF and the tables
//   __device__ real kx_sptab[KX_NP][16]   {1/M, M, M^-1/4, -, cond[0..4], -, visc[0..4], -}; padded species:
//                                         1/M = 0, M^-1/4 = 1, both quartics = 1
//   __device__ real kx_wilke[KX_NB][KX_WCHUNK]   c_kj as [kb][j][i], k = kb*TB + i, j < KX_NP (zero padded)
//   __device__ real kx_diff[n_tiles][KX_DCHUNK]  lower-triangular tiles, row-major over (kb, jb); 5 coefs + pad
#pragma once
#include <cstdint>
#include "kx_math.cuh"

#define KX_NB (KX_NP / KX_TB)
#define KX_N_DTILES (KX_NB * (KX_NB + 1) / 2)

KX_DEVICE real kx_pair_poly(const real2* __restrict__ c, real l, real l2, real l4)
{
  const real2 c01 = c[0], c23 = c[1], c4 = c[2];
  return fma(c4.x, l4, fma(fma(c23.y, l, c23.x), l2, fma(c01.y, l, c01.x)));
}
KX_DEVICE real kx_pair_rcp_d(const real2* __restrict__ c, real l, real l2, real l4)
{
#if KX_RCP_DIFF
  return kx_pair_poly(c, l, l2, l4);   // --fit-rcpdiffcoeffs: the fit is 1/D already (mix_transport.py:198-206)
#else
  return kx_rcp(kx_pair_poly(c, l, l2, l4));
#endif
}
// quartic from a species record: {c0,c1} {c2,c3} {c4,-}
KX_DEVICE real kx_rec_quartic(const real2 c01, const real2 c23, const real2 c4, real l)
{
  return fma(fma(fma(fma(c4.x, l, c23.y), l, c23.x), l, c01.y), l, c01.x);
}

#include "kx_pipe.cuh"

KX_DEVICE const real* kx_chunk_src(int c)
{
  return c < KX_NB ? kx_wilke + (size_t)c * KX_WCHUNK : kx_diff + (size_t)(c - KX_NB) * KX_DCHUNK;
}
KX_DEVICE unsigned kx_chunk_bytes(int c) { return (unsigned)((c < KX_NB ? KX_WCHUNK : KX_DCHUNK) * sizeof(real)); }

// Sum a[] over the KX_L lanes of a group and scatter: lane h receives the totals of entries KX_L*c + h.
template <int CNT>
KX_DEVICE void kx_reduce_scatter(const real (&a)[CNT], real (&mine)[CNT / KX_L], int h)
{
  static_assert(KX_L == 1 || KX_L == 2 || KX_L == 4, "lanes per state");
#pragma unroll
  for (int c = 0; c < CNT / KX_L; c++) {
    if (KX_L == 1) {
      mine[c] = a[c];
    } else if (KX_L == 2) {
      const real keep = h ? a[2 * c + 1] : a[2 * c], send = h ? a[2 * c] : a[2 * c + 1];
      mine[c] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    } else {
      const bool up = h & 2, odd = h & 1;
      real k0 = up ? a[4 * c + 2] : a[4 * c], k1 = up ? a[4 * c + 3] : a[4 * c + 1];
      const real s0 = up ? a[4 * c] : a[4 * c + 2], s1 = up ? a[4 * c + 1] : a[4 * c + 3];
      k0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      k1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const real keep = odd ? k1 : k0, send = odd ? k0 : k1;
      mine[c] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
  }
}
KX_DEVICE real kx_group_sum(real v)
{
#pragma unroll
  for (int m = KX_L / 2; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

template <typename ST>   // ST: storage type of the state / result buffers (reference: dfloat)
__global__ void __launch_bounds__(KX_BK2_BLOCK, 1)
kx_bk2(const long long n_states, const long long offsetT, const long long offset, const real pressure,
       const ST* __restrict__ state, ST* __restrict__ conductivity, ST* __restrict__ viscosity,
       ST* __restrict__ rhoD, const double Tref)
{
  extern __shared__ __align__(16) unsigned char kx_sm_raw[];
  // G groups of L lanes; every group carries P states (slots g, g + G, ...): LD = G * P states per CTA
  constexpr int L = KX_L, P = KX_P, CJ = KX_TB / L, MJ = KX_NP / L;
  constexpr int GS = KX_BK2_BLOCK / L, G = GS, LD = GS * P;
  constexpr int NW = KX_BK2_BLOCK / 32, STG = KX_STAGES;
  constexpr int N_CHUNKS = KX_NB + KX_N_DTILES;
  static_assert(KX_TB % L == 0 && (STG & (STG - 1)) == 0, "tile / ring shape");
  uint64_t* const full = reinterpret_cast<uint64_t*>(kx_sm_raw);           // STG full + STG empty barriers
  uint64_t* const empty = full + STG;
  real* const buf0 = reinterpret_cast<real*>(kx_sm_raw + 16 * STG);        // STG x KX_CHUNK_MAX reals
  const int h = threadIdx.x & (L - 1), g = threadIdx.x / L;
  real* __restrict__ X = buf0 + STG * KX_CHUNK_MAX + g;                    // X[k] of state p at X[k * LD + p * G]
  real* __restrict__ S = X + KX_NP * LD;                                   // b_k = 1/w_k, later the sums S_k

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STG; s++) { kx_mbar_init(&full[s], 1); kx_mbar_init(&empty[s], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int s = 0; s < STG; s++)
      if (s < N_CHUNKS) kx_bulk_load(buf0 + s * KX_CHUNK_MAX, kx_chunk_src(s), kx_chunk_bytes(s), &full[s]);
  }

  bool live[P];
  long long id[P];
  real lnT[P], lnT2[P], lnT4[P], sqrT[P], Mbar[P];
#pragma unroll
  for (int p = 0; p < P; p++) {
    const long long gid = (long long)blockIdx.x * (GS * P) + p * GS + g;
    live[p] = gid < n_states;
    id[p] = live[p] ? gid : n_states - 1;   // tail slots recompute the last state, store nothing
    const double Td = Tref * (double)kx_ld_stream(state + id[p]);
    lnT[p] = (real)kx_log(Td);
    sqrT[p] = kx_sqrt((real)Td);
    lnT2[p] = lnT[p] * lnT[p];
    lnT4[p] = lnT2[p] * lnT2[p];
  }
  const real2* __restrict__ rec = reinterpret_cast<const real2*>(kx_sptab) + h * 8;   // record of species L*m + h

  // ---- mole fractions (transportProps.okl:23-35): lane h loads the rows k = L*m + h ----
#pragma unroll
  for (int p = 0; p < P; p++) {
    real rcpMbar = 0;
    const ST* sp = state + id[p] + offsetT + (size_t)h * offset;
    constexpr int LB = 16;
#pragma unroll 1
    for (int m0 = 0; m0 < MJ; m0 += LB) {
      ST y[LB];
#pragma unroll
      for (int i = 0; i < LB; i++)
        if (m0 + i < MJ) y[i] = (L * (m0 + i) + h < KX_N) ? kx_ld_stream(sp + (size_t)(L * (m0 + i)) * offset) : (ST)0;
#pragma unroll
      for (int i = 0; i < LB; i++) {
        if (m0 + i < MJ) {
          const real yi = (real)y[i];
          const real w = (yi > (real)0 ? yi : (real)0) * __ldg(&rec[(m0 + i) * L * 8].x);
          X[(L * (m0 + i) + h) * LD + p * G] = w;
          rcpMbar += w;
        }
      }
    }
    Mbar[p] = kx_rcp(kx_group_sum(rcpMbar));
  }

  // ---- conductivity, and per-species viscosity factors (own species only) ----
  {
    real s1[P], s2[P];
#pragma unroll
    for (int p = 0; p < P; p++) s1[p] = s2[p] = 0;
#pragma unroll 2
    for (int m = 0; m < MJ; m++) {
      const int k = L * m + h;
      const real2* r = rec + m * L * 8;
      const real2 l01 = __ldg(r + 2), l23 = __ldg(r + 3), l4 = __ldg(r + 4);
      const real2 v01 = __ldg(r + 5), v23 = __ldg(r + 6), v4 = __ldg(r + 7);
      const real m4 = __ldg(&r[1].x);
#pragma unroll
      for (int p = 0; p < P; p++) {
        const real x = X[k * LD + p * G] * Mbar[p];
        X[k * LD + p * G] = x;
        const real lam = kx_rec_quartic(l01, l23, l4, lnT[p]);
        s1[p] = fma(x, lam, s1[p]);
        s2[p] = fma(x, kx_rcp(lam), s2[p]);
        const real v = kx_rec_quartic(v01, v23, v4, lnT[p]);
        S[k * LD + p * G] = kx_rcp(v * m4);                    // b_k = 1 / w_k
      }
    }
#pragma unroll
    for (int p = 0; p < P; p++) {
      const real t1 = kx_group_sum(s1[p]), t2 = kx_group_sum(s2[p]);
      if (live[p] && h == 0) kx_st_stream(conductivity + id[p], (ST)(sqrT[p] * ((real)0.5 * (t1 + kx_rcp(t2)))));
    }
  }
  __syncthreads();   // mbarrier inits and every lane's X rows visible to the whole CTA

  int chunk = 0;
  auto acquire = [&]() -> const real* {
    kx_mbar_wait(&full[chunk & (STG - 1)], (chunk / STG) & 1);
    return buf0 + (chunk & (STG - 1)) * KX_CHUNK_MAX;
  };
  // hand the stage back; thread 0 then refills a stage with the chunk STG - LAG ahead: with more than two
  // stages the stage of the PREVIOUS chunk (which the other warps have normally left already), else this one
  constexpr int LAG = STG > 2 ? 1 : 0;
  auto release = [&]() {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) kx_mbar_arrive(&empty[chunk & (STG - 1)]);
    if (threadIdx.x == 0 && chunk >= LAG && chunk - LAG + STG < N_CHUNKS) {
      const int c2 = chunk - LAG + STG, s2 = c2 & (STG - 1);
      kx_mbar_wait(&empty[s2], (c2 / STG - 1) & 1);
      kx_bulk_load(buf0 + s2 * KX_CHUNK_MAX, kx_chunk_src(c2), kx_chunk_bytes(c2), &full[s2]);
    }
    chunk++;
  };

  // ---- viscosity: Wilke with the three-matvec refactoring; lane h adds the terms j = L*m + h ----
  {
    real vis[P];
#pragma unroll
    for (int p = 0; p < P; p++) vis[p] = 0;
#pragma unroll 1
    for (int kb = 0; kb < KX_NB; kb++) {
      real a0[P][KX_TB], a1[P][KX_TB], a2[P][KX_TB];
#pragma unroll
      for (int p = 0; p < P; p++)
#pragma unroll
        for (int i = 0; i < KX_TB; i++) a0[p][i] = a1[p][i] = a2[p][i] = 0;
      const real* __restrict__ cw = acquire() + h * KX_TB;
#pragma unroll 2
      for (int m = 0; m < MJ; m++) {
        real x[P], xb[P], xbb[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
          x[p] = X[(L * m + h) * LD + p * G];
          const real b = S[(L * m + h) * LD + p * G];
          xb[p] = x[p] * b;
          xbb[p] = xb[p] * b;
        }
        const real* __restrict__ cj = cw + m * (L * KX_TB);
#pragma unroll
        for (int i = 0; i < KX_TB; i += 2) {
          real c0, c1;
          if ((KX_TB & 1) == 0 && sizeof(real) == 8) {
            const real2 cc = *reinterpret_cast<const real2*>(cj + i);
            c0 = cc.x; c1 = cc.y;
          } else {
            c0 = cj[i]; c1 = (i + 1 < KX_TB) ? cj[i + 1] : (real)0;
          }
#pragma unroll
          for (int p = 0; p < P; p++) {
            a0[p][i] = fma(c0, x[p], a0[p][i]);
            a1[p][i] = fma(c0, xb[p], a1[p][i]);
            a2[p][i] = fma(c0, xbb[p], a2[p][i]);
            if (i + 1 < KX_TB) {
              a0[p][i + 1] = fma(c1, x[p], a0[p][i + 1]);
              a1[p][i + 1] = fma(c1, xb[p], a1[p][i + 1]);
              a2[p][i + 1] = fma(c1, xbb[p], a2[p][i + 1]);
            }
          }
        }
      }
      release();
#pragma unroll
      for (int p = 0; p < P; p++) {
        real m0[CJ], m1[CJ], m2[CJ];
        kx_reduce_scatter<KX_TB>(a0[p], m0, h);
        kx_reduce_scatter<KX_TB>(a1[p], m1, h);
        kx_reduce_scatter<KX_TB>(a2[p], m2, h);
#pragma unroll
        for (int c = 0; c < CJ; c++) {
          const int k = kb * KX_TB + L * c + h;
          if (k < KX_N) {
            const real2* r = reinterpret_cast<const real2*>(kx_sptab) + k * 8;
            const real v = kx_rec_quartic(__ldg(r + 5), __ldg(r + 6), __ldg(r + 7), lnT[p]);
            const real w = v * __ldg(&r[1].x);
            const real phi = fma(w, fma(w, m2[c], m1[c] + m1[c]), m0[c]);
            vis[p] = fma(X[k * LD + p * G] * (v * v), kx_rcp(phi), vis[p]);
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < P; p++) {
      const real t = kx_group_sum(vis[p]);
      if (live[p] && h == 0) kx_st_stream(viscosity + id[p], (ST)(sqrT[p] * t));
    }
  }

  // ---- mixture-averaged diffusion: S_k = sum_{j != k} X_j / D_kj over tiles of the lower triangle; lane h
  //      evaluates the tile columns L*c + h; the coefficients of a pair are loaded once for the P states ----
#pragma unroll 1
  for (int kb = 0; kb < KX_NB; kb++) {
    real xk[P][KX_TB], sk[P][KX_TB], xo[P][CJ], so[P][CJ];
#pragma unroll
    for (int p = 0; p < P; p++) {
#pragma unroll
      for (int i = 0; i < KX_TB; i++) { xk[p][i] = X[(kb * KX_TB + i) * LD + p * G]; sk[p][i] = 0; }
#pragma unroll
      for (int c = 0; c < CJ; c++) { xo[p][c] = X[(kb * KX_TB + L * c + h) * LD + p * G]; so[p][c] = 0; }
    }
#pragma unroll 1
    for (int jb = 0; jb < kb; jb++) {
      real xj[P][CJ], sj[P][CJ];
#pragma unroll
      for (int p = 0; p < P; p++)
#pragma unroll
        for (int c = 0; c < CJ; c++) {   // running sums of the owned columns: loaded now, needed after the tile
          xj[p][c] = X[(jb * KX_TB + L * c + h) * LD + p * G];
          sj[p][c] = S[(jb * KX_TB + L * c + h) * LD + p * G];
        }
      const real2* __restrict__ tile = reinterpret_cast<const real2*>(acquire()) + h * 3;
#pragma unroll
      for (int i = 0; i < KX_TB; i++) {
        real d[P][CJ];
#pragma unroll
        for (int c = 0; c < CJ; c++) {
          const real2* cp = tile + (i * KX_TB + L * c) * 3;
          const real2 c01 = cp[0], c23 = cp[1], c4 = cp[2];
#pragma unroll
          for (int p = 0; p < P; p++) {
            const real q = fma(c4.x, lnT4[p], fma(fma(c23.y, lnT[p], c23.x), lnT2[p], fma(c01.y, lnT[p], c01.x)));
            d[p][c] = KX_RCP_DIFF ? q : kx_rcp(q);
          }
        }
#pragma unroll
        for (int p = 0; p < P; p++) {
          real se = 0, sod = 0;
#pragma unroll
          for (int c = 0; c < CJ; c++) {
            if (c & 1) sod = fma(xj[p][c], d[p][c], sod); else se = fma(xj[p][c], d[p][c], se);
            sj[p][c] = fma(xk[p][i], d[p][c], sj[p][c]);
          }
          sk[p][i] += CJ > 1 ? se + sod : se;
        }
      }
      release();
#pragma unroll
      for (int p = 0; p < P; p++)
#pragma unroll
        for (int c = 0; c < CJ; c++) S[(jb * KX_TB + L * c + h) * LD + p * G] = sj[p][c];
    }
    // diagonal tile: pairs i > j inside the block; column sums go to so[] (owned), row sums to sk[] (partial)
    {
      const real2* __restrict__ tile = reinterpret_cast<const real2*>(acquire()) + h * 3;
#pragma unroll
      for (int i = 1; i < KX_TB; i++) {
#pragma unroll
        for (int c = 0; c < CJ; c++) {
          if (L * c < i) {
            const real2* cp = tile + (i * KX_TB + L * c) * 3;
            const real2 c01 = cp[0], c23 = cp[1], c4 = cp[2];
#pragma unroll
            for (int p = 0; p < P; p++) {
              const real q = fma(c4.x, lnT4[p], fma(fma(c23.y, lnT[p], c23.x), lnT2[p], fma(c01.y, lnT[p], c01.x)));
              real d = KX_RCP_DIFF ? q : kx_rcp(q);
              if (L * c + L - 1 >= i) d = (L * c + h < i) ? d : (real)0;   // columns j >= i of this row: not a pair
              sk[p][i] = fma(xo[p][c], d, sk[p][i]);
              so[p][c] = fma(xk[p][i], d, so[p][c]);
            }
          }
        }
      }
      release();
    }
    // first touch of this row block's sums: later row blocks add their column contributions
#pragma unroll
    for (int p = 0; p < P; p++) {
      real tot[CJ];
      kx_reduce_scatter<KX_TB>(sk[p], tot, h);
#pragma unroll
      for (int c = 0; c < CJ; c++) S[(kb * KX_TB + L * c + h) * LD + p * G] = tot[c] + so[p][c];
    }
  }

  // ---- rho * D_km  (mix_transport.py:621-622 and transportProps.okl:43-47; p and Mbar cancel) ----
#pragma unroll
  for (int p = 0; p < P; p++) {
    if (live[p]) {
      const real f = sqrT[p] * (real)(1.0 / 8.31446261815324);      // rho*T^1.5/(p*Mbar) = sqrt(T)/R
      ST* out = rhoD + id[p] + (size_t)h * offset;
#pragma unroll 4
      for (int m = 0; m < MJ; m++) {
        const int k = L * m + h;
        if (k < KX_N) {
          const real num = fma(-__ldg(&rec[m * L * 8].y), X[k * LD + p * G], Mbar[p]);
          kx_st_stream(out + (size_t)(L * m) * offset, (ST)(f * num * kx_rcp(S[k * LD + p * G])));
        }
      }
    }
  }
  (void)pressure;
}
