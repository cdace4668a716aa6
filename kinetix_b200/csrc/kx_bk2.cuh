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
//   * the N(N-1)/2 binary-diffusion polynomials and the N^2 Wilke mass factors are TABLES (global memory,
//     L2 resident, read with warp-uniform 128-bit loads), walked by a small register-tiled loop nest:
//     TB x TB species tiles, accumulators for both the row block and the column block in registers
//     (each D_jk is evaluated once and used for S_j and S_k);  code size is a few KB -> I-cache resident;
//   * Wilke's sum is refactored:  (C1 + C2 v_k/v_j)^2 = c_kj (1 + w_k/w_j)^2 with w = v M^(-1/4) and
//     c_kj = 1/sqrt(8 (1 + M_k/M_j)), so
//        Phi_k = sum_j c_kj X_j  + 2 w_k sum_j c_kj X_j/w_j  + w_k^2 sum_j c_kj X_j/w_j^2
//     i.e. three constant-matrix x per-state-vector products: 3 N^2 DFMA instead of 4 N^2 mixed ops;
//   * divisions are MUFU.RCP64H + Newton (kx_rcp), rho*D_km is simplified algebraically (p and Mbar
//     cancel exactly as in the reference's formula, transportProps.okl:41-46).
//
// The including translation unit (generated per mechanism) defines:
//   KX_N, KX_NP (= KX_N rounded up to a multiple of KX_TB), KX_TB, KX_BK2_BLOCK, KX_BK2_MINB
//   __constant__ double kx_rcpM[KX_N], kx_M[KX_N], kx_m4[KX_N]  (1/M_k, M_k, M_k^(-1/4))
//   __constant__ double kx_cond[KX_N][5], kx_visc[KX_N][5]       quartics in ln T
//   __device__   double kx_wilke[KX_NP/KX_TB][KX_N][KX_TB]        c_kj, k-block major
//   __device__   double kx_diff[n_tiles][KX_TB*KX_TB][6]          lower-triangular tiles (kb >= jb),
//                                                                 row-major over (kb, jb); 5 coefs + pad
#pragma once
#include "kx_math.cuh"

#define KX_NB (KX_NP / KX_TB)

KX_DEVICE double kx_quartic(const double* __restrict__ c, double l)
{
  return fma(fma(fma(fma(c[4], l, c[3]), l, c[2]), l, c[1]), l, c[0]);
}

// 5 coefficients stored as 6 doubles, fetched with three 128-bit warp-uniform loads
KX_DEVICE double kx_quartic6(const double2* __restrict__ c, double l)
{
  const double2 a = __ldg(c), b = __ldg(c + 1), d = __ldg(c + 2);
  return fma(fma(fma(fma(d.x, l, b.y), l, b.x), l, a.y), l, a.x);
}

extern "C" __global__ void __launch_bounds__(KX_BK2_BLOCK, KX_BK2_MINB)
kx_bk2_f64(const long long n_states, const long long offsetT, const long long offset, const double pressure,
           const double* __restrict__ state, double* __restrict__ conductivity,
           double* __restrict__ viscosity, double* __restrict__ rhoD, const double Tref)
{
  extern __shared__ double kx_sm[];
  double* __restrict__ X = kx_sm + threadIdx.x;                          // X[k] at X[k * BLOCK]
  double* __restrict__ S = kx_sm + KX_NP * KX_BK2_BLOCK + threadIdx.x;   // b_k = 1/w_k, later the sums S_k
  constexpr int LD = KX_BK2_BLOCK;

  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < n_states;
  const long long id = live ? gid : n_states - 1;   // tail threads recompute the last state, store nothing

  const double T = Tref * kx_ld_stream(state + id);
  const double lnT = kx_log(T);
  const double rcpT = kx_rcp(T);
  const double sqrT = sqrt(T);

  // ---- mole fractions (transportProps.okl:23-35) ----
  double rcpMbar = 0.0;
  {
    const double* sp = state + id + offsetT;
#pragma unroll 8
    for (int k = 0; k < KX_N; k++) {
      const double w = fmax(0.0, kx_ld_stream(sp + k * offset)) * kx_rcpM[k];
      X[k * LD] = w;
      rcpMbar += w;
    }
  }
  const double Mbar = kx_rcp(rcpMbar);

  // ---- conductivity, and per-species viscosity factors ----
  {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll 4
    for (int k = 0; k < KX_N; k++) {
      const double x = X[k * LD] * Mbar;
      X[k * LD] = x;
      const double lam = kx_quartic(kx_cond[k], lnT);
      s1 = fma(x, lam, s1);
      s2 = fma(x, kx_rcp(lam), s2);
      const double v = kx_quartic(kx_visc[k], lnT);
      S[k * LD] = kx_rcp(v * kx_m4[k]);                       // b_k = 1 / w_k
    }
    for (int k = KX_N; k < KX_NP; k++) { X[k * LD] = 0.0; S[k * LD] = 1.0; }
    if (live) kx_st_stream(conductivity + id, sqrT * (0.5 * (s1 + kx_rcp(s2))));
  }

  // ---- viscosity: Wilke with the three-matvec refactoring ----
  {
    double vis = 0.0;
    for (int kb = 0; kb < KX_NB; kb++) {
      double a0[KX_TB], a1[KX_TB], a2[KX_TB];
#pragma unroll
      for (int i = 0; i < KX_TB; i++) a0[i] = a1[i] = a2[i] = 0.0;
      const double* __restrict__ cw = kx_wilke + (size_t)kb * KX_N * KX_TB;
#pragma unroll 2
      for (int j = 0; j < KX_N; j++) {
        const double x = X[j * LD], b = S[j * LD];
        const double xb = x * b, xbb = xb * b;
#pragma unroll
        for (int i = 0; i < KX_TB; i++) {
          const double c = __ldg(cw + j * KX_TB + i);
          a0[i] = fma(c, x, a0[i]);
          a1[i] = fma(c, xb, a1[i]);
          a2[i] = fma(c, xbb, a2[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < KX_TB; i++) {
        const int k = kb * KX_TB + i;
        if (k < KX_N) {
          const double v = kx_quartic(kx_visc[k], lnT);
          const double w = v * kx_m4[k];
          const double phi = fma(w, fma(w, a2[i], a1[i] + a1[i]), a0[i]);
          vis = fma(X[k * LD] * (v * v), kx_rcp(phi), vis);
        }
      }
    }
    if (live) kx_st_stream(viscosity + id, sqrT * vis);
  }

  // ---- mixture-averaged diffusion: S_k = sum_{j != k} X_j / D_kj, tiles of the lower triangle ----
  for (int k = 0; k < KX_NP; k++) S[k * LD] = 0.0;
  {
    const double2* __restrict__ tile = reinterpret_cast<const double2*>(kx_diff);
    for (int kb = 0; kb < KX_NB; kb++) {
      double xk[KX_TB], sk[KX_TB];
#pragma unroll
      for (int i = 0; i < KX_TB; i++) { xk[i] = X[(kb * KX_TB + i) * LD]; sk[i] = 0.0; }
      for (int jb = 0; jb < kb; jb++) {
        double xj[KX_TB], sj[KX_TB];
#pragma unroll
        for (int i = 0; i < KX_TB; i++) { xj[i] = X[(jb * KX_TB + i) * LD]; sj[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < KX_TB; i++) {
#pragma unroll
          for (int j = 0; j < KX_TB; j++) {
            const double d = kx_rcp(kx_quartic6(tile + (i * KX_TB + j) * 3, lnT));
            sk[i] = fma(xj[j], d, sk[i]);
            sj[j] = fma(xk[i], d, sj[j]);
          }
        }
#pragma unroll
        for (int i = 0; i < KX_TB; i++) S[(jb * KX_TB + i) * LD] += sj[i];
        tile += KX_TB * KX_TB * 3;
      }
      // diagonal tile: pairs i > j inside the block
#pragma unroll
      for (int i = 1; i < KX_TB; i++) {
#pragma unroll
        for (int j = 0; j < i; j++) {
          const double d = kx_rcp(kx_quartic6(tile + (i * KX_TB + j) * 3, lnT));
          sk[i] = fma(xk[j], d, sk[i]);
          sk[j] = fma(xk[i], d, sk[j]);
        }
      }
      tile += KX_TB * KX_TB * 3;
#pragma unroll
      for (int i = 0; i < KX_TB; i++) S[(kb * KX_TB + i) * LD] += sk[i];
    }
  }

  // ---- rho * D_km  (mix_transport.py:621-622 and transportProps.okl:43-47; p and Mbar cancel) ----
  if (live) {
    const double f = sqrT * (1.0 / 8.31446261815324);          // rho*T^1.5/(p*Mbar) = sqrt(T)/R
    double* out = rhoD + id;
#pragma unroll 4
    for (int k = 0; k < KX_N; k++) {
      const double num = fma(-kx_M[k], X[k * LD], Mbar);
      kx_st_stream(out + k * offset, f * num * kx_rcp(S[k * LD]));
    }
  }
  (void)pressure; (void)rcpT;
}
