// kx_thermo.cuh -- density, species heat capacities and rho*cp, one thread per state (HBM bound).
//
// Same results as the reference's `thermoCoeffs` OKL kernel around kinetix_molar_heat_capacity_R
// (reference benchmark/okl/thermoCoeffs.okl:10-40, kinetix/core/thermodynamics.py:83-97), in ONE pass
// over the state rows: sum_k w_k and sum_k cp_k/R w_k are accumulated together, so no per-thread array
// is needed (rhoCp = rho R sum_k (cp_k/R) w_k; the reference's Mbar * rcpMbar factor cancels).
//
// The including translation unit defines KX_N, the arithmetic type `real` (double, or float for the
// --single-precision module) and
//   __constant__ real kx_rcpM[KX_N], kx_Tmid[KX_N], kx_nasa[KX_N][2][7]   (low / high range)
// S is the storage type of the state / result buffers (reference: dfloat).
#pragma once
#include "kx_math.cuh"

template <typename S, bool PF>   // PF: per-state pressure field (extension)
__global__ void __launch_bounds__(256)
kx_thermo(const long long n_states, const long long offsetT, const long long offset, const real pressure_R,
          const S* __restrict__ state, S* __restrict__ rho, S* __restrict__ cp, S* __restrict__ rhoCp,
          const double Tref, const S* __restrict__ pfield)
{
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_states) return;
  const real R = (real)8.31446261815324;
  const real T = (real)(Tref * (double)kx_ld_stream(state + id));
  const S* sp = state + id + offsetT;
  S* cpo = cp + id;
  real rcpMbar = 0, cpw = 0;
#pragma unroll 8
  for (int k = 0; k < KX_N; k++) {
    const real y = (real)kx_ld_stream(sp + k * offset);
    const real w = (y > (real)0 ? y : (real)0) * kx_rcpM[k];
    const real* a = kx_nasa[k][T <= kx_Tmid[k] ? 0 : 1];
    const real cpR = fma(fma(fma(fma(a[4], T, a[3]), T, a[2]), T, a[1]), T, a[0]);
    kx_st_stream(cpo + k * offset, (S)(cpR * R * kx_rcpM[k]));
    rcpMbar += w;
    cpw = fma(cpR, w, cpw);
  }
  // per-state pressure (extension): pfield[id] = p / p_ref, pressure_R then carries p_ref / R
  const real pR = PF ? pressure_R * (real)kx_ld_stream(pfield + id) : pressure_R;
  const real rho_ = pR * kx_rcp(T) * kx_rcp(rcpMbar);
  kx_st_stream(rho + id, (S)rho_);
  kx_st_stream(rhoCp + id, (S)(rho_ * (R * cpw)));
}
