// kx_thermo.cuh -- density, species heat capacities and rho*cp, one thread per state (HBM bound).
//
// Same results as the reference's `thermoCoeffs` OKL kernel around kinetix_molar_heat_capacity_R
// (reference benchmark/okl/thermoCoeffs.okl:10-40, kinetix/core/thermodynamics.py:83-97), in ONE pass
// over the state rows: sum_k w_k and sum_k cp_k/R w_k are accumulated together, so no per-thread array
// is needed (rhoCp = rho R sum_k (cp_k/R) w_k; the reference's Mbar * rcpMbar factor cancels).
//
// The including translation unit defines KX_N and
//   __constant__ double kx_rcpM[KX_N], kx_Tmid[KX_N], kx_nasa[KX_N][2][7]   (low / high range)
#pragma once
#include "kx_math.cuh"

extern "C" __global__ void __launch_bounds__(256)
kx_thermo_f64(const long long n_states, const long long offsetT, const long long offset, const double pressure_R,
              const double* __restrict__ state, double* __restrict__ rho, double* __restrict__ cp,
              double* __restrict__ rhoCp, const double Tref)
{
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_states) return;
  const double R = 8.31446261815324;
  const double T = Tref * kx_ld_stream(state + id);
  const double* sp = state + id + offsetT;
  double* cpo = cp + id;
  double rcpMbar = 0.0, cpw = 0.0;
#pragma unroll 8
  for (int k = 0; k < KX_N; k++) {
    const double w = fmax(0.0, kx_ld_stream(sp + k * offset)) * kx_rcpM[k];
    const double* a = kx_nasa[k][T <= kx_Tmid[k] ? 0 : 1];
    const double cpR = fma(fma(fma(fma(a[4], T, a[3]), T, a[2]), T, a[1]), T, a[0]);
    kx_st_stream(cpo + k * offset, cpR * R * kx_rcpM[k]);
    rcpMbar += w;
    cpw = fma(cpR, w, cpw);
  }
  const double rho_ = pressure_R * kx_rcp(T) * kx_rcp(rcpMbar);
  kx_st_stream(rho + id, rho_);
  kx_st_stream(rhoCp + id, rho_ * (R * cpw));
}
