// kx_routine_kernels.cu -- the reference's three per-state kernels as plain CUDA __global__ functions around the
// GENERATED reference-signature routines (kinetix_b200/core/emit_routines.py).
//
// This is what a caller that force-includes the generated files gets (reference benchmark/okl/productionRates.okl:1-65,
// transportProps.okl:1-50, thermoCoeffs.okl:1-41 do exactly this through OCCA): one thread = one state, the state's
// vectors in thread-local arrays, the arithmetic inside kinetix_species_rates / kinetix_enthalpy_RT /
// kinetix_conductivity / kinetix_viscosity / kinetix_diffusivity / kinetix_molar_heat_capacity_R.  It exists to
//   (1) prove on the GPU that the emitted routines compute what the reference's routines compute
//       (tests/test_routines_gpu.py compares these kernels with oracle/_ref at 1e-10), and
//   (2) measure what the routine flavour costs against the native kernels of the same mechanism.
// Compile with -I<routine directory> (kinetix_b200.jit.ensure_routines does); exports kxr_* launchers.
#include <cuda_runtime.h>

#include "kinetix_b200_routines.cuh"

#ifndef p_R
#define p_R (1.380649e-23 * 6.02214076e23)   // kinetix.cpp:37,194
#endif
#ifndef p_BLOCKSIZE
#define p_BLOCKSIZE 128
#endif

namespace {

// mass fractions -> w_k = max(Y_k, 0) / M_k into `w`, returns sum_k w_k = 1 / Mbar
__device__ __forceinline__ cfloat load_composition(const dfloat* __restrict__ state, long long id, long long offsetT,
                                                   long long offset, cfloat* w)
{
  cfloat rcpMbar = 0;
#pragma unroll
  for (int k = 0; k < __KINETIX_NSPECIES__; k++) {
    const cfloat y = (cfloat)state[id + offsetT + k * offset];
    w[k] = (y > (cfloat)0 ? y : (cfloat)0) * kinetix_rcp_molar_mass[k];
    rcpMbar += w[k];
  }
  return rcpMbar;
}

__global__ void __launch_bounds__(p_BLOCKSIZE)
productionRates(long long n_states, long long offsetT, long long offset, double pressure_R, double pressure_,
                const dfloat* __restrict__ state, dfloat* __restrict__ rates, double Tref)
{
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_states) return;
  const cfloat T = Tref * state[id];
  const cfloat rcpT = kx_rcp(T), lnT = kx_log(T);
  const cfloat T2 = T * T, T3 = T2 * T, T4 = T2 * T2;
  cfloat conc[__KINETIX_NSPECIES__], wdot[__KINETIX_NSPECIES__];
  const cfloat rcpMbar = load_composition(state, id, offsetT, offset, conc);
  const cfloat rho = pressure_R * rcpT * kx_rcp(rcpMbar);
#pragma unroll
  for (int k = 0; k < __KINETIX_NSPECIES__; k++) {
    conc[k] *= rho;       // molar concentrations [mol/m^3]
    wdot[k] = 0;          // the routine accumulates
  }
  kinetix_species_rates(lnT, T, T2, T3, T4, rcpT, (cfloat)pressure_, (cfloat)kx_log(pressure_), conc, wdot);
#pragma unroll
  for (int k = 0; k < __KINETIX_NSPECIES__; k++) rates[id + offsetT + k * offset] = kinetix_molar_mass[k] * wdot[k];
  kinetix_enthalpy_RT(T, T2, T3, T4, rcpT, conc);   // conc[] now holds h_k / RT
  cfloat heat = 0;
#pragma unroll
  for (int k = 0; k < __KINETIX_NSPECIES__; k++) heat = fma(wdot[k], conc[k], heat);
  rates[id] = -p_R * T * heat;
}

__global__ void __launch_bounds__(p_BLOCKSIZE)
transport(long long n_states, long long offsetT, long long offset, dfloat pressure, const dfloat* __restrict__ state,
          dfloat* __restrict__ conductivity, dfloat* __restrict__ viscosity, dfloat* __restrict__ density_diffusivity,
          double Tref)
{
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_states) return;
  const cfloat T = Tref * state[id];
  const cfloat lnT = kx_log(T), rcpT = kx_rcp(T), sqrT = sqrt(T);
  const cfloat lnT2 = lnT * lnT, lnT3 = lnT2 * lnT, lnT4 = lnT2 * lnT2;
  cfloat X[__KINETIX_NSPECIES__], D[__KINETIX_NSPECIES__];
  const cfloat rcpMbar = load_composition(state, id, offsetT, offset, X);
  const cfloat Mbar = kx_rcp(rcpMbar);
#pragma unroll
  for (int k = 0; k < __KINETIX_NSPECIES__; k++) X[k] *= Mbar;   // mole fractions
  conductivity[id] = sqrT * kinetix_conductivity(rcpMbar, lnT, lnT2, lnT3, lnT4, X);
  viscosity[id] = sqrT * kinetix_viscosity(lnT, lnT2, lnT3, lnT4, X);
  kinetix_diffusivity(Mbar, pressure, T * sqrT, lnT, lnT2, lnT3, lnT4, X, D);
  const cfloat rho = pressure / p_R * rcpT * Mbar;
#pragma unroll
  for (int k = 0; k < __KINETIX_NSPECIES__; k++) density_diffusivity[k * offset + id] = rho * D[k];
}

__global__ void __launch_bounds__(p_BLOCKSIZE)
thermoCoeffs(long long n_states, long long offsetT, long long offset, double pressure_R,
             const dfloat* __restrict__ state, dfloat* __restrict__ rho, dfloat* __restrict__ cp,
             dfloat* __restrict__ rhoCp, double Tref)
{
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_states) return;
  const cfloat T = Tref * state[id];
  const cfloat rcpT = kx_rcp(T), T2 = T * T, T3 = T2 * T, T4 = T2 * T2;
  cfloat w[__KINETIX_NSPECIES__], cp_R[__KINETIX_NSPECIES__];
  const cfloat rcpMbar = load_composition(state, id, offsetT, offset, w);
  const cfloat Mbar = kx_rcp(rcpMbar);
  const cfloat density = pressure_R * rcpT * Mbar;
  rho[id] = density;
  kinetix_molar_heat_capacity_R(T, T2, T3, T4, cp_R);
  cfloat mean_cp_R = 0;
#pragma unroll
  for (int k = 0; k < __KINETIX_NSPECIES__; k++) {
    cp[k * offset + id] = cp_R[k] * p_R * kinetix_rcp_molar_mass[k];
    mean_cp_R += cp_R[k] * w[k] * Mbar;
  }
  rhoCp[id] = density * (mean_cp_R * p_R * rcpMbar);
}

unsigned grid_for(long long n) { return (unsigned)((n + p_BLOCKSIZE - 1) / p_BLOCKSIZE); }

}  // namespace

extern "C" {

int kxr_n_species() { return n_species; }
int kxr_n_reactions() { return n_reactions; }

int kxr_production_rates(long long n, long long offsetT, long long offset, double pressure_R, double pressure,
                         const double* state, double* rates, double Tref, cudaStream_t stream)
{
  if (n <= 0) return 0;
  productionRates<<<grid_for(n), p_BLOCKSIZE, 0, stream>>>(n, offsetT, offset, pressure_R, pressure, state, rates, Tref);
  return (int)cudaGetLastError();
}

int kxr_transport(long long n, long long offsetT, long long offset, double pressure, const double* state,
                  double* conductivity, double* viscosity, double* rhoD, double Tref, cudaStream_t stream)
{
  if (n <= 0) return 0;
  transport<<<grid_for(n), p_BLOCKSIZE, 0, stream>>>(n, offsetT, offset, pressure, state, conductivity, viscosity, rhoD,
                                                     Tref);
  return (int)cudaGetLastError();
}

int kxr_thermo(long long n, long long offsetT, long long offset, double pressure_R, const double* state, double* rho,
               double* cp, double* rhoCp, double Tref, cudaStream_t stream)
{
  if (n <= 0) return 0;
  thermoCoeffs<<<grid_for(n), p_BLOCKSIZE, 0, stream>>>(n, offsetT, offset, pressure_R, state, rho, cp, rhoCp, Tref);
  return (int)cudaGetLastError();
}

}  // extern "C"
