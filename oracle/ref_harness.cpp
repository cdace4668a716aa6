// TEST INFRASTRUCTURE ONLY -- serial host wrapper around the REFERENCE's own generated routines.
//
// oracle/build_ref.py runs the unmodified reference generator (/root/reference/kinetix/__main__.py) and
// compiles this file with `-include <gen>/mech.h -include <gen>/rates.cpp ...`, i.e. the arithmetic
// below the wrappers is the reference's generated code, byte for byte.  OCCA and MPI cannot be built in
// this image, so the three per-state OKL kernels are restated here as plain loops, statement by
// statement, with the macro set of the reference's SERIAL backend (benchmark/src/kinetix.cpp:215-252):
//
//   ref_production_rates  <- benchmark/okl/productionRates.okl:10-64   (p_PELETOOL == 0 branch)
//   ref_transport         <- benchmark/okl/transportProps.okl:11-49
//   ref_thermo            <- benchmark/okl/thermoCoeffs.okl:10-40
//
// It mirrors examples/KinetiX/kinetix_test.cpp:110-141,169-198, the reference's own OCCA-free harness.
// Output goes to oracle/_ref/ only.  Nothing in the product links or loads this.
//
// Types: dfloat = storage type of state/result buffers, cfloat = arithmetic type (both set by -D).

#ifndef p_R
#define p_R (1.380649e-23 * 6.02214076e23)
#endif

extern "C" {

int ref_n_species() { return n_species; }
int ref_n_active_species() { return n_active_species; }
int ref_n_reactions() { return n_reactions; }
const char* ref_species_names() { return species_names; }
void ref_molar_masses(double* out) { for (int k = 0; k < n_species; k++) out[k] = kinetix_molar_mass[k]; }

void ref_production_rates(long n_states, long offsetT, long offset, double pressure_R, double pressure_,
                          const dfloat* state, dfloat* rates, double Tref)
{
  for (long id = 0; id < n_states; ++id) {
    const cfloat T = Tref * state[id];
    const cfloat rcpT = 1 / T;
    const cfloat logT = log(T);
    const cfloat T2 = T * T;
    const cfloat T3 = T * T * T;
    const cfloat T4 = T * T * T * T;
    const cfloat P = pressure_;
    const cfloat logP = log(pressure_);

    cfloat wrk1[__KINETIX_NSPECIES__];
    cfloat Mbar;
    {
      cfloat rcpMbar = 0;
      for (int k = 0; k < __KINETIX_NSPECIES__; k++) {
        const cfloat Yi = __KINETIX_MAX((cfloat)0, (cfloat)state[id + offsetT + k * offset]);
        wrk1[k] = Yi * kinetix_rcp_molar_mass[k];
        rcpMbar += wrk1[k];
      }
      Mbar = 1 / rcpMbar;
    }
    {
      cfloat wrk2[__KINETIX_NSPECIES__];
      const cfloat rho = pressure_R * rcpT * Mbar;
      for (int k = 0; k < __KINETIX_NSPECIES__; k++) {
        const cfloat Ci = wrk1[k] * rho;
        wrk1[k] = Ci;
        wrk2[k] = 0;
      }
      kinetix_species_rates(logT, T, T2, T3, T4, rcpT, P, logP, wrk1, wrk2);
      for (int k = 0; k < __KINETIX_NSPECIES__; k++)
        rates[id + offsetT + k * offset] = kinetix_molar_mass[k] * wrk2[k];

      kinetix_enthalpy_RT(T, T2, T3, T4, rcpT, wrk1);
      cfloat sum_h_RT = 0;
      for (int k = 0; k < __KINETIX_NSPECIES__; k++)
        sum_h_RT += wrk2[k] * wrk1[k];
      cfloat ratesFactorEnergy = -p_R * T;
      rates[id] = ratesFactorEnergy * sum_h_RT;
    }
  }
}

#ifndef REF_NO_TRANSPORT
void ref_transport(long n_states, long offsetT, long offset, dfloat pressure, const dfloat* state,
                   dfloat* conductivity, dfloat* viscosity, dfloat* density_diffusivity, double Tref)
{
  for (long id = 0; id < n_states; ++id) {
    const cfloat T = Tref * state[id];
    const cfloat lnT = log(T);
    const cfloat rcpT = 1 / T;
    const cfloat sqrT = sqrt(T);
    const cfloat lnT2 = lnT * lnT;
    const cfloat lnT3 = lnT * lnT * lnT;
    const cfloat lnT4 = lnT * lnT * lnT * lnT;

    cfloat wrk1[__KINETIX_NSPECIES__];
    cfloat wrk2[__KINETIX_NSPECIES__];

    cfloat rcpMbar = 0;
    for (int k = 0; k < __KINETIX_NSPECIES__; k++) {
      const cfloat Yi = __KINETIX_MAX((cfloat)0, (cfloat)state[id + offsetT + k * offset]);
      wrk1[k] = Yi * kinetix_rcp_molar_mass[k];
      rcpMbar += wrk1[k];
    }
    const cfloat Mbar = 1 / rcpMbar;
    for (int k = 0; k < __KINETIX_NSPECIES__; k++)
      wrk1[k] *= Mbar;

    conductivity[id] = sqrT * kinetix_conductivity(rcpMbar, lnT, lnT2, lnT3, lnT4, wrk1);
    viscosity[id] = sqrT * kinetix_viscosity(lnT, lnT2, lnT3, lnT4, wrk1);
    kinetix_diffusivity(Mbar, pressure, T * sqrT, lnT, lnT2, lnT3, lnT4, wrk1, wrk2);

    const cfloat rho = pressure / p_R * rcpT * Mbar;
    for (int k = 0; k < __KINETIX_NSPECIES__; k++)
      density_diffusivity[k * offset + id] = rho * wrk2[k];
  }
}
#endif

void ref_thermo(long n_states, long offsetT, long offset, double pressure_R, const dfloat* state,
                dfloat* rho, dfloat* cp, dfloat* rhoCp, double Tref)
{
  for (long id = 0; id < n_states; ++id) {
    const cfloat T = Tref * state[id];
    const cfloat rcpT = 1 / T;
    const cfloat T2 = T * T;
    const cfloat T3 = T * T * T;
    const cfloat T4 = T * T * T * T;

    cfloat wrk1[__KINETIX_NSPECIES__];
    cfloat rcpMbar = 0;
    for (int k = 0; k < __KINETIX_NSPECIES__; k++) {
      const cfloat Yi = __KINETIX_MAX((cfloat)0, (cfloat)state[id + offsetT + k * offset]);
      wrk1[k] = Yi * kinetix_rcp_molar_mass[k];
      rcpMbar += wrk1[k];
    }
    const cfloat Mbar = 1 / rcpMbar;
    const cfloat rho_ = pressure_R * rcpT * Mbar;
    rho[id] = rho_;

    cfloat cp_R[__KINETIX_NSPECIES__];
    kinetix_molar_heat_capacity_R(T, T2, T3, T4, cp_R);

    cfloat mean_cp_R = 0;
    for (int k = 0; k < __KINETIX_NSPECIES__; k++) {
      cp[k * offset + id] = cp_R[k] * p_R * kinetix_rcp_molar_mass[k];
      mean_cp_R += cp_R[k] * wrk1[k] * Mbar;
    }
    rhoCp[id] = rho_ * (mean_cp_R * p_R * rcpMbar);
  }
}

}  // extern "C"
