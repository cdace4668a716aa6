/* kinetix_b200.h -- C ABI of libkinetix_b200.so: the drop-in boundary for KinetiX's BK1/BK2 hot path.
 *
 * Each entry point replaces one function of the reference's C++ host API
 * (reference benchmark/src/kinetix.hpp:15-107, implemented in benchmark/src/kinetix.cpp).  The
 * reference passes OCCA handles (occa::device, occa::memory) and an MPI communicator; here buffers
 * are raw CUDA device pointers, sizes are 64-bit, every call returns an int status (0 = success) and
 * kx_last_error() describes the last failure.  No OCCA, no MPI, no backend dispatch: the only backend is
 * CUDA sm_100a, and the library fails loudly (non-zero status) if the mechanism's CUDA module cannot
 * be built or loaded -- there is no CPU fallback.
 *
 * Conventions kept from the reference:
 *   - the caller owns every device buffer; the library never allocates state memory;
 *   - addressing: T/T_ref at state[id], mass fraction k at state[id + offsetT + k*offset]; results use
 *     the same layout (rates) or [k*offset + id] (rhoD, cp_i);
 *   - `pressure` arguments are NON-DIMENSIONAL (p / p_ref), exactly as in kinetix.hpp:65-94;
 *   - one mechanism per DEVICE (the reference: per process), launches are asynchronous on the given stream
 *     and the caller synchronises (the reference: default stream + device.finish()); see kx_select_device.
 *
 *   - temperature contract: the generated kernels are valid for T in [200, 6000] K (the emitter proves by interval
 *     arithmetic that every exp() argument stays in range there and folds Troe terms that are exactly 0 or 1 on that
 *     interval, kinetix_b200/core/emit_bk1.py T_VALID_LO/HI); outside it results may saturate differently from the
 *     reference's libm.  The reference's NASA-7 data are fitted for 200/300-3500/6000 K anyway.
 *
 * Calling sequence (same as benchmark/src/bk.cpp:574-760):
 *   kx_init -> getters -> kx_build -> kx_thermodynamic_props / kx_production_rates /
 *   kx_mixture_avg_transport_props ... -> kx_finalize
 */
#ifndef KINETIX_B200_H
#define KINETIX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* storage type of state/result buffers; the reference keys its kernel choice on o_state.dtype()
 * (kinetix.cpp:795-799): FP64 buffers -> FP64 kernel, or "fpmix" (FP64 storage, FP32 math) when
 * single_precision was requested at init; FP32 buffers + single_precision -> pure FP32 kernel. */
#define KX_DTYPE_F64 0
#define KX_DTYPE_F32 1

/* Options of kinetix::init (kinetix.hpp:17-35) minus the OCCA/MPI handles.  Zero-initialise, then set. */
typedef struct kx_options {
  int device_id;          /* CUDA device ordinal (reference: occa::device)                    */
  int block_size;         /* --block-size; 0 = library default (reference default 512)        */
  int single_precision;   /* --single-precision                                               */
  int unroll_loops;       /* --unroll-loops   (accepted; the sm_100a emitter is always specialised) */
  int loop_gibbsexp;      /* --loop-gibbsexp  (accepted; exp(g_k) per species is always used) */
  int group_rxn_unroll;   /* --group-rxnUnroll (accepted; (beta,Ta) sharing is always on)     */
  int group_vis;          /* --group-vis      (accepted, no effect on results)                */
  int nonsym_dij;         /* --nonsymDij      (accepted, no effect on results)                */
  int fit_rcp_diff_coeffs;/* --fit-rcpDiffCoeffs: fit 1/D_jk instead of D_jk (changes results like the reference) */
  int verbose;
  const char* cache_dir;  /* where generated/compiled modules live; NULL = <package>/_cache   */
  const char* tool;       /* "KinetiX" (default). "Pele" is not supported and is rejected.   */
} kx_options;

/* kinetix::init (kinetix.hpp:17-35 / kinetix.cpp:523-599): load the mechanism, generate + compile
 * (or load from cache) its sm_100a module, make the getters valid. */
int kx_init(const char* yaml_path, const kx_options* options);

/* Several GPUs from one process (extension; the reference's state is file-static, one device per process,
 * kinetix.cpp:21-65): every kx_init creates (or replaces) the context of options->device_id and makes it the calling
 * thread's current context; kx_select_device(d) switches the calling thread to the context initialised on device d
 * (error if none).  All other entry points act on the current context and launch on ITS device.  One host thread per
 * GPU may call concurrently, each after its own kx_select_device / kx_init; kx_finalize releases the current
 * context only.  kx_current_device() = device of the current context, -1 if none. */
int kx_select_device(int device_id);
int kx_current_device(void);

/* Generate + compile the mechanism module into the cache if it is not there yet, WITHOUT touching CUDA
 * (safe to call before fork(); a multi-process launcher calls it once, like rank 0 running the generator
 * first in kinetix.cpp:290-296,655-699).  kx_init does this implicitly. */
int kx_prepare(const char* yaml_path, const kx_options* options);

/* Extension hook = the reference's kinetixBuildKernel_t (kinetix.hpp:11-13, installed by init's optional
 * `buildKernel` argument, kinetix.cpp:499-502,542): a host application may supply its own builder for the
 * mechanism module.  When a builder is set, kx_init / kx_prepare call it INSTEAD of running the built-in
 * generator + nvcc whenever the module is not in the cache (or KINETIX_B200_REBUILD is set): it must leave
 * `<output_dir>/libkx_mech.so` (exporting the kxm_* interface) for (yaml_path, options) and return 0; any other
 * return value makes the caller fail with that code in the message.  `user` is passed through.  The builder is a
 * process-wide setting that survives kx_finalize; kx_set_module_builder(NULL, NULL) restores the default. */
typedef int (*kx_build_module_fn)(const char* yaml_path, const kx_options* options, const char* output_dir,
                                  void* user);
int kx_set_module_builder(kx_build_module_fn builder, void* user);

/* kinetix::isInitialized (kinetix.hpp:15).  Like the reference it becomes true after kx_build. */
int kx_is_initialized(void);

/* kinetix::build (kinetix.hpp:58-63 / kinetix.cpp:609-784): store the reference state used to
 * non-dimensionalise T, p and molar masses. ref_mass_fractions has n_species entries. */
int kx_build(double ref_pressure, double ref_temperature, const double* ref_mass_fractions, int transport);

/* kinetix::productionRates (kinetix.hpp:65-72 / kinetix.cpp:786-813).
 * rates[id] = heat release rate [W/m^3], rates[id+offsetT+k*offset] = M_k wdot_k [kg/m^3/s]. */
int kx_production_rates(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                        const void* d_state, void* d_rates, int dtype, void* cuda_stream);

/* kinetix::mixtureAvgTransportProps (kinetix.hpp:74-83 / kinetix.cpp:815-841).  Argument order as in
 * the reference's HOST API: viscosity, then conductivity, then rho*D_km[k*offset + id]. */
int kx_mixture_avg_transport_props(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                                   const void* d_state, void* d_viscosity, void* d_conductivity,
                                   void* d_rho_d, int dtype, void* cuda_stream);

/* kinetix::thermodynamicProps (kinetix.hpp:85-94 / kinetix.cpp:843-871):
 * rho[id], cp_i[k*offset + id] (J/kg/K), rhoCp[id]. */
int kx_thermodynamic_props(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                           const void* d_state, void* d_rho, void* d_cp_i, void* d_rho_cp,
                           int dtype, void* cuda_stream);

/* Extension beyond the reference: one pressure PER STATE instead of one per launch
 * (kinetix.cpp:802-812 dimensionalises a single scalar).  d_pressure[id] = p / p_ref, n_states entries of the
 * same storage type as the state.  With a constant field these compute exactly what kx_production_rates /
 * kx_thermodynamic_props compute with that scalar.  The transport properties of kx_mixture_avg_transport_props
 * do not depend on pressure at all (p cancels in rho * D_km, transportProps.okl:41-46), so it has no such flavour. */
int kx_production_rates_pfield(int64_t n_states, int64_t offsetT, int64_t offset, const void* d_pressure,
                               const void* d_state, void* d_rates, int dtype, void* cuda_stream);
int kx_thermodynamic_props_pfield(int64_t n_states, int64_t offsetT, int64_t offset, const void* d_pressure,
                                  const void* d_state, void* d_rho, void* d_cp_i, void* d_rho_cp,
                                  int dtype, void* cuda_stream);

/* Host-buffer entry points: same arithmetic, HOST pointers in/out (pageable or pinned).  The batch is
 * cut into chunks that are copied in, computed and copied out on alternating streams so that PCIe
 * transfers overlap the kernels.  These are what an application that keeps its fields on the host
 * (e.g. the reference's SERIAL users) calls; bench.py's `e2e` figure times them. */
int kx_production_rates_host(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                             const double* h_state, double* h_rates);
int kx_mixture_avg_transport_props_host(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                                        const double* h_state, double* h_viscosity, double* h_conductivity,
                                        double* h_rho_d);

/* BK1 and BK2 of the same host-resident states with a single upload of the state slab (what one CFD
 * time step needs).  Not in the reference API: its callers issue productionRates and
 * mixtureAvgTransportProps back to back on the same o_state (bk.cpp:697-760). */
int kx_rates_and_transport_host(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                                const double* h_state, double* h_rates, double* h_viscosity,
                                double* h_conductivity, double* h_rho_d);

/* getters (kinetix.hpp:96-107 / kinetix.cpp:873-908) */
int kx_n_species(void);
int kx_n_active_species(void);
int kx_n_reactions(void);
const char* kx_species_name(int k);              /* NULL if out of range                        */
int kx_species_index(const char* name);          /* -1 if absent                                */
int kx_molecular_weights(double* out);           /* M_k / Mbar_ref, n_species entries (needs kx_build) */
int kx_molar_masses(double* out);                /* M_k in kg/mol (mech.h's kinetix_molar_mass) */
double kx_ref_pressure(void);
double kx_ref_temperature(void);
int kx_ref_mass_fractions(double* out);
double kx_ref_mean_molecular_weight(void);

/* FNV-1a-64 step over a byte string: the hash of the `.inputs` stamp the generator writes beside a cached module and
 * kx_init re-checks (exported so that the Python generator and this library agree by construction). */
uint64_t kx_fnv1a64(uint64_t h, const void* data, size_t n);

/* diagnostics */
const char* kx_last_error(void);
const char* kx_module_path(void);                /* the compiled mechanism module in use        */
int kx_finalize(void);                           /* unload the module, reset all state          */

#ifdef __cplusplus
}
#endif
#endif /* KINETIX_B200_H */
