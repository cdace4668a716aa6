// kx_host.cpp -- libkinetix_b200.so: thin C-ABI host for the BK1/BK2 hot path (include/kinetix_b200.h).
//
// Replaces the reference's host library (reference benchmark/src/kinetix.cpp, 908 lines on top of OCCA +
// MPI): same life-cycle (init -> build -> launches -> getters), same non-dimensional pressure
// convention, same "run the Python generator through system(), cache by option hash" idea
// (kinetix.cpp:336-346, 655-699) -- but the generated artefact is a CUDA sm_100a shared object that is
// dlopen()ed and launched directly through the CUDA runtime.  No OCCA, no MPI, no backend dispatch, and
// NO CPU fallback: if the module cannot be produced or loaded every entry point fails with a message.
#include "kinetix_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

namespace {

const double R_GAS = 1.380649e-23 * 6.02214076e23;   // kinetix.cpp:37

typedef int (*fn_int_t)();
typedef const char* (*fn_str_t)();
typedef void (*fn_masses_t)(double*);
typedef int (*fn_rates_t)(long long, long long, long long, double, double, const void*, void*, double, const void*,
                          int, cudaStream_t);
typedef int (*fn_transport_t)(long long, long long, long long, double, const void*, void*, void*, void*, double,
                              int, cudaStream_t);
typedef int (*fn_thermo_t)(long long, long long, long long, double, const void*, void*, void*, void*, double,
                           const void*, int, cudaStream_t);

struct State {
  void* module = nullptr;
  std::string module_path;
  fn_rates_t rates = nullptr;
  fn_transport_t transport = nullptr;
  fn_thermo_t thermo = nullptr;
  int n_species = -1, n_active = -1, n_reactions = -1;
  std::vector<std::string> names;
  std::vector<double> m_molar;
  bool built = false;
  bool single_precision = false;
  int device_id = 0;
  double ref_pressure = 0, ref_temperature = 0, ref_mean_molar_mass = 0;
  std::vector<double> ref_mass_fractions;
  // staging for the host-buffer entry points
  static const int SLOTS = 4;
  cudaStream_t streams[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  void* d_in[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  void* d_out[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  size_t in_bytes = 0, out_bytes = 0;
} g;

std::string g_error;

int fail(const std::string& msg, int code = 1)
{
  g_error = msg;
  if (getenv("KINETIX_B200_VERBOSE")) fprintf(stderr, "[kinetix_b200] error: %s\n", msg.c_str());
  return code;
}

int cuda_fail(const char* what, int err)
{
  std::ostringstream s;
  s << what << ": CUDA error " << err << " (" << cudaGetErrorString((cudaError_t)err) << ")";
  return fail(s.str(), err ? err : 1);
}

bool exists(const std::string& p)
{
  struct stat st;
  return stat(p.c_str(), &st) == 0;
}

// directory that holds this shared object (= the kinetix_b200 python package)
std::string package_dir()
{
  Dl_info info;
  if (dladdr((void*)&package_dir, &info) && info.dli_fname) {
    std::string p(info.dli_fname);
    size_t pos = p.find_last_of('/');
    return pos == std::string::npos ? "." : p.substr(0, pos);
  }
  return ".";
}

std::string stem_of(const std::string& path)
{
  size_t a = path.find_last_of('/');
  std::string base = a == std::string::npos ? path : path.substr(a + 1);
  size_t b = base.find_last_of('.');
  return b == std::string::npos ? base : base.substr(0, b);
}

void unload()
{
  for (int s = 0; s < State::SLOTS; s++) {
    if (g.d_in[s]) cudaFree(g.d_in[s]);
    if (g.d_out[s]) cudaFree(g.d_out[s]);
    if (g.streams[s]) cudaStreamDestroy(g.streams[s]);
  }
  if (g.module) dlclose(g.module);
  g = State();
}

template <class F>
bool resolve(F& f, const char* sym)
{
  f = (F)dlsym(g.module, sym);
  return f != nullptr;
}

// host-supplied module builder (kx_set_module_builder); process-wide like the reference's static `buildKernel`
kx_build_module_fn g_builder = nullptr;
void* g_builder_user = nullptr;

// Locate the compiled module of (mechanism, options) in the cache, generating + compiling it first if it
// is missing (cf. kinetix.cpp:655-699: generator through system(), cached by option hash).  No CUDA calls.
int prepare_module(const char* yaml_path, const kx_options& opt, std::string& lib)
{
  const std::string pkg = package_dir();
  std::string cache = opt.cache_dir ? opt.cache_dir : (getenv("KINETIX_B200_CACHE") ? getenv("KINETIX_B200_CACHE")
                                                                                     : pkg + "/_cache");
  std::string tag = stem_of(yaml_path);
  if (opt.fit_rcp_diff_coeffs) tag += "-rcpdiff";
  if (opt.single_precision) tag += "-sp";
  if (opt.block_size > 0) tag += "-b" + std::to_string(opt.block_size);
  const std::string dir = cache + "/" + tag;
  lib = dir + "/libkx_mech.so";

  if ((!exists(lib) || getenv("KINETIX_B200_REBUILD")) && g_builder) {
    if (opt.verbose) fprintf(stderr, "[kinetix_b200] module builder hook -> %s\n", dir.c_str());
    const int rc = g_builder(yaml_path, &opt, dir.c_str(), g_builder_user);
    if (rc != 0) return fail("kx_init: the module builder hook returned " + std::to_string(rc) + " for " + dir);
    if (!exists(lib)) return fail("kx_init: the module builder hook did not produce " + lib);
  } else if (!exists(lib) || getenv("KINETIX_B200_REBUILD")) {
    const char* py = getenv("KINETIX_B200_PYTHON") ? getenv("KINETIX_B200_PYTHON") : "python3";
    std::string parent = pkg.substr(0, pkg.find_last_of('/'));
    std::ostringstream cmd;
    cmd << "PYTHONPATH='" << parent << "':\"$PYTHONPATH\" " << py << " -m kinetix_b200"
        << " --mechanism '" << yaml_path << "' --output '" << dir << "' --target sm_100a --compile";
    if (opt.single_precision) cmd << " --single-precision";
    if (opt.unroll_loops) cmd << " --unroll-loops";
    if (opt.loop_gibbsexp) cmd << " --loop-gibbsexp";
    if (opt.group_rxn_unroll) cmd << " --group-rxnunroll";
    if (opt.group_vis) cmd << " --group-vis";
    if (opt.nonsym_dij) cmd << " --nonsymDij";
    if (opt.fit_rcp_diff_coeffs) cmd << " --fit-rcpdiffcoeffs";
    if (opt.block_size > 0) cmd << " --block-size " << opt.block_size;
    if (opt.verbose) fprintf(stderr, "[kinetix_b200] %s\n", cmd.str().c_str());
    if (system(cmd.str().c_str()) != 0 || !exists(lib))
      return fail("kx_init: error while running the code generator / nvcc: " + cmd.str());
  }
  return 0;
}

}  // namespace

extern "C" {

const char* kx_last_error(void) { return g_error.c_str(); }
const char* kx_module_path(void) { return g.module_path.c_str(); }

int kx_finalize(void)
{
  unload();
  return 0;
}

int kx_init(const char* yaml_path, const kx_options* opt_in)
{
  unload();
  if (!yaml_path) return fail("kx_init: yaml_path is NULL");
  kx_options opt;
  memset(&opt, 0, sizeof(opt));
  if (opt_in) opt = *opt_in;
  if (opt.tool && strcmp(opt.tool, "KinetiX") != 0)
    return fail(std::string("kx_init: tool '") + opt.tool + "' is not supported (only KinetiX routines)");
  if (!exists(yaml_path)) return fail(std::string("kx_init: mechanism file not found: ") + yaml_path);

  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail("kx_init: no CUDA device available -- kinetix_b200 has no CPU path (use the reference's SERIAL "
                "backend for that)");
  if (opt.device_id < 0 || opt.device_id >= ndev) return fail("kx_init: device_id out of range");
  if ((ce = cudaSetDevice(opt.device_id)) != cudaSuccess) return cuda_fail("cudaSetDevice", ce);
  g.device_id = opt.device_id;
  g.single_precision = opt.single_precision != 0;

  std::string lib;
  if (int e = prepare_module(yaml_path, opt, lib)) return e;

  g.module = dlopen(lib.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!g.module) return fail(std::string("kx_init: dlopen failed: ") + dlerror());
  g.module_path = lib;

  fn_int_t abi = nullptr, nsp = nullptr, nact = nullptr, nrx = nullptr;
  fn_str_t names = nullptr;
  fn_masses_t masses = nullptr;
  if (!resolve(abi, "kxm_abi_version") || !resolve(nsp, "kxm_n_species") ||
      !resolve(nact, "kxm_n_active_species") || !resolve(nrx, "kxm_n_reactions") ||
      !resolve(names, "kxm_species_names") || !resolve(masses, "kxm_molar_masses") ||
      !resolve(g.rates, "kxm_production_rates") || !resolve(g.transport, "kxm_transport") ||
      !resolve(g.thermo, "kxm_thermo")) {
    unload();
    return fail("kx_init: " + lib + " does not export the kxm_* module interface");
  }
  if (abi() != 2) {
    unload();
    return fail("kx_init: module ABI version mismatch, remove the cached module " + lib);
  }
  fn_int_t msp = nullptr;
  if (!resolve(msp, "kxm_single_precision") || (msp() != 0) != g.single_precision) {
    unload();
    return fail("kx_init: cached module " + lib + " was generated for a different precision");
  }
  // equivalent of the reference's mech.okl query kernels (kinetix.cpp:352-403)
  g.n_species = nsp();
  g.n_active = nact();
  g.n_reactions = nrx();
  g.m_molar.resize(g.n_species);
  masses(g.m_molar.data());
  std::istringstream is(names());
  std::string tok;
  while (is >> tok) g.names.push_back(tok);
  if ((int)g.names.size() != g.n_species) {
    unload();
    return fail("kx_init: species name table is inconsistent");
  }
  g_error.clear();
  return 0;
}

int kx_prepare(const char* yaml_path, const kx_options* opt_in)
{
  if (!yaml_path) return fail("kx_prepare: yaml_path is NULL");
  kx_options opt;
  memset(&opt, 0, sizeof(opt));
  if (opt_in) opt = *opt_in;
  if (!exists(yaml_path)) return fail(std::string("kx_prepare: mechanism file not found: ") + yaml_path);
  std::string lib;
  return prepare_module(yaml_path, opt, lib);
}

int kx_set_module_builder(kx_build_module_fn builder, void* user)
{
  g_builder = builder;
  g_builder_user = builder ? user : nullptr;
  return 0;
}

int kx_is_initialized(void) { return g.built ? 1 : 0; }

int kx_build(double ref_pressure, double ref_temperature, const double* ref_mass_fractions, int /*transport*/)
{
  if (!g.module) return fail("kx_build: call kx_init first");
  if (!ref_mass_fractions) return fail("kx_build: ref_mass_fractions is NULL");
  g.ref_pressure = ref_pressure;
  g.ref_temperature = ref_temperature;
  g.ref_mass_fractions.assign(ref_mass_fractions, ref_mass_fractions + g.n_species);
  double sum = 0.;
  for (int k = 0; k < g.n_species; k++) sum += ref_mass_fractions[k] / g.m_molar[k];   // kinetix.cpp:625-630
  g.ref_mean_molar_mass = 1. / sum;
  g.built = true;
  return 0;
}

// kernel flavour = f(single_precision at init, storage type), as in the reference (kinetix.cpp:795-799):
// FP64 module serves FP64 buffers; the --single-precision module serves FP64 buffers ("fpmix") and FP32.
static int check_dtype(const char* who, int dtype)
{
  if (dtype == KX_DTYPE_F64) return 0;
  if (dtype == KX_DTYPE_F32 && g.single_precision) return 0;
  return fail(std::string(who) + (dtype == KX_DTYPE_F32
                                      ? ": FP32 buffers need kx_init with single_precision = 1"
                                      : ": unknown dtype"));
}

#define KX_REQUIRE_BUILT(name) \
  if (!g.built) return fail(name ": kx_init/kx_build have not been called")

int kx_production_rates(int64_t n_states, int64_t offsetT, int64_t offset, double pressure, const void* d_state,
                        void* d_rates, int dtype, void* stream)
{
  KX_REQUIRE_BUILT("kx_production_rates");
  if (n_states < 0) return fail("kx_production_rates: negative n_states");
  if (n_states && (!d_state || !d_rates)) return fail("kx_production_rates: NULL buffer");
  if (int e = check_dtype("kx_production_rates", dtype)) return e;
  const double pressure_ = pressure * g.ref_pressure;          // kinetix.cpp:802-803
  const double pressure_R = pressure_ / R_GAS;
  int e = g.rates(n_states, offsetT, offset, pressure_R, pressure_, d_state, d_rates, g.ref_temperature, nullptr,
                  dtype, (cudaStream_t)stream);
  return e ? cuda_fail("kx_production_rates", e) : 0;
}

// Extension: one pressure PER STATE.  d_pressure[id] = p / p_ref, same storage type as the
// state.  The reference has a single pressure per launch (kinetix.cpp:802-812); with a constant field this call
// computes exactly what kx_production_rates computes.
int kx_production_rates_pfield(int64_t n_states, int64_t offsetT, int64_t offset, const void* d_pressure,
                               const void* d_state, void* d_rates, int dtype, void* stream)
{
  KX_REQUIRE_BUILT("kx_production_rates_pfield");
  if (n_states < 0) return fail("kx_production_rates_pfield: negative n_states");
  if (n_states && (!d_state || !d_rates || !d_pressure)) return fail("kx_production_rates_pfield: NULL buffer");
  if (int e = check_dtype("kx_production_rates_pfield", dtype)) return e;
  int e = g.rates(n_states, offsetT, offset, g.ref_pressure / R_GAS, g.ref_pressure, d_state, d_rates,
                  g.ref_temperature, d_pressure, dtype, (cudaStream_t)stream);
  return e ? cuda_fail("kx_production_rates_pfield", e) : 0;
}

int kx_mixture_avg_transport_props(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                                   const void* d_state, void* d_viscosity, void* d_conductivity, void* d_rho_d,
                                   int dtype, void* stream)
{
  KX_REQUIRE_BUILT("kx_mixture_avg_transport_props");
  if (n_states < 0) return fail("kx_mixture_avg_transport_props: negative n_states");
  if (n_states && (!d_state || !d_viscosity || !d_conductivity || !d_rho_d))
    return fail("kx_mixture_avg_transport_props: NULL buffer");
  if (int e = check_dtype("kx_mixture_avg_transport_props", dtype)) return e;
  // the reference passes the non-dimensional pressure straight through (kinetix.cpp:832-840); note the
  // kernel argument order conductivity, viscosity
  int e = g.transport(n_states, offsetT, offset, pressure, d_state, d_conductivity, d_viscosity, d_rho_d,
                      g.ref_temperature, dtype, (cudaStream_t)stream);
  if (e == 1001) return fail("kx_mixture_avg_transport_props: module was generated without transport");
  return e ? cuda_fail("kx_mixture_avg_transport_props", e) : 0;
}

int kx_thermodynamic_props(int64_t n_states, int64_t offsetT, int64_t offset, double pressure, const void* d_state,
                           void* d_rho, void* d_cp_i, void* d_rho_cp, int dtype, void* stream)
{
  KX_REQUIRE_BUILT("kx_thermodynamic_props");
  if (n_states < 0) return fail("kx_thermodynamic_props: negative n_states");
  if (n_states && (!d_state || !d_rho || !d_cp_i || !d_rho_cp)) return fail("kx_thermodynamic_props: NULL buffer");
  if (int e = check_dtype("kx_thermodynamic_props", dtype)) return e;
  const double pressure_R = pressure * g.ref_pressure / R_GAS;   // kinetix.cpp:858
  int e = g.thermo(n_states, offsetT, offset, pressure_R, d_state, d_rho, d_cp_i, d_rho_cp, g.ref_temperature,
                   nullptr, dtype, (cudaStream_t)stream);
  return e ? cuda_fail("kx_thermodynamic_props", e) : 0;
}

// per-state pressure flavour of kx_thermodynamic_props (rho = p M / (R T) is the only pressure-dependent output)
int kx_thermodynamic_props_pfield(int64_t n_states, int64_t offsetT, int64_t offset, const void* d_pressure,
                                  const void* d_state, void* d_rho, void* d_cp_i, void* d_rho_cp, int dtype,
                                  void* stream)
{
  KX_REQUIRE_BUILT("kx_thermodynamic_props_pfield");
  if (n_states < 0) return fail("kx_thermodynamic_props_pfield: negative n_states");
  if (n_states && (!d_state || !d_rho || !d_cp_i || !d_rho_cp || !d_pressure))
    return fail("kx_thermodynamic_props_pfield: NULL buffer");
  if (int e = check_dtype("kx_thermodynamic_props_pfield", dtype)) return e;
  int e = g.thermo(n_states, offsetT, offset, g.ref_pressure / R_GAS, d_state, d_rho, d_cp_i, d_rho_cp,
                   g.ref_temperature, d_pressure, dtype, (cudaStream_t)stream);
  return e ? cuda_fail("kx_thermodynamic_props_pfield", e) : 0;
}

// ---- host-buffer entry points ---------------------------------------------------------------------
namespace {

// states per pipelined chunk (SLOTS chunks in flight: H2D | kernels | D2H).  The copies dominate (PCIe), so the
// chunk only has to be large enough for full-rate DMA (tens of MB) and small enough that the pipeline's fill and
// drain (one chunk up, one chunk down) stay a small fraction of a call: 128 Ki states = 55 MB of GRI-3.0 state.
// KX_HOST_CHUNK overrides (development).
int64_t host_chunk()
{
  static int64_t chunk = 0;
  if (!chunk) {
    const char* e = getenv("KX_HOST_CHUNK");
    chunk = e ? atoll(e) : (int64_t)1 << 17;
    if (chunk < 1024) chunk = 1024;
  }
  return chunk;
}
#define CHUNK host_chunk()

int ensure_staging(size_t in_bytes, size_t out_bytes)
{
  for (int s = 0; s < State::SLOTS; s++) {
    if (!g.streams[s]) {
      cudaError_t e = cudaStreamCreateWithFlags(&g.streams[s], cudaStreamNonBlocking);
      if (e != cudaSuccess) return cuda_fail("cudaStreamCreate", e);
    }
  }
  if (in_bytes > g.in_bytes) {
    for (int s = 0; s < State::SLOTS; s++) {
      if (g.d_in[s]) cudaFree(g.d_in[s]);
      cudaError_t e = cudaMalloc(&g.d_in[s], in_bytes);
      if (e != cudaSuccess) return cuda_fail("cudaMalloc(staging in)", e);
    }
    g.in_bytes = in_bytes;
  }
  if (out_bytes > g.out_bytes) {
    for (int s = 0; s < State::SLOTS; s++) {
      if (g.d_out[s]) cudaFree(g.d_out[s]);
      cudaError_t e = cudaMalloc(&g.d_out[s], out_bytes);
      if (e != cudaSuccess) return cuda_fail("cudaMalloc(staging out)", e);
    }
    g.out_bytes = out_bytes;
  }
  return 0;
}

// error exit of a pipelined call: wait for the chunks already in flight (they read and write the CALLER's host
// buffers, which the caller may release as soon as the call has returned), then pass the status on
int drain(int status)
{
  for (int s = 0; s < State::SLOTS; s++)
    if (g.streams[s]) cudaStreamSynchronize(g.streams[s]);
  return status;
}

// copy rows [T; Y_0..Y_{N-1}] of `len` states starting at state s0 into a dense (N+1) x len device slab
cudaError_t upload_chunk(const double* h_state, int64_t s0, int64_t len, int64_t offsetT, int64_t offset,
                         double* d, cudaStream_t st)
{
  const int N = g.n_species;
  cudaError_t e = cudaMemcpyAsync(d, h_state + s0, len * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  return cudaMemcpy2DAsync(d + len, len * sizeof(double), h_state + s0 + offsetT, offset * sizeof(double),
                           len * sizeof(double), N, cudaMemcpyHostToDevice, st);
}

}  // namespace

int kx_production_rates_host(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                             const double* h_state, double* h_rates)
{
  KX_REQUIRE_BUILT("kx_production_rates_host");
  if (n_states <= 0) return n_states == 0 ? 0 : fail("kx_production_rates_host: negative n_states");
  if (!h_state || !h_rates) return fail("kx_production_rates_host: NULL buffer");
  const int N = g.n_species;
  const int64_t chunk = std::min<int64_t>(CHUNK, n_states);
  const size_t slab = (size_t)(N + 1) * chunk * sizeof(double);
  if (int e = ensure_staging(slab, slab)) return e;
  int slot = 0;
  for (int64_t s0 = 0; s0 < n_states; s0 += chunk, slot = (slot + 1) % State::SLOTS) {
    const int64_t len = std::min<int64_t>(chunk, n_states - s0);
    cudaStream_t st = g.streams[slot];
    double* din = (double*)g.d_in[slot];
    double* dout = (double*)g.d_out[slot];
    cudaError_t e = upload_chunk(h_state, s0, len, offsetT, offset, din, st);
    if (e != cudaSuccess) return drain(cuda_fail("kx_production_rates_host: H2D", e));
    if (int r = kx_production_rates(len, len, len, pressure, din, dout, KX_DTYPE_F64, st)) return drain(r);
    e = cudaMemcpyAsync(h_rates + s0, dout, len * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(h_rates + s0 + offsetT, offset * sizeof(double), dout + len, len * sizeof(double),
                            len * sizeof(double), N, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return drain(cuda_fail("kx_production_rates_host: D2H", e));
  }
  for (int s = 0; s < State::SLOTS; s++) {
    cudaError_t e = cudaStreamSynchronize(g.streams[s]);
    if (e != cudaSuccess) return cuda_fail("kx_production_rates_host: sync", e);
  }
  return 0;
}

int kx_mixture_avg_transport_props_host(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                                        const double* h_state, double* h_viscosity, double* h_conductivity,
                                        double* h_rho_d)
{
  KX_REQUIRE_BUILT("kx_mixture_avg_transport_props_host");
  if (n_states <= 0) return n_states == 0 ? 0 : fail("kx_mixture_avg_transport_props_host: negative n_states");
  if (!h_state || !h_viscosity || !h_conductivity || !h_rho_d)
    return fail("kx_mixture_avg_transport_props_host: NULL buffer");
  const int N = g.n_species;
  const int64_t chunk = std::min<int64_t>(CHUNK, n_states);
  if (int e = ensure_staging((size_t)(N + 1) * chunk * sizeof(double), (size_t)(N + 2) * chunk * sizeof(double)))
    return e;
  int slot = 0;
  for (int64_t s0 = 0; s0 < n_states; s0 += chunk, slot = (slot + 1) % State::SLOTS) {
    const int64_t len = std::min<int64_t>(chunk, n_states - s0);
    cudaStream_t st = g.streams[slot];
    double* din = (double*)g.d_in[slot];
    double* dout = (double*)g.d_out[slot];   // [viscosity | conductivity | rhoD rows]
    cudaError_t e = upload_chunk(h_state, s0, len, offsetT, offset, din, st);
    if (e != cudaSuccess) return drain(cuda_fail("kx_mixture_avg_transport_props_host: H2D", e));
    if (int r = kx_mixture_avg_transport_props(len, len, len, pressure, din, dout, dout + len, dout + 2 * len,
                                               KX_DTYPE_F64, st))
      return drain(r);
    e = cudaMemcpyAsync(h_viscosity + s0, dout, len * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(h_conductivity + s0, dout + len, len * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(h_rho_d + s0, offset * sizeof(double), dout + 2 * len, len * sizeof(double),
                            len * sizeof(double), N, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return drain(cuda_fail("kx_mixture_avg_transport_props_host: D2H", e));
  }
  for (int s = 0; s < State::SLOTS; s++) {
    cudaError_t e = cudaStreamSynchronize(g.streams[s]);
    if (e != cudaSuccess) return cuda_fail("kx_mixture_avg_transport_props_host: sync", e);
  }
  return 0;
}

// BK1 + BK2 for the same host-resident states with ONE upload of the state slab: what a
// CFD time step needs (rates and transport coefficients of the same field).
int kx_rates_and_transport_host(int64_t n_states, int64_t offsetT, int64_t offset, double pressure,
                                const double* h_state, double* h_rates, double* h_viscosity,
                                double* h_conductivity, double* h_rho_d)
{
  KX_REQUIRE_BUILT("kx_rates_and_transport_host");
  if (n_states <= 0) return n_states == 0 ? 0 : fail("kx_rates_and_transport_host: negative n_states");
  if (!h_state || !h_rates || !h_viscosity || !h_conductivity || !h_rho_d)
    return fail("kx_rates_and_transport_host: NULL buffer");
  const int N = g.n_species;
  const int64_t chunk = std::min<int64_t>(CHUNK, n_states);
  if (int e = ensure_staging((size_t)(N + 1) * chunk * sizeof(double), (size_t)(2 * N + 3) * chunk * sizeof(double)))
    return e;
  int slot = 0;
  for (int64_t s0 = 0; s0 < n_states; s0 += chunk, slot = (slot + 1) % State::SLOTS) {
    const int64_t len = std::min<int64_t>(chunk, n_states - s0);
    cudaStream_t st = g.streams[slot];
    double* din = (double*)g.d_in[slot];
    double* drates = (double*)g.d_out[slot];                 // (N+1) x len
    double* dtr = drates + (size_t)(N + 1) * len;            // [viscosity | conductivity | rhoD rows]
    cudaError_t e = upload_chunk(h_state, s0, len, offsetT, offset, din, st);
    if (e != cudaSuccess) return drain(cuda_fail("kx_rates_and_transport_host: H2D", e));
    if (int r = kx_production_rates(len, len, len, pressure, din, drates, KX_DTYPE_F64, st)) return drain(r);
    if (int r = kx_mixture_avg_transport_props(len, len, len, pressure, din, dtr, dtr + len, dtr + 2 * len,
                                               KX_DTYPE_F64, st))
      return drain(r);
    e = cudaMemcpyAsync(h_rates + s0, drates, len * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(h_rates + s0 + offsetT, offset * sizeof(double), drates + len, len * sizeof(double),
                            len * sizeof(double), N, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(h_viscosity + s0, dtr, len * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(h_conductivity + s0, dtr + len, len * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(h_rho_d + s0, offset * sizeof(double), dtr + 2 * len, len * sizeof(double),
                            len * sizeof(double), N, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return drain(cuda_fail("kx_rates_and_transport_host: D2H", e));
  }
  for (int s = 0; s < State::SLOTS; s++) {
    cudaError_t e = cudaStreamSynchronize(g.streams[s]);
    if (e != cudaSuccess) return cuda_fail("kx_rates_and_transport_host: sync", e);
  }
  return 0;
}

// ---- getters (kinetix.cpp:873-908) ----------------------------------------------------------------
int kx_n_species(void) { return g.n_species; }
int kx_n_active_species(void) { return g.n_active; }
int kx_n_reactions(void) { return g.n_reactions; }

const char* kx_species_name(int k)
{
  if (k < 0 || k >= (int)g.names.size()) return nullptr;
  return g.names[k].c_str();
}

int kx_species_index(const char* name)
{
  if (!name) return -1;
  for (size_t k = 0; k < g.names.size(); k++)
    if (g.names[k] == name) return (int)k;
  return -1;
}

int kx_molar_masses(double* out)
{
  if (!g.module) return fail("kx_molar_masses: call kx_init first");
  std::copy(g.m_molar.begin(), g.m_molar.end(), out);
  return 0;
}

int kx_molecular_weights(double* out)
{
  KX_REQUIRE_BUILT("kx_molecular_weights");
  for (int k = 0; k < g.n_species; k++) out[k] = g.m_molar[k] / g.ref_mean_molar_mass;   // kinetix.cpp:888-895
  return 0;
}

double kx_ref_pressure(void) { return g.ref_pressure; }
double kx_ref_temperature(void) { return g.ref_temperature; }
double kx_ref_mean_molecular_weight(void) { return g.ref_mean_molar_mass; }

int kx_ref_mass_fractions(double* out)
{
  KX_REQUIRE_BUILT("kx_ref_mass_fractions");
  std::copy(g.ref_mass_fractions.begin(), g.ref_mass_fractions.end(), out);
  return 0;
}

}  // extern "C"
