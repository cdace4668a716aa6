// kinetix_bk -- benchmark / self-check driver of the B200-native KinetiX hot path.
//
// Command line, output lines and exit status follow the reference driver (reference
// benchmark/src/bk.cpp:376-781) so it can be dropped into the same scripts:
//
//   kinetix_bk --backend CUDA --yaml-file gri30.yaml --mode 1|2|0 --n-states N [--n-repetitions R]
//              [--cimode 1|2] [--debug] [--device-id i] [--block-size b] [--single-precision]
//              [--unroll-loops] [--loop-gibbsexp] [--group-rxnUnroll] [--group-vis] [--nonsymDij]
//              [--fit-rcpDiffCoeffs] [--tool KinetiX]
//   additions: --gpus G          G worker processes, one per GPU, each owning n-states/G states (what
//                                `mpirun -np G` does for the reference; there is no MPI here)
//              --random-states   seeded synthetic states (T ~ U[300,2500] K, normalised random Y, 1 atm)
//                                instead of the reference's identical states (bk.cpp:612-615)
//
// Differences: no OCCA/MPI; `--backend` must be CUDA (or B200) -- SERIAL/HIP/DPCPP are the reference's business;
// besides GRXN/s and GDOF/s the driver prints states/s and the fraction of the measured FP64 roofline.
#include <getopt.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "kinetix_b200.h"

namespace {

bool debug_flag = false;
int n_states = 100000, n_species = 0, n_reactions = 0, n_active_species = 0;

struct CiData {
  std::vector<double> rho, cp_mean, hrr, conductivity, viscosity;
  std::vector<std::vector<double>> cp_i, rates, rhoD;
} ci;

#define CUDA_OK(x)                                                                          \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) {                                                                \
      fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));                       \
      exit(EXIT_FAILURE);                                                                   \
    }                                                                                       \
  } while (0)

void kx_ok(int status, const char* what)
{
  if (status) {
    fprintf(stderr, "%s failed: %s\n", what, kx_last_error());
    exit(EXIT_FAILURE);
  }
}

double rel_err(double a, double b)
{
  if (std::isnan(a) || std::isinf(a)) printf("Detected invalid value: %e \n", a);
  return std::abs((a - b) / b);
}

std::vector<std::string> split(const std::string& s, char delim)
{
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string tok;
  while (std::getline(ss, tok, delim)) out.push_back(tok);
  return out;
}

std::vector<double> numbers(const std::string& line, int n)
{
  std::vector<double> v;
  std::stringstream ss(line);
  double x;
  while ((int)v.size() < n && ss >> x) v.push_back(x);
  return v;
}

std::string ci_dir()
{
  if (const char* p = getenv("KINETIX_CI_DATA")) return p;
  if (const char* p = getenv("KINETIX_PATH")) return std::string(p) + "/kinetix/ci_data";
  return "tests/golden/ci_data";
}

// 13-line Cantera known-answer file (reference bk.cpp:271-374)
void load_ci_state(const std::string& state, const std::string& mech, double& p, double& T,
                   std::vector<double>& X, std::vector<double>& M)
{
  const std::string path = ci_dir() + "/" + mech + "." + state + ".cantera";
  std::cout << "Reading ci data from " << path;
  std::ifstream f(path);
  std::vector<std::string> lines;
  for (std::string l; std::getline(f, l);) lines.push_back(l);
  if (lines.size() < 13) {
    std::cout << "\ndata file does not exist or is corrupt!" << std::endl;
    exit(EXIT_FAILURE);
  }
  M = numbers(lines[1], n_species);
  if ((int)M.size() != n_species) {
    std::cerr << "\nNumber of species does not match!\n";
    exit(EXIT_FAILURE);
  }
  std::vector<double> Mk(n_species);
  kx_molar_masses(Mk.data());
  for (int k = 0; k < n_species; k++) {   // species-order guard
    const double e = std::abs(M[k] - Mk[k]) / Mk[k];
    if (e > 1e-7) {
      printf("\nmolar mass mismatch [%d]: w %e cantera %e relative error %e\n", k, Mk[k], M[k], e);
      exit(EXIT_FAILURE);
    }
  }
  T = std::stod(lines[2]);
  p = std::stod(lines[3]);
  X = numbers(lines[4], n_species);
  double sum = 0;
  for (double x : X) sum += x;
  for (double& x : X) x /= sum;
  double Mbar = 0;
  for (int k = 0; k < n_species; k++) Mbar += X[k] * M[k];
  ci.rho.push_back(std::stod(lines[5]));
  ci.cp_mean.push_back(std::stod(lines[6]) / Mbar);
  auto cp = numbers(lines[7], n_species);
  for (int k = 0; k < n_species; k++) cp[k] /= M[k];
  ci.cp_i.push_back(cp);
  ci.rates.push_back(numbers(lines[8], n_species));
  ci.hrr.push_back(std::stod(lines[9]));
  ci.conductivity.push_back(std::stod(lines[10]));
  ci.viscosity.push_back(std::stod(lines[11]));
  ci.rhoD.push_back(numbers(lines[12], n_species));
  std::cout << " ... done" << std::endl;
}

template <class T>
std::vector<T> download(const T* d, size_t n)
{
  std::vector<T> h(n);
  CUDA_OK(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost));
  return h;
}

bool check_thermo(const double* d_rho, const double* d_cp, const double* d_rhoCp)   // bk.cpp:89-157
{
  auto rho = download(d_rho, n_states), cp = download(d_cp, (size_t)n_states * n_species),
       rhoCp = download(d_rhoCp, n_states);
  bool all = true;
  for (int id = 0; id < n_states; id++) {
    double e = std::max(rel_err(rho[id], ci.rho[id]), rel_err(rhoCp[id], ci.rho[id] * ci.cp_mean[id]));
    for (int k = 0; k < n_species; k++) e = std::max(e, rel_err(cp[(size_t)k * n_states + id], ci.cp_i[id][k]));
    const double rtol = 5e-7;
    const bool ok = e < rtol;
    all &= ok;
    printf("thermoCoeffs error_inf: %e < %e (%s)\n", e, rtol, ok ? "passed" : "failed");
  }
  return all;
}

bool check_rates(const double* d_rates, bool single_precision)   // bk.cpp:159-209
{
  auto rates = download(d_rates, (size_t)n_states * (n_species + 1));
  std::vector<double> mw(n_species);
  kx_molecular_weights(mw.data());
  const double Mref = kx_ref_mean_molecular_weight();
  bool all = true;
  for (int id = 0; id < n_states; id++) {
    double e = rel_err(rates[id], ci.hrr[id]);
    if (debug_flag) printf("HRR    Cantera %+.15e KinetiX %+.15e relative error %e\n", ci.hrr[id], rates[id], e);
    for (int k = 0; k < n_active_species; k++) {
      const double ref = ci.rates[id][k];
      const double molar = rates[id + (size_t)(k + 1) * n_states] / (mw[k] * Mref);
      const double ek = std::abs(ref) > 1e-50 ? rel_err(molar, ref) : std::abs(molar);
      if (debug_flag)
        printf("%-6s Cantera %+.15e KinetiX %+.15e relative error %e\n", kx_species_name(k), ref, molar, ek);
      e = std::max(e, ek);
    }
    double rtol = single_precision ? 0.02 : 2e-08;
    if (id == 2 || id == 3) rtol = single_precision ? 0.02 : 5e-5;
    const bool ok = e < rtol;
    all &= ok;
    printf("rates error_inf: %e < %e (%s)\n", e, rtol, ok ? "passed" : "failed");
  }
  return all;
}

bool check_transport(const double* d_cond, const double* d_visc, const double* d_rhoD)   // bk.cpp:211-269
{
  auto cond = download(d_cond, n_states), visc = download(d_visc, n_states),
       rhoD = download(d_rhoD, (size_t)n_states * n_species);
  bool all = true;
  for (int id = 0; id < n_states; id++) {
    double e = std::max(rel_err(cond[id], ci.conductivity[id]), rel_err(visc[id], ci.viscosity[id]));
    for (int k = 0; k < n_species; k++) e = std::max(e, rel_err(rhoD[(size_t)k * n_states + id], ci.rhoD[id][k]));
    const double rtol = 1e-3;
    const bool ok = e < rtol;
    all &= ok;
    printf("transport error_inf: %e < %e (%s)\n", e, rtol, ok ? "passed" : "failed");
  }
  return all;
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void usage()
{
  printf("Usage: ./kinetix_bk --backend CUDA --n-states n --yaml-file s"
         "[--mode 1|2] [--tool s] [--n-repetitions n] [--single-precision] [--cimode n] [--debug] "
         "[--block-size  n] [--device-id  n] [--unroll-loops] [--loop-gibbsexp] "
         "[--group-rxnUnroll] [--group-vis] [--nonsymDij] [--fit-rcpDiffCoeffs] [--gpus n] [--random-states]\n");
}

}  // namespace

int main(int argc, char** argv)
{
  int mode = 0, blockSize = 0, nRep = 50, cimode = 0, deviceId = 0, gpus = 1;
  bool deviceIdFlag = false, random_states = false;
  kx_options opt;
  memset(&opt, 0, sizeof(opt));
  std::string backend, tool = "KinetiX", mech;
  int err = 0;

  static struct option long_options[] = {
      {"mode", required_argument, 0, 'e'},      {"backend", required_argument, 0, 'd'},
      {"tool", required_argument, 0, 't'},      {"n-states", required_argument, 0, 'n'},
      {"block-size", required_argument, 0, 'b'}, {"n-repetitions", required_argument, 0, 'r'},
      {"single-precision", no_argument, 0, 'p'}, {"debug", no_argument, 0, 'g'},
      {"cimode", required_argument, 0, 'c'},    {"yaml-file", required_argument, 0, 'f'},
      {"device-id", required_argument, 0, 'i'}, {"unroll-loops", no_argument, 0, 'u'},
      {"group-rxnUnroll", no_argument, 0, 'a'}, {"loop-gibbsexp", no_argument, 0, 'x'},
      {"group-vis", no_argument, 0, 'v'},       {"nonsymDij", no_argument, 0, 's'},
      {"fit-rcpDiffCoeffs", no_argument, 0, 'o'}, {"gpus", required_argument, 0, 'G'},
      {"random-states", no_argument, 0, 'R'},   {0, 0, 0, 0}};
  for (;;) {
    int idx = 0;
    const int c = getopt_long(argc, argv, "", long_options, &idx);
    if (c == -1) break;
    switch (c) {
      case 'e': mode = std::stoi(optarg); break;
      case 'd': backend = optarg; break;
      case 't': tool = optarg; break;
      case 'n': n_states = std::stoi(optarg); break;
      case 'b': blockSize = std::stoi(optarg); break;
      case 'r': nRep = std::stoi(optarg); break;
      case 'p': opt.single_precision = 1; break;
      case 'g': debug_flag = true; break;
      case 'c': cimode = std::stoi(optarg); break;
      case 'f': mech = optarg; break;
      case 'i': deviceId = std::stoi(optarg); deviceIdFlag = true; break;
      case 'u': opt.unroll_loops = 1; break;
      case 'a': opt.group_rxn_unroll = 1; break;
      case 'x': opt.loop_gibbsexp = 1; break;
      case 'v': opt.group_vis = 1; break;
      case 's': opt.nonsym_dij = 1; break;
      case 'o': opt.fit_rcp_diff_coeffs = 1; break;
      case 'G': gpus = std::max(1, std::stoi(optarg)); break;
      case 'R': random_states = true; break;
      default: err++;
    }
  }
  if (backend.empty() || mech.empty()) err++;
  if (err) {
    usage();
    return EXIT_FAILURE;
  }
  if (backend != "CUDA" && backend != "B200") {
    fprintf(stderr, "kinetix_bk (B200 build): --backend %s is not available here; only CUDA (sm_100a). "
                    "Use the reference build for SERIAL/HIP/DPCPP.\n", backend.c_str());
    return EXIT_FAILURE;
  }

  std::vector<std::string> ciStates;
  if (cimode) {
    if (gpus != 1) {
      printf("Running ci mode requires a single worker!");
      return EXIT_FAILURE;
    }
    if (cimode == 1) ciStates = {"initial", "ignition", "final"};
    if (cimode == 2 && mech.find("gri30") != std::string::npos) ciStates = {"ignition.highP"};
    n_states = (int)ciStates.size();
    nRep = 0;
  }

  opt.block_size = blockSize;
  opt.verbose = debug_flag;
  opt.tool = tool.c_str();
  // generate + compile the module once, before any worker exists (kinetix.cpp:290-296: rank 0 first)
  kx_ok(kx_prepare(mech.c_str(), &opt), "kx_prepare");

  // ---- workers: one process per GPU (the reference: one MPI rank per GPU, bk.cpp:523-531) ----
  int rank = 0;
  std::vector<int> pipes;
  std::vector<pid_t> children;
  const int n_total = n_states;
  if (gpus > 1) {
    // the reference gives every rank n_states / size states and drops the remainder (bk.cpp:443); here the first
    // n_states % gpus workers take one more, so the aggregate really is --n-states
    n_states = n_total / gpus + (0 < n_total % gpus ? 1 : 0);
    for (int r = 1; r < gpus; r++) {
      int fd[2];
      if (pipe(fd)) return EXIT_FAILURE;
      pid_t pid = fork();
      if (pid == 0) {
        rank = r;
        n_states = n_total / gpus + (r < n_total % gpus ? 1 : 0);
        close(fd[0]);
        pipes = {fd[1]};
        break;
      }
      close(fd[1]);
      pipes.push_back(fd[0]);
      children.push_back(pid);
    }
  }
  if (!deviceIdFlag) deviceId = rank;
  opt.device_id = deviceId;

  if (rank == 0) {
    std::cout << "number of states: " << n_states << '\n';
    std::cout << "number of repetitions: " << nRep << '\n';
  }

  kx_ok(kx_init(mech.c_str(), &opt), "kx_init");
  n_species = kx_n_species();
  n_reactions = kx_n_reactions();
  n_active_species = kx_n_active_species();

  // ---- state vector (bk.cpp:608-660) ----
  const size_t S = (size_t)n_states;
  std::vector<double> states((size_t)(n_species + 1) * S);
  double pressure = 1.0, ref_pressure = 1e5, ref_temperature = 1000;
  std::vector<double> ref_Y(n_species);
  const std::string stem = mech.substr(mech.find_last_of('/') + 1, mech.find_last_of('.') - mech.find_last_of('/') - 1);
  std::mt19937_64 rng(1234 + rank);
  std::uniform_real_distribution<double> uT(300.0, 2500.0), u01(0.0, 1.0);
  for (size_t id = 0; id < S; id++) {
    double p_Pa = 1e5, T_K = 1000;
    std::vector<double> X(n_species, 1.0 / n_species), M(n_species, 1.0);
    if (cimode) load_ci_state(ciStates[id], stem, p_Pa, T_K, X, M);
    double Mbar = 0;
    for (int k = 0; k < n_species; k++) Mbar += X[k] * M[k];
    std::vector<double> Y(n_species);
    for (int k = 0; k < n_species; k++) Y[k] = X[k] * M[k] / Mbar;
    if (random_states && !cimode) {
      p_Pa = 101325.0;
      T_K = uT(rng);
      double sum = 0;
      for (int k = 0; k < n_species; k++) sum += (Y[k] = u01(rng));
      for (int k = 0; k < n_species; k++) Y[k] /= sum;
    }
    if (id == 0) {   // reference state = first state
      ref_Y = Y;
      ref_pressure = (cimode == 2) ? 101325 : p_Pa;
      ref_temperature = random_states ? 1.0 : T_K;
    }
    pressure = p_Pa / ref_pressure;
    states[id] = T_K / ref_temperature;
    for (int k = 0; k < n_species; k++) states[id + (size_t)(k + 1) * S] = Y[k];
  }
  double *d_states, *d_a, *d_b, *d_c;
  CUDA_OK(cudaMalloc(&d_states, states.size() * sizeof(double)));
  CUDA_OK(cudaMemcpy(d_states, states.data(), states.size() * sizeof(double), cudaMemcpyHostToDevice));
  kx_ok(kx_build(ref_pressure, ref_temperature, ref_Y.data(), mode == 0 || mode == 2), "kx_build");
  const std::string module_path = kx_module_path() ? kx_module_path() : "";
  if (rank == 0) {
    printf("\n================= KinetiX (B200 native) =================\n");
    printf("module: %s\nyaml-file: %s\nnSpecies: %d\nTRef: %g K\npRef: %g Pa\n", kx_module_path(), mech.c_str(),
           n_species, ref_temperature, ref_pressure);
  }

  bool pass = true;
  double t_bk1 = 0, t_bk2 = 0;
  CUDA_OK(cudaMalloc(&d_a, (size_t)(n_species + 1) * S * sizeof(double)));
  CUDA_OK(cudaMalloc(&d_b, std::max<size_t>(S, 1) * sizeof(double)));
  CUDA_OK(cudaMalloc(&d_c, std::max<size_t>(S, 1) * sizeof(double)));

  // thermo: always once (bk.cpp:677-684)
  kx_ok(kx_thermodynamic_props(S, S, S, pressure, d_states, d_b, d_a, d_c, KX_DTYPE_F64, nullptr), "thermodynamicProps");
  CUDA_OK(cudaDeviceSynchronize());
  if (cimode && rank == 0 && mode == 0) pass &= check_thermo(d_b, d_a, d_c);

  if (mode == 0 || mode == 1) {
    kx_ok(kx_production_rates(S, S, S, pressure, d_states, d_a, KX_DTYPE_F64, nullptr), "productionRates");
    CUDA_OK(cudaDeviceSynchronize());
    const double t0 = now();
    for (int i = 0; i < nRep; i++)
      kx_ok(kx_production_rates(S, S, S, pressure, d_states, d_a, KX_DTYPE_F64, nullptr), "productionRates");
    CUDA_OK(cudaDeviceSynchronize());
    t_bk1 = now() - t0;
    if (cimode && rank == 0) pass &= check_rates(d_a, opt.single_precision);
  }
  if (mode == 0 || mode == 2) {
    double* d_rhoD = d_a;
    kx_ok(kx_mixture_avg_transport_props(S, S, S, pressure, d_states, d_b, d_c, d_rhoD, KX_DTYPE_F64, nullptr),
          "mixtureAvgTransportProps");
    CUDA_OK(cudaDeviceSynchronize());
    const double t0 = now();
    for (int i = 0; i < nRep; i++)
      kx_ok(kx_mixture_avg_transport_props(S, S, S, pressure, d_states, d_b, d_c, d_rhoD, KX_DTYPE_F64, nullptr),
            "mixtureAvgTransportProps");
    CUDA_OK(cudaDeviceSynchronize());
    t_bk2 = now() - t0;
    if (cimode && rank == 0) pass &= check_transport(d_c, d_b, d_rhoD);
  }

  // ---- gather worker times (max over workers, like the barrier-bracketed MPI_Wtime of the reference) ----
  if (rank != 0) {
    double t[2] = {t_bk1, t_bk2};
    if (write(pipes[0], t, sizeof(t)) < 0) {}
    _exit(0);
  }
  for (int fd : pipes) {
    double t[2] = {0, 0};
    if (read(fd, t, sizeof(t)) == (ssize_t)sizeof(t)) {
      t_bk1 = std::max(t_bk1, t[0]);
      t_bk2 = std::max(t_bk2, t[1]);
    }
  }
  for (pid_t c : children) waitpid(c, nullptr, 0);

  const double FP64_PEAK = 1.709e13;   // measured DFMA lane-instr/s per B200 (profiles/peaks_r01.json)
  // FP64 instructions per state the loaded BK1 kernel executes: the SASS census the build wrote beside the module
  // (counts.json; straight-line kernel: static = executed).  0 when absent.
  double w_bk1 = 0;
  {
    std::string mp = module_path;
    const size_t slash = mp.rfind('/');
    std::ifstream f((slash == std::string::npos ? std::string(".") : mp.substr(0, slash)) + "/counts.json");
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string txt = ss.str();
    const size_t b = txt.find("\"bk1\"");
    const size_t q = b == std::string::npos ? b : txt.find("\"fp64\":", b);
    if (q != std::string::npos) w_bk1 = atof(txt.c_str() + q + 7);
  }
  if (!cimode) {
    if (mode == 0 || mode == 1) {
      const double sps = (double)n_total * nRep / t_bk1;
      printf("BK1 (reaction rates) results:\n");
      printf("avg elapsed time: %.5f s\n", t_bk1);
      printf("avg aggregated throughput: %.2f GRXN/s\n", sps * n_reactions / 1e9);
      printf("avg aggregated throughput: %.4e states/s on %d GPU(s)\n", sps, gpus);
      if (w_bk1 > 0 && !opt.single_precision)
        printf("fraction of FP64 roofline by the %.0f FP64 instr/state this kernel executes: %.3f\n", w_bk1,
               sps / gpus * w_bk1 / FP64_PEAK);
      if (stem == "gri30" && !opt.single_precision)
        printf("  (by the reference's minimal-form estimate W = 1.4e4: %.3f)\n", sps / gpus * 1.4e4 / FP64_PEAK);
    }
    if (mode == 0 || mode == 2) {
      const double sps = (double)n_total * nRep / t_bk2;
      printf("BK2 (transport) results:\n");
      printf("avg elapsed time: %.5f s\n", t_bk2);
      printf("avg aggregated throughput: %.2f GDOF/s\n", sps * (n_species + 2) / 1e9);
      printf("avg aggregated throughput: %.4e states/s on %d GPU(s)\n", sps, gpus);
      if (stem == "gri30" && !opt.single_precision)
        printf("fraction of FP64 roofline: %.3f by the 1.79e4 FP64 instr/state the kernel executes (ncu, profiles/counts_r02.json);"
               " %.3f by the reference's as-emitted W = 2.47e4\n", sps / gpus * 1.79e4 / FP64_PEAK, sps / gpus * 2.47e4 / FP64_PEAK);
    }
  }
  if (pass && cimode) printf("all tests passed!\n");
  return pass ? EXIT_SUCCESS : EXIT_FAILURE;
}
