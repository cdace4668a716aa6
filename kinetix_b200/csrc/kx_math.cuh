// kx_math.cuh -- FP64 / FP32 math primitives for the generated BK1/BK2 kernels (sm_100a).
//
// Both hot paths are FP64-pipe bound on B200 (64 DFMA lanes/clk/SM): what matters is the number of
// FP64-pipe instructions per transcendental, not bytes.  libdevice's exp/log/div cost 18 / 30 / ~10
// FP64 instructions plus special-case handling the kernels do not need (arguments are finite and in a
// known range).  The versions here are branch-light, accurate to <= 2 ulp (measured against mpmath,
// tools/fit_math_polys.py) and cost
//
//     kx_exp   15 FP64  (Cody-Waite reduction + degree-11 polynomial, exponent patched with integer ops)
//     kx_log   ~17 FP64 (atanh series in s=(m-1)/(m+1), reciprocal by MUFU.RCP64H + Newton)
//     kx_rcp    3 FP64 + 1 MUFU
//
// The reference relies on the backend's libm for these (benchmark/src/kinetix.cpp:246-251 maps
// __KINETIX_EXP__/LOG__/LOG10__/POW__ onto exp/log/log10/pow); parity is to 1e-10, not bit-wise.
#pragma once
#include <cuda_runtime.h>

#define KX_DEVICE __device__ __forceinline__

// ---- reciprocal ------------------------------------------------------------------------------
// MUFU.RCP64H seed (measured max relative error 9.9e-7 on B200, profiles/peaks_r01.json) + one cubic
// Newton step: error e0^3 ~ 1e-18, measured 1.1e-16 (= rounding).  3 DFMA + 1 MUFU, no range checks.
// Valid for normal, finite, non-zero |a| in about [1e-300, 1e300] -- all uses satisfy this.
KX_DEVICE double kx_rcp(double a)
{
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  e = fma(e, e, e);
  return fma(x, e, x);
}

// Reciprocal with ONE quadratic Newton step: x0 (2 - a x0), 2 FP64 ops, relative error e0^2 ~ 1e-12.
// Used only where the result feeds sums of positive terms (BK2 pair loop), 100x inside the 1e-10 contract.
KX_DEVICE double kx_rcp_fast(double a)
{
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  return x * fma(-a, x, 2.0);
}

// a / b with the reciprocal above (one extra DMUL); relative error ~2e-16.
KX_DEVICE double kx_div(double a, double b) { return a * kx_rcp(b); }

// ---- exp -------------------------------------------------------------------------------------
#ifdef KX_EXP_TABLE
// Table-driven exp for the kernels that define KX_EXP_TABLE (BK1, FP64):  x = n ln2/32 + r,  |r| <= ln2/64,
// e^x = 2^(n >> 5) * 2^((n & 31)/32) * e^r  with a degree-5 polynomial for e^r (near-minimax, max relative error
// 2.5e-16, tools/fit_math_polys.py `expt32`) and a 32-entry table of 2^(j/32) in shared memory: 10 FP64-pipe
// instructions instead of 15, plus one LDS (lanes pick different entries: at worst a 2-way bank conflict on a pipe
// this kernel leaves 90 % idle) and three integer instructions.  The table is read at the END of the dependency
// chain, so its latency is hidden behind the polynomial.
__constant__ double kx_exptab_c[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002};
__shared__ double kx_exptab[32];
// every thread of the CTA calls this once, before the first exp
__device__ __forceinline__ void kx_exptab_init()
{
  if (threadIdx.x < 32) kx_exptab[threadIdx.x] = kx_exptab_c[threadIdx.x];
  __syncthreads();
}
__device__ __forceinline__ double kx_exp_table_core(double x, int lo_clamp, int hi_clamp, bool clamp)
{
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
  double kd = fma(x, 46.16624130844683, MAGIC);            // 32 / ln 2
  int n = __double2loint(kd);
  kd -= MAGIC;
  double r = fma(kd, -6.93147180369123816490e-01 / 32.0, x);   // (ln2 / 32) high part, exact product
  r = fma(kd, -1.90821492927058770002e-10 / 32.0, r);          // low part
  const double t = kx_exptab[n & 31];
  double s = 0.008333368243545858;
  s = fma(s, r, 0.04166691103828231);
  s = fma(s, r, 0.16666666666513108);
  s = fma(s, r, 0.4999999999892509);
  s = fma(s, r, 1.0);
  const double y = fma(t, s * r, t);
  int k = n >> 5;
  if (clamp) k = max(min(k, hi_clamp), lo_clamp);
  return __hiloint2double(__double2hiint(y) + (k << 20), __double2loint(y));
}
#endif

// exp(x), x finite.  The exponent is clamped to the normal range with two integer min/max (ALU pipe),
// so results saturate at ~2^-1022 / ~2^1023 instead of going through denormals / inf: absolute
// differences of < 1e-307 against libm.  The emitter proves |x| < 700 for every call it routes here
// over the validity range of T (interval arithmetic at generation time) and uses kx_exp_wide otherwise.
KX_DEVICE double kx_exp(double x)
{
#ifdef KX_EXP_TABLE
  return kx_exp_table_core(x, -1021, 1022, true);
#endif
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
  double kd = fma(x, 1.4426950408889634, MAGIC);
  int k = __double2loint(kd);
  kd -= MAGIC;
  double r = fma(kd, -6.93147180369123816490e-01, x);   // ln2 high part (32 trailing zero bits)
  r = fma(kd, -1.90821492927058770002e-10, r);          // ln2 low part
  double p = 2.5110037605963777e-08;
  p = fma(p, r, 2.763263963904103e-07);
  p = fma(p, r, 2.755724091857897e-06);
  p = fma(p, r, 2.4801485482328494e-05);
  p = fma(p, r, 0.00019841269890047113);
  p = fma(p, r, 0.0013888888952314775);
  p = fma(p, r, 0.008333333333319601);
  p = fma(p, r, 0.0416666666664881);
  p = fma(p, r, 0.1666666666666668);
  p = fma(p, r, 0.5000000000000019);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  k = max(min(k, 1022), -1021);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// exp(x) for arguments the emitter has PROVEN to lie in [-690, 690] for every valid state: no exponent
// clamp at all (14 FP64 + 3 integer instructions).
// (Round 2 experiment: ONE copy of this body reached by CALL instead of ~22 inlined instructions per use shrinks the
// straight-line BK1 text by 19-31 % and removes nearly all register spills, but executes ~20 % more instructions
// (CALL/RET, argument moves, no scheduling across the call): GRI-3.0 958 vs 957 M states/s, heptaneLu88 467 vs 652,
// EtOHKonnov 113 vs 164 -- rejected.)
KX_DEVICE double kx_exp_nc(double x)
{
#ifdef KX_EXP_TABLE
  return kx_exp_table_core(x, 0, 0, false);
#endif
  const double MAGIC = 6755399441055744.0;
  double kd = fma(x, 1.4426950408889634, MAGIC);
  const int k = __double2loint(kd);
  kd -= MAGIC;
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
#if defined(KX_EXP_EVENODD)
  {
    // even/odd split: two Horner chains in r^2 of depth 5/6 (+1 DMUL, +1 DFMA over plain Horner)
    const double r2 = r * r;
    double e = 2.763263963904103e-07;            // c10
    e = fma(e, r2, 2.4801485482328494e-05);      // c8
    e = fma(e, r2, 0.0013888888952314775);       // c6
    e = fma(e, r2, 0.0416666666664881);          // c4
    e = fma(e, r2, 0.5000000000000019);          // c2
    e = fma(e, r2, 1.0);                         // c0
    double o = 2.5110037605963777e-08;           // c11
    o = fma(o, r2, 2.755724091857897e-06);       // c9
    o = fma(o, r2, 0.00019841269890047113);      // c7
    o = fma(o, r2, 0.008333333333319601);        // c5
    o = fma(o, r2, 0.1666666666666668);          // c3
    o = fma(o, r2, 1.0);                         // c1
    const double q = fma(o, r, e);
    return __hiloint2double(__double2hiint(q) + (k << 20), __double2loint(q));
  }
#elif defined(KX_EXP_ESTRIN)
  {
    // Estrin evaluation of the same degree-11 polynomial: 3 more multiplies, dependency depth 5 instead of 11
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a01 = fma(1.0, r, 1.0), a23 = fma(0.1666666666666668, r, 0.5000000000000019);
    const double a45 = fma(0.008333333333319601, r, 0.0416666666664881);
    const double a67 = fma(0.00019841269890047113, r, 0.0013888888952314775);
    const double a89 = fma(2.755724091857897e-06, r, 2.4801485482328494e-05);
    const double aab = fma(2.5110037605963777e-08, r, 2.763263963904103e-07);
    const double b0 = fma(a23, r2, a01), b1 = fma(a67, r2, a45), b2 = fma(aab, r2, a89);
    const double q = fma(b2, r8, fma(b1, r4, b0));
    return __hiloint2double(__double2hiint(q) + (k << 20), __double2loint(q));
  }
#endif
#ifndef KX_EXP_DEG10
  double p = 2.5110037605963777e-08;
  p = fma(p, r, 2.763263963904103e-07);
  p = fma(p, r, 2.755724091857897e-06);
  p = fma(p, r, 2.4801485482328494e-05);
  p = fma(p, r, 0.00019841269890047113);
  p = fma(p, r, 0.0013888888952314775);
  p = fma(p, r, 0.008333333333319601);
  p = fma(p, r, 0.0416666666664881);
  p = fma(p, r, 0.1666666666666668);
  p = fma(p, r, 0.5000000000000019);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
#else
  // degree 10 (Chebyshev-interpolated on |r| <= ln2/2): max relative error 4.2e-16 against mpmath
  // (tools/fit_math_polys.py), one DFMA less than degree 11 (1.4e-16) -- measured no faster (906 vs 914 M st/s)
  double p = 2.7626357241447223e-07;
  p = fma(p, r, 2.764018079620985e-06);
  p = fma(p, r, 2.4801504346997686e-05);
  p = fma(p, r, 0.00019841170270440067);
  p = fma(p, r, 0.0013888888932488599);
  p = fma(p, r, 0.008333333385667782);
  p = fma(p, r, 0.04166666666657314);
  p = fma(p, r, 0.16666666666554406);
  p = fma(p, r, 0.5000000000000006);
  p = fma(p, r, 1.0000000000000067);
  p = fma(p, r, 1.0);
#endif
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// Full-range exp (libdevice): underflow to 0 / overflow to inf exactly like the reference's libm.
KX_DEVICE double kx_exp_wide(double x) { return exp(x); }

// ---- log -------------------------------------------------------------------------------------
// log(x) for normal positive finite x.  x = 2^e * m with m in [sqrt(1/2), sqrt(2)),
// log(m) = 2 atanh(s), s = (m-1)/(m+1), atanh(s)/s = 1 + z q(z), z = s^2, |s| <= 0.1716.
KX_DEVICE double kx_log(double x)
{
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  // bring the mantissa into [sqrt(.5), sqrt(2)): subtract the high word of sqrt(.5) (0x3fe6a09e)
  int e = (hi - 0x3fe6a09e) >> 20;
  hi -= e << 20;
  const double m = __hiloint2double(hi, lo);
  const double s = (m - 1.0) * kx_rcp(m + 1.0);
  const double z = s * s;
  double q = 0.07308224842521703;
  q = fma(q, z, 0.07665860800278021);
  q = fma(q, z, 0.09091444562630861);
  q = fma(q, z, 0.1111110556739754);
  q = fma(q, z, 0.14285714312987743);
  q = fma(q, z, 0.19999999999949752);
  q = fma(q, z, 0.3333333333333335);
  const double s2 = s + s;
  const double lm = fma(s2 * z, q, s2);
  const double ed = (double)e;
  // e*ln2 split so that the high product is exact for |e| < 2^11
  return fma(ed, 6.93147180369123816490e-01, fma(ed, 1.90821492927058770002e-10, lm));
}

KX_DEVICE double kx_log10(double x) { return kx_log(x) * 0.4342944819032518; }

// x^y for x > 0 and 10^x
KX_DEVICE double kx_pow(double x, double y) { return kx_exp(y * kx_log(x)); }
KX_DEVICE double kx_exp10(double x) { return kx_exp(x * 2.302585092994046); }

// ---- streaming global access -----------------------------------------------------------------
// State rows are read once and result rows written once: keep them out of L1 and mark evict-first.
KX_DEVICE double kx_ld_stream(const double* p)
{
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// Row accesses of the species-major slabs: element `base[K * offset]`.  The address arithmetic happens INSIDE
// the volatile asm: otherwise the compiler computes all 2 x N row addresses up front, cannot keep them in
// registers and spills them (GRI-3.0: 0.75 KB, EtOHKonnov: 3 KB of local memory per thread were nothing but
// addresses).  One 64-bit multiply-add per access instead.
template <int K>
KX_DEVICE double kx_ld_row(const double* base, long long offset)
{
  double v;
  asm volatile("{\n\t.reg .s64 a;\n\tmad.lo.s64 a, %2, %3, %1;\n\tld.global.nc.L1::no_allocate.f64 %0, [a];\n\t}"
               : "=d"(v) : "l"(base), "l"(offset), "n"(K * 8));
  return v;
}
template <int K>
KX_DEVICE void kx_st_row(double* base, long long offset, double v)
{
  asm volatile("{\n\t.reg .s64 a;\n\tmad.lo.s64 a, %1, %2, %0;\n\tst.global.L1::no_allocate.f64 [a], %3;\n\t}"
               ::"l"(base), "l"(offset), "n"(K * 8), "d"(v) : "memory");
}
// base[K*offset] += v  (thread-private element: partial rate rows under live_cap)
template <int K>
KX_DEVICE void kx_add_row(double* base, long long offset, double v)
{
  asm volatile("{\n\t.reg .s64 a;\n\t.reg .f64 t;\n\tmad.lo.s64 a, %1, %2, %0;\n\tld.global.f64 t, [a];\n\t"
               "add.f64 t, t, %3;\n\tst.global.f64 [a], t;\n\t}"
               ::"l"(base), "l"(offset), "n"(K * 8), "d"(v) : "memory");
}

// read-only load that MAY stay in L1: used for the state rows BK1 reads twice (pass 1 for the mean molar
// mass, then again when a species is activated) so the second read can hit L1 instead of L2.
KX_DEVICE double kx_ld_keep(const double* p)
{
  double v;
  asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
KX_DEVICE void kx_st_stream(double* p, double v)
{
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// ---- FP32 flavour (fpmix / fp32 kernels): MUFU approximations, 2^-22 relative ---------------------
KX_DEVICE float kx_ex2f(float x)
{
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
KX_DEVICE float kx_lg2f(float x)
{
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
KX_DEVICE float kx_rcpf(float x)
{
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// overloads so that templated kernels (csrc/kx_bk2.cuh, kx_thermo.cuh) are written once
KX_DEVICE float kx_rcp(float a) { return kx_rcpf(a); }
KX_DEVICE float kx_rcp_fast(float a) { return kx_rcpf(a); }
KX_DEVICE float kx_log(float x) { return kx_lg2f(x) * 0.69314718f; }
KX_DEVICE float kx_sqrt(float x) { return sqrtf(x); }
KX_DEVICE double kx_sqrt(double x) { return sqrt(x); }
KX_DEVICE float kx_ld_stream(const float* p)
{
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
KX_DEVICE void kx_st_stream(float* p, float v)
{
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- asynchronous 8-byte global -> shared copy (LDGSTS) with FIFO completion ----------------------
// One commit group per copy; kx_cp_async_wait<N>() returns when all but the N most recent groups of
// this thread have landed; kx_ring_read() must be used to read the landed value.  All three are volatile
// asm WITHOUT a "memory" clobber: they stay ordered among themselves (which is all the protocol needs)
// while the compiler remains free to schedule every other memory operation across them.
KX_DEVICE void kx_cp_async8(unsigned smem_addr, const double* gptr)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n\tcp.async.commit_group;" ::"r"(smem_addr), "l"(gptr));
}
template <int N>
KX_DEVICE void kx_cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
KX_DEVICE double kx_ring_read(unsigned smem_addr)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_addr));
  return v;
}
