"""sm_100a emitter for BK1 with FP32 arithmetic (`--single-precision`).

The reference builds two reduced-precision flavours of its kernels from the same generated source
(reference benchmark/src/kinetix.cpp:254-281): "fpmix" (double buffers, float math -- what
`kinetix_bk --single-precision` actually runs) and pure FP32 (float buffers).  Both are
served by ONE kernel template here, `kx_bk1_f32<S>` with S = storage type.

The reference's FP32 code evaluates exp(sum nu g_k) and exp(ln A + ...) directly and returns NaN/Inf for
every state below ~615 K ("FP32 range").  This emitter works in log2 space instead:

    k_f   = ex2( l2A + beta*log2 T - (Ta*log2 e)/T )
    k_rev = ex2( [same exponent] + sum_k nu_k g2_k - (sum nu) * log2 C0 ),   g2_k = g_k/RT * log2 e

so only physical rate constants are ever exponentiated (no overflow at 300 K), with one MUFU.EX2 each.
Third-body / falloff factors multiply both.  exp / log / reciprocal are the MUFU approximations
(ex2.approx, lg2.approx, rcp.approx: 2^-22 relative), T-dependent scalars (ln T, 1/T) are computed in FP64
once per state and rounded.  Bound: per-state scaled error <= 1e-4 against the FP64 reference for
T in [300, 2500] K (tests/test_parity_gpu.py); the reference's own bound for this mode is 2e-2 (bk.cpp:198).

Scheduling (liveness-ordered units, activate / retire) is inherited from the FP64 emitter; with 4-byte
values the live C_k, g2_k and accumulators all stay in registers (no shared memory).
"""
import math

from . import constants as const
from .emit_bk1 import BK1Emitter, ConstPool, _lit

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453
LOG10_2 = 0.30102999566398120


class FloatPool(ConstPool):
    def __call__(self, v):
        v = float(v)
        if math.isinf(v) or abs(v) > 3.0e38:
            raise SystemExit(f'FP32 emitter: constant {v} does not fit single precision')
        if self.inline:
            # a 32-bit literal is an immediate operand or ONE MOV; a pool entry that is the second constant of an
            # FFMA costs an LDC through the MIO queue (ncu: 40 % of the FP32 kernel's stall samples sat on LDC)
            lit = _lit(v)
            return f'({lit}f)' if ('e' in lit or '.' in lit or 'n' in lit) else f'({lit}.f)'
        return super().__call__(v)

    def definition(self, ctype='float'):
        vals = self.values or [0.0]
        body = ',\n  '.join(', '.join(f'{_lit(v)}f' if 'e' in _lit(v) or '.' in _lit(v) else f'{_lit(v)}.f'
                                      for v in vals[i:i + 6]) for i in range(0, len(vals), 6))
        return f'__constant__ float {self.name}[{len(vals)}] = {{\n  {body}\n}};\n'


class BK1EmitterF32(BK1Emitter):
    def __init__(self, mech, K=None):
        super().__init__(mech, K or FloatPool('kcf'))

    # log2 of A T^b exp(-Ta/T) as an expression in L2T (= log2 T) and rcpT
    def l2_arrhenius(self, A, b, Ta):
        K = self.K
        expr = K(math.log2(A))
        if b != 0:
            expr = f'fmaf({K(b)}, L2T, {expr})'
        if Ta != 0:
            expr = f'fmaf({K(-Ta * LOG2E)}, rcpT, {expr})'
        return expr

    def l2_ratio(self, rx):
        """log2(k0/k_inf) (reaction_rates.py:244-259 semantics incl. its constant special case)."""
        A_inf, b_inf, E_inf = rx.rate.A, rx.rate.b, rx.rate.Ta
        A0, b0, E0 = rx.k0.A, rx.k0.b, rx.k0.Ta
        K = self.K
        if (A0 - A_inf) != 0 and ((b0 - b_inf) != 0 or (E0 - E_inf) != 0):
            expr = K(math.log2(A0) - math.log2(A_inf))
            if (b0 - b_inf) != 0:
                expr = f'fmaf({K(b0 - b_inf)}, L2T, {expr})'
            if (E0 - E_inf) != 0:
                expr = f'fmaf({K((-E0 + E_inf) * LOG2E)}, rcpT, {expr})'
            return expr
        return K(math.log2(A0 / A_inf))

    def emit(self, kernel_name='kx_bk1_f32', block=128, min_blocks=4, sync_every=0, reorder=True, nasa_indexed=False, **_):
        """nasa_indexed: False = every NASA-7 coefficient is a select between its two range values (immediates);
        'ldg' = one load from a range-indexed float table through L1 (as in the FP64 emitter)."""
        m, N, K = self.m, self.N, self.K
        self._flags = {}
        self.nasa_indexed, self.nasa_lo_tab, self.nasa_hi_tab = nasa_indexed, [], []
        body = []
        self.lines = body
        w = self.w

        units = self._units()
        usp = [self._species_of_unit(u) for u in units]
        order = self._schedule(units, usp, reorder)
        first, last = {}, {}
        for pos, ui in enumerate(order):
            for k in usp[ui]:
                first.setdefault(k, pos)
                last[k] = pos
        need_g = [False] * N
        for rx in m.reactions:
            if rx.reversible:
                for k, v in enumerate(rx.nu_net):
                    if v != 0:
                        need_g[k] = True

        eff_names = {}
        for rx in m.reactions:
            if rx.efficiencies is not None and tuple(rx.efficiencies) not in eff_names:
                eff_names[tuple(rx.efficiencies)] = f'M{len(eff_names)}'

        # ---- per-state scalars: FP64 once, then rounded ----
        w('const double Td = Tref * (double)kx_ld_stream(state + id);')
        w('const float T = (float)Td;')
        w('const float rcpT = (float)kx_rcp(Td);')
        w(f'const float L2T = (float)(kx_log(Td) * {LOG2E!r});')
        w('float rcpMbar = 0.f;')
        for name in eff_names.values():
            w(f'float {name} = 0.f;')
        w('{')
        for k in range(N):
            w(f'  const float w{k} = fmaxf(0.f, (float)kx_ld_stream(sp + {k} * offset)) * {K(1. / m.species[k].M)}; '
              f'rcpMbar += w{k};')
            for vec, name in eff_names.items():
                if vec[k] != 1:
                    w(f'  {name} = fmaf({K(vec[k] - 1)}, w{k}, {name});')
        w('}')
        w('const float rho = pR * rcpT * kx_rcpf(rcpMbar);')
        w('const float Cm = rho * rcpMbar;')
        for name in eff_names.values():
            w(f'{name} = fmaf(rho, {name}, Cm);')
        # log2 C0 = log2(p_atm/R) - log2 T
        w(f'const float L2C0 = {K(math.log2(const.ONE_ATM / const.R_GAS))} - L2T;')
        flag_pos = len(body)

        def collider(rx):
            if rx.efficiencies is not None:
                return eff_names[tuple(rx.efficiencies)]
            if rx.third_body_index >= 0:
                return f'cs{rx.third_body_index}'
            return 'Cm'

        used = sorted(first)
        w('float ' + ', '.join(f'cs{k}' for k in used) + ';')
        w('float ' + ', '.join(f'wd{k}' for k in used) + ';')
        if any(need_g[k] for k in used):
            w('float ' + ', '.join(f'g{k}' for k in used if need_g[k]) + ';')
        w('float hsum = 0.f;')

        def gcoef(a):   # g/RT*log2e = c0 + c1 L2T + c6/T + T (c2 + T (c3 + T (c4 + T c5)))
            return [(a[0] - a[6]) * LOG2E, -a[0], -a[1] / 2 * LOG2E, (1. / 3. - 1. / 2.) * a[2] * LOG2E,
                    (1. / 4. - 1. / 3.) * a[3] * LOG2E, (1. / 5. - 1. / 4.) * a[4] * LOG2E, a[5] * LOG2E]

        def hcoef(a):
            return [a[0], a[1] / 2, a[2] / 3, a[3] / 4, a[4] / 5, a[5]]

        def activate(k):
            w(f'cs{k} = fmaxf(0.f, (float)kx_ld_stream(sp + {k} * offset)) * ({K(1. / m.species[k].M)} * rho); '
              f'wd{k} = 0.f;')
            if need_g[k]:
                c, _, _ = self.nasa_select(k, gcoef)
                w(f'g{k} = fmaf(fmaf(fmaf(fmaf({c[5]}, T, {c[4]}), T, {c[3]}), T, {c[2]}), T, '
                  f'fmaf({c[1]}, L2T, fmaf({c[6]}, rcpT, {c[0]})));')

        def retire(k):
            c, _, _ = self.nasa_select(k, hcoef)
            w(f'if (live) kx_st_stream(out + {k} * offset, (S)({K(m.species[k].M)} * wd{k}));')
            w(f'hsum = fmaf(wd{k}, fmaf(fmaf(fmaf(fmaf({c[4]}, T, {c[3]}), T, {c[2]}), T, {c[1]}), T, '
              f'fmaf({c[5]}, rcpT, {c[0]})), hsum);')

        def conc_product(nu):
            terms = []
            for k, c in enumerate(nu):
                terms += [f'cs{k}'] * c
            return ' * '.join(terms)

        for k in range(N):
            if k not in first:
                w(f'if (live) kx_st_stream(out + {k} * offset, (S)0);')

        by_first, by_last = {}, {}
        for k, pos in first.items():
            by_first.setdefault(pos, []).append(k)
        for k, pos in last.items():
            by_last.setdefault(pos, []).append(k)

        emitted, live_now, peak_live = 0, 0, 0
        for pos, ui in enumerate(order):
            members = units[ui]
            for k in sorted(by_first.get(pos, [])):
                activate(k)
                live_now += 1
            peak_live = max(peak_live, live_now)
            if sync_every and emitted and emitted // sync_every != (emitted + len(members)) // sync_every:
                w('__syncthreads();')
            emitted += len(members)
            first_rx = m.reactions[members[0]]
            w(f'// ---- unit {pos}: reactions ' + ', '.join(str(i + 1) for i in members))
            w('{')
            if first_rx.kind != 'P-log':
                w(f'  const float l2base = {self.l2_arrhenius(first_rx.rate.A, first_rx.rate.b, first_rx.rate.Ta)};')
            for i in members:
                rx = m.reactions[i]
                w(f'  // {i + 1}: {rx.equation}')
                w('  {')
                if rx.kind == 'P-log':
                    self._emit_plog_f32(rx)
                elif i == members[0]:
                    w('    const float l2k = l2base;')
                else:
                    w(f'    const float l2k = l2base + {K(math.log2(rx.rate.A / first_rx.rate.A))};')
                self.stats['exp'] += 1
                w('    float fac = 1.f;')
                if rx.kind == 'three-body':
                    w(f'    fac = {collider(rx)};')
                elif rx.kind in ('pressure-modification', 'Troe', 'SRI'):
                    M = collider(rx)
                    w(f'    const float Pr = kx_ex2f({self.l2_ratio(rx)}) * {M};')
                    w('    const float rcp1Pr = kx_rcpf(1.f + Pr);')
                    self.stats['exp'] += 1
                    self.stats['rcp'] += 1
                    if rx.kind == 'pressure-modification':
                        w('    fac = Pr * rcp1Pr;')
                    elif rx.kind == 'Troe':
                        w(f'    const float logPr = kx_lg2f(Pr + 1e-37f) * {K(LOG10_2)};')
                        w(f'    const float l2Fc = kx_lg2f({self._troe_fcent_f32(rx.troe)});')
                        w(f'    const float logFc = l2Fc * {K(LOG10_2)};')
                        w('    const float tc = fmaf(-0.67f, logFc, -0.4f) + logPr;')
                        w('    const float tn = fmaf(-1.27f, logFc, 0.75f) - 0.14f * tc;')
                        w('    const float tn2 = tn * tn;')
                        w('    fac = Pr * rcp1Pr * kx_ex2f(l2Fc * tn2 * kx_rcpf(fmaf(tc, tc, tn2)));')
                        self.stats['log'] += 2
                        self.stats['exp'] += 1
                        self.stats['rcp'] += 1
                    else:
                        sr = rx.sri
                        w(f'    const float logPr = kx_lg2f(Pr) * {K(LOG10_2)};')
                        w(f'    const float sb = {K(sr["A"])} * kx_ex2f({K(-sr["B"] * LOG2E)} * rcpT) + '
                          f'kx_ex2f({K(-LOG2E / (sr["C"] + const.FLOAT_MIN))} * T);')
                        w(f'    fac = Pr * rcp1Pr * {K(sr["D"])} * '
                          f'kx_ex2f(kx_lg2f(sb) * kx_rcpf(fmaf(logPr, logPr, 1.f)) + {K(sr["E"])} * L2T);')
                        self.stats['log'] += 2
                        self.stats['exp'] += 3
                Rf = conc_product(rx.nu_reac)
                net = rx.nu_net
                w('    const float kf = kx_ex2f(l2k);')
                if not rx.reversible:
                    w(f'    const float q = fac * kf * {Rf};')
                else:
                    terms = []
                    for k, v in enumerate(net):
                        if v == 1:
                            terms.append(f'+ g{k}')
                        elif v == -1:
                            terms.append(f'- g{k}')
                        elif v != 0:
                            terms.append(f'+ {float(v)}f * g{k}')
                    sn = sum(net)
                    if sn != 0:
                        terms.append(f'- {float(sn)}f * L2C0')
                    w(f'    const float krev = kx_ex2f(l2k {" ".join(terms)});')
                    self.stats['exp'] += 1
                    Rr = conc_product(rx.nu_prod)
                    w(f'    const float q = fac * fmaf(-krev, {Rr}, kf * {Rf});')
                for k, v in enumerate(net):
                    if v == 1:
                        w(f'    wd{k} += q;')
                    elif v == -1:
                        w(f'    wd{k} -= q;')
                    elif v != 0:
                        w(f'    wd{k} = fmaf({float(v)}f, q, wd{k});')
                w('  }')
            w('}')
            for k in sorted(by_last.get(pos, [])):
                retire(k)
                live_now -= 1
        w(f'if (live) kx_st_stream(rates + id, (S)({K(-const.R_GAS)} * T * hsum));')

        self.smem_doubles_per_thread = 0
        self.schedule_stats = dict(units=len(units), peak_live=peak_live, smem_slots=0)
        body[flag_pos:flag_pos] = ['  ' + v for v in self._flags.values()]
        head = [
            f'// BK1, FP32 math (fpmix: S = double, fp32: S = float): {m.name}; {self.stats["exp"]} ex2, '
            f'{self.stats["log"]} lg2, {self.stats["rcp"]} rcp per state; peak live species {peak_live}',
            'template <typename S, bool PF>',
            f'__global__ void __launch_bounds__({block}, {min_blocks})',
            f'{kernel_name}(const long long n_states, const long long offsetT, const long long offset,',
            '           const float pressure_R, const float P, const float lnP,',
            '           const S* __restrict__ state, S* __restrict__ rates, const double Tref,',
            '           const S* __restrict__ pfield)',
            '{',
            '  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;',
            '  const bool live = gid < n_states;',
            '  const long long id = live ? gid : n_states - 1;',
            '  const S* sp = state + id + offsetT;',
            '  S* out = rates + id + offsetT;',
            '  // per-state pressure (extension): pfield[id] = p / p_ref, the scalars then carry p_ref',
            '  float pR = pressure_R, Pv = P, lnPv = lnP;',
            '  if (PF) {',
            '    const float pn = (float)kx_ld_stream(pfield + id);',
            '    pR *= pn;',
            '    Pv *= pn;',
            f'    lnPv = {"kx_log(Pv)" if any(r.kind == "P-log" for r in m.reactions) else "lnP"};',
            '  }',
        ]
        return '\n'.join(head + body + ['}', ''])

    def tmid_offset(self, tmid):
        name = 'noff_' + repr(float(tmid)).replace('.', '_').replace('-', 'm')
        if name not in self._flags:
            self._flags[name] = f'const int {name} = (T <= {_lit(float(tmid))}f) ? 0 : KX_NASA_HALF;'
        return name

    def nasa_table_definition(self):
        if not self.nasa_indexed:
            return ''
        vals = self.nasa_lo_tab + self.nasa_hi_tab
        body = ',\n  '.join(', '.join(f'{float(v)!r}f' for v in vals[i:i + 6]) for i in range(0, len(vals), 6))
        return (f'#define KX_NASA_HALF {len(self.nasa_lo_tab)}\n#define KX_NASA_LEN {len(vals)}\n'
                f'__device__ const __align__(16) float kx_nasa_tab[{len(vals)}] = {{\n  {body}\n}};\n')

    def tmid_flag(self, tmid):
        name = 'lo_' + repr(float(tmid)).replace('.', '_').replace('-', 'm')
        if name not in self._flags:
            self._flags[name] = f'const bool {name} = T <= {_lit(float(tmid))}f;'
        return name

    def _troe_fcent_f32(self, tr):
        K = self.K
        T_LO, T_HI = 200.0, 6000.0

        def texp(coef_T, weight):
            lo, hi = sorted((coef_T * T_LO, coef_T * T_HI))
            if hi < -120.0:
                return None                       # underflows to 0 in FP32 over the whole range
            if abs(lo) < 1e-9 and abs(hi) < 1e-9:
                return K(weight)
            e = f'kx_ex2f({K(coef_T * LOG2E)} * T)'
            return e if weight == 1 else f'{K(weight)} * {e}'

        A = tr['A']
        if A == 0:
            parts = [texp(-1. / (tr['T3'] + const.FLOAT_MIN), 1.0)]
        elif A == 1:
            parts = [texp(-1. / (tr['T1'] + const.FLOAT_MIN), 1.0)]
        else:
            parts = [texp(-1. / (tr['T3'] + const.FLOAT_MIN), 1 - A), texp(-1. / (tr['T1'] + const.FLOAT_MIN), A)]
        if tr['T2'] < float('inf'):
            parts.append(f'kx_ex2f({K(-tr["T2"] * LOG2E)} * rcpT)')
        parts = [p for p in parts if p is not None]
        self.stats['exp'] += len(parts)
        return ' + '.join(parts) if parts else '1e-37f'

    def _emit_plog_f32(self, rx):
        K, w = self.K, self.w

        def l2sum(ks):
            if len(ks) == 1:
                return self.l2_arrhenius(ks[0].A, ks[0].b, ks[0].Ta)
            return 'kx_lg2f(' + ' + '.join(f'kx_ex2f({self.l2_arrhenius(k.A, k.b, k.Ta)})' for k in ks) + ')'

        pl = rx.plog
        n = len(pl)
        w('    float l2k;')
        for i in range(n - 1):
            (p1, k1), (p2, k2) = pl[i], pl[i + 1]
            lnp1, lnp2 = math.log(p1), math.log(p2)
            w(f'    {"if" if i == 0 else "} else if"} ((Pv > {_lit(p1)}f) && (Pv < {_lit(p2)}f)) {{')
            w(f'      const float a1 = {l2sum(k1)}, a2 = {l2sum(k2)};')
            w(f'      l2k = fmaf((a2 - a1) * (lnPv - {K(lnp1)}), {K(1 / (lnp2 - lnp1))}, a1);')
            if i == 0:
                w(f'    }} else if (Pv <= {_lit(p1)}f) {{')
            else:
                w(f'    }} else if (Pv == {_lit(p1)}f) {{')
            w(f'      l2k = {l2sum(k1)};')
            if i == n - 2:
                w(f'    }} else if (Pv >= {_lit(p2)}f) {{')
                w(f'      l2k = {l2sum(k2)};')
        w('    } else { l2k = -150.f; }')
