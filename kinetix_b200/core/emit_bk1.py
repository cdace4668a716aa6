"""sm_100a emitter for BK1: species net production rates + heat release rate.

Emits one mechanism-specialised CUDA kernel (text) from the Mechanism IR.  What the kernel computes is
what the reference's `productionRates` OKL kernel computes around its generated `kinetix_species_rates`
and `kinetix_enthalpy_RT` (reference benchmark/okl/productionRates.okl:10-64,
kinetix/core/reaction_rates.py:291-416,549-613, thermodynamics.py:65-80); HOW it is computed is
re-designed for the B200 FP64 pipe:

  * one thread = one state; all loads/stores species-major and coalesced (row k at k*offset + id);
  * every constant lives in one `__constant__` pool `kc[]` addressed with literal indices, so FP64
    instructions take constant-bank / uniform-register operands instead of materialising 64-bit
    immediates with pairs of MOVs (the reference-style kernel spends >25 % of its issue slots on that);
  * "minimal form": one exp(g_k) and one reciprocal per species (not one exp per reversible
    reaction), equilibrium constants as products of those; rate constants that share (beta, Ta) share
    one exp; distinct third-body efficiency vectors are evaluated once; log10(Pr) of falloff reactions
    is obtained from the exponent of Pr and ln(M) per distinct collider instead of a log per reaction;
    Troe terms that are exactly 0 or 1 over the validity range are folded at generation time;
  * exp/log/reciprocal are the short sequences of csrc/kx_math.cuh.

Results agree with the reference's generated code to rounding-level differences (per-state scaled
error ~1e-14; the parity bound is 1e-10, tests/test_parity_gpu.py).
"""
import math

from . import constants as const

T_VALID_LO = 200.0      # the emitter proves exp() arguments stay in range for T in this interval
T_VALID_HI = 6000.0
EXP_SAFE = 690.0


class ConstPool:
    """De-duplicated pool of double constants -> `kc[i]` references."""

    PARAM_LIMIT = 3800              # doubles that fit the 32 KB kernel-parameter space beside the other args

    def __init__(self, name='kc', inline=False, as_param=False):
        self.name = name
        self.inline = inline        # emit literals (compiler materialises them) instead of pool loads
        self.as_param = as_param    # pass the pool BY VALUE as a __grid_constant__ kernel parameter (bank 0)
        self.values = []
        self._index = {}

    def __call__(self, v):
        v = float(v)
        if self.inline and not math.isinf(v):
            return f'({_lit(v)})'
        key = v.hex()
        if key not in self._index:
            self._index[key] = len(self.values)
            self.values.append(v)
        i = self._index[key]
        if self.as_param:
            # the first PARAM_LIMIT constants travel in the parameter struct, the rest stay __constant__
            return f'kp.v[{i}]' if i < self.PARAM_LIMIT else f'{self.name}[{i - self.PARAM_LIMIT}]'
        return f'{self.name}[{i}]'

    def definition(self, ctype='double'):
        def fmt(vals):
            vals = vals or [0.0]
            return ',\n  '.join(', '.join(_lit(v) for v in vals[i:i + 4]) for i in range(0, len(vals), 4)), len(vals)
        if not self.as_param:
            body, n = fmt(self.values)
            return f'__constant__ {ctype} {self.name}[{n}] = {{\n  {body}\n}};\n'
        head, tail = self.values[:self.PARAM_LIMIT], self.values[self.PARAM_LIMIT:]
        hb, hn = fmt(head)
        tb, tn = fmt(tail)
        return (f'struct KxParamPool {{ {ctype} v[{hn}]; }};\n'
                f'static const KxParamPool kx_param_pool = {{{{\n  {hb}\n}}}};\n'
                f'__constant__ {ctype} {self.name}[{tn}] = {{\n  {tb}\n}};\n')


def _lit(v):
    if math.isinf(v):
        return 'CUDART_INF' if v > 0 else '-CUDART_INF'
    return repr(float(v))


def _arg_range(c0, c_lnT, c_rcpT, c_T=0.0):
    """Conservative range of c0 + c_lnT*ln T + c_rcpT/T + c_T*T over the validity interval."""
    lo = hi = c0
    for c, a, b in ((c_lnT, math.log(T_VALID_LO), math.log(T_VALID_HI)),
                    (c_rcpT, 1 / T_VALID_HI, 1 / T_VALID_LO),
                    (c_T, T_VALID_LO, T_VALID_HI)):
        x, y = c * a, c * b
        lo += min(x, y)
        hi += max(x, y)
    return lo, hi


def _exp_fn(lo, hi):
    return 'kx_exp_nc' if (lo > -EXP_SAFE and hi < EXP_SAFE) else 'kx_exp_wide'


class BK1Emitter:
    def __init__(self, mech, K=None):
        self.m = mech
        self.N = mech.n_species
        self.K = K or ConstPool()
        self.lines = []
        self.stats = dict(exp=0, exp_wide=0, log=0, rcp=0)

    # ------------------------------------------------------------------------------------------
    def w(self, s=''):
        self.lines.append('  ' + s if s else '')

    def exp(self, arg_expr, lo, hi):
        fn = _exp_fn(lo, hi)
        self.stats['exp' if fn == 'kx_exp_nc' else 'exp_wide'] += 1
        return f'{fn}({arg_expr})'

    def EG(self, k):
        return self.val(self.eg_slot[k]) if self.gibbs_in_smem else f'eg{k}'

    def RG(self, k):
        return self.val(self.rg_slot[k]) if self.gibbs_in_smem else f'rg{k}'

    # ---- per-thread scratch slots: shared memory [slot][thread] or tensor memory (slot ids >= TM_BASE) --------
    TM_BASE = 10000

    def is_tm(self, slot):
        return slot >= self.TM_BASE

    def val(self, slot):
        """rvalue of a slot: the shared-memory word itself, or the temporary a preceding fetch() has loaded"""
        if self.is_tm(slot):
            assert slot in self._fetched, 'tensor-memory slot read without fetch()'
            return f'tv{slot - self.TM_BASE}'
        return f'gs[{slot} * {self.block}]'

    def store(self, slot, expr, indent=''):
        if self.is_tm(slot):
            self.w(f'{indent}kx_tm_st_d(tmb + {2 * (slot - self.TM_BASE)}, {expr});')
            self._tm_dirty = True
        else:
            self.w(f'{indent}gs[{slot} * {self.block}] = {expr};')

    def fetch(self, slots, indent=''):
        """make the tensor-memory slots among `slots` readable inside the current C scope: batched tcgen05.ld,
        one wait, values pinned behind the wait"""
        tm = [t for t in dict.fromkeys(slots) if self.is_tm(t)]
        self._fetched = set(tm)
        if not tm:
            return
        w = self.w
        if self._tm_dirty:
            w(f'{indent}kx_tm_wait_st();')
            self._tm_dirty = False
        for t in tm:
            j = t - self.TM_BASE
            w(f'{indent}unsigned tl{j}, th{j}; kx_tm_ld_d(tmb + {2 * j}, tl{j}, th{j});')
        w(f'{indent}kx_tm_wait_ld();')
        for t in tm:
            j = t - self.TM_BASE
            w(f'{indent}const double tv{j} = kx_tm_pin(tl{j}, th{j});')
        self.stats['tm_ld'] = self.stats.get('tm_ld', 0) + len(tm)

    # ---- thermo ------------------------------------------------------------------------------
    def nasa_select(self, k, make):
        """coefficient list for species k selected on T <= T_mid; `make(a)` maps the 7 NASA
        coefficients to the derived coefficients actually needed.
        nasa_indexed: the low-range sets of all species live in the first half of one table and the high-range
        sets in the second half; a per-T_mid integer offset (0 or half) picks the range, so a coefficient is ONE
        load with a register offset instead of two loads and a 64-bit select.  The offset differs between the
        lanes of a warp, so where the table lives matters (GRI-3.0, M states/s): 'ldg' = global memory through L1
        958 (default), 'smem' = per-CTA copy in shared memory 945, True / 'const' = __constant__ (the load
        serialises over the two addresses and misses the 2 KB constant cache) 907; 128-bit LDG pairs 936; round 2:
        an L1 evict-last hint on the table loads 916 (L1 hit rate 47 -> 41 %, more spills), no table at all -- both
        ranges evaluated from constant operands, the result selected -- 790 (1.7 KB of spills)."""
        s = self.m.species[k]
        lo, hi = make(s.nasa_lo), make(s.nasa_hi)
        if getattr(self, 'nasa_indexed', False):
            base = len(self.nasa_lo_tab)
            self.nasa_lo_tab += [float(v) for v in lo]
            self.nasa_hi_tab += [float(v) for v in hi]
            off = self.tmid_offset(s.T_mid)
            mode = self.nasa_indexed
            if mode == 'ldg':        # global memory through L1 (two addresses per warp at most)
                return [f'__ldg(&kx_nasa_tab[{off} + {base + i}])' for i in range(len(lo))], lo, hi
            if mode == 'smem':       # CTA-local copy in shared memory (filled in the kernel prologue)
                return [f'kx_nasa_s[{off} + {base + i}]' for i in range(len(lo))], lo, hi
            return [f'kx_nasa_tab[{off} + {base + i}]' for i in range(len(lo))], lo, hi
        flag = self.tmid_flag(s.T_mid)
        out = []
        for a, b in zip(lo, hi):
            out.append(self.K(a) if a == b else f'({flag} ? {self.K(a)} : {self.K(b)})')
        return out, lo, hi

    def tmid_offset(self, tmid):
        name = 'noff_' + repr(float(tmid)).replace('.', '_').replace('-', 'm')
        if name not in self._flags:
            self._flags[name] = f'const int {name} = (T <= {_lit(float(tmid))}) ? 0 : KX_NASA_HALF;'
        return name

    def nasa_table_definition(self):
        if not getattr(self, 'nasa_indexed', False):
            return ''
        vals = self.nasa_lo_tab + self.nasa_hi_tab
        body = ',\n  '.join(', '.join(_lit(v) for v in vals[i:i + 4]) for i in range(0, len(vals), 4))
        qual = '__constant__' if self.nasa_indexed in (True, 'const') else '__device__ const __align__(16)'
        return (f'#define KX_NASA_HALF {len(self.nasa_lo_tab)}\n#define KX_NASA_LEN {len(vals)}\n'
                f'{qual} double kx_nasa_tab[{len(vals)}] = {{\n  {body}\n}};\n')

    def tmid_flag(self, tmid):
        name = 'lo_' + repr(float(tmid)).replace('.', '_').replace('-', 'm')
        if name not in self._flags:
            self._flags[name] = f'const bool {name} = T <= {_lit(float(tmid))};'
        return name

    # ---- Arrhenius ---------------------------------------------------------------------------
    def arrhenius_group_expr(self, A, b, Ta):
        """Expression for A T^b exp(-Ta/T) with the reference's literal special cases
        (reaction_rates.py:218-241)."""
        K = self.K
        if b == 0 and Ta == 0:
            return K(A)
        if Ta == 0 and b in (-2, -1, 1, 2):
            return {-2: f'{K(A)} * rcpT * rcpT', -1: f'{K(A)} * rcpT', 1: f'{K(A)} * T', 2: f'{K(A)} * T * T'}[b]
        if not A > 0:
            # the reference folds ln A too (reaction_rates.py:218-241) and fails on math.log; say why
            raise SystemExit(f'kinetix_b200: pre-exponential factor A = {A} is not positive; negative-A (or zero) '
                             f'Arrhenius terms are not supported (neither does the reference: it folds ln A)')
        lnA = math.log(A)
        if b == 0:
            lo, hi = _arg_range(lnA, 0, -Ta)
            return self.exp(f'fma({K(-Ta)}, rcpT, {K(lnA)})', lo, hi)
        if Ta == 0:
            lo, hi = _arg_range(lnA, b, 0)
            return self.exp(f'fma({K(b)}, lnT, {K(lnA)})', lo, hi)
        lo, hi = _arg_range(lnA, b, -Ta)
        return self.exp(f'fma({K(b)}, lnT, fma({K(-Ta)}, rcpT, {K(lnA)}))', lo, hi)

    def ratio_expr(self, rx):
        """k0/k_inf as the reference forms it (one exponential of differences, or the constant
        A0/A_inf -- reaction_rates.py:244-259).  Returns (expression, ln-argument expression or None)."""
        K = self.K
        A_inf, b_inf, E_inf = rx.rate.A, rx.rate.b, rx.rate.Ta
        A0, b0, E0 = rx.k0.A, rx.k0.b, rx.k0.Ta
        if (A0 - A_inf) != 0 and ((b0 - b_inf) != 0 or (E0 - E_inf) != 0):
            c0 = math.log(A0) - math.log(A_inf)
            cl = (b0 - b_inf)
            cr = (-E0 + E_inf)
            arg = K(c0)
            if cl != 0:
                arg = f'fma({K(cl)}, lnT, {arg})'
            if cr != 0:
                arg = f'fma({K(cr)}, rcpT, {arg})'
            lo, hi = _arg_range(c0, cl, cr)
            return arg, (lo, hi)
        return None, A0 / A_inf

    # ---- reaction scheduling -----------------------------------------------------------------
    def _units(self):
        """Scheduling units: reactions sharing (beta, Ta) of their main rate constant form one unit (they
        share one exp); everything else is a unit of one."""
        groups = {}
        for i, rx in enumerate(self.m.reactions):
            if rx.kind == 'P-log':
                key = ('plog', i)
            else:
                b, Ta = rx.rate.b, rx.rate.Ta
                trivial = (b == 0 and Ta == 0) or (Ta == 0 and b in (-2, -1, 1, 2))
                key = (b, Ta, i) if trivial else (b, Ta)
            groups.setdefault(key, []).append(i)
        return list(groups.values())

    def _species_of_unit(self, unit):
        s = set()
        for i in unit:
            rx = self.m.reactions[i]
            for k in range(self.N):
                if rx.nu_reac[k] or rx.nu_prod[k]:
                    s.add(k)
            if rx.efficiencies is None and rx.third_body_index >= 0 and rx.kind != 'elementary' \
                    and rx.kind != 'irreversible':
                s.add(rx.third_body_index)
        return s

    def _schedule(self, units, usp, reorder):
        """Order the units so that few species are 'live' (first use .. last use) at any time: greedy
        choice of the unit that opens the fewest new live ranges, closes the most and touches species
        close to retirement.  For GRI-3.0 the peak live set drops from 41 to 25 species (mean 18), which is
        what lets concentrations and rate accumulators of the live species stay in registers."""
        if not reorder:
            return list(range(len(units)))
        count = [0] * self.N
        for sp in usp:
            for k in sp:
                count[k] += 1
        rem = list(count)
        live, left, order = set(), set(range(len(units))), []
        while left:
            best = None
            for ui in left:
                new = len(usp[ui] - live)
                ret = sum(1 for k in usp[ui] if rem[k] == 1)
                close = sum(1.0 / rem[k] for k in usp[ui])
                cost = new - 1.0 * ret - 2.0 * close
                if best is None or cost < best[0] or (cost == best[0] and ui < best[1]):
                    best = (cost, ui)
            ui = best[1]
            order.append(ui)
            left.discard(ui)
            live |= usp[ui]
            for k in usp[ui]:
                rem[k] -= 1
                if rem[k] == 0:
                    live.discard(k)
        return order

    # ---- main --------------------------------------------------------------------------------
    def emit(self, kernel_name='kx_bk1_f64', block=128, min_blocks=2, sync_every=8, gibbs_in_smem=True,
             reorder=True, prefetch=4, ring=0, pin_loads=False, l1_keep=False, keep_until=0, live_cap=0, eff_in_smem=True, nasa_indexed=False,
             tmem_slots=0, smem_cap=0, tmem_cols=512, cold_uses=0, cold_slot_cap=0, routine=False, kbase_ahead=0,
             cold_conc_only=False, gibbs_prefer_tm=False):
        """block / min_blocks: launch bounds.
        routine: emit the reference-signature DEVICE FUNCTION `kinetix_species_rates(lnT, T, T2, T3, T4, rcpT, P, lnP,
          Ci, wdot)` (reference reaction_rates.py:560-562) instead of the kernel: concentrations come from `Ci[]`,
          rates are ADDED to `wdot[]` (the caller pre-zeroes it, productionRates.okl:40), no state rows, no scratch
          slots, no heat release.  Same minimal-form arithmetic and liveness schedule (see emit_routines.py).
        cold_uses / cold_slot_cap (experimental, off by default): species that occur in at most `cold_uses` reactions
          keep their concentration C_k and rate accumulator wdot_k in two SHARED-MEMORY slots instead of registers
          while they are live, as long as the thread's slots in use stay below `cold_slot_cap`: explicit placement of
          the values ptxas would otherwise spill to local memory (heptaneLu88: 88 species, 29 slots, 2.7 KB of spill
          loads at 168 registers while two thirds of its shared-memory budget are idle).
        cold_conc_only: the cold placement moves only the (read-only) concentration C_k to a slot; the accumulator
          wdot_k, the read-modify-write half, stays in a register.  gibbs_prefer_tm: exp(+-g_k) take tensor-memory
          slots first, leaving the shared-memory slots to the concentrations.
        tmem_slots: > 0: that many doubles per thread of TENSOR MEMORY hold scratch slots beside at most `smem_cap`
          shared-memory slots (large mechanisms: EtOHKonnov needs 210 slots = 1.7 KB per thread, which limits a
          shared-memory-only kernel to 4 warps per SM).  Third-body sums go to tensor memory first, exp(+-g_k) to
          shared memory while it lasts.  The kernel then allocates all 512 TMEM columns (one CTA per SM).
        sync_every: a CTA-wide barrier every that many reactions keeps the warps of a CTA inside the same
          window of the straight-line code so instruction-cache fills are shared (0 = none).
          (Measured and removed: named barriers of the warps that share a SCHEDULER only, in one 384-thread CTA, so that
          they find each other's lines in that scheduler's L0 instruction cache.  It does what it says -- no_instruction
          1.10 -> 0.57 stall cycles per issue, I-cache hit rate 60 -> 72 % -- and loses: warps at the same place of the code
          wait for the same dependent chains at the same time (wait 1.78 -> 2.10, long_scoreboard 0.79 -> 1.12): 854 / 883 /
          905 / 933 M states/s for a barrier every 2 / 4 / 8 / 16 reactions against 962 for three unsynchronised CTAs.)
        gibbs_in_smem: exp(+-g_k) of live species in shared memory slots [slot][thread] (slots are recycled
          when a species retires) instead of registers.
        reorder: liveness-minimising reaction order (see _schedule).
        live_cap: > 0 limits the number of simultaneously live species (large mechanisms): when a unit would
          exceed it, the live species whose next use is farthest away is suspended after its latest use -- its
          partial M_k*wdot_k is written (first time) or added (later) to its own rate row, its heat release
          added, its shared-memory slots recycled -- and re-activated from the state row at its next use.
          This is live-range splitting with the (thread-private, L2-resident) output row as the spill slot.
        prefetch: (ring == 0) the state row of a species is re-loaded this many ACTIVATIONS before its own.
        ring: > 0: state rows are re-fetched with cp.async (LDGSTS) into a ring of that many shared-memory
          slots per thread, one commit group per species in activation order; an activation waits with
          cp.async.wait_group for ITS group only (FIFO), whereas a register load would wait on a counting
          scoreboard shared with younger prefetches."""
        m, N, K = self.m, self.N, self.K
        self.block, self.sync_every, self.gibbs_in_smem = block, sync_every, gibbs_in_smem
        self.nasa_indexed, self.nasa_lo_tab, self.nasa_hi_tab = nasa_indexed, [], []
        # keep_until: species whose first reaction comes at or before this schedule position keep their pass-1
        # value Y_k/M_k in a register until activation (no second read of the state row); early in the
        # schedule few species are live, so this does not raise the peak register demand
        self.keep_until = keep_until
        # l1_keep: pass-1 loads of the state rows allocate in L1 (evict_last) so the reload at activation can hit
        self.ld1 = 'kx_ld_keep' if l1_keep else 'kx_ld_stream'
        self.ld2 = 'kx_ld_keep' if l1_keep else 'kx_ld_stream'
        self._flags = {}
        self._tm_dirty, self._fetched = False, set()
        self.tmem_slots = tmem_slots
        body = []
        self.lines = body
        w = self.w

        units = self._units()
        usp = [self._species_of_unit(u) for u in units]
        order = self._schedule(units, usp, reorder)
        first, last = {}, {}
        uses = {}
        for pos, ui in enumerate(order):
            for k in usp[ui]:
                first.setdefault(k, pos)
                last[k] = pos
                uses.setdefault(k, []).append(pos)
        self.schedule_stats = dict(units=len(units))
        # live segments per species: [(start_pos, end_pos)], one segment unless live_cap forces suspensions
        segments = {k: [[first[k], last[k]]] for k in first}
        if live_cap:
            import bisect
            live_set, seg_start, last_use = set(), {}, {}
            segments = {k: [] for k in first}
            for pos, ui in enumerate(order):
                need = usp[ui]
                for k in need:
                    if k not in live_set:
                        live_set.add(k)
                        seg_start[k] = pos
                # suspend the species with the farthest next use while over capacity
                while len(live_set) > live_cap:
                    cand = [k for k in live_set if k not in need]
                    if not cand:
                        break

                    def next_use(k):
                        i = bisect.bisect_right(uses[k], pos)
                        return uses[k][i] if i < len(uses[k]) else 10 ** 9
                    victim = max(cand, key=next_use)
                    live_set.discard(victim)
                    segments[victim].append([seg_start[victim], last_use[victim]])
                for k in need:
                    last_use[k] = pos
                    if pos == last[k]:
                        live_set.discard(k)
                        segments[k].append([seg_start[k], pos])
            n_susp = sum(len(v) - 1 for v in segments.values())
            self.schedule_stats['suspensions'] = n_susp

        # which exp(+-g_k) are needed
        need_pos = [False] * N   # exp(+g): species is a net product of a reversible reaction
        need_neg = [False] * N   # exp(-g): net reactant
        for rx in m.reactions:
            if rx.reversible:
                for k, v in enumerate(rx.nu_net):
                    if v > 0:
                        need_pos[k] = True
                    elif v < 0:
                        need_neg[k] = True

        # ---- pass 1 over the state rows: mean molar mass, density, third-body sums -------------------
        # (productionRates.okl:11-41).  Cm = sum_k C_k = rho * sum_k Y_k/M_k; an efficiency vector is
        # Cm + sum_k (eps_k - 1) C_k (reaction_rates.py:298-301); distinct vectors are evaluated once.
        eff_names = {}
        for rx in m.reactions:
            if rx.efficiencies is not None and tuple(rx.efficiencies) not in eff_names:
                eff_names[tuple(rx.efficiencies)] = f'M{len(eff_names)}'
        # shared-memory slot allocator ([slot][thread] doubles); the Y ring (if any) takes the first slots
        self.eg_slot, self.rg_slot = {}, {}
        free_slots, n_slots = [], (ring if ring else 0)
        free_tm, n_tm = [], 0

        def take_slot(prefer_tm=False, smem_only=False):
            nonlocal n_slots, n_tm
            for pool in (('s',) if smem_only else ('t', 's') if prefer_tm else ('s', 't')):
                if pool == 's':
                    if free_slots:
                        return free_slots.pop(0)
                    if not tmem_slots or n_slots < smem_cap:
                        n_slots += 1
                        return n_slots - 1
                elif tmem_slots:
                    if free_tm:
                        return free_tm.pop(0)
                    if n_tm < tmem_slots:
                        n_tm += 1
                        return self.TM_BASE + n_tm - 1
            if smem_only:
                return None
            raise RuntimeError(f'BK1 emitter: out of scratch slots ({smem_cap} shared + {tmem_slots} tensor memory)')

        def release_slot(slot):
            (free_tm if self.is_tm(slot) else free_slots).append(slot)

        if gibbs_in_smem or ring:
            w('extern __shared__ double kx_sm[];')
            w('double* const gs = kx_sm + threadIdx.x;')
        if tmem_slots:
            # this thread's tensor-memory columns: lane quadrant of the warp, column block of the warp group
            w('__shared__ unsigned kx_tm_slot;')
            w(f'const unsigned tm_alloc = kx_tm_alloc_all<{tmem_cols}>(&kx_tm_slot);')
            w(f'const unsigned tmb = tm_alloc + (((threadIdx.x >> 5) & 3u) << 21) + (threadIdx.x >> 7) * {2 * tmem_slots}u;')
        if not routine:
            w('#ifdef KX_EXP_TABLE')
            w('kx_exptab_init();')
            w('#endif')
        if nasa_indexed == 'smem':
            w('__shared__ double kx_nasa_s[KX_NASA_LEN];')
            w('for (int i = threadIdx.x; i < KX_NASA_LEN; i += blockDim.x) kx_nasa_s[i] = kx_nasa_tab[i];')
            w('__syncthreads();')
        # third-body sums M_i (and ln M_i) are needed all along the reaction list: in shared memory they do not
        # occupy 2 registers each for the whole kernel (GRI-3.0: 10 + 5 values, EtOHKonnov: 30 + 15)
        eff_smem = bool(eff_in_smem and gibbs_in_smem)
        eff_slot = {}
        for name in eff_names.values():
            if eff_smem:
                eff_slot[name] = take_slot(prefer_tm=True)
        self.routine = routine
        if routine:
            assert not (gibbs_in_smem or ring or tmem_slots or live_cap or keep_until)
            w('const double Pv = P, lnPv = lnP;')
            w('(void)T2; (void)T3; (void)T4; (void)Pv; (void)lnPv;')
        else:
            w('const double T = Tref * kx_ld_stream(state + id);')
            w('const double rcpT = kx_rcp(T);')
            w('const double lnT = kx_log(T);')
        w('double rho, Cm;')
        if not eff_smem:
            for name in eff_names.values():
                w(f'double {name};')

        def collider(rx):
            """name of the third-body concentration of a reaction: an efficiency-vector sum 'M<i>', a single
            species 'cs<k>', or the total concentration 'Cm'"""
            if rx.efficiencies is not None:
                return eff_names[tuple(rx.efficiencies)]
            if rx.third_body_index >= 0:
                return f'cs{rx.third_body_index}'
            return 'Cm'

        # ln(M) once per distinct collider of a Troe reaction whose Pr is formed as exp(.)*M
        # (guarded like the reference's log10(Pr + CFLOAT_MIN))
        need_ln = []
        for rx in m.reactions:
            if rx.kind == 'Troe':
                name = collider(rx)
                arg, _ = self.ratio_expr(rx)
                if not name.startswith('cs') and arg is not None and name not in need_ln:
                    need_ln.append(name)
        if not eff_smem:
            for name in need_ln:
                w(f'double ln_{name};')
        kept = set(k for k in first if first[k] <= keep_until) if keep_until else set()
        self.kept = kept
        if kept:
            w('double ' + ', '.join(f'w{k}' for k in sorted(kept)) + ';')
        w('{')
        w('  double rcpMbar = 0.0;')
        for name in eff_names.values():
            w(f'  double a{name} = 0.0;')
        for k in range(N):
            if routine:      # w_k = C_k, rho = 1: Cm = sum_k C_k over ALL species (reaction_rates.py:579)
                w(f'  const double w{k} = Ci[{k}]; rcpMbar += w{k};')
            else:
                w(f'  {"" if k in kept else "const double "}w{k} = fmax(0.0, kx_ld_row<{k}>(sp, offset)) * '
                  f'{K(1. / m.species[k].M)}; rcpMbar += w{k};')
            for vec, name in eff_names.items():
                if vec[k] != 1:
                    w(f'  a{name} = fma({K(vec[k] - 1)}, w{k}, a{name});')
        if routine:
            w('  rho = 1.0;')
            w('  Cm = rcpMbar;')
        else:
            w('  rho = pR * rcpT * kx_rcp(rcpMbar);')
            w('  Cm = rho * rcpMbar;')
        ln_slot, ln_collider = {}, {}
        for name in eff_names.values():
            if eff_smem:
                w(f'  a{name} = fma(rho, a{name}, Cm);')
                self.store(eff_slot[name], f'a{name}', '  ')
            else:
                w(f'  {name} = fma(rho, a{name}, Cm);')
        for name in need_ln:
            src = ('a' + name) if (eff_smem and name != 'Cm') else name
            if eff_smem:
                ln_slot[name] = take_slot(prefer_tm=True)
                self.store(ln_slot[name], f'kx_log(fmax({src}, 1e-300))', '  ')
            else:
                ln_collider[name] = 'ln_' + name
                w(f'  ln_{name} = kx_log(fmax({src}, 1e-300));')
            self.stats['log'] += 1
        w('}')
        w(f'const double C0 = {K(const.ONE_ATM / const.R_GAS)} * rcpT;')
        w(f'const double rcpC0 = {K(const.R_GAS / const.ONE_ATM)} * T;')
        flag_pos = len(body)

        def collider_slots(rx):
            """scratch slots a reaction reads for its third body (value, ln value)"""
            out = []
            name = collider(rx)
            if rx.kind in ('three-body', 'pressure-modification', 'Troe', 'SRI'):
                if name in eff_slot:
                    out.append(eff_slot[name])
                if rx.kind == 'Troe' and name in ln_slot and self.ratio_expr(rx)[0] is not None:
                    out.append(ln_slot[name])
            return out

        def collider_val(rx):
            name = collider(rx)
            if name.startswith('cs'):
                return CS(int(name[2:]))
            return self.val(eff_slot[name]) if name in eff_slot else name

        def ln_collider_val(rx):
            """expression of ln(M) if it was precomputed for this reaction's collider, else None"""
            name = collider(rx)
            if name in ln_slot:
                return self.val(ln_slot[name])
            return ln_collider.get(name)

        # ---- per-species live state --------------------------------------------------------------
        used = sorted(first)                                  # species that occur in some reaction
        # cold species: C_k / wdot_k in shared-memory slots while live (see cold_uses above)
        mem_cs, mem_wd, seg_of = {}, {}, {}
        self.cold_activations = 0

        def CS(k):
            return f'gs[{mem_cs[k]} * {block}]' if k in mem_cs else f'cs{k}'

        def WD(k):
            return f'gs[{mem_wd[k]} * {block}]' if k in mem_wd else f'wd{k}'

        # Which live segments of which species keep C_k / wdot_k in shared-memory slots ("cold" placement): planned on
        # the slot-usage timeline of the schedule, so that the slots EVER allocated stay within the cap (the CTA's
        # shared memory, hence the CTAs per SM, must not change: GRI-3.0 with an in-use-only check grew from 63 to 79
        # slots, lost its third CTA per SM and ran at 676 instead of 962 M states/s).  A species' exp(+-g) slots are
        # busy from its activation to its retirement; a cold segment adds two more over the same span.
        cold_plan = set()
        n_cold = 1 if cold_conc_only else 2
        if tmem_slots and cold_conc_only:
            cold_slot_cap = smem_cap + tmem_slots - 2
        if cold_uses and gibbs_in_smem and (not tmem_slots or cold_conc_only):
            n_pos = len(order)
            usage = [len(eff_slot) + len([n for n in need_ln if eff_smem])] * (n_pos + 1)
            for k, segs in segments.items():
                for a, b in segs:
                    for pos in range(a, b + 1):
                        usage[pos] += int(need_pos[k]) + int(need_neg[k])
            for k in sorted(segments, key=lambda k: segments[k][0][0]):
                if len(uses[k]) > cold_uses:
                    continue
                for si, (a, b) in enumerate(segments[k]):
                    if max(usage[a:b + 1]) + n_cold <= cold_slot_cap:
                        cold_plan.add((k, si))
                        for pos in range(a, b + 1):
                            usage[pos] += n_cold

        def place(k, segment=0):
            """decide where species k lives for this live segment"""
            if not (cold_uses and gibbs_in_smem) or len(uses[k]) > cold_uses:
                return
            if tmem_slots and not cold_conc_only:
                # tensor-memory layout: shared-memory slots are capped by smem_cap, in-use check with head-room for the
                # exp(+-g) of the next activations (the overflow goes to tensor memory, not to a bigger CTA)
                if (n_slots - len(free_slots)) + 2 + 6 > smem_cap:
                    return
            elif (k, segment) not in cold_plan:
                return
            if cold_conc_only:
                a = take_slot(smem_only=True)
                if a is not None:
                    mem_cs[k] = a
                    self.cold_activations += 1
                return
            a, b = take_slot(smem_only=True), take_slot(smem_only=True)
            if a is None or b is None:
                for t in (a, b):
                    if t is not None:
                        release_slot(t)
                return
            mem_cs[k], mem_wd[k] = a, b
            self.cold_activations += 1
        w('double ' + ', '.join(f'cs{k}' for k in used) + ';')
        w('double ' + ', '.join(f'wd{k}' for k in used) + ';')
        if not ring and not routine:
            w('double ' + ', '.join(f'y{k}' for k in used) + ';')
        if not routine:
            w('double hsum = 0.0;')
        act_order = sorted(used, key=lambda k: (first[k], k))      # activation order
        rank = {k: i for i, k in enumerate(act_order)}
        if ring:
            w('const unsigned ring_base = (unsigned)__cvta_generic_to_shared(gs);')
            for k in act_order[:ring]:
                w(f'kx_cp_async8(ring_base + {rank[k] % ring} * {block} * 8, sp + {k} * offset);')
        if not gibbs_in_smem:
            names = [f'eg{k}' for k in used if need_pos[k]] + [f'rg{k}' for k in used if need_neg[k]]
            if names:
                w('double ' + ', '.join(names) + ';')

        def gcoef(a):   # g/RT = b0 + b1 lnT + b6/T + T (b2 + T (b3 + T (b4 + T b5)))   (reaction_rates.py:565-569)
            return [a[0] - a[6], -a[0], -a[1] / 2, (1. / 3. - 1. / 2.) * a[2], (1. / 4. - 1. / 3.) * a[3],
                    (1. / 5. - 1. / 4.) * a[4], a[5]]

        def hcoef(a):   # h/RT (thermodynamics.py:74-75)
            return [a[0], a[1] / 2, a[2] / 3, a[3] / 4, a[4] / 5, a[5]]

        def activate(k, reactivation=False):
            """species k becomes live: concentration, zeroed accumulator, exp(+-g_k/RT)"""
            place(k, seg_of.get(k, 0))
            if routine:
                w(f'{CS(k)} = Ci[{k}]; {WD(k)} = 0.0;')
            elif k in self.kept and not reactivation:
                w(f'{CS(k)} = w{k} * rho; {WD(k)} = 0.0;')
            elif ring and not reactivation:
                r = rank[k]
                pending = max(0, min(ring - 1, len(act_order) - 1 - r))
                w(f'kx_cp_async_wait<{pending}>();')
                w(f'const double y{k} = kx_ring_read(ring_base + {r % ring} * {block} * 8);')
                if r + ring < len(act_order):
                    nk = act_order[r + ring]
                    w(f'kx_cp_async8(ring_base + {r % ring} * {block} * 8, sp + {nk} * offset);')
            if routine or (k in self.kept and not reactivation):
                pass
            elif reactivation:
                w(f'{CS(k)} = fmax(0.0, y{k}) * ({K(1. / m.species[k].M)} * rho); {WD(k)} = 0.0;')
            elif not ring and pin_loads:
                # volatile max: keeps this activation ordered after the loads issued `prefetch` activations
                # ahead (volatile asms are not reordered among themselves), so the compiler cannot sink those
                # loads down to their first use
                w(f'{{ double t; asm volatile("max.f64 %0, %1, 0d0000000000000000;" : "=d"(t) : "d"(y{k})); '
                  f'{CS(k)} = t * ({K(1. / m.species[k].M)} * rho); {WD(k)} = 0.0; }}')
            else:
                w(f'{CS(k)} = fmax(0.0, y{k}) * ({K(1. / m.species[k].M)} * rho); {WD(k)} = 0.0;')
            if not (need_pos[k] or need_neg[k]):
                return
            c, _, _ = self.nasa_select(k, gcoef)
            s = m.species[k]
            gmin, gmax = float('inf'), float('-inf')
            for i in range(200):
                t = T_VALID_LO + (T_VALID_HI - T_VALID_LO) * i / 199
                b = gcoef(s.nasa_lo if t <= s.T_mid else s.nasa_hi)
                g = b[0] + b[1] * math.log(t) + b[6] / t + t * (b[2] + t * (b[3] + t * (b[4] + t * b[5])))
                gmin, gmax = min(gmin, g), max(gmax, g)
            glo, ghi = min(gmin * 1.05, gmin * 0.95) - 5, max(gmax * 1.05, gmax * 0.95) + 5
            if need_pos[k]:
                self.eg_slot[k] = take_slot(prefer_tm=gibbs_prefer_tm) if gibbs_in_smem else None
            if need_neg[k]:
                self.rg_slot[k] = take_slot(prefer_tm=gibbs_prefer_tm) if gibbs_in_smem else None
            w('{')
            w(f'  const double g = fma(fma(fma(fma({c[5]}, T, {c[4]}), T, {c[3]}), T, {c[2]}), T, '
              f'fma({c[1]}, lnT, fma({c[6]}, rcpT, {c[0]})));')
            def put(slots, name, expr):
                if gibbs_in_smem:
                    self.store(slots[k], expr, '  ')
                else:
                    w(f'  {name}{k} = {expr};')
            if need_pos[k]:
                w(f'  const double e = {self.exp("g", glo, ghi)};')
                put(self.eg_slot, 'eg', 'e')
                if need_neg[k]:
                    if _exp_fn(glo, ghi) == 'kx_exp_wide':
                        # e may be inf or 0 here; the Newton reciprocal would turn that into NaN, a second full-range
                        # exp saturates to 0 / inf the way the reference's single exp(sum nu g) does
                        put(self.rg_slot, 'rg', self.exp('-g', -ghi, -glo))
                    else:
                        put(self.rg_slot, 'rg', 'kx_rcp(e)')
                        self.stats['rcp'] += 1
            else:
                put(self.rg_slot, 'rg', self.exp("-g", -ghi, -glo))
            w('}')

        def retire(k, first_flush=True):
            """species k leaves the live set (for good, or suspended under live_cap): write / add its rate row,
            add its heat release, free its slots (productionRates.okl:48-62)"""
            if routine:
                w(f'wdot[{k}] += {WD(k)};')
                return
            c, _, _ = self.nasa_select(k, hcoef)
            if first_flush:
                w(f'if (live) kx_st_row<{k}>(out, offset, {K(m.species[k].M)} * {WD(k)});')
            else:
                w(f'if (live) kx_add_row<{k}>(out, offset, {K(m.species[k].M)} * {WD(k)});')
            w(f'hsum = fma({WD(k)}, fma(fma(fma(fma({c[4]}, T, {c[3]}), T, {c[2]}), T, {c[1]}), T, '
              f'fma({c[5]}, rcpT, {c[0]})), hsum);')
            for slots in (self.eg_slot, self.rg_slot):
                if gibbs_in_smem and k in slots:
                    release_slot(slots[k])
            for slots in (mem_cs, mem_wd):
                if k in slots:
                    release_slot(slots.pop(k))

        def conc_product(nu):
            terms = []
            for k, c in enumerate(nu):
                terms += [CS(k)] * c
            return ' * '.join(terms)

        # species that never occur in a reaction: rate row is zero (the reference's wdot[k] stays 0)
        for k in range(N):
            if k not in first and not routine:
                w(f'if (live) kx_st_row<{k}>(out, offset, 0.0);')

        # ---- units in schedule order ---------------------------------------------------------------
        by_first, by_last = {}, {}
        seg_index = {}
        for k, segs in segments.items():
            for si, (a, b) in enumerate(segs):
                by_first.setdefault(a, []).append(k)
                by_last.setdefault(b, []).append(k)
                seg_index[(k, a)] = si
                seg_index[(k, 'end', b)] = si
        loaded = set()

        def issue_loads(upto_rank):
            if ring or routine:
                return
            for k in act_order[:upto_rank + 1]:
                if k not in loaded and k not in self.kept:
                    loaded.add(k)
                    w(f'y{k} = kx_ld_row<{k}>(sp, offset);')

        emitted = 0
        peak_live, live_now = 0, 0

        def kbase_expr(pos):
            rx0 = m.reactions[units[order[pos]][0]]
            return None if rx0.kind == 'P-log' else self.arrhenius_group_expr(rx0.rate.A, rx0.rate.b, rx0.rate.Ta)

        # kbase_ahead = A > 0: the rate constant of a unit depends on T only, so it is written A units BEFORE the unit
        # that uses it -- its exp chain (16 dependent FP64 instructions) then sits next to independent arithmetic in
        # the source, instead of relying on the scheduler to look that far (2 registers per constant in flight)
        if kbase_ahead:
            for pos in range(min(kbase_ahead, len(order))):
                e = kbase_expr(pos)
                if e is not None:
                    w(f'double kb_{pos} = {e};')
        for pos, ui in enumerate(order):
            members = units[ui]
            if kbase_ahead and pos + kbase_ahead < len(order):
                e = kbase_expr(pos + kbase_ahead)
                if e is not None:
                    w(f'const double kb_{pos + kbase_ahead} = {e};')
            for k in sorted(by_first.get(pos, [])):
                seg_of[k] = seg_index[(k, pos)]
                if seg_index[(k, pos)] == 0:
                    issue_loads(rank[k] + prefetch)
                else:                                   # re-activation after a suspension: fresh load
                    w(f'y{k} = kx_ld_row<{k}>(sp, offset);')
                activate(k, reactivation=seg_index[(k, pos)] > 0)
                live_now += 1
            peak_live = max(peak_live, live_now)
            if sync_every and emitted and emitted // sync_every != (emitted + len(members)) // sync_every:
                w('__syncthreads();')
            emitted += len(members)
            first_rx = m.reactions[members[0]]
            w(f'// ---- unit {pos}: reactions ' + ', '.join(str(i + 1) for i in members))
            w('{')
            if first_rx.kind != 'P-log':
                if kbase_ahead:
                    w(f'  const double kbase = kb_{pos};')
                else:
                    w(f'  const double kbase = '
                      f'{self.arrhenius_group_expr(first_rx.rate.A, first_rx.rate.b, first_rx.rate.Ta)};')
            for i in members:
                rx = m.reactions[i]
                w(f'  // {i + 1}: {rx.equation}')
                w('  {')
                if gibbs_in_smem:
                    need = collider_slots(rx)
                    if rx.reversible:
                        need += [self.eg_slot[k] for k, v in enumerate(rx.nu_net) if v > 0]
                        need += [self.rg_slot[k] for k, v in enumerate(rx.nu_net) if v < 0]
                    self.fetch(need, '    ')
                if rx.kind == 'P-log':
                    self._emit_plog(rx)
                elif i == members[0]:
                    w('    double kf = kbase;')
                else:
                    w(f'    double kf = kbase * {K(rx.rate.A / first_rx.rate.A)};')

                if rx.kind == 'three-body':
                    w(f'    kf *= {collider_val(rx)};')
                elif rx.kind in ('pressure-modification', 'Troe', 'SRI'):
                    M = collider_val(rx)
                    arg, rng = self.ratio_expr(rx)
                    if arg is not None:
                        w(f'    const double lnr = {arg};')
                        w(f'    const double Pr = {self.exp("lnr", *rng)} * {M};')
                    else:
                        w(f'    const double Pr = {K(rng)} * {M};')
                    w('    const double rcp1Pr = kx_rcp(1.0 + Pr);')
                    self.stats['rcp'] += 1
                    if rx.kind == 'pressure-modification':
                        w('    kf *= Pr * rcp1Pr;')
                    elif rx.kind == 'Troe':
                        # log10(Pr + CFLOAT_MIN): from the exponent and ln(M) when M is a sum of
                        # concentrations; literally when the collider is a single species (can be 0)
                        if arg is not None and ln_collider_val(rx) is not None:
                            w(f'    const double logPr = (lnr + {ln_collider_val(rx)}) * {K(1 / math.log(10))};')
                        else:
                            w('    const double logPr = kx_log10(Pr + 1e-300);')
                            self.stats['log'] += 1
                        w(f'    const double Fc = {self._troe_fcent(rx.troe)};')
                        w('    const double lnFc = kx_log(Fc);')
                        self.stats['log'] += 1
                        w(f'    const double logFc = lnFc * {K(1 / math.log(10))};')
                        w('    const double tc = fma(-0.67, logFc, -0.4) + logPr;')
                        w('    const double tn = fma(-1.27, logFc, 0.75) - 0.14 * tc;')
                        # F = 10^(logFc/(1+(tc/tn)^2)) = exp(lnFc * tn^2/(tn^2+tc^2))
                        w('    const double tn2 = tn * tn;')
                        w('    const double F = kx_exp(lnFc * tn2 * kx_rcp(fma(tc, tc, tn2)));')
                        self.stats['exp'] += 1
                        self.stats['rcp'] += 1
                        w('    kf *= Pr * rcp1Pr * F;')
                    else:  # SRI (reaction_rates.py:346-357)
                        sr = rx.sri
                        w('    const double logPr = kx_log10(Pr);')
                        w(f'    const double sb = {K(sr["A"])} * kx_exp_wide({K(-sr["B"])} * rcpT) + '
                          f'kx_exp_wide({K(-1. / (sr["C"] + const.FLOAT_MIN))} * T);')
                        w(f'    double F = {K(sr["D"])} * kx_pow(sb, kx_rcp(fma(logPr, logPr, 1.0)));')
                        if sr['E'] != 0:
                            w(f'    F *= kx_exp({K(sr["E"])} * lnT);')
                        self.stats['log'] += 2
                        self.stats['exp'] += 3
                        w('    kf *= Pr * rcp1Pr * F;')

                Rf = conc_product(rx.nu_reac)
                net = rx.nu_net
                if not rx.reversible:
                    w(f'    const double q = kf * {Rf};')
                else:
                    # 1/Kc = prod exp(g)^nu * C0^(-sum nu); interleave +g / -g factors so partial
                    # products stay O(exp(delta g)) (no intermediate over/underflow)
                    pos_f = [self.EG(k) for k, v in enumerate(net) if v > 0 for _ in range(v)]
                    neg_f = [self.RG(k) for k, v in enumerate(net) if v < 0 for _ in range(-v)]
                    factors = []
                    while pos_f or neg_f:
                        if pos_f:
                            factors.append(pos_f.pop(0))
                        if neg_f:
                            factors.append(neg_f.pop(0))
                    sn = sum(net)
                    factors += ['C0' if sn < 0 else 'rcpC0'] * abs(sn)
                    w(f'    const double kr = {" * ".join(factors)};')
                    Rr = conc_product(rx.nu_prod)
                    w(f'    const double q = kf * fma(-kr, {Rr}, {Rf});')
                for k, v in enumerate(net):
                    if v == 1:
                        w(f'    {WD(k)} += q;')
                    elif v == -1:
                        w(f'    {WD(k)} -= q;')
                    elif v != 0:
                        w(f'    {WD(k)} = fma({float(v)}, q, {WD(k)});')
                w('  }')
            w('}')
            for k in sorted(by_last.get(pos, [])):
                retire(k, first_flush=seg_index[(k, 'end', pos)] == 0)
                live_now -= 1
        if not routine:
            w(f'if (live) kx_st_stream(rates + id, {K(-const.R_GAS)} * T * hsum);')
        if tmem_slots:
            w(f'kx_tm_free_all<{tmem_cols}>(tm_alloc);')

        self.smem_doubles_per_thread = n_slots if (gibbs_in_smem or ring) else 0
        self.schedule_stats.update(peak_live=peak_live, smem_slots=n_slots, tmem_slots=n_tm)
        if cold_uses:
            self.schedule_stats['cold_activations'] = self.cold_activations
        body[flag_pos:flag_pos] = ['  ' + v for v in self._flags.values()]

        if routine:
            head = [
                f'// kinetix_species_rates: {m.name}, {N} species / {m.n_reactions} reactions; {self.stats["exp"]} kx_exp, '
                f'{self.stats["exp_wide"]} wide exp, {self.stats["log"]} log, {self.stats["rcp"]} rcp per call; '
                f'peak live species {peak_live}',
                '__KINETIX_DEVICE__ __KINETIX_INLINE__ void kinetix_species_rates(const cfloat lnT, const cfloat T, '
                'const cfloat T2, const cfloat T3, const cfloat T4, const cfloat rcpT, const cfloat P, const cfloat lnP, '
                'const cfloat* Ci, cfloat* wdot)',
                '{',
            ]
            return '\n'.join(head + body + ['}', ''])
        head = [
            f'// BK1 (species production rates): {m.name}, {N} species / {m.n_reactions} reactions; '
            f'{self.stats["exp"]} kx_exp, {self.stats["exp_wide"]} wide exp, {self.stats["log"]} log, '
            f'{self.stats["rcp"]} rcp per state; peak live species {peak_live}, {n_slots} smem slots' +
            (f', {n_tm} tensor-memory slots ({self.stats.get("tm_ld", 0)} loads)' if tmem_slots else ''),
            '// PF: per-state pressure field (extension); the reference flavour PF = false carries no trace of it',
            'template <bool PF>',
            f'__global__ void __launch_bounds__({block}, {min_blocks})',
            f'{kernel_name}(const long long n_states, const long long offsetT, const long long offset,',
            '           const double pressure_R, const double P, const double lnP,',
            '           const double* __restrict__ state, double* __restrict__ rates, const double Tref,',
            '           const double* __restrict__ pfield' +
            (', const __grid_constant__ KxParamPool kp)' if getattr(K, 'as_param', False) else ')'),
            '{',
            '  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;',
            '  const bool live = gid < n_states;',
            '  const long long id = live ? gid : n_states - 1;   // tail threads recompute the last state, store nothing',
            '  const double* sp = state + id + offsetT;',
            '  double* out = rates + id + offsetT;',
            '  // per-state pressure (extension; the reference has one pressure per launch, kinetix.cpp:802-812):',
            '  // pfield[id] = p / p_ref of this state, the scalar arguments then carry p_ref',
            '  double pR = pressure_R, Pv = P, lnPv = lnP;',
            '  if (PF) {',
            '    const double pn = kx_ld_stream(pfield + id);',
            '    pR *= pn;',
            '    Pv *= pn;',
            f'    lnPv = {"kx_log(Pv)" if any(r.kind == "P-log" for r in m.reactions) else "lnP"};',
            '  }',
        ]
        return '\n'.join(head + body + ['}', ''])

    # ------------------------------------------------------------------------------------------
    def _troe_fcent(self, tr):
        """(1-A) exp(-T/T3) + A exp(-T/T1) [+ exp(-T2/T)] with the reference's A in {0,1} special
        cases (reaction_rates.py:331-340); terms that are exactly 0 or 1 in double over the whole
        validity range of T are folded."""
        K = self.K
        terms = []

        def texp(coef_T, weight):
            # weight * exp(coef_T * T)
            lo, hi = sorted((coef_T * T_VALID_LO, coef_T * T_VALID_HI))
            if hi < -750.0:
                return None                       # exactly 0
            if abs(lo) < 1e-17 and abs(hi) < 1e-17:
                return K(weight)                  # exp() == 1 exactly
            e = self.exp(f'{K(coef_T)} * T', lo, hi)
            return e if weight == 1 else f'{K(weight)} * {e}'

        A = tr['A']
        if A == 0:
            parts = [texp(-1. / (tr['T3'] + const.FLOAT_MIN), 1.0)]
        elif A == 1:
            parts = [texp(-1. / (tr['T1'] + const.FLOAT_MIN), 1.0)]
        else:
            parts = [texp(-1. / (tr['T3'] + const.FLOAT_MIN), 1 - A), texp(-1. / (tr['T1'] + const.FLOAT_MIN), A)]
        if tr['T2'] < float('inf'):
            lo, hi = sorted((-tr['T2'] / T_VALID_LO, -tr['T2'] / T_VALID_HI))
            parts.append(self.exp(f'{K(-tr["T2"])} * rcpT', lo, hi))
        parts = [p for p in parts if p is not None]
        return ' + '.join(parts) if parts else '0.0'

    def _emit_plog(self, rx):
        """Pressure-dependent Arrhenius: ln k linear in ln P between tabulated pressures; the pressure
        is uniform over a launch so the branch chain is warp-uniform (reaction_rates.py:358-388)."""
        K, w = self.K, self.w

        def ksum(ks):
            return ' + '.join(self.arrhenius_group_expr(k.A, k.b, k.Ta) for k in ks)

        pl = rx.plog
        n = len(pl)
        if n == 1:                                   # a single tabulated pressure: no interpolation, one rate
            w(f'    double kf = {ksum(pl[0][1])};')
            return
        w('    double kf;')
        for i in range(n - 1):
            (p1, k1), (p2, k2) = pl[i], pl[i + 1]
            lnp1, lnp2 = math.log(p1), math.log(p2)
            w(f'    {"if" if i == 0 else "} else if"} ((Pv > {_lit(p1)}) && (Pv < {_lit(p2)})) {{')
            w(f'      const double l1 = kx_log({ksum(k1)}), l2 = kx_log({ksum(k2)});')
            w(f'      kf = kx_exp(fma((l2 - l1) * (lnPv - {K(lnp1)}), {K(1 / (lnp2 - lnp1))}, l1));')
            self.stats['log'] += 2
            self.stats['exp'] += 1
            if i == 0:
                w(f'    }} else if (Pv <= {_lit(p1)}) {{')
                w(f'      kf = {ksum(k1)};')
            else:
                w(f'    }} else if (Pv == {_lit(p1)}) {{')
                w(f'      kf = {ksum(k1)};')
            if i == n - 2:
                w(f'    }} else if (Pv >= {_lit(p2)}) {{')
                w(f'      kf = {ksum(k2)};')
        w('    } else { kf = 0.0; }')
