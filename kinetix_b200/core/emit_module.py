"""Assemble the per-mechanism CUDA translation unit (`kx_mech.cu`) for sm_100a.

One generated file per (mechanism, option set) holds
  * the mechanism constants (`__constant__` pool + transport tables),
  * the BK1 kernel text from emit_bk1.BK1Emitter,
  * the BK2 / thermo kernels, which are hand-written templates (csrc/kx_bk2.cuh, csrc/kx_thermo.cuh)
    specialised through macros and tables emitted here,
  * `extern "C"` launchers + metadata getters (`kxm_*`) that the C-ABI host library
    (csrc/kx_host.cpp, libkinetix_b200.so) resolves with dlsym after dlopen()ing the compiled module.

It plays the role of the reference's generated `mech.h/rates.cpp/...` files *plus* the OKL wrappers that
textually include them (reference kinetix/core/generate.py:79-120, benchmark/src/kinetix.cpp:409-475).
"""
import math

from . import constants as const
from .emit_bk1 import BK1Emitter, ConstPool, _lit
from .emit_bk1_f32 import BK1EmitterF32, FloatPool

ABI_VERSION = 2      # 2: per-state pressure field argument in kxm_production_rates / kxm_thermo


def _table(name, rows, qualifier='__constant__', dims=None, ctype='real'):
    """`real` is typedef'd per module (double, or float in the --single-precision module); literals are
    written in full double precision and rounded by the compiler."""
    flat = [int(v) for v in rows] if ctype == 'int' else [float(v) for v in rows]
    body = ',\n  '.join(', '.join(f'(real){_lit(v)}' if ctype == 'real' else (str(v) if ctype == 'int' else _lit(v)) for v in flat[i:i + 4])
                        for i in range(0, len(flat), 4))
    dims = dims or f'[{len(flat)}]'
    return f'{qualifier} {ctype} {name}{dims} = {{\n  {body}\n}};\n'


def choose_tile(N, forced=None):
    """Tile edge for the BK2 pair loops: minimise padded pair count (ties: larger tiles first)."""
    best = None
    for tb in ([forced] if forced else (9, 8, 10, 7, 6)):
        NP = -(-N // tb) * tb
        cost = NP * NP / float(N * N)
        if best is None or cost < best[0] - 1e-9:
            best = (cost, tb, NP)
    return best[1], best[2]



def _emit_bk2_single(out, mech, fits, opt, options, sp, rsize, tq):
    """Tables + macros for csrc/kx_bk2.cuh (one thread per state).  Returns (dynamic smem bytes, states per CTA)."""
    N = mech.n_species
    M = mech.molar_masses
    tb, NP = choose_tile(N, opt.get('tile_bk2'))
    NB = NP // tb
    align = 16 // rsize                            # chunk sizes are multiples of 16 bytes (bulk copy)
    wchunk = -(-(N * tb) // align) * align         # reals per Wilke chunk
    dchunk = -(-(tb * tb * 6) // align) * align    # reals per diffusion tile
    limit = 227 * 1024 - 1024
    n_arrays = 1 if opt['bk2_scratch'] else 2      # [k][thread] arrays in shared memory: X (+ b_k / S_k)
    explicit = 'block_bk2' in (options or {}) or 'minb_bk2' in (options or {})
    max_block = opt['block_bk2'] * (2 if sp and 'block_bk2' not in (options or {}) else 1)

    def plan(split):
        """(resident threads, CTAs/SM, block, split, r0, j0, chunk_max) for a streaming granularity: whole
        table blocks (split 1) or half blocks (split 2: tile rows [0,r0)/[r0,tb), Wilke columns [0,j0)/[j0,N),
        all sub-chunks 16-byte multiples)."""
        r0 = j0 = 0
        if split == 2:
            r0 = next((r for r in range((tb + 1) // 2, tb) if (r * tb * 6) % align == 0), 0)
            j0 = next((j for j in range((N + 1) // 2, N) if (j * tb) % align == 0), 0)
            if not r0 or not j0:
                return None
            cmax = max(r0 * tb * 6, dchunk - r0 * tb * 6, j0 * tb, wchunk - j0 * tb)
        else:
            cmax = max(wchunk, dchunk)
        best = None
        for minb in ((opt['minb_bk2'],) if explicit else (2, 1)):
            top = max_block if explicit else (256 if minb == 2 else 512)
            for b in range(min(top, 1024), 31, -32):
                smem = 16 + (2 * cmax + n_arrays * NP * b) * rsize
                if smem * minb + 1024 * (minb - 1) <= limit and b * minb <= 640:
                    cand = (b * minb, minb, b, split, r0, j0, cmax)
                    if best is None or cand > best:
                        best = cand
                    break
        return best

    # occupancy: as many resident threads per SM as shared memory allows.  Half-block streaming halves the
    # staging buffers, which is what lets two CTAs share an SM (decoupled barriers / prologues: +45 % on the
    # 9-species mechanism); with a single CTA per SM it only adds barriers, so whole blocks are kept then.
    plans = [p for p in (plan(s_) for s_ in ((opt['bk2_split'],) if opt['bk2_split'] == 1 else (2, 1))) if p]
    two = [p for p in plans if p[1] >= 2 and p[3] == 2]
    one = [p for p in plans if p[3] == 1]
    chosen = max(two) if two and (not one or max(two)[0] >= max(one)[0]) else max(one or plans)
    _, opt['minb_bk2'], block2, split, r0, j0, chunk_max = chosen

    def smem_for(b):
        return 16 + (2 * chunk_max + n_arrays * NP * b) * rsize

    bk2_smem = smem_for(block2)
    opt['block_bk2'] = block2
    # Wilke mass factors: dense N x N, or the rank-r factorisation when it is cheaper (6 N r < 3 N^2)
    U, V, rank = wilke_low_rank(M)
    low_rank = opt.get('bk2_low_rank', True) and 2 * rank < N and rank % 2 == 0
    if low_rank:
        wrows = min(NP, chunk_max // rank)         # species rows per streamed chunk
        out.append(f'#define KX_WR {rank}')
        out.append(f'#define KX_WROWS {wrows}')
        out.append(f'#define KX_NWC {-(-NP // wrows)}')
    out.append(f'#define KX_TB {tb}')
    out.append(f'#define KX_NP {NP}')
    out.append(f'#define KX_BK2_BLOCK {block2}')
    out.append(f'#define KX_BK2_MINB {opt["minb_bk2"]}')
    out.append(f'#define KX_BK2_SCRATCH {1 if opt["bk2_scratch"] else 0}')
    out.append(f'#define KX_SPLIT {split}')
    out.append(f'#define KX_R0 {r0}')
    out.append(f'#define KX_J0 {j0}')
    out.append(f'#define KX_CHUNK_MAX {chunk_max}')
    out.append(f'#define KX_WCHUNK {wchunk}')
    out.append(f'#define KX_DCHUNK {dchunk}')
    out.append(_table('kx_m4', [m ** -0.25 for m in M]))
    out.append(_table('kx_cond', [c for k in range(N) for c in fits.conductivity[k]], qualifier=tq, dims=f'[{N}][5]'))
    out.append(_table('kx_visc', [c for k in range(N) for c in fits.viscosity[k]], qualifier=tq, dims=f'[{N}][5]'))
    if low_rank:
        for name, F in (('kx_wilke_v', V), ('kx_wilke_u', U)):
            rows = []
            for k in range(NP):
                rows += [float(x) for x in F[k]] if k < N else [0.0] * rank
            out.append(_table(name, rows, qualifier='__device__ const __align__(16)'))
    # Wilke mass factors c_kj = 1/sqrt(8 (1 + M_k/M_j)), one chunk per k-block: [kb][j][i], k = kb*tb + i
    wil = []
    for kb in range(NB if not low_rank else 0):
        chunk = []
        for j in range(N):
            for i in range(tb):
                k = kb * tb + i
                chunk.append(1.0 / math.sqrt(8.0 * (1.0 + M[k] / M[j])) if k < N else 0.0)
        chunk += [0.0] * (wchunk - len(chunk))
        wil += chunk
    if not low_rank:
        out.append(_table('kx_wilke', wil, qualifier='__device__ const __align__(16)'))
    # binary diffusion quartics, lower-triangular tiles; padded pairs evaluate to D = 1
    rcp = fits.reciprocal_diffusivity
    dif = []
    for kb in range(NB):
        for jb in range(kb + 1):
            for i in range(tb):
                for j in range(tb):
                    k, jj = kb * tb + i, jb * tb + j
                    if k < N and jj < N and k > jj:
                        dif += list(fits.diffusivity[k][jj]) + [0.0]
                    else:
                        dif += [1.0, 0.0, 0.0, 0.0, 0.0, 0.0]
            dif += [0.0] * (dchunk - tb * tb * 6)
    out.append(f'#define KX_RCP_DIFF {1 if rcp else 0}')
    out.append(_table('kx_diff', dif, qualifier='__device__ const __align__(16)'))
    out.append('#include "kx_bk2.cuh"')
    return bk2_smem, block2


def _emit_bk2_tmem(out, mech, fits, opt, tq):
    """Tables + macros for csrc/kx_bk2_tmem.cuh (persistent CTAs, KX_P states per thread, S_k in tensor memory,
    FP64).  Returns (smem bytes, states per CTA round, persistent=True) or None when the mechanism does not fit."""
    N = mech.n_species
    M = mech.molar_masses
    tb, NP = choose_tile(N, opt.get('tile_bk2'))
    NB = NP // tb
    U, V, rank = wilke_low_rank(M)
    wr = rank + (rank & 1)
    limit = 227 * 1024
    # doubles per state in tensor memory: the row blocks run in descending order and the top block's sums are
    # final (and consumed) before anything is stored, so only NB - 1 blocks ever live there
    ns = max(NP - tb, 1)
    # a diffusion tile is stored column by column, the rows of a column in PAIRS with their coefficients interleaved
    # (c0(i), c0(i+1), c1(i), c1(i+1), ...): the kernel walks a tile one column at a time and fetches the coefficients
    # of two pairs with five 16-byte loads (a 5-double record per pair would be 8-byte aligned every other pair)
    col_size = (tb // 2) * 10 + (tb % 2) * 6
    dchunk = tb * col_size
    cmax0 = max(dchunk, tb * (wr + 12))            # at least one species block per chunk of species rows

    def rows(width):
        return tb * min(NB, max(1, cmax0 // (tb * width)))
    vrows, urows = rows(wr + 12), rows(wr + 6)
    cmax = -(-max(dchunk, vrows * (wr + 12), urows * (wr + 6)) // 2) * 2
    # (threads, states per thread): as many resident states as shared memory (X_k: N doubles per state) and tensor
    # memory (2 ns columns per state, ceil(warps / 4) x spt states per lane) allow; two states per thread (half the
    # coefficient wavefronts per state) whenever that still leaves 8 warps.  Measured (M states/s): GRI-3.0 256 x 2:
    # 644, 256 x 1: 469; heptaneLu88 256 x 1: 236, 128 x 2: 197; EtOHKonnov (129 species) 128 x 1: 97-103, and with the
    # top block out of tensor memory 192 x 1 fits (two warps on lane quadrants 0 and 1).
    # CTAs of 256 or 128 threads (whole warps per scheduler: 192 one-state threads measured no faster than 128 on
    # EtOHKonnov, 98.4 vs 101.4 M states/s).  `lanes` = warps per state group: when only 128 one-state threads fit
    # (EtOHKonnov: X_k alone is 1 KB per state), 256 threads work as two halves that SHARE the 128 states -- warps w
    # and w + 4 own the same TMEM lanes and the same X_k -- and split every loop over species between them.
    nx = max(3 * wr + 2, tb)                        # doubles one half hands to the other at a time
    plans = []
    for spt in ((opt['bk2_spt'],) if opt.get('bk2_spt') else (2, 1)):
        for threads, lanes in ((256, 1), (256, 2), (128, 1)):
            if opt.get('bk2_threads') and threads != opt['bk2_threads']:
                continue
            if opt.get('bk2_lanes') and lanes != opt['bk2_lanes']:
                continue
            if lanes == 2 and (spt != 1 or tb % 2 or not opt.get('bk2_split_warps', True)):
                continue
            states = threads // lanes * spt
            slots = -(-(threads // lanes // 32) // 4)
            if slots * spt * 2 * (ns + 2 + (lanes * nx if lanes > 1 else 0)) > 512:
                continue                               # + 2 doubles per state: Mbar and sqrt(T) are parked there
            # two stages of the coefficient ring are enough (GRI-3.0: 632 vs 616 M states/s with four)
            for stages in ((opt['bk2_stages'],) if opt.get('bk2_stages') else (2,)):
                smem = 16 * stages + 16 + (stages * cmax + N * states) * 8
                if smem <= limit:
                    plans.append((threads, spt, stages, smem, lanes))
                    break
    plan = None
    if plans:
        # preference: the most resident states, then the most warps (256 x 2 > 256 x 1 > 256 as two halves > 128 x 1)
        plan = max(plans, key=lambda pl: (pl[0] // pl[4] * pl[1], pl[0]))
    # small mechanisms are latency / bandwidth leaning and run faster as two 128-thread CTAs per SM of the
    # one-state-per-thread kernel (LiDryer 10.4 vs 7.9, gri30-20 2.96 vs 2.71 G states/s)
    if plan is None or plan[0] < opt.get('bk2_tmem_min_threads', 128) or N < opt.get('bk2_tmem_min_species', 25):
        return None
    threads, spt, stages, smem, lanes = plan
    # every CTA allocates all 512 tensor-memory columns of its SM: never let two of them share an SM (the second
    # would wait in tcgen05.alloc until the first, persistent, CTA exits)
    smem = max(smem, 117 * 1024)

    # the coefficient stream of one batch, chunk by chunk (each chunk padded to a 16-byte multiple)
    stream, offs = [], [0]

    def add_chunk(vals):
        vals = list(vals)
        vals += [0.0] * (len(vals) & 1)
        assert len(vals) <= cmax
        stream.extend(vals)
        offs.append(len(stream))

    def species_row(k):
        if k < N:
            return list(fits.conductivity[k]) + list(fits.viscosity[k]) + [M[k] ** -0.25, 0.0]
        return [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0.0]
    for r0 in range(0, NP, vrows):
        add_chunk(v for k in range(r0, min(NP, r0 + vrows))
                  for v in (species_row(k) + ([float(x) for x in V[k]] if k < N else [0.0] * rank) + [0.0] * (wr - rank)))
    for r0 in range(0, NP, urows):
        add_chunk(v for k in range(r0, min(NP, r0 + urows))
                  for v in (([float(x) for x in U[k]] + [0.0] * (wr - rank) + list(fits.viscosity[k]) + [M[k] ** -0.25])
                            if k < N else [0.0] * wr + [1.0, 0, 0, 0, 0, 1.0]))
    nvc, nuc = -(-NP // vrows), -(-NP // urows)
    for kb in range(NB - 1, -1, -1):                   # row blocks in the order the kernel walks them: descending
        for jb in range(kb + 1):
            def quartic(i, j):
                k, jj = kb * tb + i, jb * tb + j
                return list(fits.diffusivity[k][jj]) if (k < N and jj < N and k > jj) else [1.0, 0.0, 0.0, 0.0, 0.0]
            tile = []
            for j in range(tb):
                for i in range(0, tb - 1, 2):
                    a, b2 = quartic(i, j), quartic(i + 1, j)
                    for c in range(5):
                        tile += [a[c], b2[c]]
                if tb % 2:
                    tile += quartic(tb - 1, j) + [0.0]
            assert len(tile) == dchunk
            add_chunk(tile)
    n_chunks = len(offs) - 1
    out.append(f'#define KX_TB {tb}')
    out.append(f'#define KX_NP {NP}')
    out.append(f'#define KX_P {spt}')
    out.append(f'#define KX_L {lanes}')
    out.append(f'#define KX_NS {ns}')
    out.append(f'#define KX_BK2_BLOCK {threads}')
    out.append(f'#define KX_STAGES {stages}')
    out.append(f'#define KX_CHUNK_MAX {cmax}')
    out.append(f'#define KX_COL {col_size}')
    # columns per iteration of the pair loop's body: one for two-state threads (3.5 KB of code, L0-resident; unrolling
    # further measured no faster: GRI-3.0 633 / 629 / 589 M states/s for 1 / 3 / 9), more for one-state threads whose
    # column is half the work (EtOHKonnov 5: 96.6 vs 91.5 M, heptaneLu88 8: 244 vs 232 M)
    cu = opt.get('bk2_col_unroll') or (1 if spt == 2 else max(d for d in range(1, 9) if tb % d == 0))
    if lanes > 1:
        cu = tb // lanes                               # each half takes its half of a tile in one go
    assert tb % cu == 0, 'bk2_col_unroll must divide the tile edge'
    out.append(f'#define KX_COL_UNROLL {cu}')
    out.append(f'#define KX_WR {wr}')
    out.append(f'#define KX_VROWS {vrows}')
    out.append(f'#define KX_UROWS {urows}')
    out.append(f'#define KX_NVC {nvc}')
    out.append(f'#define KX_NUC {nuc}')
    out.append(f'#define KX_N_CHUNKS {n_chunks}')
    out.append(f'#define KX_RCP_DIFF {1 if fits.reciprocal_diffusivity else 0}')
    out.append(_table('kx_chunk_off', offs, qualifier='__constant__', ctype='int'))
    out.append(_table('kx_bk2_stream', stream, qualifier='__device__ const __align__(16)'))
    out.append('#include "kx_bk2_tmem.cuh"')
    return smem, threads // lanes * spt, True


def wilke_low_rank(M, tol=1e-13):
    """Factor the Wilke mass-factor matrix c_kj = (8 (1 + M_k/M_j))^-1/2 as U V^T (truncated SVD, singular values
    folded into U).  c is a smooth kernel in ln M_k - ln M_j, so its numerical rank is ~12-14 whatever the number
    of species; the rank is raised until every entry is reproduced to `tol` (relative; the SVD itself bottoms out at ~1e-14)."""
    import numpy as np
    M = np.asarray(M, dtype=np.float64)
    C = 1.0 / np.sqrt(8.0 * (1.0 + M[:, None] / M[None, :]))
    u, s, vt = np.linalg.svd(C)
    for rank in range(1, len(M) + 1):
        U = u[:, :rank] * s[:rank]
        V = vt[:rank].T
        if np.max(np.abs(U @ V.T - C) / C) <= tol:
            break
    if rank & 1 and rank < len(M):                 # the kernels read the factors two columns at a time
        rank += 1
        U, V = u[:, :rank] * s[:rank], vt[:rank].T
    return U, V, rank


def emit_module(mech, fits, options=None, single_precision=False):
    """Return the CUDA source text.  `fits` is a TransportFits or None (BK2 omitted).
    single_precision: FP32 arithmetic; the module then serves FP64 buffers ("fpmix") and FP32 buffers."""
    sp = bool(single_precision)
    rsize = 4 if sp else 8
    opt = dict(block_bk1=128, minb_bk1=3, sync_every=16, gibbs_in_smem=True, reorder=True, prefetch=4, ring=0, pin_loads=False,
               inline_constants=False, param_constants=True, l1_keep=False, keep_until=0, live_cap=0, eff_in_smem=True, nasa_indexed='ldg',
               block_bk2=128, minb_bk2=2, bk2_scratch=False, bk2_split=2)
    opt.update(options or {})
    N = mech.n_species
    K = FloatPool('kcf', inline=opt.get('inline_constants_f32', True)) if sp else ConstPool('kc', inline=opt['inline_constants'], as_param=opt['param_constants'])
    # FP32 BK1 (GRI-3.0 fpmix, M states/s): constants as 32-bit immediates instead of pool loads 1800 -> 2640 (40 % of the
    # stall samples sat on LDC: the second constant of an FFMA is an explicit load through the MIO queue), then
    # 3 CTAs x 168 registers instead of 4 x 128: 2800; NASA table through L1 2580, barriers 2760-2810, 2 CTAs 1930

    def emit_bk1(kernel_name='kx_bk1_f64'):
        if sp:
            e = BK1EmitterF32(mech, K)
            return e, e.emit('kx_bk1_f32', opt['block_bk1'], opt['minb_bk1'], opt.get('sync_every_f32', 0), opt['reorder'],
                             nasa_indexed=opt.get('nasa_indexed_f32', False))
        e = BK1Emitter(mech, K)
        src = e.emit(kernel_name, opt['block_bk1'], opt['minb_bk1'], opt['sync_every'], opt['gibbs_in_smem'],
                     opt['reorder'], opt['prefetch'], opt['ring'], opt['pin_loads'], opt['l1_keep'], opt['keep_until'], opt['live_cap'], opt['eff_in_smem'], opt['nasa_indexed'],
                     tmem_slots=bk1_tm['slots'], smem_cap=bk1_tm['smem_cap'], tmem_cols=bk1_tm.get('cols', 512),
                     cold_uses=opt.get('cold_uses', 0), cold_slot_cap=opt.get('cold_slot_cap', 0),
                     cold_conc_only=opt.get('cold_conc_only', False), gibbs_prefer_tm=opt.get('gibbs_prefer_tm', False),
                     kbase_ahead=opt.get('kbase_ahead', 0))
        return e, src

    bk1_tm = dict(slots=0, smem_cap=0)
    bk1, bk1_src = emit_bk1()
    # FOUR warps per scheduler for mid-size mechanisms (round 2): two 256-thread CTAs per SM at 128 registers.  What makes
    # 128 registers enough: the (read-only) concentrations C_k of the live species sit in shared-memory slots, the
    # exp(+-g_k) and third-body sums go to TENSOR memory first (two CTAs x 256 columns = 64 doubles per thread) and
    # overflow to shared memory (53 slots per thread), so the registers hold the accumulators wdot_k and temporaries
    # only (GRI-3.0: 0.6 KB of spill loads, as at 168 registers before).  Two instruction streams per SM instead of
    # three, four warps per scheduler instead of three to cover dependent-issue and instruction-fetch stalls.
    # M states/s, classic layout (3-4 CTAs x 128 threads, 168 / 128 registers) -> this one:
    #   gri30 956 -> 1020, gri30-35 1243 -> 1444, NH3Konnov_edit 880 -> 1289 (its 81 slots had allowed two 128-thread CTAs),
    #   chempolimi_edit 1347 -> 1416, gri30-27 2538 -> 2656, gri30-20 3577 -> 3807, H2_new_mech 5324 -> 5826, H2_Konnov 5968 -> 6054;
    #   not applied: LiDryer 13661 -> 12120 (8 live species), heptaneLu88 718 -> 709 (41 live species: 2.9 KB of spills).
    # Around it on GRI-3.0: 4 x 128 threads 892 (four streams: no_instruction 1.86), 1 x 512 983, a CTA barrier every
    # 12 / 16 / 24 / 32 reactions 994 / 1017 / 1007 / 993 and none 1022-1024, exp(+-g) of often-used species in shared
    # memory instead 859-970 (more spills), wdot_k of rarely used species in slots too 900-995, C_k of rarely used species
    # only 906, TMEM base address in a uniform register (563 R2UR fewer, 1.1 KB of spills) 1001.
    layout_keys = ('block_bk1', 'minb_bk1', 'bk1_tmem', 'bk1_tmem_block', 'bk1_tmem_ctas', 'bk1_smem_cap', 'cold_uses',
                   'sync_every', 'live_cap', 'gibbs_in_smem', 'ring', 'prefetch', 'eff_in_smem')
    wide = (not sp and opt['gibbs_in_smem'] and not opt['ring'] and opt.get('bk1_layout', 'auto') != 'classic'
            and not any(k in (options or {}) for k in layout_keys)
            and 10 <= bk1.schedule_stats.get('peak_live', 0) <= 36 and bk1.smem_doubles_per_thread <= 53 + 64 - 2)
    if wide:
        opt.update(bk1_tmem=True, bk1_tmem_block=256, bk1_tmem_ctas=2, cold_uses=1 << 20, cold_conc_only=True,
                   gibbs_prefer_tm=True, sync_every=0)
    # small mechanisms: few live species need few registers, more resident CTAs pay (M states/s, 8 Mi states;
    # FP64 3 -> 4 CTAs: LiDryer 13580 -> 13750, H2_Konnov 5470 -> 5980, gri30-20 3090 -> 3580; FP32 math 3 -> 6 / 8:
    # LiDryer 29100 -> 34100 / 35600 = 5.7 TB/s of HBM traffic, H2_Konnov 14660 -> 16880 / 16160, gri30-20 9200 -> 10170 / 9140)
    if 'minb_bk1' not in (options or {}) and 'block_bk1' not in (options or {}) and not wide:
        peak = bk1.schedule_stats.get('peak_live', N)
        small = 4 if not sp else (8 if peak <= 8 else 6)
        # FP64, 17-25 live species: chempolimi_edit (21) 1180 -> 1340, gri30-27 (19) and gri30-35 (25) unchanged
        if peak <= (21 if not sp else 16) and small != opt['minb_bk1']:
            opt['minb_bk1'] = small
            bk1, bk1_src = emit_bk1()
    budget = 220 * 1024
    # mid-size mechanisms whose live set does not fit the registers but whose scratch slots leave most of the shared
    # memory idle (heptaneLu88: 41 live species, 29 slots): the live species keep their (read-only) concentration C_k in
    # a shared-memory slot, only the accumulators wdot_k stay in registers -- explicit placement of what ptxas would
    # spill (spill loads 2.7 KB -> 0.66 KB per state).  M states/s: no placement 652, C_k and wdot_k of species with
    # <= 40 uses in slots 676-682, C_k alone of those species 695, C_k of every species 712.  (GRI-3.0, 3 CTAs per SM
    # kept: 942-947 vs 952; EtOHKonnov in the tensor-memory layout: 129-165 vs 162 -- not applied to either.)
    if (not sp and opt['gibbs_in_smem'] and 'cold_uses' not in (options or {}) and not wide
            and bk1.schedule_stats.get('peak_live', 0) > 30 and bk1.smem_doubles_per_thread <= 40
            and bk1.smem_doubles_per_thread * 8 * 128 * 2 <= budget):
        opt['cold_uses'], opt['cold_slot_cap'], opt['cold_conc_only'] = 1 << 20, 70, True
        bk1, bk1_src = emit_bk1()
    # large mechanisms: when the scratch slots (exp(+-g_k) of live species, third-body sums) allow at most one
    # 128-thread CTA per SM in shared memory (EtOHKonnov: 210 slots), spread them over shared AND tensor memory
    # and run ONE 256-thread CTA per SM: twice the warps to hide latencies, and the (MB-sized) straight-line
    # instruction stream is fetched once for 8 warps instead of 4.  'auto' | True | False
    want_tm = opt.get('bk1_tmem', 'auto')
    if not sp and opt['gibbs_in_smem'] and want_tm and not opt['ring']:
        tm_block = opt.get('bk1_tmem_block', 256)
        tm_ctas = opt.get('bk1_tmem_ctas', 1)            # CTAs per SM sharing the 512 columns
        tm_cols = 512 // tm_ctas
        tm_slots = tm_cols // 2 // -(-tm_block // 128)    # doubles per thread: the columns are shared by block/128 warp groups
        cap = min(budget // (tm_block * tm_ctas * 8), opt.get('bk1_smem_cap', 1 << 30))
        need = bk1.smem_doubles_per_thread
        if (want_tm is True or need * 8 * 128 * 2 > budget) and need <= tm_slots + cap:
            bk1_tm = dict(slots=tm_slots, smem_cap=cap, cols=tm_cols)
            opt['block_bk1'], opt['minb_bk1'] = tm_block, tm_ctas
            # the 8 warps share one pass over ~0.75 MB of straight-line code: a barrier every 4 reactions keeps
            # them inside the same few KB of it (EtOHKonnov, M states/s: every 16: 118, 8: 135, 4: 157, 1: 148)
            if 'sync_every' not in (options or {}) and not wide:
                opt['sync_every'] = 4
            # cap the live set (live-range splitting through the output rows, emit_bk1.py) so that the register
            # spills shrink (EtOHKonnov: 87 -> 60 live species, 10-13 KB -> 6 KB of spill loads per state for 57
            # more exp): without it the throughput falls with the batch size (158 M st/s at 1 Mi states, 127 M at
            # 4 Mi: spill lines compete with the streamed rows for L2); with it 163 M at both sizes
            if 'live_cap' not in (options or {}) and bk1.schedule_stats.get('peak_live', 0) > 60:
                opt['live_cap'] = 60
            bk1, bk1_src = emit_bk1()
    # occupancy: the shared-memory slots per thread are a property of the schedule; if the requested CTAs per SM
    # do not fit, lower the CTA count (then the CTA size) and emit again
    per_thread = bk1.smem_doubles_per_thread * 8
    changed = False
    while per_thread * opt['block_bk1'] * opt['minb_bk1'] > budget and (opt['minb_bk1'] > 1 or opt['block_bk1'] > 32):
        if opt['minb_bk1'] > 1:
            opt['minb_bk1'] -= 1
        else:
            opt['block_bk1'] -= 32
        changed = True
    if changed:
        bk1, bk1_src = emit_bk1()
    bk1_smem = bk1.smem_doubles_per_thread * 8 * opt['block_bk1']
    # small launches of the four-warp layout: below one wave the latency of ONE pass is all there is, and the classic
    # layout's pass is shorter (GRI-3.0, 1 Ki - 16 Ki states: 51 vs 70 us; from 64 Ki states on the wide layout wins,
    # 836 vs 579 M states/s).  The module carries the classic kernel as `kx_bk1_f64s` for launches of at most one wave of
    # it.  Same schedule, hence the same NASA table (checked), one constant pool.
    bk1_small = None
    if wide and opt.get('bk1_small', True):
        saved_opt, saved_tm = dict(opt), bk1_tm
        opt.update(block_bk1=128, minb_bk1=4 if bk1.schedule_stats.get('peak_live', N) <= 21 else 3, sync_every=16,
                   bk1_tmem=False, cold_uses=0, cold_conc_only=False, gibbs_prefer_tm=False)
        bk1_tm = dict(slots=0, smem_cap=0)
        e_s, src_s = emit_bk1('kx_bk1_f64s')
        while e_s.smem_doubles_per_thread * 8 * 128 * opt['minb_bk1'] > budget and opt['minb_bk1'] > 1:
            opt['minb_bk1'] -= 1
            e_s, src_s = emit_bk1('kx_bk1_f64s')
        if (e_s.nasa_lo_tab, e_s.nasa_hi_tab) == (bk1.nasa_lo_tab, bk1.nasa_hi_tab):
            bk1_small = dict(src=src_s, block=128, minb=opt['minb_bk1'], smem=e_s.smem_doubles_per_thread * 8 * 128)
        opt.clear()
        opt.update(saved_opt)
        bk1_tm = saved_tm
    out = []
    out.append('// GENERATED by kinetix_b200 -- mechanism-specialised sm_100a kernels. Do not edit.')
    out.append(f'// mechanism: {mech.name}  species: {N} (active {mech.n_active})  reactions: {mech.n_reactions}')
    out.append('#include <cuda_runtime.h>')
    out.append('#include <math_constants.h>')
    # 32-entry table exp in BK1 (csrc/kx_math.cuh): 17 % fewer FP64 instructions but measured SLOWER on GRI-3.0
    # (879 vs 909 M states/s: more spills at 168 registers, and the kernel is issue/latency bound, not FP64 bound)
    if opt.get('exp_table', False) and not sp:
        out.append('#define KX_EXP_TABLE 1')
    for d in opt.get('defines', ()):               # development switches (tools/build_variants.py)
        out.append(f'#define {d}')
    out.append('#include "kx_math.cuh"')
    if bk1_tm['slots']:
        out.append('#include "kx_tm.cuh"')
    out.append(f'#define KX_N {N}')
    out.append(f'#define KX_SINGLE_PRECISION {1 if sp else 0}')
    out.append('typedef float real;\ntypedef float2 real2;' if sp else 'typedef double real;\ntypedef double2 real2;')
    out.append('')
    out.append(K.definition())
    if hasattr(bk1, 'nasa_table_definition'):
        out.append(bk1.nasa_table_definition())
    out.append(bk1_src)
    if bk1_small:
        out.append(bk1_small['src'])

    M = mech.molar_masses
    # constant bank budget (64 KB): pool + per-species tables; large mechanisms keep the tables that are
    # indexed at run time (thermo, transport pure-species fits) in global memory instead (L1/L2 cached,
    # warp-uniform addresses)
    const_bytes = rsize * (len(K.values) + N * (3 + 14 + 1 + 10))
    tq = '__constant__' if const_bytes < 56 * 1024 else '__device__ const'
    out.append(_table('kx_rcpM', [1. / m for m in M]))
    out.append(_table('kx_M', M))
    # ---- thermo tables ----
    out.append(_table('kx_Tmid', [s.T_mid for s in mech.species]))
    nasa = []
    for s in mech.species:
        nasa += list(s.nasa_lo) + list(s.nasa_hi)
    out.append(_table('kx_nasa', nasa, qualifier=tq, dims=f'[{N}][2][7]'))
    out.append('#include "kx_thermo.cuh"')

    has_bk2 = fits is not None
    bk2_smem = 0
    bk2_states_per_cta = 0
    if has_bk2:
        limit = 227 * 1024 - 1024
        align = 16 // rsize                            # chunk sizes are multiples of 16 bytes (bulk copy)
        planned = None
        if opt.get('bk2_tmem', True) and not sp:
            planned = _emit_bk2_tmem(out, mech, fits, opt, tq)
        bk2_persistent = False
        if planned:
            bk2_smem, bk2_states_per_cta, bk2_persistent = planned
        else:
            bk2_smem, bk2_states_per_cta = _emit_bk2_single(out, mech, fits, opt, options, sp, rsize, tq)

    names = ' '.join(mech.species_names)
    out.append(f'''
// ---- host side: metadata + launchers resolved by libkinetix_b200.so (csrc/kx_host.cpp) ----
// dtype = storage type of the buffers: 0 = FP64, 1 = FP32.  This module computes in {"FP32" if sp else "FP64"}:
// {"dtype 0 is the reference's fpmix flavour, dtype 1 its pure FP32 flavour (kinetix.cpp:254-281)" if sp else "only dtype 0 is served; other precisions live in the --single-precision module"}.
extern "C" {{
int kxm_abi_version() {{ return {ABI_VERSION}; }}
int kxm_single_precision() {{ return {1 if sp else 0}; }}
int kxm_n_species() {{ return {N}; }}
int kxm_n_active_species() {{ return {mech.n_active}; }}
int kxm_n_reactions() {{ return {mech.n_reactions}; }}
int kxm_has_transport() {{ return {1 if has_bk2 else 0}; }}
const char* kxm_species_names() {{ return "{names}"; }}
const char* kxm_mechanism_name() {{ return "{mech.name}"; }}
void kxm_molar_masses(double* out) {{
  static const double M[{N}] = {{{', '.join(_lit(float(m)) for m in M)}}};
  for (int k = 0; k < {N}; k++) out[k] = M[k];
}}
}}  // extern "C"

template <typename K>
static int kxm_set_smem(K kernel, size_t smem) {{
  if (smem <= 48 * 1024) return 0;
  return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}}
// launch-time state is kept PER DEVICE (function attributes belong to a device's context, SM counts differ): one
// process may drive several GPUs through kx_select_device (csrc/kx_host.cpp)
#define KXM_MAX_DEVICES 64
static int kxm_device() {{
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < KXM_MAX_DEVICES ? dev : 0;
}}
''')
    small_cfg = small_launch = ''
    if bk1_small:
        pool_arg = ', kx_param_pool' if opt['param_constants'] else ''
        small_cfg = (f"    if (int e = kxm_set_smem(kx_bk1_f64s<false>, {bk1_small['smem']})) return e;\n"
                     f"    if (int e = kxm_set_smem(kx_bk1_f64s<true>, {bk1_small['smem']})) return e;\n")
        small_launch = f'''  if (n <= (long long)n_sm[dev] * {bk1_small['block'] * bk1_small['minb']}) {{   // at most one wave of the classic layout
    const unsigned g = (unsigned)((n + {bk1_small['block']} - 1) / {bk1_small['block']});
    if (pfield)
      kx_bk1_f64s<true><<<g, {bk1_small['block']}, {bk1_small['smem']}, stream>>>(n, offsetT, offset, pressure_R, pressure, log(pressure),
                                                     (const double*)state, (double*)rates, Tref, (const double*)pfield{pool_arg});
    else
      kx_bk1_f64s<false><<<g, {bk1_small['block']}, {bk1_small['smem']}, stream>>>(n, offsetT, offset, pressure_R, pressure, log(pressure),
                                                      (const double*)state, (double*)rates, Tref, nullptr{pool_arg});
    return (int)cudaGetLastError();
  }}
'''
    if sp:
        out.append(f'''
template <typename S>
static int launch_bk1(long long n, long long offsetT, long long offset, double pressure_R, double pressure,
                      const void* state, void* rates, double Tref, const void* pfield, cudaStream_t stream) {{
  const int block = {opt['block_bk1']};
  const unsigned grid = (unsigned)((n + block - 1) / block);
  if (pfield)
    kx_bk1_f32<S, true><<<grid, block, 0, stream>>>(n, offsetT, offset, (float)pressure_R, (float)pressure,
                                                    (float)log(pressure), (const S*)state, (S*)rates, Tref, (const S*)pfield);
  else
    kx_bk1_f32<S, false><<<grid, block, 0, stream>>>(n, offsetT, offset, (float)pressure_R, (float)pressure,
                                                     (float)log(pressure), (const S*)state, (S*)rates, Tref, nullptr);
  return (int)cudaGetLastError();
}}
''')
    else:
        out.append(f'''
template <typename S>
static int launch_bk1(long long n, long long offsetT, long long offset, double pressure_R, double pressure,
                      const void* state, void* rates, double Tref, const void* pfield, cudaStream_t stream) {{
  const int block = {opt['block_bk1']};
  const size_t smem = {bk1_smem};
  static bool configured[KXM_MAX_DEVICES] = {{}};
  static int n_sm[KXM_MAX_DEVICES] = {{}};
  const int dev = kxm_device();
  if (!configured[dev]) {{
    if (int e = kxm_set_smem(kx_bk1_f64<false>, smem)) return e;
    if (int e = kxm_set_smem(kx_bk1_f64<true>, smem)) return e;
    if (cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1002;
{small_cfg}    configured[dev] = true;
  }}
{small_launch}  const unsigned grid = (unsigned)((n + block - 1) / block);
  if (pfield)
    kx_bk1_f64<true><<<grid, block, smem, stream>>>(n, offsetT, offset, pressure_R, pressure, log(pressure),
                                                    (const double*)state, (double*)rates, Tref, (const double*)pfield{', kx_param_pool' if opt['param_constants'] else ''});
  else
    kx_bk1_f64<false><<<grid, block, smem, stream>>>(n, offsetT, offset, pressure_R, pressure, log(pressure),
                                                     (const double*)state, (double*)rates, Tref, nullptr{', kx_param_pool' if opt['param_constants'] else ''});
  return (int)cudaGetLastError();
}}
''')
    out.append(f'''
template <typename S>
static int launch_thermo(long long n, long long offsetT, long long offset, double pressure_R, const void* state,
                         void* rho, void* cp, void* rhoCp, double Tref, const void* pfield, cudaStream_t stream) {{
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (pfield)
    kx_thermo<S, true><<<grid, 256, 0, stream>>>(n, offsetT, offset, (real)pressure_R, (const S*)state, (S*)rho,
                                                 (S*)cp, (S*)rhoCp, Tref, (const S*)pfield);
  else
    kx_thermo<S, false><<<grid, 256, 0, stream>>>(n, offsetT, offset, (real)pressure_R, (const S*)state, (S*)rho,
                                                  (S*)cp, (S*)rhoCp, Tref, nullptr);
  return (int)cudaGetLastError();
}}
''')
    if has_bk2:
        if bk2_persistent:
            out.append(f'''
template <typename S>
static int launch_bk2(long long n, long long offsetT, long long offset, double pressure, const void* state,
                      void* conductivity, void* viscosity, void* rhoD, double Tref, cudaStream_t stream) {{
  // persistent CTAs, one per SM, each loops over batches of KX_BK2_BLOCK x P states.  Launches with fewer than one
  // full batch per SM use one state per thread: half-size batches, twice as many SMs at work.
  const int block = KX_BK2_BLOCK;
  const size_t smem = {bk2_smem};
  static bool configured[KXM_MAX_DEVICES] = {{}};
  static int n_sm[KXM_MAX_DEVICES] = {{}};
  const int dev = kxm_device();
  if (!configured[dev]) {{
    if (int e = kxm_set_smem(kx_bk2<S, KX_P, KX_L>, smem)) return e;
    if (KX_P > 1) if (int e = kxm_set_smem(kx_bk2<S, 1, KX_L>, smem)) return e;
    if (cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1002;
    configured[dev] = true;
  }}
  const bool small = KX_P > 1 && n <= (long long)n_sm[dev] * (block / KX_L) * (KX_P - 1);
  const int per_cta = (block / KX_L) * (small ? 1 : KX_P);
  unsigned grid = (unsigned)((n + per_cta - 1) / per_cta);
  if (grid > (unsigned)n_sm[dev]) grid = (unsigned)n_sm[dev];
  if (small) {{
    kx_bk2<S, 1, KX_L><<<grid, block, smem, stream>>>(n, offsetT, offset, (real)pressure, (const S*)state, (S*)conductivity,
                                               (S*)viscosity, (S*)rhoD, Tref);
    return (int)cudaGetLastError();
  }}
  // The batches of a launch are dealt round-robin to the persistent CTAs: when their number is not a multiple of the
  // grid, a last round keeps a few SMs busy for a whole batch time.  If that remainder is at most half the grid, the
  // states of the last round go to a second launch of the one-state-per-thread instantiation instead: twice as many
  // CTAs with half-size batches (about 0.6 of a full batch time).  1 Mi GRI-3.0 states = 13 rounds + 29 batches: 14
  // batch times become 13.6.
  const long long nb = (n + per_cta - 1) / per_cta;
  const long long rem = nb % grid;
  long long n_main = n;
  if (KX_P > 1 && nb > (long long)grid && rem > 0 && 2 * rem <= (long long)grid) n_main = (nb - rem) * per_cta;
  kx_bk2<S, KX_P, KX_L><<<grid, block, smem, stream>>>(n_main, offsetT, offset, (real)pressure, (const S*)state,
                                                (S*)conductivity, (S*)viscosity, (S*)rhoD, Tref);
  if (n_main < n) {{
    const long long n_tail = n - n_main;
    const int per_tail = block / KX_L;
    unsigned g2 = (unsigned)((n_tail + per_tail - 1) / per_tail);
    if (g2 > (unsigned)n_sm[dev]) g2 = (unsigned)n_sm[dev];
    kx_bk2<S, 1, KX_L><<<g2, block, smem, stream>>>(n_tail, offsetT, offset, (real)pressure, (const S*)state + n_main,
                                             (S*)conductivity + n_main, (S*)viscosity + n_main, (S*)rhoD + n_main, Tref);
  }}
  return (int)cudaGetLastError();
}}
''')
        else:
            out.append(f'''
template <typename S>
static int launch_bk2(long long n, long long offsetT, long long offset, double pressure, const void* state,
                      void* conductivity, void* viscosity, void* rhoD, double Tref, cudaStream_t stream) {{
  const int block = KX_BK2_BLOCK, per_cta = {bk2_states_per_cta};   // states per CTA
  const size_t smem = {bk2_smem};
  static bool configured[KXM_MAX_DEVICES] = {{}};
  const int dev = kxm_device();
  if (!configured[dev]) {{
    if (int e = kxm_set_smem(kx_bk2<S>, smem)) return e;
    configured[dev] = true;
  }}
  const unsigned grid = (unsigned)((n + per_cta - 1) / per_cta);
  kx_bk2<S><<<grid, block, smem, stream>>>(n, offsetT, offset, (real)pressure, (const S*)state, (S*)conductivity,
                                           (S*)viscosity, (S*)rhoD, Tref);
  return (int)cudaGetLastError();
}}
''')
    f32_case = 'if (dtype == 1) return {fn}<float>({args});' if sp else 'if (dtype == 1) return 1000;'
    a1 = 'n, offsetT, offset, pressure_R, pressure, state, rates, Tref, pfield, stream'
    a2 = 'n, offsetT, offset, pressure, state, conductivity, viscosity, rhoD, Tref, stream'
    a3 = 'n, offsetT, offset, pressure_R, state, rho, cp, rhoCp, Tref, pfield, stream'
    out.append(f'''
extern "C" {{
// pfield: NULL, or n per-state pressures p / p_ref (storage type of the state); the scalars then carry p_ref
int kxm_production_rates(long long n, long long offsetT, long long offset, double pressure_R, double pressure,
                         const void* state, void* rates, double Tref, const void* pfield, int dtype,
                         cudaStream_t stream) {{
  if (n <= 0) return 0;
  {f32_case.format(fn='launch_bk1', args=a1)}
  if (dtype != 0) return 1000;
  return launch_bk1<double>({a1});
}}

int kxm_thermo(long long n, long long offsetT, long long offset, double pressure_R, const void* state,
               void* rho, void* cp, void* rhoCp, double Tref, const void* pfield, int dtype, cudaStream_t stream) {{
  if (n <= 0) return 0;
  {f32_case.format(fn='launch_thermo', args=a3)}
  if (dtype != 0) return 1000;
  return launch_thermo<double>({a3});
}}

int kxm_transport(long long n, long long offsetT, long long offset, double pressure, const void* state,
                  void* conductivity, void* viscosity, void* rhoD, double Tref, int dtype, cudaStream_t stream) {{
''' + (f'''  if (n <= 0) return 0;
  {f32_case.format(fn='launch_bk2', args=a2)}
  if (dtype != 0) return 1000;
  return launch_bk2<double>({a2});
''' if has_bk2 else '  return 1001;\n') + '}\n')
    out.append('}  // extern "C"\n')
    return '\n'.join(out), dict(bk1=bk1.stats, bk1_schedule=bk1.schedule_stats, n_const=len(K.values))
