#!/usr/bin/env python3
"""bench.py -- BK1/BK2 throughput of the B200-native KinetiX hot path (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]                 our CUDA path (one process per GPU)
  python bench.py --impl reference [--gpus N] --steps K --warmup W    the reference's own CPU implementation
                                                                       (oracle/_ref) on the host's cores

A *step* is one pass of the hot path over one batch of synthetic states resident in HBM: one BK1 launch
(kinetix::productionRates) and one BK2 launch (kinetix::mixtureAvgTransportProps) over
`--n-states` states per GPU (default 16 Mi, BASELINE.json's target size), GRI-Mech 3.0, FP64.
value = (states per GPU x N GPUs) / (device time of one step, max over ranks)  [states/s through BK1+BK2];
the per-kernel rates are reported beside it.  States are independent: the batch is sharded over the
GPUs, there is no collective on the data path ("scaling": "weak").

Beside the headline the same line carries `configs`: BASELINE.json's other configurations measured in the same
process (LiDryer BK1 8 Mi states/GPU in FP64 and with FP32 math -- the bandwidth-leaning regime --, EtOHKonnov
BK1+BK2 16 Mi states/GPU, GRI-3.0 BK1 with FP32 math, and a STRONG-scaling point: 16 Mi GRI states in total cut
into per-rank shards by kinetix_b200/sharding.py), each with the roofline fraction of the pipe that bounds it.

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'BK1+BK2 states/sec (GRI-3.0, FP64)'
UNIT = 'states/s'
P_ATM = 101325.0

# FP64-pipe lane instructions per state (DFMA/DMUL/DADD/DSETP, one per lane) the REFERENCE's arithmetic costs as its
# generator emits it, with libdevice-class transcendental costs (probed SASS of the reference-style
# kernel).  Kept for continuity with round 1 (`frac_reference_W`); the roofline's `frac` uses the count of the
# MINIMAL form known today = what our kernels execute (shared exps, low-rank Wilke, 2-instruction pair reciprocal),
# so that redundant arithmetic can never inflate the fraction ("use the minimal-form W").
W_REFERENCE = {'gri30': {'bk1': 1.89e4, 'bk2': 2.47e4}}


def read_json(path):
    try:
        with open(path) as fh:
            return json.load(fh)
    except Exception:
        return None


def fp64_peak():
    """Measured DFMA issue peak of this pool's B200 [lane-instr/s] (tools/peaks.cu -> profiles/peaks_r01.json);
    nominal 148 SM x 64 lanes x 1.965 GHz if the measurement is absent.  MEASURED_PEAKS.json has no FP64 entry."""
    pk = read_json(os.path.join(ROOT, 'profiles', 'peaks_r01.json'))
    if pk and 'dfma_lane_instr_per_s_sustained_3s' in pk:
        return float(pk['dfma_lane_instr_per_s_sustained_3s']), 'measured DFMA issue rate (tools/peaks.cu -> profiles/peaks_r01.json)'
    return 148 * 64 * 1.965e9, 'nominal 148 SM x 64 lanes x 1.965 GHz'


def hbm_peak():
    pk = read_json(os.path.join(ROOT, 'MEASURED_PEAKS.json'))
    if pk and 'hbm_gbs' in pk:
        return float(pk['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def mufu_peak():
    pk = read_json(os.path.join(ROOT, 'profiles', 'peaks_r01.json'))
    if pk and 'mufu_ex2_lane_instr_per_s' in pk:
        return float(pk['mufu_ex2_lane_instr_per_s'])
    return 148 * 16 * 1.965e9


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {'sm_mhz': float(np.median(busy)), 'sm_max_mhz': max(mx), 'power_w_max': max(pw),
                'samples': len(sm), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's generated SERIAL code (oracle/_ref), all host cores
# ---------------------------------------------------------------------------------------------------
class CpuReference:
    """W independent workers, each owning its own slab of states -- the semantics of
    `mpirun -np W kinetix_bk --backend SERIAL` (bk.cpp:443,705-718); MPI is absent, so threads + join."""

    PER_WORKER = 16384     # states per worker per pass

    def __init__(self, mech):
        from oracle import build_ref
        from oracle.port import Port, synthetic_states
        self.kind = 'reference'
        lib = build_ref.lib_path(mech, 'serial')
        if os.path.isdir(build_ref.gen_dir(mech, 'serial')):
            lib = build_ref.ensure_serial_native(mech)       # the reference's -march=native, on THIS host
        if not os.path.exists(lib):
            lib = None
        self.port = None
        if lib is None:
            self.kind = 'port'
            self.port = Port(mech)
            self.N = self.port.N
        else:
            self.lib = ctypes.CDLL(lib)
            self.N = self.lib.ref_n_species()
        self.libpath = lib
        try:
            self.cores = len(os.sched_getaffinity(0))
        except AttributeError:
            self.cores = os.cpu_count() or 1
        if self.port is not None:
            self.cores = 1
        S = self.PER_WORKER
        self.slabs = []
        for w in range(self.cores):
            st = np.ascontiguousarray(synthetic_states(self.N, S, seed=1234 + w))
            self.slabs.append((st, np.empty_like(st), np.empty(S), np.empty(S), np.empty((self.N, S))))

    def _worker(self, w, passes):
        from oracle.port import R
        st, rates, cond, visc, rhoD = self.slabs[w]
        S = st.shape[1]
        if self.port is not None:
            for _ in range(passes):
                self.port.production_rates(st, P_ATM / R, P_ATM)
                self.port.transport(st, 1.0)
            return
        dp = ctypes.POINTER(ctypes.c_double)
        p = lambda a: a.ctypes.data_as(dp)
        L = ctypes.c_long
        for _ in range(passes):
            self.lib.ref_production_rates(L(S), L(S), L(S), ctypes.c_double(P_ATM / R), ctypes.c_double(P_ATM),
                                          p(st), p(rates), ctypes.c_double(1.0))
            self.lib.ref_transport(L(S), L(S), L(S), ctypes.c_double(1.0), p(st), p(cond), p(visc), p(rhoD),
                                   ctypes.c_double(1.0))

    def run(self, passes=1):
        """one timed pass-set over all workers; returns (seconds, states processed)"""
        threads = [threading.Thread(target=self._worker, args=(w, passes)) for w in range(self.cores)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
        return dt, self.cores * self.PER_WORKER * passes

    def describe(self, passes):
        src = 'rolled SERIAL code from the unmodified reference generator, g++ -O3 -march=native -ffast-math' \
            if self.kind == 'reference' else 'numpy port (oracle/port.py)'
        return (f'{self.cores} workers x {self.PER_WORKER} states x {passes} pass(es) of BK1+BK2, GRI-3.0 FP64; {src}')


def cpu_baseline(mech, budget_s=12.0):
    ref = CpuReference(mech)
    dt, n = ref.run(1)                          # warm-up + calibration
    passes = max(1, min(50, int(budget_s / max(dt, 1e-3))))
    dt, n = ref.run(passes)
    return {'value': n / dt, 'unit': UNIT, 'cores': ref.cores, 'kind': ref.kind, 'sample': ref.describe(passes)}


def run_reference_arm(args, out=sys.stdout):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    ref = CpuReference(args.mechanism)
    dt, _ = ref.run(1)
    passes = max(1, min(20, int(1.0 / max(dt, 1e-3))))     # ~1 s of CPU work per step
    for _ in range(args.warmup):
        ref.run(passes)
    total_t, total_n = 0.0, 0
    for _ in range(args.steps):
        dt, n = ref.run(passes)
        total_t += dt
        total_n += n
    value = total_n / total_t
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_t / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, per_step_states=total_n // args.steps),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': ref.cores, 'kind': ref.kind,
                         'sample': ref.describe(passes)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), file=out, flush=True)
    return 0


def workload_config(args, per_step_states=None):
    return {'workload': f'BK1 (species production rates) + BK2 (mixture-averaged transport), GRI-Mech 3.0 '
                        f'(53 species / 325 reactions), FP64, {args.n_states} states per GPU, seeded synthetic '
                        f'states T~U[300,2500] K, p = 1 atm, normalised random Y',
            'mechanism': args.mechanism, 'n_states_per_gpu': args.n_states,
            'states_per_step': per_step_states if per_step_states is not None else args.n_states * args.gpus,
            'l2': 'inputs (>= 0.4 GB per launch) larger than the 126 MB L2; no flush needed',
            'parallelism': f'{args.gpus} x independent state shards, no collective'}


# ---------------------------------------------------------------------------------------------------
def _claim_stdout():
    """Only the final JSON line may reach stdout (libraries such as NCCL print banners there): keep the real
    stdout aside and point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    return real


def bind_to_gpu_numa_node(local):
    """Pin this rank's threads to the CPUs NVML reports as local to its GPU, BEFORE the pinned host buffers of the
    end-to-end leg are allocated (first touch puts them on that NUMA node): what `mpirun --bind-to` / numactl does
    for the reference's MPI ranks.  Returns a description for the JSON line, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        index = int(visible.split(',')[local]) if visible and visible.split(',')[local].isdigit() else local
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        cpus = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f'{len(cpus)} CPUs local to GPU {index}'
    except Exception:
        pass
    return None


class Bench:
    """per-process state of the CUDA arm: torch, the host mirror, rank / world, collectives for timing only"""

    def __init__(self, args):
        import torch
        import kinetix_b200.host as kinetix
        self.torch, self.kinetix, self.args = torch, kinetix, args
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device; the product has no CPU path (use --impl reference)')
        torch.cuda.set_device(self.local)
        self.all_cpus = os.sched_getaffinity(0) if hasattr(os, 'sched_getaffinity') else None
        self.numa = bind_to_gpu_numa_node(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
            self.dist = dist
        self.launches = 0

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device='cuda')
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device='cuda')
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def init_mechanism(self, mech, single_precision=False):
        """rank 0 generates / compiles a missing module first (kinetix.cpp:290-296), then everybody loads it"""
        k = self.kinetix
        path = os.path.join(ROOT, 'kinetix_b200', 'mechanisms', mech + '.yaml')
        if self.rank == 0:
            k.prepare(path, single_precision=single_precision)
        self.barrier()
        k.init(path, device_id=self.local, single_precision=single_precision)
        N = k.nSpecies()
        k.build(P_ATM, 1.0, [1.0 / N] * N, True)
        return N

    def synthetic_states(self, N, S, dtype=None, seed_offset=0):
        torch = self.torch
        gen = torch.Generator(device='cuda')
        gen.manual_seed(1234 + self.rank + 1000 * seed_offset)
        state = torch.empty((N + 1, S), dtype=torch.float64, device='cuda')
        state[0].uniform_(300.0, 2500.0, generator=gen)
        state[1:].uniform_(0.0, 1.0, generator=gen)
        state[1:] /= state[1:].sum(dim=0, keepdim=True)
        return state if dtype in (None, torch.float64) else state.to(dtype)

    def time_kernels(self, N, S, steps, warmup, state, storage_dtype=0, bk2=True, sampler=None):
        """`warmup` untimed + `steps` timed passes of BK1 (+ BK2) over `state`; CUDA events on the launching stream,
        barrier + synchronize on both sides, max over ranks.  Returns (ms per step, ms BK1, ms BK2, buffers)."""
        torch, k = self.torch, self.kinetix
        rates = torch.empty_like(state)
        visc = torch.empty(max(S, 1), dtype=state.dtype, device='cuda')
        cond = torch.empty_like(visc)
        rhoD = torch.empty((N, max(S, 1)), dtype=state.dtype, device='cuda') if bk2 else None

        def step(ev=None):
            if ev:
                ev[0].record()
            k.productionRates(S, S, S, 1.0, state, rates, dtype=storage_dtype)
            if ev:
                ev[1].record()
            if bk2:
                k.mixtureAvgTransportProps(S, S, S, 1.0, state, visc, cond, rhoD, dtype=storage_dtype)
            if ev:
                ev[2].record()

        for _ in range(warmup):
            step()
        self.barrier()
        if sampler is not None:
            sampler.start()
            time.sleep(0.3)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        self.barrier()
        for i in range(steps):
            step(evs[i])
        self.barrier()
        self.launches += steps * (2 if bk2 else 1)
        total_ms = evs[0][0].elapsed_time(evs[-1][2])
        t1 = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
        t2 = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
        total_ms, t1, t2 = self.max_over_ranks([total_ms, t1, t2])
        assert bool(torch.isfinite(rates[:, :1024].double()).all()), 'BK1 produced non-finite rates'
        if bk2:
            assert bool(torch.isfinite(rhoD[:, :1024].double()).all()), 'BK2 produced non-finite coefficients'
        return total_ms / steps, t1, t2, (rates, visc, cond, rhoD)


def module_counts(kinetix):
    """static SASS census written next to the loaded module by the build (kinetix_b200/sass.py); None if absent or
    if it does not belong to this module's source"""
    from kinetix_b200 import sass
    return sass.read_counts(os.path.dirname(kinetix.modulePath()))


def profiled_counts(kinetix):
    """executed FP64 instructions / DRAM traffic per state measured with ncu (tools/ncu_counts.py ->
    profiles/counts_r02.json), valid ONLY for the module source they were measured on"""
    from kinetix_b200 import sass
    tab = read_json(os.path.join(ROOT, 'profiles', 'counts_r02.json')) or {}
    try:
        sha = sass.source_hash(os.path.dirname(kinetix.modulePath()))
    except Exception:
        return None, 'module source not found'
    entry = tab.get(sha)
    if entry is None:
        return None, ('stale: profiles/counts_r02.json has no entry for this module source (kernels changed since the '
                      'last tools/ncu_counts.py run)')
    return entry, 'profiles/counts_r02.json (ncu, same module source)'


def fp64_roofline(kernel, rate, N, W_exec, W_ref, traffic_per_state, S, note):
    """FP64-pipe roofline of one kernel: achieved = W x states/s (as TFLOP/s, one lane instruction = one DFMA = 2
    flop) against the measured DFMA issue peak."""
    peak, peak_src = fp64_peak()
    hbm, hbm_src = hbm_peak()
    alg_bytes = {'bk1': 2 * (N + 1) * 8, 'bk2': (2 * N + 3) * 8}[kernel]
    out = {'bound': 'fp64', 'kernel': 'kx_bk1_f64' if kernel == 'bk1' else 'kx_bk2<double>', 'unit': 'TFLOP/s',
           'peak': peak * 2 / 1e12, 'peak_source': peak_src, 'states_per_s': rate,
           'fp64_lane_instr_per_state': W_exec, 'work_source': note,
           'traffic': traffic_per_state * S if traffic_per_state else None,
           'algorithmic_bytes_per_launch': alg_bytes * S,
           'hbm': {'achieved': rate * alg_bytes / 1e9, 'peak': hbm, 'unit': 'GB/s', 'frac': rate * alg_bytes / 1e9 / hbm,
                   'bytes_per_state': alg_bytes, 'peak_source': hbm_src}}
    if W_exec:
        out['achieved'] = rate * W_exec * 2 / 1e12
        out['frac'] = rate * W_exec / peak
        out['roofline_states_per_s'] = peak / W_exec
    else:
        out['achieved'] = out['frac'] = None
    if W_ref:
        out['frac_reference_W'] = rate * W_ref / peak
        out['reference_fp64_lane_instr_per_state'] = W_ref
    return out


def bandwidth_probe(b, nbytes=256 << 20, reps=4):
    """host<->device copy rate of this rank's GPU from pinned memory allocated on its NUMA-local cores: each rank
    ALONE (the others wait) and ALL ranks at once -- what bounds the end-to-end path when several GPUs share a host"""
    torch = b.torch
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device='cuda')

    def rate(direction):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if direction == 'h2d':
                d.copy_(h, non_blocking=True)
            else:
                h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        return nbytes * reps / (time.perf_counter() - t0) / 1e9

    rate('h2d'); rate('d2h')
    alone = [0.0, 0.0]
    for r in range(b.world):
        b.barrier()
        if r == b.rank:
            alone = [rate('h2d'), rate('d2h')]
    b.barrier()
    together = [rate('h2d'), rate('d2h')]
    b.barrier()
    # bidirectional, all ranks: what the pipelined e2e path actually does
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2 = h  # same buffer is fine for a rate probe
            h2.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    bidir = nbytes * reps / (time.perf_counter() - t0) / 1e9
    sums = b.sum_over_ranks(alone + together + [bidir])
    mins = [-x for x in b.max_over_ranks([-v for v in alone + together + [bidir]])]
    del h, d
    return {'unit': 'GB/s', 'bytes_per_copy': nbytes,
            'alone_h2d_mean': sums[0] / b.world, 'alone_d2h_mean': sums[1] / b.world,
            'alone_h2d_min': mins[0], 'alone_d2h_min': mins[1],
            'all_ranks_h2d_sum': sums[2], 'all_ranks_d2h_sum': sums[3],
            'all_ranks_h2d_min': mins[2], 'all_ranks_d2h_min': mins[3],
            'all_ranks_bidirectional_each_way_sum': sums[4], 'all_ranks_bidirectional_each_way_min': mins[4]}


def extra_configs(b, main, steps, warmup):
    """BASELINE.json configs 3, 4, 5 next to the headline (1, 2): same timing rules, fewer steps."""
    torch, k, args = b.torch, b.kinetix, b.args
    peak, _ = fp64_peak()
    hbm, _ = hbm_peak()
    out = []

    def entry(name, mech, S, sp, storage, bk2, note):
        N = b.init_mechanism(mech, single_precision=sp)
        tdt = torch.float32 if storage == 1 else torch.float64
        state = b.synthetic_states(N, S, dtype=tdt)
        ms, t1, t2, bufs = b.time_kernels(N, S, steps, warmup, state, storage_dtype=storage, bk2=bk2)
        counts = module_counts(k) or {}
        prof, _ = profiled_counts(k)
        size = 4 if storage == 1 else 8
        e = {'name': name, 'mechanism': mech, 'n_species': N, 'n_reactions': k.nReactions(),
             'precision': ('fp32 math, fp32 buffers' if storage == 1 else 'fp32 math, fp64 buffers (fpmix)') if sp else 'fp64',
             'states_per_gpu': S, 'n_gpus': b.world, 'steps': steps, 'warmup': warmup, 'note': note,
             'bk1_ms': t1, 'bk1_states_per_s': S * b.world / (t1 * 1e-3),
             'module': os.path.relpath(k.modulePath(), ROOT)}
        bytes1 = 2 * (N + 1) * size
        r1 = S / (t1 * 1e-3)
        e['bk1_hbm'] = {'achieved': r1 * bytes1 / 1e9, 'peak': hbm, 'unit': 'GB/s', 'frac': r1 * bytes1 / 1e9 / hbm,
                        'bytes_per_state': bytes1}
        c1 = counts.get('bk1') or {}
        if not sp and c1.get('fp64'):
            e['bk1_fp64'] = {'fp64_lane_instr_per_state': c1['fp64'], 'frac': r1 * c1['fp64'] / peak,
                             'work_source': 'static SASS count of the loaded module (straight-line kernel)'}
        if sp and c1.get('mufu'):
            e['bk1_mufu'] = {'mufu_lane_instr_per_state': c1['mufu'], 'frac': r1 * c1['mufu'] / mufu_peak(),
                             'fp32_lane_instr_per_state': c1.get('ffma')}
        e['bk1_bound'] = max((('hbm', e['bk1_hbm']['frac']), ('fp64', e.get('bk1_fp64', {}).get('frac', 0)),
                              ('mufu', e.get('bk1_mufu', {}).get('frac', 0))), key=lambda x: x[1])[0]
        if bk2:
            r2 = S / (t2 * 1e-3)
            bytes2 = (2 * N + 3) * size
            e.update({'bk2_ms': t2, 'bk2_states_per_s': S * b.world / (t2 * 1e-3), 'ms_per_step': ms,
                      'states_per_s': S * b.world / (ms * 1e-3),
                      'bk2_hbm': {'achieved': r2 * bytes2 / 1e9, 'peak': hbm, 'unit': 'GB/s',
                                  'frac': r2 * bytes2 / 1e9 / hbm, 'bytes_per_state': bytes2}})
            if prof and prof.get('bk2_fp64_per_state'):
                e['bk2_fp64'] = {'fp64_lane_instr_per_state': prof['bk2_fp64_per_state'],
                                 'frac': r2 * prof['bk2_fp64_per_state'] / peak,
                                 'work_source': 'ncu smsp__inst_executed_pipe_fp64 on this module source'}
        del state, bufs
        torch.cuda.empty_cache()
        out.append(e)

    free = torch.cuda.mem_get_info()[0]
    entry('config 3: small H2/O2 mechanism, BK1, 8 Mi states per GPU (64 Mi on 8), FP64', 'LiDryer', 1 << 23, False, 0, False,
          'bandwidth-leaning regime starts with FP32 math (next entry)')
    entry('config 3 (FP32 math, FP64 buffers)', 'LiDryer', 1 << 23, True, 0, False, 'HBM-bound: see bk1_hbm.frac')
    etoh_S = 1 << 24
    need = 3 * 130 * etoh_S * 8 * 1.05
    if free < need:
        etoh_S = 1 << 22
    entry('config 4: largest shipped hydrocarbon mechanism, BK1 + BK2, 16 Mi states per GPU, FP64', 'EtOHKonnov', etoh_S,
          False, 0, True, '' if etoh_S == 1 << 24 else 'reduced to 4 Mi states: not enough free device memory')
    entry('config 5: GRI-3.0 BK1, FP32 math with FP64 buffers (fpmix), 16 Mi states per GPU', 'gri30', args.n_states, True, 0,
          False, 'FP64 point of the sweep = the headline bk1_states_per_s')
    entry('config 5: GRI-3.0 BK1, FP32 math with FP32 buffers, 16 Mi states per GPU', 'gri30', args.n_states, True, 1,
          False, '')
    # strong scaling: a FIXED total of 16 Mi GRI-3.0 states cut into contiguous per-rank shards
    from kinetix_b200.sharding import shard_of
    total = 1 << 24
    lo, hi = shard_of(total, b.rank, b.world)
    N = b.init_mechanism('gri30')
    state = b.synthetic_states(N, hi - lo, seed_offset=7)
    ms, t1, t2, bufs = b.time_kernels(N, hi - lo, steps, warmup, state)
    out.append({'name': 'strong scaling: 16 Mi GRI-3.0 states in total, BK1 + BK2, FP64', 'mechanism': 'gri30',
                'scaling': 'strong', 'total_states': total, 'n_gpus': b.world, 'states_this_rank': hi - lo,
                'ms_per_step': ms, 'states_per_s': total / (ms * 1e-3), 'bk1_states_per_s': total / (t1 * 1e-3),
                'bk2_states_per_s': total / (t2 * 1e-3), 'steps': steps, 'warmup': warmup,
                'partition': 'kinetix_b200.sharding.shard_of: contiguous ranges, remainder spread over the first ranks'})
    del state, bufs
    torch.cuda.empty_cache()
    # what a single fused BK1+BK2 kernel could save at most: it would read the state slab once instead
    # of twice.  Measured, not estimated: the thermo kernel streams exactly that slab (reads (N+1) rows, writes N+2) at
    # HBM speed; its read share is the ceiling of the saving, and only if none of it were already hidden under FP64 work.
    S = args.n_states
    state = b.synthetic_states(N, S, seed_offset=9)
    rho = torch.empty(S, dtype=torch.float64, device='cuda')
    rcp = torch.empty_like(rho)
    cp = torch.empty((N, S), dtype=torch.float64, device='cuda')
    for _ in range(3):
        k.thermodynamicProps(S, S, S, 1.0, state, rho, cp, rcp)
    b.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        k.thermodynamicProps(S, S, S, 1.0, state, rho, cp, rcp)
    e1.record()
    b.barrier()
    b.launches += steps
    t_th = b.max_over_ranks([e0.elapsed_time(e1) / steps])[0]
    read_share = (N + 1) / (2 * N + 3)
    out.append({'name': 'fusion bound: state slab streamed once (thermo kernel, GRI-3.0, same states per GPU)',
                'mechanism': 'gri30', 'states_per_gpu': S, 'thermo_ms': t_th,
                'thermo_gb_per_s': (2 * N + 3) * 8 * S / (t_th * 1e-3) / 1e9,
                'state_read_ms_upper_bound': t_th * read_share,
                'note': 'a fused BK1+BK2 kernel saves at most state_read_ms_upper_bound per step (compare ms_per_step)'})
    del state, rho, rcp, cp
    torch.cuda.empty_cache()
    return out


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--mechanism', default='gri30')
    ap.add_argument('--n-states', type=int, default=1 << 24, help='states per GPU')
    ap.add_argument('--e2e-states', type=int, default=1 << 22, help='states per GPU per end-to-end step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--configs', default='all', choices=['all', 'none'],
                    help="'none' skips the extra BASELINE configurations (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    if args.impl == 'reference':
        return run_reference_arm(args, real_stdout)

    b = Bench(args)
    torch, kinetix, rank, world = b.torch, b.kinetix, b.rank, b.world
    N = b.init_mechanism(args.mechanism)

    # ---- headline: synthetic states resident in HBM (this rank's shard), BK1 + BK2 ----
    S = args.n_states
    state = b.synthetic_states(N, S)
    sampler = ClockSampler(b.local) if rank == 0 else None
    ms_per_step, t_bk1, t_bk2, (rates, visc, cond, rhoD) = b.time_kernels(N, S, args.steps, args.warmup, state, sampler=sampler)
    clocks = sampler.stop() if rank == 0 else None
    value = S * world / (ms_per_step * 1e-3)
    headline_launches = b.launches
    counts = module_counts(kinetix)
    prof, prof_src = profiled_counts(kinetix)
    module_path = kinetix.modulePath()

    # ---- end to end through the C ABI with HOST buffers (H2D + kernels + D2H inside the timed region) ----
    Se = min(args.e2e_states, S)
    h_state = state[:, :Se].contiguous().cpu().pin_memory()
    h_rates = torch.empty_like(h_state).pin_memory()
    h_visc = torch.empty(Se, dtype=torch.float64).pin_memory()
    h_cond = torch.empty(Se, dtype=torch.float64).pin_memory()
    h_rhoD = torch.empty((N, Se), dtype=torch.float64).pin_memory()

    def e2e_two_calls():
        kinetix.productionRatesHost(Se, Se, Se, 1.0, h_state, h_rates)
        kinetix.mixtureAvgTransportPropsHost(Se, Se, Se, 1.0, h_state, h_visc, h_cond, h_rhoD)

    def e2e_one_upload():
        kinetix.ratesAndTransportHost(Se, Se, Se, 1.0, h_state, h_rates, h_visc, h_cond, h_rhoD)

    e2e_steps = max(3, min(args.steps, 5))

    def time_e2e(fn):
        fn()
        b.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        b.barrier()
        t = b.max_over_ranks([time.perf_counter() - t0])[0]
        return Se * world * e2e_steps / t

    e2e_two = time_e2e(e2e_two_calls)
    e2e_one = time_e2e(e2e_one_upload)
    assert torch.equal(h_rates[:, :256], rates[:, :256].cpu())
    probe = bandwidth_probe(b)
    del h_state, h_rates, h_visc, h_cond, h_rhoD, state, rates, visc, cond, rhoD
    torch.cuda.empty_cache()

    configs = None
    if args.configs == 'all':
        try:
            configs = extra_configs(b, None, steps=max(3, min(args.steps, 5)), warmup=3)
        except Exception as e:          # the extra configurations never take the headline down with them
            configs = [{'error': f'{type(e).__name__}: {e}'}]

    if rank != 0:
        if b.dist is not None:
            b.dist.destroy_process_group()
        return 0

    bk1_rate = S / (t_bk1 * 1e-3)
    bk2_rate = S / (t_bk2 * 1e-3)
    dominant = 'bk2' if t_bk2 >= t_bk1 else 'bk1'
    W_ref = W_REFERENCE.get(args.mechanism, {})
    W1 = (counts or {}).get('bk1', {}).get('fp64') if counts else None
    W2 = prof.get('bk2_fp64_per_state') if prof else None
    tr1 = prof.get('bk1_dram_bytes_per_state') if prof else None
    tr2 = prof.get('bk2_dram_bytes_per_state') if prof else None
    roof = {
        'bk1': fp64_roofline('bk1', bk1_rate, N, W1, W_ref.get('bk1'), tr1, S,
                             'static SASS count of the loaded module (straight-line kernel: static = executed)'
                             if W1 else 'unknown: counts.json missing beside the module'),
        'bk2': fp64_roofline('bk2', bk2_rate, N, W2, W_ref.get('bk2'), tr2, S, prof_src),
    }
    # the four-warp BK1 layout re-reads concentrations from shared memory and no longer shares their products across
    # reactions: 2 % more FP64 instructions than the classic kernel carried in the same module.  Quote the fraction by
    # that smaller count as well, so that re-computed work cannot flatter the headline fraction.
    W1_min = ((counts or {}).get('bk1_small') or {}).get('fp64')
    if W1_min and W1 and W1_min < W1 and roof['bk1'].get('frac'):
        roof['bk1']['fp64_lane_instr_per_state_minimal_form'] = W1_min
        roof['bk1']['frac_minimal_form'] = roof['bk1']['frac'] * W1_min / W1
    other = 'bk1' if dominant == 'bk2' else 'bk2'
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args),
        'bk1_states_per_s': bk1_rate * world, 'bk2_states_per_s': bk2_rate * world,
        'bk1_ms': t_bk1, 'bk2_ms': t_bk2,
        'bk1_grxn_per_s': bk1_rate * world * 325 / 1e9 if args.mechanism == 'gri30' else None,
        'bk2_gdof_per_s': bk2_rate * world * (N + 2) / 1e9,
        'roofline': roof[dominant],
        'roofline_other': roof[other],
        'e2e': {'value': e2e_one, 'unit': UNIT,
                'h2d_bytes_per_step': (N + 1) * Se * 8 * world, 'd2h_bytes_per_step': ((N + 1) + (N + 2)) * Se * 8 * world,
                'states_per_step': Se * world,
                'api': 'kx_rates_and_transport_host: BK1 + BK2 of the same host-resident states, ONE upload of the state '
                       'slab, 128 Ki-state chunks pipelined H2D | kernels | D2H on 4 streams',
                'two_call_value': e2e_two,
                'two_call_api': 'kx_production_rates_host + kx_mixture_avg_transport_props_host (the reference\'s two calls: '
                                'the state slab is uploaded twice)',
                'two_call_h2d_bytes_per_step': 2 * (N + 1) * Se * 8 * world,
                'bytes_are': 'whole job (all ranks)',
                'host_link_probe': probe},
        'gpu_launches': headline_launches,
        'gpu_launches_all_configs': b.launches,
        'clocks': clocks,
        'module': os.path.relpath(module_path, ROOT),
        'host_binding': b.numa,
        'configs': configs,
    }
    if world == 1 and not args.no_cpu_baseline:
        if b.all_cpus:
            os.sched_setaffinity(0, b.all_cpus)      # the CPU baseline runs on ALL host cores, not the GPU-local ones
        try:
            line['cpu_baseline'] = cpu_baseline(args.mechanism)
        except Exception as e:     # the baseline is reported, never required for the product number
            line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'unavailable', 'sample': str(e)}
    print(json.dumps(line), file=real_stdout, flush=True)
    if b.dist is not None:
        b.dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
