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

# Algorithmic FP64-pipe work per state (lane instructions: DFMA/DMUL/DADD), minimal form with the
# reference's libdevice-class costs -- SURVEY.md 8(d), derivation in DESIGN.md "Roofline".
W_FP64 = {'gri30': {'bk1': 1.4e4, 'bk2': 2.47e4}}
# FP64-pipe lane instructions our kernels actually EXECUTE per state (ncu smsp__inst_executed_pipe_fp64 x 32 / states,
# profiles/ncu_r01_bk1_v6.txt, ncu_r01_bk2_v4.txt): fewer than W (shared exps in BK1; rank-12 Wilke factorisation and
# the 2-instruction pair reciprocal in BK2), so `frac` (by W) exceeds the hardware-side pipe utilisation,
# which is reported next to it as `frac_executed`.
EXEC_FP64 = {'gri30': {'bk1': 1.13e4, 'bk2': 1.82e4}}


def read_json(path):
    try:
        with open(path) as fh:
            return json.load(fh)
    except Exception:
        return None


def fp64_peak():
    """Measured DFMA issue peak of this pool's B200 [lane-instr/s] (tools/peaks.cu -> profiles/peaks_r01.json);
    nominal 148 SM x 64 lanes x 1.965 GHz if the measurement is absent."""
    pk = read_json(os.path.join(ROOT, 'profiles', 'peaks_r01.json'))
    if pk and 'dfma_lane_instr_per_s_sustained_3s' in pk:
        return float(pk['dfma_lane_instr_per_s_sustained_3s']), 'measured (profiles/peaks_r01.json)'
    return 148 * 64 * 1.965e9, 'nominal'


def hbm_peak():
    pk = read_json(os.path.join(ROOT, 'MEASURED_PEAKS.json'))
    if pk and 'hbm_gbs' in pk:
        return float(pk['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    if args.impl == 'reference':
        return run_reference_arm(args, real_stdout)

    import torch
    import kinetix_b200.host as kinetix

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU path (use --impl reference)')
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0) if hasattr(os, 'sched_getaffinity') else None
    numa = bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    mech_yaml = os.path.join(ROOT, 'kinetix_b200', 'mechanisms', args.mechanism + '.yaml')
    kinetix.init(mech_yaml, device_id=local)
    N = kinetix.nSpecies()
    kinetix.build(P_ATM, 1.0, [1.0 / N] * N, True)

    # ---- synthetic states, resident in HBM (this rank's shard) ----
    S = args.n_states
    gen = torch.Generator(device='cuda')
    gen.manual_seed(1234 + rank)
    state = torch.empty((N + 1, S), dtype=torch.float64, device='cuda')
    state[0].uniform_(300.0, 2500.0, generator=gen)
    state[1:].uniform_(0.0, 1.0, generator=gen)
    state[1:] /= state[1:].sum(dim=0, keepdim=True)
    rates = torch.empty_like(state)
    visc = torch.empty(S, dtype=torch.float64, device='cuda')
    cond = torch.empty_like(visc)
    rhoD = torch.empty((N, S), dtype=torch.float64, device='cuda')

    def step(ev=None):
        if ev:
            ev[0].record()
        kinetix.productionRates(S, S, S, 1.0, state, rates)
        if ev:
            ev[1].record()
        kinetix.mixtureAvgTransportProps(S, S, S, 1.0, state, visc, cond, rhoD)
        if ev:
            ev[2].record()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        step(evs[i])
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = evs[0][0].elapsed_time(evs[-1][2])
    t_bk1 = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    t_bk2 = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    tt = torch.tensor([total_ms, t_bk1, t_bk2], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, t_bk1, t_bk2 = tt.tolist()
    ms_per_step = total_ms / args.steps
    value = S * world / (ms_per_step * 1e-3)

    # sanity: results are finite (catches a kernel that did not run)
    assert bool(torch.isfinite(rates[:, :1024]).all()) and bool(torch.isfinite(rhoD[:, :1024]).all())

    # ---- end to end through the C ABI with HOST buffers (H2D + kernels + D2H inside the timed region) ----
    Se = min(args.e2e_states, S)
    h_state = state[:, :Se].contiguous().cpu().pin_memory()
    h_rates = torch.empty_like(h_state).pin_memory()
    h_visc = torch.empty(Se, dtype=torch.float64).pin_memory()
    h_cond = torch.empty(Se, dtype=torch.float64).pin_memory()
    h_rhoD = torch.empty((N, Se), dtype=torch.float64).pin_memory()

    def e2e_step():
        kinetix.productionRatesHost(Se, Se, Se, 1.0, h_state, h_rates)
        kinetix.mixtureAvgTransportPropsHost(Se, Se, Se, 1.0, h_state, h_visc, h_cond, h_rhoD)

    def e2e_fused_step():
        kinetix.ratesAndTransportHost(Se, Se, Se, 1.0, h_state, h_rates, h_visc, h_cond, h_rhoD)

    e2e_steps = max(3, min(args.steps, 5))

    def time_e2e(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return Se * world * e2e_steps / float(t.item())

    e2e_value = time_e2e(e2e_step)
    e2e_fused = time_e2e(e2e_fused_step)
    h2d = 2 * (N + 1) * Se * 8                      # the state slab is uploaded once per kernel
    d2h = ((N + 1) + (N + 2)) * Se * 8
    assert torch.equal(h_rates[:, :256], rates[:, :256].cpu())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = fp64_peak()
    hbm, hbm_src = hbm_peak()
    W = W_FP64.get(args.mechanism, {'bk1': float('nan'), 'bk2': float('nan')})
    bk1_rate = S / (t_bk1 * 1e-3)
    bk2_rate = S / (t_bk2 * 1e-3)
    dominant = 'bk2' if t_bk2 >= t_bk1 else 'bk1'
    dom_rate, dom_t = (bk2_rate, t_bk2) if dominant == 'bk2' else (bk1_rate, t_bk1)
    alg_bytes = {'bk1': 2 * (N + 1) * 8, 'bk2': (2 * N + 3) * 8}

    traffic_tab = read_json(os.path.join(ROOT, 'profiles', 'traffic_r01.json')) or {}

    def traffic(kernel):
        """DRAM bytes per launch from the committed ncu --set full capture, scaled to this launch's state count"""
        t = traffic_tab.get({'bk1': 'kx_bk1_f64', 'bk2': 'kx_bk2'}[kernel]) if args.mechanism == 'gri30' else None
        if not t:
            return None
        return (t['dram_read_bytes'] + t['dram_write_bytes']) / t['states'] * S

    def fp64_roofline(kernel, rate):
        ach = rate * W[kernel] * 2 / 1e12        # FP64 TFLOP/s counting one lane instruction as 2 flop (DFMA)
        pk = peak * 2 / 1e12
        return {'bound': 'fp64', 'kernel': 'kx_bk1_f64' if kernel == 'bk1' else 'kx_bk2<double>', 'achieved': ach, 'peak': pk, 'unit': 'TFLOP/s',
                'frac': ach / pk, 'traffic': traffic(kernel), 'fp64_lane_instr_per_state': W[kernel],
                'executed_fp64_lane_instr_per_state': EXEC_FP64.get(args.mechanism, {}).get(kernel),
                'frac_executed': (rate * EXEC_FP64[args.mechanism][kernel] / peak
                                  if args.mechanism in EXEC_FP64 else None),
                'states_per_s': rate, 'roofline_states_per_s': peak / W[kernel], 'peak_source': peak_src,
                'hbm': {'achieved': rate * alg_bytes[kernel] / 1e9, 'peak': hbm, 'unit': 'GB/s',
                        'frac': rate * alg_bytes[kernel] / 1e9 / hbm, 'bytes_per_state': alg_bytes[kernel],
                        'peak_source': hbm_src}}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args),
        'bk1_states_per_s': bk1_rate * world, 'bk2_states_per_s': bk2_rate * world,
        'bk1_ms': t_bk1, 'bk2_ms': t_bk2,
        'bk1_grxn_per_s': bk1_rate * world * kinetix.nReactions() / 1e9,
        'bk2_gdof_per_s': bk2_rate * world * (N + 2) / 1e9,
        'roofline': fp64_roofline(dominant, dom_rate),
        'roofline_other': fp64_roofline('bk1' if dominant == 'bk2' else 'bk2',
                                        bk1_rate if dominant == 'bk2' else bk2_rate),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'states_per_step': Se * world, 'api': 'kx_production_rates_host + kx_mixture_avg_transport_props_host',
                'fused_value': e2e_fused, 'fused_api': 'kx_rates_and_transport_host (one state upload)',
                'fused_h2d_bytes_per_step': (N + 1) * Se * 8},
        'gpu_launches': 2 * args.steps,
        'clocks': clocks,
        'module': os.path.relpath(kinetix.modulePath(), ROOT),
        'host_binding': numa,
    }
    if world == 1 and not args.no_cpu_baseline:
        if all_cpus:
            os.sched_setaffinity(0, all_cpus)      # the CPU baseline runs on ALL host cores, not the GPU-local ones
        try:
            line['cpu_baseline'] = cpu_baseline(args.mechanism)
        except Exception as e:     # the baseline is reported, never required for the product number
            line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'unavailable', 'sample': str(e)}
    print(json.dumps(line), file=real_stdout, flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
