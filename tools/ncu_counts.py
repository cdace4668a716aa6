#!/usr/bin/env python3
"""Executed FP64 instructions and DRAM traffic per state of the shipped kernels, measured with ncu, keyed by the
SHA-256 of the module source they were measured on -> profiles/counts_r02.json (bench.py refuses the entry when the
loaded module's source differs: the roofline line cannot quote the work of another build).

  on the GPU box:
    ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/counts.csv python tools/ncu_counts.py --run [mech ...]
    python tools/ncu_counts.py --parse gpurun_out/counts.csv gpurun_out/counts_run.json >> merges into profiles/counts_r02.json
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P_ATM = 101325.0
STATES = {'gri30': 1 << 21, 'EtOHKonnov': 1 << 19, 'LiDryer': 1 << 22, 'heptaneLu88': 1 << 20}


def run(mechs):
    """one BK1 and one BK2 launch per mechanism (run this under ncu); writes which module / how many states"""
    import torch
    import kinetix_b200.host as kx
    from kinetix_b200 import sass
    meta = []
    for mech in mechs:
        kx.init(os.path.join(ROOT, 'kinetix_b200', 'mechanisms', mech + '.yaml'))
        N = kx.nSpecies()
        kx.build(P_ATM, 1.0, [1.0 / N] * N, True)
        S = STATES.get(mech, 1 << 20)
        gen = torch.Generator(device='cuda')
        gen.manual_seed(7)
        st = torch.empty((N + 1, S), dtype=torch.float64, device='cuda')
        st[0].uniform_(300.0, 2500.0, generator=gen)
        st[1:].uniform_(0.0, 1.0, generator=gen)
        st[1:] /= st[1:].sum(dim=0, keepdim=True)
        rates = torch.empty_like(st)
        visc = torch.empty(S, dtype=torch.float64, device='cuda')
        cond = torch.empty_like(visc)
        rhoD = torch.empty((N, S), dtype=torch.float64, device='cuda')
        torch.cuda.synchronize()
        kx.productionRates(S, S, S, 1.0, st, rates)
        kx.mixtureAvgTransportProps(S, S, S, 1.0, st, visc, cond, rhoD)
        torch.cuda.synchronize()
        d = os.path.dirname(kx.modulePath())
        meta.append(dict(mechanism=mech, states=S, n_species=N, source_sha256=sass.source_hash(d),
                         module=os.path.relpath(kx.modulePath(), ROOT)))
        kx.finalize()
        del st, rates, visc, cond, rhoD
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'counts_run.json'), 'w') as fh:
        json.dump(meta, fh, indent=1)


def parse(csv_path, meta_path, out_path):
    meta = json.load(open(meta_path))
    rows = []
    with open(csv_path) as fh:
        lines = [ln for ln in fh if not ln.startswith('==')]
    for r in csv.DictReader(lines):
        name = r.get('Kernel Name', '')
        if 'kx_bk1' in name or 'kx_bk2' in name:
            rows.append(r)
    # launches appear in order: per mechanism BK1 then BK2; one CSV row per (launch, metric)
    launches, seen = [], {}
    for r in rows:
        key = r['ID']
        if key not in seen:
            seen[key] = dict(kernel=r['Kernel Name'])
            launches.append(seen[key])
        seen[key][r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
        seen[key]['unit:' + r['Metric Name']] = r.get('Metric Unit', '')
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    it = iter(launches)
    for m in meta:
        k1, k2 = next(it), next(it)
        assert 'kx_bk1' in k1['kernel'] and 'kx_bk2' in k2['kernel'], (k1['kernel'], k2['kernel'])
        S = m['states']

        def dram(k):
            tot = 0.0
            for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                v, u = k[name], k['unit:' + name].lower()
                tot += v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
            return tot

        def dur(k):
            v, u = k['gpu__time_duration.sum'], k['unit:gpu__time_duration.sum'].lower()
            return v * {'ns': 1e-9, 'us': 1e-6, 'usecond': 1e-6, 'ms': 1e-3, 'msecond': 1e-3, 'nsecond': 1e-9, 's': 1}.get(u, 1e-9)
        table[m['source_sha256']] = dict(
            mechanism=m['mechanism'], module=m['module'], states_measured=S,
            bk1_fp64_per_state=k1['smsp__inst_executed_pipe_fp64.sum'] * 32 / S,
            bk1_instr_per_state=k1['smsp__inst_executed.sum'] * 32 / S,
            bk1_dram_bytes_per_state=dram(k1) / S, bk1_ncu_seconds=dur(k1),
            bk2_fp64_per_state=k2['smsp__inst_executed_pipe_fp64.sum'] * 32 / S,
            bk2_instr_per_state=k2['smsp__inst_executed.sum'] * 32 / S,
            bk2_dram_bytes_per_state=dram(k2) / S, bk2_ncu_seconds=dur(k2),
            source='ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_*.sum (tools/ncu_counts.py)')
        print(m['mechanism'], json.dumps(table[m['source_sha256']]))
    with open(out_path, 'w') as fh:
        json.dump(table, fh, indent=1)


if __name__ == '__main__':
    if sys.argv[1] == '--run':
        run(sys.argv[2:] or ['gri30', 'EtOHKonnov', 'LiDryer', 'heptaneLu88'])
    else:
        parse(sys.argv[2], sys.argv[3], os.path.join(ROOT, 'profiles', 'counts_r02.json'))
