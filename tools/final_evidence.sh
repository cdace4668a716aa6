#!/bin/bash
# Round-end evidence on ONE B200: bench line, reference arm, ncu launch list of the bench command, full captures of
# the two kernels (summaries only: the .ncu-rep files stay on the box), sweeps for BASELINE configs 3-5.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --n-states 4194304 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/launches.csv 'ncu --metrics gpu__time_duration.sum -c 400 python bench.py --steps 2 --warmup 3 --n-states 4194304' > gpurun_out/launches_summary.txt
for k in bk1 bk2; do
  timeout 300 ncu --set full --clock-control none -k regex:kx_$k -c 1 -o /tmp/full_$k python tools/quick_time.py --mech gri30 --n 4194304 --reps 1 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/full_$k.ncu-rep > gpurun_out/ncu_full_$k.txt 2>&1
done
python tools/sweep.py --mech LiDryer --modes f64,fpmix,f32 --min 8388608 --max 8388608 --out gpurun_out/sweep_lidryer.jsonl > /dev/null 2>&1
python tools/sweep.py --mech EtOHKonnov --modes f64 --min 16777216 --max 16777216 --out gpurun_out/sweep_etoh.jsonl > /dev/null 2>&1
python tools/sweep.py --mech heptaneLu88 --modes f64 --min 16777216 --max 16777216 --out gpurun_out/sweep_heptane.jsonl > /dev/null 2>&1
python tools/sweep.py --mech gri30 --modes f64,fpmix,f32 --min 134217728 --max 134217728 --kernels bk1 --out gpurun_out/sweep_gri128m.jsonl > /dev/null 2>&1
cat gpurun_out/bench_n1.json | cut -c1-300
cat gpurun_out/launches_summary.txt
