#!/bin/bash
# GPU call r02a: full GPU test-suite, the new bench line, executed-work table (ncu), first kernel experiments.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:kx_bk --csv --log-file gpurun_out/r02a_counts.csv python tools/ncu_counts.py --run > gpurun_out/r02a_counts.log 2>&1
for v in base expcall expcall_s8 cold40 expcall_cold40 bk2_128; do
  timeout 300 python tools/quick_time.py --mech EtOHKonnov --n 4194304 --reps 3 --cache build/variants/$v --tag $v --check >> gpurun_out/r02a_variants.log 2>&1
done
for v in base expcall cold40 bk2_320 bk2_384; do
  timeout 300 python tools/quick_time.py --mech heptaneLu88 --n 4194304 --reps 3 --cache build/variants/$v --tag $v --check >> gpurun_out/r02a_variants.log 2>&1
done
for v in base expcall expcall_s32 expcall_4cta bk2_st2 bk2_fullrcp; do
  timeout 300 python tools/quick_time.py --mech gri30 --n 8388608 --reps 5 --cache build/variants/$v --tag $v --check >> gpurun_out/r02a_variants.log 2>&1
done
# full capture of the restructured BK2 kernel (GRI-3.0) and of EtOHKonnov BK1 / BK2
for spec in "gri30 bk2 2097152" "EtOHKonnov bk1 524288" "EtOHKonnov bk2 524288"; do
  set -- $spec
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:kx_$2 -c 1 -o /tmp/full_$1_$2 python tools/quick_time.py --mech $1 --n $3 --reps 1 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/full_$1_$2.ncu-rep > gpurun_out/r02a_ncu_$1_$2.txt 2>&1
  python tools/ncu_lines.py /tmp/full_$1_$2.ncu-rep > gpurun_out/r02a_lines_$1_$2.txt 2>&1
  python tools/ncu_stalls.py /tmp/full_$1_$2.ncu-rep > gpurun_out/r02a_stalls_$1_$2.txt 2>&1
done
cat gpurun_out/r02a_variants.log | grep -v "^$" | cut -c1-220
cut -c1-700 gpurun_out/r02a_bench.json
