#!/bin/bash
# per-phase FP64-pipe utilisation of the heptaneLu88 BK2 kernel (one state per thread, 256 threads)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kx_bk2 -c 1 -o /tmp/full_hept_bk2 python tools/quick_time.py --mech heptaneLu88 --n 2097152 --reps 1 > /dev/null 2>&1
python tools/ncu_phase_util.py /tmp/full_hept_bk2.ncu-rep 334:batch_start 340:mole_fractions 406:wilke_setup 419:wilke_pass1 495:conductivity 523:wilke_pass2 595:viscosity 607:diffusion_setup 625:row_block_load 654:tiles 734:diagonal_tile 762:rhoD_store 797:prefetch > gpurun_out/r02ao_phases_heptane_bk2.txt 2>&1
cat gpurun_out/r02ao_phases_heptane_bk2.txt
python tools/ncu_summary.py /tmp/full_hept_bk2.ncu-rep | tail -12
python tools/ncu_lines.py /tmp/full_hept_bk2.ncu-rep 2>/dev/null | head -14
