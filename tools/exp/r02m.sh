#!/bin/bash
# per-phase FP64-pipe utilisation of the GRI-3.0 BK2 kernel (source-level samples of one full capture)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kx_bk2 -c 1 -o /tmp/full_gri_bk2 python tools/quick_time.py --mech gri30 --n 4194304 --reps 1 --cache build/variants/swp > /dev/null 2>&1
python tools/ncu_phase_util.py /tmp/full_gri_bk2.ncu-rep 334:batch_start 340:mole_fractions 406:wilke_setup 419:wilke_pass1 495:conductivity 523:wilke_pass2 595:viscosity 607:diffusion_setup 625:row_block_load 654:tiles 732:diagonal_tile 755:rhoD_store 780:prefetch > gpurun_out/r02m_phases_gri_bk2.txt 2>&1
cat gpurun_out/r02m_phases_gri_bk2.txt
python tools/ncu_lines.py /tmp/full_gri_bk2.ncu-rep 2>/dev/null | head -40 > gpurun_out/r02m_lines_gri_bk2.txt
cat gpurun_out/r02m_lines_gri_bk2.txt
