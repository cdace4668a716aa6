#!/bin/bash
# the kinetix_bk driver's own output lines (reference CLI, reference units) for the record
mkdir -p gpurun_out
Y=kinetix_b200/mechanisms/gri30.yaml
L=gpurun_out/r02aj_kinetix_bk.log; : > $L
for mode in 1 2; do
  echo "\$ ./benchmark/kinetix_bk --backend CUDA --yaml-file $Y --mode $mode --n-states 16777216 --n-repetitions 20 --random-states" >> $L
  timeout 300 ./benchmark/kinetix_bk --backend CUDA --yaml-file $Y --mode $mode --n-states 16777216 --n-repetitions 20 --random-states >> $L 2>&1
done
echo "\$ ./benchmark/kinetix_bk --backend CUDA --yaml-file $Y --mode 1 --n-states 1000000 --n-repetitions 50   (BASELINE config 1: the stock benchmark's identical states)" >> $L
timeout 300 ./benchmark/kinetix_bk --backend CUDA --yaml-file $Y --mode 1 --n-states 1000000 --n-repetitions 50 >> $L 2>&1
echo "\$ ./benchmark/kinetix_bk --backend CUDA --yaml-file $Y --mode 2 --n-states 1000000 --n-repetitions 50   (BASELINE config 2)" >> $L
timeout 300 ./benchmark/kinetix_bk --backend CUDA --yaml-file $Y --mode 2 --n-states 1000000 --n-repetitions 50 >> $L 2>&1
echo "\$ ./benchmark/kinetix_bk --backend CUDA --yaml-file $Y --cimode 1" >> $L
timeout 300 ./benchmark/kinetix_bk --backend CUDA --yaml-file $Y --cimode 1 >> $L 2>&1; echo "exit status $?" >> $L
cat $L
