for m in LiDryer NH3Konnov_edit chempolimi_edit gri30-20; do
  n=4194304; [ $m = LiDryer ] && n=8388608
  for v in old single_lr; do timeout 120 python tools/quick_time.py --mech $m --n $n --cache build/variants/$v --tag $v --check 2>&1 | tail -1; done
  timeout 120 python tools/quick_time.py --mech $m --n $n --tag default --check 2>&1 | tail -1
done
timeout 200 python tools/quick_time.py --mech gri30 --tag default --check 2>&1 | tail -1
timeout 200 python tools/quick_time.py --mech EtOHKonnov --n 1048576 --tag default --check 2>&1 | tail -1
timeout 200 python tools/quick_time.py --mech heptaneLu88 --n 1048576 --tag default --check 2>&1 | tail -1
