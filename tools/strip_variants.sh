#!/bin/bash
# Times the experiment builds of the strip window kernel (make -C tidal-wave_b200/csrc var NAME=.. VARFLAGS=..): gauss_iter ms/step.
export TW_WINDOW=strip
for so in tidal-wave_b200/libtidalwave_b200_var_*.so; do
  n=${so##*_var_}; n=${n%.so}
  TW_LIB=$PWD/$so python -m pytest tests/test_gpu_parity.py -x -q -k strip 2>&1 | tail -1
  TW_LIB=$PWD/$so python bench.py --no-cpu --no-e2e --steps 60 ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['roofline']['families']
print('$n gauss_iter ms/step %.3f  gauss_last %.3f  value %.0f' % (f['gauss_iter']['ms_per_step'], f['gauss_last']['ms_per_step'], d['value']))"
done
