#!/bin/bash
# Role-isolation timing of the strip window kernel (dev build `make -C tidal-wave_b200/csrc sdbg`): gauss_iter ms/step with the
# V taps, H + solve, U switched off in turn.  usage (through gpurun): tools/strip_roles.sh  -> stdout
export TW_LIB=$PWD/tidal-wave_b200/libtidalwave_b200_sdbg.so TW_WINDOW=strip
for d in 0 11 10 9 3 8 1 2; do
  TW_STRIP_DBG=$d python bench.py --no-cpu --no-e2e --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['roofline']['families']
print('dbg=$d gauss_iter ms/step %.3f  gauss_last %.3f  value %.0f' % (f['gauss_iter']['ms_per_step'], f['gauss_last']['ms_per_step'], d['value']))"
done
