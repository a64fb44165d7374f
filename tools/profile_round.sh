#!/bin/bash
# Round-end measurement set on a B200 box (run through gpurun from the repo root): tests, bench, parity report, ncu launch list,
# ncu --set full of one step (summarised on the box: the .ncu-rep of 21 launches exceeds the 64 MiB pull limit) and a
# source-level capture of the level-0 window kernel.  usage: tools/profile_round.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-rXX}; o=gpurun_out; B=32  # bench.py's default batch
python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err; echo "reference rc=$?"
python bench.py --batch 16 --no-cpu --e2e-pairs 4000 > $o/${tag}_bench_b16.json 2>> $o/${tag}_bench.err; echo "b16 rc=$?"
python tools/parity_report.py > $o/${tag}_parity_fullsize.jsonl 2> $o/${tag}_parity.err; echo "parity rc=$?"
export TW_GRAPH=0
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > $o/${tag}_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_batch${B}.csv $CMD > $o/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?"
$CMD > $o/${tag}_plain2.log 2>&1 && ncu --set full --clock-control none -k regex:"gauss_strip|gauss_iter2|gauss_last_sparse|polyexp|first_update|level_" -s 63 -c 21 -o /tmp/prof_${tag} $CMD > $o/${tag}_ncu_full.log 2>&1; echo "full rc=$?"
python tools/ncu_summary.py /tmp/prof_${tag}.ncu-rep $o/${tag}_ncu_full_step_batch${B}.csv
python tools/ncu_traffic.py $o/${tag}_ncu_full_step_batch${B}.csv $B relaxed > $o/${tag}_traffic.json
# source-level capture of a level-0 launch of the window kernel (tile kernel: 11 launches per step incl. the dense lasts of the coarser scales; 9, 10 = finest)
$CMD > $o/${tag}_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gauss_iter2" -s 42 -c 1 -f -o $o/${tag}_window_level0 $CMD > $o/${tag}_ncu_window.log 2>&1; echo "window rc=$?"
# the same launch of the opt-in strip kernel (TW_WINDOW=strip: 11 launches per step, indices 9 and 10 are the finest scale)
TW_WINDOW=strip ncu --set full --clock-control none --import-source on -k regex:"gauss_strip" -s 20 -c 1 -f -o $o/${tag}_strip_level0 $CMD > $o/${tag}_ncu_strip.log 2>&1; echo "strip rc=$?"
TW_WINDOW=strip python bench.py --no-cpu --no-e2e --steps 100 > $o/${tag}_bench_strip.json 2>> $o/${tag}_bench.err
python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python tools/cfg_bench.py cfg3 > $o/${tag}_cfg3.json 2>> $o/${tag}_bench.err; python tools/cfg_bench.py cfg4 > $o/${tag}_cfg4.json 2>> $o/${tag}_bench.err
ls -la $o | grep ${tag}
