#!/bin/bash
# One GPU-box call that produces everything profiles/ is made from (1 GPU):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile_run.sh r1'
# then here: python scripts/make_profiles.py r1
# Numbers printed under ncu are never bench values; the bench JSON lines come
# from the un-profiled runs at the end.
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
rm -f $out/*_${tag}.ncu-rep $out/launches_${tag}.csv $out/bench_${tag}_*.json

# 1. launch list of the bench command, captured deep into the timed region
#    (9 launches per training step; ncu slows every skipped launch, so only 70 steps are skipped)
RECUR_B200_PLAIN_LAUNCH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 630 -c 90 --csv \
  --log-file $out/launches_${tag}.csv \
  python bench.py --steps 100 --warmup 30 --no-cpu-baseline --trained-after 0 > $out/ncu_launches_${tag}.log 2>&1

# 2. one --set full capture per hot kernel, 120 training steps in (ncu slows every
#    intercepted launch, deeper captures cost minutes of box time each)
# (the persistent chain kernel is launched cooperatively on clusters, which ncu's
#  kernel replay cannot do: RECUR_B200_PLAIN_LAUNCH drops the cooperative attribute for the
#  captures, nothing else changes)
export RECUR_B200_PLAIN_LAUNCH=1
for k in k_tc_chain_persistent k_tc_dw_pair k_tc_nt k_update_split k_out_multi; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 120 -c 1 \
    -f -o $out/${k}_${tag} python bench.py --steps 140 --warmup 30 --no-cpu-baseline --trained-after 0 \
    > $out/ncu_${k}_${tag}.log 2>&1
done

unset RECUR_B200_PLAIN_LAUNCH

# 2b. the cell automaton's frame kernel (BASELINE configs[4])
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cells_frame_tc -s 8 -c 1 \
  -f -o $out/k_cells_frame_tc_${tag} python bench.py --config rnnca --steps 20 --warmup 5 --no-cpu-baseline \
  > $out/ncu_k_cells_frame_tc_${tag}.log 2>&1

# 3. the bench lines themselves, un-profiled
timeout 900 python bench.py > $out/bench_${tag}_n1.json 2> $out/bench_${tag}_n1.err
timeout 900 python bench.py --impl reference > $out/bench_${tag}_ref.json 2> $out/bench_${tag}_ref.err
timeout 600 python bench.py --config rnnca > $out/bench_${tag}_rnnca.json 2> $out/bench_${tag}_rnnca.err
timeout 600 python bench.py --config rnnca --impl reference > $out/bench_${tag}_rnnca_ref.json 2> $out/bench_${tag}_rnnca_ref.err
tail -c 600 $out/bench_${tag}_n1.json
tail -c 400 $out/bench_${tag}_ref.json
