#!/bin/bash
# The measurement pass behind profiles/r2_*: GPU tests, smoke, the three bench lines + the reference arm, the ncu
# launch list of bench.py, one --set full capture of bidir_ctc, and the side benches.  One B200, under gpurun.
mkdir -p gpurun_out
o=gpurun_out/r2z
python -m pytest tests -m gpu -q > ${o}_tests.log 2>&1; tail -2 ${o}_tests.log
python __graft_entry__.py smoke > ${o}_smoke.log 2>&1; tail -1 ${o}_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > ${o}_ref.json 2> ${o}_ref.err
python bench.py --steps 20 --warmup 5 > ${o}_bench.json 2> ${o}_bench.err
python bench.py --workload asg --steps 10 --warmup 3 > ${o}_asg.json 2> ${o}_asg.err
python bench.py --workload viterbi --steps 10 --warmup 3 > ${o}_vit.json 2> ${o}_vit.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${o}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-api-path > ${o}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bidir -s 1 -c 1 -o ${o}_bidir -f \
  python scripts/one_step_bidir.py 3 > ${o}_ncu_full.log 2>&1; tail -1 ${o}_ncu_full.log
python scripts/bench_variants.py > ${o}_variants.json 2> ${o}_variants.err
python scripts/bench_gcompose.py > ${o}_gcompose.json 2> ${o}_gcompose.err; tail -4 ${o}_gcompose.err
python scripts/bench_api.py > ${o}_api.log 2>&1; tail -2 ${o}_api.log
python scripts/bench_configs.py > ${o}_configs.json 2> ${o}_configs.err
head -c 600 ${o}_bench.json
