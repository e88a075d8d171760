#!/bin/bash
# Short GPU call: tests, smoke, bench (both arms) and the ncu launch list of the bench command.
#   gpurun --timeout 1500 -- 'bash scripts/capture_lite.sh r02'
TAG=${1:-r02}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -8 ) 2>&1 | tee gpurun_out/${TAG}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${TAG}_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
tail -2 gpurun_out/${TAG}_bench_1gpu.err
cat gpurun_out/${TAG}_bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> /dev/null
cat gpurun_out/${TAG}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-extra-configs --no-fast-mode --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
