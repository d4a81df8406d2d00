#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench line, ncu launch list, ncu --set full of the traversal kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag] [what...]'   what ⊆ {tests smoke bench launches ncu}
set -u
TAG=${1:-r01}; shift || true
WHAT=${*:-tests smoke bench launches ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
for w in $WHAT; do
  case $w in
    tests) timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log;;
    bench) timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err;;
    benchref) timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_bench.log 2>&1; echo "launches rc=$?";;
    ncu) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s ${NCU_SKIP:-10} -c ${NCU_COUNT:-5} -f -o $OUT/prof_trace \
        env MODE=northstar FRAMES=2 python tools/quick_bench.py ${SCENES:-c2} > $OUT/ncu_trace.log 2>&1; echo "ncu rc=$?";;
    bench2) for sh in samples rows; do timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus ${NGPU:-2} --steps 20 --warmup 3 --shard $sh > $OUT/bench_${NGPU:-2}gpu_$sh.json 2> $OUT/bench_${NGPU:-2}gpu_$sh.err; echo "bench2 $sh rc=$?"; cat $OUT/bench_${NGPU:-2}gpu_$sh.json; tail -3 $OUT/bench_${NGPU:-2}gpu_$sh.err; done;;
    quick) timeout 600 python tools/quick_bench.py ${SCENES:-c2 c2flat} > $OUT/quick.log 2>&1; echo "quick rc=$?"; cat $OUT/quick.log;;
    quickstatic) SNKPT_TRACE_STATIC=1 timeout 600 python tools/quick_bench.py ${SCENES:-c2} > $OUT/quick_static.log 2>&1; cat $OUT/quick_static.log;;
    ncurestir) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_restir -s 4 -c 2 -f -o $OUT/prof_restir \
        python tools/quick_bench.py ${SCENES:-c2} > $OUT/ncu_restir.log 2>&1; echo "ncu rc=$?";;
  esac
done
ls -la $OUT
