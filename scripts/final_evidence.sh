#!/bin/bash
# Round evidence in one gpurun call (one GPU): GPU tests, the bench line, the reference arm, the launch list, one
# ncu --set full capture of the two heavy kernels, the other configs and the streaming shim.  Everything lands in
# gpurun_out/<tag>_*; ncu runs only after the same command has exited 0 without it.
#   gpurun --timeout 1500 -- 'bash scripts/final_evidence.sh r2z'
tag=${1:-ev}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/${tag}_tests.log 2>&1; tail -2 $o/${tag}_tests.log
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference > $o/${tag}_bench_ref.json 2>> $o/${tag}_bench.err; echo "ref rc=$?"
small="bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs"
if python $small > $o/${tag}_small.json 2>> $o/${tag}_bench.err; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv python $small > $o/${tag}_ncu.log 2>&1
  ncu --set full --import-source on --clock-control none -k regex:'kzg_exec_kernel|kzg_fse_kernel' -c 2 -o $o/${tag}_full python $small > $o/${tag}_ncu2.log 2>&1
fi
python scripts/gpu_configs.py 1.0 logs json literal random mixed mixed:64 > $o/${tag}_cfg.log 2>&1
python scripts/gpu_stream_bench.py --native > $o/${tag}_stream.txt 2>&1
tail -1 $o/${tag}_bench.json | cut -c1-400
