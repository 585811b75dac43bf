#!/bin/bash
# Round-2 multi-GPU evidence on one 8-GPU box: driver-format lines for c4 (weak, the named configuration) and c5,
# the in-process multi-GPU arm (the library's own ngpus argument), and the multi-GPU tests.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "two_gpus or sharded or ngpus" 2>&1 | tail -3 > gpurun_out/r2_multi_tests.log
$TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_c4_${N}gpu.json 2> gpurun_out/r2_bench_c4_${N}gpu.err
$TR bench.py --gpus $N --steps 2 --warmup 1 --workload c5 > gpurun_out/r2_bench_c5_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err
python bench.py --inproc --gpus $N --points-per-gpu 625000 --steps 2 --warmup 1 > gpurun_out/r2_bench_inproc_${N}gpu.json 2> gpurun_out/r2_bench_inproc_${N}gpu.err
python bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_${N}gpu.json 2>/dev/null
cat gpurun_out/r2_multi_tests.log
for f in gpurun_out/r2_bench_*_${N}gpu.json; do echo "== $f"; head -c 400 $f; echo; done
tail -n 5 gpurun_out/r2_bench_*_${N}gpu.err
