#!/bin/bash
# N-GPU runs (inference sweep point + data-parallel training step) through torchrun; usage: scale_run.sh N
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 exit $?"; grep -c '^{' gpurun_out/$2.json; }
run 29521 bench_n${N} --steps 20 --warmup 3 --no-cpu-baseline
run 29522 bench_train_n${N} --mode train --steps 10 --warmup 3
