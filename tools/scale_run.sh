#!/bin/bash
# N-GPU runs through torchrun, as the driver launches them: the bench line with every BASELINE config (inference B=1024 L=1 +
# the `configs` block: L=5, hi-res, B=256, training B=32 / B=128 per GPU), and the reference arm.  usage: scale_run.sh N [tag]
N=${1:-8}
R=${2:-r02}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 exit $?"; grep -c '^{' gpurun_out/$2.json; }
run 29521 bench_n${N}_${R} --steps 20 --warmup 3
run 29522 bench_train_n${N}_${R} --mode train --steps 20 --warmup 5 --no-configs
run 29523 bench_ref_n${N}_${R} --impl reference --steps 2 --warmup 1
nvidia-smi topo -m > gpurun_out/topo_n${N}_${R}.txt 2>&1
(numactl -H || lscpu | grep -i numa) > gpurun_out/numa_n${N}_${R}.txt 2>&1
