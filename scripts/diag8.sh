#!/bin/bash
# dev helper: 8-GPU diagnostics of the sharded sweep (timelines incl. commits, commit paths, e2e phases)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1"
B="bench.py --gpus ${NG:-8} --steps 20 --warmup 5 --no-parity --no-weak --no-e2e"
BENCH_DUMP_TIMELINE=1 $T --master-port 29701 $B > gpurun_out/d8_lag2.json 2> gpurun_out/d8.err
XCOLUMNS_B200_LAG=1 BENCH_DUMP_TIMELINE=1 $T --master-port 29702 $B > gpurun_out/d8_lag1.json 2>> gpurun_out/d8.err
XCOLUMNS_B200_P2P=0 $T --master-port 29703 $B > gpurun_out/d8_nccl.json 2>> gpurun_out/d8.err
BENCH_E2E_PHASES=1 $T --master-port 29704 bench.py --gpus ${NG:-8} --steps 20 --warmup 5 --no-parity --no-weak > gpurun_out/d8_e2e.json 2>> gpurun_out/d8.err
$T --master-port 29705 scripts/multi_gpu_check.py > gpurun_out/d8_multi.log 2>&1
