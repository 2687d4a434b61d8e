#!/bin/bash
# 2-GPU probe of the opt-in NCCL-cycle-under-CUDA-graph path (LULESH_B200_NCCL_GRAPH=1): what fails, and
# whether NCCL's graph buffer registration is involved.  Every run under its own timeout.
out=gpurun_out/${1:-r02_ncclgraph}; mkdir -p $out
B=./lulesh_b200/bin/lulesh_b200
export LULESH_B200_HALO=nccl NCCL_DEBUG=WARN
echo "== eager (baseline)"; timeout 120 $B --gpus 2 --global 20 -q > $out/eager.log 2>&1; echo "rc $?"; tail -n 3 $out/eager.log
echo "== graph, defaults"; LULESH_B200_NCCL_GRAPH=1 timeout 120 $B --gpus 2 --global 20 > $out/graph_default.log 2>&1; echo "rc $?"; tail -n 12 $out/graph_default.log
echo "== graph, NCCL_GRAPH_REGISTER=0"; NCCL_GRAPH_REGISTER=0 LULESH_B200_NCCL_GRAPH=1 timeout 120 $B --gpus 2 --global 20 > $out/graph_noreg.log 2>&1; echo "rc $?"; tail -n 12 $out/graph_noreg.log
echo "== graph, NCCL_GRAPH_REGISTER=0 NCCL_GRAPH_MIXING_SUPPORT=0"; NCCL_GRAPH_MIXING_SUPPORT=0 NCCL_GRAPH_REGISTER=0 LULESH_B200_NCCL_GRAPH=1 timeout 120 $B --gpus 2 --global 20 > $out/graph_nomix.log 2>&1; echo "rc $?"; tail -n 12 $out/graph_nomix.log
echo "== graph, separate processes (torchrun bench, -s 64)"
LULESH_B200_NCCL_GRAPH=1 timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --size 64 --steps 50 --warmup 5 --no-extras --no-parity > $out/graph_torchrun.json 2> $out/graph_torchrun.err; echo "rc $?"
tail -n 6 $out/graph_torchrun.err; cut -c1-300 $out/graph_torchrun.json
echo "== graph + NCCL_GRAPH_REGISTER=0, separate processes"
NCCL_GRAPH_REGISTER=0 LULESH_B200_NCCL_GRAPH=1 timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
   bench.py --gpus 2 --size 64 --steps 50 --warmup 5 --no-extras --no-parity > $out/graph_torchrun_noreg.json 2> $out/graph_torchrun_noreg.err; echo "rc $?"
tail -n 6 $out/graph_torchrun_noreg.err; cut -c1-300 $out/graph_torchrun_noreg.json
