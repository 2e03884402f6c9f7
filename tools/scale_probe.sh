#!/bin/bash
# one multi-GPU probe: topology, concurrent pinned H2D bandwidth, the 2-GPU NCCL test, bench at N GPUs.
# usage: tools/scale_probe.sh N [extra bench args]
N=${1:-8}
shift
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name" >> $O/topo_n$N.txt
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node)" >> $O/topo_n$N.txt; fi; done
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_bench.py --concurrent > $O/h2d_n$N.jsonl 2> $O/h2d_n$N.err
timeout 200 python -m pytest tests/test_gpu_infer_dist.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 > $O/dist_test_n$N.log
NCCL_DEBUG=INFO timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/bench_n${N}.json 2> $O/bench_n${N}.err
grep -c "NCCL INFO" $O/bench_n${N}.err; tail -c 600 $O/bench_n${N}.json; echo; cat $O/dist_test_n$N.log; cat $O/h2d_n$N.jsonl | tail -12
