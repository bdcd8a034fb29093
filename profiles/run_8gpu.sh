#!/bin/bash
# everything that needs the full 8-GPU node, in one lease: topology, the NCCL / in-process replication parity tests, the box's
# host<->device ceiling, the scaling bench (with extras and all-rank parity) and the end-to-end variants
out=gpurun_out/r2_8gpu; mkdir -p $out
N=$(nvidia-smi -L | wc -l)
(nvidia-smi topo -m; echo "nproc $(nproc)"; lscpu | grep -i "numa\|model name\|socket"; free -g | head -2) > $out/topology.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharding_nccl.py tests/test_gpu_replicate.py -m gpu -x -q -rs -v > $out/pytest_${N}gpu.log 2>&1; tail -3 $out/pytest_${N}gpu.log
for k in 1 2 4 8; do [ $k -le $N ] && python profiles/exp_pcie_multi.py --single $k; done > $out/pcie_ceiling.txt 2>&1
for k in 2 4 8; do [ $k -le $N ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2951$k profiles/exp_pcie_multi.py 2>/dev/null | grep mode; done >> $out/pcie_ceiling.txt
[ 8 -le $N ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 profiles/exp_pcie_multi.py --split-affinity 2>/dev/null | grep mode >> $out/pcie_ceiling.txt
cat $out/pcie_ceiling.txt
for k in 8 4; do
  [ $k -le $N ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2953$k bench.py --gpus $k 2> $out/bench_n$k.err | grep '^{' > $out/bench_n$k.json; echo "bench n=$k rc=$?"
done
if [ 8 -le $N ]; then
  for aff in none nvml; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-extras --affinity $aff 2>/dev/null | grep '^{' > $out/bench_n8_aff_$aff.json
  done
  BRG_NARROW_OUT=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --no-extras 2>/dev/null | grep '^{' > $out/bench_n8_narrow.json
  BRG_HOST_THREADS=3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --no-extras 2>/dev/null | grep '^{' > $out/bench_n8_threads3.json
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.3g" % d["value"], "e2e %.3g (%.1f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "pageable %.3g" % d["e2e_pageable"]["value"],
              "bcast", d["build"].get("broadcast_ms"), "parity", d["parity"]["all_ranks_all_rows"] if d.get("parity") else None)
    except Exception as ex:
        print(f, "unreadable", ex)
PY
