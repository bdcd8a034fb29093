out=gpurun_out/r2_4gpu; mkdir -p $out
N=$(nvidia-smi -L | wc -l); echo "gpus $N"
timeout 600 python -m pytest tests/test_gpu_sharding_nccl.py tests/test_gpu_replicate.py -m gpu -x -q -rs > $out/pytest.log 2>&1; tail -2 $out/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --no-cpu --no-e2e 2> $out/bench_n4.err | grep '^{' > $out/bench_n4_spread.json; echo "rc $?"; tail -c 300 $out/bench_n4.err
BRG_REPLICATE=broadcast timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --no-cpu --no-e2e --no-extras 2>/dev/null | grep '^{' > $out/bench_n4_broadcast.json
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.3g" % d["value"], "frac %.3f" % d["roofline"]["frac"], "bcast", d["build"].get("broadcast_ms"), d["build"].get("broadcast_gbs"), "bytes", d["build"]["index_bytes"], "parity", d.get("parity"))
        for k, v in d.get("extras", {}).items():
            if isinstance(v, dict) and "ms_per_step" in v: print("   ", k, round(v["ms_per_step"], 4), "value %.3g" % v["value"], v.get("parity_all"), (v.get("build") or {}).get("broadcast_ms"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
