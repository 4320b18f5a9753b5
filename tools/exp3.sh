mkdir -p gpurun_out/r1k
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r1k/bench_n2.json 2> gpurun_out/r1k/bench_n2.err; echo "n2 exit $?"
tail -3 gpurun_out/r1k/bench_n2.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r1k/bench_n1.json 2> gpurun_out/r1k/bench_n1.err; echo "n1 exit $?"
python - <<'PY'
import json
for f in ("bench_n1","bench_n2"):
    try:
        d=json.loads(open(f"gpurun_out/r1k/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), {k: round(v["avg_us"],1) for k,v in d["kernels"].items()})
    except Exception as e: print(f, "ERR", e)
PY
