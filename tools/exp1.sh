mkdir -p gpurun_out/r1i
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r1i/bench.json 2> gpurun_out/r1i/bench.err; echo "bench exit $?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-shard 32 > gpurun_out/r1i/bench_s32.json 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-shard 128 > gpurun_out/r1i/bench_s128.json 2>&1
python - <<'PY'
import json
for f in ("bench","bench_s32","bench_s128"):
    d=json.loads(open(f"gpurun_out/r1i/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), {k: round(v["avg_us"],1) for k,v in d["kernels"].items()})
PY
