mkdir -p gpurun_out/r1v
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r1v/bench_n8.json 2> gpurun_out/r1v/bench_n8.err; echo "n8 exit $?"
tail -2 gpurun_out/r1v/bench_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r1v/bench_n8.json").read().strip().splitlines()[-1])
    print("bench_n8", d.get("n_gpus"), round(d["value"]), d.get("ms_per_step"), round(d["e2e"]["value"]), d["clocks"])
except Exception as e: print("ERR", e)
PY
