mkdir -p gpurun_out/r1u
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r1u/bench_n4.json 2> gpurun_out/r1u/bench_n4.err; echo "n4 exit $?"
tail -2 gpurun_out/r1u/bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r1u/ref_n4.json 2> gpurun_out/r1u/ref_n4.err; echo "ref n4 exit $?"
python - <<'PY'
import json
for f in ("bench_n4","ref_n4"):
    try:
        d=json.loads(open(f"gpurun_out/r1u/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), round(d["value"]), d.get("ms_per_step"), round(d["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
