mkdir -p gpurun_out/r1j
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1j/bench_n2.json 2> gpurun_out/r1j/bench_n2.err; echo "n2 exit $?"
tail -3 gpurun_out/r1j/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r1j/ref_n2.json 2> gpurun_out/r1j/ref_n2.err; echo "ref n2 exit $?"
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r1j/bench_n1.json 2> gpurun_out/r1j/bench_n1.err; echo "n1 exit $?"
python - <<'PY'
import json
for f in ("bench_n1","bench_n2","ref_n2"):
    try:
        d=json.loads(open(f"gpurun_out/r1j/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), round(d["value"]), round(d["e2e"]["value"]), d.get("pcie"), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
