#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: canonicalized img/s at 3x224x224, C8 (configs[1]).

    python bench.py --gpus N --steps K --warmup W                  # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...         # the UNMODIFIED reference on the host cores
    python bench.py --impl reference-gpu ...                        # (informational) the same reference classes, eager, on cuda:0

One "step" = one pass of the hot path over one batch of 512 synthetic images per GPU:
    y = canonicalizer(x)                        crop+antialiased resize -> fused C8 group-conv stack ->
                                                group pool/select (+ prior statistic) -> inverse-rotation warp
    z = canonicalizer.invert_canonicalization(y, induced_rep_type="scalar")     forward warp of the output
    prior loss + identity metric                (ONE 3-float NCCL all-reduce when N > 1)
The prediction network between canonicalize and invert (ResNet-50 in configs[1]) is the caller's
PyTorch module and is NOT part of the hot path, so it is not run or timed here.

Timed region (`value`): K replays of the step captured ONCE as a CUDA graph through the public
`canonicalizer.capture_step()` (kernels + the all-reduce; one launch per step, so host jitter of any rank cannot stall
the others), one CUDA event per step -> min / median / max step time; `--no-graph` times the eager public calls instead
(also reported as `eager`).  Per-kernel times (`kernels`, `roofline`) come from CUDA events around every C-ABI call in an
eager pass of the same K steps right after the timed region (events cannot sit between the nodes of a graph launch).
`e2e` = the same step through the public classes from/to pinned HOST buffers (H2D + D2H inside the timed region);
`cpu_baseline` / `--impl reference` = the reference's own classes (baseline/_ref, installed by baseline/install_ref.sh;
the oracle port when that directory is absent) on this box's host cores on a bounded sample; `configs` = the other
BASELINE.json configurations (optimised D4 orbit, SO(3) point clouds, E(3) n-body) measured at this N.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import statistics
import subprocess
import sys
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "canonicalized img/s at 3x224x224 C8"
N_ROT, IN_SHAPE, CROP, RESIZE = 8, (3, 224, 224), 0.8, 96
OUT_CH, KSIZE, LAYERS = 32, 5, 3
IMG_BYTES = 2 * 3 * 224 * 224 * 4                      # warp: 1 read + 1 write (SURVEY.md 8d "W")
STACK_FLOP_EXECUTED = 2 * 92 * 92 * 256 * (75 + 256)   # lift + one 1x1 layer; the last layer is folded
STACK_FLOP_REFERENCE = 2 * 92 * 92 * 256 * (75 + 256 + 256)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(ROOT, "oracle", "shims")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (batch 512) of the kernels behind each C-ABI call, from the
    NEWEST committed ncu export under profiles/ (profiles/*_traffic.json, written by tools/ncu_traffic.py from an
    `ncu --set full` capture of this very command); {} when there is none."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return {}, None
    try:
        return json.load(open(files[-1])), os.path.basename(files[-1])
    except (OSError, ValueError):
        return {}, None


def make_layers(seed=0):
    """Weights exactly as the reference initialises them (kaiming-uniform a=sqrt(5), bias 0), seed 0, on CPU."""
    from equiadapt_b200.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork
    torch.manual_seed(seed)
    return CustomEquivariantNetwork((3, RESIZE, RESIZE), OUT_CH, KSIZE, "rotation", N_ROT, LAYERS, device="cpu")


def host_batch(batch, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, *IN_SHAPE, generator=g)


def hp_image():
    return SimpleNamespace(beta=1.0, input_crop_ratio=CROP, resize_shape=RESIZE)


# ---------------------------------------------------------------------------------------------------
# the reference itself (baseline/_ref) and the oracle port of it
# ---------------------------------------------------------------------------------------------------
def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "equiadapt"))


def import_reference():
    """The UNMODIFIED reference package from baseline/_ref behind the stand-in modules of oracle/shims (kornia, e2cnn,
    torch_scatter, omegaconf are not installable here; only the kornia stand-in carries arithmetic)."""
    if SHIMS not in sys.path:
        sys.path[:0] = [SHIMS, REF_DIR]
    import equiadapt  # noqa: F401
    return equiadapt


def reference_image_canonicalizer(device="cpu"):
    """The reference's GroupEquivariantImageCanonicalization + CustomEquivariantNetwork for configs[1]; same seed, hence
    the same weights as make_layers().  `group_type` / `num_rotations` are set on the network by hand: the reference
    class forgets them and its wrapper reads them -- driver glue, not a modification."""
    import_reference()
    from equiadapt.images.canonicalization.discrete_group import GroupEquivariantImageCanonicalization
    from equiadapt.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork
    torch.manual_seed(0)
    net = CustomEquivariantNetwork((3, RESIZE, RESIZE), OUT_CH, KSIZE, "rotation", N_ROT, LAYERS, device="cpu")
    if str(device) != "cpu":
        # the reference keeps its group-permutation index tables as plain attributes created on the constructor's
        # `device` (not buffers: .to() leaves them behind), so a GPU instance is built there and given the CPU-seeded weights
        sd = net.state_dict()
        net = CustomEquivariantNetwork((3, RESIZE, RESIZE), OUT_CH, KSIZE, "rotation", N_ROT, LAYERS, device=str(device))
        net.load_state_dict(sd)
    net.group_type, net.num_rotations = "rotation", N_ROT
    return GroupEquivariantImageCanonicalization(net, hp_image(), IN_SHAPE).to(device).eval()


def public_step(can, x):
    y = can(x)
    z = can.invert_canonicalization(y, induced_rep_type="scalar")
    return z, can.get_prior_regularization_loss(), can.get_identity_metric()


def oracle_step(O, x, layers):
    x_pre = O.pre_network_transform(x, IN_SHAPE, CROP, RESIZE)
    act = O.custom_equivariant_network(x_pre, layers, N_ROT, False)
    el = O.activations_to_group_element(act, N_ROT, False)
    y = O.canonicalize_image(x, el["rotation"], None)
    z = O.invert_image_features(y, el["rotation"], None, N_ROT, N_ROT, "scalar")
    return z, O.prior_loss_discrete(act), O.identity_metric_discrete(act)


def time_cpu_reference(sample_batch, steps, warmup, budget_s=None):
    """The reference path on the host cores, all threads: its own classes when baseline/_ref is there (kind
    "reference"), else the oracle port (kind "port")."""
    import warnings
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    x = host_batch(sample_batch, seed=1)
    if reference_available():
        can = reference_image_canonicalizer("cpu")
        kind, run = "reference", (lambda: public_step(can, x))
    else:
        from oracle import reference_path as O
        net = make_layers()
        layers = [(m.weights.detach(), m.bias.detach()) for m in net.eqv_network if hasattr(m, "weights")]
        kind, run = "port", (lambda: oracle_step(O, x, layers))
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            run()
        t_all = time.perf_counter()
        for _ in range(steps):
            t0 = time.perf_counter()
            run()
            times.append(time.perf_counter() - t0)
            if budget_s is not None and time.perf_counter() - t_all > budget_s and len(times) >= 3:
                break
    total = sum(times)
    return {"img_s": sample_batch * len(times) / total, "ms_per_step": 1e3 * total / len(times), "steps": len(times),
            "cores": torch.get_num_threads(), "sample_batch": sample_batch, "kind": kind}


def cpu_sample_text(r):
    what = ("the unmodified reference classes from baseline/_ref (kornia restated by oracle/shims)" if r["kind"] == "reference"
            else "the oracle port of the reference path (baseline/_ref absent)")
    return (f"{r['steps']} steps x {r['sample_batch']} images of the same workload (the reference's B=512 transients exceed "
            f"40 GB on CPU, SURVEY.md 8d), eval/no_grad, {r['cores']} torch threads: {what}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_cpu_reference(args.cpu_sample_batch, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["img_s"], "unit": "img/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus),
        "cpu_baseline": {"value": r["img_s"], "unit": "img/s", "cores": r["cores"], "kind": r["kind"], "sample": cpu_sample_text(r)},
        "e2e": {"value": r["img_s"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_reference_gpu(args):
    """Informational second baseline: the reference's own classes, unchanged, eager PyTorch on cuda:0 -- what a
    user gets today by calling .cuda() on equiadapt.  The batch is chunked (its transients are ~40 GB per 512 images)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not reference_available() or not torch.cuda.is_available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "needs baseline/_ref (baseline/install_ref.sh) and a CUDA device"}))
        return
    dev = torch.device("cuda", 0)
    can = reference_image_canonicalizer(dev)
    B, chunk = args.batch, args.ref_gpu_chunk
    x = host_batch(B, seed=1).to(dev)

    def step():
        out = []
        for lo in range(0, B, chunk):
            out.append(public_step(can, x[lo:lo + chunk]))
        return out

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    line = {"impl": "reference-gpu", "metric": METRIC, "value": B / (ms * 1e-3), "unit": "img/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(B, 1),
            "how": f"unmodified equiadapt classes from baseline/_ref on cuda:0, eager, {chunk}-image chunks, inputs resident in HBM, "
                   "CUDA events; kornia restated by oracle/shims (torch ops: affine_grid / grid_sample)",
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    print(json.dumps(line))


def workload_config(batch, n_gpus):
    return {
        "workload": ("BASELINE configs[1]: C8 GroupEquivariantImageCanonicalization canonicalize + invert(scalar) warp, "
                     "synthetic U[0,1) 3x224x224, batch 512 per GPU; CustomEquivariantNetwork out_channels 32, "
                     "kernel 5, 3 layers, input_crop_ratio 0.8, resize 96; prediction network excluded"),
        "per_gpu_batch": batch, "global_batch": batch * n_gpus, "parallelism": f"dp{n_gpus} (batch-sharded)",
        "l2": "inputs 308 MB per GPU > 126 MB L2 (no flush needed)",
        "collective": "none" if n_gpus == 1 else "one 3-float NCCL all-reduce per step (prior statistic), forked beside the warp kernels inside the captured graph",
    }


# ---------------------------------------------------------------------------------------------------
# B200 arm helpers
# ---------------------------------------------------------------------------------------------------
def _cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out += list(range(int(lo), int(hi or lo) + 1))
    return out


def place_rank(local, world):
    """Bind this rank's threads (and therefore its first-touch pinned buffers) to the CPUs next to its GPU: the GPU's
    NUMA node when sysfs names one, and within it a private slice per local rank so eight ranks do not migrate over each
    other.  Returns what was done (reported in the JSON line)."""
    info = {"numa_node": None, "cpus": None}
    try:
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read())
        local_cpus = _cpulist(open(base + "/local_cpulist").read())
        info["numa_node"] = node
    except Exception:
        local_cpus = []
    try:
        allowed = sorted(os.sched_getaffinity(0))
        pool = [c for c in allowed if c in set(local_cpus)] or allowed
        if world > 1:
            # ranks that share a pool (all GPUs report the same node on these boxes) take disjoint slices of it
            per = max(len(pool) // world, 1)
            mine = pool[(local * per) % len(pool):(local * per) % len(pool) + per] or pool
        else:
            mine = pool
        os.sched_setaffinity(0, mine)
        info["cpus"] = f"{mine[0]}-{mine[-1]} ({len(mine)})"
    except Exception as e:  # noqa: BLE001
        info["cpus"] = f"unchanged ({type(e).__name__})"
    return info


class ClockSampler:
    """SM clock / throttle reasons of ONE GPU, sampled every ~10 ms by an NVML thread while the timed region runs.  Only rank
    0 samples (NVML calls take driver locks that every rank's launches also need; round 1 polled from every rank every
    5 ms); nvidia-smi is the fallback when NVML is not importable."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        import threading
        self.proc, self.thread, self.samples, self.stop_flag = None, None, [], False
        self.interval = float(os.environ.get("EQB_CLOCK_SAMPLE_MS", "10")) / 1e3
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = torch.cuda.get_device_properties(index).uuid
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
                h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml, self.h = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._sample()
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def _sample(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
            mask = int(reasons_fn(self.h))
            try:
                watts = n.nvmlDeviceGetPowerUsage(self.h) / 1e3
            except Exception:
                watts = None
            self.samples.append((mhz, mask, watts, time.perf_counter()))
        except Exception:
            pass

    def _run(self):
        while not self.stop_flag:
            time.sleep(self.interval)
            self._sample()

    def stop(self, t_begin=None, t_end=None):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            inside = [s for s in self.samples if t_begin is None or t_begin <= s[3] <= t_end] or self.samples
            sm = sorted(s[0] for s in inside)
            reasons = sorted(k for k, b in bits.items() if any(s[1] & b for s in inside))
            power = [s[2] for s in inside if s[2] is not None]
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                    "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi"}


def timed_steps(fn, steps, barrier):
    """K calls of fn with one CUDA event per step on the launching stream; -> (total ms, [per-step ms]).  The host does not
    synchronise inside the region."""
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    t_begin = time.perf_counter()
    evs[0].record()
    for i in range(steps):
        fn()
        evs[i + 1].record()
    barrier()
    t_end = time.perf_counter()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return evs[0].elapsed_time(evs[steps]), per, (t_begin, t_end)


def step_stats(per):
    return {"min": min(per), "median": statistics.median(per), "max": max(per)}


# ---------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (configs[2..4]) at this N
# ---------------------------------------------------------------------------------------------------
def conv_network(in_shape, out_channels, kernel_size, num_layers, out_vector_size):
    """The caller-side torch network the optimisation-based variant scores its orbit with (same architecture as the
    reference's example ConvNetwork: stride-2 k x k Conv2d + BatchNorm2d + GELU per layer, channels doubled on every
    third, then BatchNorm1d / Dropout1d / ReLU / Linear): cuDNN, NOT part of the hot path -- timed with and without."""
    nn = torch.nn
    layers, c = [], in_shape[0]
    for i in range(num_layers):
        if i == 0:
            layers.append(nn.Conv2d(c, out_channels, kernel_size, 2)); c = out_channels
        elif i % 3 == 2:
            layers.append(nn.Conv2d(c, 2 * c, kernel_size, 2, 1)); c = 2 * c
        else:
            layers.append(nn.Conv2d(c, c, kernel_size, 2))
        layers += [nn.BatchNorm2d(c), nn.GELU()]
    enc = nn.Sequential(*layers)
    out_dim = enc(torch.zeros(1, *in_shape)).flatten(1).shape[1]

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.enc_network = enc
            self.final_fc = nn.Sequential(nn.BatchNorm1d(out_dim), nn.Dropout1d(0.5), nn.ReLU(), nn.Linear(out_dim, out_vector_size))
            self.out_vector_size = out_vector_size

        def forward(self, x):
            return self.final_fc(self.enc_network(x).reshape(x.shape[0], -1))

    return Net()


def measure(fn, inputs, steps, warmup, barrier, world, dev, use_graph=True):
    """-> (ms per step max over ranks, mode): fn(*inputs) timed as a captured graph when it captures, eagerly otherwise."""
    import torch.distributed as dist
    from equiadapt_b200 import graphed
    mode, g = "eager", None
    call = lambda: fn(*inputs)  # noqa: E731
    with torch.no_grad():
        if use_graph:
            try:
                g = graphed.capture(fn, inputs, warmup=max(warmup, 3))
                call, mode = (lambda: g()), "graph"
            except Exception as e:  # noqa: BLE001
                torch.cuda.synchronize()
                mode = f"eager (capture failed: {type(e).__name__})"
        for _ in range(max(warmup, 3)):
            call()
        total, per, _ = timed_steps(call, steps, barrier)
        call = g = None
    t = torch.tensor([total / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), mode


def bench_other_configs(args, dev, rank, world, barrier, pk):
    """configs[2] optimised D4 orbit (B = 256 global), configs[3] SO(3) point clouds (128 x 1024), configs[4] E(3) n-body
    (10 000 systems): each STRONG (the configuration's global batch sharded over the N ranks, as BASELINE.json states it) and
    WEAK (that batch per GPU), prior statistic all-reduced (one 3-float NCCL call per step), samples/s = all ranks'
    samples / max-over-ranks device time."""
    from equiadapt_b200 import distributed as D
    from equiadapt_b200.images.canonicalization.discrete_group import OptimizedGroupEquivariantImageCanonicalization
    from equiadapt_b200.nbody.canonicalization.euclidean_group import EuclideanGroupNBody
    from equiadapt_b200.nbody.canonicalization_networks.custom_equivariant_networks import VNDeepSets
    from equiadapt_b200.pointcloud.canonicalization.continuous_group import EquivariantPointcloudCanonicalization
    from equiadapt_b200.pointcloud.canonicalization_networks.equivariant_networks import VNSmall
    out = {}
    K, W = args.steps, args.warmup
    sync = world > 1

    # ---- configs[2]: OptimizedGroupEquivariant D4, 3x224x224 ------------------------------------------------------
    torch.manual_seed(0)
    net = conv_network((3, RESIZE, RESIZE), 16, 7, 3, 128).to(dev).eval()
    hp = SimpleNamespace(beta=1.0, input_crop_ratio=CROP, resize_shape=RESIZE, group_type="roto-reflection", num_rotations=4,
                         artifact_err_wt=0, learn_ref_vec=False)
    can = OptimizedGroupEquivariantImageCanonicalization(net, hp, IN_SHAPE).to(dev).eval()
    can.sync_prior_across_ranks = can.prefetch_prior_allreduce = sync
    hot_only = OptimizedGroupEquivariantImageCanonicalization(_FixedVectors(128), hp, IN_SHAPE).to(dev).eval()
    hot_only.sync_prior_across_ranks = hot_only.prefetch_prior_allreduce = sync
    orbit_bytes = (3 * 180 * 180 + 3 * 96 * 96) * 4 + (3 * 96 * 96 + 8 * 3 * 96 * 96) * 4 + 2 * IMG_BYTES   # resize + expand + 2 warps
    for tag, b in (("strong", max(256 // world, 1)), ("weak", 256)):
        x = torch.rand(b, *IN_SHAPE, generator=torch.Generator().manual_seed(11 + rank)).to(dev)
        ms, mode = measure(lambda t: public_step(can, t), (x,), K, W, barrier, world, dev)
        ms_hot, mode_hot = measure(lambda t: public_step(hot_only, t), (x,), K, W, barrier, world, dev)
        out.setdefault("cfg3_opt_d4_images", {})[tag] = {
            "per_gpu_batch": b, "global_batch": b * world, "samples_per_s": b * world / (ms * 1e-3), "us": 1e3 * ms, "mode": mode,
            "hot_path_only": {"samples_per_s": b * world / (ms_hot * 1e-3), "us": 1e3 * ms_hot, "mode": mode_hot,
                              "achieved_gbs": orbit_bytes * b / (ms_hot * 1e-3) / 1e9,
                              "note": "crop+resize, |G|=8 orbit expand, cosine activations, select, canonicalize + invert warps; the "
                                      "orbit is scored by a constant stand-in instead of the caller's ConvNetwork (cuDNN)"}}
    out["cfg3_opt_d4_images"]["workload"] = ("BASELINE configs[2]: OptimizedGroupEquivariantImageCanonicalization, D4 (|G| = 8 orbit expand), "
                                             "3x224x224, ConvNetwork(16, k7, 3 layers, 128) as the caller's torch module; step = canonicalize "
                                             "+ invert(scalar) + prior loss + identity metric")
    del can, hot_only, net

    # ---- configs[3]: SO(3) point clouds ---------------------------------------------------------------------------
    torch.manual_seed(0)
    pc = EquivariantPointcloudCanonicalization(VNSmall(SimpleNamespace(n_knn=20, pooling="mean")).to(dev).eval(), SimpleNamespace()).eval()
    pc.sync_prior_across_ranks = sync

    def pc_step(x):
        return pc(x), pc.get_prior_regularization_loss()

    for tag, b in (("strong", max(128 // world, 1)), ("weak", 128)):
        x = torch.randn(b, 3, 1024, generator=torch.Generator().manual_seed(21 + rank)).to(dev)
        ms, mode = measure(pc_step, (x,), K, W, barrier, world, dev)
        out.setdefault("cfg4_pointcloud_so3", {})[tag] = {
            "per_gpu_batch": b, "global_batch": b * world, "samples_per_s": b * world / (ms * 1e-3), "us": 1e3 * ms, "mode": mode,
            "achieved_gbs": (2 * 3 * 1024 * 4 + 36) * b / (ms * 1e-3) / 1e9}
    out["cfg4_pointcloud_so3"]["workload"] = ("BASELINE configs[3]: EquivariantPointcloudCanonicalization, VNSmall(n_knn 20, mean) -> Gram-Schmidt -> "
                                              "R x on 1024x3 clouds + prior (MSE to identity); latency-bound (24 612 B/cloud, SURVEY 8d 'P'): "
                                              "GB/s is reported, not a roofline fraction")
    del pc

    # ---- configs[4]: E(3) n-body ------------------------------------------------------------------------------------
    for tag, S in (("strong", max(10000 // world, 1)), ("weak", 10000)):
        torch.manual_seed(0)
        hpn = SimpleNamespace(out_dim=4, hidden_dim=16, layer_pooling="mean", final_pooling="mean", num_layers=4, nonlinearity="relu",
                              canon_feature="p", canon_translation=False, angular_feature=0, dropout=0.5, batch_size=S)
        nb = EuclideanGroupNBody(VNDeepSets(hpn, device=str(dev)).eval()).eval()
        nb.sync_prior_across_ranks = sync
        g = torch.Generator().manual_seed(31 + rank)
        loc, vel = torch.randn(5 * S, 3, generator=g).to(dev), torch.randn(5 * S, 3, generator=g).to(dev)
        ch = (torch.randint(0, 2, (5 * S, 1), generator=g) * 2 - 1).float().to(dev)
        base = torch.tensor([[i, j] for i in range(5) for j in range(5) if i != j]).t()
        edges = (base[:, None, :] + 5 * torch.arange(S)[None, :, None]).reshape(2, -1).to(dev)
        nodes = torch.sqrt(torch.sum(vel ** 2, dim=1)).unsqueeze(1)

        def nb_step(nodes, loc, vel, ch, edges=edges, nb=nb):
            cl, cv = nb(nodes, None, loc=loc, edges=edges, vel=vel, edge_attr=None, charges=ch)
            back = nb.invert_canonicalization(cl)
            return cl, cv, back, nb.get_prior_regularization_loss()

        ms, mode = measure(nb_step, (nodes, loc, vel, ch), K, W, barrier, world, dev)
        out.setdefault("cfg5_nbody_e3", {})[tag] = {
            "per_gpu_systems": S, "global_systems": S * world, "samples_per_s": S * world / (ms * 1e-3), "us": 1e3 * ms, "mode": mode,
            "achieved_gbs": (288 + 168) * S / (ms * 1e-3) / 1e9}
        del nb
    out["cfg5_nbody_e3"]["workload"] = ("BASELINE configs[4]: EuclideanGroupNBody, VNDeepSets(4 layers, hidden 16) -> modified Gram-Schmidt -> "
                                        "canonicalize loc/vel + invert + prior statistic all-reduced over NCCL; 5-particle systems; "
                                        "latency-bound (456 B/system, SURVEY 8d 'B')")
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu"] = cpu_other_configs()
    return out


class _FixedVectors(torch.nn.Module):
    """Stand-in consumer for the hot-path-only reading of configs[2]: returns one constant row per orbit member."""

    def __init__(self, v):
        super().__init__()
        self.out_vector_size = v
        self.register_buffer("table", torch.randn(4096, v, generator=torch.Generator().manual_seed(5)))

    def forward(self, x):
        reps = -(-x.shape[0] // self.table.shape[0])
        return self.table.repeat(reps, 1)[: x.shape[0]]


def cpu_other_configs():
    """The reference's own classes (baseline/_ref) for configs[2..4] on the host cores, bounded samples; the oracle port of
    the same path when baseline/_ref is absent."""
    res = {"cores": os.cpu_count()}
    torch.set_num_threads(os.cpu_count() or 1)

    def clock(fn, reps):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    with torch.no_grad():
        if reference_available():
            import_reference()
            from equiadapt.images.canonicalization.discrete_group import OptimizedGroupEquivariantImageCanonicalization as RefOpt
            from equiadapt.images.canonicalization_networks.custom_nonequivariant_networks import ConvNetwork
            from equiadapt.nbody.canonicalization.euclidean_group import EuclideanGroupNBody as RefNB
            from equiadapt.nbody.canonicalization_networks.custom_equivariant_networks import VNDeepSets as RefDS
            from equiadapt.pointcloud.canonicalization.continuous_group import EquivariantPointcloudCanonicalization as RefPC
            from equiadapt.pointcloud.canonicalization_networks.equivariant_networks import VNSmall as RefVN
            res["kind"] = "reference"
            torch.manual_seed(0)
            hp = SimpleNamespace(beta=1.0, input_crop_ratio=CROP, resize_shape=RESIZE, group_type="roto-reflection", num_rotations=4,
                                 artifact_err_wt=0, learn_ref_vec=False)
            can = RefOpt(ConvNetwork((3, RESIZE, RESIZE), 16, 7, 3, 128), hp, IN_SHAPE).eval()
            x = torch.rand(16, *IN_SHAPE)
            t = clock(lambda: public_step(can, x), 2)
            res["cfg3_opt_d4_images"] = {"samples_per_s": 16 / t, "sample": "16 images x 2 steps"}
            pc = RefPC(RefVN(SimpleNamespace(n_knn=20, pooling="mean")).eval(), SimpleNamespace()).eval()
            xc = torch.randn(8, 3, 1024)
            t = clock(lambda: (pc(xc), pc.get_prior_regularization_loss()), 1)
            res["cfg4_pointcloud_so3"] = {"samples_per_s": 8 / t, "sample": "8 clouds x 1 step"}
            S = 1000
            hpn = SimpleNamespace(out_dim=4, hidden_dim=16, layer_pooling="mean", final_pooling="mean", num_layers=4, nonlinearity="relu",
                                  canon_feature="p", canon_translation=False, angular_feature=0, dropout=0.5, batch_size=S)
            nb = RefNB(RefDS(hpn, device="cpu").eval()).eval()
            loc, vel = torch.randn(5 * S, 3), torch.randn(5 * S, 3)
            ch = (torch.randint(0, 2, (5 * S, 1)) * 2 - 1).float()
            base = torch.tensor([[i, j] for i in range(5) for j in range(5) if i != j]).t()
            edges = (base[:, None, :] + 5 * torch.arange(S)[None, :, None]).reshape(2, -1)
            nodes = torch.sqrt(torch.sum(vel ** 2, dim=1)).unsqueeze(1)

            def nb_step():
                cl, cv = nb(nodes, None, loc=loc, edges=edges, vel=vel, edge_attr=None, charges=ch)
                return nb.invert_canonicalization(cl)      # (the reference has no working prior loss here: SURVEY 8e)

            t = clock(nb_step, 5)
            res["cfg5_nbody_e3"] = {"samples_per_s": S / t, "sample": "1 000 systems x 5 steps (canonicalize + invert)"}
        else:
            res["kind"] = "unavailable (baseline/_ref absent; run baseline/install_ref.sh)"
    return res


def index_agreement(can, x_dev, n=32):
    """ours vs the fp32 and fp64 oracle port on the first n images of THIS bench batch (the reference's own fp32 path
    disagrees with exact arithmetic on near-ties too: SURVEY.md 7, hard part 1)."""
    from oracle import reference_path as O
    net = make_layers()
    layers = [(m.weights.detach(), m.bias.detach()) for m in net.eqv_network if hasattr(m, "weights")]
    xs = x_dev[:n]
    sync, can.sync_prior_across_ranks, can.prefetch_prior_allreduce = can.sync_prior_across_ranks, False, False   # rank 0 alone calls this
    with torch.no_grad():
        can(xs)
        act = can.canonicalization_info_dict["group_activations"].float().cpu()
        ours = can.canonicalization_info_dict["group_element"].index.cpu().long()
        xc = xs.cpu()
        a32 = O.custom_equivariant_network(O.pre_network_transform(xc, IN_SHAPE, CROP, RESIZE), layers, N_ROT, False)
        l64 = [(w.double(), b.double()) for w, b in layers]
        a64 = O.custom_equivariant_network(O.pre_network_transform(xc.double(), IN_SHAPE, CROP, RESIZE), l64, N_ROT, False)
    can.sync_prior_across_ranks = can.prefetch_prior_allreduce = sync
    i32, i64 = a32.argmax(-1), a64.argmax(-1)
    top2 = a64.topk(2, dim=-1).values
    return {"n": n, "ours_vs_fp64": float((ours == i64).float().mean()), "fp32_port_vs_fp64": float((i32 == i64).float().mean()),
            "ours_vs_fp32_port": float((ours == i32).float().mean()),
            "act_max_abs_err_ours_vs_fp64": float((act.double() - a64).abs().max()),
            "act_max_abs_err_fp32_port_vs_fp64": float((a32.double() - a64).abs().max()),
            "fp64_top2_gap_median": float((top2[:, 0] - top2[:, 1]).median()), "fp64_top2_gap_min": float((top2[:, 0] - top2[:, 1]).min())}


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from equiadapt_b200 import native, ops
    from equiadapt_b200.images.canonicalization.discrete_group import GroupEquivariantImageCanonicalization

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (B200 arm) needs a CUDA device; there is no CPU fallback")
    if not os.path.exists(native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    placement = place_rank(local, world)      # before any pinned allocation: first touch decides the NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    net = make_layers().to(dev)
    can = GroupEquivariantImageCanonicalization(net, hp_image(), IN_SHAPE).eval()
    # N > 1: the prior statistic is all-reduced (opt-in), started right behind the select kernel so it runs beside the warps
    can.sync_prior_across_ranks = can.prefetch_prior_allreduce = world > 1
    B = args.batch
    x_host = host_batch(B, seed=1 + rank).pin_memory()
    z_host = torch.empty_like(x_host).pin_memory()
    x = x_host.to(dev)

    def step(xin):
        return public_step(can, xin)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step(x)
        barrier()
        l0 = ops.launch_count
        step(x)
        launches = ops.launch_count - l0
        graphed_step = None
        if not args.no_graph:
            graphed_step = can.capture_step(x, induced_rep_type="scalar", warmup=1)
            for _ in range(max(args.warmup, 3)):
                graphed_step()
        # ---------------- device-timed region: K steps, inputs resident in HBM ----------------
        clocks = ClockSampler(local) if rank == 0 else None
        run = (lambda: graphed_step()) if graphed_step is not None else (lambda: step(x))
        elapsed_ms, per_step, (t_begin, t_end) = timed_steps(run, args.steps, barrier)
        clock_info = clocks.stop(t_begin, t_end) if clocks is not None else None
        if graphed_step is not None:
            y_g, z_g, loss_g, ident_g = graphed_step.outputs
            graph_out = (z_g.clone(), float(loss_g), float(ident_g))
        # ---------------- the same K steps through the eager public calls, with per-call events ----------------
        ops.event_log = {}
        eager_ms, eager_per, _ = timed_steps(lambda: step(x), args.steps, barrier)
        log, ops.event_log = ops.event_log, None
        kernels_pass = "the eager pass"
        if world > 1:
            # Per-kernel times for `kernels` / `rooflines` WITHOUT the collective in flight: the all-reduce kernel, issued beside
            # the warp kernels, waits for the slowest rank ON the SMs, and in an eager pass that wait is another rank's Python
            # thread (warps measured 125 -> 170 us at N = 2..8 that way; the graph-timed step shows +40 us for the whole step).
            s0, p0 = can.sync_prior_across_ranks, can.prefetch_prior_allreduce
            can.sync_prior_across_ranks = can.prefetch_prior_allreduce = False
            ops.event_log = {}
            timed_steps(lambda: step(x), args.steps, barrier)
            log, ops.event_log = ops.event_log, None
            can.sync_prior_across_ranks, can.prefetch_prior_allreduce = s0, p0
            kernels_pass = "a second eager pass with the prior all-reduce switched off (rank-local statistic)"
        z_e, loss_e, ident_e = step(x)
        graph_equal = None
        if graphed_step is not None:
            graph_equal = bool(torch.equal(graph_out[0], z_e)) and graph_out[1] == float(loss_e)
            del graph_out
        # ---------------- end-to-end: pinned host -> device -> step -> pinned host ----------------
        for _ in range(2):
            xd = x_host.to(dev, non_blocking=True)
            z, loss, ident = step(xd)
            z_host.copy_(z, non_blocking=True)
            float(loss)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            xd = x_host.to(dev, non_blocking=True)
            z, loss, ident = step(xd)
            z_host.copy_(z, non_blocking=True)
            loss_v, ident_v = float(loss), float(ident)   # device -> host read of the step's metrics (syncs)
        torch.cuda.synchronize()
        e2e_serial_s = time.perf_counter() - t0
        # the package's host-buffer front end: same step, batch sharded over a 3-stream copy/compute pipeline
        from equiadapt_b200.host_pipeline import HostStreamedCanonicalizer
        pipe = HostStreamedCanonicalizer(can, None, "scalar", shard=args.e2e_shard, slots=3, device=dev, ramp=bool(args.e2e_ramp))
        z_serial = z_host.clone()
        z_host.zero_()
        for _ in range(2):
            loss, ident = pipe(x_host, z_host)
            float(loss)
        torch.cuda.synchronize()
        pipe_equal = bool(torch.equal(z_serial, z_host))   # sharding the batch must not change any output bit
        del z_serial
        barrier()
        t0 = time.perf_counter()
        e2e_per = []
        for _ in range(args.e2e_steps):
            t1 = time.perf_counter()
            loss, ident = pipe(x_host, z_host)
            loss_p, ident_p = float(loss), float(ident)
            torch.cuda.synchronize()
            e2e_per.append(1e3 * (time.perf_counter() - t1))
        e2e_s = time.perf_counter() - t0
        # informational: the training-loop shape of the same step - inputs from pinned host memory, only the loss and
        # the metric read back (the canonicalized / inverted batches stay on the device for their consumer)
        for _ in range(2):
            loss, ident = pipe(x_host, None)
            float(loss)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            loss, ident = pipe(x_host, None)
            float(loss), float(ident)
            torch.cuda.synchronize()
        e2e_in_s = time.perf_counter() - t0
        # what the link allows: plain pinned copies of the same buffers (the e2e step moves exactly these bytes),
        # every rank at the same time
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        zd = torch.empty_like(x)

        def timed(fn, reps=3):
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps

        def both():
            with torch.cuda.stream(s1):
                x.copy_(x_host, non_blocking=True)
            with torch.cuda.stream(s2):
                z_host.copy_(zd, non_blocking=True)

        gb = x_host.numel() * 4 / 1e9
        t_h2d = timed(lambda: x.copy_(x_host, non_blocking=True))
        t_d2h = timed(lambda: z_host.copy_(zd, non_blocking=True))
        t_both = timed(both)
        del zd

    t = torch.tensor([elapsed_ms, e2e_s, e2e_serial_s, eager_ms, t_both, max(per_step), max(e2e_per)], dtype=torch.float64, device=dev)
    tmin = torch.tensor([min(per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    elapsed_ms, e2e_s, e2e_serial_s, eager_ms, t_both_max = float(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[4])
    pcie = {"h2d_gbs": gb / t_h2d, "d2h_gbs": gb / t_d2h, "bidir_ms": 1e3 * t_both_max, "bound_img_s": B * world / t_both_max,
            "note": "plain pinned-memory copies of one batch each way on two streams, all ranks at once, max over ranks: the floor of any e2e step"}
    kernels = {}
    for name, evs in log.items():
        ms = [a.elapsed_time(b) for a, b in evs]
        kernels[name] = {"avg_us": 1e3 * sum(ms) / len(ms), "calls_per_step": len(ms) / args.steps}

    pk = peaks()
    other = None
    if not args.no_configs:
        with torch.no_grad():
            other = bench_other_configs(args, dev, rank, world, barrier, pk)

    if rank == 0:
        traffic, traffic_src = ncu_traffic()
        ms_per_step = elapsed_ms / args.steps
        value = B * world * args.steps / (elapsed_ms * 1e-3)
        step_us = sum(k["avg_us"] * k["calls_per_step"] for k in kernels.values())
        dom = max(kernels, key=lambda n: kernels[n]["avg_us"] * kernels[n]["calls_per_step"])
        rooflines = {}
        stack_key = next((k for k in kernels if k.startswith("eqb_gconv_stack_run")), None)
        if stack_key:
            # SURVEY.md 8d "N": algorithmic work = the contraction as the REFERENCE computes it (2.544 GFLOP/img);
            # the kernel executes 1.434 GFLOP/img (last layer folded through the mean) x 3 (fp16 hi/lo operand split).
            # Peak = the BURST bf16 figure (a kernel timed alone over a 40 ms region), the sustained one beside it.
            us = kernels[stack_key]["avg_us"]
            ach = STACK_FLOP_REFERENCE * B / (us * 1e-6) / 1e12
            rooflines[stack_key] = {
                "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops"], "frac_of_sustained_peak": ach / pk["bf16_tflops_sustained"],
                "traffic": traffic.get("eqb_gconv_stack_run"),
                "executed_tflops": 3 * STACK_FLOP_EXECUTED * B / (us * 1e-6) / 1e12,
                "note": ("algorithmic FLOPs = the reference's dense contraction, 2.544 GFLOP/img (SURVEY 8d), against the bf16 "
                         "tensor peak (" + pk["source"] + ", burst); executed_tflops = what the tensor pipe runs: 1.434 GFLOP/img "
                         "(last layer folded) x 3 fp16 hi/lo products of kind::f16 MMA; the call = CTA-pair tcgen05 stack kernel + "
                         "finish kernel (fold, group select, prior statistic)")}
        for prefix, byt in (("eqb_warp_canonicalize", IMG_BYTES), ("eqb_warp_invert", IMG_BYTES),
                            ("eqb_crop_resize_aa", (3 * 180 * 180 + 3 * 96 * 96) * 4)):
            name = next((k for k in kernels if k.startswith(prefix)), None)
            if name:
                ach = byt * B / (kernels[name]["avg_us"] * 1e-6) / 1e9
                rooflines[name] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                   "frac": ach / pk["hbm_gbs"], "traffic": traffic.get(prefix)}
        for name in kernels:
            kernels[name]["share_of_step"] = kernels[name]["avg_us"] * kernels[name]["calls_per_step"] / step_us
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = time_cpu_reference(args.cpu_sample_batch, steps=12, warmup=1, budget_s=20.0)
            cpu = {"value": r["img_s"], "unit": "img/s", "cores": r["cores"], "kind": r["kind"], "sample": cpu_sample_text(r)}
        checks = {"prior_loss": loss_v, "identity_metric": ident_v, "prior_loss_pipelined": loss_p,
                  "identity_metric_pipelined": ident_p, "pipelined_output_equals_unsharded": pipe_equal,
                  "graph_output_equals_eager": graph_equal}
        if not args.no_cpu_baseline:
            checks["index_agreement"] = index_agreement(can, x, 32)
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(B, world),
            "timed_region": "cuda graph replay of the public step (canonicalizer.capture_step)" if graphed_step is not None else "eager public calls",
            "step_ms": dict(step_stats(per_step), max_over_ranks=float(t[5]), min_over_ranks=float(tmin[0])),
            "eager": {"value": B * world * args.steps / (eager_ms * 1e-3), "ms_per_step": eager_ms / args.steps,
                      "step_ms": step_stats(eager_per),
                      "note": "same K steps as eager ctypes calls with two CUDA events around each; `kernels` / `rooflines` come from " + kernels_pass},
            "clocks": clock_info,
            "e2e": {"value": B * world * args.e2e_steps / e2e_s, "unit": "img/s",
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": z_host.numel() * 4 + 8,
                    "ms_per_step": 1e3 * e2e_s / args.e2e_steps, "step_ms": dict(step_stats(e2e_per), max_over_ranks=float(t[6])),
                    "frac_of_pcie_bound": (B * world * args.e2e_steps / e2e_s) / pcie["bound_img_s"],
                    "how": f"HostStreamedCanonicalizer: pinned host -> {args.e2e_shard}-image shards over h2d/compute/d2h "
                           "streams -> pinned host, loss + metric read back every step",
                    "unpipelined_value": B * world * args.e2e_steps / e2e_serial_s,
                    "inputs_only_value": B * world * args.e2e_steps / e2e_in_s,
                    "inputs_only_note": "same step with only loss + metric read back (8 bytes D2H): not max-reduced over ranks"},
            "pcie": pcie,
            "placement": placement,
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "roofline": dict(rooflines.get(dom, {}), kernel=dom),
            "rooflines": rooflines,
            "traffic_source": traffic_src,
            "kernels": kernels,
            "cpu_baseline": cpu,
            "configs": other,
            "checks": checks,
        }
        print(json.dumps(line))
    # captured graphs hold NCCL kernels: they must be gone before the communicator is torn down (ncclCommDestroy waits on them)
    graphed_step = None
    shutdown(world)


def shutdown(world):
    import gc
    import threading
    import torch.distributed as dist
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    if world > 1:
        dist.barrier()
        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start()
        th.join(20.0)
        if th.is_alive():          # never let a teardown problem turn a finished measurement into a hung job
            sys.stderr.write("bench.py: destroy_process_group did not return within 20 s; exiting\n")
            sys.stderr.flush()
            os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--batch", type=int, default=512, help="images per GPU (weak scaling)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --batch is the GLOBAL batch, each of the N ranks takes batch / N images (SURVEY 8e asks "
                         "for both readings; the headline and the driver's runs are weak)")
    ap.add_argument("--no-graph", action="store_true", help="time the eager public calls instead of the captured step")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs[2..4]")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-shard", type=int, default=64, help="images per stage of the host-buffer pipeline")
    ap.add_argument("--e2e-ramp", type=int, default=0, help="1: quarter / half shards at both ends of the pipeline (shorter fill and drain)")
    ap.add_argument("--cpu-sample-batch", type=int, default=32)
    ap.add_argument("--ref-gpu-chunk", type=int, default=32, help="--impl reference-gpu: images per eager call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.scaling = "weak"
    if args.strong:
        if args.batch % max(args.gpus, 1):
            ap.error("--strong needs --batch divisible by --gpus")
        args.batch //= max(args.gpus, 1)
        args.scaling = "strong"
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
