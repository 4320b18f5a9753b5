#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: canonicalized img/s at 3x224x224, C8 (configs[1]).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

One "step" = one pass of the hot path over one batch of 512 synthetic images per GPU:
    y = canonicalizer(x)                        crop+antialiased resize -> fused C8 group-conv stack ->
                                                group pool/select (+ prior statistic) -> inverse-rotation warp
    z = canonicalizer.invert_canonicalization(y, induced_rep_type="scalar")     forward warp of the output
    prior loss + identity metric                (ONE 3-float NCCL all-reduce when N > 1)
The prediction network between canonicalize and invert (ResNet-50 in configs[1]) is the caller's
PyTorch module and is NOT part of the hot path, so it is not run or timed here.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM;
`e2e` = the same step through the public classes from/to pinned HOST buffers (H2D + D2H inside the
timed region); `roofline` = the dominant kernel against MEASURED_PEAKS.json; `cpu_baseline` = the
oracle's restatement of the reference path (torch CPU ops, as the reference itself runs) on this
box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch at the default batch of 512, from the ncu --set full
# captures summarised under profiles/r1h_summary.md (pass S: conv stack and warp kernels; pass Q: crop + resize)
NCU_TRAFFIC = {"eqb_gconv_stack_run": 57.0e6 + 2.0e6, "eqb_warp_canonicalize": 292.6e6 + 261.4e6,
               "eqb_warp_invert": 292.6e6 + 263.5e6, "eqb_crop_resize_aa": 247.8e6 + 43.3e6}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def make_layers(seed=0):
    """Weights exactly as the reference initialises them (kaiming-uniform a=sqrt(5), bias 0), seed 0, on CPU."""
    from equiadapt_b200.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork
    torch.manual_seed(seed)
    return CustomEquivariantNetwork((3, RESIZE, RESIZE), OUT_CH, KSIZE, "rotation", N_ROT, LAYERS, device="cpu")


def host_batch(batch, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, *IN_SHAPE, generator=g)


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------------
def oracle_step(O, x, layers):
    x_pre = O.pre_network_transform(x, IN_SHAPE, CROP, RESIZE)
    act = O.custom_equivariant_network(x_pre, layers, N_ROT, False)
    el = O.activations_to_group_element(act, N_ROT, False)
    y = O.canonicalize_image(x, el["rotation"], None)
    z = O.invert_image_features(y, el["rotation"], None, N_ROT, N_ROT, "scalar")
    return z, O.prior_loss_discrete(act), O.identity_metric_discrete(act)


def time_oracle(sample_batch, steps, warmup, budget_s=None):
    from oracle import reference_path as O
    import warnings
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    net = make_layers()
    layers = [(m.weights.detach(), m.bias.detach()) for m in net.eqv_network if hasattr(m, "weights")]
    x = host_batch(sample_batch, seed=1)
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            oracle_step(O, x, layers)
        t_all = time.perf_counter()
        for _ in range(steps):
            t0 = time.perf_counter()
            oracle_step(O, x, layers)
            times.append(time.perf_counter() - t0)
            if budget_s is not None and time.perf_counter() - t_all > budget_s and len(times) >= 3:
                break
    total = sum(times)
    return {"img_s": sample_batch * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "steps": len(times), "cores": torch.get_num_threads(), "sample_batch": sample_batch}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_oracle(args.cpu_sample_batch, args.steps, args.warmup)
    sample = (f"{r['steps']} steps x {r['sample_batch']} images of the same workload (the reference's B=512 transients "
              f"exceed 40 GB on CPU, SURVEY.md 8d), eval/no_grad, torch CPU ops as the reference itself dispatches")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["img_s"], "unit": "img/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus),
        "cpu_baseline": {"value": r["img_s"], "unit": "img/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["img_s"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(batch, n_gpus):
    return {
        "workload": ("BASELINE configs[1]: C8 GroupEquivariantImageCanonicalization canonicalize + invert(scalar) warp, "
                     "synthetic U[0,1) 3x224x224, batch 512 per GPU; CustomEquivariantNetwork out_channels 32, "
                     "kernel 5, 3 layers, input_crop_ratio 0.8, resize 96; prediction network excluded"),
        "per_gpu_batch": batch, "global_batch": batch * n_gpus, "parallelism": f"dp{n_gpus} (batch-sharded)",
        "l2": "inputs 308 MB per GPU > 126 MB L2 (no flush needed)",
        "collective": "none" if n_gpus == 1 else "one async 3-float NCCL all-reduce per step (prior statistic), off the warp's critical path",
    }


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons of one GPU, sampled every ~5 ms by an NVML thread while the timed region runs
    (EQB_CLOCK_SAMPLE_MS overrides; every rank polls its own GPU, so the rate is kept moderate: NVML calls take driver locks
    that CUDA launches of all ranks on the box also need)
    (nvidia-smi -lms needs ~100 ms to start and cannot resolve a 40 ms region; it is the fallback when NVML is
    not importable)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        import threading
        self.proc, self.thread, self.samples, self.stop_flag = None, None, [], False
        self.interval = float(os.environ.get("EQB_CLOCK_SAMPLE_MS", "5")) / 1e3
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

    def _run(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                mask = int(reasons_fn(self.h))
                try:
                    watts = n.nvmlDeviceGetPowerUsage(self.h) / 1e3
                except Exception:
                    watts = None
                self.samples.append((mhz, mask, watts))
            except Exception:
                pass
            time.sleep(self.interval)

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = sorted(s[0] for s in self.samples)
            reasons = sorted(k for k, b in bits.items() if any(s[1] & b for s in self.samples))
            power = [s[2] for s in self.samples if s[2] is not None]
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    net = make_layers().to(dev)
    can = GroupEquivariantImageCanonicalization(
        net, SimpleNamespace(beta=1.0, input_crop_ratio=CROP, resize_shape=RESIZE), IN_SHAPE).eval()
    # N > 1: start the 3-float all-reduce right behind the select kernel so it runs beside the warp kernels
    can.prefetch_prior_allreduce = world > 1
    B = args.batch
    x_host = host_batch(B, seed=1 + rank).pin_memory()
    z_host = torch.empty_like(x_host).pin_memory()
    x = x_host.to(dev)

    def step(xin):
        y = can(xin)
        z = can.invert_canonicalization(y, induced_rep_type="scalar")
        loss = can.get_prior_regularization_loss()
        ident = can.get_identity_metric()
        return z, loss, ident

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step(x)
        barrier()
        # ---------------- device-timed region: K steps, inputs resident in HBM ----------------
        clocks = ClockSampler(local)
        ops.event_log = {}
        launches0 = ops.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            z, loss, ident = step(x)
        e1.record()
        barrier()
        elapsed_ms = e0.elapsed_time(e1)
        launches = (ops.launch_count - launches0) // args.steps
        log, ops.event_log = ops.event_log, None
        clock_info = clocks.stop()
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
        for _ in range(args.e2e_steps):
            loss, ident = pipe(x_host, z_host)
            loss_p, ident_p = float(loss), float(ident)
            torch.cuda.synchronize()
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
        # what the link allows: plain pinned copies of the same buffers (the e2e step moves exactly these bytes)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        zd = torch.empty_like(x)

        def timed(fn, reps=3):
            torch.cuda.synchronize()
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
        pcie = {"h2d_gbs": gb / t_h2d, "d2h_gbs": gb / t_d2h, "bidir_ms": 1e3 * t_both,
                "bound_img_s": B * world / t_both,
                "note": "plain pinned-memory copies of one batch each way on two streams: the floor of any e2e step"}
        del zd

    t = torch.tensor([elapsed_ms, e2e_s, e2e_serial_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_s, e2e_serial_s = float(t[0]), float(t[1]), float(t[2])
    kernels = {}
    for name, evs in log.items():
        ms = [a.elapsed_time(b) for a, b in evs]
        kernels[name] = {"avg_us": 1e3 * sum(ms) / len(ms), "calls_per_step": len(ms) / args.steps}

    if rank == 0:
        pk = peaks()
        ms_per_step = elapsed_ms / args.steps
        value = B * world * args.steps / (elapsed_ms * 1e-3)
        step_us = sum(k["avg_us"] * k["calls_per_step"] for k in kernels.values())
        dom = max(kernels, key=lambda n: kernels[n]["avg_us"] * kernels[n]["calls_per_step"])
        rooflines = {}
        if "eqb_gconv_stack_run" in kernels:
            # SURVEY.md 8d "N": algorithmic work = the contraction as the REFERENCE computes it (2.544 GFLOP/img);
            # the kernel executes 1.434 GFLOP/img (last layer folded through the mean) x 3 (fp16 hi/lo operand split)
            us = kernels["eqb_gconv_stack_run"]["avg_us"]
            ach = STACK_FLOP_REFERENCE * B / (us * 1e-6) / 1e12
            rooflines["eqb_gconv_stack_run"] = {
                "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"], "traffic": NCU_TRAFFIC.get("eqb_gconv_stack_run"),
                "note": ("algorithmic FLOPs = the reference's dense contraction, 2.544 GFLOP/img (SURVEY 8d), against the bf16 "
                         "tensor peak (" + pk["source"] + ", sustained); executed on the tensor pipe: 1.434 GFLOP/img (last "
                         "layer folded) x 3 fp16 hi/lo products = "
                         f"{3 * STACK_FLOP_EXECUTED * B / (us * 1e-6) / 1e12:.1f} TFLOP/s of kind::f16 MMA; "
                         "call = absmax + tcgen05 CTA-pair stack (cta_group::2, stationary weights) + finish kernels")}
        for name, byt in (("eqb_warp_canonicalize", IMG_BYTES), ("eqb_warp_invert", IMG_BYTES),
                          ("eqb_crop_resize_aa", (3 * 180 * 180 + 3 * 96 * 96) * 4)):
            if name in kernels:
                ach = byt * B / (kernels[name]["avg_us"] * 1e-6) / 1e9
                rooflines[name] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                   "frac": ach / pk["hbm_gbs"], "traffic": NCU_TRAFFIC.get(name)}
        for name in kernels:
            kernels[name]["share_of_step"] = kernels[name]["avg_us"] * kernels[name]["calls_per_step"] / step_us
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = time_oracle(args.cpu_sample_batch, steps=12, warmup=1, budget_s=20.0)
            cpu = {"value": r["img_s"], "unit": "img/s", "cores": r["cores"], "kind": "port",
                   "sample": f"{r['steps']} steps x {r['sample_batch']} images of the same workload on the host cores "
                             "(oracle restatement of the reference path on torch CPU ops)"}
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(B, world),
            "clocks": clock_info,
            "e2e": {"value": B * world * args.e2e_steps / e2e_s, "unit": "img/s",
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": z_host.numel() * 4 + 8,
                    "ms_per_step": 1e3 * e2e_s / args.e2e_steps,
                    "how": f"HostStreamedCanonicalizer: pinned host -> {args.e2e_shard}-image shards over h2d/compute/d2h "
                           "streams -> pinned host, loss + metric read back every step",
                    "unpipelined_value": B * world * args.e2e_steps / e2e_serial_s,
                    "inputs_only_value": B * world * args.e2e_steps / e2e_in_s,
                    "inputs_only_note": "same step with only loss + metric read back (8 bytes D2H): not max-reduced over ranks"},
            "pcie": pcie,
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "roofline": dict(rooflines.get(dom, {}), kernel=dom),
            "rooflines": rooflines,
            "kernels": kernels,
            "cpu_baseline": cpu,
            "checks": {"prior_loss": loss_v, "identity_metric": ident_v, "prior_loss_pipelined": loss_p,
                       "identity_metric_pipelined": ident_p, "pipelined_output_equals_unsharded": pipe_equal},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="images per GPU (weak scaling)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --batch is the GLOBAL batch, each of the N ranks takes batch / N images (SURVEY 8e asks "
                         "for both readings; the headline and the driver's runs are weak)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-shard", type=int, default=64, help="images per stage of the host-buffer pipeline")
    ap.add_argument("--e2e-ramp", type=int, default=0, help="1: quarter / half shards at both ends of the pipeline (shorter fill and drain)")
    ap.add_argument("--cpu-sample-batch", type=int, default=32)
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
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
