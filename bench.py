#!/usr/bin/env python
"""Benchmark of the CTC loss+gradient hot path (BASELINE.json metric) on 1..8 B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg4]
                  [--variant simplified|classic] [--scaling strong|weak] [--ragged] [--e2e-grad]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one fused loss + d/dlogits call over one synthetic batch (logits ~ N(0,1), labels ~ U{1..V-1}).
Rank 0 prints ONE JSON line.  See DESIGN.md section "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, T, V, L, default variant) -- BASELINE.json configs[1], [2], [4] (per-GPU slice for cfg4)
    "cfg1": (32, 500, 29, 100, "classic"),
    "cfg2": (256, 1000, 1024, 200, "simplified"),
    "cfg4": (256, 1600, 5000, 400, "classic"),
    # BASELINE.json configs[4] at its full batch on one GPU: 65.5 GB logits + 65.5 GB gradient + 11 GB workspace; inputs are
    # generated on the device (no host copy exists, so this workload has no e2e / CPU leg)
    "cfg4full": (2048, 1600, 5000, 400, "classic"),
}
METRIC = "CTC loss+grad samples/s at B=256 T=1000 V=1024 L=200"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default=None, choices=["simplified", "classic"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): the named batch is cut into N contiguous slices, one per GPU (SURVEY.md 8e); "
                         "weak: every GPU processes the whole named batch.  The other one is reported under 'secondary'.")
    ap.add_argument("--staged", action="store_true", help="force the three staged kernels instead of the fused one")
    ap.add_argument("--ragged", action="store_true", help="logit_length~U[T/2,T], label_length~U[L/2,L]")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--dtype-in", default="f32", choices=["f32", "bf16"],
                    help="bf16: logits are bfloat16 in HBM / host memory (CTCB200_LOGITS_BF16, an extension of the reference's "
                         "float32-only interface); arithmetic stays fp32.  A SECONDARY line: the headline is f32")
    ap.add_argument("--dtype-grad", default="f32", choices=["f32", "bf16"], help="with --dtype-in bf16: gradient format")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other scaling mode's measurement at N > 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=64, help="utterances timed by the CPU baseline leg")
    return ap.parse_args()


def synth(B, T, V, L, seed, ragged):
    import torch
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn((B, T, V), generator=g, dtype=torch.float32)
    labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32)
    if ragged:
        logit_length = torch.randint(T // 2, T + 1, (B,), generator=g, dtype=torch.int32)
        label_length = torch.randint(L // 2, L + 1, (B,), generator=g, dtype=torch.int32)
    else:
        logit_length = torch.full((B,), T, dtype=torch.int32)
        label_length = torch.full((B,), L, dtype=torch.int32)
    return logits, labels, label_length, logit_length


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in names.items():
                    if bit and (mask & bit):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.004)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


def cpu_port_samples_per_s(variant_id, logits, labels, label_length, logit_length, reps):
    """The oracle's C restatement (float32 arithmetic like the reference, one pthread per host core)."""
    from oracle import c_oracle
    import numpy as np
    x, lab = logits.numpy(), labels.numpy()
    ll, tl = label_length.numpy(), logit_length.numpy()
    c_oracle.loss_grad(lab[:2], x[:2], ll[:2], tl[:2], 0, variant_id, dtype=np.float32)     # load + warm
    t0 = time.perf_counter()
    for _ in range(reps):
        c_oracle.loss_grad(lab, x, ll, tl, 0, variant_id, dtype=np.float32)
    dt = (time.perf_counter() - t0) / reps
    return x.shape[0] / dt, dt, c_oracle.max_threads()


def run_reference(args, rank, world, out):
    """--impl reference: the reference's algorithm on the host cores (oracle port; TensorFlow is not installable
    here, so the reference itself cannot run).  Rank 0 only."""
    if rank != 0:
        return
    B, T, V, L, dvar = WORKLOADS[args.workload]
    variant = args.variant or dvar
    vid = 1 if variant == "simplified" else 0
    nb = min(B, args.cpu_sample)
    logits, labels, ll, tl = synth(nb, T, V, L, 0, args.ragged)
    from oracle import c_oracle
    import numpy as np
    x, lab = logits.numpy(), labels.numpy()
    for _ in range(max(args.warmup, 1)):
        c_oracle.loss_grad(lab[:8], x[:8], ll.numpy()[:8], tl.numpy()[:8], 0, vid, dtype=np.float32)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.loss_grad(lab, x, ll.numpy(), tl.numpy(), 0, vid, dtype=np.float32)
    dt = (time.perf_counter() - t0) / args.steps
    value = nb / dt
    cores = c_oracle.max_threads()
    sample = f"{nb} of {B} utterances per step (same T,V,L), C port of the reference algorithm in float32, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC if args.workload == "cfg2" else f"CTC loss+grad samples/s ({args.workload})",
        "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {variant}_ctc_loss B={B} T={T} V={V} L={L}", "variant": variant,
                   "ragged": bool(args.ragged)},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=out, flush=True)


def bind_to_gpu_numa_node(index):
    """Best effort: run this rank's host threads (and so its first-touch pinned allocations) on the CPUs NVML lists as
    local to the GPU.  On boxes where every GPU reports the same node this is a no-op."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def claim_stdout():
    """stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version banner on communicator
    creation): from here on file descriptor 1 points at stderr and the returned handle is the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    args = parse_args()
    out = claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch
    import torch.distributed as dist
    from tf_seq2seq_losses_b200 import _lib, shard_bounds

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, T, V, L, dvar = WORKLOADS[args.workload]
    variant = args.variant or dvar
    vid = _lib.SIMPLIFIED if variant == "simplified" else _lib.CLASSIC
    if args.scaling == "weak":
        local_B, global_B = B, B * world           # every GPU processes the named batch
    else:
        b0, b1 = shard_bounds(B, world, rank)      # the named batch is cut into contiguous slices
        local_B, global_B = b1 - b0, B
    on_device = local_B * T * V * 4 > 16e9           # too large for a host copy: generate on the device
    if on_device:
        args.no_e2e = args.no_cpu_baseline = True
        gen = torch.Generator(device=dev).manual_seed(1000 + rank)
        logits = torch.empty((local_B, T, V), dtype=torch.float32, device=dev)
        for b0 in range(0, local_B, 64):
            logits[b0:b0 + 64].normal_(generator=gen)
        labels = torch.randint(1, V, (local_B, L), generator=gen, dtype=torch.int32, device=dev)
        tl = torch.full((local_B,), T, dtype=torch.int32, device=dev)
        ll = torch.full((local_B,), L, dtype=torch.int32, device=dev)
        if args.ragged:
            tl = torch.randint(T // 2, T + 1, (local_B,), generator=gen, dtype=torch.int32, device=dev)
            ll = torch.randint(L // 2, L + 1, (local_B,), generator=gen, dtype=torch.int32, device=dev)
    else:
        logits_h, labels_h, ll_h, tl_h = synth(local_B, T, V, L, 1000 + rank, args.ragged)
        logits, labels, ll, tl = logits_h.to(dev), labels_h.to(dev), ll_h.to(dev), tl_h.to(dev)
    es_in, es_out = (2 if args.dtype_in == "bf16" else 4), (2 if args.dtype_grad == "bf16" else 4)
    assert es_out == 4 or es_in == 2, "--dtype-grad bf16 needs --dtype-in bf16"
    if es_in == 2:
        logits = logits.to(torch.bfloat16)
        if not on_device:
            logits_h = logits_h.to(torch.bfloat16)
    dflags = (_lib.FORCE_STAGED if args.staged else 0) | (_lib.GRAD_BF16 if es_out == 2 else 0)
    desc = _lib.make_desc(logits, labels, 0, vid, L + 1, dflags)
    lib = _lib.load()
    ws = torch.empty(max(lib.ctcb200_workspace_bytes(ctypes.byref(desc), _lib.WS_LOSS_GRAD_LOGITS), 256), dtype=torch.uint8, device=dev)
    loss = torch.empty((local_B,), dtype=torch.float32, device=dev)
    grad = torch.empty(logits.shape, dtype=torch.bfloat16 if es_out == 2 else torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())

    def step(d=desc):
        _lib.check(lib.ctcb200_loss_grad(ctypes.byref(d), P(logits), P(labels), P(ll), P(tl), None, P(loss), P(grad), None,
                                         P(ws), ws.numel(), ctypes.c_void_p(stream.cuda_stream)))

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        sync_all()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = global_B / (ms_step * 1e-3)

    # ---- per-stage timing for the roofline of the dominant kernel (stages re-run on the filled workspace) ----
    launches_per_step = int(lib.ctcb200_launches_per_call(ctypes.byref(desc))) if hasattr(lib, "ctcb200_launches_per_call") else 3
    stage_ms = {}
    if hasattr(lib, "ctcb200_stage_names"):
        names = lib.ctcb200_stage_names(ctypes.byref(desc)).decode().split(",")
        for i, name in enumerate(names):
            d = _lib.Desc(desc.B, desc.T, desc.V, desc.Lw, desc.blank, desc.variant, desc.U, desc.flags | ((1 << i) << 8))
            for _ in range(3):
                step(d)
            torch.cuda.synchronize(dev)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            for _ in range(args.steps):
                step(d)
            s1.record(stream)
            torch.cuda.synchronize(dev)
            stage_ms[name] = s0.elapsed_time(s1) / args.steps
        step()                                   # leave the workspace consistent
        torch.cuda.synchronize(dev)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # logits read once + gradient written once (+ labels, lengths, loss).  With ragged lengths only the
    # frames below logit_length are read; the gradient is still written for all T frames (zeros beyond the length).
    frames_read = int(tl.clamp(0, T).sum().item())
    alg_bytes = V * (es_in * frames_read + es_out * local_B * T) + local_B * (4 * L + 12)
    if stage_ms:
        dom = max(stage_ms, key=stage_ms.get)
        dom_ms = stage_ms[dom]
    else:
        dom, dom_ms = "whole call", ms_step
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json), if this
    # workload / kernel was captured
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = f"{args.workload}:{variant}:{dom}"
        if key in tr and local_B == B and not args.ragged:
            traffic = tr[key]["traffic_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": dom_ms, "stage_ms": stage_ms,
                "path_achieved": alg_bytes / (ms_step * 1e-3) / 1e9, "path_frac": alg_bytes / (ms_step * 1e-3) / 1e9 / peak}

    # ---- the other scaling mode, measured in the same run (N > 1 only; at N = 1 the two coincide) ----
    secondary = None
    if world > 1 and not args.no_secondary and not on_device:
        if args.scaling == "strong":
            sB, s_global, s_name = B, B * world, "weak"
        else:
            b0, b1 = shard_bounds(B, world, rank)
            sB, s_global, s_name = b1 - b0, B, "strong"
        g2 = torch.Generator(device=dev).manual_seed(2000 + rank)
        x2 = torch.empty((sB, T, V), dtype=torch.float32, device=dev).normal_(generator=g2)
        if es_in == 2:
            x2 = x2.to(torch.bfloat16)
        lab2 = torch.randint(1, V, (sB, L), generator=g2, dtype=torch.int32, device=dev)
        tl2 = torch.full((sB,), T, dtype=torch.int32, device=dev)
        ll2 = torch.full((sB,), L, dtype=torch.int32, device=dev)
        if args.ragged:
            tl2 = torch.randint(T // 2, T + 1, (sB,), generator=g2, dtype=torch.int32, device=dev)
            ll2 = torch.randint(L // 2, L + 1, (sB,), generator=g2, dtype=torch.int32, device=dev)
        d2 = _lib.make_desc(x2, lab2, 0, vid, L + 1, dflags)
        ws2 = torch.empty(max(lib.ctcb200_workspace_bytes(ctypes.byref(d2), _lib.WS_LOSS_GRAD_LOGITS), 256), dtype=torch.uint8, device=dev)
        loss2, grad2 = torch.empty((sB,), dtype=torch.float32, device=dev), torch.empty(x2.shape, dtype=grad.dtype, device=dev)

        def step2():
            _lib.check(lib.ctcb200_loss_grad(ctypes.byref(d2), P(x2), P(lab2), P(ll2), P(tl2), None, P(loss2), P(grad2), None,
                                             P(ws2), ws2.numel(), ctypes.c_void_p(stream.cuda_stream)))
        for _ in range(max(args.warmup, 3)):
            step2()
        sync_all()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            step2()
        f1.record(stream)
        sync_all()
        t2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item()) / args.steps
        secondary = {"scaling": s_name, "value": s_global / (ms2 * 1e-3), "unit": "samples/s", "ms_per_step": ms2,
                     "per_gpu_batch": sB, "global_batch": s_global}
        del x2, grad2, ws2

    # ---- end to end: pinned host buffers -> C ABI host entry point -> loss back on the host ----
    e2e = None
    if not args.no_e2e:
        del grad, ws
        torch.cuda.empty_cache()
        affinity_before = os.sched_getaffinity(0)
        bind_to_gpu_numa_node(local_rank)
        n_slices = max(1, min(8, local_B // 16))     # >= 16 utterances per slice: below that the T-step chain, not the copy, paces a slice
        ctx = _lib.HostContext(local_B, T, V, L, 0, vid, L + 1, device=local_rank, num_slices=n_slices,
                               flags=desc.flags & (_lib.LOGITS_BF16 | _lib.GRAD_BF16))
        pin = [t.pin_memory() for t in (logits_h, labels_h, ll_h, tl_h)]
        loss_pin = torch.empty((local_B,), dtype=torch.float32).pin_memory()

        def timed_e2e(grad_out):
            for _ in range(2):
                ctx.loss_grad(*pin, loss_pin, grad_out)
            sync_all()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                ctx.loss_grad(*pin, loss_pin, grad_out)          # blocks until the results are in host memory
            dt = (time.perf_counter() - t0) / args.e2e_steps
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        dt = timed_e2e(None)
        h2d = int(sum(p.numel() * p.element_size() for p in pin))
        e2e = {"value": global_B / dt, "unit": "samples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(loss_pin.numel() * 4), "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "returns": "loss [B] to the host; the gradient stays resident on the device for the optimizer "
                          "(ctcb200_host_grad_device_ptr) -- see with_gradient_to_host for the full round trip",
               "api": f"ctcb200_host_loss_grad (pinned host buffers, {n_slices} slices; H2D stream + kernel stream + D2H stream)"}
        assert torch.equal(loss_pin, loss.cpu()), "host entry point disagrees with the device entry point"
        # the same call with the [B,T,V] gradient copied back to pinned host memory as well (PCIe is full duplex: the
        # read-back of slice i overlaps the upload of slice i+1)
        grad_pin = torch.empty((local_B, T, V), dtype=torch.bfloat16 if es_out == 2 else torch.float32).pin_memory()
        dtg = timed_e2e(grad_pin)
        e2e["with_gradient_to_host"] = {"value": global_B / dtg, "unit": "samples/s", "ms_per_step": dtg * 1e3,
                                        "d2h_bytes_per_step": int(loss_pin.numel() * 4 + grad_pin.numel() * es_out)}
        # the ceiling next to it: the bare host->device copy of the same pinned bytes, all ranks at once, no kernel
        dst = torch.empty_like(logits)
        for _ in range(2):
            dst.copy_(pin[0], non_blocking=True)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            dst.copy_(pin[0], non_blocking=True)
        torch.cuda.synchronize(dev)
        tc = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        e2e["h2d_copy_only"] = {"ms_per_step": float(tc.item()) * 1e3, "gb_per_s_per_gpu": pin[0].numel() * es_in / float(tc.item()) / 1e9,
                                "note": "cudaMemcpyAsync of the same logits from pinned memory on every rank at once (max over ranks): "
                                        "the PCIe / host-memory ceiling of the e2e figure"}
        del dst, grad_pin
        ctx.close()
        os.sched_setaffinity(0, affinity_before)      # the CPU baseline below uses every host core

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # the CPU leg is an N=1 report (the reference arm re-times it at every N)
        nb = min(local_B, args.cpu_sample)
        v, dt, cores = cpu_port_samples_per_s(1 if variant == "simplified" else 0, logits_h[:nb].float(), labels_h[:nb], ll_h[:nb], tl_h[:nb], 2)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{nb} of {local_B} utterances (same T,V,L), 2 repetitions, C port of the reference algorithm in float32"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC if args.workload == "cfg2" else f"CTC loss+grad samples/s ({args.workload})",
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32" if es_in == 4 else f"f32 arithmetic on bf16 logits ({args.dtype_grad} gradient) -- secondary line, not the headline",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {variant}_ctc_loss B={B} T={T} V={V} L={L}", "variant": variant,
                       "per_gpu_batch": local_B, "global_batch": global_B, "ragged": bool(args.ragged),
                       "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": "inputs exceed L2 (logits+grad per step >> 126 MB), no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "secondary": secondary,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks.summary(),
        }), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
