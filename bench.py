#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: quantized-conv images/sec (ResNet-56 CIFAR-10,
PO2 4-bit QAT forward + STE backward + SGD, batch 128 per GPU) and quantizer GB/s vs HBM peak.

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One JSON line on rank 0.  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same
step fed from pinned host memory with the loss read back every step; `roofline` = the quantizer's
streaming kernel timed live with CUDA events against MEASURED_PEAKS.json; `cpu_baseline` = the
oracle's torch-CPU restatement of the same training step on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "resnet56_cifar10_po2_4bit_qat_fwd_bwd_sgd"
METRIC = "quantized-conv images/sec (ResNet-56 PO2 4-bit QAT fwd+STE bwd); quantizer GB/s vs HBM peak in `roofline`"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [t.strip() for t in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
def build_training(device, world, local_rank, batch):
    import po2_quantization_b200 as P
    from workloads import resnet_cifar
    torch.manual_seed(8)
    model = resnet_cifar(56, 10, P.PowerOfTwoQuantizer, 4).to(device).train()
    if os.environ.get("PO2_PREFETCH", "1") == "1":
        # quantize all 56 weights in ONE multi-tensor launch at the start of each forward instead of one
        # by one in front of each conv (po2_quantization_b200/prefetch.py); same arithmetic, same results
        P.enable_weight_prefetch(model)
    if world > 1:
        if os.environ.get("PO2_DDP", "0") == "1":
            # torch's DistributedDataParallel, as the reference wraps its model (train.py:153-155)
            model = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], gradient_as_bucket_view=True,
                                                        broadcast_buffers=False)
        else:
            # same arithmetic, one coalesced NCCL all-reduce(AVG) per step (distributed.BatchSharded)
            from po2_quantization_b200.distributed import BatchSharded
            model = BatchSharded(model)
    # reference train.py:51-56: SGD momentum 0.9, wd 1e-4, lr 0.1 * world
    opt = torch.optim.SGD(model.parameters(), lr=0.1 * world, momentum=0.9, weight_decay=1e-4)
    crit = nn.CrossEntropyLoss()
    return model, opt, crit


def _parallelism_note():
    from po2_quantization_b200 import batchnorm
    grads = ("torch DDP buckets" if os.environ.get("PO2_DDP", "0") == "1"
             else "coalesced NCCL all-reduce(AVG) of the gradients in 4 buckets, issued from grad hooks under backward")
    ex = [e for e in batchnorm._exchanges.values()]
    mode = os.environ.get("PO2_BN_EXCHANGE", "peer")
    bn = ("SyncBatchNorm statistics exchanged inside the BN kernels over NVLink peer stores" if any(e is not None for e in ex)
          else "per-rank BatchNorm statistics" if mode == "local" else "SyncBatchNorm statistics over NCCL all_gather/all_reduce")
    return f" (batch-sharded; {grads}; {bn})"


def _exchange_timeouts():
    """number of peer mailboxes whose poll ever timed out (must be 0)"""
    from po2_quantization_b200 import batchnorm
    return sum(1 for e in batchnorm._exchanges.values() if e is not None and e.error_flag() != 0)


def activation_bytes_estimate(batch):
    # saved activations of ResNet-56 at 32x32: 19 layer-1 convs+bn+relu at 16ch/32^2, 18 at 32ch/16^2, 18 at 64ch/8^2
    per_img = (19 * 16 * 32 * 32 + 18 * 32 * 16 * 16 + 18 * 64 * 8 * 8) * 4 * 3
    return per_img * batch


def run_ours(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from po2_quantization_b200 import ops

    # Everything (model/DDP construction, warm-up, capture, replay, timing events) runs on ONE side
    # stream: CUDA-graph capture is illegal on the legacy default stream, and DDP's AccumulateGrad
    # hooks must be created on the stream the captured step later runs on.
    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _run_ours_on_stream(a, ops, world, rank, local_rank, device)
    if world > 1:
        # A captured graph keeps NCCL work objects alive and destroy_process_group() can then wait
        # forever at interpreter exit; every rank is past its last collective here, so leave hard.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _run_ours_on_stream(a, ops, world, rank, local_rank, device):
    B = a.batch
    model, opt, crit = build_training(device, world, local_rank, B)
    g = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.randn(B, 3, 32, 32, generator=g).pin_memory()
    y_host = torch.randint(0, 10, (B,), generator=g).pin_memory()
    x_dev = x_host.to(device)
    y_dev = y_host.to(device)
    loss_buf = torch.zeros((), device=device)

    def step():
        opt.zero_grad()  # reference train.py:81 (set_to_none=True: no fill / accumulate kernels)
        loss = crit(model(x_dev), y_dev)
        loss.backward()
        if hasattr(model, "average_gradients"):
            model.average_gradients()
        opt.step()
        loss_buf.copy_(loss.detach())

    if a.torch_profile:
        # kernel-time table of eager steps from torch.profiler (works under torchrun, where ncu does not):
        # rank 0 writes the per-kernel totals of 5 steps to the given path
        from torch.profiler import ProfilerActivity, profile
        for _ in range(12):
            step()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                step()
            torch.cuda.synchronize()
        if rank == 0:
            rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
            rows.sort(key=lambda r: -r[2])
            tot = sum(r[2] for r in rows)
            with open(a.torch_profile, "w") as f:
                f.write(f"# 5 eager steps, world={world}; total device time {tot / 5:.1f} us per step\n")
                for k, c, t in rows:
                    f.write(f"{t / 5:10.1f} us/step {c // 5:5d} x  {100 * t / tot:5.1f}%  {k[:110]}\n")
            print(json.dumps({"torch_profile": a.torch_profile, "device_us_per_step": tot / 5}))
        return

    if a.profile_step:
        # one eager step between cudaProfilerStart/Stop: `ncu --profile-from-start off` lists exactly
        # the kernels of a step (profiles/ launch list)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        print(json.dumps({"profile_step": True, "launches_of_libpo2b200": ops.LAUNCHES}))
        return

    # ---- CUDA graph of the whole step.  With DDP (N>1) the NCCL all-reduce and the SyncBatchNorm
    # collectives are captured too (ProcessGroupNCCL supports capture); DDP needs its bucket rebuild
    # (iteration 2) to have happened, hence 11 eager warm-up steps on the capture stream first.
    graph = None
    launches_per_step = None
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if not a.no_graph:
        try:
            for _ in range(11 if world > 1 else 3):
                step()
            torch.cuda.current_stream().synchronize()
            if world > 1:
                torch.distributed.barrier()
            graph = torch.cuda.CUDAGraph()
            ops.LAUNCHES = 0
            with torch.cuda.graph(graph, stream=torch.cuda.current_stream()):
                step()
            launches_per_step = ops.LAUNCHES
            torch.cuda.synchronize()
        except Exception as e:  # pragma: no cover
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            graph = None
    if launches_per_step is None:
        ops.LAUNCHES = 0
        step()
        launches_per_step = ops.LAUNCHES
    run = graph.replay if graph is not None else step

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident timing
    for _ in range(max(a.warmup, 3)):
        run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for _ in range(a.steps):
            run()
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        # ---- end to end: pinned host -> device every step, loss read back every step
        for _ in range(2):
            x_dev.copy_(x_host, non_blocking=True); y_dev.copy_(y_host, non_blocking=True); run(); loss_buf.item()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        last = 0.0
        for _ in range(a.steps):
            x_dev.copy_(x_host, non_blocking=True)
            y_dev.copy_(y_host, non_blocking=True)
            run()
            last = loss_buf.item()            # device -> host read of the step's result
        f1.record()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))
        roof, extra = (quantizer_roofline(device, a) if rank == 0 else (None, None))
        conv_roof = conv_forward_roofline(device, B) if rank == 0 else None
    clocks = clk.summary()
    if world > 1:
        torch.distributed.barrier()

    if rank != 0:
        return
    ms_step = ms_total / a.steps
    out = {
        "metric": METRIC, "value": world * B * a.steps / (ms_total / 1e3), "unit": "images/s",
        "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (randn images 3x32x32, randint labels; kaiming-init weights, seed 8)",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}" + (_parallelism_note() if world > 1 else ""),
                   "cuda_graph": graph is not None,
                   "l2": "working set per step ~%d MB of saved activations > 126 MB L2; no flush needed"
                         % (activation_bytes_estimate(B) // 2 ** 20),
                   "conv_backend": ops.conv_backend_name(), "last_loss": last,
                   "norm_backend": "po2 FusedSyncBatchNorm kernels (norm + residual add + ReLU, forward and backward)",
                   "weight_quantization": ("one multi-tensor launch per step (prefetch)"
                                           if os.environ.get("PO2_PREFETCH", "1") == "1" else "one launch per layer"),
                   "bn_exchange_timeouts": _exchange_timeouts()},
        "e2e": {"value": world * B * a.steps / (ms_e2e / 1e3), "unit": "images/s",
                "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches_per_step * a.steps,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clocks,
        "roofline": roof, "roofline_extra": extra, "roofline_conv_forward": conv_roof,
    }
    if world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(B, steps=2)
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def quantizer_roofline(device, a):
    """Quantizer streaming kernels at 2^log2n elements, timed one launch at a time with CUDA events
    on the launching stream.  Algorithmic bytes: absmax reads es, quantize reads es
    and writes es per element."""
    from po2_quantization_b200 import ops
    pk = peaks()
    res = []
    for dt, es, name in ((torch.float32, 4, "f32"), (torch.bfloat16, 2, "bf16")):
        n = 1 << a.sweep_log2
        x = torch.randn(n, device=device, dtype=torch.float32).to(dt)
        y = torch.empty_like(x)
        s = torch.empty((), dtype=torch.float32, device=device)

        def timed(fn, iters=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
            for p, q in ev:
                p.record(); fn(); q.record()
            torch.cuda.synchronize()
            return sum(p.elapsed_time(q) for p, q in ev) / iters

        t_abs = timed(lambda: ops.absmax_out(x, s))
        t_q = timed(lambda: ops.quantize_out(x, y, s, 4, 1, False))
        t_all = timed(lambda: ops.quantize_fused_out(x, y, s, 4, 1, False))
        res.append({"kernel": f"po2::quantize_kernel<{name}> (pass 2 of po2_quantize_fused)", "dtype": name,
                    "elements": n, "bytes_per_element": 2 * es, "ms": t_q,
                    "achieved": 2 * es * n / t_q / 1e6, "absmax_ms": t_abs,
                    "absmax_GBs": es * n / t_abs / 1e6, "both_passes_ms": t_all,
                    "both_passes_GBs": 3 * es * n / t_all / 1e6})
        del x, y
    r0 = res[0]
    roof = {"bound": "hbm", "kernel": r0["kernel"], "achieved": r0["achieved"], "peak": pk["hbm_gbs"],
            "unit": "GB/s", "frac": r0["achieved"] / pk["hbm_gbs"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size from the committed
            # `ncu --set full` captures (profiles/r01_ncu_quantize_kernel_f32_2p28.csv: 1.074 + 1.028 GB;
            # profiles/r01_ncu_quantize_kernel_f32_2p30.csv: 4.295 + 4.248 GB)
            "traffic": {28: 2.102e9, 30: 8.543e9}.get(a.sweep_log2), "algorithmic_bytes": 8.0 * (1 << a.sweep_log2),
            "peak_source": pk["source"],
            "note": "algorithmic bytes = 8 B/element (4 read + 4 written) x 2^%d fp32 elements per launch; "
                    "inputs (%.1f GB) larger than L2; BASELINE's sweep tops out at 2^32 elements, see "
                    "profiles/r01_quantizer_sweep_2p30_2p32.json" % (a.sweep_log2, 4 * (1 << a.sweep_log2) / 1e9)}
    for r in res:
        r["frac_quantize_pass"] = r["achieved"] / pk["hbm_gbs"]
        r["frac_both_passes"] = r["both_passes_GBs"] / pk["hbm_gbs"]
        r["frac_absmax_pass"] = r["absmax_GBs"] / pk["hbm_gbs"]
    return roof, res


def conv_forward_roofline(device, batch):
    """Quantized-conv FORWARD of the ResNet-56 layer classes, timed live in CUDA graphs (the pack +
    conv kernels of one QuantizedConv2d.forward), with a cold L2 (a 320 MB buffer is rewritten before
    every conv).  These layers have 36-144 flop/B at the fp32 NCHW module boundary, so the binding
    roof is HBM: algorithmic bytes = 4*(B*C*H*W + B*K*P*Q)."""
    import torch.nn.functional as F
    from po2_quantization_b200 import ops
    pk = peaks()
    flush = torch.zeros(320 * 1024 * 1024 // 4, dtype=torch.int32, device=device)

    def graph_ms(body, reps=10, iters=5):
        body()
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
            for _ in range(reps):
                body()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2] / reps

    t_flush = graph_ms(lambda: flush.add_(1))
    layers = [("16->16 3x3 @32x32", 16, 32, 16, 18), ("32->32 3x3 @16x16", 32, 16, 32, 17), ("64->64 3x3 @8x8", 64, 8, 64, 17)]
    rows, tot_us, tot_cudnn, tot_flop, tot_bytes = [], 0.0, 0.0, 0.0, 0.0
    for name, C, HW, K, count in layers:
        x = torch.randn(batch, C, HW, HW, device=device)
        w = torch.randn(K, C, 3, 3, device=device) * 0.1
        y, _, scale, _, _ = torch.ops.po2.quantize_full(w, 4, 1, False)
        out = torch.empty(batch, K, HW, HW, device=device)

        def ours():
            flush.add_(1)
            ops.conv2d_out(x, y, scale, out, 1, 1, 1, 0)

        def cudnn():
            flush.add_(1)
            F.conv2d(x, y, None, 1, 1)
        us = (graph_ms(ours) - t_flush) * 1e3
        us_c = (graph_ms(cudnn) - t_flush) * 1e3
        flop = 2.0 * batch * K * HW * HW * C * 9
        byts = 4.0 * (x.numel() + out.numel())
        rows.append({"layer": name, "count_in_resnet56": count, "us": us, "us_cudnn_tf32": us_c,
                     "TFLOPs": flop / us / 1e6, "io_GBs": byts / us / 1e3, "frac_hbm": byts / us / 1e3 / pk["hbm_gbs"]})
        tot_us += us * count; tot_cudnn += us_c * count; tot_flop += flop * count; tot_bytes += byts * count
    return {"bound": "hbm", "kernel": "po2::conv_umma_kernel<9> (+pack_weights_kernel)", "unit": "GB/s",
            "achieved": tot_bytes / tot_us / 1e3, "peak": pk["hbm_gbs"], "frac": tot_bytes / tot_us / 1e3 / pk["hbm_gbs"],
            "traffic": None, "TFLOPs": tot_flop / tot_us / 1e6, "frac_of_bf16_peak": tot_flop / tot_us / 1e6 / pk["bf16_tflops"],
            "resnet56_3x3_forward_us": tot_us, "resnet56_3x3_forward_us_cudnn_tf32": tot_cudnn,
            "images_per_s_forward_qconv_only": batch / (tot_us * 1e-6), "layers": rows,
            "note": "52 stride-1 3x3 quantized convs of ResNet-56 at batch %d, cold L2, CUDA-graph timed" % batch}


# ------------------------------------------------------------------------------------------------
def cpu_training_step_fn(batch):
    """The same training step on host cores through the oracle's torch-CPU restatement of the
    reference (oracle/po2_oracle_torch.py): the one place bench.py executes oracle/."""
    from oracle.po2_oracle_torch import PO2, QuantizedConv2dOracle
    from workloads import resnet_cifar
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(8)
    model = resnet_cifar(56, 10, PO2, 4, conv_cls=QuantizedConv2dOracle).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    crit = nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(1000)
    x = torch.randn(batch, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (batch,), generator=g)

    def step():
        opt.zero_grad()  # reference train.py:81 (set_to_none=True: no fill / accumulate kernels)
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        return float(loss.item())
    return step


def cpu_baseline(batch, steps=2):
    step = cpu_training_step_fn(batch)
    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} full training steps (fwd+bwd+SGD) of ResNet-56 PO2 4-bit QAT at batch {batch}, "
                      f"oracle/po2_oracle_torch.py (stock ATen CPU ops == what the reference runs), 1 warm-up",
            "ms_per_step": dt / steps * 1e3}


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the step.  The reference is
    Python and cannot travel to the GPU box, so this is the oracle's op-for-op torch restatement
    (pinned bit-exactly to the reference by tests/test_oracle_golden.py) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    B = a.batch
    step = cpu_training_step_fn(B)
    t0 = time.perf_counter()
    step()
    probe = time.perf_counter() - t0
    sample = f"full batch {B}"
    if probe > 6.0:                      # slow host: keep the whole run within a few minutes
        B = 32
        step = cpu_training_step_fn(B)
        sample = f"reduced batch {B} (a batch-{a.batch} step took {probe:.1f} s on this host)"
    for _ in range(max(0, min(a.warmup, 3) - 1)):
        step()
    steps = max(1, min(a.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = B * steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": world,
           "steps": steps, "warmup": min(a.warmup, 3), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic (randn images 3x32x32, randint labels; kaiming-init weights, seed 8)",
           "config": {"workload": WORKLOAD, "batch_per_gpu": B, "device": "cpu"},
           "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": f"{steps} training steps, {sample}; oracle/po2_oracle_torch.py on torch CPU kernels"},
           "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-log2", type=int, default=30)
    ap.add_argument("--torch-profile", default=None, help="write a torch.profiler kernel table of eager steps and exit")
    ap.add_argument("--profile-step", action="store_true", help="run one eager step inside cudaProfilerStart/Stop and exit")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
