#!/usr/bin/env python
"""bench.py — headline benchmark of the one hot path: mini-batch forward / backprop / momentum update.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c1|c4] [--precision tf32|fp32]

One "step" = one mini-batch step (forward, output delta, delta backprop, weight gradients, momentum update) over one
batch of synthetic input.  N = 1 runs BASELINE.json config C3 (784-4096-4096-10, batch 8192); N > 1 runs the same
per-GPU batch on every rank (weak scaling) with the weight gradients all-reduced by NCCL inside the engine.
Rank 0 prints ONE JSON line.  `--impl reference` times the UNMODIFIED reference (its cuBLAS build, oracle/_ref) on the
same config instead.
"""
from __future__ import annotations

import argparse
import ctypes
import importlib.util
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _load_b2nn():
    spec = importlib.util.spec_from_file_location("b2nn_binding", os.path.join(ROOT, "cublas-ml_b200", "b2nn.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b2nn_binding"] = mod
    spec.loader.exec_module(mod)
    return mod


WORKLOADS = {
    # name: (layers, per-GPU batch, description)
    "c3": ([784, 4096, 4096, 10], 8192, "C3 synthetic wide MLP 784-4096-4096-10, batch 8192 per GPU"),
    "c1": ([784, 50, 10], 4200, "C1 MNIST-shaped classifier 784-50-10, batch 4200 (main.cpp:126-157), synthetic pixels"),
    "c4": ([784] + [8192] * 8 + [10], 8192, "C4 synthetic deep MLP 8x8192 hidden, batch 8192 per GPU"),
}


def step_flops(layers, B):
    """Algorithmic FLOPs of one step: forward + weight gradient = 2 * sum 2B(l_j+1)l_{j+1};
    delta backprop = sum_{j>=1} 2B l_j l_{j+1}."""
    fwd = sum(2.0 * B * (layers[j] + 1) * layers[j + 1] for j in range(len(layers) - 1))
    bwd = sum(2.0 * B * layers[j] * layers[j + 1] for j in range(1, len(layers) - 1))
    return 2 * fwd + bwd


def num_params(layers):
    return sum((layers[j] + 1) * layers[j + 1] for j in range(len(layers) - 1))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int, uuid: str | None):
        super().__init__(daemon=True)
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
                except Exception:
                    h = None
            self.h = h or pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self.stop_flag.wait(0.004)

    def summary(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "note": "NVML unavailable or no samples"}
        return {"sm_mhz": int(statistics.median(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "power_w_max": round(max(self.power), 1) if self.power else None}


def synth_batches(layers, B, nb, seed, rank):
    """nb batches of B samples, unit-major like the reference's column-major host layout: X (n+1, m) with the ones
    column FIRST, Y (K, m) one-hot; labels = argmax of a fixed random teacher so the loss can fall."""
    import torch
    n, K = layers[0], layers[-1]
    m = B * nb
    g = torch.Generator().manual_seed(seed * 1000 + rank)
    x = torch.empty(n + 1, m, dtype=torch.float32)
    x[0].fill_(1.0)
    x[1:].normal_(generator=g)
    teacher = torch.randn(K, n, generator=torch.Generator().manual_seed(2))
    lab = (teacher @ x[1:]).argmax(0)
    y = torch.zeros(K, m, dtype=torch.float32)
    y[lab, torch.arange(m)] = 1.0
    return x, y, lab


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(v, world):
    if world == 1:
        return v
    import torch
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_ours(args):
    import torch
    b2nn = _load_b2nn()
    rank, world, local = dist_setup(args.gpus)
    layers, B, desc_text = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    K, W = args.steps, max(args.warmup, 3)
    prec = {"tf32": b2nn.PREC_TF32, "fp32": b2nn.PREC_FP32, "tf32x3": b2nn.PREC_TF32X3}[args.precision]

    net = b2nn.Net(local)
    net.set_layers(layers, b2nn.INIT_2, 42)
    p2p_on = False
    if world > 1:
        import torch.distributed as dist
        uid = [b2nn.Net.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        net.comm_init(rank, world, uid[0])
        # gradient exchange of the large layers over NVLink peer memory, fused into the GEMM / update kernels
        if os.environ.get("B2NN_P2P", "1") != "0" and 2 <= world <= 8:
            ok = 1
            try:
                mine = net.comm_p2p_export()
            except Exception:  # noqa: BLE001
                mine, ok = b"", 0
            blobs = [None] * world
            dist.all_gather_object(blobs, mine)
            if ok and all(len(b) == 192 for b in blobs):
                try:
                    net.comm_p2p_open(b"".join(blobs))
                except Exception:  # noqa: BLE001
                    ok = 0
            else:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            p2p_on = bool(int(flag.item()))
            if not p2p_on and ok:
                # some rank could not map its peers (no CUDA IPC in this container): every rank stays on the NCCL path
                net.comm_p2p_open(None)

    nb = K
    iters = 1
    if K > 64:
        nb = max(d for d in range(1, 65) if K % d == 0)
        iters = K // nb
    x, y, _lab = synth_batches(layers, B, nb, seed=1, rank=rank)
    m = B * nb
    net.set_train_data(x.numpy().reshape(-1), y.numpy().reshape(-1), m)

    def desc(iters_, display_=1 << 30):
        return b2nn.make_desc(momentum=0.9, rate=0.01, iters=iters_, display=display_, batch_num=nb, classify=True,
                              act=b2nn.ACT_TANH, out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY, precision=prec,
                              skip_final_cost=True)

    # ---- warm-up (>= 3 steps), untimed -----------------------------------------------------------------------------
    wd = b2nn.make_desc(momentum=0.9, rate=0.01, iters=1, display=1 << 30, batch_num=nb, classify=True, act=b2nn.ACT_TANH,
                        out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY, precision=prec, skip_final_cost=True)
    done = 0
    while done < W:
        net.train(wd)
        done += nb

    # ---- timed region: exactly K steps, device-timed on the engine's stream (CUDA events inside b2nn_train) --------
    dev_uuid = None
    try:
        dev_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        pass
    sampler = ClockSampler(local, dev_uuid)
    barrier(world)
    sampler.start()
    out = net.train(desc(iters))
    barrier(world)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    assert out.steps == K, (out.steps, K)
    secs = max_over_ranks(out.loop_seconds, world)
    value = (B * world * K) / secs

    # ---- roofline leg: same K steps with per-launch CUDA events on the engine stream ---------------------------------
    sampler2 = ClockSampler(local, dev_uuid)
    sampler2.start()
    net.profile_enable(True)
    net.train(desc(iters))
    prof = net.profile_read()
    net.profile_enable(False)
    sampler2.stop_flag.set()
    sampler2.join(timeout=2)
    pk, pk_src = peaks()
    # MEASURED_PEAKS.json holds a burst and a sustained (power-capped) bf16 figure.  The leg is compared with the one
    # whose regime it ran in: sustained once the median SM clock of the leg has dropped below 0.85 x max, burst otherwise.
    c2 = sampler2.summary()
    # (conservative: a power-cap flag alone does not switch to the lower sustained figure; the median SM clock of the leg must
    # actually have dropped to the capped regime — MEASURED_PEAKS.json saw 0.66 x max under its sustained load)
    capped = bool(c2.get("sm_mhz") and c2.get("sm_max_mhz") and c2["sm_mhz"] < 0.85 * c2["sm_max_mhz"])
    bf16_key = "bf16_tflops_sustained" if capped else "bf16_tflops"
    bf16_peak = pk.get(bf16_key, pk.get("bf16_tflops", 1400.0))
    gemm_ms = sum(prof[k]["ms"] for k in ("forward", "backprop", "wgrad"))
    gemm_fl = sum(prof[k]["flops"] for k in ("forward", "backprop", "wgrad"))
    gemm_launches = sum(prof[k]["launches"] for k in ("forward", "backprop", "wgrad"))
    total_ms = sum(v["ms"] for v in prof.values())
    upd = prof["update"]
    if args.precision == "tf32":
        # MEASURED_PEAKS.json holds dense bf16; kind::tf32 runs at half the bf16 rate on the same tensor pipe.
        peak = bf16_peak / 2.0
        peak_note = f"0.5 x {bf16_key} ({pk_src}): tf32 is half-rate on the tcgen05 pipe"
    elif args.precision == "tf32x3":
        peak = bf16_peak / 6.0
        peak_note = f"{bf16_key} / 6 ({pk_src}): three tf32 MMAs per algorithmic multiply-add"
    else:
        peak = 75.0
        peak_note = "fp32 FFMA planning value (BASELINE.md); not measured"
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    try:        # DRAM bytes per tcgen05 launch from the committed ncu --set full capture of this workload (profiles/)
        if args.workload == "c3" and args.precision == "tf32" and B == 8192:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["gemm_tc_kernel"]["dram_bytes_per_launch_mean"]
    except Exception:
        traffic = None
    roofline = {
        "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 kind::tf32) x3 contractions" if args.precision != "fp32" else "gemm_simt_kernel",
        "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
        "traffic": traffic, "traffic_source": "profiles/r1_traffic.json (profiles/r1_step.csv: ncu dram__bytes_read+write per launch, mean over the 7 tcgen05 launches of a step)" if traffic else None,
        "algorithmic_flops_per_launch_mean": round(gemm_fl / max(gemm_launches, 1)), "peak_source": peak_note,
        "regime": "sustained (power-capped)" if capped else "burst (boost clocks)", "clocks": c2,
        "launches": gemm_launches, "avg_launch_ms": round(gemm_ms / max(gemm_launches, 1), 4),
        "share_of_step": round(gemm_ms / total_ms, 4) if total_ms > 0 else None,
        "by_kind": {k: {"ms_per_step": round(v["ms"] / K, 4), "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 and v["flops"] > 0 else None,
                        "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 and v["bytes"] > 0 else None,
                        "launches": v["launches"], "tc_launches": v["tc_launches"]} for k, v in prof.items()},
        "update_hbm": {"achieved_gbs": round(upd["bytes"] / (upd["ms"] * 1e-3) / 1e9, 1) if upd["ms"] > 0 else None,
                       "peak_gbs": pk.get("hbm_gbs"), "frac": round(upd["bytes"] / (upd["ms"] * 1e-3) / 1e9 / pk.get("hbm_gbs", 6650.0), 4) if upd["ms"] > 0 else None,
                       "bytes_per_param": 20},
        "algorithmic_flops_per_step": step_flops(layers, B),
    }

    # ---- end-to-end leg: host buffers through the C ABI.  b2nn_train_host streams every batch host -> device (pinned
    # source, double-buffered, copy of batch b+1 overlapping the compute of batch b) and reads every step's loss back.
    n1, Kout = layers[0] + 1, layers[-1]
    if True:
        xh = torch.empty(n1 * m, dtype=torch.float32).pin_memory()
        yh = torch.empty(Kout * m, dtype=torch.float32).pin_memory()
        xh.copy_(x.reshape(-1))
        yh.copy_(y.reshape(-1))
        sd = b2nn.make_desc(momentum=0.9, rate=0.01, iters=iters, display=1, batch_num=nb, classify=True, act=b2nn.ACT_TANH,
                            out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY, precision=prec, skip_final_cost=True)
        wdesc = b2nn.make_desc(momentum=0.9, rate=0.01, iters=1, display=1, batch_num=nb, classify=True, act=b2nn.ACT_TANH,
                               out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY, precision=prec, skip_final_cost=True)
        net.train_host(wdesc, ctypes.c_void_p(xh.data_ptr()), ctypes.c_void_p(yh.data_ptr()), m)      # warm-up (>= 3 steps)
        barrier(world)
        t0 = time.perf_counter()
        Jsteps, er = net.train_host(sd, ctypes.c_void_p(xh.data_ptr()), ctypes.c_void_p(yh.data_ptr()), m)
        barrier(world)
        e2e_secs = max_over_ranks(time.perf_counter() - t0, world)
        assert er.steps == K
        e2e = {"value": round(B * world * K / e2e_secs, 1), "unit": "samples/s",
               "h2d_bytes_per_step": int((layers[0] + Kout) * B * 4), "d2h_bytes_per_step": 8,
               "ms_per_step": round(e2e_secs / K * 1e3, 4), "last_loss": round(float(Jsteps[-1]), 6),
               "first_loss": round(float(Jsteps[0]), 6),
               "api": "b2nn_train_host: pinned host dataset in, every batch H2D (double-buffered, overlapped), every step's loss D2H"}
    else:
        # data parallel: per-step host entry point (H2D of the rank's batch slice, step, global loss back), every rank
        n_host = min(nb, 4)
        xh = [torch.empty(n1 * B, dtype=torch.float32).pin_memory() for _ in range(n_host)]
        yh = [torch.empty(Kout * B, dtype=torch.float32).pin_memory() for _ in range(n_host)]
        for i in range(n_host):
            xh[i].copy_(x[:, i * B:(i + 1) * B].reshape(-1))
            yh[i].copy_(y[:, i * B:(i + 1) * B].reshape(-1))
        sd = b2nn.make_desc(momentum=0.9, rate=0.01, iters=1, display=1, batch_num=1, classify=True, act=b2nn.ACT_TANH,
                            out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY, precision=prec, skip_final_cost=True)
        for i in range(3):
            net.train_step_host(sd, ctypes.c_void_p(xh[i % n_host].data_ptr()), ctypes.c_void_p(yh[i % n_host].data_ptr()), B, B * world)
        barrier(world)
        t0 = time.perf_counter()
        J = 0.0
        for i in range(K):
            J = net.train_step_host(sd, ctypes.c_void_p(xh[i % n_host].data_ptr()), ctypes.c_void_p(yh[i % n_host].data_ptr()), B, B * world)
        barrier(world)
        e2e_secs = max_over_ranks(time.perf_counter() - t0, world)
        e2e = {"value": round(B * world * K / e2e_secs, 1), "unit": "samples/s",
               "h2d_bytes_per_step": int((layers[0] + Kout) * B * 4), "d2h_bytes_per_step": 8,
               "ms_per_step": round(e2e_secs / K * 1e3, 4), "last_loss": round(J, 6),
               "api": "b2nn_train_step_host on every rank (pinned host batch in, gradients all-reduced, global loss out, every step)"}

    # ---- secondary: the fp32-grade tensor-core mode (3xTF32) on the same workload, same K steps ---------------------
    x3 = None
    if world == 1 and args.precision == "tf32":
        try:
            d3 = b2nn.make_desc(momentum=0.9, rate=0.01, iters=iters, display=1 << 30, batch_num=nb, classify=True,
                                act=b2nn.ACT_TANH, out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY,
                                precision=b2nn.PREC_TF32X3, skip_final_cost=True)
            net.train(d3)
            o3 = net.train(d3)
            x3 = {"value": round(B * K / o3.loop_seconds, 1), "unit": "samples/s", "ms_per_step": round(o3.loop_seconds / K * 1e3, 4),
                  "note": "fp32-grade: hi*hi + hi*lo + lo*hi tcgen05.mma per product, held to the fp32 parity tolerances in tests/"}
        except Exception as ex:  # noqa: BLE001
            x3 = {"value": None, "note": f"failed: {ex}"}

    line = None
    if rank == 0:
        # ---- CPU baseline: the oracle port on the host cores, bounded sample of the same workload ---------------------
        # rank 0 at N = 1 only (torchrun pins OMP_NUM_THREADS=1 for N > 1, which would misreport the host cores)
        cpu = cpu_baseline(layers, min(B, 2048 if args.workload != "c1" else B)) if world == 1 else \
            {"value": None, "unit": "samples/s", "cores": None, "kind": "port", "sample": "measured at N=1 only"}
        line = {
            "metric": "train_samples_per_sec", "value": round(value, 1), "unit": "samples/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": round(secs / K * 1e3, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": desc_text, "layers": layers, "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "precision_mode": args.precision,
                       "gradient_exchange": ("peer-memory: reduce-scatter fused in the wgrad GEMM epilogue, update+all-gather fused on the owner (NCCL for the 10-unit layer)" if p2p_on else ("NCCL all-reduce per layer, overlapped with backprop" if world > 1 else "none")),
                       "hidden": "tanh", "output": "softmax", "cost": "cross-entropy", "momentum": 0.9,
                       "l2": "every step reads a different batch; per-step working set (weights+velocity+gradients+activations) exceeds the 126 MB L2"},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(out.kernel_launches),
            "gpu_launches_tcgen05": int(out.gemm_tc_launches),
            "tflops": round(step_flops(layers, B) * K * world / secs / 1e12, 2),
            "roofline": roofline, "cpu_baseline": cpu, "modes": {"tf32x3": x3},
        }
        print(json.dumps(line), flush=True)
    net.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return line


def cpu_baseline(layers, sample_B, target_seconds=12.0):
    """The oracle port (oracle/mlp_oracle.cpp, fp32 + OpenMP) timed on this box's host cores on a BOUNDED sample of the
    same workload: momentum steps at batch `sample_B` of the same network — one step to calibrate, then as many as fit in
    about `target_seconds` (at most 16) — reported as samples/s."""
    try:
        from oracle import oracle as O
        n, K = layers[0], layers[-1]
        rng = np.random.default_rng(1)
        th0 = O.init_weights(layers, 2, 42)

        def run(steps):
            m = sample_B * steps
            x = rng.standard_normal((m, n)).astype(np.float32)
            lab = rng.integers(0, K, m)
            x_cm, y_cm = O.with_bias_colmajor(x), O.one_hot_colmajor(lab, K)
            return O.train(layers, x_cm, y_cm, m, batch_num=steps, iters=1, display=1 << 30, momentum=0.9, rate=0.01,
                           hidden_act=O.TANH, classify=True, out_act=O.OUT_SOFTMAX, cost=O.COST_CROSSENTROPY, theta0=th0,
                           precision="f32").seconds
        t1 = run(1)
        steps = int(max(1, min(16, target_seconds // max(t1, 1e-3))))
        secs = run(steps) if steps > 1 else t1
        return {"value": round(sample_B * steps / secs, 1), "unit": "samples/s", "cores": int(O.lib().oracle_num_threads()),
                "kind": "port", "sample": f"{steps} momentum steps at batch {sample_B} of the same network (loop time only)",
                "seconds": round(secs, 3)}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "unit": "samples/s", "cores": None, "kind": "port", "sample": f"failed: {ex}"}


def run_reference(args):
    """The UNMODIFIED reference (cuBLAS Sgemm + its own element-wise kernels) on the same config and GPU.  The class
    only reports the time of a whole train*Momentum call (setup + loop + final cost, cublasNN.cpp:586-639), so the
    per-step time is the difference of two calls that differ only in the number of epochs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from oracle import oracle as O
    layers, B, desc_text = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    K, W = args.steps, max(args.warmup, 3)
    if not O.ref_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcublasml_ref.so was not built (needs /root/reference)"}))
        return None
    nb = 4 if K % 4 == 0 else (2 if K % 2 == 0 else 1)
    it_timed = K // nb
    it_warm = max(1, (W + nb - 1) // nb)
    x, y, lab = synth_batches(layers, B, nb, seed=1, rank=0)
    x_rows = np.ascontiguousarray(x[1:].numpy().T)
    y_rows = lab.numpy().astype(np.float32).reshape(-1, 1)
    kw = dict(batch_num=nb, display=1 << 30, momentum=0.9, rate=0.01, hidden_act=O.TANH, classify=True,
              out_act=O.OUT_SOFTMAX, cost=O.COST_CROSSENTROPY, init_kind=2, seed=42)
    env_note = "NVIDIA_TF32_OVERRIDE=" + os.environ.get("NVIDIA_TF32_OVERRIDE", "unset")
    O.ref_train(layers, x_rows, y_rows, iters=it_warm, **kw)                       # warm-up call (also warms cuBLAS)
    t_short = O.ref_train(layers, x_rows, y_rows, iters=it_warm, **kw).seconds
    t_long = O.ref_train(layers, x_rows, y_rows, iters=it_warm + it_timed, **kw).seconds
    secs = max(t_long - t_short, 1e-9)
    value = B * K / secs
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": round(value, 1), "unit": "samples/s",
            "n_gpus": int(getattr(args, "gpus", 1) or 1), "gpus_used": 1, "steps": K, "warmup": it_warm * nb,
            "ms_per_step": round(secs / K * 1e3, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32" if os.environ.get("NVIDIA_TF32_OVERRIDE", "0") != "1" else "tf32(cuBLAS override)",
            "data": "synthetic",
            "config": {"workload": desc_text, "layers": layers, "batch_per_gpu": B, "global_batch": B,
                       "parallelism": "dp1 (the reference is single-GPU: it runs on cuda:0 whatever --gpus says, see gpus_used)", "math": env_note,
                       "timing": "difference of two trainClassifyMomentum calls (epochs E and E+K/nb): setup and final cost cancel"},
            "tflops": round(step_flops(layers, B) * K / secs / 1e12, 2),
            "cpu_baseline": {"value": round(value, 1), "unit": "samples/s", "cores": 0, "kind": "reference",
                             "sample": "unmodified reference built from /root/reference (oracle/_ref), its cuBLAS path on cuda:0; the reference has no CPU path"},
            "e2e": {"value": round(value, 1), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32", "tf32x3"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
