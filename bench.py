#!/usr/bin/env python
"""bench.py — headline benchmark of the one hot path: mini-batch forward / backprop / momentum update.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c1|c4]
                    [--precision tf32|fp32|tf32x3] [--no-secondary]

One "step" = one mini-batch step (forward, output delta, delta backprop, weight gradients, momentum update) over one
batch of synthetic input.  N = 1 runs BASELINE.json config C3 (784-4096-4096-10, batch 8192); N > 1 runs the same
per-GPU batch on every rank (weak scaling) with the weight gradients exchanged inside the engine.  Rank 0 prints ONE
JSON line.  Besides the headline the line carries `secondary`: the other four BASELINE.json configs (C1 MNIST net, C2
bikeshare, C4 deep net, C5 inference sweep) and the fp32-grade mode of C3, each with its own clocks and roofline
fraction; for N > 1, C4 at global batch 8192 x N and a replica-identity check.  `--impl reference` times the UNMODIFIED
reference (its cuBLAS build, oracle/_ref) on the same configs instead.
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
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


def _load_cases():
    """tests/cases.py holds the seeded input generators shared with the parity tests (it reads only tests/golden)."""
    spec = importlib.util.spec_from_file_location("b2nn_cases", os.path.join(ROOT, "tests", "cases.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b2nn_cases"] = mod
    spec.loader.exec_module(mod)
    return mod


WORKLOADS = {
    # name: (layers, per-GPU batch, description)
    "c3": ([784, 4096, 4096, 10], 8192, "C3 synthetic wide MLP 784-4096-4096-10, batch 8192 per GPU"),
    "c1": ([784, 50, 10], 4200, "C1 MNIST-shaped classifier 784-50-10, batch 4200 (main.cpp:126-157), synthetic pixels"),
    "c4": ([784] + [8192] * 8 + [10], 8192, "C4 synthetic deep MLP 8x8192 hidden, batch 8192 per GPU"),
}
C2_LAYERS = [10, 40, 160, 1]


def step_flops(layers, B):
    """Algorithmic FLOPs of one step: forward + weight gradient = 2 * sum 2B(l_j+1)l_{j+1};
    delta backprop = sum_{j>=1} 2B l_j l_{j+1}."""
    fwd = sum(2.0 * B * (layers[j] + 1) * layers[j + 1] for j in range(len(layers) - 1))
    bwd = sum(2.0 * B * layers[j] * layers[j + 1] for j in range(1, len(layers) - 1))
    return 2 * fwd + bwd


def num_params(layers):
    return sum((layers[j] + 1) * layers[j + 1] for j in range(len(layers) - 1))


def step_bytes_small(layers, B):
    """Algorithmic HBM bytes of one step of a small net: the X batch once, Y once, 20 B per weight."""
    return 4.0 * B * (layers[0] + 1) + 4.0 * B * layers[-1] + 20.0 * num_params(layers)


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


class Clocks:
    """`with Clocks(dev) as c: ...; c.summary` — clocks sampled over exactly the enclosed region."""

    def __init__(self, local, uuid):
        self.s = ClockSampler(local, uuid)
        self.summary = None

    def __enter__(self):
        self.s.start()
        return self

    def __exit__(self, *exc):
        self.s.stop_flag.set()
        self.s.join(timeout=2)
        self.summary = self.s.summary()
        return False


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


def yardsticks():
    """cuBLAS TF32 / FP32 8192^3 throughput measured once on this pool by tools/measure_peaks.py (yardstick only: the
    roofline denominators stay MEASURED_PEAKS.json's)."""
    p = os.path.join(ROOT, "profiles", "r2_peaks.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def tensor_peak(precision, clocks, pk, pk_src):
    """Roofline denominator of a tensor-core leg.  MEASURED_PEAKS.json holds a burst and a sustained (power-capped) bf16
    figure; the leg is compared with the one whose regime it ran in: sustained once its median SM clock dropped below
    0.85 x max, burst otherwise.  kind::tf32 runs at half the bf16 rate on the same pipe; 3xTF32 issues three MMAs."""
    capped = bool(clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] < 0.85 * clocks["sm_max_mhz"])
    key = "bf16_tflops_sustained" if capped else "bf16_tflops"
    bf16 = pk.get(key, pk.get("bf16_tflops", 1400.0))
    if precision == "tf32":
        return bf16 / 2.0, f"0.5 x {key} ({pk_src}): tf32 is half-rate on the tcgen05 pipe", capped
    if precision == "tf32x3":
        return bf16 / 6.0, f"{key} / 6 ({pk_src}): three tf32 MMAs per algorithmic multiply-add", capped
    if precision == "bf16":
        return bf16, f"{key} ({pk_src}): kind::f16 with bf16 inputs runs at the measured cuBLAS bf16 rate", capped
    y = yardsticks()
    if y and y.get("fp32_ffma_tflops"):
        return float(y["fp32_ffma_tflops"]), "cuBLAS SGEMM 8192^3 measured by tools/measure_peaks.py (profiles/r2_peaks.json)", capped
    return 75.0, "fp32 FFMA planning value (BASELINE.md); not measured", capped


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


def make_net(b2nn, layers, local, rank, world, init=None, seed=42):
    """A context with its communicator and, when the devices allow it, the peer-memory gradient exchange opened."""
    import torch
    net = b2nn.Net(local)
    net.set_layers(layers, b2nn.INIT_2 if init is None else init, seed)
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
    return net, p2p_on


def replicas_identical(net, world):
    """Every rank must hold bit-identical weights after data-parallel training: compare a SHA-256 of the weight arena."""
    th = net.get_weights()
    digest = hashlib.sha256(th.tobytes()).digest()
    if world == 1:
        return True, th
    import torch
    import torch.distributed as dist
    mine = torch.tensor(list(digest), dtype=torch.uint8, device="cuda")
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    return all(bool(torch.equal(allv[0], v)) for v in allv[1:]), th


def classify_desc(b2nn, prec, nb, iters, display=1 << 30, rate=0.01):
    return b2nn.make_desc(momentum=0.9, rate=rate, iters=iters, display=display, batch_num=nb, classify=True,
                          act=b2nn.ACT_TANH, out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY, precision=prec,
                          skip_final_cost=True)


def traffic_record(b2nn_dir):
    """DRAM bytes per tcgen05 launch from the committed ncu --set full capture (profiles/), valid only while the kernel
    source is the one that was profiled: the capture records the SHA-256 of csrc/gemm_tc.cu."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            rec = json.load(open(p))
        except Exception:
            continue
        src = os.path.join(b2nn_dir, "csrc", "gemm_tc.cu")
        sha = hashlib.sha256(open(src, "rb").read()).hexdigest()[:16] if os.path.exists(src) else None
        if rec.get("gemm_tc_cu_sha16") and rec["gemm_tc_cu_sha16"] == sha:
            return rec["gemm_tc_kernel"]["dram_bytes_per_launch_mean"], {"profiled": f"profiles/{name}@{rec.get('commit', '?')}",
                                                                        "gemm_tc_cu_sha16": sha}
        return None, {"profiled": None, "note": f"profiles/{name} was captured on a different gemm_tc.cu; not reported"}
    return None, {"profiled": None}


# ---- secondary configs ---------------------------------------------------------------------------------------------------

def sec_c1(b2nn, local, uuid, pk):
    """C1: MNIST-shaped 784-50-10, m = 42 000 (the repo's real labels, synthetic pixels), 10 batches of 4 200, tanh /
    softmax / cross-entropy, mu 0.9, rate 0.075 (main.cpp:126-157).  HBM roofline: 14.15 MB per step."""
    cases = _load_cases()
    layers, m, nb = [784, 50, 10], 42000, 10
    lab = cases.mnist_labels()
    xr = cases.zscore(cases.synth_digits(lab, 784, seed=1234))
    x = np.ascontiguousarray(np.concatenate([np.ones((m, 1), np.float32), xr], axis=1).T).reshape(-1)     # column-major, ones first
    y = np.zeros((10, m), np.float32)
    y[lab, np.arange(m)] = 1.0
    y = y.reshape(-1)
    del xr
    out = {}
    net = b2nn.Net(local)
    net.set_layers(layers, b2nn.INIT_2, 42)
    net.set_train_data(x, y, m)
    for name, prec in (("default_fp32grade", b2nn.PREC_AUTO), ("tf32", b2nn.PREC_TF32), ("fp32", b2nn.PREC_FP32)):
        net.train(classify_desc(b2nn, prec, nb, 3, rate=0.075))
        with Clocks(local, uuid) as ck:
            r = net.train(classify_desc(b2nn, prec, nb, 40, rate=0.075))
        us = r.loop_seconds / r.steps * 1e6
        by = step_bytes_small(layers, m // nb)
        out[name] = {"us_per_step": round(us, 2), "samples_per_s": round(m / nb / us * 1e6), "steps": int(r.steps),
                     "launches_per_step": round(r.kernel_launches / r.steps, 2), "clocks": ck.summary,
                     "roofline": {"bound": "hbm", "achieved": round(by / us / 1e3, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                                  "frac": round(by / us / 1e3 / pk["hbm_gbs"], 4), "algorithmic_bytes_per_step": int(by)}}
    net.close()
    return {"workload": WORKLOADS["c1"][2], "batch": m // nb, **out}


def sec_c2(b2nn, local, uuid, pk):
    """C2: bikeshare regression 10-40-160-1, m = 4337, full batch, sigmoid, mu 0.9, rate 0.003 (main.cpp:54-96), the
    committed fixture of the repo's CSVs.  Latency bound (0.2 MB per step): us/step is the number."""
    cases = _load_cases()
    b = cases.bikeshare()
    m = b["x"].shape[0]
    xn = cases.zscore(b["x"])
    x = np.ascontiguousarray(np.concatenate([np.ones((m, 1), np.float32), xn], axis=1).T).reshape(-1)
    y = np.ascontiguousarray(b["y"].astype(np.float32).T).reshape(-1)
    net = b2nn.Net(local)
    net.set_layers(C2_LAYERS, b2nn.INIT_1, 42)
    net.set_train_data(x, y, m)

    def d(iters):
        return b2nn.make_desc(momentum=0.9, rate=0.003, iters=iters, display=1 << 30, batch_num=1, classify=False,
                              act=b2nn.ACT_SIGMOID, out=b2nn.OUT_LINEAR, cost=b2nn.COST_SQUARED, precision=b2nn.PREC_AUTO,
                              skip_final_cost=True)
    net.train(d(50))
    with Clocks(local, uuid) as ck:
        r = net.train(d(3000))
    net.close()
    us = r.loop_seconds / r.steps * 1e6
    by = step_bytes_small(C2_LAYERS, m)
    return {"workload": "C2 bikeshare regression 10-40-160-1, m 4337, full batch (main.cpp:54-96)", "us_per_step": round(us, 2),
            "samples_per_s": round(m / us * 1e6), "steps": int(r.steps), "launches_per_step": round(r.kernel_launches / r.steps, 2),
            "precision": "default (every contraction below the tensor-core threshold: exact fp32)", "clocks": ck.summary,
            "roofline": {"bound": "hbm", "achieved": round(by / us / 1e3, 2), "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": round(by / us / 1e3 / pk["hbm_gbs"], 5), "algorithmic_bytes_per_step": int(by),
                         "note": "latency bound: 0.2 MB per step; the reference issues 23 launches per step, this engine 2"}}


def sec_c4(b2nn, local, rank, world, uuid, pk, pk_src, steps=4):
    """C4: 784 - 8 x 8192 - 10, 8192 rows per GPU (global batch 8192 x N: 65 536 at 8 GPUs, north_star's config), tf32."""
    layers, B, _ = WORKLOADS["c4"]
    nb = 2
    iters = max(1, steps // nb)
    net, p2p_on = make_net(b2nn, layers, local, rank, world)
    x, y, _ = synth_batches(layers, B, nb, seed=3, rank=rank)
    net.set_train_data(x.numpy().reshape(-1), y.numpy().reshape(-1), B * nb)
    del x, y
    net.train(classify_desc(b2nn, b2nn.PREC_TF32, nb, 2))            # 4 warm-up steps
    barrier(world)
    with Clocks(local, uuid) as ck:
        r = net.train(classify_desc(b2nn, b2nn.PREC_TF32, nb, iters))
        barrier(world)
    secs = max_over_ranks(r.loop_seconds, world)
    same, _ = replicas_identical(net, world)
    net.close()
    fl = step_flops(layers, B) * world
    peak, note, capped = tensor_peak("tf32", ck.summary, pk, pk_src)
    burst = pk.get("bf16_tflops", 1590.0) / 2.0
    tf = fl * r.steps / secs / 1e12
    return {"workload": WORKLOADS["c4"][2], "global_batch": B * world, "n_gpus": world, "ms_per_step": round(secs / r.steps * 1e3, 3),
            "samples_per_s": round(B * world * r.steps / secs), "steps": int(r.steps), "tflops": round(tf, 1),
            "precision": "tf32", "clocks": ck.summary, "replicas_identical": bool(same),
            "gradient_exchange": "peer-memory (fused)" if p2p_on else ("NCCL" if world > 1 else "none"),
            "roofline": {"bound": "tensor", "achieved": round(tf / world, 1), "peak": round(peak, 1), "unit": "TFLOP/s per GPU",
                         "frac": round(tf / world / peak, 4), "frac_of_burst": round(tf / world / burst, 4), "peak_source": note,
                         "regime": "sustained (power-capped)" if capped else "burst (boost clocks)",
                         "algorithmic_flops_per_step": fl}}


def sec_c5(b2nn, local, uuid, pk):
    """C5: forward-only sweep on the C1 net.  Device-resident (b2nn_predict_device, CUDA events on the engine stream) and
    end to end through b2nn_predict from pinned host memory (H2D + forward + D2H of the class ids, chunks overlapped)."""
    import torch
    layers = [784, 50, 10]
    net = b2nn.Net(local)
    net.set_layers(layers, b2nn.INIT_2, 42)
    stream = torch.cuda.ExternalStream(net.stream())
    res = {"workload": "C5 forward-only inference on the 784-50-10 net, batch 1 .. 2^20", "resident": {}, "e2e": {}}
    for name, prec in (("default_fp32grade", b2nn.PREC_AUTO), ("tf32", b2nn.PREC_TF32)):
        d = b2nn.make_desc(classify=True, act=b2nn.ACT_TANH, out=b2nn.OUT_SOFTMAX, precision=prec)
        rows = {}
        with Clocks(local, uuid) as ck:
            for p in (0, 6, 12, 16, 18, 20):
                Bn = 1 << p
                ld = (Bn + 31) // 32 * 32
                xd = torch.randn(785, ld, device="cuda")
                xd[784] = 1.0
                outd = torch.empty(ld, device="cuda")
                for _ in range(3):
                    net.predict_device(d, xd.data_ptr(), Bn, ld, outd.data_ptr())
                reps = 20 if Bn <= 65536 else 5
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(reps):
                    net.predict_device(d, xd.data_ptr(), Bn, ld, outd.data_ptr())
                e1.record(stream)
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / reps * 1e3
                gbs = Bn * 3144.0 / us / 1e3
                rows[str(Bn)] = {"us": round(us, 2), "samples_per_s": round(Bn / us * 1e6), "hbm_gbs": round(gbs, 1),
                                 "frac_of_hbm_peak": round(gbs / pk["hbm_gbs"], 4)}
                del xd, outd
        best = max(rows.values(), key=lambda v: v["hbm_gbs"])
        res["resident"][name] = {"by_batch": rows, "clocks": ck.summary,
                                 "roofline": {"bound": "hbm", "achieved": best["hbm_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                                              "frac": best["frac_of_hbm_peak"], "algorithmic_bytes_per_sample": 3144}}
    d = b2nn.make_desc(classify=True, act=b2nn.ACT_TANH, out=b2nn.OUT_SOFTMAX, precision=b2nn.PREC_AUTO)
    with Clocks(local, uuid) as ck:
        for Bn in (1, 4096, 1 << 18):
            xh = torch.empty(785 * Bn, dtype=torch.float32).pin_memory()
            xh.normal_()
            xh[:Bn] = 1.0
            oh = torch.empty(Bn, dtype=torch.float32).pin_memory()
            xp, op = ctypes.c_void_p(xh.data_ptr()), ctypes.c_void_p(oh.data_ptr())
            for _ in range(2):
                net.predict_ptr(d, xp, Bn, op)
            reps = 20 if Bn <= 4096 else 3
            t0 = time.perf_counter()
            for _ in range(reps):
                net.predict_ptr(d, xp, Bn, op)
            dt = (time.perf_counter() - t0) / reps
            res["e2e"][str(Bn)] = {"us": round(dt * 1e6, 1), "samples_per_s": round(Bn / dt), "h2d_gbs": round(Bn * 3136.0 / dt / 1e9, 2),
                                   "h2d_bytes": Bn * 3136, "d2h_bytes": Bn * 4}
            del xh, oh
    res["e2e"]["clocks"] = ck.summary
    res["e2e"]["api"] = "b2nn_predict: pinned host rows in, class ids out; rows travel in chunks, H2D / compute / D2H overlapped"
    net.close()
    return res


def run_ours(args):
    import torch
    b2nn = _load_b2nn()
    rank, world, local = dist_setup(args.gpus)
    layers, B, desc_text = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    K, W = args.steps, max(args.warmup, 3)
    prec = {"tf32": b2nn.PREC_TF32, "fp32": b2nn.PREC_FP32, "tf32x3": b2nn.PREC_TF32X3, "bf16": b2nn.PREC_BF16}[args.precision]
    pk, pk_src = peaks()
    dev_uuid = None
    try:
        dev_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        pass

    net, p2p_on = make_net(b2nn, layers, local, rank, world)

    nb = K
    iters = 1
    if K > 64:
        nb = max(d for d in range(1, 65) if K % d == 0)
        iters = K // nb
    x, y, _lab = synth_batches(layers, B, nb, seed=1, rank=rank)
    m = B * nb
    net.set_train_data(x.numpy().reshape(-1), y.numpy().reshape(-1), m)

    # ---- warm-up (>= 3 steps), untimed -----------------------------------------------------------------------------
    done = 0
    while done < W:
        net.train(classify_desc(b2nn, prec, nb, 1))
        done += nb

    # ---- timed region: exactly K steps, device-timed on the engine's stream (CUDA events inside b2nn_train) --------
    barrier(world)
    with Clocks(local, dev_uuid) as ck_main:
        out = net.train(classify_desc(b2nn, prec, nb, iters))
        barrier(world)
    assert out.steps == K, (out.steps, K)
    secs = max_over_ranks(out.loop_seconds, world)
    value = (B * world * K) / secs
    same_replicas, theta_after = replicas_identical(net, world)

    # ---- roofline leg: same K steps with per-launch CUDA events on the engine stream ---------------------------------
    with Clocks(local, dev_uuid) as ck_roof:
        net.profile_enable(True)
        net.train(classify_desc(b2nn, prec, nb, iters))
        prof = net.profile_read()
        net.profile_enable(False)
    c2 = ck_roof.summary
    # Headline denominator: ALWAYS the burst figure (the timed region is tens of milliseconds; NVML's 4 ms samples of it
    # flip between regimes from run to run).  The sustained figure is reported beside it (`frac_of_sustained`).
    _, _, capped = tensor_peak(args.precision, c2, pk, pk_src)
    peak, peak_note, _ = tensor_peak(args.precision, {"sm_mhz": None}, pk, pk_src)
    peak_sus, _, _ = tensor_peak(args.precision, {"sm_mhz": 1, "sm_max_mhz": 1000}, pk, pk_src)
    gemm_ms = sum(prof[k]["ms"] for k in ("forward", "backprop", "wgrad"))
    gemm_fl = sum(prof[k]["flops"] for k in ("forward", "backprop", "wgrad"))
    gemm_launches = sum(prof[k]["launches"] for k in ("forward", "backprop", "wgrad"))
    total_ms = sum(v["ms"] for v in prof.values())
    upd = prof["update"]
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic, traffic_src = (None, {"profiled": None})
    if args.workload == "c3" and args.precision == "tf32" and B == 8192:
        traffic, traffic_src = traffic_record(os.path.join(ROOT, "cublas-ml_b200"))
    yard = yardsticks()
    roofline = {
        "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 kind::tf32) x3 contractions" if args.precision != "fp32" else "gemm_simt_kernel",
        "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
        "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_flops_per_launch_mean": round(gemm_fl / max(gemm_launches, 1)), "peak_source": peak_note,
        "frac_of_sustained": round(achieved / peak_sus, 4), "peak_sustained": round(peak_sus, 1),
        "regime_seen": "sustained (power-capped)" if capped else "burst (boost clocks)", "clocks": c2,
        "launches": gemm_launches, "avg_launch_ms": round(gemm_ms / max(gemm_launches, 1), 4),
        "share_of_step": round(gemm_ms / total_ms, 4) if total_ms > 0 else None,
        "whole_step": {"tflops": round(step_flops(layers, B) * K / secs / 1e12, 2),
                       "frac": round(step_flops(layers, B) * K / secs / 1e12 / peak, 4),
                       "note": "algorithmic FLOPs of the step / device-timed step (per GPU), every launch included"},
        "by_kind": {k: {"ms_per_step": round(v["ms"] / K, 4), "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 and v["flops"] > 0 else None,
                        "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 and v["bytes"] > 0 else None,
                        "launches": v["launches"], "tc_launches": v["tc_launches"]} for k, v in prof.items()},
        "update_hbm": {"achieved_gbs": round(upd["bytes"] / (upd["ms"] * 1e-3) / 1e9, 1) if upd["ms"] > 0 else None,
                       "peak_gbs": pk.get("hbm_gbs"), "frac": round(upd["bytes"] / (upd["ms"] * 1e-3) / 1e9 / pk.get("hbm_gbs", 6650.0), 4) if upd["ms"] > 0 else None,
                       "bytes_per_param": 20},
        "algorithmic_flops_per_step": step_flops(layers, B),
        "cublas_yardstick": yard,
    }

    # ---- end-to-end leg: host buffers through the C ABI.  b2nn_train_host streams every batch host -> device (pinned
    # source, double-buffered, copy of batch b+1 overlapping the compute of batch b) and reads every step's loss back.
    n1, Kout = layers[0] + 1, layers[-1]
    xh = torch.empty(n1 * m, dtype=torch.float32).pin_memory()
    yh = torch.empty(Kout * m, dtype=torch.float32).pin_memory()
    xh.copy_(x.reshape(-1))
    yh.copy_(y.reshape(-1))
    xp, yp = ctypes.c_void_p(xh.data_ptr()), ctypes.c_void_p(yh.data_ptr())

    def e2e_leg(precision_code):
        sd = classify_desc(b2nn, precision_code, nb, iters, display=1)
        net.train_host(classify_desc(b2nn, precision_code, nb, 1, display=1), xp, yp, m)      # warm-up (>= 3 steps)
        barrier(world)
        t0 = time.perf_counter()
        Jsteps, er = net.train_host(sd, xp, yp, m)
        barrier(world)
        e2e_secs = max_over_ranks(time.perf_counter() - t0, world)
        assert er.steps == K
        return {"value": round(B * world * K / e2e_secs, 1), "unit": "samples/s",
                "h2d_bytes_per_step": int((layers[0] + Kout) * B * 4), "d2h_bytes_per_step": 8,
                "ms_per_step": round(e2e_secs / K * 1e3, 4),
                "device_loop_ms_per_step": round(max_over_ranks(er.loop_seconds, world) / K * 1e3, 4), "last_loss": round(float(Jsteps[-1]), 6),
                "first_loss": round(float(Jsteps[0]), 6),
                "api": "b2nn_train_host: pinned host dataset in, every batch H2D (double-buffered, overlapped), every step's loss D2H"}
    e2e = e2e_leg(prec)

    secondary = {}
    # ---- secondary: the other tensor-core modes on the same workload, same K steps, each with its own e2e ------------
    if args.precision == "tf32" and not args.no_secondary:
        for key, code, pname, note in (
                ("c3_tf32x3", b2nn.PREC_TF32X3, "tf32x3", "fp32-grade (the engine's DEFAULT precision): hi*hi + hi*lo + lo*hi tcgen05.mma per product, held to the fp32 parity tolerances in tests/"),
                ("c3_bf16", b2nn.PREC_BF16, "bf16", "opt-in: tcgen05 kind::f16 on bf16 copies of the operands (fp32 accumulate, fp32 master weights / activations / deltas); stated bound 6e-2 on the weights (tests/)")):
            try:
                net.train(classify_desc(b2nn, code, nb, 1))
                barrier(world)
                with Clocks(local, dev_uuid) as ck3:
                    o3 = net.train(classify_desc(b2nn, code, nb, iters))
                    barrier(world)
                s3 = max_over_ranks(o3.loop_seconds, world)
                p3, n3, cap3 = tensor_peak(pname, ck3.summary, pk, pk_src)
                burst3 = pk.get("bf16_tflops", 1590.0) / {"tf32x3": 6.0, "bf16": 1.0}[pname]
                tf3 = step_flops(layers, B) * K / s3 / 1e12
                secondary[key] = {
                    "value": round(B * world * K / s3, 1), "unit": "samples/s", "ms_per_step": round(s3 / K * 1e3, 4), "clocks": ck3.summary,
                    "roofline": {"bound": "tensor", "achieved": round(tf3, 2), "peak": round(p3, 1), "unit": "TFLOP/s", "frac": round(tf3 / p3, 4),
                                 "peak_source": n3, "regime": "sustained (power-capped)" if cap3 else "burst (boost clocks)",
                                 "frac_of_burst": round(tf3 / burst3, 4),
                                 "note": "whole step (every launch) over the mode's tensor peak; the sustained figure of MEASURED_PEAKS.json was taken at 1297 MHz, a leg that ran between the two regimes can read above 1.0 of it"},
                    "e2e": e2e_leg(code), "note": note}
            except Exception as ex:  # noqa: BLE001
                secondary[key] = {"value": None, "note": f"failed: {ex}"}

    # ---- data parallel: the NCCL path must land on the same weights as the peer-memory path ---------------------------
    dp_check = None
    if world > 1:
        dp_check = {"replicas_identical": bool(same_replicas), "gradient_exchange": "peer-memory" if p2p_on else "NCCL"}
        if p2p_on:
            try:
                d_chk = classify_desc(b2nn, prec, nb, 1)
                net.set_weights(theta_after)
                net.train(d_chk)
                same_p, th_p = replicas_identical(net, world)
                net.set_weights(theta_after)
                net.comm_p2p_open(None)                   # same steps from the same weights over NCCL all-reduce
                net.train(d_chk)
                same_n, th_n = replicas_identical(net, world)
                dp_check.update({"extra_steps": nb, "replicas_identical_after_extra_steps": bool(same_p and same_n),
                                 "p2p_vs_nccl_theta_max_rel": float(np.abs(th_p - th_n).max() / max(np.abs(th_n).max(), 1e-30))})
            except Exception as ex:  # noqa: BLE001
                dp_check["p2p_vs_nccl"] = f"failed: {ex}"
    launches_headline, tc_headline = int(out.kernel_launches), int(out.gemm_tc_launches)
    net.close()
    del xh, yh, x, y

    # ---- secondary: the other configs (every rank runs C4; rank-0-only configs at N = 1) ------------------------------
    if not args.no_secondary and args.workload == "c3":
        try:
            secondary["c4"] = sec_c4(b2nn, local, rank, world, dev_uuid, pk, pk_src)
        except Exception as ex:  # noqa: BLE001
            secondary["c4"] = {"failed": str(ex)}
        if world == 1:
            for name, fn in (("c1", sec_c1), ("c2", sec_c2), ("c5", sec_c5)):
                try:
                    secondary[name] = fn(b2nn, local, dev_uuid, pk)
                except Exception as ex:  # noqa: BLE001
                    secondary[name] = {"failed": str(ex)}

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
                       "gradient_exchange": ("peer-memory: reduce-scatter fused in the wgrad GEMM epilogue, update+all-gather fused on the owner, slot buffers for the 10-unit layer" if p2p_on else ("NCCL all-reduce per layer, overlapped with backprop" if world > 1 else "none")),
                       "hidden": "tanh", "output": "softmax", "cost": "cross-entropy", "momentum": 0.9,
                       "l2": "every step reads a different batch; per-step working set (weights+velocity+gradients+activations) exceeds the 126 MB L2"},
            "clocks": ck_main.summary, "e2e": e2e, "gpu_launches": launches_headline,
            "gpu_launches_tcgen05": tc_headline,
            "tflops": round(step_flops(layers, B) * K * world / secs / 1e12, 2),
            "roofline": roofline, "cpu_baseline": cpu, "secondary": secondary,
        }
        if dp_check is not None:
            line["replicas_identical"] = dp_check["replicas_identical"]
            line["dp_check"] = dp_check
        print(json.dumps(line), flush=True)
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


# ---- reference arm --------------------------------------------------------------------------------------------------------

def _ref_step_seconds(O, layers, x_rows, y_rows, nb, it_short, it_long, repeats=3, **kw):
    """The class only reports the time of a whole train*Momentum call (allocVarGPU + splitData + loop + calcFinalCost,
    cublasNN.cpp:586-639): per-step time = difference of two calls that differ only in the number of epochs.  Setup time
    jitters by more than a short loop lasts, so each of the two is the MINIMUM over `repeats` calls (positive noise only)."""
    O.ref_train(layers, x_rows, y_rows, iters=it_short, batch_num=nb, **kw)                    # warm-up call (also warms cuBLAS)
    t_short = min(O.ref_train(layers, x_rows, y_rows, iters=it_short, batch_num=nb, **kw).seconds for _ in range(repeats))
    t_long = min(O.ref_train(layers, x_rows, y_rows, iters=it_long, batch_num=nb, **kw).seconds for _ in range(repeats))
    return max(t_long - t_short, 1e-9) / ((it_long - it_short) * nb)


def run_reference(args):
    """The UNMODIFIED reference (cuBLAS Sgemm + its own element-wise kernels) on the same configs and GPU."""
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
    kw = dict(display=1 << 30, momentum=0.9, rate=0.01, hidden_act=O.TANH, classify=True,
              out_act=O.OUT_SOFTMAX, cost=O.COST_CROSSENTROPY, init_kind=2, seed=42)
    env_note = "NVIDIA_TF32_OVERRIDE=" + os.environ.get("NVIDIA_TF32_OVERRIDE", "unset")
    step_s = _ref_step_seconds(O, layers, x_rows, y_rows, nb, it_warm, it_warm + it_timed, **kw)
    secs = step_s * K
    value = B * K / secs
    del x, y, x_rows

    secondary = {}
    if not args.no_secondary and args.workload == "c3":
        try:        # C1 (main.cpp:126-157)
            cases = _load_cases()
            labs = cases.mnist_labels()
            xr = cases.zscore(cases.synth_digits(labs, 784, seed=1234))
            t = _ref_step_seconds(O, [784, 50, 10], xr, labs.astype(np.float32).reshape(-1, 1), 10, 5, 45,
                                  **dict(kw, rate=0.075))
            secondary["c1"] = {"us_per_step": round(t * 1e6, 2), "samples_per_s": round(4200 / t), "batch": 4200}
            del xr
        except Exception as ex:  # noqa: BLE001
            secondary["c1"] = {"failed": str(ex)}
        try:        # C2 (main.cpp:54-96)
            cases = _load_cases()
            b = cases.bikeshare()
            t = _ref_step_seconds(O, C2_LAYERS, cases.zscore(b["x"]), b["y"], 1, 500, 4500,
                                  **dict(kw, rate=0.003, hidden_act=O.SIGMOID, classify=False, init_kind=1))
            secondary["c2"] = {"us_per_step": round(t * 1e6, 2), "samples_per_s": round(b["x"].shape[0] / t)}
        except Exception as ex:  # noqa: BLE001
            secondary["c2"] = {"failed": str(ex)}
        try:        # C4 at 8192 rows (the reference's int arena sizes overflow at >= 32768 rows, cublasNN.cpp:400, 444-452)
            l4, B4, _ = WORKLOADS["c4"]
            x4, _y4, lab4 = synth_batches(l4, B4, 2, seed=3, rank=0)
            t = _ref_step_seconds(O, l4, np.ascontiguousarray(x4[1:].numpy().T), lab4.numpy().astype(np.float32).reshape(-1, 1),
                                  2, 1, 3, repeats=2, **kw)
            secondary["c4"] = {"ms_per_step": round(t * 1e3, 3), "samples_per_s": round(B4 / t), "batch": B4,
                               "tflops": round(step_flops(l4, B4) / t / 1e12, 1),
                               "note": "single GPU, 8192 rows: the reference is single-GPU and its int arena sizes overflow at >= 32768 rows"}
            del x4
        except Exception as ex:  # noqa: BLE001
            secondary["c4"] = {"failed": str(ex)}
        try:        # C5 end to end: predictClassify incl. its cudaMalloc + pageable H2D (cublasNN.cpp:906-977)
            th = O.init_weights([784, 50, 10], 2, 42)
            c5 = {}
            for Bn in (1, 4096, 1 << 18):
                xr = np.random.default_rng(0).standard_normal((Bn, 784)).astype(np.float32)
                O.ref_predict([784, 50, 10], th, xr)
                reps = 10 if Bn <= 4096 else 2
                t0 = time.perf_counter()
                for _ in range(reps):
                    O.ref_predict([784, 50, 10], th, xr)
                dt = (time.perf_counter() - t0) / reps
                c5[str(Bn)] = {"us": round(dt * 1e6, 1), "samples_per_s": round(Bn / dt)}
            secondary["c5"] = {"e2e": c5, "note": "predictClassify through the harness: host rows in, class ids out (wall clock)"}
        except Exception as ex:  # noqa: BLE001
            secondary["c5"] = {"failed": str(ex)}

    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": round(value, 1), "unit": "samples/s",
            "n_gpus": int(getattr(args, "gpus", 1) or 1), "gpus_used": 1, "steps": K, "warmup": it_warm * nb,
            "ms_per_step": round(secs / K * 1e3, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32" if os.environ.get("NVIDIA_TF32_OVERRIDE", "0") != "1" else "tf32(cuBLAS override)",
            "data": "synthetic",
            "config": {"workload": desc_text, "layers": layers, "batch_per_gpu": B, "global_batch": B,
                       "parallelism": "dp1 (the reference is single-GPU: it runs on cuda:0 whatever --gpus says, see gpus_used)", "math": env_note,
                       "timing": "difference of two trainClassifyMomentum calls (epochs E and E+K/nb): setup and final cost cancel; each the minimum of 3 calls"},
            "tflops": round(step_flops(layers, B) * K / secs / 1e12, 2),
            "cpu_baseline": {"value": round(value, 1), "unit": "samples/s", "cores": 0, "kind": "reference",
                             "sample": "unmodified reference built from /root/reference (oracle/_ref), its cuBLAS path on cuda:0; the reference has no CPU path"},
            "e2e": {"value": round(value, 1), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "secondary": secondary}
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32", "tf32x3", "bf16"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-secondary", action="store_true", help="headline only (skip the C1/C2/C4/C5/tf32x3 legs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
