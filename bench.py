"""Benchmark of the hot path: MLP train samples/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload wide|mnist] [--batch-per-gpu B]

Workload at N=1 (default): BASELINE.json configs[3] -- 4 x Dense(4096) with 3 ReLU, D_in = C =
4096, batch 8192, fp32 storage, 3xTF32 tensor-core GEMMs, fused global-softmax CE, fused Adam;
synthetic data and reference-style Xavier-uniform weights (np.random.seed(0)).
At N>1 (launched with torchrun, one rank per GPU): the same model, batch 8192 PER GPU (weak
scaling; N=8 is configs[4]'s global batch 65536), rows sharded by rank, one NCCL SUM all-reduce
of the flat gradient arena per step plus the 2-float CE-normaliser all-gather.

A "step" = zero_grad + forward + loss + backward + (all-reduce) + optimizer step.
  value : steps timed with CUDA events on the compute stream, inputs resident in HBM
  e2e   : the same step driven from HOST buffers: per step the batch (inputs + one-hot labels)
          is copied from pinned host memory, and the loss is read back to the host
  roofline     : the tcgen05 GEMM kernel, per-launch time from CUDA events inside the timed steps
  cpu_baseline : oracle/ref_numpy.py (numpy restatement of the reference) on the host cores, on a
                 bounded sample (rank 0, N=1 only)
--impl reference times that CPU implementation alone, on the same config/metric/unit.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDE = dict(name="wide_mlp_4x4096", widths=[4096, 4096, 4096, 4096], d_in=4096, batch=8192)
MNIST = dict(name="mnist_mlp_784-200-100-70-30-10", widths=[200, 100, 70, 30, 10], d_in=784, batch=128)


# ------------------------------------------------------------------------------------------------
def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"],
                    bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons sampled every 100 ms while the timed region runs"""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """index of the next sample: brackets the timed region inside a longer-running sampler"""
        return len(self.lines)

    def stop(self, first=0, last=None):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = self.lines[first:last] or self.lines[-3:]
        for ln in window:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        busy = [s for s in sm if s > 0]
        return dict(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)),
                    reasons=sorted(reasons), samples=len(sm))


def make_config(cfg, batch_per_gpu, world):
    """the `config` object both arms print: the workload the metric is quoted on"""
    return {"workload": cfg["name"], "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * world,
            "d_in": cfg["d_in"], "widths": cfg["widths"], "optimizer": "Adam(1e-3)",
            "parallelism": "dp%d" % world,
            "l2": "per-step working set (parameters, activations, tf32 planes) far exceeds the 126 MB L2 "
                  "for the wide MLP; no flush between steps"}


def gemm_flops_per_step(cfg, batch):
    """2*M*N*K over the forward, dX (all layers but the first) and dW products"""
    dims = [cfg["d_in"]] + cfg["widths"]
    fl = 0
    for i in range(len(cfg["widths"])):
        mnk = batch * dims[i] * dims[i + 1]
        fl += 2 * mnk * (3 if i > 0 else 2)
    return fl


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the numpy oracle on the host cores
# ------------------------------------------------------------------------------------------------
def _oracle_model(cfg, sample_batch, seed=0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_numpy as R
    np.random.seed(seed)
    rng = np.random.RandomState(seed)
    C = cfg["widths"][-1]
    x = rng.rand(sample_batch, cfg["d_in"]).astype(np.float32)
    labels = np.eye(C, dtype=np.float32)[rng.randint(0, C, sample_batch)]
    mlp = R.RefMLP(cfg["widths"], R.RefAdam(lr=1e-3))
    return mlp, x, labels


def cpu_reference_throughput(cfg, sample_batch, steps, warmup):
    mlp, x, labels = _oracle_model(cfg, sample_batch)
    for _ in range(warmup):
        mlp.train_step(x, labels)
    t0 = time.perf_counter()
    for _ in range(steps):
        mlp.train_step(x, labels)
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps


def pick_reference_sample(cfg, total_steps, budget_s=150.0):
    """largest power-of-two sample batch (<= the config's batch) whose steps fit the time budget"""
    probe = 32 if cfg["batch"] >= 32 else cfg["batch"]
    mlp, x, labels = _oracle_model(cfg, probe)
    mlp.train_step(x, labels)                    # float32 parameters
    t0 = time.perf_counter()
    mlp.train_step(x, labels)                    # float64 parameters from here on
    t = time.perf_counter() - t0
    b = probe
    # step time grows (sub-)linearly with the batch; the optimiser part is batch independent
    while b * 2 <= min(cfg["batch"], 1024) and t * 2 * total_steps <= budget_s:
        b *= 2
        t *= 2
    return b


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = pick_reference_sample(cfg, args.steps + args.warmup)
    value, s_per_step = cpu_reference_throughput(cfg, sample, args.steps, args.warmup)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    desc = "train steps of the %s at batch %d (numpy oracle, all host threads)" % (cfg["name"], sample)
    line = {
        "impl": "reference", "metric": "MLP train samples/sec", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(make_config(cfg, cfg["batch"], args.gpus), sample_batch=sample),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def build_model(cfg):
    from core.layers import Dense, ReLU
    from core.losses import SoftmaxCrossEntropyLoss
    from core.model import Model
    from core.nn import Net
    from core.optimizer import Adam
    layers = []
    for i, w in enumerate(cfg["widths"]):
        layers.append(Dense(w))
        if i + 1 < len(cfg["widths"]):
            layers.append(ReLU())
    net = Net(layers)
    return net, Model(net=net, loss=SoftmaxCrossEntropyLoss(), optimizer=Adam(lr=1e-3))


def run_b200_arm(args, cfg):
    import core._backend as be
    import core._dist as dist
    from core.losses import SoftmaxCrossEntropyLoss
    from core.tensor import Tensor

    dist.init_process_group()
    rank, world = dist.rank(), dist.world_size()
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch N>1 with torchrun)" % (args.gpus, world))
    B = args.batch_per_gpu or cfg["batch"]
    if args.global_batch:
        lo, hi = dist.shard_bounds(args.global_batch, rank, world)
        if (hi - lo) * world != args.global_batch:
            raise SystemExit("--global-batch must be divisible by the number of GPUs")
        B = hi - lo
    C = cfg["widths"][-1]
    peaks = read_peaks()

    # synthetic inputs, one shard per rank; identical reference-style init on every rank
    rng = np.random.RandomState(1000 + rank)
    x_host = be.PinnedArray((2, B, cfg["d_in"]), np.float32)
    y_host = be.PinnedArray((2, B, C), np.float32)
    for k in range(2):
        x_host.array[k] = rng.rand(B, cfg["d_in"]).astype(np.float32)
        y_host.array[k] = 0.0
        y_host.array[k][np.arange(B), rng.randint(0, C, B)] = 1.0
    np.random.seed(0)
    net, model = build_model(cfg)
    loss_layer = SoftmaxCrossEntropyLoss()

    x_dev = [be.from_numpy(x_host.array[k]) for k in range(2)]
    y_dev = [be.from_numpy(y_host.array[k]) for k in range(2)]

    use_graph = args.graph == "on" or (args.graph == "auto" and args.workload == "mnist")

    def train_step(x_t, y_t):
        if use_graph:   # the same five calls, recorded once per batch shape and replayed
            return model.train_step(x_t, y_t)
        model.zero_grad()
        pred = model.forward(x_t)
        loss = loss_layer.loss(pred, y_t)
        loss.backward()
        model.step()
        return loss

    # ---- device-resident throughput ("value") --------------------------------------------------
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()                       # nvidia-smi needs ~0.3 s to produce its first sample
    for i in range(args.warmup):
        loss = train_step(Tensor(x_dev[i % 2]), Tensor(y_dev[i % 2]))
    float(loss.values)
    dist.barrier()
    s_first = sampler.mark()
    if not use_graph:
        be.prof_enable(1)     # per-launch events around the tcgen05 GEMM (not recordable in a graph)
    launches0 = be.launch_count()
    ev0, ev1 = be.Event(), be.Event()
    ev0.record()
    for i in range(args.steps):
        loss = train_step(Tensor(x_dev[i % 2]), Tensor(y_dev[i % 2]))
    ev1.record()
    dist.barrier()
    ms = ev1.elapsed_ms_since(ev0)
    s_last = sampler.mark() + 1
    launches = be.launch_count() - launches0
    gemm_ms, gemm_n = be.prof_collect()
    be.prof_enable(0)
    clocks = sampler.stop(s_first, s_last)
    last_loss = float(loss.values)

    # ---- end to end from host buffers ("e2e"): the public input pipeline ----------------------
    # utils.data_iterator.PrefetchIterator over a host data set of 4 batches (pinned in place):
    # batch i+1 is DMA'd on the copy stream while step i runs; the loss is read back every step.
    from utils.data_iterator import PrefetchIterator
    n_host_batches = 4
    x_data = np.empty((n_host_batches * B, cfg["d_in"]), np.float32)
    y_data = np.zeros((n_host_batches * B, C), np.float32)
    for k in range(n_host_batches):
        x_data[k * B:(k + 1) * B] = x_host.array[k % 2]
        y_data[k * B:(k + 1) * B] = y_host.array[(k + 1) % 2]
    del x_host, y_host
    e2e_steps = args.steps
    feed = iter(PrefetchIterator(batch_size=B, loop=True)(x_data, y_data))
    for _ in range(2):                                   # e2e warm-up
        batch = next(feed)
        float(train_step(batch.inputs, batch.targets).values)
    dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = be.Event(), be.Event()
    e0.record()
    batch = next(feed)
    for _ in range(e2e_steps):
        loss = train_step(batch.inputs, batch.targets)   # queued; the GPU starts on it
        batch = next(feed)                               # H2D of a following batch is queued in here,
                                                         # while the GPU runs the step just queued
        float(loss.values)                               # D2H read of this step's loss
    e1.record()
    dist.barrier()
    e2e_ms = e1.elapsed_ms_since(e0)
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_ms, e2e_wall)
    feed.close()

    # ---- max over ranks ------------------------------------------------------------------------
    if world > 1:
        import torch
        import torch.distributed as td
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    global_batch = B * world
    value = global_batch * args.steps / (ms * 1e-3)
    e2e_value = global_batch * e2e_steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 3xTF32 GEMM) ---------------------------------
    flops_step = gemm_flops_per_step(cfg, B)
    roofline = None
    if gemm_n:
        per_launch_flops = flops_step * args.steps / gemm_n
        achieved = per_launch_flops / (gemm_ms / gemm_n * 1e-3) / 1e12
        # tensor-pipe cost of one algorithmic MMA, in bf16-MMA equivalents (a TF32 MMA costs two):
        #   mix    : 1 TF32 + 2 BF16 = 4        tf32x3 : 3 TF32 = 6
        mix = be.TC_SPLIT == "mix"
        cost = 4.0 if mix else 6.0
        peak = peaks["bf16_sustained"] / cost
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tpath):   # dram bytes per launch from the committed ncu --set full capture
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "frac_of_burst": achieved / (peaks["bf16_burst"] / cost),
                    "frac_of_nominal": achieved / (2250.0 / cost),
                    "kernel": "gemm_tf32x3_kernel<MIX=%d>" % int(mix), "launches": int(gemm_n),
                    "avg_launch_ms": gemm_ms / gemm_n, "share_of_step": gemm_ms / ms,
                    "split": be.TC_SPLIT,
                    "peak_source": "%s bf16 sustained %.1f TFLOP/s / %d (%s)" % (
                        peaks["source"], peaks["bf16_sustained"], int(cost),
                        "1 TF32 + 2 BF16 MMAs per algorithmic MMA" if mix else "3 TF32 MMAs per algorithmic MMA")}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = 256 if cfg["batch"] >= 256 else cfg["batch"]
        v, s_per = cpu_reference_throughput(cfg, sample, steps=2, warmup=0)
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        cpu_baseline = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                        "sample": "2 train steps of the %s at batch %d with oracle/ref_numpy.py "
                                  "(%.2f s/step)" % (cfg["name"], sample, s_per)}

    if rank == 0:
        line = {
            "metric": "MLP train samples/sec", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.global_batch else "weak",
            "vs_baseline": None,
            "dtype": "f32 (tensor-core GEMMs: %s, fp32 accumulate)" % (
                "tf32 main term + bf16 cross terms" if be.TC_SPLIT == "mix" else "3xTF32"),
            "data": "synthetic",
            "config": dict(make_config(cfg, B, world),
                           step_mode="cuda_graph_replay" if use_graph else "eager_launches"),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "samples/s",
                    "h2d_bytes_per_step": int(B * (cfg["d_in"] + C) * 4), "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks, "final_loss": last_loss,
            "gemm_tflops_algorithmic": flops_step * args.steps / (ms * 1e-3) / 1e12,
        }
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="wide", choices=["wide", "mnist"])
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: fix the global batch (BASELINE config 5 uses 65536) and "
                         "split it over the ranks; reported with \"scaling\": \"strong\"")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step as a CUDA graph (Model.train_step); auto = on for the "
                         "launch-bound mnist workload, off for the wide MLP")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = dict(WIDE if args.workload == "wide" else MNIST)
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_b200_arm(args, cfg)


if __name__ == "__main__":
    main()
