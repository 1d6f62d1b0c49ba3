#!/usr/bin/env python
"""Headline benchmark: reactions/s of the CGR-MPNN-3D forward (BASELINE.json configs) on B200.

    python bench.py --gpus N --steps K --warmup W                     # cfg-2 (the headline): d4 h400 batch 64
    python bench.py --config cfg4 ...                                  # 1 M reactions, batch 8192, sharded over the ranks
    python bench.py --config cfg5 ...                                  # drug-like, d6 h1024, batch 1024
    python bench.py --impl reference --gpus N --steps K --warmup W     # the reference's CPU path (oracle port)

A step is one forward pass of the hot path over one batch of synthetic reactions (Fa = 78 + 768 synthetic MACE
columns, Fb = 14, random-init weights in the reference .pth layout).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from cgr_mpnn_3d_b200.data import Batch, make_batch, make_reactions  # noqa: E402

L2_BYTES = 126e6

# BASELINE.json configs: [1] cfg-2 (headline), [3] cfg-4, [4] cfg-5
CONFIGS = {
    "cfg2": dict(fa=846, fb=14, depth=4, hidden=400, batch=64, kind="t1x", n_params=1485205,
                 name="cfg-2: CGR-MPNN-3D d4 h400 learnable-skip forward (inference), batch 64/GPU"),
    "cfg4": dict(fa=846, fb=14, depth=4, hidden=400, batch=8192, kind="t1x", n_params=1485205,
                 name="cfg-4: high-throughput screening, CGR-MPNN-3D d4 h400 learnable-skip forward, batch 8192, "
                      "reactions sharded over the GPUs with no communication"),
    "cfg5": dict(fa=846, fb=14, depth=6, hidden=1024, batch=1024, kind="drug", n_params=9096199,
                 name="cfg-5: stress shape, drug-like reactions (80-120 atoms), d6 h1024 learnable-skip forward, "
                      "batch 1024/GPU"),
}


def workload_string(cfg) -> str:
    return (f"{cfg['name']}, Fa={cfg['fa']} Fb={cfg['fb']}, "
            f"{'T1x-shaped' if cfg['kind'] == 't1x' else 'drug-like'} synthetic reactions, random-init weights in the "
            f"reference .pth layout")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--engine", default=os.environ.get("CGR_ENGINE", "auto"))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fast"],
                    help="fp32: FP16x3 split (1e-4 parity mode, the headline); fast: single-pass fp16 operands, "
                         "reported with its measured error")
    ap.add_argument("--batch", type=int, default=None, help="override the config's batch size")
    ap.add_argument("--pool", type=int, default=None, help="distinct resident batches rotated through (> L2)")
    ap.add_argument("--repeats", type=int, default=50, help="timed regions of --steps steps; the median is reported")
    ap.add_argument("--no-graph", action="store_true", help="time eager calls instead of CUDA-graph replay")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--leg-seconds", type=float, default=0.6, help="work per e2e / store measurement")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=16, help="CUDA streams the independent steps are pipelined over")
    ap.add_argument("--group", type=int, default=20,
                    help="steps issued per GNN.forward_group call (two launches for the whole group); 1 = one "
                         "GNN.forward call per step, pipelined over --streams")
    ap.add_argument("--coalesce", type=int, default=8,
                    help="host batches predict_stream submits together in the e2e leg (1 = one submission per batch)")
    ap.add_argument("--policy", default="auto", choices=["auto", "latency", "throughput"],
                    help="kernel configuration of the pipelined region (auto: throughput when several streams are used)")
    ap.add_argument("--no-store", dest="store", action="store_false", default=True)
    ap.add_argument("--no-train", dest="train", action="store_false", default=True)
    ap.add_argument("--no-collate", dest="collate", action="store_false", default=True)
    ap.add_argument("--job-reactions", type=int, default=1_000_000, help="cfg4: reactions of the sharded screening job")
    ap.add_argument("--job-unique", type=int, default=65536, help="cfg4: distinct reactions held per GPU (drawn with repeats)")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return (float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), float(d.get("bf16_tflops_sustained", 1400.0)),
                "measured")
    return 6650.0, 1590.0, 1400.0, "fallback"


def algorithmic_bytes_fwd(n, e, b, cfg, s=4):
    """SURVEY.md §8(d) forward byte formula (layer-wise formulation, fp32, int64 indices)."""
    fa, fb, h, d = cfg["fa"], cfg["fb"], cfg["hidden"], cfg["depth"]
    return (s * (n * fa + e * fb) + (16 * e + 8 * n) + s * e * h + d * s * (3 * e * h + 2 * n * h)
            + s * (e * h + n * fa + 2 * b * h) + 4 * b + 4 * cfg["n_params"])


def algorithmic_flops_fwd(n, e, b, cfg):
    fa, fb, h, d = cfg["fa"], cfg["fb"], cfg["hidden"], cfg["depth"]
    return 2 * e * (fa + fb) * h + d * 2 * e * h * h + 2 * n * (fa + h) * h + 2 * b * h


def stage_bytes(stage, n, e, b, cfg, s=4):
    """Algorithmic bytes ONE launch of a stage accounts for: the matching terms of the §8(d) formula."""
    fa, fb, h, d = cfg["fa"], cfg["fb"], cfg["hidden"], cfg["depth"]
    per_depth = s * (3 * e * h + 2 * n * h)              # read h_l, read h0, write h_{l+1}, write + read atom sums
    readout = s * (e * h + 2 * b * h) + 4 * b             # read h_d, pooled write + read, energies
    return {
        "bond_layer": per_depth,
        "tc_readout": readout,
        # edge initialisation (bond features, indices, write h0) + every bond layer + the readout in one launch
        "tc_fwd_fused": s * (e * fb + e * h) + 16 * e + 8 * n + d * per_depth + readout,
        "tc_atom_proj": s * (2 * n * fa) + 4 * 2 * h * fa,   # x for the edge initialisation and for the readout, W_x / W_ox
        "tc_edge_init": s * (e * fb + e * h) + 16 * e + 8 * n,   # bond features, indices, write h0
        "gemm_bond_update": s * (3 * e * h) + 4 * h * h,
    }.get(stage)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons streamed (-lms) while the GPU legs run (B200_PROFILING.md);
    samples carry a host timestamp so the ones inside the timed region can be told apart."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self._t = None
        self.t_begin = self.t_end = None

    def _run(self):
        try:
            for line in self.proc.stdout:
                cols = [c.strip() for c in line.strip().split(",")]
                if len(cols) >= 7:
                    self.rows.append((time.perf_counter(), cols))
        except Exception:
            pass

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception:
            self.proc = None
        return self

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
                self.proc.wait(timeout=5)
            except Exception:
                pass
        if self._t is not None:
            self._t.join(timeout=5)

    def summary(self):
        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[0])); mx.append(float(r[1]))
                except Exception:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        timed = [x for x in self.rows if self.t_begin is not None and self.t_begin <= x[0] <= (self.t_end or 1e30)]
        window = "timed regions of `value`"
        if len(timed) < 3:      # timed regions shorter than the sampler period: use every sample of the GPU-busy legs
            timed, window = self.rows, "all GPU legs of this run (timed regions shorter than the sampling period)"
        sm, mx, reasons = parse(timed)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": window}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def build_model(cfg, engine: str, device, precision: str = "fp32"):
    from cgr_mpnn_3D.models.GNN import GNN
    torch.manual_seed(0)
    d = cfg["depth"]
    m = GNN(cfg["fa"], cfg["fb"], depth=d, hidden_sizes=[cfg["hidden"]] * d, dropout_ps=[0.0] * d,
            activation_fn=F.relu, use_learnable_skip=True)
    m.engine = engine
    m.precision = precision
    return m.to(device)


def build_oracle(cfg, dtype=torch.float32):
    from oracle.gnn_oracle import OracleGNN
    torch.manual_seed(0)
    d = cfg["depth"]
    return OracleGNN(cfg["fa"], cfg["fb"], depth=d, hidden_sizes=[cfg["hidden"]] * d, dropout_ps=[0.0] * d,
                     activation_fn=F.relu, use_learnable_skip=True).to(dtype).eval()


def cpu_reference_leg(cfg, batch_size: int, steps: int, warmup: int, budget_s: float, threads: int, passes_per_step: int = 1):
    """The reference's CPU implementation of the path (oracle port of GNN.py, bit-identical to the reference's fp32
    output) on the host cores: `steps` forward passes, bounded by `budget_s` seconds.  cfg-5 steps are sampled at 32
    reactions per pass (one full 1024-reaction pass is minutes of CPU time); throughput is per reaction either way."""
    torch.set_num_threads(threads)
    model = build_oracle(cfg)
    sample_b = batch_size if cfg["kind"] == "t1x" else min(batch_size, 32)
    n_b = 4 if sample_b <= 1024 else 1
    batches = [make_batch(sample_b, seed=9000 + i, kind=cfg["kind"], fa=cfg["fa"]) for i in range(n_b)]
    pps = max(1, passes_per_step)
    with torch.no_grad():
        t0 = time.perf_counter()
        for i in range(max(0, warmup) * pps):
            model(batches[i % n_b])
            if time.perf_counter() - t0 > budget_s:           # a slow config: do not spend the whole budget warming up
                break
        t0 = time.perf_counter()
        done = 0
        while done < steps * pps and (done == 0 or (time.perf_counter() - t0) < budget_s):
            model(batches[done % n_b])
            done += 1
        dt = time.perf_counter() - t0
    return {"value": sample_b * done / dt, "unit": "reactions/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": f"{done} forward passes of one {sample_b}-reaction batch "
                                      f"(oracle/gnn_oracle.py, fp32, {torch.get_num_threads()} threads, {dt:.1f} s"
                                      + (f"; {pps} passes per step" if pps > 1 else "") + ")",
            "ms_per_step": 1e3 * dt / done * (batch_size / sample_b), "steps_done": max(1, done // pps)}


def pin_rank_to_cores(local_rank: int, world: int) -> str:
    """One process per GPU on a shared host: give every rank its own slice of the cores this process may use, so the
    submitting threads of the ranks do not migrate over each other."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(cores) < 2 * world:
            return "unpinned"
        per = len(cores) // world
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return f"cores {mine[0]}-{mine[-1]}"
    except Exception:
        return "unpinned"


def bench_config(workload: str, cfg: dict, B: int) -> dict:
    """The `config` object of the JSON line: the WORKLOAD only, static text and sizes, identical in this framework's arm
    and in the reference arm (the driver compares the two); everything that describes how THIS arm ran it (engine, CUDA
    graph, streams, group, measured side figures, the realised L2 pool) goes to the line's `run` object."""
    return {"workload": workload, "batch_per_gpu": B, "depth": cfg["depth"], "hidden": cfg["hidden"], "fa": cfg["fa"],
            "fb": cfg["fb"], "parameters": cfg["n_params"],
            "l2": "GPU arm: inputs LARGER than the 126 MB L2 -- the timed regions rotate through a pool of distinct "
                  "resident batches whose bytes exceed it (pool size and bytes in `run.l2`), only the weights "
                  "(%.1f MB) stay resident; reference arm: CPU, one batch" % (4 * cfg["n_params"] / 1e6)}


def timed_loop(seconds: float, fn):
    """Call fn() until `seconds` of wall clock have passed (at least 3 calls); returns (calls, elapsed)."""
    t0 = time.perf_counter()
    n = 0
    while n < 3 or time.perf_counter() - t0 < seconds:
        fn()
        n += 1
    torch.cuda.synchronize()
    return n, time.perf_counter() - t0


def main():
    args = parse()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["batch"] = args.batch
        cfg["name"] = cfg["name"].replace("batch 64", f"batch {args.batch}").replace("batch 8192", f"batch {args.batch}") \
            .replace("batch 1024", f"batch {args.batch}")
    B = cfg["batch"]
    workload = workload_string(cfg)
    metric = "reactions/sec (CGR-MPNN-3D d%d h%d fwd)" % (cfg["depth"], cfg["hidden"])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ---------------------------------------------------------------- reference arm (CPU) ----
    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        # a step of the reference arm is a bounded SAMPLE of the workload: 16 passes over 64-reaction batches (one pass
        # is 6 ms -- twenty of them measure thread start-up, not the path: 6.0 k against 10.5 k reactions/s in steady
        # state); ms_per_step stays per 64-reaction batch like this framework's
        pps = 16 if (B <= 1024 and cfg["kind"] == "t1x") else 1
        leg = cpu_reference_leg(cfg, B, args.steps, args.warmup, budget_s=150.0, threads=cores, passes_per_step=pps)
        line = {
            "impl": "reference", "metric": metric, "value": leg["value"],
            "unit": "reactions/s", "n_gpus": args.gpus, "steps": leg["steps_done"], "warmup": args.warmup,
            "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(workload, cfg, B),
            "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": "reactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        _emit(line)
        return 0

    # ---------------------------------------------------------------- this framework ---------
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the hot path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = pin_rank_to_cores(local_rank, world)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    from cgr_mpnn_3d_b200 import _lib
    from cgr_mpnn_3d_b200.collate import build_plan, plan_for, split_features_for
    lib = _lib.load()
    hbm_peak, tc_peak, tc_sustained, peak_kind = load_peaks()

    engine = args.engine
    small = B <= 1024 and cfg["kind"] == "t1x"
    model = build_model(cfg, engine, dev, args.precision).eval()
    n_streams_req = max(1, args.streams) if (small and not args.no_graph) else 1
    model.tile_policy = ("throughput" if n_streams_req > 1 else "latency") if args.policy == "auto" else args.policy
    lat_model = build_model(cfg, engine, dev, args.precision).eval()      # latency configuration: a lone forward
    if args.pool:
        n_pool = max(2, args.pool)
    else:
        n_pool = 60 if small else (4 if cfg["kind"] == "t1x" else 3)
    host = [make_batch(B, seed=1000 + 997 * rank + i, kind=cfg["kind"], fa=cfg["fa"]) for i in range(n_pool)]
    for hb in host:
        hb.y = None
    pool = [hb.to(dev) for hb in host]
    if not small:
        host = host[:1]                       # large batches: keep host memory bounded
    for b in pool:
        plan_for(b)
    torch.cuda.synchronize()
    resident = sum(b.x.numel() * 4 + b.edge_attr.numel() * 4 + b.edge_index.numel() * 8 + b.batch.numel() * 8
                   for b in pool)
    n_atoms = sum(b.num_nodes for b in pool) / n_pool
    n_bonds = sum(b.num_edges for b in pool) / n_pool

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, repeats = max(1, args.steps), max(1, args.repeats)
    # ---- leg 1: device-resident throughput.  ONE CUDA graph holds the whole timed region: `steps` forwards, batch
    # i % n_pool, spread over n_streams parallel branches (a screening job is a stream of independent batches), so a
    # rank's host issues one launch per region whatever the step count ----
    outs = [None] * n_pool
    with torch.no_grad():
        for i in range(n_pool):                   # eager pass: builds every plan (tile packing syncs once per batch)
            outs[i] = model(pool[i])
        model.check_numerics()
        torch.cuda.synchronize()
        c0 = lib.cgr_launch_count()
        model(pool[0])
        launches_per_step = lib.cgr_launch_count() - c0
        group = max(1, min(args.group, 24, steps)) if (small and engine != "simt") else 1
        n_chunks = -(-steps // group)
        n_streams = max(1, min(n_streams_req, n_chunks if group > 1 else steps))
        streams = [torch.cuda.Stream() for _ in range(n_streams)]
        side = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        # distinct timed regions rotate through the pool, so a batch comes back only after more than an L2 of other
        # inputs went through (20 steps x 3.5 MB = 70 MB per region, 60 batches = 210 MB in all)
        n_regions = max(1, min(4, n_pool // steps))

        def issue_region(r):
            """`steps` forwards of region r: batches (r * steps + i) % n_pool."""
            ids = [(r * steps + i) % n_pool for i in range(steps)]
            if group > 1:
                for c in range(n_chunks):
                    chunk = ids[c * group:(c + 1) * group]
                    with torch.cuda.stream(streams[c % n_streams]):
                        for k, o in zip(chunk, model.forward_group([pool[k] for k in chunk])):
                            outs[k] = o
            else:
                for i, k in enumerate(ids):
                    with torch.cuda.stream(streams[i % n_streams]):
                        outs[k] = model(pool[k])

        if group > 1:
            model.forward_group(pool[:group])           # warm: workspaces, tensor maps
            torch.cuda.synchronize()
            c0 = lib.cgr_launch_count()
            model.forward_group(pool[:group])
            launches_per_step = (lib.cgr_launch_count() - c0) / group
        regions = []
        if not args.no_graph:
            for r in range(n_regions):
                region = torch.cuda.CUDAGraph()
                with torch.cuda.graph(region, stream=side):
                    fork_c = torch.cuda.Event()
                    fork_c.record(side)
                    for st in streams:
                        st.wait_event(fork_c)
                    issue_region(r)
                    for st in streams:
                        ev_c = torch.cuda.Event()
                        ev_c.record(st)
                        side.wait_event(ev_c)
                regions.append(region)
        region_no = [0]

        def run_region():
            r = region_no[0] % n_regions
            region_no[0] += 1
            if regions:
                regions[r].replay()
            else:
                issue_region(r)
                for st in streams:
                    main.wait_stream(st)

        clocks = ClockSampler(local_rank).start()
        for _ in range(max(1, -(-max(3, args.warmup) // steps))):       # >= W untimed warm-up steps
            run_region()
        region_ms = []
        clocks.mark_begin()
        for _ in range(repeats):
            barrier()
            # a short device-side sleep lets the host enqueue the region before the clock starts: the events then
            # bracket device time, not the host's launch latency
            torch.cuda._sleep(200_000)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(main)
            run_region()
            ev1.record(main)
            barrier()
            region_ms.append(ev0.elapsed_time(ev1))
        clocks.mark_end()

        # single-stream latency of one step with the latency kernel configuration (one forward at a time)
        n_lg = min(8, n_pool)
        for i in range(n_lg):
            lat_model(pool[i])
        lat_graphs = []
        if not args.no_graph:
            for i in range(n_lg):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    lat_model(pool[i])
                lat_graphs.append(g)
        n_lat = min(max(steps, 50), 400) if small else min(steps, 20)
        for i in range(n_lg):
            lat_graphs[i].replay() if lat_graphs else lat_model(pool[i])
        torch.cuda.synchronize()
        lat0, lat1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lat0.record(main)
        for i in range(n_lat):
            if lat_graphs:
                lat_graphs[i % n_lg].replay()
            else:
                lat_model(pool[i % n_lg])
        lat1.record(main)
        torch.cuda.synchronize()
        single_stream_ms = lat0.elapsed_time(lat1) / n_lat
        # eager public-API rate on device batches: model(batch) issued call by call (host-bound for small batches)
        n_e, dt_e = timed_loop(min(args.leg_seconds, 0.5), lambda: lat_model(pool[0]))
        eager_device_rate = B * n_e / dt_e

    # ---- leg 2: per-stage CUDA-event timing (events recorded by the library around every launch of a stage, on the
    # launching stream, the stream kept busy so that an event pair brackets kernel time, not host launch latency) ----
    prof_steps = min(steps, 40) if small else min(steps, 6)
    stage_ms, stage_cnt = {}, {}
    prof_units = 0                          # steps the profiled launches processed (group mode: `group` per launch)
    with torch.no_grad():
        def prof_call(i):
            nonlocal prof_units
            if group > 1:
                ids = [(i * group + k) % n_pool for k in range(group)]
                model.forward_group([pool[k] for k in ids])
                prof_units += group
            else:
                lat_model(pool[i % n_pool])
                prof_units += 1
        for i in range(3):
            prof_call(i)
        prof_units = 0
        torch.cuda.synchronize()
        lib.cgr_profile_enable(1)
        n_prof = max(3, prof_steps // group) if group > 1 else prof_steps
        for i in range(n_prof):
            torch.cuda._sleep(2_000_000 if small else 200_000)
            prof_call(i)
        torch.cuda.synchronize()
        name = ctypes.create_string_buffer(64)
        ms = ctypes.c_float()
        for i in range(lib.cgr_profile_count()):
            if lib.cgr_profile_get(i, name, 64, ctypes.byref(ms)) == 0:
                k = name.value.decode()
                stage_ms[k] = stage_ms.get(k, 0.0) + ms.value
                stage_cnt[k] = stage_cnt.get(k, 0) + 1
        lib.cgr_profile_enable(0)
    dominant = max(stage_ms, key=stage_ms.get) if stage_ms else None

    # ---- leg 3: end to end through the public API with HOST buffers (H2D + index build + kernels + D2H per step) ----
    e2e = None
    if not args.skip_e2e and small:
        pinned = [hb.pin_memory() for hb in host]
        h2d = sum(t.numel() * t.element_size() for t in (pinned[0].x, pinned[0].edge_attr, pinned[0].edge_index,
                                                         pinned[0].batch, pinned[0].ptr))
        with torch.no_grad():
            for i in range(3):
                res = model(pinned[i % n_pool])
            barrier()
            n_s, dt_s = timed_loop(args.leg_seconds, lambda: model(pinned[0]))     # GNN.forward(host batch), call by call

            def stream_rate(coalesce):
                """GNN.predict_stream over pinned host batches for >= leg_seconds of work (pipeline fill / drain included)."""
                list(model.predict_stream((pinned[i % n_pool] for i in range(64)), depth=4, coalesce=coalesce))
                barrier()
                t0 = time.perf_counter()

                def feed():
                    i = 0
                    while i < 64 or time.perf_counter() - t0 < args.leg_seconds:
                        yield pinned[i % n_pool]
                        i += 1
                n_done = 0
                for r_ in model.predict_stream(feed(), depth=4, coalesce=coalesce):
                    n_done += 1
                    last = r_
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                assert last.numel() == B
                return n_done, dt
            n_u, dt_u = stream_rate(1)
            n_c, dt_c = stream_rate(args.coalesce)
            # what the host link can do at best: plain pinned H2D copies (4 x 64 MiB, 4 streams in flight)
            raw_h = [torch.empty(64 << 20, dtype=torch.uint8).pin_memory() for _ in range(4)]
            raw_d = [torch.empty_like(t, device=dev) for t in raw_h]
            cstreams = [torch.cuda.Stream() for _ in range(4)]
            for q, h_, d_ in zip(cstreams, raw_h, raw_d):
                with torch.cuda.stream(q):
                    d_.copy_(h_, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(8):
                for q, h_, d_ in zip(cstreams, raw_h, raw_d):
                    with torch.cuda.stream(q):
                        d_.copy_(h_, non_blocking=True)
            torch.cuda.synchronize()
            h2d_peak_gbs = 8 * 4 * raw_h[0].numel() / (time.perf_counter() - t0) / 1e9
            del raw_h, raw_d
        e2e = {"rate": B * n_c / dt_c, "steps": n_c, "single_rate": B * n_s / dt_s, "uncoalesced_rate": B * n_u / dt_u,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(res.numel() * 4), "h2d_peak_gbs": h2d_peak_gbs}
    elif not args.skip_e2e:
        # large batches: one host batch through GNN.forward (H2D of the whole batch, kernels, D2H), call by call
        pinned = [host[0].pin_memory()]
        h2d = sum(t.numel() * t.element_size() for t in (pinned[0].x, pinned[0].edge_attr, pinned[0].edge_index,
                                                         pinned[0].batch, pinned[0].ptr))
        with torch.no_grad():
            res = model(pinned[0])
            barrier()
            n_s, dt_s = timed_loop(args.leg_seconds, lambda: model(pinned[0]))
        e2e = {"rate": B * n_s / dt_s, "steps": n_s, "single_rate": B * n_s / dt_s, "uncoalesced_rate": None,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(res.numel() * 4), "h2d_peak_gbs": None}

    # ---- leg: collation cost of a FRESH batch (north-star subsystem 1): CSR a2b / b2a arrays, tile plan, feature
    # split -- everything `value` keeps cached on its resident batches ----
    collate_leg = None
    if args.collate and small:
        fresh = [Batch(b.x, b.edge_index, b.edge_attr, b.batch, b.ptr, None) for b in pool[:16]]

        def prep_once(k=[0]):
            b = fresh[k[0] % len(fresh)]
            k[0] += 1
            plan = build_plan(b.edge_index, b.num_nodes, b.batch, b.ptr)
            plan.ensure_tiles()                      # one host synchronisation (tile count)
            b.__dict__.pop("_cgr_xsplit", None)
            split_features_for(b, plan)
        for _ in range(5):
            prep_once()
        n_p, dt_p = timed_loop(0.3, prep_once)
        idx_bytes = 16 * n_bonds + 8 * n_atoms + 4 * (3 * n_bonds + n_atoms + 1)      # int64 edge_index + batch in, int32 CSR out
        collate_leg = {"us_per_batch": 1e6 * dt_p / n_p, "batches_per_s": n_p / dt_p,
                       "what": "fresh device batch -> cgr_csr_build (src, dst, in_ptr, in_idx, validity flags) + tile "
                               "plan (cgr_tc_plan_build/check, one host sync) + cgr_tc_split_features (x -> fp16 hi/lo)",
                       "index_bytes_per_batch": int(idx_bytes), "feature_bytes_per_batch": int(8 * n_atoms * cfg["fa"])}

    # ---- leg: device-resident reaction store (features stay in HBM, batches assembled by a kernel) ----
    store_leg = None
    if args.store and small:
        import numpy as np
        from cgr_mpnn_3d_b200.store import ReactionStore
        n_store = max(4096, 4 * B)
        store = ReactionStore.from_graphs(make_reactions(n_store, seed=9000 + rank, kind="t1x", fa=cfg["fa"]), device=dev)
        with torch.no_grad():
            for bt in store.loader(B, shuffle=True, seed=0):          # warm-up pass
                model(bt)
            barrier()
            t0 = time.perf_counter()
            n_rx, ep = 0, 0
            while time.perf_counter() - t0 < args.leg_seconds:
                for bt in store.loader(B, shuffle=True, seed=1 + ep):
                    model(bt)
                    n_rx += int(bt.y.numel())
                ep += 1
            torch.cuda.synchronize()
            dt_loader = time.perf_counter() - t0
            rng_o = np.random.default_rng(0)
            for _ in range(2):
                store.predict(model, batch_size=B, order=rng_o.permutation(len(store)))
            barrier()
            n_pred, dt_pred = timed_loop(args.leg_seconds,
                                         lambda: store.predict(model, batch_size=B, order=rng_o.permutation(len(store))))
        store_leg = {"rate": n_rx / dt_loader, "predict_rate": n_pred * len(store) / dt_pred,
                     "store_bytes": store.nbytes(), "store_reactions": len(store)}
        del store

    # ---- cfg4: the sharded screening job itself: job_reactions reactions, batch 8192, ReactionStore.predict per rank ----
    job = None
    if args.config == "cfg4":
        import numpy as np
        from cgr_mpnn_3d_b200.parallel import shard_range
        from cgr_mpnn_3d_b200.store import ReactionStore
        lo, hi = shard_range(args.job_reactions, rank, world)
        n_unique = min(args.job_unique, hi - lo)
        store = ReactionStore.from_graphs(make_reactions(n_unique, seed=7000 + rank, kind="t1x", fa=cfg["fa"]), device=dev)
        order = np.random.default_rng(rank).integers(0, n_unique, size=hi - lo)
        with torch.no_grad():
            store.predict(model, batch_size=B, order=order[: 8 * B])       # warm-up (workspaces of all 8 slots)
            barrier()
            j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            j0.record()
            t0 = time.perf_counter()
            out_job = store.predict(model, batch_size=B, order=order)
            j1.record()
            torch.cuda.synchronize()
            dt_job = time.perf_counter() - t0
            # spot check of a slice against the oracle (rank 0): the job's energies are the reference's
            job_err = None
            if rank == 0:
                from cgr_mpnn_3d_b200.data import collate_host
                from oracle.gnn_oracle import scale_normalised_error
                gs = make_reactions(n_unique, seed=7000 + rank, kind="t1x", fa=cfg["fa"])
                sl = [int(v) for v in order[:24]]
                ref = build_oracle(cfg)(collate_host([gs[k] for k in sl]))
                job_err = scale_normalised_error(out_job[:24].cpu(), ref)
        job = {"seconds": dt_job, "reactions": int(hi - lo), "store_bytes": store.nbytes(), "unique": n_unique,
               "oracle_slice_error": job_err}
        del store

    # ---- leg: training step (forward + MSE(sum) + explicit backward + gradient SUM over the replicas) ----
    train = None
    if args.train and args.config == "cfg2":
        train = train_leg(args, cfg, dev, rank, world, barrier, lib)

    # ---- reduce over ranks (max time), assemble the line ----
    med_ms = statistics.median(region_ms)
    vals = [med_ms, min(region_ms), (1.0 / e2e["rate"]) if e2e else 0.0, train["ms_step"] if train else 0.0,
            job["seconds"] if job else 0.0]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    med_ms, min_ms, e2e_inv, train_ms, job_s = (float(v) for v in t)
    ms_per_step = med_ms / steps
    value = B * world / (ms_per_step * 1e-3)

    cpu = cpu1 = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_reference_leg(cfg, B, 10 ** 9, 3, args.cpu_seconds, os.cpu_count() or 1)
        cpu1 = cpu_reference_leg(cfg, B, 10 ** 9, 1, min(args.cpu_seconds, 6.0), 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["single_thread"] = {k: cpu1[k] for k in ("value", "unit", "cores", "sample")}

    # fast mode: report the error it costs (scale-normalised, vs an fp64 evaluation of the same graphs)
    precision_info = None
    if rank == 0 and small:
        with torch.no_grad():
            hb = make_batch(B, seed=1000 + 997 * rank, kind=cfg["kind"], fa=cfg["fa"])
            o64 = build_oracle(cfg, torch.float64)
            ref = o64(Batch(hb.x.double(), hb.edge_index, hb.edge_attr.double(), hb.batch, hb.ptr, None))
            from oracle.gnn_oracle import scale_normalised_error
            precision_info = {"mode": args.precision,
                              "ea_error_vs_fp64": scale_normalised_error(lat_model(hb.to(dev)).cpu(), ref)}

    clocks.stop()
    roofline = None
    if dominant:
        per_unit = stage_bytes(dominant, n_atoms, n_bonds, B, cfg)
        if per_unit is None:
            per_unit = algorithmic_bytes_fwd(n_atoms, n_bonds, B, cfg) / max(1, launches_per_step)
        units_per_launch = prof_units / stage_cnt[dominant]              # group mode: `group` batches per launch
        # stage_bytes() is per launch for one batch: a stage launched once per layer (units_per_launch < 1) needs no scaling
        per_launch = per_unit * max(1.0, units_per_launch)
        lps = stage_cnt[dominant] / prof_units
        dur_us = 1e3 * stage_ms[dominant] / stage_cnt[dominant]          # CUDA events around the launch, busy stream
        share = stage_ms[dominant] / sum(stage_ms.values())
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                traffic = json.load(fh).get(dominant, {}).get(f"batch_{B}")
        except Exception:
            pass
        if traffic is not None and dominant == "tc_fwd_fused" and abs(units_per_launch - 20.0) > 1e-6 and B == 64:
            traffic = None                                               # the capture is of a 20-batch group launch
        pipe_us = ms_per_step * 1e3 * share / lps                        # machine time the pipelined region spends per launch
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": per_launch / (dur_us * 1e-6) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": per_launch / (dur_us * 1e-6) / 1e9 / hbm_peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": per_launch, "algorithmic_bytes_per_batch": per_unit,
                    "batches_per_launch": units_per_launch, "peak_kind": peak_kind,
                    "avg_launch_us": dur_us, "launches_per_step": lps,
                    "pipelined_us_per_launch": pipe_us,
                    "pipelined_frac": per_launch / (pipe_us * 1e-6) / 1e9 / hbm_peak,
                    "stage_share": {k: round(v / sum(stage_ms.values()), 4) for k, v in sorted(stage_ms.items())},
                    "note": "avg_launch_us / achieved / frac: the kernel's own duration (CUDA events around each launch on the "
                            "launching stream, stream kept busy) against its algorithmic bytes; pipelined_*: the kernel's "
                            "event share of a step x the timed region's time per step (independent forwards overlapped "
                            "over streams) -- machine time per launch, not a launch duration"}
        if dominant == "gemm_bond_update":
            # layer-wise path (graphs that do not tile, cfg-5): the dominant launch is a plain [E, H] x [H, H] GEMM whose
            # arithmetic intensity (H / 6 flop per algorithmic byte, x3 executed by the FP16x3 split) is past the ridge
            # of the machine at H = 1024: the tensor pipe bounds it, not HBM
            fl = 2.0 * n_bonds * cfg["hidden"] * cfg["hidden"] * max(1.0, units_per_launch)
            tensor_peak = tc_sustained                      # a kernel timed inside a long step: the sustained figure
            ridge = tensor_peak * 1e12 / (hbm_peak * 1e9)
            if 3.0 * fl / per_launch > ridge:
                tf = fl / (dur_us * 1e-6) / 1e12
                roofline.update({"bound": "tensor", "achieved": tf, "peak": tensor_peak, "unit": "TFLOP/s",
                                 "frac": tf / tensor_peak, "algorithmic_flops_per_launch": fl,
                                 "executed_tensor_flops_factor": 3, "executed_frac": 3.0 * tf / tensor_peak,
                                 "hbm_frac": per_launch / (dur_us * 1e-6) / 1e9 / hbm_peak,
                                 "peak_kind": peak_kind + " (dense bf16, sustained)"})
    if rank == 0:
        alg = algorithmic_bytes_fwd(n_atoms, n_bonds, B, cfg)
        flops = algorithmic_flops_fwd(n_atoms, n_bonds, B, cfg)
        step_s = ms_per_step * 1e-3
        line = {
            "metric": metric, "value": value, "unit": "reactions/s",
            "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (FP16x3-split tcgen05 MMA, fp32 accumulate; 2e-6 of fp64)" if args.precision == "fp32"
                     else "f16 single-pass tcgen05 MMA, fp32 accumulate (fast mode, error reported in `precision`)",
            "data": "synthetic",
            "config": bench_config(workload, cfg, B),
            "run": {"engine": engine, "precision": args.precision, "cuda_graph": not args.no_graph, "streams": n_streams,
                       "group": group,
                       "timed_region": (f"{steps} forwards issued as {n_chunks} GNN.forward_group call(s) of up to {group} "
                                        f"batches (two launches per call: atom projection + fused cluster kernel)"
                                        if group > 1 else f"{steps} GNN.forward calls") +
                                       f", captured as ONE CUDA graph over {n_streams} parallel branch(es); {n_regions} "
                                       f"such regions rotate through the batch pool; median of {repeats} regions (min "
                                       f"{min_ms / steps * 1e3:.2f} us/step), each between barrier + synchronize, max over ranks",
                       "single_stream_ms_per_step": single_stream_ms,
                       "eager_device_batch_rate": eager_device_rate,
                       "parallelism": f"replicas x{world}, no collective", "cpu_affinity": affinity,
                       "l2": f"inputs rotate over {n_pool} distinct resident batches ({resident / 1e6:.0f} MB > "
                             f"{L2_BYTES / 1e6:.0f} MB L2); weights ({4 * cfg['n_params'] / 1e6:.1f} MB) stay resident",
                       "atoms_per_batch": n_atoms, "bonds_per_batch": n_bonds},
            "whole_forward": {"algorithmic_bytes": alg, "achieved_gbs": alg / step_s / 1e9,
                              "hbm_frac": alg / step_s / 1e9 / hbm_peak,
                              "algorithmic_flops": flops, "achieved_tflops": flops / step_s / 1e12,
                              "tensor_frac_of_bf16_sustained": flops / step_s / 1e12 / tc_sustained,
                              "executed_tensor_flops_factor": 3 if args.precision == "fp32" else 1,
                              "peak_kind": peak_kind},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks.summary(),
            "gpu_launches": int(round(launches_per_step * steps)) * repeats,
            "launches_per_step": launches_per_step,
            "precision": precision_info,
        }
        if train:
            line["train_step"] = dict(train["line"], value=B * world / (train_ms * 1e-3), ms_per_step=train_ms)
        if collate_leg:
            line["collate"] = collate_leg
        if store_leg:
            line["resident_store"] = {"value": store_leg["rate"] * world, "unit": "reactions/s",
                                      "store_reactions": store_leg["store_reactions"], "store_bytes": store_leg["store_bytes"],
                                      "predict_value": store_leg["predict_rate"] * world,
                                      "predict_api": "ReactionStore.predict(model, batch_size): the per-batch loop in C "
                                                     "(cgr_store_infer), 8 streams, results stay on the device; consecutive "
                                                     "batches are assembled as super-batches of <= 1024 reactions, in best-fit order "
                                                     "for the 128-row tiles (a reaction's energy does not depend on its batch "
                                                     "or position: same result vector)",
                                      "what": "shuffled epochs over a ReactionStore held in HBM for >= %.1f s: per step one "
                                              "small index upload, the gather kernel, one-launch CSR and the forward (eager "
                                              "launches); no host-to-device copy of features" % args.leg_seconds}
        if job:
            line["screening_job"] = {"value": args.job_reactions / job_s, "unit": "reactions/s", "seconds": job_s,
                                     "reactions": args.job_reactions, "reactions_per_gpu": job["reactions"],
                                     "per_gpu_value": job["reactions"] / job_s, "batch": B,
                                     "unique_reactions_per_gpu": job["unique"], "store_bytes_per_gpu": job["store_bytes"],
                                     "hbm_frac_per_gpu": (alg / B) * job["reactions"] / job_s / 1e9 / hbm_peak,
                                     "oracle_slice_error": job["oracle_slice_error"],
                                     "what": "ReactionStore.predict over this rank's contiguous shard (parallel.shard_range) of "
                                             "the job: ids drawn with repeats from a resident set (a 1 M-reaction set is "
                                             "58 GB of HBM; host generation of it is the slow part), batches assembled on "
                                             "the device, no collective, max over ranks"}
        if e2e:
            e2e_rate = (1.0 / e2e_inv) * world if e2e_inv > 0 else None
            line["e2e"] = {"value": e2e_rate, "unit": "reactions/s",
                           "api": ("GNN.predict_stream(host batches of %d, depth=4, workers=2, coalesce=%d) for >= %.1f s: "
                                   "every step's H2D from its own pinned buffers + index build + kernels + D2H; up to %d "
                                   "consecutive batches share one submission" % (B, args.coalesce, args.leg_seconds,
                                                                                  args.coalesce)) if small
                                  else "GNN.forward(host batch of %d): slices of 1024 whole reactions pipelined over 4 streams (H2D of one slice "
                                       "overlaps the kernels of another) + index build + kernels + D2H" % B,
                           "steps": e2e["steps"],
                           "uncoalesced_value": e2e["uncoalesced_rate"] * world if e2e["uncoalesced_rate"] else None,
                           "uncoalesced_api": "GNN.predict_stream(..., coalesce=1): one submission per batch",
                           "single_call_value": e2e["single_rate"] * world, "single_call_api": "GNN.forward(host batch)",
                           "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                           "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                           "h2d_gbs": (e2e["h2d_bytes_per_step"] / B) * (e2e_rate / world) / 1e9 if e2e_rate else None,
                           "h2d_copy_peak_gbs": e2e["h2d_peak_gbs"],
                           "bound": "host link: h2d_gbs is the per-GPU input traffic the e2e rate implies, "
                                    "h2d_copy_peak_gbs plain pinned 64 MiB cudaMemcpyAsync copies on 4 streams of this box"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def train_leg(args, cfg, dev, rank, world, barrier, lib):
    """cfg-3: the data-parallel training step at batch 64 per GPU.  The timed step is the COMPLETE step the reference's
    trainer runs (trainer.py:139-144): forward, MSELoss(sum), explicit backward, gradient SUM over the replicas and the
    Adam(amsgrad) update -- forward + loss + backward replayed as one CUDA graph, then the exchange + update."""
    import torch.distributed as dist
    from cgr_mpnn_3d_b200.optim import FusedAdam
    from cgr_mpnn_3d_b200.parallel import allreduce_gradients_
    B = cfg["batch"]
    tm = build_model(cfg, "auto", dev).train()
    tb = [make_batch(B, seed=5000 + 997 * rank + i, kind="t1x", fa=cfg["fa"]).to(dev) for i in range(8)]
    n_t = min(max(args.steps, 20), 100)

    def train_step(d):
        loss = torch.nn.functional.mse_loss(tm(d), d.y, reduction="sum")      # train.py:120
        loss.backward()

    side_t = torch.cuda.Stream()
    side_t.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side_t):
        for d in tb:                      # every batch once: builds its index arrays / tile plan before capture
            tm.zero_grad(set_to_none=True)
            train_step(d)
    torch.cuda.current_stream().wait_stream(side_t)
    torch.cuda.synchronize()
    tgraphs = []
    tm.zero_grad(set_to_none=True)
    for d in tb:
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph, stream=side_t):
            train_step(d)
        tgraphs.append(gph)
    opt = FusedAdam(tm.parameters(), lr=1e-4, weight_decay=1e-5, amsgrad=True)

    def timed(fn, n):
        for i in range(5):
            fn(i)
        ms = []
        for _ in range(5):
            barrier()
            torch.cuda._sleep(200_000)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(n):
                fn(i)
            a1.record()
            barrier()
            ms.append(a0.elapsed_time(a1) / n)
        return statistics.median(ms)

    def step_fb(i):                       # forward + loss + backward + gradient exchange (optimizer excluded)
        tgraphs[i % len(tb)].replay()
        if world > 1:
            allreduce_gradients_(tm.parameters())

    def step_full(i):                     # the complete step: + Adam
        step_fb(i)
        opt.step()
    ms_fb = timed(step_fb, n_t)
    ms_full = timed(step_full, n_t)
    ms_captured = None
    if world > 1:
        # the same complete step with the NCCL all-reduce captured INSIDE the step's CUDA graph (issued by the graph right
        # behind the last backward kernel: no host round trip between backward and exchange)
        try:
            cgraphs = []
            tm.zero_grad(set_to_none=True)
            with torch.cuda.stream(side_t):
                for d in tb[:2]:
                    tm.zero_grad(set_to_none=True)
                    train_step(d)
                    allreduce_gradients_(tm.parameters())
            torch.cuda.current_stream().wait_stream(side_t)
            torch.cuda.synchronize()
            tm.zero_grad(set_to_none=True)
            for d in tb:
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=side_t):
                    train_step(d)
                    allreduce_gradients_(tm.parameters())
                cgraphs.append(gph)

            def step_captured(i):
                cgraphs[i % len(tb)].replay()
                opt.step()
            ms_captured = timed(step_captured, n_t)
        except Exception as exc:                      # capture of the collective is not available in this build
            print(f"[bench] NCCL capture in the step graph failed: {exc}", file=sys.stderr)
            ms_captured = None
    # the same step issued eagerly, as a plain training loop does: host-bound, wall clock between two synchronisations
    for i in range(5):
        tm.zero_grad(set_to_none=True)
        train_step(tb[i % len(tb)])
    barrier()
    t0 = time.perf_counter()
    for i in range(n_t):
        tm.zero_grad(set_to_none=True)
        train_step(tb[i % len(tb)])
        if world > 1:
            allreduce_gradients_(tm.parameters())
    barrier()
    eager_ms = (time.perf_counter() - t0) * 1e3 / n_t

    def time_opt(o, n=100):
        for _ in range(5):
            o.step()
        torch.cuda.synchronize()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(n):
            o.step()
        o1.record()
        torch.cuda.synchronize()
        return o0.elapsed_time(o1) / n
    adam_fused_ms = time_opt(opt)
    adam_torch_ms = time_opt(torch.optim.Adam(tm.parameters(), lr=1e-4, weight_decay=1e-5, amsgrad=True))
    dp = None
    if world > 1:
        # the same complete step with the gradient exchange and the update as ONE kernel over NVLink peer memory
        from cgr_mpnn_3d_b200 import ops as _ops
        from cgr_mpnn_3d_b200.optim import PeerFusedAdam
        tm2 = build_model(cfg, "auto", dev).train()
        popt = PeerFusedAdam(tm2.parameters(), lr=1e-4, weight_decay=1e-5, amsgrad=True)

        def train_step2(d):
            torch.nn.functional.mse_loss(tm2(d), d.y, reduction="sum").backward()
        with torch.cuda.stream(side_t):
            for d in tb[:2]:
                tm2.zero_grad(set_to_none=True)
                train_step2(d)
        torch.cuda.current_stream().wait_stream(side_t)
        torch.cuda.synchronize()
        pgraphs = []
        for j, d in enumerate(tb):           # graph j writes its gradients into arena j % 2: replayed in order
            tm2.zero_grad(set_to_none=True)
            popt._cur = j % 2
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=side_t):
                train_step2(d)
            pgraphs.append(gph)
        popt._cur = 0
        state = {"i": 0}

        def peer_step(_):
            j = state["i"] % len(tb)
            pgraphs[j].replay()
            popt.step(arena=j % 2)
            state["i"] += 1
        ms_peer = timed(peer_step, n_t)
        _ops.set_grad_arena(None)
        dp = {"nccl_allreduce_plus_fused_adam_ms": ms_full, "nccl_in_graph_plus_fused_adam_ms": ms_captured,
              "peer_fused_adam_ms": ms_peer,
              "what": "graph-replayed fwd+loss+bwd, then gradient SUM over replicas and Adam(amsgrad): NCCL all-reduce + "
                      "one-launch Adam vs ONE kernel over NVLink peer memory (no NCCL)"}
    ms_step = ms_full if dp is None else min(v for v in (ms_full, ms_captured, dp["peer_fused_adam_ms"]) if v)
    line = {"unit": "reactions/s", "steps": n_t,
            "fwd_loss_bwd_exchange_ms": ms_fb, "eager_ms_per_step": eager_ms,
            "dp_step_with_optimizer": dp,
            "optimizer": {"fused_adam_ms": adam_fused_ms, "torch_adam_ms": adam_torch_ms,
                          "what": "Adam(weight_decay, amsgrad=True) over all parameters, timed apart (train.py:117-119)"},
            "what": "COMPLETE data-parallel training step, batch %d/GPU: forward + MSE(sum) + explicit backward (one CUDA "
                    "graph) + gradient SUM over the replicas + Adam(amsgrad) update; the faster of NCCL all-reduce + "
                    "FusedAdam and PeerFusedAdam when N > 1; median of 5 regions, max over ranks" % B}
    return {"ms_step": ms_step, "line": line}


def _emit(line: dict) -> None:
    """The one JSON line goes to the process's ORIGINAL stdout (see _quiet_stdout)."""
    data = (json.dumps(line) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    os.write(fd, data)


_REAL_STDOUT = None


def _quiet_stdout() -> None:
    """Libraries (NCCL's version banner, for one) write to file descriptor 1 behind Python's back; the contract is ONE
    JSON line on stdout, so everything else is routed to stderr at the descriptor level."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


if __name__ == "__main__":
    _quiet_stdout()
    sys.exit(main())
