#!/usr/bin/env python
"""Headline benchmark: reactions/s of the CGR-MPNN-3D d4 h400 forward (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this framework on N B200s
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A step is one forward pass of the hot path over one batch of 64 synthetic T1x-shaped reactions
(Fa = 78 + 768 synthetic MACE columns, Fb = 14, depth 4, hidden 400, learnable skip, random-init
weights in the reference .pth layout).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from cgr_mpnn_3d_b200.data import Batch, make_batch  # noqa: E402

FA, FB, DEPTH, HID, BATCH = 846, 14, 4, 400, 64
L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default=os.environ.get("CGR_ENGINE", "auto"))
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--pool", type=int, default=48, help="distinct resident batches rotated through (> L2)")
    ap.add_argument("--no-group", action="store_true",
                    help="replay one single-batch graph per step instead of one graph per group of --streams batches")
    ap.add_argument("--no-graph", action="store_true", help="time eager custom-op calls instead of CUDA-graph replay")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=8, help="CUDA streams the independent steps are pipelined over")
    ap.add_argument("--coalesce", type=int, default=8,
                    help="host batches predict_stream submits together in the e2e leg (1 = one submission per batch)")
    ap.add_argument("--store", action="store_true", default=True,
                    help="also time inference over a device-resident reaction store (no per-step feature copies) [default]")
    ap.add_argument("--no-store", dest="store", action="store_false")
    ap.add_argument("--train", action="store_true", default=True,
                    help="also time the training step (fwd+loss+bwd[+allreduce]) [default]")
    ap.add_argument("--no-train", dest="train", action="store_false")
    ap.add_argument("--peer-adam", action="store_true",
                    help="N > 1: also time the complete data-parallel step with NCCL all-reduce + FusedAdam and with "
                         "PeerFusedAdam (gradient sum over NVLink peer memory + Adam in one kernel)")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def algorithmic_bytes_fwd(n, e, b, fa=FA, fb=FB, h=HID, d=DEPTH, n_params=1485205, s=4):
    """SURVEY.md §8(d) forward byte formula (layer-wise formulation, fp32, int64 indices)."""
    return (s * (n * fa + e * fb) + (16 * e + 8 * n) + s * e * h + d * s * (3 * e * h + 2 * n * h)
            + s * (e * h + n * fa + 2 * b * h) + 4 * b + 4 * n_params)


def bond_update_bytes(n, e, h=HID, s=4):
    """per-depth term of the same formula: read h_l, read h0, write h_{l+1}, write+read atom sums."""
    return s * (3 * e * h + 2 * n * h)


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
        window = "timed region"
        if len(timed) < 3:      # timed region shorter than the sampler period: use every sample of the GPU-busy legs
            timed, window = self.rows, "all GPU legs of this run (timed region shorter than the sampling period)"
        sm, mx, reasons = parse(timed)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": window}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def build_model(engine: str, device):
    from cgr_mpnn_3D.models.GNN import GNN
    torch.manual_seed(0)
    m = GNN(FA, FB, depth=DEPTH, hidden_sizes=[HID] * DEPTH, dropout_ps=[0.0] * DEPTH, activation_fn=F.relu,
            use_learnable_skip=True)
    m.engine = engine
    return m.to(device)


def build_oracle():
    from oracle.gnn_oracle import OracleGNN
    torch.manual_seed(0)
    return OracleGNN(FA, FB, depth=DEPTH, hidden_sizes=[HID] * DEPTH, dropout_ps=[0.0] * DEPTH,
                     activation_fn=F.relu, use_learnable_skip=True).eval()


def cpu_reference_leg(batch_size: int, steps: int, warmup: int, budget_s: float):
    """The reference's CPU implementation of the path (oracle port of GNN.py) on the host cores."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_oracle()
    batches = [make_batch(batch_size, seed=9000 + i, kind="t1x", fa=FA) for i in range(4)]
    with torch.no_grad():
        for i in range(max(1, min(warmup, 3))):
            model(batches[i % 4])
        t0 = time.perf_counter()
        done = 0
        while done < steps and (time.perf_counter() - t0) < budget_s:
            model(batches[done % 4])
            done += 1
        dt = time.perf_counter() - t0
    return {"value": batch_size * done / dt, "unit": "reactions/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": f"{done} forward passes of one {batch_size}-reaction batch "
                                      f"(oracle/gnn_oracle.py, fp32, {dt:.1f} s)", "ms_per_step": 1e3 * dt / done,
            "steps_done": done}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ---------------------------------------------------------------- reference arm (CPU) ----
    if args.impl == "reference":
        if rank != 0:
            return 0
        leg = cpu_reference_leg(args.batch, args.steps, args.warmup, budget_s=max(30.0, args.cpu_seconds * 10))
        line = {
            "impl": "reference", "metric": "reactions/sec (CGR-MPNN-3D d4 h400 fwd)", "value": leg["value"],
            "unit": "reactions/s", "n_gpus": args.gpus, "steps": leg["steps_done"], "warmup": min(args.warmup, 3),
            "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg-2: CGR-MPNN-3D d{DEPTH} h{HID} learnable-skip forward, batch {args.batch}, "
                                   f"Fa={FA} Fb={FB}, T1x-shaped synthetic reactions", "device": "host CPU"},
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
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    from cgr_mpnn_3d_b200 import _lib
    from cgr_mpnn_3d_b200.collate import plan_for
    lib = _lib.load()
    hbm_peak, _, peak_kind = load_peaks()

    engine = args.engine
    model = build_model(engine, dev).eval()
    model.tile_policy = "throughput" if (args.streams > 1 and not args.no_graph) else "latency"
    lat_model = build_model(engine, dev).eval()      # latency configuration for the single-stream figure
    n_pool = max(2, args.pool)
    host = [make_batch(args.batch, seed=1000 + 997 * rank + i, kind="t1x", fa=FA) for i in range(n_pool)]
    for hb in host:
        hb.y = None
    pool = [hb.to(dev) for hb in host]
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

    # ---- leg 1: device-resident throughput (CUDA-graph replay of the forward, one graph per batch) ----
    outs = [None] * n_pool
    with torch.no_grad():
        for i in range(n_pool):                   # eager pass: builds every plan (tile packing syncs once per batch)
            outs[i] = model(pool[i])
        model.check_numerics()
        torch.cuda.synchronize()
        launches_per_step = 0
        graphs = None
        if not args.no_graph:
            graphs = []
            side = torch.cuda.Stream()
            for i in range(n_pool):
                g = torch.cuda.CUDAGraph()
                c0 = lib.cgr_launch_count()
                with torch.cuda.graph(g, stream=side):
                    outs[i] = model(pool[i])
                launches_per_step = lib.cgr_launch_count() - c0
                graphs.append(g)
        else:
            c0 = lib.cgr_launch_count()
            model(pool[0])
            launches_per_step = lib.cgr_launch_count() - c0

        n_streams = max(1, args.streams) if graphs is not None else 1
        while n_pool % n_streams:          # a graph must always replay on the same stream
            n_streams -= 1
        streams = [torch.cuda.Stream() for _ in range(n_streams)]
        main = torch.cuda.current_stream()

        # One graph per GROUP of n_streams consecutive batches, their forwards on parallel branches (fork / join inside
        # the capture): the same launches as replaying n_streams single-batch graphs on n_streams streams, with one host
        # call instead of n_streams -- with one process per GPU on a shared host the submitting threads are the first
        # thing to saturate.  Consecutive groups alternate between two streams so they overlap at their boundaries.
        group_graphs = []
        if graphs is not None and n_streams > 1 and not args.no_group:
            for g0 in range(0, n_pool, n_streams):
                gg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gg, stream=side):
                    fork_c = torch.cuda.Event()
                    fork_c.record(side)
                    for k in range(n_streams):
                        st = streams[k]
                        st.wait_event(fork_c)
                        with torch.cuda.stream(st):
                            outs[g0 + k] = model(pool[g0 + k])
                        ev_c = torch.cuda.Event()
                        ev_c.record(st)
                        side.wait_event(ev_c)
                group_graphs.append(gg)
        gstreams = [torch.cuda.Stream(), torch.cuda.Stream()]

        def run_steps(count):
            """`count` independent forward passes: batch i % n_pool, n_streams of them in flight."""
            if graphs is None:
                for i in range(count):
                    outs[i % n_pool] = model(pool[i % n_pool])
                return
            fork = torch.cuda.Event()
            fork.record(main)
            used = streams + gstreams
            for st in used:
                st.wait_event(fork)
            i = 0
            if group_graphs:
                n_groups = count // n_streams
                for j in range(n_groups):
                    with torch.cuda.stream(gstreams[j & 1]):
                        group_graphs[j % len(group_graphs)].replay()
                i = n_groups * n_streams
            for k in range(i, count):                 # remainder: single-batch graphs on their own streams
                with torch.cuda.stream(streams[k % n_streams]):
                    graphs[k % n_pool].replay()
            for st in used:
                ev = torch.cuda.Event()
                ev.record(st)
                main.wait_event(ev)

        if graphs is not None:             # initialisation, like the capture itself: instantiate/upload every graph once
            for gph in graphs + group_graphs:
                gph.replay()
            torch.cuda.synchronize()
        clocks = ClockSampler(local_rank).start()
        run_steps(max(3, args.warmup))
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks.mark_begin()
        ev0.record(main)
        run_steps(args.steps)
        ev1.record(main)
        barrier()
        clocks.mark_end()
        ms_total = ev0.elapsed_time(ev1)
        # single-stream latency of one step with the latency kernel configuration (one CTA per SM), for reference
        n_lg = min(8, n_pool)
        lat_graphs = []
        for i in range(n_lg):
            lat_model(pool[i])
        if graphs is not None:
            for i in range(n_lg):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    lat_model(pool[i])
                lat_graphs.append(g)
        lat0, lat1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_lat = min(args.steps, 400)
        for i in range(n_lg):
            lat_graphs[i].replay() if lat_graphs else lat_model(pool[i])
        torch.cuda.synchronize()
        lat0.record(main)
        for i in range(n_lat):
            if lat_graphs:
                lat_graphs[i % n_lg].replay()
            else:
                lat_model(pool[i % n_lg])
        lat1.record(main)
        torch.cuda.synchronize()
        single_stream_ms = lat0.elapsed_time(lat1) / n_lat

    # ---- leg 2: per-stage CUDA-event timing of the same steps (eager, events on the launching stream) ----
    prof_steps = min(args.steps, 40)
    stage_ms, stage_cnt = {}, {}
    with torch.no_grad():
        prof_model = lat_model if n_streams == 1 else model
        for i in range(3):
            prof_model(pool[i % n_pool])
        torch.cuda.synchronize()
        lib.cgr_profile_enable(1)
        for i in range(prof_steps):
            # keep the stream busy while the host enqueues the step, so the kernels (and the events between
            # them) execute back to back: the event pairs then bracket kernel time, not host launch latency
            torch.cuda._sleep(2_000_000)
            prof_model(pool[i % n_pool])
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
    roofline = None
    if dominant:
        avg_ms = stage_ms[dominant] / stage_cnt[dominant]
        per_launch = {
            "gemm_bond_update": 4 * (3 * n_bonds * HID) + 4 * HID * HID,
            "bond_layer": bond_update_bytes(n_atoms, n_bonds),
            "gather_bonds": 4 * (2 * n_bonds * HID),
            "gemm_atom_proj": 4 * (n_atoms * FA + n_atoms * HID + HID * FA),
            "gemm_readout_x": 4 * (n_atoms * FA + n_atoms * HID + HID * FA),
            "atom_proj": 4 * (n_atoms * FA + 2 * n_atoms * HID + 2 * HID * FA),
        }.get(dominant, algorithmic_bytes_fwd(n_atoms, n_bonds, args.batch) / max(1, launches_per_step))
        achieved = per_launch / (avg_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                traffic = json.load(fh).get(dominant, {}).get(f"batch_{args.batch}")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch": per_launch,

                    "peak_kind": peak_kind,
                    "avg_launch_us": avg_ms * 1e3, "launches_per_step": stage_cnt[dominant] / prof_steps,
                    "stage_share": {k: round(v / sum(stage_ms.values()), 4) for k, v in sorted(stage_ms.items())}}

    # ---- leg 3: end to end through the public API with HOST buffers (H2D + forward + D2H per step) ----
    e2e = None
    if not args.skip_e2e:
        pinned = [hb.pin_memory() for hb in host]
        h2d = sum(t.numel() * t.element_size() for t in (pinned[0].x, pinned[0].edge_attr, pinned[0].edge_index,
                                                         pinned[0].batch, pinned[0].ptr))
        with torch.no_grad():
            for i in range(3):
                model(pinned[i % n_pool])
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                res = model(pinned[i % n_pool])      # stages inputs, runs the kernels, copies Ea back to the host
            barrier()
            e2e_single_s = time.perf_counter() - t0
            # the pipelined public API for a stream of host batches (H2D of batch i+1 overlaps batch i's kernels)
            def stream_seconds(coalesce):
                list(model.predict_stream((pinned[i % n_pool] for i in range(32)), depth=4, coalesce=coalesce))
                barrier()
                t0 = time.perf_counter()
                n_done = 0
                for res in model.predict_stream((pinned[i % n_pool] for i in range(args.steps)), depth=4,
                                                coalesce=coalesce):
                    n_done += 1
                barrier()
                assert n_done == args.steps and res.numel() == args.batch
                return time.perf_counter() - t0
            e2e_uncoalesced_s = stream_seconds(1)
            e2e_s = stream_seconds(args.coalesce)
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
        e2e = {"seconds": e2e_s, "single_call_seconds": e2e_single_s, "uncoalesced_seconds": e2e_uncoalesced_s,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(res.numel() * 4), "h2d_peak_gbs": h2d_peak_gbs}

    # ---- optional leg: device-resident reaction store (features stay in HBM, batches assembled by a kernel) ----
    store_leg = None
    if args.store:
        from cgr_mpnn_3d_b200.data import make_reactions
        from cgr_mpnn_3d_b200.store import ReactionStore
        n_store = max(4096, 4 * args.batch)
        store = ReactionStore.from_graphs(make_reactions(n_store, seed=9000 + rank, kind="t1x", fa=FA), device=dev)
        with torch.no_grad():
            for bt in store.loader(args.batch, shuffle=True, seed=0):          # warm-up pass
                model(bt)
            barrier()
            t0 = time.perf_counter()
            n_rx, ep = 0, 0
            while n_rx < args.batch * args.steps:
                for bt in store.loader(args.batch, shuffle=True, seed=1 + ep):
                    out_r = model(bt)
                    n_rx += int(bt.y.numel())
                ep += 1
            barrier()
            store_leg = {"seconds": time.perf_counter() - t0, "reactions": n_rx, "store_bytes": store.nbytes(),
                         "store_reactions": len(store)}
            # the same screening job through the one-call C loop (cgr_store_infer), shuffled order
            import numpy as np
            rng_o = np.random.default_rng(0)
            n_pass = max(1, (args.batch * args.steps) // len(store))
            store.predict(model, batch_size=args.batch, order=rng_o.permutation(len(store)))
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_pass):
                store.predict(model, batch_size=args.batch, order=rng_o.permutation(len(store)))
            barrier()
            store_leg["predict_seconds"] = time.perf_counter() - t0
            store_leg["predict_reactions"] = n_pass * len(store)
        del store

    # ---- optional leg 4: training step (forward + MSE(sum) + explicit backward + gradient SUM all-reduce) ----
    train = None
    if args.train:
        from cgr_mpnn_3d_b200.parallel import allreduce_gradients_
        tm = build_model("auto", dev).train()
        tb = [make_batch(args.batch, seed=5000 + 997 * rank + i, kind="t1x", fa=FA).to(dev) for i in range(8)]
        n_t = min(args.steps, 100)

        def train_step(d):
            loss = torch.nn.functional.mse_loss(tm(d), d.y, reduction="sum")      # train.py:120
            loss.backward()

        # whole-step CUDA graphs (one per resident batch): the step is launch-bound when issued eagerly
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

        def run_train(i):
            tgraphs[i % len(tb)].replay()
            if world > 1:
                allreduce_gradients_(tm.parameters())

        for i in range(5):
            run_train(i)
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for i in range(n_t):
            run_train(i)
        t1e.record()
        barrier()
        train = {"ms_total": t0e.elapsed_time(t1e), "steps": n_t}
        # the same step issued eagerly, as a plain training loop does (trainer.py:139-144 without the optimizer):
        # host-bound, so wall clock between two synchronisations
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
        train["eager_ms"] = (time.perf_counter() - t0) * 1e3 / n_t
        # optimizer step, reported separately (SURVEY.md section 8 d-ii / f-1): one fused launch vs torch's foreach Adam
        from cgr_mpnn_3d_b200.optim import FusedAdam
        for p_ in tm.parameters():
            if p_.grad is None:
                p_.grad = torch.zeros_like(p_)

        def time_opt(opt, n=200):
            for _ in range(5):
                opt.step()
            torch.cuda.synchronize()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record()
            for _ in range(n):
                opt.step()
            o1.record()
            torch.cuda.synchronize()
            return o0.elapsed_time(o1) / n
        l_before = lib.cgr_launch_count()
        train["adam_fused_ms"] = time_opt(FusedAdam(tm.parameters(), lr=1e-3, weight_decay=1e-5, amsgrad=True))
        train["adam_fused_launches"] = (lib.cgr_launch_count() - l_before) / 205
        train["adam_torch_ms"] = time_opt(torch.optim.Adam(tm.parameters(), lr=1e-3, weight_decay=1e-5, amsgrad=True))

        # opt-in: the COMPLETE data-parallel step (forward + loss + backward + gradient SUM + Adam) two ways -- NCCL all-reduce
        # then the one-launch Adam, vs PeerFusedAdam (sum over NVLink peer memory + Adam in one kernel, no NCCL)
        if args.peer_adam and world > 1:
            from cgr_mpnn_3d_b200 import ops as _ops
            from cgr_mpnn_3d_b200.optim import PeerFusedAdam

            def timed_steps(step_fn, n):
                for i in range(6):
                    step_fn(i)
                barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for i in range(n):
                    step_fn(i)
                a1.record()
                barrier()
                return a0.elapsed_time(a1) / n
            fa_opt = FusedAdam(tm.parameters(), lr=1e-4, weight_decay=1e-5, amsgrad=True)

            def nccl_step(i):
                tgraphs[i % len(tb)].replay()
                allreduce_gradients_(tm.parameters())
                fa_opt.step()
            train["dp_step_nccl_ms"] = timed_steps(nccl_step, n_t)
            tm2 = build_model("auto", dev).train()
            popt = PeerFusedAdam(tm2.parameters(), lr=1e-4, weight_decay=1e-5, amsgrad=True)

            def train_step2(d):
                torch.nn.functional.mse_loss(tm2(d), d.y, reduction="sum").backward()
            with torch.cuda.stream(side_t):
                for j, d in enumerate(tb[:2]):
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
            train["dp_step_peer_ms"] = timed_steps(peer_step, n_t)
            _ops.set_grad_arena(None)

    # ---- reduce over ranks (max time), assemble the line ----
    t = torch.tensor([ms_total, e2e["seconds"] if e2e else 0.0, train["ms_total"] if train else 0.0,
                      e2e["single_call_seconds"] if e2e else 0.0, e2e["uncoalesced_seconds"] if e2e else 0.0],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, train_ms, e2e_single_s, e2e_unco_s = (float(v) for v in t)
    total_rxn = args.batch * args.steps * world
    value = total_rxn / (ms_total * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_reference_leg(args.batch, 10 ** 9, 3, args.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    clocks.stop()
    if roofline:
        # Average launch duration of the dominant kernel OVER THE TIMED REGION of `value` = its CUDA-event share of a step
        # (recorded by the library on the launching stream, the stream kept busy so event pairs bracket kernel time) x the
        # timed region's time per step / launches per step -- the machine time the pipelined region spends per launch.
        # Beside it: the same share of the un-pipelined single-stream step (latency of one launch in a dependent chain) and
        # one isolated launch between two events (launch latency and event overhead included).
        share = roofline["stage_share"].get(roofline["kernel"], 0.0)
        lps = max(1e-9, roofline["launches_per_step"])
        per_launch = roofline["algorithmic_bytes_per_launch"]
        roofline["event_bracketed_us"] = roofline.pop("avg_launch_us")
        dur_us = (ms_total / args.steps) * 1e3 * share / lps
        roofline["avg_launch_us"] = dur_us
        roofline["achieved"] = per_launch / (dur_us * 1e-6) / 1e9
        roofline["frac"] = roofline["achieved"] / hbm_peak
        lat_us = single_stream_ms * 1e3 * share / lps
        roofline["single_stream_us_per_launch"] = lat_us
        roofline["single_stream_frac"] = per_launch / (lat_us * 1e-6) / 1e9 / hbm_peak
        roofline["note"] = ("avg_launch_us / achieved / frac: kernel's CUDA-event share of a step x time per step of the timed "
                            "region (value: independent steps pipelined over streams) / launches per step; single_stream_*: "
                            "same share of the un-pipelined CUDA-graph step (latency of a launch in a dependent chain); "
                            "event_bracketed_us: one isolated launch between two events (launch latency included)")
    if rank == 0:
        alg = algorithmic_bytes_fwd(n_atoms, n_bonds, args.batch)
        line = {
            "metric": "reactions/sec (CGR-MPNN-3D d4 h400 fwd)", "value": value, "unit": "reactions/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg-2: CGR-MPNN-3D d{DEPTH} h{HID} learnable-skip forward (inference), batch "
                                   f"{args.batch}/GPU, Fa={FA} Fb={FB}, T1x-shaped synthetic reactions, random-init "
                                   f"weights in the reference .pth layout",
                       "engine": engine, "cuda_graph": not args.no_graph, "streams": n_streams,
                       "graph_grouping": (f"one graph per {n_streams} batches on parallel branches, groups alternate over 2 streams"
                                          if group_graphs else "one graph per batch"),
                       "single_stream_ms_per_step": single_stream_ms,
                       "parallelism": f"replicas x{world}, no collective",
                       "l2": f"inputs rotate over {n_pool} distinct resident batches ({resident / 1e6:.0f} MB > "
                             f"{L2_BYTES / 1e6:.0f} MB L2); weights (5.9 MB) stay resident",
                       "atoms_per_batch": n_atoms, "bonds_per_batch": n_bonds},
            "whole_forward": {"algorithmic_bytes": alg, "achieved_gbs": alg / (ms_total / args.steps * 1e-3) / 1e9,
                              "hbm_frac": alg / (ms_total / args.steps * 1e-3) / 1e9 / hbm_peak,
                              "peak_kind": peak_kind},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks.summary(),
            "gpu_launches": int(launches_per_step) * args.steps,
        }
        if train:
            line["train_step"] = {"value": args.batch * train["steps"] * world / (train_ms * 1e-3), "unit": "reactions/s",
                                  "ms_per_step": train_ms / train["steps"], "steps": train["steps"],
                                  "eager_ms_per_step": train["eager_ms"],
                                  "dp_step_with_optimizer": ({"nccl_allreduce_plus_fused_adam_ms": train["dp_step_nccl_ms"],
                                                              "peer_fused_adam_ms": train["dp_step_peer_ms"],
                                                              "what": "graph-replayed fwd+loss+bwd, then gradient SUM over "
                                                                      "replicas and Adam(amsgrad): NCCL all-reduce + one-launch "
                                                                      "Adam vs ONE kernel over NVLink peer memory (no NCCL)"}
                                                             if "dp_step_peer_ms" in train else None),
                                  "optimizer": {"fused_adam_ms": train["adam_fused_ms"],
                                                "fused_adam_launches_per_step": train["adam_fused_launches"],
                                                "torch_adam_ms": train["adam_torch_ms"],
                                                "what": "Adam(weight_decay, amsgrad=True) step over all parameters, eager "
                                                        "launches, timed apart from the step above (train.py:117-119)"},
                                  "what": "forward + MSE(sum) + explicit backward + flat gradient SUM all-reduce "
                                          "(optimizer excluded), batch %d/GPU, whole step replayed as a CUDA graph" % args.batch}
        if store_leg:
            line["resident_store"] = {"value": store_leg["reactions"] * world / store_leg["seconds"], "unit": "reactions/s",
                                      "store_reactions": store_leg["store_reactions"], "store_bytes": store_leg["store_bytes"],
                                      "predict_value": store_leg["predict_reactions"] * world / store_leg["predict_seconds"],
                                      "predict_api": "ReactionStore.predict(model, batch_size): the per-batch loop in C "
                                                     "(cgr_store_infer), 8 streams, results stay on the device",
                                      "what": "shuffled epochs over a ReactionStore held in HBM: per step one small index "
                                              "upload, the gather kernel, one-launch CSR and the forward (eager launches); "
                                              "no host-to-device copy of features"}
        if e2e:
            line["e2e"] = {"value": total_rxn / e2e_s, "unit": "reactions/s",
                           "api": "GNN.predict_stream(host batches of %d, depth=4, workers=2, coalesce=%d): every step's "
                                  "H2D from its own pinned buffers + index build + kernels + D2H; up to %d consecutive "
                                  "batches share one submission" % (args.batch, args.coalesce, args.coalesce),
                           "uncoalesced_value": total_rxn / e2e_unco_s,
                           "uncoalesced_api": "GNN.predict_stream(..., coalesce=1): one submission per batch",
                           "single_call_value": total_rxn / e2e_single_s, "single_call_api": "GNN.forward(host batch)",
                           "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                           "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                           "h2d_gbs": e2e["h2d_bytes_per_step"] * args.steps / e2e_s / 1e9,
                           "h2d_copy_peak_gbs": e2e["h2d_peak_gbs"],
                           "bound": "host link: h2d_gbs is the per-GPU input traffic the e2e rate implies, "
                                    "h2d_copy_peak_gbs plain pinned 64 MiB cudaMemcpyAsync copies on 4 streams of this box"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


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
