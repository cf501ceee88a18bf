#!/usr/bin/env python
"""bench.py — headline benchmark of the batched self-play MCTS path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): PyRat 7x7 open maze, 10 cheese, 50 turns, `7x7_rust_tuned`
(1897 sims, c_puct 0.512, fpu 0.459, force_k 0.103, batch 16), uniform priors, Dirichlet noise 0,
9472 resident game trees per GPU (two per warp, `--tree-engine half`: 148 SMs x 32 warps x 2; `--tree-engine warp` runs 4736,
one per warp; BASELINE names 4096, `--concurrent 4096`).  One step = one pass of the hot path over one batch of
synthetic games (`--games-per-step` per GPU, default 131072, fresh games every step, played to completion).
Steps are fed continuously (ar_stream_*: three to eight batches in flight, the next batch's blocks take over the
SMs as the previous batch's last long games finish), so throughput does not depend on the batch size.

Metric: self-play MCTS simulations/sec, counted as S_new = descents performed
(nn_evals + terminals); games/hour and the reference's own S_ref (sum of root visits,
selfplay.rs:547) are reported alongside.

  value     kernel-only: every step's inputs are uploaded to its device buffer before its launch is enqueued;
            device clock (CUDA events) from the start of the first timed launch to the end of the last
  e2e       the same K steps through the public streaming C-ABI calls with HOST buffers
            (ar_stream_submit / ar_stream_collect): H2D of games + seeds and D2H of every record inside
            the timed region, wall clock between two device synchronisations
  roofline  algorithmic tree bytes (288 B per node visit + 240 B per new node) of the timed
            launches / their device span, against the measured HBM copy bandwidth
  cpu_baseline  the oracle (restated reference, C++) on all host cores, bounded sample
  nn_configs    BASELINE configs 3 and 4 (MLP / SymmetricMLP / CNN-gpool guided search), short runs, N = 1 only
  config5       BASELINE config 5: 65536 games in total split over the N ranks, records + stats gathered on
                rank 0 over NCCL (device to device) inside the wall time — the strong-scaling point

`--impl reference` times the reference arm: the oracle (the Rust reference cannot be built
here: no cargo, pyrat-rust not vendored) on all host cores for the same workload.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import queue
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(width=7, height=7, cheese_count=10, max_turns=50)
SEARCH = dict(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103,
              noise_epsilon=0.0)
BYTES_PER_NODE_VISIT = 288
BYTES_PER_NEW_NODE = 240


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops_sustained": float(j.get("bf16_tflops_sustained", 1387.3)),
                "kind": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1387.3, "kind": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int) -> None:
        self.device = device
        self.lines: list[str] = []
        self.proc = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self) -> None:
        assert self.proc and self.proc.stdout
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


WORKLOAD_NAME = ("PyRat 7x7 open, 10 cheese, 50 turns, 7x7_rust_tuned (1897 sims, c_puct 0.512, "
                 "fpu 0.459, force_k 0.103, batch 16), uniform priors, noise 0")


def bench_config(args, world: int) -> dict:
    """The `config` of the JSON line: identical for the CUDA arm and the reference arm."""
    return {
        "workload": WORKLOAD_NAME,
        "concurrent_games_per_gpu": args.concurrent, "games_per_step_per_gpu": args.games_per_step,
        "tree_engine": args.tree_engine,
        "parallelism": f"games sharded over {world} GPU(s), no data-path collective",
        "l2": "per-GPU node pools (GBs) exceed the 126 MB L2; fresh games every step",
        "simulations_definition": "S_new = nn_evals + terminals (descents performed)",
        "feed": f"continuous: {n_buffers(args)} batches in flight (ar_stream_*), one tail per run instead of one per step",
    }


def n_buffers(args) -> int:
    """Batches in flight.  A batch lives much longer than its share of the throughput (its last games run alone while
    the next batches fill the machine), so small batches need more of them in flight to keep the GPU fed."""
    return max(3, min(8, (262144 + args.games_per_step - 1) // max(args.games_per_step, 1)))


def make_batch(n: int, first_index: int):
    from alpharat_b200.games import make_games, pods_array

    specs = make_games(n, first_index=first_index, **WORKLOAD)
    seeds = (C.c_uint64 * max(n, 1))(*[first_index + i for i in range(n)])
    return pods_array(specs), seeds


def make_batch_bytes(n: int, first_index: int) -> bytes:
    """Worker-process half of the batch feeder: the pods of one batch as bytes (building 131072 games takes about as
    long in Python as the GPU needs to play them, so the feeder builds several batches in parallel)."""
    pods, _ = make_batch(n, first_index)
    return bytes(memoryview(pods).cast("B"))


def batch_from_bytes(raw: bytes, n: int, first_index: int):
    from alpharat_b200 import _native as N

    pods = (N.GamePod * max(n, 1)).from_buffer_copy(raw)
    seeds = (C.c_uint64 * max(n, 1))(*range(first_index, first_index + n))
    return pods, seeds


def cpu_baseline(seconds: float, threads: int, first_index: int = 10_000_000) -> dict:
    """Oracle self-play on the host cores over the same workload, bounded by wall time.  Chunks of 64 games
    per thread: the CPU arm's own end-of-chunk tail is amortised like the GPU arm's."""
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import load_oracle, oracle_selfplay  # checker, used here as the CPU arm only
    from alpharat_b200.engine import search_cfg

    lib = load_oracle()
    cfg = search_cfg(**SEARCH)
    chunk = 64 * threads
    sims = games = positions = sref = 0
    busy = 0.0
    while busy < seconds:
        pods, seeds = make_batch(chunk, first_index + games)
        t1 = time.perf_counter()
        _, _, _, st = oracle_selfplay(lib, pods, cfg, list(seeds), n_threads=threads)
        busy += time.perf_counter() - t1
        sims += st.total_nn_evals + st.total_terminals
        sref += st.total_simulations
        positions += st.total_positions
        games += chunk
    return {"value": sims / busy, "unit": "simulations/s", "cores": threads, "kind": "port",
            "sample": f"{games} games ({positions} positions) of the bench workload in chunks of {chunk} "
                      f"(64 per thread), {busy:.1f} s",
            "games_per_hour": games / busy * 3600.0, "sref_per_s": sref / busy, "_busy": busy,
            "_sims": sims}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(args.ref_seconds, threads, first_index=20_000_000 + i * 1_000_000)
        if i >= args.warmup:
            per_step.append(r)
    sims = sum(r["_sims"] for r in per_step)
    busy = sum(r["_busy"] for r in per_step)
    v = sims / busy
    line = {
        "impl": "reference", "metric": "self-play MCTS simulations/sec", "value": v, "unit": "simulations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": busy / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, args.gpus),
        "cpu_baseline": {"value": v, "unit": "simulations/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of ~{args.ref_seconds:.0f} s each of the same workload "
                                   "(restated reference, C++ oracle, all host cores, 64 games per thread per chunk)"},
        "e2e": {"value": v, "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# NN-guided configurations (BASELINE configs 3 and 4)
# ---------------------------------------------------------------------------------------------------------
NN_FLOPS = {"mlp": 315_904, "symmetric": 976_896, "cnn": 22.2e6}


def nn_runs():
    sys.path.insert(0, str(ROOT / "tests"))
    from alpharat_b200 import _native as N
    from nn_ref import make_cnn_state_dict, make_mlp_state_dict, make_symmetric_state_dict

    return {
        "mlp": dict(config="3: 7x7_rust_tuned + MLP(hidden 256), 16384 resident trees", arch=N.AR_ARCH_MLP,
                    sd=lambda: make_mlp_state_dict(0, 349), conc=16384, n=16384, sims=1897, fpu=0.459, fk=0.103),
        # config 4 runs are short (2048 games, one per resident tree) to bound the bench time; a run lasts as long as its
        # longest game (the NN-guided loop is bulk-synchronous, one batch per tree and step), so S_new/s grows with the
        # number of games per tree: 4096 games through 4096 trees gave 1.2e7 (SymmetricMLP) / 1.7e7 (CNN), cache off
        "symmetric": dict(config="4: 7x7_rust_strong + SymmetricMLP(hidden 256), 2048 resident trees",
                          arch=N.AR_ARCH_SYMMETRIC, sd=lambda: make_symmetric_state_dict(2, 7, 7), conc=2048, n=2048,
                          sims=2693, fpu=0.479, fk=0.025),
        "cnn": dict(config="4: 7x7_rust_strong + CNN res-res-gpool32 (64 ch), 2048 resident trees",
                    arch=N.AR_ARCH_CNN, sd=lambda: make_cnn_state_dict(3, ("res", "res", "gpool")), conc=2048, n=2048,
                    sims=2693, fpu=0.479, fk=0.025),
    }


def nn_parity(name: str, run: dict, sd, n_pos: int) -> dict:
    """Policy L1 / value error of the bf16 tcgen05 evaluator inside the search against the oracle driven by the
    fp32 restatement of the reference model (tests/nn_ref.py), at the configuration's real simulation count."""
    import numpy as np

    from alpharat_b200 import _native as N
    from alpharat_b200.engine import Engine, search_cfg
    from alpharat_b200.games import make_games, pods_array
    from conftest import EVAL_CB, load_oracle, oracle_search
    from nn_ref import cnn_forward, mlp_forward, symmetric_forward

    fwd = {"mlp": lambda o: mlp_forward(sd, o), "symmetric": lambda o: symmetric_forward(sd, o, 7, 7),
           "cnn": lambda o: cnn_forward(sd, o, 7, 7)}[name]
    lib = load_oracle()
    pods = pods_array(make_games(n_pos, first_index=4242, **WORKLOAD))
    cfg = search_cfg(simulations=run["sims"], batch_size=16, c_puct=0.512, fpu_reduction=run["fpu"], force_k=run["fk"])
    seeds = list(range(n_pos))

    def cb(user, states, n, p1, p2, v1, v2):
        sub = (N.GamePod * n).from_address(C.addressof(states.contents))
        obs = np.zeros((n, 349), np.float32)
        lib.orc_encode(sub, n, obs.ctypes.data_as(C.POINTER(C.c_float)))
        for dst, src in zip((p1, p2, v1, v2), fwd(obs)):
            src = np.ascontiguousarray(src, np.float32)
            C.memmove(dst, src.ctypes.data, src.nbytes)
        return 0

    cbp = EVAL_CB(cb)
    with Engine(concurrent_games=max(n_pos, 4), max_turns=50, max_batch_size=16, max_simulations=run["sims"],
                pool_nodes=run["sims"] + 64) as eng:
        eng.load_weights(run["arch"], 7, 7, sd)
        out = eng.search_batch(pods, cfg, seeds)
    l1, dv = [], []
    for i in range(n_pos):
        rc, ref, _ = oracle_search(lib, pods[i], cfg, seeds[i], eval_cb=cbp)
        if rc != 0:
            continue
        for a, b in ((out[i].policy_p1, ref.policy_p1), (out[i].policy_p2, ref.policy_p2)):
            l1.append(float(np.abs(np.asarray(a[:]) - np.asarray(b[:])).sum()))
        for a, b in ((out[i].value_p1, ref.value_p1), (out[i].value_p2, ref.value_p2)):
            dv.append(abs(a - b))
    q = lambda x, p: float(np.percentile(x, p)) if x else None
    return {"positions": n_pos, "simulations": run["sims"], "reference": "oracle + fp32 numpy restatement of the model",
            "policy_l1": {"median": q(l1, 50), "p90": q(l1, 90), "max": q(l1, 100)},
            "abs_value_err": {"median": q(dv, 50), "p90": q(dv, 90), "max": q(dv, 100)}}


def nn_cpu_arm(run: dict, sd, seconds: float) -> dict:
    """CPU arm of config 3: the oracle's search with the fp32 MLP restatement as its evaluator (numpy, through
    the oracle's predict_fn-style callback: one host thread, the callback holds the GIL)."""
    import numpy as np

    from alpharat_b200 import _native as N
    from alpharat_b200.engine import search_cfg
    from alpharat_b200.games import make_games, pods_array
    from conftest import EVAL_CB, load_oracle, oracle_selfplay
    from nn_ref import mlp_forward

    lib = load_oracle()
    cfg = search_cfg(simulations=run["sims"], batch_size=16, c_puct=0.512, fpu_reduction=run["fpu"], force_k=run["fk"])

    def cb(user, states, n, p1, p2, v1, v2):
        sub = (N.GamePod * n).from_address(C.addressof(states.contents))
        obs = np.zeros((n, 349), np.float32)
        lib.orc_encode(sub, n, obs.ctypes.data_as(C.POINTER(C.c_float)))
        for dst, src in zip((p1, p2, v1, v2), mlp_forward(sd, obs)):
            src = np.ascontiguousarray(src, np.float32)
            C.memmove(dst, src.ctypes.data, src.nbytes)
        return 0

    cbp = EVAL_CB(cb)
    sims = games = 0
    busy = 0.0
    while busy < seconds:
        pods = pods_array(make_games(2, first_index=777_000 + games, **WORKLOAD))
        t1 = time.perf_counter()
        _, _, _, st = oracle_selfplay(lib, pods, cfg, [games, games + 1], n_threads=1, eval_cb=cbp)
        busy += time.perf_counter() - t1
        sims += st.total_nn_evals + st.total_terminals
        games += 2
    return {"value": sims / busy, "unit": "simulations/s", "cores": 1, "kind": "port",
            "sample": f"{games} games, {busy:.1f} s; oracle + numpy fp32 MLP forward through a Python callback "
                      "(one thread). SymmetricMLP / CNN: uniform-only CPU arm (top-level cpu_baseline)"}


def run_nn_configs(args, local: int, peaks: dict) -> list:
    from alpharat_b200.engine import Engine, search_cfg
    from alpharat_b200.games import make_games, pods_array

    out = []
    for name, r in nn_runs().items():
        if args.nn_only and name not in args.nn_only.split(","):
            continue
        sd = r["sd"]()
        pods = pods_array(make_games(r["n"], first_index=50_000_000, **WORKLOAD))
        cfg = search_cfg(simulations=r["sims"], batch_size=16, c_puct=0.512, fpu_reduction=r["fpu"], force_k=r["fk"])
        entry = {"evaluator": name, "config": r["config"], "games": r["n"], "simulations": r["sims"], "runs": []}
        with Engine(device=local, concurrent_games=r["conc"], max_turns=50, max_batch_size=16,
                    max_simulations=r["sims"]) as eng:
            eng.load_weights(r["arch"], 7, 7, sd)
            for cache in (0, 4096):
                eng.set_eval_cache(cache)
                eng.selfplay_upload(pods, list(range(r["n"])))
                st = eng.selfplay_run_resident(cfg)
                summ, _ = eng.selfplay_download(r["n"], 50)
                nn = sum(summ[i].total_nn_evals for i in range(r["n"]))
                term = sum(summ[i].total_terminals for i in range(r["n"]))
                sec = st.device_ms * 1e-3
                rows = int(st.cache_misses) if cache else nn
                tflops = rows / sec * NN_FLOPS[name] / 1e12
                entry["runs"].append({
                    "eval_cache_entries_per_tree": cache, "device_s": sec, "S_new_per_s": (nn + term) / sec,
                    "nn_evals_per_s": nn / sec, "games_per_hour": r["n"] / sec * 3600.0,
                    "evaluator_rows_per_s": rows / sec, "leaf_eval_tflops_in_loop": tflops,
                    "tensor_frac_in_loop": tflops / peaks["bf16_tflops_sustained"],
                    "cache_hit_rate": (st.cache_hits / max(st.cache_hits + st.cache_misses, 1)) if cache else None,
                    "kernel_launches": int(st.kernel_launches)})
        if not args.no_nn_parity:
            entry["parity"] = nn_parity(name, r, sd, {"mlp": 32, "symmetric": 16, "cnn": 2}[name])
        if name == "mlp" and not args.no_cpu:
            entry["cpu_baseline"] = nn_cpu_arm(r, sd, args.nn_cpu_seconds)
        else:
            entry["cpu_baseline"] = "uniform-only (top-level cpu_baseline): no batched CPU evaluator for this model here"
        out.append(entry)
    return out


# ---------------------------------------------------------------------------------------------------------
# The CUDA arm
# ---------------------------------------------------------------------------------------------------------
def run_cuda(args) -> None:
    import torch
    import torch.distributed as dist

    from alpharat_b200 import _native as N
    from alpharat_b200.engine import Engine, search_cfg
    from alpharat_b200.parallel import device_bytes, gather_device, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL announces its version on stdout when it initialises; stdout carries the one JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = search_cfg(**SEARCH)
    n = args.games_per_step
    stride = WORKLOAD["max_turns"]
    eng = Engine(device=local, concurrent_games=args.concurrent, max_turns=stride,
                 max_batch_size=SEARCH["batch_size"], max_simulations=SEARCH["simulations"],
                 tree_engine=args.tree_engine)
    N_BUFFERS = n_buffers(args)
    eng.stream_open(N_BUFFERS, n, stride)

    def first_index(step: int) -> int:  # fresh games every step, disjoint across ranks
        return (step * world + rank) * n

    # Batches are built ahead of the step that plays them by worker processes (building one in Python takes about
    # as long as playing it), a few at a time, so the GPU never waits for the host and host memory does not grow
    # with --steps.
    total_steps = args.warmup + args.steps
    feed: queue.Queue = queue.Queue(maxsize=2)

    def feeder() -> None:
        try:
            with ProcessPoolExecutor(max_workers=args.feed_workers, mp_context=mp.get_context("spawn")) as pool:
                ahead = []
                nxt = 0
                for i in range(total_steps):
                    while nxt < total_steps and len(ahead) < args.feed_workers + 1:
                        ahead.append(pool.submit(make_batch_bytes, n, first_index(nxt)))
                        nxt += 1
                    feed.put(batch_from_bytes(ahead.pop(0).result(), n, first_index(i)))
        except BaseException as exc:  # surfaces in the consumer instead of hanging it
            feed.put(exc)

    threading.Thread(target=feeder, daemon=True).start()

    def next_batch():
        b = feed.get()
        if isinstance(b, BaseException):
            raise b
        return b

    # ---- warm-up: W steps through the same streaming path, drained before the timed region -----------------
    for i in range(args.warmup):
        b = i % N_BUFFERS
        if i >= N_BUFFERS:
            eng.stream_wait(b)
        pods, seeds = next_batch()
        eng.stream_submit(b, pods, None, seeds)
        eng.stream_launch(b, cfg)
    for i in range(max(0, args.warmup - N_BUFFERS), args.warmup):
        eng.stream_wait(i % N_BUFFERS)

    # ---- kernel-only: K steps, continuous feed, device clock ---------------------------------------------------
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    kept = []  # the timed batches are played again by the end-to-end leg
    sims = sref = positions = path_nodes = new_nodes = launches = 0
    t_first, t_last, busy_ms = None, 0.0, 0.0

    def finish(buf: int) -> None:
        nonlocal sims, sref, positions, path_nodes, new_nodes, launches, t_first, t_last, busy_ms
        st = eng.stream_wait(buf)
        a, z = eng.stream_times(buf)
        t_first = a if t_first is None else min(t_first, a)
        t_last = max(t_last, z)
        busy_ms += st.device_ms
        sims += st.total_nn_evals + st.total_terminals
        sref += st.total_simulations
        positions += st.total_positions
        path_nodes += st.path_nodes
        new_nodes += st.new_nodes
        launches += st.kernel_launches

    for i in range(args.steps):
        b = i % N_BUFFERS
        if i >= N_BUFFERS:
            finish(b)
        batch = next_batch()
        if len(kept) < args.steps and len(kept) < args.e2e_keep:
            kept.append(batch)
        eng.stream_submit(b, batch[0], None, batch[1])  # inputs resident before the launch is enqueued
        eng.stream_launch(b, cfg)
        del batch
    for i in range(max(0, args.steps - N_BUFFERS), args.steps):
        finish(i % N_BUFFERS)
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    dev_ms = (t_last - t_first) if t_first is not None else 0.0

    # ---- end to end through the public streaming calls with host buffers ---------------------------------------
    e2e_sims = h2d = d2h = 0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        b = i % N_BUFFERS
        if i >= N_BUFFERS:
            _, _, _, st = eng.stream_collect(b, n, stride)
            e2e_sims += st.total_nn_evals + st.total_terminals
            h2d += st.h2d_bytes
            d2h += st.d2h_bytes
        pods, seeds = kept[i % len(kept)]
        eng.stream_submit(b, pods, cfg, seeds)
    for i in range(max(0, args.steps - N_BUFFERS), args.steps):
        _, _, _, st = eng.stream_collect(i % N_BUFFERS, n, stride)
        e2e_sims += st.total_nn_evals + st.total_terminals
        h2d += st.h2d_bytes
        d2h += st.d2h_bytes
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    eng.stream_close()

    # ---- config 5: a fixed total of 65536 games split over the ranks, gather on rank 0 inside the wall time ----
    c5 = None
    if args.config5_games > 0:
        lo, hi = shard_range(args.config5_games, rank, world)
        m = hi - lo
        pods, seeds = make_batch(m, 900_000_000 + lo)
        barrier()
        t0 = time.perf_counter()
        eng.selfplay_upload(pods, seeds)
        st5 = eng.selfplay_run_resident(cfg)
        t_play = time.perf_counter()
        summ, d_summ, d_rec, n_rec = eng.selfplay_pack_device(m)
        parts_s = gather_device(device_bytes(d_summ, m * C.sizeof(N.GameSummary)))
        parts_r = gather_device(device_bytes(d_rec, n_rec * C.sizeof(N.PositionRecord)))
        gathered = 0
        if parts_s is not None:  # rank 0: the records end in host memory, like the reference's Vec<GameRecord>
            host = [(a.cpu(), b.cpu()) for a, b in zip(parts_s, parts_r)]
            gathered = sum(a.numel() + b.numel() for a, b in host)
        sims5 = sum(summ[g].total_nn_evals + summ[g].total_terminals for g in range(m))
        tot5 = torch.tensor([float(sims5), float(n_rec)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot5)  # the stats vector (SelfPlayStats aggregation, selfplay.rs:212-224)
        barrier()
        t1 = time.perf_counter()
        tv = torch.tensor([t1 - t0, t1 - t_play, st5.device_ms * 1e-3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        wall, gather_s, play_s = tv.tolist()
        c5 = {"total_games": args.config5_games, "games_per_gpu": m, "scaling": "strong",
              "wall_s": wall, "simulations_per_s": tot5[0].item() / wall, "games_per_hour": args.config5_games / wall * 3600,
              "play_device_s_max": play_s, "pack_gather_d2h_s_max": gather_s, "gathered_bytes_rank0": int(gathered),
              "records": int(tot5[1].item()),
              "includes": "H2D of the shard, play, device-side record packing, NCCL gather of summaries + position "
                          "records to rank 0 (device to device), D2H on rank 0, all-reduce of the stats vector",
              "residual": "a fixed total leaves 65536/N games per GPU for the resident trees: the run ends with the "
                          "tail of its longest games (no next batch to overlap), the collective is a few ms"}

    # ---- max over ranks, sum of work -------------------------------------------------------------------------
    vec = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([sims, sref, positions, path_nodes, new_nodes, launches, e2e_sims, h2d, d2h, busy_ms],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = vec.tolist()
    sims, sref, positions, path_nodes, new_nodes, launches, e2e_sims, h2d, d2h, busy_ms = tot.tolist()

    if rank == 0:
        peaks = measured_peaks()
        peak = peaks["hbm_gbs"]
        algo_bytes = BYTES_PER_NODE_VISIT * path_nodes + BYTES_PER_NEW_NODE * new_nodes
        # every rank runs its own launches: per-GPU achieved bandwidth = per-GPU algorithmic bytes / device span
        achieved = algo_bytes / world / (dev_ms * 1e-3) / 1e9
        # `traffic`: DRAM bytes per launch.  The ncu --set full capture is taken on a bounded launch of the same
        # kernel and workload (profiles/traffic_r2.json: measured DRAM bytes and that launch's algorithmic bytes);
        # the ratio is applied to this run's algorithmic bytes per launch.
        traffic = None
        tp = ROOT / "profiles" / "traffic_r2.json"
        if tp.exists():
            tj = json.loads(tp.read_text())
            tj = tj.get(args.tree_engine, tj)  # one capture per tree kernel
            traffic = tj["dram_bytes"] / tj["algorithmic_bytes"] * algo_bytes / max(launches, 1)
        base = cpu_baseline(args.cpu_seconds, os.cpu_count() or 1) if world == 1 and not args.no_cpu else None
        if base:
            base = {k: v for k, v in base.items() if not k.startswith("_")}
        nn_blocks = None
        if world == 1 and not args.no_nn:
            eng.close()
            eng = None
            nn_blocks = run_nn_configs(args, local, peaks)
        line = {
            "metric": "self-play MCTS simulations/sec", "value": sims / (dev_ms * 1e-3),
            "unit": "simulations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, world),
            "games_per_hour": n * args.steps * world / (dev_ms * 1e-3) * 3600.0,
            "sref_per_s": sref / (dev_ms * 1e-3),
            "positions": positions,
            "e2e": {"value": e2e_sims / (e2e_ms * 1e-3), "unit": "simulations/s",
                    "h2d_bytes_per_step": h2d / args.steps / world, "d2h_bytes_per_step": d2h / args.steps / world,
                    "steps": args.steps, "api": "ar_stream_submit / ar_stream_collect (host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_kind": peaks["kind"],
                         "algorithmic_bytes_per_launch": algo_bytes / max(launches, 1),
                         "kernel": "selfplay_half_kernel" if args.tree_engine == "half" else "selfplay_uniform_kernel",
                         "launch_ms_avg": busy_ms / max(launches, 1),
                         "note": "achieved = algorithmic bytes of the timed launches / their device span (launches "
                                 "overlap: the sum of launch durations exceeds the span). The kernel is instruction-issue "
                                 "bound, not HBM bound (profiles/r2_summary.md); traffic = DRAM/algorithmic ratio of the "
                                 "ncu capture x this run's algorithmic bytes per launch"},
            "cpu_baseline": base,
            "clocks": clocks,
            "config5": c5,
            "nn_configs": nn_blocks,
        }
        print(json.dumps(line), flush=True)
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--games-per-step", type=int, default=131072)
    ap.add_argument("--tree-engine", default="half", choices=["half", "warp"],
                    help="uniform-prior tree kernel: two trees per warp (default) or one")
    ap.add_argument("--concurrent", type=int, default=None,
                    help="resident trees per GPU; default 148 SMs x 32 warps x trees per warp = 9472 (half) / 4736 "
                         "(warp); BASELINE names 4096: --concurrent 4096")
    ap.add_argument("--feed-workers", type=int, default=3, help="processes that build the synthetic batches")
    ap.add_argument("--e2e-keep", type=int, default=4, help="distinct batches kept in host memory for the e2e leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=10.0)
    ap.add_argument("--config5-games", type=int, default=65536)
    ap.add_argument("--nn-cpu-seconds", type=float, default=6.0)
    ap.add_argument("--nn-only", default="")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-nn", action="store_true")
    ap.add_argument("--no-nn-parity", action="store_true")
    args = ap.parse_args()
    if args.concurrent is None:
        args.concurrent = 9472 if args.tree_engine == "half" else 4736
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
