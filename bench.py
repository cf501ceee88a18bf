#!/usr/bin/env python
"""bench.py — headline benchmark of the batched self-play MCTS path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): PyRat 7x7 open maze, 10 cheese, 50 turns, `7x7_rust_tuned`
(1897 sims, c_puct 0.512, fpu 0.459, force_k 0.103, batch 16), uniform priors, Dirichlet noise 0,
4096 concurrent game trees per GPU.  One step = one pass of the hot path over one batch of
synthetic games (`--games-per-step` per GPU, default 131072, played to completion through the
4096 resident trees; fresh games every step).  A step ends with a tail in which the last, longest games
run alone (games last 12.5 turns on average, up to 50), so throughput grows with the batch: 3.25 / 3.93 /
4.54 / 4.80 x 10^8 simulations/s at 16k / 32k / 64k / 128k games per step (profiles/r1_summary.md).

Metric: self-play MCTS simulations/sec, counted as S_new = descents performed
(nn_evals + terminals); games/hour and the reference's own S_ref (sum of root visits,
selfplay.rs:547) are reported alongside.

  value     kernel-only: inputs resident in HBM, CUDA events on the engine's stream
  e2e       the same metric through the public API (`ar_selfplay_run`): host buffers, H2D of the
            games + seeds and D2H of every record inside the timed region
  roofline  algorithmic tree bytes (288 B per node visit + 240 B per new node)
            per launch / launch duration, against the measured HBM copy bandwidth
  cpu_baseline  the oracle (restated reference, C++) on all host cores, bounded sample

`--impl reference` times the reference arm: the oracle (the Rust reference cannot be built
here: no cargo, pyrat-rust not vendored) on all host cores for the same workload.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import queue
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(width=7, height=7, cheese_count=10, max_turns=50)
SEARCH = dict(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103,
              noise_epsilon=0.0)
BYTES_PER_NODE_VISIT = 288
BYTES_PER_NEW_NODE = 240


def measured_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


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
        "parallelism": f"games sharded over {world} GPU(s), no data-path collective",
        "l2": "per-GPU node pools (GBs) exceed the 126 MB L2; fresh games every step",
        "simulations_definition": "S_new = nn_evals + terminals (descents performed)",
    }


def make_batch(n: int, first_index: int):
    from alpharat_b200.games import make_games, pods_array

    specs = make_games(n, first_index=first_index, **WORKLOAD)
    seeds = [first_index + i for i in range(n)]
    return pods_array(specs), seeds


def cpu_baseline(seconds: float, threads: int, first_index: int = 10_000_000) -> dict:
    """Oracle self-play on the host cores over the same workload, bounded by wall time."""
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import load_oracle, oracle_selfplay  # checker, used here as the CPU arm only
    from alpharat_b200.engine import search_cfg

    lib = load_oracle()
    cfg = search_cfg(**SEARCH)
    chunk = max(64, 16 * threads)
    sims = games = positions = sref = 0
    busy = 0.0
    while busy < seconds:
        pods, seeds = make_batch(chunk, first_index + games)
        t1 = time.perf_counter()
        _, _, _, st = oracle_selfplay(lib, pods, cfg, seeds, n_threads=threads)
        busy += time.perf_counter() - t1
        sims += st.total_nn_evals + st.total_terminals
        sref += st.total_simulations
        positions += st.total_positions
        games += chunk
    return {"value": sims / busy, "unit": "simulations/s", "cores": threads, "kind": "port",
            "sample": f"{games} games ({positions} positions) of the bench workload, {busy:.1f} s",
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
                                   "(restated reference, C++ oracle, all host cores)"},
        "e2e": {"value": v, "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_cuda(args) -> None:
    import torch
    import torch.distributed as dist

    from alpharat_b200 import _native as N
    from alpharat_b200.engine import Engine, search_cfg

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
    eng = Engine(device=local, concurrent_games=args.concurrent, max_turns=WORKLOAD["max_turns"],
                 max_batch_size=SEARCH["batch_size"], max_simulations=SEARCH["simulations"])
    stride = WORKLOAD["max_turns"]

    def first_index(step: int) -> int:  # fresh games every step, disjoint across ranks
        return (step * world + rank) * n

    # ---- kernel-only: inputs resident, CUDA-event device time -----------------------------
    # Batches are built by a feeder thread, two ahead of the step that plays them (building one takes about as
    # long as playing it; the engine call releases the GIL, and the timed quantity is device time), so host memory
    # and start-up time do not grow with --steps.
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    feed: queue.Queue = queue.Queue(maxsize=2)

    def feeder() -> None:
        try:
            for i in range(args.warmup + args.steps):
                feed.put(make_batch(n, first_index(i)))
        except BaseException as exc:  # surfaces in the consumer instead of hanging it
            feed.put(exc)

    threading.Thread(target=feeder, daemon=True).start()

    def next_batch():
        b = feed.get()
        if isinstance(b, BaseException):
            raise b
        return b

    for i in range(args.warmup):
        eng.selfplay_upload(*next_batch())
        eng.selfplay_run_resident(cfg)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    dev_ms = 0.0
    sims = sref = positions = path_nodes = new_nodes = launches = 0
    kept = []  # the first timed batches are played again by the end-to-end leg
    for i in range(args.warmup, args.warmup + args.steps):
        batch = next_batch()
        if len(kept) < e2e_steps:
            kept.append(batch)
        eng.selfplay_upload(*batch)
        del batch
        st = eng.selfplay_run_resident(cfg)
        dev_ms += st.device_ms
        path_nodes += st.path_nodes
        new_nodes += st.new_nodes
        launches += st.kernel_launches
        summ, _ = eng.selfplay_download(n, stride)
        for g in range(n):
            sims += summ[g].total_nn_evals + summ[g].total_terminals
            sref += summ[g].total_simulations
            positions += summ[g].n_positions
    barrier()
    clocks = sampler.stop() if rank == 0 else {}

    # ---- end to end through the public C-ABI call with host buffers -------------------------
    e2e_ms = 0.0
    e2e_sims = h2d = d2h = 0
    e2e_steps = min(e2e_steps, max(len(kept), 1))
    for i in range(min(e2e_steps, len(kept))):
        pods, seeds = kept[i]
        barrier()
        t0 = time.perf_counter()
        summ, pos, _, st = eng.selfplay(pods, cfg, seeds, stride=stride)
        torch.cuda.synchronize()
        e2e_ms += (time.perf_counter() - t0) * 1e3
        e2e_sims += st.total_nn_evals + st.total_terminals
        h2d += st.h2d_bytes
        d2h += st.d2h_bytes

    # ---- max over ranks, sum of work ---------------------------------------------------------
    vec = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([sims, sref, positions, path_nodes, new_nodes, launches, e2e_sims, h2d, d2h],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        # NCCL gather of the per-game result summaries (recorded batches + stats), as in config 5
        summ_t = torch.frombuffer(bytearray(bytes(summ)), dtype=torch.uint8).cuda()
        gathered = [torch.empty_like(summ_t) for _ in range(world)] if rank == 0 else None
        dist.gather(summ_t, gathered, dst=0)
    dev_ms, e2e_ms = vec.tolist()
    sims, sref, positions, path_nodes, new_nodes, launches, e2e_sims, h2d, d2h = tot.tolist()

    if rank == 0:
        peak, peak_kind = measured_peaks()
        algo_bytes = BYTES_PER_NODE_VISIT * path_nodes + BYTES_PER_NEW_NODE * new_nodes
        # every rank runs its own kernel: per-GPU achieved bandwidth = per-GPU bytes / time
        achieved = algo_bytes / world / (dev_ms * 1e-3) / 1e9
        # `traffic`: DRAM bytes of one launch.  The ncu --set full capture is taken on a bounded launch
        # of the same kernel and workload (profiles/traffic_r1.json: measured DRAM bytes and that
        # launch's algorithmic bytes); the ratio is applied to this run's algorithmic bytes per launch.
        traffic = None
        tp = ROOT / "profiles" / "traffic_r1.json"
        if tp.exists():
            tj = json.loads(tp.read_text())
            traffic = tj["dram_bytes"] / tj["algorithmic_bytes"] * algo_bytes / max(launches, 1)
        base = cpu_baseline(args.cpu_seconds, os.cpu_count() or 1) if world == 1 and not args.no_cpu else None
        if base:
            base = {k: v for k, v in base.items() if not k.startswith("_")}
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
                    "h2d_bytes_per_step": h2d / e2e_steps / world, "d2h_bytes_per_step": d2h / e2e_steps / world,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": algo_bytes / max(launches, 1),
                         "note": "tree kernel is instruction-issue bound (68 % of issue slots, DRAM < 2 % of peak); "
                                 "traffic = DRAM/algorithmic ratio of the ncu capture x this run's algorithmic bytes; "
                                 "see profiles/r1_summary.md"},
            "cpu_baseline": base,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
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
    ap.add_argument("--concurrent", type=int, default=4096)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
