#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Harmonies engine (contract: one JSON line).

Workload (BASELINE.json configs[1]): batched uniform-random playouts of 65,536 concurrent
2-player games per GPU, engine only.  One "step" = one wave: every game is played from a
fresh initial state to the end by the fused playout kernel (hz_playout).  `value` = engine
steps/s (1 step = one apply_move-equivalent action on one game) over all ranks, with the
initial states resident in HBM; `e2e` = the same metric through the C-ABI with HOST buffers
(pinned H2D of the initial states and D2H of the final states inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--games G]
    torchrun ... bench.py --gpus N ...        (one rank per GPU, games sharded by rank)
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "engine_steps_per_sec"
UNIT = "steps/s"
ALGO_BYTES_PER_STEP = 258  # 128 B state read + 128 B write + 2 B action


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=65536, help="concurrent games per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mcts", action="store_true", help="skip the secondary MCTS sims/s measurement")
    ap.add_argument("--mcts-games", type=int, default=4096, help="concurrent games (trees) per GPU")
    ap.add_argument("--mcts-sims", type=int, default=100)
    ap.add_argument("--mcts-streams", type=int, default=1, help="slot groups on separate CUDA streams")
    ap.add_argument("--mcts-leaves", type=int, default=1, help="simulations in flight per tree and step (1 = reference-exact; >1 = virtual loss)")
    ap.add_argument("--mcts-steps", type=int, default=5, help="timed moves (each = games x sims simulations)")
    ap.add_argument("--mcts-no-graph", action="store_true", help="launch the simulation steps eagerly instead of replaying a CUDA graph")
    ap.add_argument("--mcts-tower", default="hand", choices=["hand", "cudnn"], help="residual tower: the hand-written sm_100a kernel or cuDNN (A/B)")
    ap.add_argument("--mcts-play-games", type=int, default=8192, help="whole self-play games per GPU for the configs[3] measurement (0 = skip)")
    ap.add_argument("--no-python-reference", action="store_true", help="skip timing the staged Python reference (baseline/_ref)")
    return ap.parse_args()


def kernel_counters(name, files):
    """ncu-derived per-launch counters (profiles/regen.sh), accepted only if the kernel's sources still
    hash to what was profiled."""
    import hashlib

    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None, f"{name} missing (run profiles/regen.sh on a B200)"
    d = json.load(open(p))
    h = hashlib.sha256()
    for f in sorted(os.path.join(ROOT, "harmonies_alphazero_b200", "csrc", x) for x in files):
        h.update(open(f, "rb").read())
    if h.hexdigest() != d.get("source_sha256"):
        return None, f"{name} is stale: the kernel sources changed since it was captured (run profiles/regen.sh)"
    return d, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval=0.0002):
        # interval: seconds between NVML polls.  The engine leg lasts milliseconds and is polled as fast as NVML answers;
        # the seconds-long MCTS legs are polled every 20 ms (NVML calls take driver locks that kernel launches need too:
        # eight ranks polling at kHz rates slowed each other's whole-game loops)
        self.interval = interval
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.samples, self.stop_flag = None, [], False
        try:  # in-process NVML polls every ~1 ms: the timed region of this bench is milliseconds long
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        R = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, [k for k, bit in R.items() if rs & bit]))
            except Exception:
                break
            time.sleep(self.interval)

    def start(self):
        if self.nvml:
            self.t0 = time.perf_counter()
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.th.join(timeout=1)
            nv = self.nvml
            sm = [s[1] for s in self.samples]
            reasons = sorted({r for s in self.samples for r in s[2]})
            try:
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            except Exception:
                mx = None
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                    "samples": len(sm), "source": f"nvml, polled every {self.interval * 1e3:g} ms (plus the call) during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def headline_roofline(n_games, steps_per_launch, launch_s, clocks, pk, pk_src, at4=None):
    """hz::k_playout keeps a game's state in registers for all of its ~62 steps, so DRAM sees one 128-byte
    record per GAME and the kernel is bound by instruction issue, not HBM.  The roofline is therefore the
    issue roofline (warp instructions per second against SMs x 4 schedulers x clock), from the warp-instruction
    count ncu measured for exactly these sources (profiles/regen.sh) and the launch time measured live here.
    The contract's HBM figure (258 algorithmic bytes per step) is kept beside it as `hbm_algorithmic`."""
    algo = ALGO_BYTES_PER_STEP * steps_per_launch / launch_s / 1e9
    hbm = {"achieved": algo, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": algo / pk["hbm_gbs"], "peak_source": pk_src + " (burst copy)",
           "note": "258 B/step x steps per launch / launch time (SURVEY.md 8d): what an unfused legal+apply loop would move; "
                   "above 1 because the fused kernel does not move it (see traffic), NOT because work is skipped — every final "
                   "record equals the CPU oracle's bit for bit (tests/test_gpu_engine.py)"}
    c, why = kernel_counters("r02_playout_counters.json", ["hz_engine.cu", "hz_core.cuh", "hz_tables.inc"])
    if c is None or n_games != 65536:
        return {"bound": "issue", "achieved": None, "peak": None, "unit": "Gwarp-inst/s", "frac": None, "traffic": None,
                "kernel": "hz::k_playout", "note": why or "counters were captured at 65,536 games per launch", "hbm_algorithmic": hbm}
    m = {k: v["value"] for k, v in c["metrics"].items()}
    winst = m["smsp__inst_executed.sum"]
    sm_mhz = (clocks or {}).get("sm_mhz") or pk.get("sm_max_mhz") or 1965.0
    peak = 148 * 4 * sm_mhz * 1e6 / 1e9
    ach = winst / launch_s / 1e9
    if at4:
        # warp instructions scale with the games (same kernel, same game-length distribution): 4x the counted launch
        a4 = 4.0 * winst / (at4["us_per_launch"] * 1e-6) / 1e9
        at4 = dict(at4, achieved=a4, frac=a4 / peak,
                   note="the same kernel with 4x the games per launch (14 warps per scheduler instead of 3.5); warp instructions "
                        "taken as 4x the counted 65,536-game launch")
    return {"bound": "issue", "achieved": ach, "peak": peak, "unit": "Gwarp-inst/s", "frac": ach / peak, "at_4x_batch": at4,
            "traffic": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
            "kernel": "hz::k_playout", "peak_source": f"148 SMs x 4 schedulers x {sm_mhz:.0f} MHz (SM clock sampled during the timed region)",
            "warp_inst_per_launch": winst, "active_lanes_per_warp_inst": m["smsp__thread_inst_executed.sum"] / winst,
            "thread_inst_per_engine_step": m["smsp__thread_inst_executed.sum"] / steps_per_launch,
            "frac_meaning": "issue-slot utilisation of the instructions the kernel executes: it falls when instructions are removed "
                            "at constant latency.  The round-1 kernel executed 338 thread instructions per engine step at frac 0.46 "
                            "(95 us per wave); thread_inst_per_engine_step and the launch time say what this build does (DESIGN.md §8)",
            "ncu": {"duration_us": m.get("gpu__time_duration.sum", 0) / 1e3, "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "warps_active_pct": m.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "alu_pipe_pct": m.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                    "fma_pipe_pct": m.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                    "source": "profiles/r02_playout_counters.json"},
            "note": "achieved = warp instructions per launch (ncu, same sources by sha256) / CUDA-event launch time measured in this run",
            "hbm_algorithmic": hbm}


# --------------------------------------------------------------------------------------
def cpu_port_run(n_games, seed, threads):
    """the oracle (C port of the reference's engine) playing n_games random playouts"""
    from oracle import oracle as orc

    init = orc.init_states(n_games, seed=seed)
    t0 = time.perf_counter()
    _, _, total = orc.playout(init, n_threads=threads)
    return total, time.perf_counter() - t0


def cpu_baseline(budget_s=12.0):
    threads = os.cpu_count() or 1
    total, dt = cpu_port_run(2000 * threads, 1, threads)          # calibration
    rate = total / dt
    n_games = int(max(2000 * threads, min(4_000_000, rate * budget_s / 62.0)))
    total, dt = cpu_port_run(n_games, 2, threads)
    return {"value": total / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_games} random-playout games ({total} steps, {dt:.1f} s) of the same workload by "
                      f"oracle/hz_oracle.c on {threads} pthreads; the reference itself is pure Python "
                      f"(~8.3k steps/s/core, BASELINE.md §2)"}


def python_reference(args):
    """The unmodified Python reference (baseline/_ref, staged by baseline/stage_ref.py) on this box's host
    cores: SURVEY.md §8(d) "CPU baseline timing"."""
    try:
        from baseline import ref_timing as rt

        if not rt.available():
            from baseline import stage_ref

            stage_ref.stage()              # works where /root/reference exists (the dev container)
        return rt.measure(engine_budget_s=4.0, moves=2, sims=args.mcts_sims)
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the Python reference cannot
    travel to the GPU box) on all host threads, same metric/config, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import oracle as orc

    orc.build()
    total, dt = cpu_port_run(1000 * threads, 1, threads)
    n_games = int(max(1000 * threads, min(args.games, (total / dt) * 3.0 / 62.0)))   # ~3 s per step
    for w in range(args.warmup):
        cpu_port_run(n_games, 100 + w, threads)
    steps_total, t_total = 0, 0.0
    for k in range(args.steps):
        s, dt = cpu_port_run(n_games, 200 + k, threads)
        steps_total += s
        t_total += dt
    v = steps_total / t_total
    sample = f"{n_games} random-playout games per step on {threads} pthreads (oracle/hz_oracle.c, C port of harmonies_engine.py)"
    # the sims/s half of the metric: the C port's tree-only rate and the Python reference's own self-play layout
    mcts_ref = {"port": cpu_mcts_baseline(budget_s=6.0)}
    pyref = None if args.no_python_reference else python_reference(args)
    if pyref is not None:
        mcts_ref["python_reference"] = pyref
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "random playouts, 65,536 concurrent 2-player games per GPU, engine only (configs[1])",
                   "games_per_step": n_games},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "python_reference": (pyref or {}).get("engine")},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mcts": {"metric": "mcts_sims_per_sec", "unit": "sims/s",
                 "value": ((pyref or {}).get("mcts_loggers_on") or {}).get("value") or mcts_ref["port"]["with_network_estimate"]["value"],
                 "kind": "reference" if (pyref or {}).get("mcts_loggers_on") else "port estimate",
                 "detail": mcts_ref},
    }))


# --------------------------------------------------------------------------------------
def mcts_measure(args, dev, world, rank, dist, with_collectives=True, tower=None, play_games=0):
    """configs[3]: MCTS self-play with the model.py net (random-init, bf16), 100 sims/move,
    4,096 concurrent games per GPU.  One step = one move of every game = 4,096 x 100
    simulations (select -> network forward -> expand+backup, CUDA-graph replayed)."""
    import torch

    from harmonies_alphazero_b200 import batched as hb
    from harmonies_alphazero_b200 import net as hznet
    from harmonies_alphazero_b200 import selfplay as sp

    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval()
    inf = hznet.InferenceNet(model, device=dev, dtype=torch.bfloat16, tower=tower or args.mcts_tower)
    B, S = args.mcts_games, args.mcts_sims
    cfg = sp.SelfPlayConfig(n_slots=B, num_simulations=S, cpuct=2.0, dirichlet_alpha=0.4, dirichlet_epsilon=0.25,
                            turns_until_tau0=15, seed=77, first_game_id=rank * B, n_streams=args.mcts_streams, leaves_per_step=args.mcts_leaves,
                            use_cuda_graph=not args.mcts_no_graph)
    drv = sp.BatchedSelfPlay(inf, cfg, device=dev)
    states = hb.init_states(B, device=dev, seed=77, first_id=rank * B)
    hb.playout(states, max_steps=8)                 # a few moves in: realistic branching
    u01 = torch.rand(B, device=dev)

    def one_move():
        drv.search(states)
        hb.apply(states, drv.choose(u01, None))

    one_move()                                      # warm-up: cuDNN autotune + graph capture
    one_move()
    K = args.mcts_steps
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = hb.launch_count()
    sampler = ClockSampler(dev.index if dev.index is not None else 0, interval=0.02)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        one_move()
    e1.record()
    torch.cuda.synchronize()
    mcts_clocks = sampler.stop()
    if dist is not None:
        dist.barrier()
    direct_launches = hb.launch_count() - l0
    drv.check_status()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    collectives = None
    if dist is not None and with_collectives:
        # configs[4]: NCCL weight broadcast + packed trajectory gather, once per training step
        from harmonies_alphazero_b200 import dist as hzdist

        model_dev = model.to(dev)
        n_ex = B * 62 // 8                        # examples a rank produces between two training steps (B/8 games)
        traj = sp.Trajectories(
            states=states[:1].repeat(n_ex, 1), visits=torch.zeros((n_ex, 143), dtype=torch.int16, device=dev),
            z=torch.zeros(n_ex, device=dev), game_id=torch.arange(n_ex, device=dev, dtype=torch.int64),
            move_no=torch.zeros(n_ex, dtype=torch.int32, device=dev))
        for _ in range(2):
            hzdist.broadcast_weights(model_dev, src=0, dtype=torch.bfloat16)
            hzdist.gather_trajectories(traj, dst=0)
        torch.cuda.synchronize(); dist.barrier()
        c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        c0.record()
        wbytes = hzdist.broadcast_weights(model_dev, src=0, dtype=torch.bfloat16)
        c1.record()
        g = hzdist.gather_trajectories(traj, dst=0)
        c2.record()
        torch.cuda.synchronize()
        tms = torch.tensor([c0.elapsed_time(c1), c1.elapsed_time(c2)], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        collectives = {"weight_broadcast_ms": float(tms[0]), "weight_bytes": wbytes,
                       "trajectory_gather_ms": float(tms[1]), "examples_per_rank": n_ex,
                       "bytes_per_example": g.stats.get("bytes_per_example"), "backend": "nccl"}
    sims = B * S * K * world
    v_step = sims / (ms * 1e-3)
    pk, pk_src = peaks()
    flops = hznet.flops_per_position()
    hand = inf.hand is not None
    # select, expand + (queue init, tower [with the head-convolution items], FC heads [+ the separate head-conv kernel]) / fused heads
    own_per_sim = (2 + ((3 if getattr(inf, "heads_in_tower", False) else 4) if hand else (inf.heads is not None))) if drv.graph is not None else 0
    peak_step = {"value": v_step, "unit": "sims/s", "ms_per_move": ms / K, "moves": K,
                 "what": f"{K} moves of all {B} games from ply 8 (every slot live): the step rate of the search itself"}
    whole = None
    if play_games > 0:
        # configs[3] as specified: whole self-play games (trainer.py:468-509), finished games replaced in place
        pcfg = sp.SelfPlayConfig(n_slots=B, num_simulations=S, cpuct=2.0, dirichlet_alpha=0.4, dirichlet_epsilon=0.25,
                                 turns_until_tau0=15, seed=78, first_game_id=rank * play_games, n_streams=args.mcts_streams,
                                 leaves_per_step=args.mcts_leaves, use_cuda_graph=not args.mcts_no_graph)
        drv.cfg = pcfg
        # warm-up of the whole-game loop itself: a tiny driver plays more games than it has slots, so every torch kernel of the
        # refill / compaction / bookkeeping path is loaded before the timed region (first uses cost up to 100 ms each)
        wcfg = sp.SelfPlayConfig(n_slots=32, num_simulations=4, seed=5, n_streams=args.mcts_streams, use_cuda_graph=not args.mcts_no_graph)
        sp.BatchedSelfPlay(inf, wcfg, device=dev).play(80)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        l1 = hb.launch_count()
        sampler2 = ClockSampler(dev.index if dev.index is not None else 0, interval=0.02)
        sampler2.start()
        traj = drv.play(play_games)
        torch.cuda.synchronize()
        play_clocks = sampler2.stop()
        st = traj.stats
        tt = torch.tensor([st["seconds"]], dtype=torch.float64, device=dev)
        cnt = torch.tensor([st["sims"], st["games"], len(traj), st["move_steps"]], dtype=torch.int64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        secs = float(tt.item())
        live_sims, games, examples, move_steps = (int(x) for x in cnt.tolist())
        whole = {"value": live_sims / secs, "unit": "sims/s", "games_per_s": games / secs, "examples_per_s": examples / secs,
                 "games": games, "seconds": secs, "move_steps_per_gpu": move_steps // world,
                 "slot_utilisation": live_sims / (S * B * max(1, move_steps)),
                 "vs_peak_step": (live_sims / secs) / v_step, "clocks": play_clocks,
                 "launches_direct": (hb.launch_count() - l1) * world,
                 "what": f"BatchedSelfPlay.play: {play_games} complete games per GPU on {B} slots (refill in place), live simulations only; "
                         "includes refill, trajectory bookkeeping, move sampling and the game tail; live games are kept in a dense slot prefix "
                         "and only those are searched (SelfPlayConfig.compact_live)",
                 "searched_slot_steps": int(st.get("searched_slots", 0))}
        direct_launches += hb.launch_count() - l1
    v = whole["value"] if whole else v_step
    ach = (v / world) * flops / 1e12
    n_steps_graph = S * K + (S * (whole["move_steps_per_gpu"] if whole else 0))
    out = {"metric": "mcts_sims_per_sec", "value": v, "unit": "sims/s", "ms_per_move": ms / K,
           "config": {"workload": "MCTS self-play, model.py net 128f x 8 blocks random-init, bf16, "
                                  f"{S} sims/move, {B} concurrent games per GPU (configs[3])"
                                  + (f", {play_games} whole games per GPU" if whole else ", fixed positions"),
                      "tower": "hand-written sm_100a tcgen05 (csrc/hz_tower.cu)" if hand else "cuDNN",
                      "fused_conv": inf.fused, "fused_heads": inf.heads is not None, "cuda_graph": drv.graph is not None, "streams": drv.n_groups, "leaves_per_step": args.mcts_leaves},
           "dtype": "bf16", "collectives": collectives, "clocks": mcts_clocks, "peak_step": peak_step,
           # direct C-ABI launches + this library's kernels replayed inside the CUDA graph per simulation step
           "gpu_launches_own": (direct_launches + own_per_sim * drv.n_groups * n_steps_graph) * world,
           "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / pk["bf16_tflops_sustained"], "traffic": None,
                        "peak_source": pk_src + " (sustained cuBLAS bf16)",
                        "peak_step_frac": (v_step / world) * flops / 1e12 / pk["bf16_tflops_sustained"],
                        "note": f"{flops / 1e6:.2f} MFLOP per simulation (the network forward as the reference defines it, zero padding "
                                "included) x sims/s per GPU" + ("; the hand-written tower executes 247/315 of the convolution MACs (taps that "
                                "fall off the 5x7 board are skipped)" if hand else "")}}
    if whole:
        out["whole_games"] = whole
    return out


def cpu_mcts_baseline(budget_s=10.0):
    """the oracle's MCTS (C port of MCTS.py) with the synthetic evaluator on all host threads:
    tree work only, no network — an upper bound for the CPU side."""
    from oracle import oracle as orc
    import numpy as np

    threads = os.cpu_count() or 1
    roots = orc.init_states(8 * threads, seed=5)
    keys = np.arange(len(roots), dtype=np.uint64)
    t0 = time.perf_counter()
    orc.search_batch(roots, keys, 100, 2.0, n_threads=threads)
    dt = time.perf_counter() - t0
    n = int(max(len(roots), min(200000, len(roots) * budget_s / max(dt, 1e-3))))
    roots = orc.init_states(n, seed=6)
    keys = np.arange(n, dtype=np.uint64)
    t0 = time.perf_counter()
    orc.search_batch(roots, keys, 100, 2.0, n_threads=threads)
    dt = time.perf_counter() - t0
    tree_rate = n * 100 / dt
    # the reference evaluates its network on the CPU at batch 1 once per simulation
    # (MCTS.py:302, trainer.py:104-107: one model per worker process, one thread each): time
    # that forward here and combine it with the tree-only rate into a whole-simulation estimate
    import torch

    from harmonies_alphazero_b200 import net as hznet

    old_threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        torch.manual_seed(0)
        m = hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval()
        b, g = torch.zeros(1, 38, 5, 7), torch.zeros(1, 42)
        with torch.no_grad():
            for _ in range(3):
                m(b, g)
            t0 = time.perf_counter()
            for _ in range(20):
                m(b, g)
            net_s = (time.perf_counter() - t0) / 20
    finally:
        torch.set_num_threads(old_threads)
    per_core_tree_s = threads / tree_rate
    with_net = threads / (per_core_tree_s + net_s)
    return {"value": tree_rate, "unit": "sims/s", "cores": threads, "kind": "port",
            "with_network_estimate": {"value": with_net, "unit": "sims/s", "net_forward_ms_batch1_1thread": net_s * 1e3,
                                      "how": "cores / (tree seconds per sim per core + fp32 forward at batch 1 on one thread), "
                                             "the reference's worker layout"},
            "sample": f"{n} searches x 100 sims, synthetic evaluator (no network), oracle/hz_oracle.c on {threads} pthreads; "
                      "the Python reference with its net on CPU does ~75 sims/s/core (BASELINE.md §2)"}


def run_b200(args):
    import torch
    import torch.distributed as dist

    from harmonies_alphazero_b200 import batched as hb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # harness convenience: (re)build a missing/stale library once per node; the ops themselves never do
    from harmonies_alphazero_b200 import build as hz_build

    if hz_build.needs_build():
        if local == 0:
            hz_build.build()
        else:
            for _ in range(240):
                if not hz_build.needs_build():
                    break
                time.sleep(0.5)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    if world > 1:
        # one slice of the host's cores per rank: the ranks' launch threads and NCCL proxies do not migrate onto each other
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except (AttributeError, OSError):
            pass
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.games
    K, W = args.steps, args.warmup
    stream = torch.cuda.current_stream()
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    steps_buf = torch.empty(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    pinned_in = torch.empty((n, 32), dtype=torch.int32).pin_memory()
    pinned_out = torch.empty((n, 32), dtype=torch.int32).pin_memory()
    states = torch.empty((n, 32), dtype=torch.int32, device=dev)

    def fresh(step):
        # seed differs per step and rank: every wave plays new games
        return hb.init_states(n, device=dev, seed=1000 + step, first_id=rank * n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks over the WHOLE engine leg (warm-up, timed waves, end-to-end legs, unfused comparison: tens of milliseconds,
    # i.e. more than the one or two NVML samples that fit in the 2 ms timed region itself)
    leg_sampler = ClockSampler(local, interval=0.005)     # (kHz polling here slowed the end-to-end legs of eight ranks)
    leg_sampler.start()
    # ---- kernel-resident timing: inputs already in HBM
    for w in range(W):
        st = fresh(-1 - w)
        hb.playout(st, steps=steps_buf, total=total)
    total.zero_()
    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    sampler.start()
    launches0 = hb.launch_count()
    wall0 = time.perf_counter()
    timed_launches = 0
    for k in range(K):
        st = fresh(k)                       # untimed: synthetic input generation
        flush.fill_(k & 0xFF)               # untimed: evict the 126 MB L2
        l0 = hb.launch_count()
        ev[k][0].record(stream)
        hb.playout(st, steps=steps_buf, total=total)
        ev[k][1].record(stream)
        timed_launches += hb.launch_count() - l0
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    steps_done = int(total.item())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    s = torch.tensor([steps_done, timed_launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    ms_max, steps_all, launches_all = float(t.item()), int(s[0].item()), int(s[1].item())
    value = steps_all / (ms_max * 1e-3)

    # ---- the same kernel at 4x the batch (rank 0, 3 launches): how much of the issue roofline it reaches when every
    # scheduler has 14 warps instead of 3.5 (the configuration's 65,536 games are the only parallelism it offers)
    at4 = None
    if rank == 0 and n == 65536:
        n4 = 4 * n
        steps4 = torch.empty(n4, dtype=torch.int32, device=dev)
        tot4 = torch.zeros(1, dtype=torch.int64, device=dev)
        for _ in range(2):
            hb.playout(hb.init_states(n4, device=dev, seed=5, first_id=0), steps=steps4, total=tot4)
        tot4.zero_()
        ms4 = 0.0
        for k in range(3):
            st4 = hb.init_states(n4, device=dev, seed=2000 + k, first_id=0)
            flush.fill_(k)
            a4, b4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a4.record(stream)
            hb.playout(st4, steps=steps4, total=tot4)
            b4.record(stream)
            torch.cuda.synchronize()
            ms4 += a4.elapsed_time(b4)
        at4 = {"games_per_launch": n4, "us_per_launch": ms4 / 3 * 1e3, "steps_per_s": int(tot4.item()) / (ms4 * 1e-3)}
        del steps4, st4

    # ---- end to end through the public host-buffer API (HostPlayout.run): every step copies its
    # inputs from pinned host memory (H2D) and reads the final records back (D2H) inside the
    # timed region; the call pipelines 4 chunks on 4 streams so copies overlap the kernel
    host_api = hb.HostPlayout(n, device=dev, chunks=4)
    pinned_inputs = [fresh(10_000 + k).cpu().pin_memory() for k in range(min(K, 4))]
    for w in range(2):
        host_api.run(pinned_inputs[0], pinned_out)
    torch.cuda.synchronize()
    host_api.total.zero_()
    barrier()
    pinned_outs = [pinned_out] + [torch.empty_like(pinned_out).pin_memory() for _ in range(min(K, 4) - 1)]
    in_stream = [pinned_inputs[k % len(pinned_inputs)] for k in range(K)]
    out_stream = [pinned_outs[k % len(pinned_outs)] for k in range(K)]
    for _ in range(2):                            # same buffer ring as the timed call: eager pass, then graph capture
        host_api.run_many(in_stream, out_stream)
    torch.cuda.synchronize()
    host_api.total.zero_()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    host_api.run_many(in_stream, out_stream)      # returns after the last D2H completed
    e1.record(stream)
    barrier()
    total.copy_(host_api.total)
    e2e_ms = e0.elapsed_time(e1)
    e2e_steps = int(total.item())
    t2 = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    s2 = torch.tensor([e2e_steps], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        dist.all_reduce(s2, op=dist.ReduceOp.SUM)
    e2e_full_value = int(s2.item()) / (float(t2.item()) * 1e-3)

    # ---- end to end, the natural call for this workload: fresh games are defined by their 64-bit
    # keys (HarmoniesGameState() takes no input), so each step sends n keys from pinned host
    # memory and reads back (winner/meta, final scores, length) per game: HostPlayout.run_keys
    import numpy as np

    rng = np.random.default_rng(1234 + rank)
    pinned_keys = [torch.from_numpy(rng.integers(-2**63, 2**63 - 1, size=n, dtype=np.int64)).pin_memory() for _ in range(min(K, 4))]
    pinned_ress = [torch.empty((n, 3), dtype=torch.int32).pin_memory() for _ in range(min(K, 4))]
    pinned_res = pinned_ress[(K - 1) % len(pinned_ress)]
    key_stream = [pinned_keys[k % len(pinned_keys)] for k in range(K)]
    res_stream = [pinned_ress[k % len(pinned_ress)] for k in range(K)]
    for w in range(2):
        host_api.run_keys(pinned_keys[0], pinned_res)
    for _ in range(2):                            # same buffer ring as the timed call: eager pass, then graph capture
        host_api.run_keys_many(key_stream, res_stream)
    # one batch at a time (each call returns after its own D2H): the latency-bound form
    host_api.total.zero_()
    barrier()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record(stream)
    for k in range(K):
        host_api.run_keys(key_stream[k], res_stream[k])
    q1.record(stream)
    barrier()
    e2e_sync_value = int(host_api.total.item()) / (q0.elapsed_time(q1) * 1e-3)
    # the K batches as one pipelined stream (HostPlayout.run_keys_many): every batch is still copied
    # in from pinned host memory and read back, the copies of neighbouring batches overlap the kernel
    # The call lasts ~2 ms, so one descheduled host thread on one rank is visible in the max over ranks: the K-batch call is
    # timed three times (each bracketed by barriers, max over ranks) and the MEDIAN repetition is reported; all three are kept.
    e2e_reps = []
    for _rep in range(3):
        host_api.total.zero_()
        barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        host_api.run_keys_many(key_stream, res_stream)      # returns after the last D2H completed
        k1.record(stream)
        barrier()
        t3 = torch.tensor([k0.elapsed_time(k1)], dtype=torch.float64, device=dev)
        s3 = host_api.total.clone()
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            dist.all_reduce(s3, op=dist.ReduceOp.SUM)
        e2e_reps.append(int(s3.item()) / (float(t3.item()) * 1e-3))
    e2e_value = sorted(e2e_reps)[1]
    assert int((pinned_res[:, 0] >> 29).min()) >= 1, "every game must have a winner code"

    # ---- the same wave with unfused per-step kernels (K1 legal_mask + policy + K2 apply): the
    # state makes an HBM/L2 round trip every step, which is what the 258 B/step roofline models
    st = fresh(20_000)
    abuf = torch.empty(n, dtype=torch.int16, device=dev)
    sbuf = torch.empty(n, dtype=torch.uint8, device=dev)
    mbuf = torch.empty((n, 5), dtype=torch.int32, device=dev)
    for _ in range(3):
        hb.legal_mask(st, out=mbuf); hb.random_actions(st, out=abuf); hb.apply(st, abuf, status=sbuf)
    torch.cuda.synchronize()
    ugraph = torch.cuda.CUDAGraph()          # 228 launches in one graph: device time, not Python launch overhead
    with torch.cuda.graph(ugraph):
        for _ in range(76):
            hb.legal_mask(st, out=mbuf)
            hb.random_actions(st, out=abuf)
            hb.apply(st, abuf, status=sbuf)  # finished games reject the move and stay untouched
    st.copy_(fresh(20_001))
    torch.cuda.synchronize()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record(stream)
    ugraph.replay()
    u1.record(stream)
    torch.cuda.synchronize()
    unfused_steps = int(st[:, 27].sum().item())
    unfused_ms = u0.elapsed_time(u1)
    leg_clocks = leg_sampler.stop()

    pk, pk_src = peaks()
    avg_launch_s = (ms / K) * 1e-3
    algo_bytes = ALGO_BYTES_PER_STEP * (steps_done / K)
    achieved = algo_bytes / avg_launch_s / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "random playouts, 65,536 concurrent 2-player games per GPU, engine only (configs[1])",
                   "games_per_gpu": n, "engine_steps_per_wave": steps_all // K,
                   "l2": "flushed between timed iterations (256 MiB write)", "parallelism": f"games sharded x{world}, no collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 12,
                "repetitions": e2e_reps, "reported": "median of three K-batch calls",
                "api": "HostPlayout.run_keys_many: K batches of pinned host keys in, (meta, scores, length) per game out, "
                       "copies of neighbouring batches overlap the kernel (3 device buffer pairs); the call replays the "
                       "whole copy/kernel pipeline as one CUDA graph when it is given the same pinned buffer ring again",
                "one_batch_per_call": {"value": e2e_sync_value, "unit": UNIT, "scope": "rank 0",
                                       "api": "HostPlayout.run_keys, returns after its own D2H"}},
        "e2e_full_records": {"value": e2e_full_value, "unit": UNIT, "h2d_bytes_per_step": n * 128, "d2h_bytes_per_step": n * 128,
                             "api": "HostPlayout.run_many: K batches of 128-byte records in and out, pipelined over 3 device buffers (PCIe-bound)"},
        "gpu_launches": launches_all,
        "clocks": clocks, "clocks_engine_leg": leg_clocks,
        "roofline": headline_roofline(n, steps_done / K, avg_launch_s, clocks, pk, pk_src, at4),
        "wall_s": wall,
        "unfused": {"value": unfused_steps / (unfused_ms * 1e-3), "unit": UNIT, "launches": 76 * 3, "ms": unfused_ms,
                    "achieved_GBps": unfused_steps * ALGO_BYTES_PER_STEP / (unfused_ms * 1e-3) / 1e9,
                    "note": "one wave with hz_legal_mask + hz_random_actions + hz_apply per step (rank 0)"},
    }
    if not args.no_mcts:
        # second half of BASELINE.json's metric: MCTS sims/s (configs[3]), reported alongside
        out["mcts"] = mcts_measure(args, dev, world, rank, dist if world > 1 else None, play_games=args.mcts_play_games)
        if args.mcts_tower == "hand":
            # A/B on the same box in the same run: the identical search with the cuDNN tower
            ab = mcts_measure(args, dev, world, rank, dist if world > 1 else None, with_collectives=False, tower="cudnn")
            out["mcts"]["cudnn_tower_peak_step"] = {"value": ab["peak_step"]["value"], "unit": "sims/s", "ms_per_move": ab["ms_per_move"]}
        if args.mcts_leaves == 1:
            # the same leg with 4 simulations in flight per tree (virtual loss, north_star's tree mode;
            # not visit-for-visit identical to the reference's sequential search, hence reported aside)
            import copy

            vl_args = copy.copy(args)
            vl_args.mcts_leaves = 4
            vl = mcts_measure(vl_args, dev, world, rank, dist if world > 1 else None, with_collectives=False)
            out["mcts"]["virtual_loss"] = {"leaves_per_step": 4, "value": vl["peak_step"]["value"], "unit": vl["unit"],
                                           "ms_per_move": vl["ms_per_move"], "roofline_frac": vl["roofline"]["peak_step_frac"],
                                           "what": "peak step rate with 4 simulations in flight per tree"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline()
        pyref = None if args.no_python_reference else python_reference(args)
        if pyref is not None:
            out["cpu_baseline"]["python_reference"] = pyref.get("engine", pyref)
        if "mcts" in out:
            out["mcts"]["cpu_baseline"] = cpu_mcts_baseline()
            if pyref is not None:
                out["mcts"]["cpu_baseline"]["python_reference"] = {k: v for k, v in pyref.items() if k != "engine"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
