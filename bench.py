#!/usr/bin/env python
"""bench.py -- headline benchmark of the Tetris-piclim hot path on B200 (contract: see DESIGN.md "Measurement").

One "step" = one pass of the hot path over one batch of envs, exactly what a rollout does between two value-net
calls:   Tetris.move with the chosen action  ->  auto-reset of finished episodes from the prescribed-config pool  ->
40-slot afterstate enumeration + features of the new state, written to HBM.  It is ONE kernel launch
(tpl_step_observe); the three stand-alone kernels it fuses (tpl_step, tpl_reset_from_pool, tpl_afterstates) are timed
separately after the timed region and reported under "kernels".
Workload at N GPUs: 2^20 envs per GPU (BASELINE.json configs[2] at N=1, configs[4] = 8M envs at N=8), L=10, M=30,
pool = 4096 synthetic prescribed boards + 4096 carve-generated configs; weak scaling by default, envs sharded
by global env id, one NCCL all-reduce of the 64-byte episode-stats vector per rollout (timed separately: "collective_us").
`--envs-total T` fixes the total instead (strong scaling: BASELINE configs[4] as written, 8 M envs over 2 / 4 / 8 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--envs-per-gpu E | --envs-total T]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200"

L_LINES, M_MOVES = 10, 30
SEED = 0
ALG_BYTES_AFTERSTATES = 64 + 40 * 4           # read one 64 B record, write 40 x 4 B (features with the flags packed in byte 0)
ALG_BYTES_STEP = 64 + 2 + 64 + 3              # record in, action in, record out, (dlines, flags, state) out
ALG_BYTES_FUSED = 64 + 2 + 64 + 3 + 40 * 4    # the fused step: record in/out once, action, results, 40 packed feature words
# distinct-placements form: 23.14 words per env on average (uniform pieces: (17 + 3 * 34 + 17 + 17 + 9) / 7) + the 4-byte run descriptor
ALG_BYTES_FUSED_DISTINCT_FIXED = 64 + 2 + 64 + 3 + 4
GREEDY_W = [760, -360, -180, -510, 100000, -100000]

_emit = None        # set by main(): writes the JSON line to the process's real stdout


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def csrc_hash() -> str:
    """sha256 over the kernel sources: ncu-derived numbers are only quoted while they describe THIS build."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, PKG, "csrc")
    for name in ("piclim_core.cuh", "piclim_env.cuh", "piclim_kernels.cu", "piclim_value.cu"):       # device code only
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def ncu_numbers():
    """{kernel key: {traffic, alu_pipe_pct, issue_active_pct, ...}} from profiles/r02_ncu_current.json -- written by
    scripts/ncu_to_json.py from an `ncu --set full` capture and stamped with the source hash of csrc/.  If the sources have
    changed since, every ncu-derived field of the line is null (never a stale constant)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_current.json")
    try:
        with open(p) as f:
            d = json.load(f)
    except Exception:
        return {}, "no profiles/r02_ncu_current.json"
    if d.get("csrc_sha") != csrc_hash():
        return {}, f"profiles/r02_ncu_current.json describes csrc {d.get('csrc_sha')}, this build is {csrc_hash()}: ncu fields withheld"
    return d.get("kernels", {}), f"profiles/r02_ncu_current.json (csrc {d['csrc_sha']})"


def make_pool(tp, L=L_LINES, M=M_MOVES, carve=4096):
    """carve-generated prescribed configs (native generator, bit-identical to the reference's
    random.seed(k); Tetris(L, M, warm_reset=False) for k = 0..) + 4096 synthetic boards."""
    cp = tp.carve_pool(carve, L, M, seed0=0, with_solutions=False)
    return tp.concat_pools(tp.synthetic_pool(4096, seed=SEED, M=M), cp)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.err = index, [], set(), False, None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                     "hw_power_brake": 0x80, "sync_boost": 0x10, "display_clock": 0x100}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.004)
        except Exception as e:          # NVML missing: report it, do not fail the bench
            self.err = repr(e)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "error": self.err}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def pin_to_gpu_cores(local: int, nlocal: int):
    """Pin this rank to the host cores NVML reports for its GPU (its NUMA node), split between the local ranks, so that the
    pinned staging buffers and the copy-issuing thread sit next to the PCIe root of the GPU they feed."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cores = sorted(c for c in range(os.cpu_count()) if (mask[c // 64] >> (c % 64)) & 1)
        cores = [c for c in cores if c in os.sched_getaffinity(0)] or sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(nlocal, 1))
        mine = cores[(local * per) % len(cores):][:per] or cores
        os.sched_setaffinity(0, mine)
        return {"gpu_cores": f"{cores[0]}-{cores[-1]}", "pinned_to": f"{mine[0]}-{mine[-1]}", "count": len(mine)}
    except Exception as e:
        return {"error": repr(e)}


# =====================================================================================================
# reference arm / CPU baseline (the only place bench.py may execute oracle/ and baseline/_ref)
# =====================================================================================================
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "game")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "tetris.py"))


def _np_features(board):
    """(holes, bumpiness, aggregate height) of a bool[20,10] board with plain numpy (the reference has none)."""
    import numpy as np
    filled = board.any(axis=0)
    h = 20 - np.where(filled, board.argmax(axis=0), 20)
    return int(h.sum() - board.sum()), int(np.abs(np.diff(h)).sum()), int(h.sum())


def _reference_worker(args):
    """The UNMODIFIED reference (game/tetris.py staged under baseline/_ref/ by __graft_entry__.build()): per env-step, 40 x
    (clone -> Tetris.move -> features) + one Tetris.move with a random action; prescribed configs injected the way
    load_warm_reset does (game/tetris.py:447)."""
    wid, budget_s, pool_rows, pool_pieces, pool_np = args
    import copy
    import random
    import numpy as np
    sys.dont_write_bytecode = True
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import tetris as ref                                   # the reference module, unmodified
    rnd = random.Random(1000 + wid)
    K = len(pool_rows)
    t0 = time.perf_counter()
    slots = moves = episodes = 0
    while time.perf_counter() - t0 < budget_s:
        k = rnd.randrange(K)
        g = ref.Tetris.__new__(ref.Tetris)
        g.L, g.M, g.warm_reset, g.render, g.debug = L_LINES, M_MOVES, False, False, False
        g.lines_cleared, g.moves_used, g.state = 0, 0, None
        g.board = ((pool_rows[k][:, None] >> np.arange(10)) & 1).astype(bool)
        g.pieces = [int(p) for p in pool_pieces[k][:pool_np[k]]]
        episodes += 1
        while g.state is None and g.pieces and time.perf_counter() - t0 < budget_s:
            for r in range(4):
                for c in range(10):
                    h = copy.copy(g); h.board = g.board.copy(); h.pieces = list(g.pieces)
                    h.move(r, c)
                    _np_features(h.board)
                    slots += 1
            g.move(rnd.randint(0, 3), rnd.randint(0, 9))
            moves += 1
    return slots, moves, time.perf_counter() - t0, episodes


def _py_port_worker(args):
    seed, budget_s, pool_rows, pool_pieces, pool_np = args
    from oracle import piclim_oracle as po
    K = len(pool_rows)
    t0 = time.perf_counter()
    slots = moves = 0
    ep = 0
    while time.perf_counter() - t0 < budget_s:
        k = po.config_index(SEED, seed, ep, K)
        env = po.OracleEnv(L_LINES, M_MOVES).load([int(x) for x in pool_rows[k]], [int(x) for x in pool_pieces[k][:pool_np[k]]])
        t = 0
        while env.state == po.RUNNING and env.pieces and time.perf_counter() - t0 < budget_s:
            po.afterstates(env)                      # 40 x (clone + move + features), the composed reference path
            slots += 40
            rot, loc = po.random_action(SEED, seed, ep, t)
            env.move(rot, loc)
            moves += 1
            t += 1
        ep += 1
    return slots, moves, time.perf_counter() - t0, ep


def cpu_python(pool, budget_s: float, procs: int, use_reference: bool):
    """The path as the reference implements it -- Python objects, one env at a time -- in `procs` processes
    (multiprocessing, like the reference's own generators): the reference itself when it is staged, else its port."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    worker = _reference_worker if use_reference else _py_port_worker
    with ctx.Pool(procs) as p:
        res = p.map(worker, [(i, budget_s, pool.rows, pool.pieces, pool.npieces) for i in range(procs)])
    wall = max(r[2] for r in res)
    return sum(r[0] for r in res) / wall, sum(r[1] for r in res) / wall, sum(r[3] for r in res)


def cpu_c_port(pool, budget_s: float, threads: int):
    """The same path in the plain-C oracle (oracle/piclim_oracle.c), all host threads: per env-step 40 afterstate
    evaluations + features, one move, auto-reset (its greedy rollout does exactly that work)."""
    from oracle import c_oracle
    n = 4096 * max(1, threads)
    st = c_oracle.BatchState(n)
    ep, ts, _ = c_oracle.rollout(st, 0, SEED, L_LINES, M_MOVES, pool.rows, pool.pieces, pool.npieces, 0, True)
    steps_done, t0 = 0, time.perf_counter()
    chunk = 4
    while True:
        c_oracle.rollout(st, 0, SEED, L_LINES, M_MOVES, pool.rows, pool.pieces, pool.npieces, chunk, False, ep, ts,
                         nthreads=threads, weights=GREEDY_W)
        steps_done += chunk
        el = time.perf_counter() - t0
        if el > budget_s:
            break
    return n * steps_done * 40 / el, n * steps_done / el, n, steps_done


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores -- game/tetris.py itself when
    build() staged it under baseline/_ref/ (kind "reference"), else its Python restatement (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tetris_piclim as tp
    pool = make_pool(tp)
    cores = os.cpu_count() or 1
    use_ref = reference_available()
    per_step_budget = 2.0
    vals = []
    for i in range(args.warmup + args.steps):
        a, m, eps = cpu_python(pool, per_step_budget if i >= args.warmup else 0.5, cores, use_ref)
        if i >= args.warmup:
            vals.append((a, m, eps))
    a = sum(v[0] for v in vals) / len(vals)
    m = sum(v[1] for v in vals) / len(vals)
    episodes = sum(v[2] for v in vals)
    ca, cm, cn, cs = cpu_c_port(pool, 5.0, cores)
    cfg = workload_config(args, args.envs_per_gpu * max(args.gpus, 1))
    cfg["reference_sample"] = (f"a RATE, not the 2^20-env batch: {cores} processes x {per_step_budget:.0f} s per step of fresh episodes drawn from "
                               f"the same pool (L, M, actions as the B200 arm); {episodes} episodes, {int(m * per_step_budget * len(vals))} env-steps sampled in the timed steps")
    what = ("game/tetris.py of the reference, unmodified (baseline/_ref/, staged by build()): Tetris objects with prescribed configs "
            "injected as load_warm_reset does (:447); per env-step 40 x (copy -> Tetris.move -> numpy features) + 1 Tetris.move"
            if use_ref else
            "oracle/piclim_oracle.py (Python restatement of game/tetris.py; baseline/_ref/ not staged), one env object per episode, clone+move+features per slot")
    line = {
        "impl": "reference", "metric": "afterstates/sec", "value": a, "unit": "afterstates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_budget * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "python-int", "data": "synthetic",
        "config": cfg,
        "env_steps_per_sec": m,
        "cpu_baseline": {"value": a, "unit": "afterstates/s", "cores": cores, "kind": "reference" if use_ref else "port",
                         "sample": f"{what}; {cores} processes x {per_step_budget:.0f} s per step, same pool/L/M"},
        "c_port": {"value": ca, "unit": "afterstates/s", "env_steps_per_sec": cm, "cores": cores,
                   "sample": f"oracle/piclim_oracle.c greedy rollout, {cn} envs x {cs} steps"},
        "e2e": {"value": a, "unit": "afterstates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(json.dumps(line))


def workload_config(args, n_total):
    return {"workload": "2^20 envs/GPU x [40-slot afterstate enumeration + features -> move -> auto-reset], "
                        "prescribed-config pool (4096 synthetic + 4096 carve-generated), L=10 M=30 (BASELINE configs[2]; configs[4] at 8 GPUs)",
            "envs_per_gpu": args.envs_per_gpu, "envs_total": n_total, "L": L_LINES, "M": M_MOVES, "pool": 8192,
            "l2": "working set per step (64 MiB state r+w, 160 MiB afterstate outputs) exceeds the 126 MB L2",
            "parallelism": f"envs sharded by global env id over {args.gpus} GPU(s); one 64-byte NCCL all-reduce per rollout"}


# =====================================================================================================
# B200 arm
# =====================================================================================================
def run_b200(args):
    import numpy as np
    import torch
    import tetris_piclim as tp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nlocal = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    affinity = pin_to_gpu_cores(local, nlocal)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = args.envs_per_gpu
    K, W = args.steps, args.warmup
    pool = make_pool(tp)
    env = tp.BatchedTetris(n, L_LINES, M_MOVES, device=dev, seed=SEED, config_pool=pool, env_base=rank * n)
    env.reset()
    env.rollout_random(8)                      # decorrelate episode phases so the mix of boards is stationary
    env.reset(done_only=True)
    env.stats.zero_()
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    total = W + K
    rot = torch.randint(0, 4, (total, n), device=dev, dtype=torch.uint8, generator=g)
    loc = torch.randint(0, 10, (total, n), device=dev, dtype=torch.uint8, generator=g)

    def one_step(i):
        env.step_observe(rot[i], loc[i], packed=True)          # ONE launch: move -> auto-reset -> afterstates

    sampler = ClockSampler(local); sampler.start()     # samples through warm-up, the timed region and the e2e leg
    for i in range(W):
        one_step(i)
    warm_stats = env.stats.clone()
    if dist: dist.all_reduce(warm_stats)               # the collective is warmed up too (NCCL connects lazily)
    torch.cuda.synchronize()
    if dist: dist.barrier()
    launches0 = tp.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    # The K launches are queued while the GPU spins in a few-millisecond delay kernel that sits BEFORE the start event: the
    # timed region then runs back to back from the device's queue, and a host thread that loses its core for a millisecond
    # (8 ranks + helper threads on one VM) no longer shows up as GPU idle time inside a 2 ms measurement.
    if hasattr(torch.cuda, "_sleep"):
        torch.cuda._sleep(int(2.0e6 * max(1.0, K / 20.0)))
    # ... and the W warm-up steps run once more right in front of the start event, queued behind the same delay kernel: the first
    # steps after the GPU has sat idle through the host-side synchronisation above (and the near-idle delay kernel) run 3-5 % slow
    # for about five launches (scripts/bench_curve.py: 0.105, 0.102, 0.104, 0.103, 0.102, then 0.100 with per-step events), so
    # warm-up that is separated from the timed region by an idle gap does not warm anything.
    for i in range(W):
        one_step(i)
    launches0 = tp.launch_count()                      # (host-side counter: the launches counted below are the K timed ones)
    t_start.record()
    for i in range(K):
        one_step(W + i)
    t_end.record()
    # the one collective of the path (64 bytes of episode statistics per rollout): after the last step, timed on its own
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = env.stats.clone()
    c0.record()
    if dist: dist.all_reduce(stats)
    c1.record()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    launches = tp.launch_count() - launches0
    ms = torch.tensor([t_start.elapsed_time(t_end), c0.elapsed_time(c1)], device=dev, dtype=torch.float64)
    if dist: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total, collective_ms = float(ms[0].item()), float(ms[1].item())

    # the three stand-alone kernels the fused step replaces (explanatory numbers): each is launched `reps` times back to
    # back between two CUDA events, so the ~5 us an event pair adds around a single 25 us launch does not count
    reps = 10
    def timed(fn, reps=reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for j in range(reps):
            fn(j)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    env.count_stats = False
    k_ms = [0.0, 0.0, 0.0]
    env.reset(); env.rollout_random(8); env.reset(done_only=True)
    k_ms[0] = timed(lambda j: env.afterstates(packed=True))
    k_ms[1] = timed(lambda j: env.move(rot[j % total], loc[j % total]))          # 10 moves: most episodes still running
    k_ms[2] = timed(lambda j: env.reset(done_only=True))                            # first call resets, the rest only scan
    # the fused step in the distinct-placements form (same state distribution: continue the random-action rollout)
    env.reset(); env.rollout_random(8); env.reset(done_only=True)
    for j in range(3):
        env.step_observe_distinct(rot[j % total], loc[j % total])
    dsteps = min(K, 20)
    d_ms = timed(lambda j: env.step_observe_distinct(rot[(3 + j) % total], loc[(3 + j) % total]), dsteps)
    d_used = int(env.step_observe_distinct(rot[0], loc[0])[5].item())

    # ---- end-to-end through the host-buffer C ABI (pinned host buffers, H2D + D2H inside the timed region) ----
    henv = tp.HostBatchedTetris(n, L_LINES, M_MOVES, device=local, seed=SEED, env_base=rank * n, config_pool=pool)
    henv.reset()
    cap = henv.distinct_capacity()
    e2e_steps = max(3, min(K, 10))
    # every step's actions wait in pinned host memory (one row per step, as a host-side policy would leave them); the results
    # land in pinned host buffers
    pin = {k: tp.PinnedArray(s, d) for k, (s, d) in dict(rot=((e2e_steps + 2, n), np.uint8), loc=((e2e_steps + 2, n), np.uint8),
           dl=((n,), np.int8), fl=((n,), np.uint8), st=((n,), np.int8), feats=((40, n, 4), np.uint8), rows=((cap,), np.uint32),
           runs=((n,), np.uint32)).items()}
    idx = [i % total for i in range(e2e_steps + 2)]
    pin["rot"].array[:] = rot[idx].cpu().numpy(); pin["loc"].array[:] = loc[idx].cpu().numpy()
    out3 = [pin[k].array for k in ("dl", "fl", "st")]

    def e2e_leg(call):
        for i in range(2):
            call(pin["rot"].array[i], pin["loc"].array[i])
        if dist: dist.barrier()
        t0 = time.perf_counter()
        extra = 0
        for i in range(e2e_steps):
            extra += call(pin["rot"].array[2 + i], pin["loc"].array[2 + i]) or 0
        s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if dist: dist.all_reduce(s, op=dist.ReduceOp.MAX)
        return float(s.item()), extra

    # 40-slot compact form: every slot of the grid crosses PCIe (163 B per env-step)
    e2e_s, _ = e2e_leg(lambda r, c: henv.step_observe(r, c, *out3, pin["feats"].array, None))
    # distinct-placements form: only the placements that differ (+ a 4-byte run descriptor per env)
    e2e_d_s, d_words = e2e_leg(lambda r, c: henv.step_observe_distinct(r, c, *out3, pin["rows"].array, pin["runs"].array))
    # the same call with the features left in HBM (a policy on the GPU reads them there, as train.py does): H2D actions,
    # kernel, D2H of (rows cleared, flags, state) only
    e2e_dev_s, _ = e2e_leg(lambda r, c: henv.step_observe(r, c, *out3, None, None))
    chunks = henv.chunks()
    henv.close()
    # PCIe reference for the e2e figures: D2H of a pinned buffer the size of one step's 40-slot output, all ranks at once
    pcie = {}
    try:
        nb = 163 * n
        hb, db = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8, device=dev)
        for _ in range(2):
            hb.copy_(db, non_blocking=True)
        torch.cuda.synchronize()
        if dist: dist.barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            hb.copy_(db, non_blocking=True)
        torch.cuda.synchronize()
        bw = torch.tensor([nb * 5 / (time.perf_counter() - t0) / 1e9], device=dev, dtype=torch.float64)
        lo, sm = bw.clone(), bw.clone()
        if dist:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        pcie = {"d2h_GBps_per_gpu_min": float(lo.item()), "d2h_GBps_all_gpus": float(sm.item()),
                "how": f"D2H of a pinned {nb >> 20} MiB buffer x 5, all {world} rank(s) concurrently"}
        del hb, db
    except Exception as e:          # never fail the bench on the microbenchmark
        pcie = {"error": repr(e)}
    sampler.stop_flag = True; sampler.join(timeout=2)

    if rank != 0:
        if dist: dist.destroy_process_group()
        return

    n_total = n * world
    hbm_peak, peak_src = load_peaks()
    ncu, ncu_src = ncu_numbers()
    nf = ncu.get("step_observe_kernel<0, 1>", {}) if n == (1 << 20) else {}
    nd = ncu.get("step_observe_kernel<4, 0>", {}) if n == (1 << 20) else {}
    na = ncu.get("afterstates_kernel<0, 1>", {}) if n == (1 << 20) else {}
    as_gbs = ALG_BYTES_AFTERSTATES * n / (k_ms[0] * 1e-3) / 1e9
    fused_ms = ms_total / K
    fused_gbs = ALG_BYTES_FUSED * n / (fused_ms * 1e-3) / 1e9
    d_bytes = ALG_BYTES_FUSED_DISTINCT_FIXED * n + 4 * d_used
    d_gbs = d_bytes / (d_ms * 1e-3) / 1e9
    d2h_distinct = (4 * d_words / e2e_steps) + 4 * n + 3 * n
    pcie_bw = pcie.get("d2h_GBps_all_gpus")

    def e2e_block(seconds, d2h_bytes, api, **kw):
        b = {"value": n_total * 40 * e2e_steps / seconds, "unit": "afterstates/s", "env_steps_per_sec": n_total * e2e_steps / seconds,
             "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps, "api": api,
             "d2h_GBps_all_gpus": d2h_bytes * world * e2e_steps / seconds / 1e9}
        if pcie_bw:
            b["frac_of_measured_pcie_d2h"] = b["d2h_GBps_all_gpus"] / pcie_bw
        b.update(kw)
        return b

    line = {
        "metric": "afterstates/sec", "value": n_total * 40 * K / (ms_total * 1e-3), "unit": "afterstates/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
        "scaling": "strong" if args.envs_total else "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, n_total),
        "env_steps_per_sec": n_total * K / (ms_total * 1e-3),
        "collective_us": collective_ms * 1e3,
        "roofline": {"bound": "hbm", "kernel": "step_observe_kernel<0, 1> (fused move + auto-reset + afterstates)",
                     "achieved": fused_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fused_gbs / hbm_peak,
                     "traffic": nf.get("traffic"), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALG_BYTES_FUSED * n, "avg_launch_ms": fused_ms,
                     "ncu_source": ncu_src,
                     "integer_pipe": {"alu_pipe_pct_of_peak_ncu": nf.get("alu_pipe_pct"), "issue_active_pct_ncu": nf.get("issue_active_pct"),
                                      "warp_instructions_per_launch_ncu": nf.get("warp_instructions"),
                                      "afterstates_kernel_alu_pipe_pct_of_peak_ncu": na.get("alu_pipe_pct"),
                                      "note": "the enumeration is bound by integer instruction issue (ALU + FMA pipes), not by HBM (DESIGN.md section 3); "
                                              "percentages are ncu sm__inst_executed_pipe_alu / smsp__issue_active of peak, captures under profiles/"}},
        "distinct_form": {"kernel": "step_observe_kernel<4, 0> (the same fused step, distinct placements only: tpl_step_observe_distinct)",
                          "ms_per_step": d_ms, "env_steps_per_sec": n / (d_ms * 1e-3), "grid_equivalent_afterstates_per_sec": n * 40 / (d_ms * 1e-3),
                          "distinct_afterstates_per_sec": d_used / (d_ms * 1e-3), "words_per_env": d_used / n,
                          "roofline": {"bound": "hbm", "achieved": d_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": d_gbs / hbm_peak,
                                       "traffic": nd.get("traffic"), "algorithmic_bytes_per_launch": d_bytes,
                                       "alu_pipe_pct_of_peak_ncu": nd.get("alu_pipe_pct"), "issue_active_pct_ncu": nd.get("issue_active_pct"),
                                       "warp_instructions_per_launch_ncu": nd.get("warp_instructions")},
                          "note": "per GPU (rank 0); not part of `value`, which stays the 40-slot compact form"},
        "kernels": {
            "note": "stand-alone kernels (3 launches per step); their sum is what the fused step replaces",
            "afterstates": {"ms": k_ms[0], "afterstates_per_s": n * 40 / (k_ms[0] * 1e-3), "GBps": as_gbs,
                            "frac_of_hbm_peak": as_gbs / hbm_peak},
            "step": {"ms": k_ms[1], "env_steps_per_s": n / (k_ms[1] * 1e-3), "GBps": ALG_BYTES_STEP * n / (k_ms[1] * 1e-3) / 1e9},
            "reset_done": {"ms": k_ms[2]},
        },
        "e2e": e2e_block(e2e_d_s, d2h_distinct, "tpl_env_step_observe_distinct (host-buffer C ABI, pinned buffers, pipelined over env chunks)",
                         form="distinct placements: per env the 9 / 17 / 34 placements that differ + a run descriptor; the 40-slot grid is "
                              "their expansion (rot % n_rot, min(loc, 10 - w) -- tests: expand(distinct) == oracle grid), so `value` counts 40 "
                              "grid slots per env-step like the reference arm, which evaluates all 40 clone+move per step",
                         distinct_afterstates_per_sec=(d_words / e2e_steps) * world / (e2e_d_s / e2e_steps), chunks=chunks),
        "e2e_40slot": e2e_block(e2e_s, 163 * n, "tpl_env_step_observe (host-buffer C ABI, compact 40-slot form, pinned buffers, pipelined over env chunks)",
                                note="round 1's headline form: every one of the 40 slots crosses PCIe, aliases included"),
        "e2e_features_on_device": e2e_block(e2e_dev_s, 3 * n, "tpl_env_step_observe with feats=NULL",
                                            note="the features stay in HBM for a policy on the GPU; not the headline e2e (that one ships the observation to the host)"),
        "pcie": pcie,
        "host_affinity": affinity,
        "gpu_launches": int(launches),
        "clocks": sampler.result(),
        "episode_stats": dict(zip(("episodes", "wins", "topouts", "movelimit_losses", "lines", "moves", "steps", "resets"),
                                  (int(v) for v in stats.tolist()))),
    }

    if world == 1:
        line.update(extra_single_gpu(tp, torch, dev, pool, args))
    _emit(json.dumps(line))
    if dist: dist.destroy_process_group()


def extra_single_gpu(tp, torch, dev, pool, args):
    """N=1 extras: BASELINE configs[1] (4096 envs, latency-bound: looped in one CUDA graph), the fused rollouts, the fused
    step on policy-driven states and at (L=15, M=40), and the CPU baseline (C port on all host cores, bounded sample)."""
    out = {}
    n = args.envs_per_gpu

    def timed_steps(env, steps=20, warm=3):
        g = torch.Generator(device=dev); g.manual_seed(99)
        r = torch.randint(0, 4, (steps + warm, n), device=dev, dtype=torch.uint8, generator=g)
        c = torch.randint(0, 10, (steps + warm, n), device=dev, dtype=torch.uint8, generator=g)
        for i in range(warm):
            env.step_observe(r[i], c[i], packed=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()                                    # (no idle gap between the warm-up steps and the timed ones)
        for i in range(steps):
            env.step_observe(r[warm + i], c[warm + i], packed=True)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    # configs[1]: 4096 envs x 40 slots, 1000 iterations inside one CUDA graph (inputs fit L2: says so)
    env = tp.BatchedTetris(4096, L_LINES, M_MOVES, device=dev, seed=SEED, config_pool=pool)
    env.reset(); env.rollout_random(6); env.reset(done_only=True)
    s = torch.cuda.Stream(device=dev)
    iters = 1000
    with torch.cuda.stream(s):
        env.afterstates(packed=True)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                env.afterstates(packed=True)
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); e1.record(s); s.synchronize()
    ms = e0.elapsed_time(e1)
    out["config_4096_envs"] = {"afterstates_per_s": 4096 * 40 * iters / (ms * 1e-3), "us_per_call": ms * 1e3 / iters,
                               "note": "BASELINE configs[1]; sub-wave problem (latency-bound): one thread per (env, rotation), 128 CTAs; 1000 calls in one CUDA graph; inputs L2-resident"}
    # fused rollouts (state in registers across steps)
    env = tp.BatchedTetris(n, L_LINES, M_MOVES, device=dev, seed=SEED, config_pool=pool)
    env.reset()
    for name, fn, steps in (("rollout_random", lambda k: env.rollout_random(k), 64),
                            ("rollout_greedy", lambda k: env.rollout_greedy(k, GREEDY_W), 16)):
        fn(2); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(steps); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"env_steps_per_s": n * steps / (ms * 1e-3), "ms": ms, "steps": steps}
        if name == "rollout_greedy":
            out[name]["afterstates_per_s"] = n * steps * 40 / (ms * 1e-3)
    # the fused step on policy-driven states: the greedy rollout above left every env mid-episode under a competent policy
    # (dense low boards: row-completing slots -- the deferred path -- are far more frequent than under random actions)
    env.count_stats = False
    ms = timed_steps(env)
    out["policy_driven_states"] = {"ms_per_step": ms, "afterstates_per_s": n * 40 / (ms * 1e-3), "env_steps_per_s": n / (ms * 1e-3),
                                   "frac_of_hbm_peak": ALG_BYTES_FUSED * n / (ms * 1e-3) / 1e9 / load_peaks()[0],
                                   "note": "tpl_step_observe (compact form) timed right after 18 greedy-policy moves per env; the timed "
                                           "moves themselves are uniform random (they degrade the boards only gradually over the 20 steps)"}
    # (L=15, M=40): the reference's other (L, M) pair (game/main.py:33,50)
    pool2 = make_pool(tp, 15, 40, carve=256)
    env2 = tp.BatchedTetris(n, 15, 40, device=dev, seed=SEED, config_pool=pool2)
    env2.reset(); env2.rollout_random(8); env2.reset(done_only=True)
    env2.count_stats = False
    ms = timed_steps(env2)
    out["L15_M40"] = {"ms_per_step": ms, "afterstates_per_s": n * 40 / (ms * 1e-3), "env_steps_per_s": n / (ms * 1e-3),
                      "frac_of_hbm_peak": ALG_BYTES_FUSED * n / (ms * 1e-3) / 1e9 / load_peaks()[0],
                      "note": "same fused step at L=15, M=40; pool = 4096 synthetic + 256 carve-generated configs for (15, 40)"}
    del env, env2
    if not args.no_dqn:
        try:
            out["dqn_loop_65536"] = dqn_leg(tp, torch, dev)
        except Exception as e:      # the extras never fail the headline
            out["dqn_loop_65536"] = {"error": repr(e)}
    cores = os.cpu_count() or 1
    ca, cm, cn, cs = cpu_c_port(pool, 10.0, cores)
    out["cpu_baseline"] = {"value": ca, "unit": "afterstates/s", "cores": cores, "kind": "port", "env_steps_per_sec": cm,
                           "sample": f"oracle/piclim_oracle.c (plain-C restatement) greedy rollout = 40 afterstate evals + 1 move "
                                     f"+ auto-reset per env-step, {cn} envs x {cs} steps, same pool/L/M, {cores} threads; the "
                                     f"reference's own Python speed is what --impl reference reports"}
    return out


def dqn_leg(tp, torch, dev):
    """BASELINE configs[3]: the DQN afterstate-value loop (model/train.py's constants) driving 65 536 GPU envs -- measured by
    scripts/dqn_bench.py in a FRESH process on the same GPU: capturing the optimiser block into a CUDA graph empties the caching
    allocator, which inside this process (gigabytes of cached blocks from the legs above) costs seconds of the timed run."""
    import subprocess
    env = dict(os.environ)
    env.setdefault("CUDA_VISIBLE_DEVICES", str(dev.index or 0))           # the same GPU as this process
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dqn_bench.py")], env=env, capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        return {"error": r.stderr[-400:]}
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--envs-total", type=int, default=0, help="strong scaling: this many envs in total, split over the GPUs")
    ap.add_argument("--no-dqn", action="store_true", help="skip the DQN-loop extra at N=1")
    args = ap.parse_args()
    if args.envs_total:
        world = max(int(os.environ.get("WORLD_SIZE", "1")), 1)
        args.envs_per_gpu = args.envs_total // world
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # Rank 0 must print ONE JSON line on stdout.  Libraries write there too (NCCL's version banner comes out of C code at
    # the first collective), so file descriptor 1 points at stderr for the whole run and the line goes to the saved one.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit
    def _emit(text):
        os.write(real_stdout, (text + "\n").encode())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
