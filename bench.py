#!/usr/bin/env python
"""bench.py -- headline benchmark of the Tetris-piclim hot path on B200 (contract: see DESIGN.md "Measurement").

One "step" = one pass of the hot path over one batch of envs, exactly what a rollout does between two value-net
calls:   Tetris.move with the chosen action  ->  auto-reset of finished episodes from the prescribed-config pool  ->
40-slot afterstate enumeration + features of the new state, written to HBM.  It is ONE kernel launch
(tpl_step_observe); the three stand-alone kernels it fuses (tpl_step, tpl_reset_from_pool, tpl_afterstates) are timed
separately after the timed region and reported under "kernels".
Workload at N GPUs: 2^20 envs per GPU (BASELINE.json configs[2] at N=1, configs[4] = 8M envs at N=8), L=10, M=30,
pool = 4096 synthetic prescribed boards + 4096 carve-generated configs; weak scaling, envs sharded
by global env id, one NCCL all-reduce of the 64-byte episode-stats vector per rollout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--envs-per-gpu E]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L_LINES, M_MOVES = 10, 30
SEED = 0
ALG_BYTES_AFTERSTATES = 64 + 40 * 4           # read one 64 B record, write 40 x 4 B (features with the flags packed in byte 0)
ALG_BYTES_STEP = 64 + 2 + 64 + 3              # record in, action in, record out, (dlines, flags, state) out
ALG_BYTES_FUSED = 64 + 2 + 64 + 3 + 40 * 4    # the fused step: record in/out once, action, results, 40 packed feature words
# dram__bytes_read.sum + dram__bytes_write.sum of one afterstates_kernel<0> launch at 2^20 envs, from the ncu --set full
# capture summarised in profiles/r01_ncu_full_v8_afterstates_step.txt (69.9 MB + 115.0 MB; the rest of the 160 MiB of
# output is still in L2 when the kernel ends)
NCU_TRAFFIC_AFTERSTATES_2P20 = 184.9e6
# the fused step_observe_kernel<0, 1>, same kind of capture (profiles/r01_ncu_full_v8_fused_step_observe.txt):
# 78.4 MB read + 187.0 MB write per launch against 307 MB algorithmic (the tail of the writes is still in L2)
NCU_TRAFFIC_FUSED_2P20 = 265.3e6
NCU_ALU_PIPE_PCT = 62.9                       # sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active, fused kernel (v8 capture)
NCU_ISSUE_ACTIVE_PCT = 67.3                   # smsp__issue_active.avg.pct_of_peak_sustained_active, same capture
NCU_ALU_PIPE_PCT_AFTERSTATES = 67.2           # same metric, stand-alone afterstates_kernel<0, 1> (profiles/r01_ncu_full_v8_afterstates_step.txt; issue-active 71 %)


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


def make_pool(tp):
    """4096 carve-generated prescribed configs (native generator, bit-identical to the reference's
    random.seed(k); Tetris(10, 30, warm_reset=False) for k = 0..4095) + 4096 synthetic boards."""
    carve = tp.carve_pool(4096, L_LINES, M_MOVES, seed0=0, with_solutions=False)
    return tp.concat_pools(tp.synthetic_pool(4096, seed=SEED, M=M_MOVES), carve)


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


# =====================================================================================================
# reference arm / CPU baseline (the only place bench.py may execute oracle/)
# =====================================================================================================
def _py_port_worker(args):
    seed, budget_s, pool_rows, pool_pieces, pool_np = args
    from oracle import piclim_oracle as po
    import numpy as np
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
    return slots, moves, time.perf_counter() - t0


def cpu_python_port(pool, budget_s: float, procs: int):
    """The path as the reference implements it -- Python objects, one env at a time -- restated in oracle/piclim_oracle.py,
    run in `procs` processes (multiprocessing, like the reference's own generators)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as p:
        res = p.map(_py_port_worker, [(i, budget_s, pool.rows, pool.pieces, pool.npieces) for i in range(procs)])
    wall = max(r[2] for r in res)
    return sum(r[0] for r in res) / wall, sum(r[1] for r in res) / wall


def cpu_c_port(pool, budget_s: float, threads: int):
    """The same path in the plain-C oracle (oracle/piclim_oracle.c), all host threads: per env-step 40 afterstate
    evaluations + features, one move, auto-reset (its greedy rollout does exactly that work)."""
    import numpy as np
    from oracle import c_oracle
    n = 4096 * max(1, threads)
    st = c_oracle.BatchState(n)
    ep, ts, _ = c_oracle.rollout(st, 0, SEED, L_LINES, M_MOVES, pool.rows, pool.pieces, pool.npieces, 0, True)
    w = [760, -360, -180, -510, 100000, -100000]
    steps_done, t0 = 0, time.perf_counter()
    chunk = 4
    while True:
        c_oracle.rollout(st, 0, SEED, L_LINES, M_MOVES, pool.rows, pool.pieces, pool.npieces, chunk, False, ep, ts,
                         nthreads=threads, weights=w)
        steps_done += chunk
        el = time.perf_counter() - t0
        if el > budget_s:
            break
    return n * steps_done * 40 / el, n * steps_done / el, n, steps_done


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (its Python restatement, all host cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tetris_piclim as tp
    pool = make_pool(tp)
    cores = os.cpu_count() or 1
    per_step_budget = 2.0
    vals = []
    for i in range(args.warmup + args.steps):
        a, m = cpu_python_port(pool, per_step_budget if i >= args.warmup else 0.5, cores)
        if i >= args.warmup:
            vals.append((a, m))
    a = sum(v[0] for v in vals) / len(vals)
    m = sum(v[1] for v in vals) / len(vals)
    ca, cm, cn, cs = cpu_c_port(pool, 5.0, cores)
    line = {
        "impl": "reference", "metric": "afterstates/sec", "value": a, "unit": "afterstates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_budget * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "python-int", "data": "synthetic",
        "config": workload_config(args, args.envs_per_gpu * max(args.gpus, 1)),
        "env_steps_per_sec": m,
        "cpu_baseline": {"value": a, "unit": "afterstates/s", "cores": cores, "kind": "port",
                         "sample": f"oracle/piclim_oracle.py (Python restatement of game/tetris.py, one env object per episode, "
                                   f"clone+move+features per slot), {cores} processes x {per_step_budget:.0f} s per step, same pool/L/M"},
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
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
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
    torch.cuda.synchronize()
    if dist: dist.barrier()
    launches0 = tp.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_start.record()
    for i in range(K):
        one_step(W + i)
    stats = env.stats.clone()
    if dist: dist.all_reduce(stats)            # the one collective of the path: 64 bytes per rollout
    t_end.record()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    launches = tp.launch_count() - launches0
    ms = torch.tensor([t_start.elapsed_time(t_end)], device=dev, dtype=torch.float64)
    if dist: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # the three stand-alone kernels the fused step replaces (explanatory numbers): each is launched `reps` times back to
    # back between two CUDA events, so the ~5 us an event pair adds around a single 25 us launch does not count
    reps = 10
    def timed(fn):
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

    # ---- end-to-end through the host-buffer C ABI (pinned host buffers, H2D + D2H inside the timed region) ----
    henv = tp.HostBatchedTetris(n, L_LINES, M_MOVES, device=local, seed=SEED, env_base=rank * n, config_pool=pool)
    henv.reset()
    pin = {k: tp.PinnedArray(s, d) for k, (s, d) in dict(rot=((n,), np.uint8), loc=((n,), np.uint8), dl=((n,), np.int8),
           fl=((n,), np.uint8), st=((n,), np.int8), feats=((40, n, 4), np.uint8)).items()}
    hrot, hloc = rot.cpu().numpy(), loc.cpu().numpy()
    e2e_steps = max(3, min(K, 10))
    bufs = [pin[k].array for k in ("rot", "loc", "dl", "fl", "st", "feats")] + [None]      # compact afterstate form
    for i in range(2):
        pin["rot"].array[:] = hrot[i]; pin["loc"].array[:] = hloc[i]
        henv.step_observe(*bufs)
    if dist: dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        pin["rot"].array[:] = hrot[(2 + i) % total]; pin["loc"].array[:] = hloc[(2 + i) % total]
        henv.step_observe(*bufs)
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if dist: dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    # the same call with the features left in HBM (a policy on the GPU reads them there, as train.py does): H2D actions,
    # kernel, D2H of (rows cleared, flags, state) only
    bufs_dev = bufs[:5] + [None, None]
    henv.step_observe(*bufs_dev)
    if dist: dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        pin["rot"].array[:] = hrot[(2 + i) % total]; pin["loc"].array[:] = hloc[(2 + i) % total]
        henv.step_observe(*bufs_dev)
    e2e_dev_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if dist: dist.all_reduce(e2e_dev_s, op=dist.ReduceOp.MAX)
    e2e_dev_s = float(e2e_dev_s.item())
    sampler.stop_flag = True; sampler.join(timeout=2)
    henv.close()

    if rank != 0:
        if dist: dist.destroy_process_group()
        return

    n_total = n * world
    hbm_peak, peak_src = load_peaks()
    as_gbs = ALG_BYTES_AFTERSTATES * n / (k_ms[0] * 1e-3) / 1e9
    fused_ms = ms_total / K
    fused_gbs = ALG_BYTES_FUSED * n / (fused_ms * 1e-3) / 1e9
    line = {
        "metric": "afterstates/sec", "value": n_total * 40 * K / (ms_total * 1e-3), "unit": "afterstates/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, n_total),
        "env_steps_per_sec": n_total * K / (ms_total * 1e-3),
        "roofline": {"bound": "hbm", "kernel": "step_observe_kernel<0, 1> (fused move + auto-reset + afterstates)",
                     "achieved": fused_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fused_gbs / hbm_peak,
                     "traffic": NCU_TRAFFIC_FUSED_2P20 if n == (1 << 20) else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALG_BYTES_FUSED * n, "avg_launch_ms": fused_ms,
                     "integer_pipe": {"alu_pipe_pct_of_peak_ncu": NCU_ALU_PIPE_PCT, "issue_active_pct_ncu": NCU_ISSUE_ACTIVE_PCT,
                                      "afterstates_kernel_alu_pipe_pct_of_peak_ncu": NCU_ALU_PIPE_PCT_AFTERSTATES,
                                      "note": "the enumeration is bound by integer instruction issue (ALU + FMA pipes), not by HBM (DESIGN.md section 3); "
                                              "percentages are ncu sm__inst_executed_pipe_alu of peak, captures under profiles/"}},
        "kernels": {
            "note": "stand-alone kernels (3 launches per step); their sum is what the fused step replaces",
            "afterstates": {"ms": k_ms[0], "afterstates_per_s": n * 40 / (k_ms[0] * 1e-3), "GBps": as_gbs,
                            "frac_of_hbm_peak": as_gbs / hbm_peak},
            "step": {"ms": k_ms[1], "env_steps_per_s": n / (k_ms[1] * 1e-3), "GBps": ALG_BYTES_STEP * n / (k_ms[1] * 1e-3) / 1e9},
            "reset_done": {"ms": k_ms[2]},
        },
        "e2e": {"value": n_total * 40 * e2e_steps / e2e_s, "unit": "afterstates/s", "env_steps_per_sec": n_total * e2e_steps / e2e_s,
                "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": 163 * n, "steps": e2e_steps,
                "api": "tpl_env_step_observe (host-buffer C ABI, pinned buffers)"},
        "e2e_features_on_device": {"value": n_total * 40 * e2e_steps / e2e_dev_s, "unit": "afterstates/s",
                                   "env_steps_per_sec": n_total * e2e_steps / e2e_dev_s, "h2d_bytes_per_step": 2 * n,
                                   "d2h_bytes_per_step": 3 * n, "steps": e2e_steps,
                                   "note": "same host call with feats=NULL: the 160 B/env of features stay in HBM for a policy on the "
                                           "GPU; not the headline e2e (that one ships every feature byte to the host)"},
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
    """N=1 extras: BASELINE configs[1] (4096 envs, latency-bound: looped in one CUDA graph), the fused rollouts, and
    the CPU baseline (C port on all host cores, bounded sample)."""
    out = {}
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
    n = args.envs_per_gpu
    env = tp.BatchedTetris(n, L_LINES, M_MOVES, device=dev, seed=SEED, config_pool=pool)
    env.reset()
    for name, fn, steps in (("rollout_random", lambda k: env.rollout_random(k), 64),
                            ("rollout_greedy", lambda k: env.rollout_greedy(k, [760, -360, -180, -510, 100000, -100000]), 16)):
        fn(2); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(steps); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"env_steps_per_s": n * steps / (ms * 1e-3), "ms": ms, "steps": steps}
        if name == "rollout_greedy":
            out[name]["afterstates_per_s"] = n * steps * 40 / (ms * 1e-3)
    cores = os.cpu_count() or 1
    ca, cm, cn, cs = cpu_c_port(pool, 10.0, cores)
    out["cpu_baseline"] = {"value": ca, "unit": "afterstates/s", "cores": cores, "kind": "port", "env_steps_per_sec": cm,
                           "sample": f"oracle/piclim_oracle.c (plain-C restatement) greedy rollout = 40 afterstate evals + 1 move "
                                     f"+ auto-reset per env-step, {cn} envs x {cs} steps, same pool/L/M, {cores} threads; the "
                                     f"reference's own Python speed is what --impl reference reports"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    args = ap.parse_args()
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
