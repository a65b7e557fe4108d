"""DQN afterstate-value training loop driving the batched GPU envs (BASELINE config 4: 65 536 envs).

The reference's ``model/train.py`` stops after building the network and the optimiser (``:25-27``); this is that
loop completed with the reference's own hyper-parameters (``model/train.py:15-21``), the optimiser it names
(``AdamW(lr=LR, amsgrad=True)``, ``:27``) and the standard DQN pieces those constants imply (replay memory,
exponentially decaying epsilon-greedy, Huber loss, soft target update with TAU).

Afterstate formulation: the action value of placing the current piece at slot (rot, loc) is
``r(slot) + GAMMA * V(afterstate(slot))``; ``V`` is ``ValueNet`` on the four features the env kernel emits.
Everything stays on the GPU: env state, the afterstate enumeration, the chosen move, auto-reset (one fused kernel,
``tpl_step_observe*``), the replay memory and the network.

Two rollout paths share the optimiser half:
* ``value_kernel=True`` (default): the env reports only the DISTINCT placements (``tpl_step_observe_distinct``: 23 instead of
  40 rows per env), the ranking forward pass over them is the fused tensor-core kernel (``tpl_value_rows``, weights and
  activations on chip) and the epsilon-greedy arg-max is ``tpl_select_action`` -- three launches per step, no [40 N, 128]
  activations in HBM.  The network, the loss and the optimiser are plain PyTorch.
* ``value_kernel=False``: the 40-slot form with the forward pass in PyTorch (``ValueNet.rank_bf16``), kept as the reference.
"""
from __future__ import annotations

import ctypes
import math
import time
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.optim as optim

from . import distinct as _distinct
from .batched import FLAG_ALIAS, FLAG_LOSE, FLAG_NOPIECE, FLAG_TOPOUT, FLAG_WIN, BatchedTetris
from .configs import ConfigPool, synthetic_pool
from .distinct import DISTINCT_MAX
from .model import ValueNet

# the reference's constants, model/train.py:15-21
BATCH_SIZE = 128
GAMMA = 0.99
EPS_START = 0.9
EPS_END = 0.05
EPS_DECAY = 1000
TAU = 0.005
LR = 1e-4


def reward_from(dlines: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """Reward as a pure function of the pinned integers the env reports: rows cleared by the move, +10 for
    reaching L lines, -10 for topping out or running out of moves (the reference has no reward)."""
    r = dlines.to(torch.float32)
    r = r + 10.0 * ((flags & FLAG_WIN) != 0).to(torch.float32)
    r = r - 10.0 * ((flags & (FLAG_LOSE | FLAG_TOPOUT)) != 0).to(torch.float32)
    return r


class ReplayMemory:
    """Ring buffer on the device: chosen afterstate (4 x u8), reward, done, and the next state's 40 afterstates
    (features + flags, u8) for the max in the TD target.  (``r``, the reward of the stored move itself, is kept for
    inspection only: the afterstate value V(x) bootstraps from the NEXT state's placements, whose rewards come from ``nx``.)  The 40-slot blocks are kept slot-major ([40, capacity, ...]),
    the layout the env kernel writes, so a push is 40 contiguous row copies and no transposition."""

    def __init__(self, capacity: int, device):
        self.capacity, self.size, self.pos = capacity, 0, 0
        self.x = torch.zeros((capacity, 4), dtype=torch.uint8, device=device)
        self.r = torch.zeros(capacity, dtype=torch.float32, device=device)
        self.done = torch.zeros(capacity, dtype=torch.bool, device=device)
        self.nx = torch.zeros((40, capacity, 4), dtype=torch.uint8, device=device)
        self.nfl = torch.zeros((40, capacity), dtype=torch.uint8, device=device)

    def push(self, x, r, done, nx, nfl):
        """x [n,4], r [n], done [n], nx [40,n,4], nfl [40,n]."""
        n = min(x.shape[0], self.capacity)
        first = min(n, self.capacity - self.pos)
        for lo, hi, at in ((0, first, self.pos), (first, n, 0)):
            if hi > lo:
                k = hi - lo
                self.x[at:at + k], self.r[at:at + k], self.done[at:at + k] = x[lo:hi], r[lo:hi], done[lo:hi]
                self.nx[:, at:at + k], self.nfl[:, at:at + k] = nx[:, lo:hi], nfl[:, lo:hi]
        self.pos = (self.pos + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch: int, gen, size_t: Optional[torch.Tensor] = None):
        """-> x [B,4], r [B], done [B], nx [B,40,4], nfl [B,40].  ``size_t`` (a device scalar holding ``self.size``)
        selects the CUDA-graph-safe index draw: no Python integer is baked into the captured kernels."""
        if size_t is None:
            idx = torch.randint(0, self.size, (batch,), device=self.x.device, generator=gen)
        else:
            idx = (torch.rand(batch, device=self.x.device) * size_t).long().clamp_(max=self.capacity - 1)
        return (self.x[idx], self.r[idx], self.done[idx], self.nx[:, idx].permute(1, 0, 2), self.nfl[:, idx].t())


class DistinctReplay:
    """Replay ring for the distinct-placements path, filled with ONE kernel per rollout step (``tpl_replay_push``) straight from
    the outputs of ``tpl_step_observe_distinct`` / ``tpl_select_action``, in the form the TD target consumes as is:
    ``x`` chosen placement (feature word, flags cleared: its four bytes are the features), ``r`` reward of the stored move (kept
    for inspection; the TD target does not need it), ``live`` 0 where the episode ended, and for the next state's placements ``nw`` [capacity, 34] feature words, ``nr`` their rewards (-inf past the
    end of the run) and ``ng`` = GAMMA where a placement does not end the episode (else 0): Q = nr + ng * V(nw)."""

    def __init__(self, capacity: int, device):
        self.capacity, self.size, self.pos = capacity, 0, 0
        self.x = torch.zeros(capacity, dtype=torch.int32, device=device)
        self.r = torch.zeros(capacity, dtype=torch.float32, device=device)
        self.live = torch.zeros(capacity, dtype=torch.float32, device=device)
        self.nw = torch.zeros((capacity, DISTINCT_MAX), dtype=torch.int32, device=device)
        self.nr = torch.zeros((capacity, DISTINCT_MAX), dtype=torch.float32, device=device)
        self.ng = torch.zeros((capacity, DISTINCT_MAX), dtype=torch.float32, device=device)

    def push_step(self, env, rows, runs, dlines, mflags, state, chosen) -> None:
        n = min(int(runs.numel()), self.capacity)
        p = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731
        env._call(env._L.tpl_replay_push, "tpl_replay_push", p(rows), p(runs), p(dlines), p(mflags), p(state), p(chosen), n, self.pos,
                  self.capacity, GAMMA, 10.0, -10.0, p(self.x), p(self.r), p(self.live), p(self.nw), p(self.nr), p(self.ng), env._stream())
        self.pos = (self.pos + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch: int, gen, size_t: Optional[torch.Tensor] = None):
        if size_t is None:
            idx = torch.randint(0, self.size, (batch,), device=self.x.device, generator=gen)
        else:
            idx = (torch.rand(batch, device=self.x.device) * size_t).long().clamp_(max=self.capacity - 1)
        return self.x[idx], self.live[idx], self.nw[idx], self.nr[idx], self.ng[idx]


def clean_word_features(words: torch.Tensor) -> torch.Tensor:
    """feature words with the flag bits cleared (int32) -> uint8 [..., 4] = (rows cleared, holes, bumpiness, aggregate height):
    a reinterpretation, no arithmetic"""
    return words.contiguous().view(torch.uint8).view(*words.shape, 4)


@dataclass
class TrainStats:
    env_steps: int = 0
    optim_steps: int = 0
    loss: float = float("nan")
    eps: float = EPS_START
    env_seconds: float = 0.0
    total_seconds: float = 0.0
    episodes: int = 0
    wins: int = 0

    @property
    def env_steps_per_s(self):
        return self.env_steps / max(self.env_seconds, 1e-9)

    @property
    def e2e_steps_per_s(self):
        return self.env_steps / max(self.total_seconds, 1e-9)


def select_slots(values: torch.Tensor, flags: torch.Tensor, eps, gen) -> torch.Tensor:
    """Epsilon-greedy over the distinct placements: values/flags are [40, N]; alias / no-piece slots are never
    chosen.  ``eps`` is a float or a device scalar tensor.  Returns the slot index int64[N]."""
    invalid = (flags & (FLAG_ALIAS | FLAG_NOPIECE)) != 0
    greedy = values.masked_fill(invalid, float("-inf")).argmax(dim=0)
    noise = torch.rand(values.shape, device=values.device, generator=gen).masked_fill(invalid, -1.0)
    explore = torch.rand(values.shape[1], device=values.device, generator=gen) < eps
    return torch.where(explore, noise.argmax(dim=0), greedy)


def train(num_envs: int = 65536, iterations: int = 200, L: int = 10, M: int = 30, device="cuda", seed: int = 0,
          config_pool: Optional[ConfigPool] = None, replay_capacity: int = 1 << 20, optim_steps_per_iter: int = 1,
          batch_size: int = BATCH_SIZE, log_every: int = 0, bf16_inference: bool = True, log_fn=None,
          cuda_graphs: bool = True, value_kernel: bool = True) -> tuple:
    """Run `iterations` env steps of all envs with `optim_steps_per_iter` optimiser steps each (default 1: one optimisation
    step per environment step, the structure of the DQN loop the reference's constants -- ``model/train.py:15-21`` -- come from).
    Returns (policy_net, TrainStats).  The action-selection forward pass over the 40 * num_envs afterstate rows runs
    under bf16 autocast (it only ranks slots); the optimiser step stays fp32.

    ``cuda_graphs``: the loop is launch-bound in eager PyTorch (about 370 small launches per iteration against 3.4 ms of
    GPU work at 65 536 envs), so after three eager iterations the two PyTorch blocks -- action selection, and the
    `optim_steps_per_iter` optimiser steps -- are captured into CUDA graphs and replayed; the env kernel
    (``tpl_step_observe``) and the replay push stay ordinary launches between them."""
    dev = torch.device(device)
    if value_kernel and dev.type == "cuda":
        return _train_value_kernel(num_envs, iterations, L, M, dev, seed, config_pool, replay_capacity, optim_steps_per_iter, batch_size,
                                   log_every, log_fn, cuda_graphs)
    pool = config_pool if config_pool is not None else synthetic_pool(4096, seed=seed, M=M)
    env = BatchedTetris(num_envs, L, M, device=dev, seed=seed, config_pool=pool)
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    torch.manual_seed(seed)
    policy_net = ValueNet().to(dev)
    target_net = ValueNet().to(dev)
    target_net.load_state_dict(policy_net.state_dict())
    use_graphs = bool(cuda_graphs) and dev.type == "cuda"
    # model/train.py:27 (fused: one launch; capturable: its step counter lives on the device, as graph capture needs)
    optimizer = optim.AdamW(policy_net.parameters(), lr=LR, amsgrad=True, fused=True, capturable=use_graphs)
    p_params, t_params = list(policy_net.parameters()), list(target_net.parameters())
    loss_fn = nn.SmoothL1Loss()
    memory = ReplayMemory(replay_capacity, dev)
    st = TrainStats()
    ar = torch.arange(num_envs, device=dev)
    t_all = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env_events, prev_stats = [], {}
    eps_t = torch.zeros((), device=dev)                     # epsilon and the replay fill level as device scalars: graph inputs
    size_t = torch.zeros((), device=dev)
    last_loss = torch.full((), float("nan"), device=dev)
    g_gen = None if use_graphs else gen                      # captured regions draw from the default CUDA generator

    feats, flags, ff = env.afterstates(f32=True, raw=True)  # slot-major: [40,N,4] u8, [40,N] u8, [40N,4] f32 (static buffers)
    act_out = {}

    def act_block():
        if bf16_inference and hasattr(torch, "_addmm_activation"):
            values = policy_net.rank_bf16(ff).float().view(40, num_envs)        # GEMMs with fused bias + ReLU epilogues
        else:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16_inference, cache_enabled=not use_graphs):
                values = policy_net(ff).float().view(40, num_envs)
        with torch.no_grad():
            slot = select_slots(reward_from(feats[..., 0], flags) + GAMMA * values, flags, eps_t, g_gen)
            rot, loc = (slot // 10).to(torch.uint8), (slot % 10).to(torch.uint8)
            x = feats[slot, ar].clone()                                             # chosen afterstate features [N,4]
        for k, v in (("rot", rot), ("loc", loc), ("x", x)):
            if k in act_out: act_out[k].copy_(v)
            else: act_out[k] = v.clone()

    def optim_block():
        for _ in range(optim_steps_per_iter):
            bx, br, bdone, bnx, bnfl = memory.sample(batch_size, g_gen, size_t if use_graphs else None)
            with torch.no_grad():
                nv = target_net(bnx.reshape(-1, 4)).view(batch_size, 40)
                nr = reward_from(bnx[..., 0], bnfl)
                q = (nr + GAMMA * nv * ((bnfl & (FLAG_WIN | FLAG_LOSE | FLAG_TOPOUT)) == 0)).masked_fill(
                    (bnfl & (FLAG_ALIAS | FLAG_NOPIECE)) != 0, float("-inf"))
                best_next = q.max(dim=1).values
                # V(afterstate) = value of the best continuation from the state it leads to (0 if terminal)
                target = torch.where(bdone, torch.zeros_like(best_next), best_next)
            loss = loss_fn(policy_net(bx), target)
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_value_(p_params, 100, foreach=True)
            optimizer.step()
            with torch.no_grad():                                               # soft update, TAU (two foreach launches)
                torch._foreach_mul_(t_params, 1 - TAU)
                torch._foreach_add_(t_params, p_params, alpha=TAU)
        last_loss.copy_(loss.detach())

    def run_or_capture(block, graph, it):
        """eager for the first iterations; then once more on a side stream (what capture wants) and captured; then replayed"""
        if graph is not None:
            graph.replay()
            return graph
        if not use_graphs or it < 3:
            block()
            return None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            block()
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            block()
        return g

    g_act = g_opt = None
    for it in range(iterations):
        eps = EPS_END + (EPS_START - EPS_END) * math.exp(-1.0 * it / EPS_DECAY)
        eps_t.fill_(eps)
        g_act = run_or_capture(act_block, g_act, it)
        x = act_out["x"]
        ev0.record()
        # one launch: move -> auto-reset of finished episodes -> afterstates of the new states (tpl_step_observe)
        dlines, mflags, state, feats2, flags2, ff2 = env.step_observe(act_out["rot"], act_out["loc"], packed=False, f32=True)
        assert feats2.data_ptr() == feats.data_ptr() and ff2.data_ptr() == ff.data_ptr()     # the env reuses its output buffers
        reward = reward_from(dlines, mflags)
        done = state != 0
        ev1.record()
        memory.push(x, reward, done, feats, flags)
        size_t.fill_(float(memory.size))
        st.env_steps += num_envs
        if memory.size >= batch_size:
            g_opt = run_or_capture(optim_block, g_opt, it)
            st.optim_steps += optim_steps_per_iter
        env_events.append((ev0, ev1))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.eps = eps
        if log_every and (it + 1) % log_every == 0:
            s = env.reduce_stats()
            st.loss = float(last_loss.item())
            if log_fn is not None:
                log_fn(it + 1, eps, st.loss, s, prev_stats)
                prev_stats = dict(s)
                continue
            print(f"it {it + 1}: eps {eps:.3f} loss {st.loss:.4f} episodes {s['episodes']} wins {s['wins']} "
                  f"lines/episode {s['lines'] / max(s['episodes'], 1):.2f}")
    torch.cuda.synchronize(dev)
    st.total_seconds = time.perf_counter() - t_all
    st.env_seconds = sum(a.elapsed_time(b) for a, b in env_events) * 1e-3
    st.loss = float(last_loss.item())
    s = env.reduce_stats()
    st.episodes, st.wins = s["episodes"], s["wins"]
    return policy_net, st


def _train_value_kernel(num_envs, iterations, L, M, dev, seed, config_pool, replay_capacity, optim_steps_per_iter, batch_size, log_every,
                        log_fn, cuda_graphs):
    """The rollout half on the library's own kernels: per step  pack weights -> tpl_value_rows (tensor cores) -> tpl_select_action ->
    tpl_step_observe_distinct;  the optimiser half (replay sample, TD target with the target net, Huber loss, AdamW, soft update)
    in PyTorch, captured into one CUDA graph after three eager iterations."""
    from .value_kernel import ValueKernel
    pool = config_pool if config_pool is not None else synthetic_pool(4096, seed=seed, M=M)
    env = BatchedTetris(num_envs, L, M, device=dev, seed=seed, config_pool=pool)
    env.reset()
    torch.manual_seed(seed)
    policy_net = ValueNet().to(dev)
    target_net = ValueNet().to(dev)
    target_net.load_state_dict(policy_net.state_dict())
    use_graphs = bool(cuda_graphs)
    optimizer = optim.AdamW(policy_net.parameters(), lr=LR, amsgrad=True, fused=True, capturable=use_graphs)        # model/train.py:27
    p_params, t_params = list(policy_net.parameters()), list(target_net.parameters())
    loss_fn = nn.SmoothL1Loss()
    memory = DistinctReplay(replay_capacity, dev)
    vk = ValueKernel(policy_net)
    st = TrainStats()
    size_t = torch.zeros((), device=dev)
    last_loss = torch.full((), float("nan"), device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    g_gen = None if use_graphs else gen

    def optim_block():
        for _ in range(optim_steps_per_iter):
            bx, blive, bnw, bnr, bng = memory.sample(batch_size, g_gen, size_t if use_graphs else None)
            with torch.no_grad():
                nv = target_net(clean_word_features(bnw).reshape(-1, 4)).view(batch_size, DISTINCT_MAX)
                # V(afterstate) = value of the best continuation from the state it leads to (0 if terminal or nothing to place):
                # Q(placement) = reward + GAMMA * V unless the placement ends the episode; padding carries -inf
                target = torch.addcmul(bnr, bng, nv).max(dim=1).values * blive
            loss = loss_fn(policy_net(clean_word_features(bx)), target)
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_value_(p_params, 100, foreach=True)
            optimizer.step()
            with torch.no_grad():                                               # soft update, TAU
                torch._foreach_mul_(t_params, 1 - TAU)
                torch._foreach_add_(t_params, p_params, alpha=TAU)
        last_loss.copy_(loss.detach())

    rows, runs, used = env.afterstates_distinct()
    vals = torch.zeros(rows.numel(), dtype=torch.float32, device=dev)
    sel = None
    g_opt = None
    # The optimiser half runs on its own stream, overlapped with the next rollout step: while the tensor-core kernel ranks the
    # placements of step k + 1, the (many, tiny) PyTorch kernels of optimiser block k fill the gaps.  Step k + 1 therefore acts
    # with the parameters block k - 1 produced -- the usual one-step policy lag of an actor / learner split.  Ordering: before
    # the transitions of step k are pushed, block k - 1 must be complete (it samples the memory); then the weights it produced
    # are packed for the kernel (vk.sync); then block k starts (it waits for the push and the pack).
    main = torch.cuda.current_stream(dev)
    opt_stream = torch.cuda.Stream(device=dev)
    opt_done = torch.cuda.Event()
    opt_done.record(main)
    t_all = time.perf_counter()
    env_events, prev_stats = [], {}
    for it in range(iterations):
        eps = EPS_END + (EPS_START - EPS_END) * math.exp(-1.0 * it / EPS_DECAY)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        vk.values(rows, used.reshape(1), out=vals)                                # (weights packed at the end of the previous iteration)
        sel = vk.select(rows, runs, vals, GAMMA, eps, seed, it, out=sel)
        rot, loc, chosen, _ = sel
        ev0.record()
        dlines, mflags, state, rows, runs, used = env.step_observe_distinct(rot, loc)
        ev1.record()
        main.wait_event(opt_done)                                                 # block it - 1 is complete: it no longer samples the memory ...
        memory.push_step(env, rows, runs, dlines, mflags, state, chosen)
        vk.sync()                                                                 # ... and its parameters are the ones the next step ranks with
        st.env_steps += num_envs
        if memory.size >= batch_size:
            opt_stream.wait_stream(main)                                          # the pack has read the weights; the push is in memory
            with torch.cuda.stream(opt_stream):
                size_t.fill_(float(memory.size))
                if g_opt is not None:
                    g_opt.replay()
                elif not use_graphs or it < 3:
                    optim_block()
                else:
                    optim_block()                                                 # once more outside capture (what capture wants), then captured
                    g_opt = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_opt, stream=opt_stream):
                        optim_block()
                opt_done.record(opt_stream)
            st.optim_steps += optim_steps_per_iter
        env_events.append((ev0, ev1))
        st.eps = eps
        if log_every and (it + 1) % log_every == 0:
            torch.cuda.synchronize(dev)
            s = env.reduce_stats()
            st.loss = float(last_loss.item())
            if log_fn is not None:
                log_fn(it + 1, eps, st.loss, s, prev_stats)
                prev_stats = dict(s)
                continue
            print(f"it {it + 1}: eps {eps:.3f} loss {st.loss:.4f} episodes {s['episodes']} wins {s['wins']} "
                  f"lines/episode {s['lines'] / max(s['episodes'], 1):.2f}")
    torch.cuda.synchronize(dev)
    st.total_seconds = time.perf_counter() - t_all
    st.env_seconds = sum(a.elapsed_time(b) for a, b in env_events) * 1e-3
    st.loss = float(last_loss.item())
    s = env.reduce_stats()
    st.episodes, st.wins = s["episodes"], s["wins"]
    return policy_net, st


if __name__ == "__main__":
    net, stats = train(log_every=20)
    print(stats, f"env-only {stats.env_steps_per_s:.3e} steps/s, end-to-end {stats.e2e_steps_per_s:.3e} steps/s")
