"""``BatchedTetris``: N Tetris-piclim envs resident on one B200, driven through the C ABI.

Same verbs as the reference's ``Tetris`` (``game/tetris.py:140``: ``reset`` ``:438``, ``move`` ``:354``,
``get_state`` ``:435``, ``terminate`` ``:451``) on tensors instead of one board, plus ``afterstates()`` and the
fused rollouts.  PyTorch is used for device memory and streams only; every computation is a kernel of
``libtetris_piclim_sm100.so`` queued on torch's current stream.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import distinct as _distinct
from .configs import MAX_PIECES, ConfigPool

RUNNING, WON, LOST = 0, 1, 2
FLAG_TOPOUT, FLAG_WIN, FLAG_LOSE, FLAG_ALIAS, FLAG_NOPIECE = 1, 2, 4, 8, 16
RESET_ALL, RESET_MASK, RESET_DONE = 0, 1, 2

STAT_NAMES = ("episodes", "wins", "topouts", "movelimit_losses", "lines", "moves", "steps", "resets")


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class BatchedTetris:
    """N envs with shared (L, M).  ``env_base`` is the global id of env 0 (multi-GPU sharding keeps the RNG
    streams a function of the *global* env id, so results do not depend on the number of ranks).
    ``gen_pieces`` > 0: every reset replaces the pool's piece list by that many pieces of the counter-based 7-bag sequence
    of (seed, env, episode); it may exceed the 42 pieces the env record holds (e.g. M + 1 = 61): the kernels that know the RNG
    keys (``step_observe*``, the rollouts, ``reset(done_only=True)``) refill the queue when it runs dry mid-episode.  After a
    plain ``move`` call ``reset(done_only=True)`` to get the refill (``move`` alone reports FLAG_NOPIECE on an empty queue)."""

    def __init__(self, num_envs: int, L: int, M: int, device="cuda", seed: int = 0,
                 config_pool: Optional[ConfigPool] = None, env_base: int = 0, gen_pieces: int = 0):
        if num_envs <= 0:
            raise ValueError("num_envs must be positive")
        if not (0 <= L <= 65535 and 0 <= M <= 65535):
            raise ValueError("L and M must fit 16 bits")
        if not 0 <= int(gen_pieces) <= 42 * 256:
            raise ValueError("gen_pieces must be in 0..10752")
        self._L = _lib.lib()                                    # raises if the CUDA library is missing
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedTetris is CUDA-only (B200, sm_100a); there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs, self.L, self.M = int(num_envs), int(L), int(M)
        self.seed, self.env_base, self.gen_count = int(seed), int(env_base), int(gen_pieces)
        self.stride = (self.num_envs + 31) // 32 * 32
        with torch.cuda.device(self.device):
            self.state = torch.zeros((4, self.stride, 4), dtype=torch.int32, device=self.device)
            self.episode = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
            self.tstep = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
            self.stats = torch.zeros(8, dtype=torch.int64, device=self.device)
        self.pool = None
        self.pool_size = 0
        self.count_stats = True          # move() accumulates episode statistics into self.stats
        self._out = {}
        if config_pool is not None:
            self.set_pool(config_pool)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, fn, what, *args):
        """Run one C-ABI entry point with this env's device current (the launches inside use the CUDA *current* device;
        an env on cuda:1 driven while cuda:0 is current would otherwise fail with an invalid resource handle)."""
        if torch.cuda.current_device() == self.device.index:
            _lib.check(fn(*args), what)
        else:
            with torch.cuda.device(self.device):
                _lib.check(fn(*args), what)

    def _dev(self, x, dtype, shape=None) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.device, dtype=dtype)
        else:
            a = np.ascontiguousarray(x)
            if dtype == torch.uint16:
                t = torch.from_numpy(a.astype(np.uint16).view(np.int16)).to(self.device).view(torch.uint16)
            else:
                t = torch.as_tensor(a).to(device=self.device, dtype=dtype)
        t = t.contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def _buf(self, name, shape, dtype) -> torch.Tensor:
        t = self._out.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            if name == "feats":
                # the kernels advance their store pointer with a 32-bit add when the array does not cross a 4 GB-aligned
                # address boundary (GlobalSink<.., P32>): if this allocation happens to straddle one, take another
                held = []
                while (t.data_ptr() >> 32) != ((t.data_ptr() + t.numel() * t.element_size() - 1) >> 32) and len(held) < 4:
                    held.append(t)
                    t = torch.empty(shape, dtype=dtype, device=self.device)
                del held
            self._out[name] = t
        return t

    # ------------------------------------------------------------------ config pool / reset (E)
    def set_pool(self, pool: ConfigPool) -> None:
        rows = np.ascontiguousarray(pool.rows, np.uint16)
        K = rows.shape[0]
        if rows.shape != (K, 20) or K == 0:
            raise ValueError("pool rows must be uint16[K, 20] with K > 0")
        pieces = np.ascontiguousarray(pool.pieces, np.uint8)
        npieces = np.ascontiguousarray(pool.npieces, np.uint8)
        if pieces.ndim != 2 or pieces.shape[0] != K or npieces.shape != (K,):
            raise ValueError("pool pieces must be uint8[K, P] and npieces uint8[K]")
        if int(npieces.max()) > min(MAX_PIECES, pieces.shape[1]) or int(npieces.min()) < 1:
            raise ValueError("each config needs 1..42 pieces (and no more than the pieces array holds)")
        if int(pieces.max()) > 6:
            raise ValueError("piece ids must be 0..6")
        d_rows, d_p, d_np = self._dev(rows, torch.uint16), self._dev(pieces, torch.uint8), self._dev(npieces, torch.uint8)
        self.pool = torch.empty((K, 4, 4), dtype=torch.int32, device=self.device)
        self._call(self._L.tpl_pack, "tpl_pack(pool)", _ptr(self.pool), 0, 1, K, _ptr(d_rows), _ptr(d_p), pieces.shape[1], _ptr(d_np),
                                    None, None, None, None, self._stream())
        self.pool_size = K

    def reset(self, mask=None, idx=None, boards=None, pieces=None, npieces=None, done_only: bool = False) -> None:
        """Prescribed reset (``game/tetris.py:438-449``) with ctor-fresh counters (``:149-151``).

        * ``boards``/``pieces`` given: install exactly these (uint16[N,20] bitrows or bool[N,20,10]; uint8[N,P]).
        * otherwise draw from the pool: ``idx`` int32[N] picks configs, else the counter RNG keyed by
          (seed, global env id, episode) does.  ``mask`` restricts the reset to some envs; ``done_only`` resets
          exactly the envs whose episode has ended (auto-reset).  A reset that draws its configs (no ``idx``) from a
          mask or with ``done_only`` starts a new episode: the per-env episode counter is bumped before the draw.
          Every reset zeroes the reset envs' ``tstep`` (the rollouts' per-episode action counter)."""
        n = self.num_envs
        if boards is not None:
            if mask is not None or done_only:
                raise ValueError("explicit boards reset every env")
            self.load(boards, pieces, npieces)
            self.episode.zero_(); self.tstep.zero_()
            return
        if self.pool is None:
            raise RuntimeError("reset() needs a config pool (set_pool) or explicit boards")
        mode = RESET_DONE if done_only else (RESET_MASK if mask is not None else RESET_ALL)
        d_mask = self._dev(mask, torch.uint8, (n,)) if mask is not None else None
        d_idx = self._dev(idx, torch.int32, (n,)) if idx is not None else None
        if mode == RESET_ALL:
            self.episode.zero_()
        # the kernel zeroes tstep of every env it resets and bumps the episode counter of the envs whose config it draws
        # itself (done_only, or mask without idx): `env.reset(mask=done)` gives each env a NEW config, not the old one again
        self._call(self._L.tpl_reset_from_pool, "tpl_reset_from_pool", _ptr(self.state), self.stride, n, _ptr(self.pool), self.pool_size,
                                               _ptr(d_idx), _ptr(d_mask), mode, _ptr(self.episode), _ptr(self.tstep), self.seed,
                                               self.env_base, self.gen_count, self._stream())

    def load(self, boards, pieces, npieces=None, lines=None, moves=None, state=None, head=None) -> None:
        """Install explicit env states (boundary format).  Optional counters let tests resume mid-episode."""
        n = self.num_envs
        b = boards.detach().cpu().numpy() if isinstance(boards, torch.Tensor) else np.asarray(boards)
        if b.dtype == bool or b.ndim == 3:
            from .configs import rows_from_bool
            b = rows_from_bool(b)
        d_rows = self._dev(b, torch.uint16, (n, 20))
        p = pieces.detach().cpu().numpy() if isinstance(pieces, torch.Tensor) else np.asarray(pieces)
        if p.ndim != 2 or p.shape[0] != n:
            raise ValueError("pieces must be [N, P]")
        if npieces is None:
            npieces = np.full(n, p.shape[1], np.uint8)
        npn = np.asarray(npieces)
        if npn.max() > min(MAX_PIECES, p.shape[1]):
            raise ValueError("npieces exceeds the 42-piece queue or the pieces array")
        if p.size and (int(p.max()) > 6 or int(p.min()) < 0):
            raise ValueError("piece ids must be 0..6")
        d_p, d_np = self._dev(p, torch.uint8), self._dev(npn, torch.uint8, (n,))
        opt = lambda x, dt: self._dev(x, dt, (n,)) if x is not None else None   # noqa: E731
        d_lines, d_moves, d_st, d_head = opt(lines, torch.int32), opt(moves, torch.int32), opt(state, torch.int8), opt(head, torch.uint8)
        self._call(self._L.tpl_pack, "tpl_pack", _ptr(self.state), self.stride, 0, n, _ptr(d_rows), _ptr(d_p), p.shape[1], _ptr(d_np),
                                    _ptr(d_lines), _ptr(d_moves), _ptr(d_st), _ptr(d_head), self._stream())

    # ------------------------------------------------------------------ move (C)
    def move(self, rot, loc) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """``Tetris.move(rotations, location)`` (``game/tetris.py:354-422``) on every env.

        ``rot`` may be any integers (reduced with Python's ``%`` like ``:61``); ``loc`` must be >= 0 (a negative
        location raises in the reference too) and is clamped to ``10 - width`` (``:364``).
        Returns (rows cleared int8[N], flags uint8[N], state int8[N]).  The three tensors are this object's output buffers:
        valid until the next ``move`` / ``step_observe`` call, which overwrites them in place (``.clone()`` to keep)."""
        n = self.num_envs
        if isinstance(rot, torch.Tensor):
            # uint8 tensors go straight to the kernel (it applies rot & 3, and 256 % 4 == 0)
            r = rot if rot.dtype == torch.uint8 else torch.remainder(rot.to(self.device), 4).to(torch.uint8)
        else:
            r = np.mod(np.asarray(rot, np.int64), 4).astype(np.uint8)
        if isinstance(loc, torch.Tensor):
            if loc.dtype == torch.uint8:
                l = loc
            else:
                if bool((loc < 0).any()):
                    raise ValueError("location must be >= 0")
                l = loc.to(self.device).clamp(max=255).to(torch.uint8)
        else:
            la = np.asarray(loc, np.int64)
            if (la < 0).any():
                raise ValueError("location must be >= 0")
            l = np.minimum(la, 255).astype(np.uint8)
        d_rot, d_loc = self._dev(r, torch.uint8, (n,)), self._dev(l, torch.uint8, (n,))
        dl = self._buf("dlines", (n,), torch.int8)
        fl = self._buf("mflags", (n,), torch.uint8)
        st = self._buf("mstate", (n,), torch.int8)
        self._call(self._L.tpl_step, "tpl_step", _ptr(self.state), self.stride, n, _ptr(d_rot), _ptr(d_loc), _ptr(dl), _ptr(fl), _ptr(st),
                                    _ptr(self.stats) if self.count_stats else None, self.L, self.M, self._stream())
        return dl, fl, st

    step = move

    # ------------------------------------------------------------------ observation (D)
    def get_state(self, bool_boards: bool = False):
        """Batched ``Tetris.get_state`` (``game/tetris.py:435-436``):
        (boards uint16[N,20] or bool[N,20,10], current piece uint8[N], next piece uint8[N] (255 = none),
        L - lines_cleared int32[N], M - moves_used int32[N], state int8[N]).  Boards / pieces / state alias reused
        output buffers (valid until the next ``get_state`` / ``fields`` call)."""
        f = self.fields()
        boards = f["rows"]
        if bool_boards:
            boards = ((boards.to(torch.int32)[..., None] >> torch.arange(10, device=self.device)) & 1).bool()
        return boards, f["cur"], f["next"], self.L - f["lines"], self.M - f["moves"], f["state"]

    def fields(self, queue: bool = False) -> dict:
        """Raw per-env fields (boundary format).  The tensors are reused output buffers: valid until the next ``fields`` /
        ``get_state`` call."""
        n = self.num_envs
        rows = self._buf("rows", (n, 20), torch.uint16)
        cur, nxt = self._buf("cur", (n,), torch.uint8), self._buf("next", (n,), torch.uint8)
        lines, moves = self._buf("lines", (n,), torch.int32), self._buf("moves", (n,), torch.int32)
        st, head, npc = self._buf("state", (n,), torch.int8), self._buf("head", (n,), torch.uint8), self._buf("np", (n,), torch.uint8)
        q = self._buf("queue", (n, MAX_PIECES), torch.uint8) if queue else None
        self._call(self._L.tpl_unpack, "tpl_unpack", _ptr(self.state), self.stride, n, _ptr(rows), _ptr(cur), _ptr(nxt), _ptr(lines), _ptr(moves),
                                      _ptr(st), _ptr(head), _ptr(npc), _ptr(q), self._stream())
        out = dict(rows=rows, cur=cur, next=nxt, lines=lines, moves=moves, state=st, head=head, npieces=npc)
        if queue:
            out["queue"] = q
        return out

    # ------------------------------------------------------------------ afterstates (F)
    def afterstates(self, f32: bool = False, u8: bool = True, packed: bool = False, raw: bool = False):
        """All 40 afterstates of every env: slot (r, c) == clone(env).move(r, c).

        Returns (feats, flags[, feats_f32]): ``feats`` uint8 viewed as [N, 4, 10, 4] = (rows cleared, holes,
        bumpiness, aggregate height), ``flags`` uint8 [N, 4, 10]; both are strided views of slot-major buffers
        ([40, N, 4] / [40, N]) so no copy is made.  ``feats_f32`` (if requested) is float32 [40*N, 4] in
        slot-major row order (row s*N + i), ready to be fed to the value net.

        ``packed=True`` is the compact 160 B/env form: only ``feats`` is written and its byte 0 holds
        ``rows cleared | flags << 3``; returns (feats, None).  ``raw=True`` returns the slot-major buffers themselves
        (uint8 [40, N, 4] and [40, N]) instead of the [N, 4, 10, ...] views.  All results alias this object's output buffers:
        the next ``afterstates`` / ``step_observe`` call overwrites them in place."""
        n = self.num_envs
        if packed and f32:
            raise ValueError("the packed form has no float output")
        feats = self._buf("feats", (40, n, 4), torch.uint8) if (u8 or packed) else None
        flags = None if packed else self._buf("aflags", (40, n), torch.uint8)
        ff = self._buf("feats_f32", (40 * n, 4), torch.float32) if f32 else None
        self._call(self._L.tpl_afterstates, "tpl_afterstates", _ptr(self.state), self.stride, n, _ptr(feats), _ptr(flags), _ptr(ff), self.L, self.M,
                                           self._stream())
        if raw:
            return (feats, flags, ff) if f32 else (feats, flags)
        fv = feats.view(4, 10, n, 4).permute(2, 0, 1, 3) if feats is not None else None
        gv = flags.view(4, 10, n).permute(2, 0, 1) if flags is not None else None
        return (fv, gv, ff) if f32 else (fv, gv)

    # ------------------------------------------------------------------ fused hot-path step
    def step_observe(self, rot, loc, packed: bool = True, f32: bool = False, auto_reset: bool = True):
        """move -> auto-reset of finished envs -> afterstates of the resulting state, in ONE kernel launch
        (``tpl_step_observe``): the rollout inner loop between two value-net calls.  ``rot``/``loc`` must be uint8
        CUDA tensors [N].  Returns (rows cleared, move flags, state after the move, feats, afterstate flags or None,
        feats_f32 or None); ``feats`` is the raw slot-major uint8 [40, N, 4] buffer (compact form when ``packed``).
        All results alias this object's output buffers (valid until the next call that writes them)."""
        n = self.num_envs
        if auto_reset:
            self._need_pool()
        if packed and f32:
            raise ValueError("the packed form has no float output")
        d_rot, d_loc = self._dev(rot, torch.uint8, (n,)), self._dev(loc, torch.uint8, (n,))
        dl, fl, st = self._buf("dlines", (n,), torch.int8), self._buf("mflags", (n,), torch.uint8), self._buf("mstate", (n,), torch.int8)
        feats = self._buf("feats", (40, n, 4), torch.uint8)
        aflags = None if packed else self._buf("aflags", (40, n), torch.uint8)
        ff = self._buf("feats_f32", (40 * n, 4), torch.float32) if f32 else None
        self._call(self._L.tpl_step_observe, "tpl_step_observe", _ptr(self.state), self.stride, n, _ptr(d_rot), _ptr(d_loc), _ptr(dl), _ptr(fl), _ptr(st),
                                            _ptr(self.stats) if self.count_stats else None,
                                            _ptr(self.pool) if auto_reset else None, self.pool_size, _ptr(self.episode), _ptr(self.tstep),
                                            self.seed, self.env_base, self.gen_count, _ptr(feats), _ptr(aflags), _ptr(ff), self.L, self.M,
                                            self._stream())
        return dl, fl, st, feats, aflags, ff

    # ------------------------------------------------------------------ distinct-placements (alias-free) forms
    def _distinct_bufs(self):
        n = self.num_envs
        rows = self._buf("drows", (_distinct.capacity(n),), torch.int32)
        runs = self._buf("druns", (n,), torch.int32)
        if "dcursor" not in self._out:
            self._out["dcursor"] = torch.zeros(2, dtype=torch.int32, device=self.device)
            self._phase = 0
        phase = self._phase
        self._phase ^= 1
        return rows, runs, self._out["dcursor"], phase

    def afterstates_distinct(self):
        """The afterstates of every env in the distinct-placements form (``tpl_afterstates_distinct``): only the placements
        that differ -- 9 / 17 / 34 per env instead of 40 aliased slots.  Returns (rows int32[capacity], runs int32[N],
        used): ``rows`` holds one word per placement (byte 0 = rows cleared | flags << 3, holes, bumpiness, aggregate
        height; view it as uint8 [-1, 4]), ``runs[i]`` = word offset of env i's run | piece << 29 (``distinct.run_offset`` /
        ``run_piece`` / ``gather_index`` / ``expand``), ``used`` a 0-d device tensor = words of ``rows`` in use.
        Valid until the next distinct call (the buffers are reused)."""
        n = self.num_envs
        rows, runs, cur, phase = self._distinct_bufs()
        self._call(self._L.tpl_afterstates_distinct, "tpl_afterstates_distinct", _ptr(self.state), self.stride, n, _ptr(rows), rows.numel(),
                   _ptr(runs), 0, _ptr(cur), phase, self.L, self.M, self._stream())
        return rows, runs, cur[phase]

    def step_observe_distinct(self, rot, loc, auto_reset: bool = True):
        """``step_observe`` with the new states' afterstates in the distinct-placements form (``tpl_step_observe_distinct``).
        Returns (rows cleared, move flags, state after the move, rows, runs, used) -- see ``afterstates_distinct``."""
        n = self.num_envs
        if auto_reset:
            self._need_pool()
        d_rot, d_loc = self._dev(rot, torch.uint8, (n,)), self._dev(loc, torch.uint8, (n,))
        dl, fl, st = self._buf("dlines", (n,), torch.int8), self._buf("mflags", (n,), torch.uint8), self._buf("mstate", (n,), torch.int8)
        rows, runs, cur, phase = self._distinct_bufs()
        self._call(self._L.tpl_step_observe_distinct, "tpl_step_observe_distinct", _ptr(self.state), self.stride, n, _ptr(d_rot), _ptr(d_loc),
                   _ptr(dl), _ptr(fl), _ptr(st), _ptr(self.stats) if self.count_stats else None,
                   _ptr(self.pool) if auto_reset else None, self.pool_size, _ptr(self.episode), _ptr(self.tstep), self.seed, self.env_base,
                   self.gen_count, _ptr(rows), rows.numel(), _ptr(runs), 0, _ptr(cur), phase, self.L, self.M, self._stream())
        return dl, fl, st, rows, runs, cur[phase]

    def expand_distinct(self, rows, runs):
        """rows / runs of a distinct call -> the compact 40-slot form uint8 [40, N, 4] (``tpl_expand_distinct``)."""
        n = self.num_envs
        out = torch.empty((40, n, 4), dtype=torch.uint8, device=self.device)
        self._call(self._L.tpl_expand_distinct, "tpl_expand_distinct", _ptr(rows), _ptr(runs), n, _ptr(out), self._stream())
        return out

    # ------------------------------------------------------------------ fused rollouts
    def _need_pool(self):
        if self.pool is None:
            raise RuntimeError("rollouts auto-reset from a config pool: call set_pool first")

    def rollout_random(self, steps: int) -> None:
        """``steps`` uniformly random moves per env with auto-reset; state stays in registers between moves.
        Episode statistics accumulate in ``self.stats`` (see STAT_NAMES)."""
        self._need_pool()
        self._call(self._L.tpl_rollout_random, "tpl_rollout_random", _ptr(self.state), self.stride, self.num_envs, _ptr(self.pool), self.pool_size,
                                              _ptr(self.episode), _ptr(self.tstep), _ptr(self.stats), int(steps), self.seed,
                                              self.env_base, self.gen_count, self.L, self.M, self._stream())

    def rollout_greedy(self, steps: int, weights: Sequence[int]) -> None:
        """``steps`` greedy moves per env: arg-max over the 40 afterstates of the integer linear value
        w0*dlines + w1*holes + w2*bumpiness + w3*agg_height (+ w4 on a win, + w5 on a loss / top-out)."""
        self._need_pool()
        w = (ctypes.c_int32 * 6)(*[int(x) for x in weights])
        self._call(self._L.tpl_rollout_greedy, "tpl_rollout_greedy", _ptr(self.state), self.stride, self.num_envs, _ptr(self.pool), self.pool_size,
                                              _ptr(self.episode), _ptr(self.tstep), _ptr(self.stats), int(steps),
                                              ctypes.cast(w, ctypes.c_void_p), self.seed, self.env_base, self.gen_count,
                                              self.L, self.M, self._stream())

    def gen_pieces(self, count: int, episode: int = 0) -> torch.Tensor:
        out = torch.empty((self.num_envs, count), dtype=torch.uint8, device=self.device)
        self._call(self._L.tpl_gen_pieces, "tpl_gen_pieces", _ptr(out), self.num_envs, count, self.seed, self.env_base, None, episode, self._stream())
        return out

    def reduce_stats(self) -> dict:
        """Episode statistics, summed over all ranks when torch.distributed is initialised (the one collective
        of the path: a 64-byte all-reduce per rollout)."""
        s = self.stats.clone()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
        return dict(zip(STAT_NAMES, (int(v) for v in s.tolist())))

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self) -> dict:
        """Everything a rollout continues from: the env records, per-env episode / step counters, the statistics and the
        keys of the counter RNG.  (The reference has no persistence; its env state is Python objects.  Here it is four flat
        tensors, so ``torch.save(env.state_dict(), path)`` is a checkpoint.)  The config pool is not included."""
        return {"state": self.state.clone(), "episode": self.episode.clone(), "tstep": self.tstep.clone(),
                "stats": self.stats.clone(),
                "meta": {"num_envs": self.num_envs, "L": self.L, "M": self.M, "seed": self.seed, "env_base": self.env_base,
                         "gen_count": self.gen_count, "stride": self.stride}}

    def load_state_dict(self, d: dict) -> None:
        m = d["meta"]
        if (m["num_envs"], m["stride"]) != (self.num_envs, self.stride):
            raise ValueError(f"checkpoint holds {m['num_envs']} envs, this object {self.num_envs}")
        self.L, self.M, self.seed = int(m["L"]), int(m["M"]), int(m["seed"])
        self.env_base, self.gen_count = int(m["env_base"]), int(m["gen_count"])
        for k in ("state", "episode", "tstep", "stats"):
            getattr(self, k).copy_(d[k].to(self.device))

    def terminate(self) -> None:
        """API parity with ``Tetris.terminate`` (``game/tetris.py:451``): there are no worker processes to join."""
        return None
