"""Two interchangeable engines behind one numpy-level interface, used by the parity cases:

* ``GpuEngine``  -- the product: torch CUDA tensors + the C ABI of libtetris_piclim_sm100.so (``-m gpu`` tests).
* ``EmulEngine`` -- tests/emul/libpiclim_emul.so: the same per-env device code compiled for the host by g++
  (CPU tier; test infrastructure only, never used by the product).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG_DIR = os.path.join(ROOT, "reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200")


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


class States:
    def __init__(self, n, planes, stride):
        self.n, self.planes, self.stride = n, planes, stride


class EmulEngine:
    name = "emul"

    def __init__(self):
        src = os.path.join(HERE, "emul", "emul.cpp")
        so = os.path.join(HERE, "emul", "libpiclim_emul.so")
        deps = [src, os.path.join(HERE, "emul", "host_shim.h"), os.path.join(PKG_DIR, "csrc", "piclim_core.cuh"),
                os.path.join(PKG_DIR, "csrc", "piclim_env.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
        self.L = ctypes.CDLL(so)
        P, I, I64, U64, U32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32
        sigs = {
            "pack": [P, I64, I, I, P, P, I, P, P, P, P, P],
            "unpack": [P, I64, I, P, P, P, P, P, P, P, P, P],
            "reset_from_pool": [P, I64, I, P, I, P, P, I, P, P, U64, U64, I],
            "afterstates_distinct": [P, I64, I, P, P, U32, P, I, I, I],
            "step": [P, I64, I, P, P, P, P, P, I, I],
            "afterstates": [P, I64, I, P, P, P, I, I],
            "afterstates_uniform": [P, I64, I, P, I, I, I],
            "afterstates_split": [P, I64, I, P, I, I],
            "afterstates_cursor": [P, I64, I, P, I, I],
            "gen_pieces": [P, I, I, U64, U64, P, U32],
            "step_observe": [P, I64, I, P, P, P, P, P, P, P, I, P, P, U64, U64, I, P, P, P, I, I, P, P, P],
            "rollout_random": [P, I64, I, P, I, P, P, P, I, U64, U64, I, I, I],
            "rollout_greedy": [P, I64, I, P, I, P, P, P, I, P, U64, U64, I, I, I],
        }
        for k, a in sigs.items():
            f = getattr(self.L, "emul_" + k)
            f.restype = I
            f.argtypes = a

    # -- helpers
    def _c(self, x, dt):
        return np.ascontiguousarray(x, dt) if x is not None else None

    def pack(self, rows, pieces, npieces, lines=None, moves=None, state=None, head=None):
        rows = self._c(rows, np.uint16); n = rows.shape[0]
        pieces = self._c(pieces, np.uint8); npieces = self._c(npieces, np.uint8)
        stride = (n + 31) // 32 * 32
        planes = np.zeros((4, stride, 4), np.uint32)
        a = [self._c(lines, np.int32), self._c(moves, np.int32), self._c(state, np.int8), self._c(head, np.uint8)]
        self.L.emul_pack(_np_ptr(planes), stride, 0, n, _np_ptr(rows), _np_ptr(pieces), pieces.shape[1], _np_ptr(npieces),
                         *[_np_ptr(x) for x in a])
        return States(n, planes, stride)

    def make_pool(self, rows, pieces, npieces):
        rows = self._c(rows, np.uint16); K = rows.shape[0]
        pieces = self._c(pieces, np.uint8); npieces = self._c(npieces, np.uint8)
        pool = np.zeros((K, 4, 4), np.uint32)
        self.L.emul_pack(_np_ptr(pool), 0, 1, K, _np_ptr(rows), _np_ptr(pieces), pieces.shape[1], _np_ptr(npieces),
                         None, None, None, None)
        return pool

    def empty_states(self, n):
        stride = (n + 31) // 32 * 32
        return States(n, np.zeros((4, stride, 4), np.uint32), stride)

    def raw(self, s):
        """the 64-byte records as uint32[n, 16]"""
        return np.ascontiguousarray(np.transpose(s.planes[:, :s.n, :], (1, 0, 2))).reshape(s.n, 16)

    def unpack(self, s):
        n = s.n
        out = dict(rows=np.zeros((n, 20), np.uint16), cur=np.zeros(n, np.uint8), next=np.zeros(n, np.uint8),
                   lines=np.zeros(n, np.int32), moves=np.zeros(n, np.int32), state=np.zeros(n, np.int8),
                   head=np.zeros(n, np.uint8), npieces=np.zeros(n, np.uint8), queue=np.zeros((n, 42), np.uint8))
        self.L.emul_unpack(_np_ptr(s.planes), s.stride, n, *[_np_ptr(out[k]) for k in
                           ("rows", "cur", "next", "lines", "moves", "state", "head", "npieces", "queue")])
        return out

    def step(self, s, rot, loc, L, M):
        n = s.n
        rot = np.mod(np.asarray(rot, np.int64), 4).astype(np.uint8)
        loc = np.minimum(np.asarray(loc, np.int64), 255).astype(np.uint8)
        dl, fl, st = np.zeros(n, np.int8), np.zeros(n, np.uint8), np.zeros(n, np.int8)
        self.L.emul_step(_np_ptr(s.planes), s.stride, n, _np_ptr(rot), _np_ptr(loc), _np_ptr(dl), _np_ptr(fl), _np_ptr(st), L, M)
        return dl, fl, st

    def afterstates(self, s, L, M, f32=False):
        n = s.n
        feats, flags = np.zeros((40, n, 4), np.uint8), np.zeros((40, n), np.uint8)
        ff = np.zeros((40, n, 4), np.float32) if f32 else None
        self.L.emul_afterstates(_np_ptr(s.planes), s.stride, n, _np_ptr(feats), _np_ptr(flags), _np_ptr(ff), L, M)
        out = (feats.reshape(4, 10, n, 4).transpose(2, 0, 1, 3), flags.reshape(4, 10, n).transpose(2, 0, 1))
        return out + (ff.reshape(4, 10, n, 4).transpose(2, 0, 1, 3),) if f32 else out

    def afterstates_packed(self, s, L, M):
        feats = np.zeros((40, s.n, 4), np.uint8)
        self.L.emul_afterstates(_np_ptr(s.planes), s.stride, s.n, _np_ptr(feats), None, None, L, M)
        for defer in (0, 1):                            # the alias-skipping variant must give the same bytes
            uni = np.zeros((40, s.n, 4), np.uint8)
            self.L.emul_afterstates_uniform(_np_ptr(s.planes), s.stride, s.n, _np_ptr(uni), L, M, defer)
            assert np.array_equal(uni, feats), f"warp-uniform afterstate variant (defer={defer}) differs from the plain one"
        cur = np.full((40, s.n, 4), 0xEE, np.uint8)     # the compact-form path of the persistent kernels (cursor sink, deferred slots)
        self.L.emul_afterstates_cursor(_np_ptr(s.planes), s.stride, s.n, _np_ptr(cur), L, M)
        assert np.array_equal(cur, feats), "compact-form (cursor sink) afterstate variant differs from the plain one"
        spl = np.full((40, s.n, 4), 0xEE, np.uint8)     # and so must the one-thread-per-rotation variant
        self.L.emul_afterstates_split(_np_ptr(s.planes), s.stride, s.n, _np_ptr(spl), L, M)
        assert np.array_equal(spl, feats), "rotation-split afterstate variant differs from the plain one"
        return feats.reshape(4, 10, s.n, 4).transpose(2, 0, 1, 3)

    def reset(self, s, pool, idx=None, mask=None, mode=0, episode=None, seed=0, env_base=0, gen_count=0, tstep=None):
        idx = self._c(idx, np.int32); mask = self._c(mask, np.uint8)
        self.L.emul_reset_from_pool(_np_ptr(s.planes), s.stride, s.n, _np_ptr(pool), pool.shape[0], _np_ptr(idx), _np_ptr(mask),
                                    mode, _np_ptr(episode), _np_ptr(tstep), seed, env_base, gen_count)

    def step_observe(self, s, rot, loc, pool, episode, seed, env_base, L, M, packed=False, tstep=None, distinct=False, gen_count=0):
        """distinct=True: the afterstates come back as (rows uint32[used], runs uint32[n]) in place of (feats, afl)."""
        n = s.n
        rot = np.mod(np.asarray(rot, np.int64), 4).astype(np.uint8)
        loc = np.minimum(np.asarray(loc, np.int64), 255).astype(np.uint8)
        dl, fl, st = np.zeros(n, np.int8), np.zeros(n, np.uint8), np.zeros(n, np.int8)
        feats, afl = np.zeros((40, n, 4), np.uint8), (None if packed else np.zeros((40, n), np.uint8))
        drows, druns, dcur = (np.zeros(34 * n + 4, np.uint32), np.zeros(n, np.uint32), np.zeros(1, np.uint32)) if distinct else (None, None, None)
        stats = np.zeros(8, np.int64)
        self.L.emul_step_observe(_np_ptr(s.planes), s.stride, n, _np_ptr(rot), _np_ptr(loc), _np_ptr(dl), _np_ptr(fl), _np_ptr(st),
                                 _np_ptr(stats), _np_ptr(pool), pool.shape[0], _np_ptr(episode), _np_ptr(tstep), seed, env_base, gen_count,
                                 _np_ptr(feats), _np_ptr(afl), None, L, M, _np_ptr(drows), _np_ptr(druns), _np_ptr(dcur))
        if distinct:
            return dl, fl, st, drows[:int(dcur[0])], druns, stats
        return dl, fl, st, feats, afl, stats

    def afterstates_distinct(self, s, L, M):
        """(rows uint32[used], runs uint32[n]); the in-place and the deferred resolution of row-completing slots must agree"""
        n = s.n
        out = []
        for defer in (0, 1):
            rows, runs, cur = np.zeros(34 * n + 4, np.uint32), np.zeros(n, np.uint32), np.zeros(1, np.uint32)
            self.L.emul_afterstates_distinct(_np_ptr(s.planes), s.stride, n, _np_ptr(rows), _np_ptr(runs), 0, _np_ptr(cur), L, M, defer)
            out.append((rows[:int(cur[0])], runs))
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
        return out[0]

    def gen_pieces(self, n, count, seed, env_base, episode0=0, episode=None):
        out = np.zeros((n, count), np.uint8)
        self.L.emul_gen_pieces(_np_ptr(out), n, count, seed, env_base, _np_ptr(episode), episode0)
        return out

    def rollout(self, s, pool, episode, tstep, steps, seed, env_base, gen_count, L, M, weights=None):
        stats = np.zeros(8, np.int64)
        if weights is None:
            self.L.emul_rollout_random(_np_ptr(s.planes), s.stride, s.n, _np_ptr(pool), pool.shape[0], _np_ptr(episode),
                                       _np_ptr(tstep), _np_ptr(stats), steps, seed, env_base, gen_count, L, M)
        else:
            w = np.ascontiguousarray(weights, np.int32)
            self.L.emul_rollout_greedy(_np_ptr(s.planes), s.stride, s.n, _np_ptr(pool), pool.shape[0], _np_ptr(episode),
                                       _np_ptr(tstep), _np_ptr(stats), steps, _np_ptr(w), seed, env_base, gen_count, L, M)
        return stats


class GpuEngine:
    """Calls the product's C ABI directly (device-pointer entry points) with torch CUDA tensors."""
    name = "gpu"

    def __init__(self):
        import torch
        import tetris_piclim  # noqa: F401  (alias module: puts the package on sys.path)
        from importlib import import_module
        self.torch = torch
        self._lib = import_module(tetris_piclim.__name__ + "._lib")
        self.L = self._lib.lib()
        self.dev = torch.device("cuda", 0)

    def _t(self, x, dt):
        if x is None:
            return None
        a = np.ascontiguousarray(x, dt)
        if a.dtype == np.uint16:
            return self.torch.from_numpy(a.view(np.int16)).to(self.dev)
        if a.dtype == np.uint32:
            return self.torch.from_numpy(a.view(np.int32)).to(self.dev)
        return self.torch.from_numpy(a).to(self.dev)

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def _chk(self, code, what):
        self._lib.check(code, what)

    def pack(self, rows, pieces, npieces, lines=None, moves=None, state=None, head=None):
        t = self.torch
        rows = np.ascontiguousarray(rows, np.uint16); n = rows.shape[0]
        pieces = np.ascontiguousarray(pieces, np.uint8)
        stride = (n + 31) // 32 * 32
        planes = t.zeros((4, stride, 4), dtype=t.int32, device=self.dev)
        a = [self._t(rows, np.uint16), self._t(pieces, np.uint8), self._t(npieces, np.uint8), self._t(lines, np.int32),
             self._t(moves, np.int32), self._t(state, np.int8), self._t(head, np.uint8)]
        self._chk(self.L.tpl_pack(self._p(planes), stride, 0, n, self._p(a[0]), self._p(a[1]), pieces.shape[1], self._p(a[2]),
                                  self._p(a[3]), self._p(a[4]), self._p(a[5]), self._p(a[6]), self._stream()), "tpl_pack")
        t.cuda.synchronize()
        return States(n, planes, stride)

    def make_pool(self, rows, pieces, npieces):
        t = self.torch
        rows = np.ascontiguousarray(rows, np.uint16); K = rows.shape[0]
        pieces = np.ascontiguousarray(pieces, np.uint8)
        pool = t.zeros((K, 4, 4), dtype=t.int32, device=self.dev)
        a = [self._t(rows, np.uint16), self._t(pieces, np.uint8), self._t(npieces, np.uint8)]
        self._chk(self.L.tpl_pack(self._p(pool), 0, 1, K, self._p(a[0]), self._p(a[1]), pieces.shape[1], self._p(a[2]),
                                  None, None, None, None, self._stream()), "tpl_pack")
        t.cuda.synchronize()
        return pool

    def empty_states(self, n):
        t = self.torch
        stride = (n + 31) // 32 * 32
        return States(n, t.zeros((4, stride, 4), dtype=t.int32, device=self.dev), stride)

    def raw(self, s):
        return s.planes[:, :s.n, :].permute(1, 0, 2).contiguous().cpu().numpy().view(np.uint32).reshape(s.n, 16)

    def unpack(self, s):
        t = self.torch; n = s.n
        mk = lambda shape, dt: t.zeros(shape, dtype=dt, device=self.dev)   # noqa: E731
        o = dict(rows=mk((n, 20), t.int16), cur=mk(n, t.uint8), next=mk(n, t.uint8), lines=mk(n, t.int32), moves=mk(n, t.int32),
                 state=mk(n, t.int8), head=mk(n, t.uint8), npieces=mk(n, t.uint8), queue=mk((n, 42), t.uint8))
        self._chk(self.L.tpl_unpack(self._p(s.planes), s.stride, n, *[self._p(o[k]) for k in
                  ("rows", "cur", "next", "lines", "moves", "state", "head", "npieces", "queue")], self._stream()), "tpl_unpack")
        out = {k: v.cpu().numpy() for k, v in o.items()}
        out["rows"] = out["rows"].view(np.uint16)
        return out

    def step(self, s, rot, loc, L, M):
        t = self.torch; n = s.n
        rot = self._t(np.mod(np.asarray(rot, np.int64), 4).astype(np.uint8), np.uint8)
        loc = self._t(np.minimum(np.asarray(loc, np.int64), 255).astype(np.uint8), np.uint8)
        dl, fl, st = t.zeros(n, dtype=t.int8, device=self.dev), t.zeros(n, dtype=t.uint8, device=self.dev), t.zeros(n, dtype=t.int8, device=self.dev)
        self._chk(self.L.tpl_step(self._p(s.planes), s.stride, n, self._p(rot), self._p(loc), self._p(dl), self._p(fl), self._p(st),
                                  None, L, M, self._stream()), "tpl_step")
        return dl.cpu().numpy(), fl.cpu().numpy(), st.cpu().numpy()

    def afterstates(self, s, L, M, f32=False):
        t = self.torch; n = s.n
        feats = t.zeros((40, n, 4), dtype=t.uint8, device=self.dev)
        flags = t.zeros((40, n), dtype=t.uint8, device=self.dev)
        ff = t.zeros((40, n, 4), dtype=t.float32, device=self.dev) if f32 else None
        self._chk(self.L.tpl_afterstates(self._p(s.planes), s.stride, n, self._p(feats), self._p(flags), self._p(ff), L, M,
                                         self._stream()), "tpl_afterstates")
        out = (feats.cpu().numpy().reshape(4, 10, n, 4).transpose(2, 0, 1, 3), flags.cpu().numpy().reshape(4, 10, n).transpose(2, 0, 1))
        return out + (ff.cpu().numpy().reshape(4, 10, n, 4).transpose(2, 0, 1, 3),) if f32 else out

    def afterstates_packed(self, s, L, M):
        t = self.torch
        feats = t.zeros((40, s.n, 4), dtype=t.uint8, device=self.dev)
        self._chk(self.L.tpl_afterstates(self._p(s.planes), s.stride, s.n, self._p(feats), None, None, L, M, self._stream()),
                  "tpl_afterstates(packed)")
        return feats.cpu().numpy().reshape(4, 10, s.n, 4).transpose(2, 0, 1, 3)

    def reset(self, s, pool, idx=None, mask=None, mode=0, episode=None, seed=0, env_base=0, gen_count=0, tstep=None):
        d_idx, d_mask = self._t(idx, np.int32), self._t(mask, np.uint8)
        d_ep = self._t(episode, np.uint32) if episode is not None else None
        d_ts = self._t(tstep, np.uint32) if tstep is not None else None
        self._chk(self.L.tpl_reset_from_pool(self._p(s.planes), s.stride, s.n, self._p(pool), pool.shape[0], self._p(d_idx),
                                             self._p(d_mask), mode, self._p(d_ep), self._p(d_ts), seed, env_base, gen_count, self._stream()),
                  "tpl_reset_from_pool")
        if episode is not None:
            episode[:] = d_ep.cpu().numpy().view(np.uint32)
        if tstep is not None:
            tstep[:] = d_ts.cpu().numpy().view(np.uint32)

    def _distinct_args(self, n):
        t = self.torch
        cap = 34 * n + 4 * ((n + 31) // 32)
        rows = t.full((cap,), -1, dtype=t.int32, device=self.dev)
        runs = t.zeros(n, dtype=t.int32, device=self.dev)
        if not hasattr(self, "_cursor2"):
            self._cursor2, self._phase = t.zeros(2, dtype=t.int32, device=self.dev), 0
        phase = self._phase; self._phase ^= 1
        return rows, cap, runs, self._cursor2, phase

    def _distinct_result(self, rows, runs, phase):
        used = int(self._cursor2[phase].item())
        assert int(self._cursor2[phase ^ 1].item()) == 0, "the kernel must clear the other phase's counter"
        return rows.cpu().numpy().view(np.uint32)[:used], runs.cpu().numpy().view(np.uint32)

    def afterstates_distinct(self, s, L, M):
        rows, cap, runs, cur, phase = self._distinct_args(s.n)
        self._chk(self.L.tpl_afterstates_distinct(self._p(s.planes), s.stride, s.n, self._p(rows), cap, self._p(runs), 0, self._p(cur), phase,
                                                  L, M, self._stream()), "tpl_afterstates_distinct")
        r, d = self._distinct_result(rows, runs, phase)
        # the device-side expansion must agree with the host helper
        if s.n:
            t = self.torch
            out = t.zeros((40, s.n, 4), dtype=t.uint8, device=self.dev)
            self._chk(self.L.tpl_expand_distinct(self._p(rows), self._p(runs), s.n, self._p(out), self._stream()), "tpl_expand_distinct")
            from importlib import import_module
            import tetris_piclim
            dm = import_module(tetris_piclim.__name__ + ".distinct")
            assert np.array_equal(out.cpu().numpy().transpose(1, 0, 2), dm.expand(r, d)), "tpl_expand_distinct != distinct.expand"
        return r, d

    def step_observe(self, s, rot, loc, pool, episode, seed, env_base, L, M, packed=False, tstep=None, distinct=False, gen_count=0):
        t = self.torch; n = s.n
        rot = self._t(np.mod(np.asarray(rot, np.int64), 4).astype(np.uint8), np.uint8)
        loc = self._t(np.minimum(np.asarray(loc, np.int64), 255).astype(np.uint8), np.uint8)
        z = lambda shape, dt: t.zeros(shape, dtype=dt, device=self.dev)     # noqa: E731
        dl, fl, st = z(n, t.int8), z(n, t.uint8), z(n, t.int8)
        stats = z(8, t.int64)
        d_ep = self._t(episode, np.uint32)
        d_ts = self._t(tstep, np.uint32) if tstep is not None else None
        if distinct:
            rows, cap, runs, cur, phase = self._distinct_args(n)
            self._chk(self.L.tpl_step_observe_distinct(self._p(s.planes), s.stride, n, self._p(rot), self._p(loc), self._p(dl), self._p(fl),
                                                       self._p(st), self._p(stats), self._p(pool), pool.shape[0], self._p(d_ep), self._p(d_ts),
                                                       seed, env_base, gen_count, self._p(rows), cap, self._p(runs), 0, self._p(cur), phase, L, M,
                                                       self._stream()), "tpl_step_observe_distinct")
            a, b = self._distinct_result(rows, runs, phase)
        else:
            feats, afl = z((40, n, 4), t.uint8), (None if packed else z((40, n), t.uint8))
            self._chk(self.L.tpl_step_observe(self._p(s.planes), s.stride, n, self._p(rot), self._p(loc), self._p(dl), self._p(fl),
                                              self._p(st), self._p(stats), self._p(pool), pool.shape[0], self._p(d_ep), self._p(d_ts), seed,
                                              env_base, gen_count, self._p(feats), self._p(afl), None, L, M, self._stream()), "tpl_step_observe")
            a, b = feats.cpu().numpy(), (None if afl is None else afl.cpu().numpy())
        episode[:] = d_ep.cpu().numpy().view(np.uint32)
        if tstep is not None:
            tstep[:] = d_ts.cpu().numpy().view(np.uint32)
        return dl.cpu().numpy(), fl.cpu().numpy(), st.cpu().numpy(), a, b, stats.cpu().numpy()

    def gen_pieces(self, n, count, seed, env_base, episode0=0, episode=None):
        t = self.torch
        out = t.zeros((n, count), dtype=t.uint8, device=self.dev)
        d_ep = self._t(episode, np.uint32) if episode is not None else None
        self._chk(self.L.tpl_gen_pieces(self._p(out), n, count, seed, env_base, self._p(d_ep), episode0, self._stream()), "tpl_gen_pieces")
        return out.cpu().numpy()

    def rollout(self, s, pool, episode, tstep, steps, seed, env_base, gen_count, L, M, weights=None):
        t = self.torch
        d_ep, d_ts = self._t(episode, np.uint32), self._t(tstep, np.uint32)
        stats = t.zeros(8, dtype=t.int64, device=self.dev)
        if weights is None:
            self._chk(self.L.tpl_rollout_random(self._p(s.planes), s.stride, s.n, self._p(pool), pool.shape[0], self._p(d_ep),
                                                self._p(d_ts), self._p(stats), steps, seed, env_base, gen_count, L, M, self._stream()),
                      "tpl_rollout_random")
        else:
            w = np.ascontiguousarray(weights, np.int32)
            self._chk(self.L.tpl_rollout_greedy(self._p(s.planes), s.stride, s.n, self._p(pool), pool.shape[0], self._p(d_ep),
                                                self._p(d_ts), self._p(stats), steps, ctypes.c_void_p(w.ctypes.data), seed, env_base,
                                                gen_count, L, M, self._stream()), "tpl_rollout_greedy")
        episode[:] = d_ep.cpu().numpy().view(np.uint32)
        tstep[:] = d_ts.cpu().numpy().view(np.uint32)
        return stats.cpu().numpy()
