"""BatchedOvercookedEnv — N independent Overcooked environments advanced by one CUDA launch.

The tensor-level API of the engine.  State lives in ONE int32 tensor
``state[N, S]`` (record layout in include/ovc_b200.h); ``step`` / ``rollout`` / ``reset`` /
``lossless_state_encoding`` / ``featurize_state`` are thin calls into the C ABI on torch's current
CUDA stream, so they compose with CUDA graphs and user streams.  torch owns every buffer; the
native library allocates nothing.

Semantics follow the reference per environment:
  step     OvercookedEnv.step      (src/overcooked_ai_py/mdp/overcooked_env.py:244-274) on top of
           OvercookedGridworld.get_state_transition (overcooked_mdp.py:1375-1430)
  reset    OvercookedEnv.reset     (overcooked_env.py:288-319), standard start state
  done     OvercookedEnv.is_done   (overcooked_env.py:321-325): timestep >= horizon
Stepping an environment that is already done leaves it untouched and sets EVF_STEPPED_DONE in its
event words (the reference raises AssertionError, overcooked_env.py:255); with ``auto_reset=True``
an environment that reaches the horizon is put back to its start state in the same launch (its
``done`` output is still 1 for that transition), which is what rollout collection wants.
"""
import ctypes

import numpy as np
import torch

from overcooked_ai_b200 import _native
from overcooked_ai_b200 import layout as L

_TORCH_DT = {torch.float32: _native.DT_F32, torch.uint8: _native.DT_U8, torch.int32: _native.DT_I32,
             torch.bfloat16: _native.DT_BF16}


def _as_layouts(layouts, mdp_params):
    if isinstance(layouts, (str, L.CompiledLayout)) or hasattr(layouts, "compiled"):
        layouts = [layouts]
    out = []
    for l in layouts:
        if isinstance(l, str):
            l = L.compile_layout(l, **(mdp_params or {}))
        elif hasattr(l, "compiled"):  # an overcooked_ai_b200.mdp.OvercookedGridworld
            l = l.compiled
        out.append(l)
    return out


class BatchedOvercookedEnv(object):
    def __init__(self, layouts, n_envs, horizon=400, device="cuda", auto_reset=False, state_words=None,
                 io=_native.IO_DEFAULT, env_layout=None, mdp_params=None, pdl=True,
                 random_start_pos=False, rnd_obj_prob_thresh=0.0, seed=0, random_layout=False):
        """
        layouts      layout name / CompiledLayout / OvercookedGridworld, or a list of them (mixed batch)
        n_envs       number of environments on THIS device
        horizon      episode length (OvercookedEnv's ``horizon``); <= 0 means no horizon
        env_layout   optional int array [n_envs] of layout indices; default: contiguous, near-equal
                     segments, one per layout (a warp then sees one layout)
        io           record I/O strategy of the step kernel (_native.IO_*); 0 = library default
        pdl          launch the step kernel with programmatic dependent launch: back-to-back transitions overlap
                     the next launch's prologue with the current kernel (3.58 -> 3.20 us per launch at 65 536 envs)
        random_start_pos, rnd_obj_prob_thresh, seed
                     start every episode from the reference's randomised start states
                     (get_random_start_state_fn, overcooked_mdp.py:1307-1369) instead of the standard one;
                     drawn on the device with a counter-based generator (see ovc_random_start_t)
        random_layout
                     variable MDP — OvercookedEnv(mdp_generator_fn, num_mdp > 1) whose reset draws a new MDP
                     (overcooked_env.py:288-302): every (auto-)reset redraws the environment's layout uniformly
                     from ``layouts`` (e.g. a pool from layout_generator.generate_layout_pool); ``layout_ids()``
                     gives the current assignment, ``env_layout`` only the initial one
        """
        self._lib = _native.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedOvercookedEnv needs a CUDA device: this engine has no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.layouts = _as_layouts(layouts, mdp_params)
        self.n_layouts = len(self.layouts)
        self.n_envs = int(n_envs)
        self.horizon = int(horizon) if horizon < 2**31 else 0
        self.auto_reset = bool(auto_reset)
        self.io = int(io)
        self.pdl = bool(pdl)  # programmatic dependent launch for back-to-back step() calls / graphs
        tab, starts, S = L.build_tables(self.layouts, state_words)
        assert tab.shape[1] == self._lib.ovc_layout_table_size(), "layout table size mismatch with the native library"
        self.state_words = S
        self._tab_host, self._starts_host = tab, starts
        with torch.cuda.device(self.device):
            self.tables = torch.from_numpy(tab).to(self.device)
            self.start_records = torch.from_numpy(starts).to(self.device)
            if env_layout is None:
                bounds = [self.n_envs * i // self.n_layouts for i in range(self.n_layouts + 1)]
                env_layout = np.zeros(self.n_envs, np.int32)
                for i in range(self.n_layouts):
                    env_layout[bounds[i]:bounds[i + 1]] = i
            env_layout = np.ascontiguousarray(env_layout, dtype=np.int32)
            assert env_layout.shape == (self.n_envs,) and (env_layout >= 0).all() and (env_layout < self.n_layouts).all()
            self.env_layout_host = env_layout
            self.env_layout = torch.from_numpy(env_layout).to(self.device)
            self.state = torch.zeros((self.n_envs, S), dtype=torch.int32, device=self.device)
            self.sparse = torch.zeros(self.n_envs, dtype=torch.int32, device=self.device)
            self.shaped = torch.zeros((self.n_envs, 2), dtype=torch.int32, device=self.device)
            self.done = torch.zeros(self.n_envs, dtype=torch.int32, device=self.device)
            self.events = torch.zeros((self.n_envs, 2), dtype=torch.int32, device=self.device)
        self._rs = None
        self.random_layout = bool(random_layout)
        if random_start_pos or rnd_obj_prob_thresh > 0 or random_layout:
            thr = min(int(float(rnd_obj_prob_thresh) * 4294967296.0), 0xFFFFFFFF)  # 0xFFFFFFFF = always (ovc_rng.cuh)
            self._rs = _native.RandomStart(int(seed) & 0xFFFFFFFFFFFFFFFF, thr, int(bool(random_start_pos)),
                                           int(self.random_layout), 0)
        self._lut = None
        self._segments = None
        with torch.cuda.device(self.device):
            self.reset()

    # ---------------------------------------------------------------------------------------------
    def _flags(self):
        return ((_native.F_AUTO_RESET if self.auto_reset else 0) | (_native.F_PDL if self.pdl else 0)
                | (self.io << _native.F_IO_SHIFT))

    def _rs_ptr(self):
        return ctypes.byref(self._rs) if self._rs is not None else None

    def _stream(self):
        # the C ABI launches on the calling thread's current device: make a mismatch loud instead of a fault
        if torch.cuda.current_device() != self.device.index:
            raise RuntimeError("this environment lives on %s but the current CUDA device is cuda:%d; wrap the call in "
                               "torch.cuda.device(env.device) (one process per GPU sets it once)" % (self.device, torch.cuda.current_device()))
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, mask=None):
        """Put all environments (or those with mask != 0; int32 tensor [N]) back to their layout's
        standard start state (overcooked_mdp.py:1297-1305)."""
        if mask is not None:
            assert mask.dtype == torch.int32 and mask.is_cuda and mask.is_contiguous() and mask.numel() == self.n_envs
        _native.check(self._lib.ovc_reset(
            self.tables.data_ptr(), self.n_layouts, self.start_records.data_ptr(), self.state.data_ptr(),
            self.env_layout.data_ptr(), 0 if mask is None else mask.data_ptr(), self.n_envs, self.state_words,
            self._rs_ptr(), self._stream()))

    def step(self, actions, out=None):
        """One joint transition of every environment.

        actions  int32 CUDA tensor [N, 2], action indices 0..5 (Action.INDEX_TO_ACTION order)
        out      optional (sparse[N], shaped[N,2], done[N], events[N,2]) int32 CUDA tensors to write
        returns  (sparse, shaped, done, events); without ``out`` these are buffers owned by the env and
                 overwritten by the next call.  ``sparse`` is the summed delivery reward
                 (what OvercookedEnv.step returns), ``shaped`` is shaped_reward_by_agent,
                 ``events`` holds one bit per EVENT_TYPES entry per agent (layout.EVENT_TYPES).
        """
        assert actions.dtype == torch.int32 and actions.is_cuda and actions.is_contiguous()
        assert actions.numel() == 2 * self.n_envs
        sparse, shaped, done, events = (self.sparse, self.shaped, self.done, self.events) if out is None else out
        if out is not None:
            for o, n in zip(out, (1, 2, 1, 2)):
                assert o.dtype == torch.int32 and o.is_cuda and o.is_contiguous() and o.numel() == n * self.n_envs, \
                    "step(out=...) takes int32 CUDA tensors (sparse[N], shaped[N,2], done[N], events[N,2])"
        _native.check(self._lib.ovc_step(
            self.tables.data_ptr(), self.n_layouts, self.start_records.data_ptr(), self.state.data_ptr(),
            actions.data_ptr(), sparse.data_ptr(), shaped.data_ptr(), done.data_ptr(),
            events.data_ptr(), self.n_envs, self.state_words, self.horizon, self._flags(), self._rs_ptr(), self._stream()))
        return sparse, shaped, done, events

    def narrow_ok(self):
        """True if every layout's rewards fit the narrow transfer formats (int16 sparse, int8 shaped)."""
        return all(int(l.deliver_value.max()) * 2 <= 32767 and max(
            int(l.reward_shaping_params[k]) for k in ("PLACEMENT_IN_POT_REW", "DISH_PICKUP_REWARD", "SOUP_PICKUP_REWARD")) <= 127
            for l in self.layouts)

    def n_groups(self):
        """Groups of 32 consecutive environments (one warp each) — the unit of the sparse event stream."""
        return (self.n_envs + 31) // 32

    def alloc_stream_out(self, T, cap, n_chunks=1, pin=False, dense_backup=False):
        """Buffers of the sparse event stream (OVC_F_OUT_STREAM): (masks uint32 as int32 [T, G], values int16 [n_chunks, G, cap],
        dense code words int16 [T, N] or None)."""
        G = self.n_groups()
        mk = (lambda sh, dt: torch.zeros(sh, dtype=dt, pin_memory=True)) if pin else (lambda sh, dt: torch.zeros(sh, dtype=dt, device=self.device))
        return (mk((T, G), torch.int32), mk((n_chunks, G, cap), torch.int16), mk((T, self.n_envs), torch.int16) if dense_backup else None)

    def rollout_stream(self, actions, cap, out=None, dense_backup=False):
        """T transitions in one launch with the result as a sparse event stream (include/ovc_b200.h OVC_F_OUT_STREAM):
        per transition and group of 32 environments one lane mask of the non-zero code words, plus each group's
        non-zero words compacted in (transition, lane) order, at most ``cap`` per group (the masks count the rest).
        actions: int32 / uint8 [T, N, 2] or one-byte joint actions uint8 [T, N].  Returns (masks, values, dense or None);
        expand with ``expand_stream``."""
        assert actions.dtype in (torch.int32, torch.uint8) and actions.is_cuda and actions.is_contiguous() and actions.dim() in (2, 3)
        T = actions.shape[0]
        assert actions.shape[1] == self.n_envs and 1 <= cap <= _native.STREAM_CAP_MAX
        if out is None:
            out = self.alloc_stream_out(T, cap, dense_backup=dense_backup)
        masks, values, dense = out
        assert masks.dtype == torch.int32 and tuple(masks.shape) == (T, self.n_groups()) and masks.is_cuda and masks.is_contiguous()
        assert values.dtype == torch.int16 and values.numel() == self.n_groups() * cap and values.is_cuda and values.is_contiguous()
        flags = self._flags() | _native.F_OUT_STREAM | (int(cap) << _native.F_STREAM_CAP_SHIFT)
        if actions.dim() == 2:
            assert actions.dtype == torch.uint8
            flags |= _native.F_ACT_PACKED
        elif actions.dtype == torch.uint8:
            flags |= _native.F_ACT_U8
        if flags >= 2**31:  # the C int carries the capacity in its upper half
            flags -= 2**32
        _native.check(self._lib.ovc_rollout(
            self.tables.data_ptr(), self.n_layouts, self.start_records.data_ptr(), self.state.data_ptr(),
            actions.data_ptr(), values.data_ptr(), 0, 0 if dense is None else dense.data_ptr(), masks.data_ptr(),
            self.n_envs, T, self.state_words, self.horizon, flags, self._rs_ptr(), self._stream()))
        return out

    def expand_stream(self, masks, values, chunk=None, sparse=True, shaped=True, done=True, events=False, n_threads=0, out=None):
        """Dense host arrays from a HOST sparse event stream (ovc_expand_stream_host): masks int32 [T, G], values int16
        [n_chunks, G, cap] with ``chunk`` transitions per launch (default: all T in one).  Returns (dict of arrays as
        expand_codes, number of (chunk, group) slices that overflowed their capacity)."""
        assert masks.dtype == torch.int32 and not masks.is_cuda and masks.is_contiguous() and masks.dim() == 2
        assert values.dtype == torch.int16 and not values.is_cuda and values.is_contiguous() and values.dim() == 3
        T, G = masks.shape
        assert G == self.n_groups() and values.shape[1] == G
        chunk = T if chunk is None else int(chunk)
        assert values.shape[0] == -(-T // chunk)
        N, cap = self.n_envs, values.shape[2]
        tbl = self.code_reward_table()
        lay = self.env_layout_host
        if self.random_layout:
            assert (tbl == tbl[:1]).all(), "random_layout with different reward tables: the codes alone do not name the layout"
            lay = None
        if out is None:
            out = {}
            if sparse:
                out["sparse"] = torch.empty((T, N), dtype=torch.int16)
            if shaped:
                out["shaped"] = torch.empty((T, N, 2), dtype=torch.int8)
            if done:
                out["done"] = torch.empty((T, N), dtype=torch.uint8)
            if events:
                out["events"] = torch.empty((T, N, 2), dtype=torch.int32)
        ptr = lambda k: out[k].data_ptr() if k in out else 0
        tbl = np.ascontiguousarray(tbl, dtype=np.int32)
        over = ctypes.c_int64(0)
        _native.check(self._lib.ovc_expand_stream_host(
            masks.data_ptr(), values.data_ptr(), T, chunk, cap, N, 0 if lay is None else lay.ctypes.data, tbl.ctypes.data,
            self.n_layouts, ptr("sparse"), ptr("shaped"), ptr("done"), ptr("events"), int(n_threads), ctypes.byref(over)))
        return out, int(over.value)

    def alloc_rollout_out(self, T, narrow=False, pin=False, packed=False, codes=False):
        """Output tensors for rollout(): (sparse[T,N], shaped[T,N,2], done[T,N], events[T,N,2]); int32, or with
        ``narrow`` int16 / int8 / uint8 / int32 (13 bytes per env-step), or with ``packed`` (6 bytes per env-step)
        (sparse int16 [T,N], shaped int8 [T,N,2], None, evcode int16 [T,N]) where evcode carries both agents' 5-bit
        event codes + done (include/ovc_b200.h OVC_F_OUT_PACKED; expand with wire.decode_event_codes), or with
        ``codes`` (2 bytes per env-step) (None, None, None, evcode int16 [T,N]): the same word plus one
        "shaped reward granted" bit per agent, from which rewards follow by table (OVC_F_OUT_CODES; expand_codes)."""
        N = self.n_envs
        mk = (lambda sh, dt: torch.empty(sh, dtype=dt, pin_memory=True)) if pin else (lambda sh, dt: torch.empty(sh, dtype=dt, device=self.device))
        if codes:
            return (None, None, None, mk((T, N), torch.int16))
        if packed:
            return (mk((T, N), torch.int16), mk((T, N, 2), torch.int8), None, mk((T, N), torch.int16))
        dts = (torch.int16, torch.int8, torch.uint8, torch.int32) if narrow else (torch.int32,) * 4
        shapes = ((T, N), (T, N, 2), (T, N), (T, N, 2))
        if pin:
            return tuple(torch.empty(sh, dtype=dt, pin_memory=True) for sh, dt in zip(shapes, dts))
        return tuple(torch.empty(sh, dtype=dt, device=self.device) for sh, dt in zip(shapes, dts))

    def rollout(self, actions, out=None):
        """T transitions in one launch (state stays on chip between them).

        actions  int32 or uint8 CUDA tensor [T, N, 2], or uint8 [T, N] with both agents' indices in one byte
                 (agent 0 in bits 0-3, agent 1 in bits 4-7; wire.pack_actions)
        out      optional (sparse[T,N], shaped[T,N,2], done[T,N], events[T,N,2]): all int32, or one of the narrower
                 sets of alloc_rollout_out(narrow= / packed= / codes=).
        Equivalent to T calls of step() with the same actions.
        """
        assert actions.dtype in (torch.int32, torch.uint8) and actions.is_cuda and actions.is_contiguous() and actions.dim() in (2, 3)
        T = actions.shape[0]
        assert actions.shape[1] == self.n_envs
        if out is None:
            out = self.alloc_rollout_out(T)
        sparse, shaped, done, events = out
        flags = self._flags()
        if actions.dim() == 2:
            assert actions.dtype == torch.uint8, "one-byte joint actions are uint8 [T, N]"
            flags |= _native.F_ACT_PACKED
        else:
            assert actions.shape[2] == 2
            if actions.dtype == torch.uint8:
                flags |= _native.F_ACT_U8
        if sparse is None:  # codes: 2 bytes per env-step
            assert shaped is None and done is None and events.dtype == torch.int16 and events.dim() == 2
            assert tuple(events.shape) == (T, self.n_envs) and events.is_cuda and events.is_contiguous()
            flags |= _native.F_OUT_CODES
            _native.check(self._lib.ovc_rollout(
                self.tables.data_ptr(), self.n_layouts, self.start_records.data_ptr(), self.state.data_ptr(),
                actions.data_ptr(), 0, 0, 0, events.data_ptr(), self.n_envs, T, self.state_words, self.horizon, flags,
                self._rs_ptr(), self._stream()))
            return out
        if done is None:  # packed: 6 bytes per env-step
            assert sparse.dtype == torch.int16 and shaped.dtype == torch.int8 and events.dtype == torch.int16 and events.dim() == 2
            assert self.narrow_ok(), "rewards of these layouts do not fit the narrow formats"
            flags |= _native.F_OUT_PACKED
        elif sparse.dtype == torch.int16:
            assert shaped.dtype == torch.int8 and done.dtype == torch.uint8 and events.dtype == torch.int32
            assert self.narrow_ok(), "rewards of these layouts do not fit the narrow formats"
            flags |= _native.F_OUT_NARROW
        else:
            assert sparse.dtype == shaped.dtype == done.dtype == events.dtype == torch.int32
        _native.check(self._lib.ovc_rollout(
            self.tables.data_ptr(), self.n_layouts, self.start_records.data_ptr(), self.state.data_ptr(),
            actions.data_ptr(), sparse.data_ptr(), shaped.data_ptr(), 0 if done is None else done.data_ptr(), events.data_ptr(),
            self.n_envs, T, self.state_words, self.horizon, flags, self._rs_ptr(), self._stream()))
        return out

    def code_reward_table(self):
        """int32 numpy [n_layouts, 2, 32]: delivery reward and shaped reward of each event code (OVC_F_OUT_CODES)."""
        from overcooked_ai_b200 import wire

        return wire.code_reward_table(self.layouts)

    def expand_codes(self, evcode, sparse=True, shaped=True, done=True, events=False, n_threads=0, out=None):
        """Dense host arrays from a HOST int16 [T, N] tensor of OVC_F_OUT_CODES words, on the host cores
        (ovc_expand_codes_host): dict with the requested int16 sparse [T,N], int8 shaped [T,N,2], uint8 done [T,N],
        int32 events [T,N,2].  With ``random_layout`` every layout of the pool must share one reward table."""
        assert evcode.dtype == torch.int16 and not evcode.is_cuda and evcode.is_contiguous() and evcode.dim() == 2
        T, N = evcode.shape
        assert N == self.n_envs
        tbl = self.code_reward_table()
        lay = self.env_layout_host
        if self.random_layout:
            assert (tbl == tbl[:1]).all(), "random_layout with different reward tables: the codes alone do not name the layout"
            lay = None
        if out is None:
            out = {}
            if sparse:
                out["sparse"] = torch.empty((T, N), dtype=torch.int16)
            if shaped:
                out["shaped"] = torch.empty((T, N, 2), dtype=torch.int8)
            if done:
                out["done"] = torch.empty((T, N), dtype=torch.uint8)
            if events:
                out["events"] = torch.empty((T, N, 2), dtype=torch.int32)
        ptr = lambda k: out[k].data_ptr() if k in out else 0
        tbl = np.ascontiguousarray(tbl, dtype=np.int32)
        _native.check(self._lib.ovc_expand_codes_host(
            evcode.data_ptr(), T, N, 0 if lay is None else lay.ctypes.data, tbl.ctypes.data, self.n_layouts,
            ptr("sparse"), ptr("shaped"), ptr("done"), ptr("events"), int(n_threads)))
        return out

    # ---------------------------------------------------------------------------------------------
    def segments(self):
        """[(begin, end, layout index)] maximal runs of equal layout in env order."""
        if self._segments is None:
            el = self.env_layout_host
            cuts = [0] + (np.nonzero(np.diff(el))[0] + 1).tolist() + [self.n_envs]
            self._segments = [(cuts[i], cuts[i + 1], int(el[cuts[i]])) for i in range(len(cuts) - 1) if cuts[i] < cuts[i + 1]]
        return self._segments

    def layout_ids(self):
        """int32 [N]: the layout each environment is on NOW (the id lives in word 3 of its record)."""
        return (self.state[:, 3] & 0xFF).to(torch.int32)

    def _layout_ids_host(self):
        return self.layout_ids().cpu().numpy() if self.random_layout else self.env_layout_host

    def obs_shape(self, layout_index=0):
        l = self.layouts[layout_index]
        return (l.width, l.height, 26)

    def lossless_state_encoding(self, out=None, dtype=torch.float32, view_swap=None):
        """lossless_state_encoding (overcooked_mdp.py:2385-2561) of every environment, both players:
        tensor [N, 2, W, H, 26] (index order [x][y][channel], as the reference) when all layouts share
        one grid shape, else a list of such tensors, one per layout segment.  dtype float32 (what the
        reference's RLlib consumer casts to), bfloat16, uint8 or int32.  ``view_swap`` (int32 CUDA tensor [N]):
        where non-zero, ``out[env, 0]`` is player 1's view (primary-agent-first order of the gym wrapper)."""
        if view_swap is not None:
            assert view_swap.dtype == torch.int32 and view_swap.is_cuda and view_swap.is_contiguous() and view_swap.numel() == self.n_envs
        shapes = {(l.width, l.height) for l in self.layouts}
        if len(shapes) == 1:
            runs = [(0, self.n_envs, 0)]
        else:
            assert not self.random_layout, "random_layout needs layouts of one grid shape (pad them, LayoutGenerator does)"
            runs = self.segments()
        outs = []
        for k, (b, e, li) in enumerate(runs):
            W, H = self.layouts[li].width, self.layouts[li].height
            o = out[k] if isinstance(out, (list, tuple)) else out
            if o is None:
                o = torch.empty((e - b, 2, W, H, 26), dtype=dtype, device=self.device)
            assert o.is_cuda and o.is_contiguous() and o.numel() == (e - b) * 2 * W * H * 26
            assert o.dtype in _TORCH_DT, "lossless_state_encoding writes float32, bfloat16, uint8 or int32, not %s" % o.dtype
            _native.check(self._lib.ovc_encode_lossless(
                self.tables.data_ptr(), self.n_layouts, self.state.data_ptr() + 4 * self.state_words * b,
                0 if view_swap is None else view_swap.data_ptr() + 4 * b, o.data_ptr(),
                _TORCH_DT[o.dtype], e - b, self.state_words, W, H, self.horizon if self.horizon > 0 else 2**31 - 1,
                self._stream()))
            outs.append(o)
        return outs[0] if len(shapes) == 1 else outs

    def encoded_linear(self, wt, bias, out=None, neg_slope=0.2, view_swap=None):
        """First policy layer on ``lossless_state_encoding`` without the observation tensor (kernel K7, ovc_encode_linear):
        ``out[2 env + view] = leaky_relu(obs[env, view].flatten() @ wt + bias, neg_slope)`` as bfloat16 ``[2N, n_out]``.
        ``wt``: bfloat16 CUDA tensor ``[W*H*26, n_out]`` (the layer's matrix TRANSPOSED, rows in the observation's own
        element order ``[x][y][plane]`` — for a convolution, the matrix ``selfplay.DenseGridPolicy`` builds), ``bias``
        float32 ``[n_out]``, ``n_out`` a multiple of 64.  All environments must share one grid shape."""
        assert len({(l.width, l.height) for l in self.layouts}) == 1, "one grid shape per call (group envs by layout)"
        W, H = self.layouts[0].width, self.layouts[0].height
        assert wt.is_cuda and wt.dtype == torch.bfloat16 and wt.is_contiguous() and wt.shape[0] == W * H * 26, wt.shape
        n_out = wt.shape[1]
        assert bias.is_cuda and bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == n_out
        if view_swap is not None:
            assert view_swap.dtype == torch.int32 and view_swap.is_cuda and view_swap.is_contiguous() and view_swap.numel() == self.n_envs
        if out is None:
            out = torch.empty((2 * self.n_envs, n_out), dtype=torch.bfloat16, device=self.device)
        assert out.is_cuda and out.dtype == torch.bfloat16 and out.is_contiguous() and out.numel() == 2 * self.n_envs * n_out
        _native.check(self._lib.ovc_encode_linear(
            self.tables.data_ptr(), self.n_layouts, self.state.data_ptr(), 0 if view_swap is None else view_swap.data_ptr(),
            wt.data_ptr(), bias.data_ptr(), out.data_ptr(), self.n_envs, self.state_words, W, H,
            self.horizon if self.horizon > 0 else 2**31 - 1, n_out, float(neg_slope), self._stream()))
        return out

    def sample_actions(self, scores, counter, seed=0, out=None):
        """Joint actions drawn from the policy's logits (ovc_sample_actions: Gumbel-max on Philox draws, one kernel).
        ``scores`` float32 ``[2N, ld]`` (rows ordered [env][agent], the first 6 columns are the logits), ``counter`` an
        int64 CUDA tensor of 2 zeros that the kernel advances (one step per call; graph-replay safe).  Returns int32 [N, 2]."""
        assert scores.is_cuda and scores.dtype == torch.float32 and scores.dim() == 2 and scores.stride(1) == 1 and scores.shape[0] == 2 * self.n_envs
        assert counter.is_cuda and counter.dtype == torch.int64 and counter.numel() == 2 and counter.is_contiguous()
        if out is None:
            out = torch.empty((self.n_envs, 2), dtype=torch.int32, device=self.device)
        assert out.is_cuda and out.dtype == torch.int32 and out.is_contiguous() and out.numel() == 2 * self.n_envs
        _native.check(self._lib.ovc_sample_actions(scores.data_ptr(), scores.stride(0), 6, 2 * self.n_envs, int(seed) & (2**64 - 1),
                                                   counter.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def accumulate_returns(self, ret_sparse, ret_mixed, factor=1.0):
        """``ret_sparse += sparse`` (int64 [N]) and ``ret_mixed += sparse + factor * (shaped[:, 0] + shaped[:, 1])`` (float32 [N])
        from the last ``step``'s outputs, in one kernel (ovc_accumulate_returns; rllib.py:328-329)."""
        for t, dt in ((ret_sparse, torch.int64), (ret_mixed, torch.float32)):
            assert t is None or (t.is_cuda and t.dtype == dt and t.is_contiguous() and t.numel() == self.n_envs)
        _native.check(self._lib.ovc_accumulate_returns(self.sparse.data_ptr(), self.shaped.data_ptr(), float(factor), self.n_envs,
                                                       0 if ret_sparse is None else ret_sparse.data_ptr(),
                                                       0 if ret_mixed is None else ret_mixed.data_ptr(), self._stream()))

    def feature_lut(self):
        if self._lut is None:
            lut = np.stack([l.feature_lut() for l in self.layouts]).view(np.uint8).reshape(self.n_layouts, -1)
            assert lut.shape[1] == 1024 * self._lib.ovc_feat_lut_entry_size()
            self._lut = torch.from_numpy(lut).to(self.device)
        return self._lut

    def featurize_state(self, num_pots=2, out=None, view_swap=None):
        """featurize_state (overcooked_mdp.py:2579-2898; default NO_COUNTERS_PARAMS planner):
        float32 [N, 2, 2*(10*num_pots+28)]."""
        F = 2 * (10 * num_pots + 28)
        if view_swap is not None:
            assert view_swap.dtype == torch.int32 and view_swap.is_cuda and view_swap.is_contiguous() and view_swap.numel() == self.n_envs
        if out is None:
            out = torch.empty((self.n_envs, 2, F), dtype=torch.float32, device=self.device)
        assert out.dtype == torch.float32 and out.is_cuda and out.is_contiguous() and out.numel() == self.n_envs * 2 * F
        _native.check(self._lib.ovc_featurize(
            self.tables.data_ptr(), self.n_layouts, self.feature_lut().data_ptr(), self.state.data_ptr(),
            0 if view_swap is None else view_swap.data_ptr(), out.data_ptr(), self.n_envs, self.state_words, num_pots,
            self._stream()))
        return out

    def potential(self, gamma=0.99, out=None):
        """potential_function (overcooked_mdp.py:2920-3250) of every environment: float64 [N], bit-identical
        to the reference's Python floats (planner costs from the default NO_COUNTERS_PARAMS planner)."""
        if getattr(self, "_pot_gamma", None) != gamma:
            pt, cl, gpow = L.build_potential_tables(self.layouts, gamma)
            assert pt.shape[1] == self._lib.ovc_potential_table_size()
            self._pot = (torch.from_numpy(pt).to(self.device), torch.from_numpy(cl).to(self.device),
                         torch.from_numpy(gpow).to(self.device))
            self._pot_gamma = gamma
        if out is None:
            out = torch.empty(self.n_envs, dtype=torch.float64, device=self.device)
        assert out.dtype == torch.float64 and out.is_cuda and out.is_contiguous() and out.numel() == self.n_envs
        pt, cl, gpow = self._pot
        _native.check(self._lib.ovc_potential(
            self.tables.data_ptr(), self.n_layouts, pt.data_ptr(), cl.data_ptr(), gpow.data_ptr(), gpow.numel(),
            self.state.data_ptr(), out.data_ptr(), self.n_envs, self.state_words, self._stream()))
        return out

    # ---------------------------------------------------------------------------------------------
    def sparse_by_agent(self, events, layout_ids=None):
        """Per-agent delivery reward int32 [..., N, 2] from the event words (bits 25-28 carry the delivered recipe).
        ``layout_ids`` (int [N]): the layouts the events were produced on; default the initial assignment — with
        ``random_layout`` pass ``layout_ids()`` taken BEFORE the step (a finished environment has moved on)."""
        val = torch.from_numpy(np.stack([l.deliver_value for l in self.layouts]).astype(np.int32)).to(events.device)
        rec = (events >> L.EV_RECIPE_SHIFT) & 15
        ids = self.env_layout if layout_ids is None else layout_ids
        lid = ids.long().view(*([1] * (events.dim() - 2)), -1, 1).expand_as(rec)
        return val[lid, rec.long()] * ((events >> 15) & 1)

    def get_states(self, indices=None):
        """Unpack records into OvercookedState objects (host; for debugging / the drop-in adapters)."""
        recs = self.state.cpu().numpy()
        idx = range(self.n_envs) if indices is None else indices
        return [L.unpack_state(self.layouts[int(recs[i][3]) & 0xFF], recs[i]) for i in idx]

    def set_states(self, states, indices=None):
        idx = list(range(self.n_envs)) if indices is None else list(indices)
        lids = self._layout_ids_host()
        recs = np.stack([L.pack_state(self.layouts[lids[i]], s, int(lids[i]), self.state_words) for i, s in zip(idx, states)])
        self.state[torch.as_tensor(idx, device=self.device)] = torch.from_numpy(recs).to(self.device)


class PassTicket(object):
    """Completion handle of one HostRolloutPipeline pass submitted with wait=False."""

    def __init__(self, pipe, ticket):
        self._pipe, self.ticket = pipe, ticket

    def synchronize(self):
        """Block the host until the pass's last device->host copy has landed."""
        _native.check(self._pipe._lib.ovc_pipeline_wait(self._pipe._handle, self.ticket))


class HostRolloutPipeline(object):
    """Rollout collection with HOST buffers: the end-to-end path a host-side policy / learner sees.

    ``run(actions_host[T,N,2])`` hands pinned host buffers to the native driver (``ovc_pipeline_run``), which
    copies the action trace host->device in chunks of ``chunk`` steps (copy stream), advances the environments with
    the fused rollout kernel (compute stream) and copies sparse / shaped / done / events device->host (second copy
    stream), the three stages overlapped across chunks with double buffering — and across successive passes with
    ``wait=False``.  Returns pinned host tensors (sparse[T,N], shaped[T,N,2], done[T,N], events[T,N,2]).  Per
    environment-step this moves 8 bytes host->device and 24 bytes device->host (2 + 13 with ``narrow``, 2 + 6 with
    ``packed``, 1 + 2 with ``codes``).
    ``run`` is stream ordered like every other call here: the returned tensors are complete once the current
    stream has been synchronised (``torch.cuda.current_stream().synchronize()``), not when ``run`` returns.
    """

    def __init__(self, env, n_steps, chunk=50, narrow=False, packed=False, codes=False, host_buffers=1, stream=False,
                 stream_fill=0.25, packed_actions=True):
        """narrow=True: uint8 actions in, int16 sparse / int8 shaped / uint8 done / int32 events out — the same
        values in 2 + 13 instead of 8 + 24 bytes per env-step.  packed=True: uint8 actions in, int16 sparse / int8
        shaped / int16 event codes (+done) out — 2 + 6 bytes per env-step, lossless (wire.decode_event_codes).
        codes=True: one byte of joint action in (wire.pack_actions, actions_host uint8 [T,N]), one int16 word of
        event codes + done + reward-grant bits out — 1 + 2 bytes per env-step, lossless (env.expand_codes).
        host_buffers: number of pinned output sets, used round robin by successive run() calls (2 lets a consumer
        read pass i while pass i+1 is in flight, see run(wait=False))."""
        """stream=True: the result as a sparse event stream (OVC_F_OUT_STREAM): per transition one lane mask per 32
        environments + the non-zero code words, ``stream_fill`` x 32 x chunk value slots per group and chunk (0.125 + 2 x
        stream_fill bytes per env-step device->host instead of 2); run() returns (masks, values) host tensors and
        ``expand(result)`` rebuilds dense arrays, falling back to the dense code words kept on the device for any chunk
        whose group overflowed.  packed_actions=False (with codes / stream): uint8 [T, N, 2] actions instead of one byte
        per joint action."""
        codes = codes or stream
        narrow = narrow or packed or codes
        self.env, self.T, self.chunk, self.narrow, self.packed = env, int(n_steps), int(chunk), bool(narrow), bool(packed)
        self.codes, self.stream = bool(codes), bool(stream)
        self.packed_actions = bool(packed_actions) and self.codes
        if narrow:
            assert env.narrow_ok(), "rewards of these layouts do not fit the narrow formats"
        N, dev = env.n_envs, env.device
        self._lib = env._lib
        self.act_dtype = torch.uint8 if narrow else torch.int32
        self.act_shape = (N,) if self.packed_actions else (N, 2)
        self.n_chunks = -(-self.T // self.chunk)
        self.stream_cap = max(1, min(_native.STREAM_CAP_MAX, int(round(self.chunk * 32 * float(stream_fill))))) if stream else 0
        with torch.cuda.device(dev):
            self.d_act = [torch.empty((self.chunk,) + self.act_shape, dtype=self.act_dtype, device=dev) for _ in range(2)]
            if stream:
                G = env.n_groups()
                self.d_out = [(torch.empty((G, self.stream_cap), dtype=torch.int16, device=dev), None, None,
                               torch.empty((self.chunk, G), dtype=torch.int32, device=dev)) for _ in range(2)]
                # dense code words of a whole pass stay on the device (two sets: pass k uses set k & 1): overflow backup
                self.d_codes_full = [torch.empty((self.T, N), dtype=torch.int16, device=dev) for _ in range(2)]
                self.h_outs = [(torch.zeros((self.n_chunks, G, self.stream_cap), dtype=torch.int16, pin_memory=True), None, None,
                                torch.zeros((self.T, G), dtype=torch.int32, pin_memory=True)) for _ in range(max(1, int(host_buffers)))]
            else:
                self.d_out = [env.alloc_rollout_out(self.chunk, narrow=narrow, packed=packed, codes=codes) for _ in range(2)]
                self.h_outs = [env.alloc_rollout_out(self.T, narrow=narrow, pin=True, packed=packed, codes=codes)
                               for _ in range(max(1, int(host_buffers)))]
        self.h_out = self.h_outs[0]
        self._runs = 0
        self.h2d_bytes_per_step = N * (1 if self.packed_actions else 2) * self.d_act[0].element_size()
        if stream:
            self.d2h_bytes_per_step = (self.h_out[3].numel() * 4 + self.h_out[0].numel() * 2) / float(self.T)
        else:
            self.d2h_bytes_per_step = sum(N * (2 if o.dim() == 3 else 1) * o.element_size() for o in self.h_out if o is not None)
        flags = env._flags() | (_native.F_ACT_PACKED if self.packed_actions else _native.F_ACT_U8 if narrow else 0)
        flags |= (_native.F_OUT_STREAM if stream else _native.F_OUT_CODES if codes else _native.F_OUT_PACKED if packed
                  else _native.F_OUT_NARROW if narrow else 0)
        d = _native.PipelineDesc()
        d.layouts, d.n_layouts, d.state_words = env.tables.data_ptr(), env.n_layouts, env.state_words
        d.start_records, d.state, d.n_envs = env.start_records.data_ptr(), env.state.data_ptr(), N
        d.horizon, d.flags, d.chunk = env.horizon, flags, self.chunk
        d.has_random_start = int(env._rs is not None)
        if env._rs is not None:
            d.random_start = env._rs
        ptr = lambda t: 0 if t is None else t.data_ptr()
        for b in range(2):
            d.d_actions[b] = self.d_act[b].data_ptr()
            d.d_sparse[b], d.d_shaped[b], d.d_done[b], d.d_events[b] = (ptr(t) for t in self.d_out[b])
            d.d_codes_full[b] = self.d_codes_full[b].data_ptr() if stream else 0
        d.stream_cap = self.stream_cap
        h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _native.check(self._lib.ovc_pipeline_create(ctypes.byref(d), ctypes.byref(h)))
        self._handle = h

    def run(self, actions_host, wait=True):
        """wait=True: the current stream waits for the pass (stream-ordered like every other call); returns the
        host tensors.  wait=False: returns (host tensors, PassTicket) without joining the current stream, so the
        next run() starts its copies while this pass is still draining — the steady state of a collection
        loop; ``ticket.synchronize()`` before reading, and ``join()`` before touching the env from the current
        stream again."""
        assert actions_host.dtype == self.act_dtype and actions_host.is_pinned() and actions_host.is_contiguous()
        assert tuple(actions_host.shape) == (self.T,) + self.act_shape
        h_out = self.h_outs[self._runs % len(self.h_outs)]
        self._last_set = self._runs & 1  # which dense-backup set this pass writes (the native side counts the same way)
        self._runs += 1
        ptr = lambda t: 0 if t is None else t.data_ptr()
        ticket = ctypes.c_int64(-1)
        _native.check(self._lib.ovc_pipeline_run(
            self._handle, actions_host.data_ptr(), ptr(h_out[0]), ptr(h_out[1]), ptr(h_out[2]), ptr(h_out[3]), self.T,
            self.env._stream(), int(bool(wait)), ctypes.byref(ticket)))
        self._last_actions = actions_host  # keep the source alive until the copies have run
        if wait:
            return h_out
        return h_out, PassTicket(self, ticket.value)

    def expand(self, h_out, codes_set=None, out=None, n_threads=0, **which):
        """stream=True: dense host arrays (dict, as env.expand_codes) from one pass's (values, -, -, masks) host tensors,
        AFTER the pass has landed.  ``codes_set``: the dense-backup set that pass wrote (``self._last_set`` right after
        its run()); if a group overflowed its value slots the dense words are fetched from that set and expanded instead
        (correct as long as no later pass has reused the set: passes k and k + 2 share one)."""
        assert self.stream
        dense, over = self.env.expand_stream(h_out[3], h_out[0], chunk=self.chunk, out=out, n_threads=n_threads, **which)
        self.last_overflow = over
        if over:
            cs = self._last_set if codes_set is None else codes_set
            words = self.d_codes_full[cs].cpu()  # synchronising copy: the rare slow path
            dense = self.env.expand_codes(words, out=dense, n_threads=n_threads)
        return dense

    def join(self):
        """Make the current stream wait for everything the pipeline has in flight."""
        _native.check(self._lib.ovc_pipeline_join(self._handle, self.env._stream()))

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.ovc_pipeline_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EpisodeStats(object):
    """Running per-environment episode statistics on the device — the batched form of OvercookedEnv.game_stats
    (overcooked_env.py:308-319, 382-401) and of the ``episode`` entry of the info dict (:363-380).

    The reference keeps, per event and agent, the LIST of timesteps at which it fired; its consumers only ever
    take ``len()`` of those lists (rllib.py:480-483), so counts are kept: ``event_counts[N, 2, 25]``, plus
    ``cumulative_sparse_rewards_by_agent[N, 2]``, ``cumulative_shaped_rewards_by_agent[N, 2]`` and ``ep_length[N]``.
    ``update`` returns the finished environments' statistics (a dict of tensors, rows selected by ``done``) and
    clears them for the next episode.
    """

    def __init__(self, env):
        self.env = env
        N, dev = env.n_envs, env.device
        self.event_counts = torch.zeros((N, 2, 25), dtype=torch.int32, device=dev)
        self.cumulative_sparse_rewards_by_agent = torch.zeros((N, 2), dtype=torch.int64, device=dev)
        self.cumulative_shaped_rewards_by_agent = torch.zeros((N, 2), dtype=torch.int64, device=dev)
        self.ep_length = torch.zeros(N, dtype=torch.int32, device=dev)
        self._bits = torch.arange(25, device=dev, dtype=torch.int32)
        self._lid = env.layout_ids()  # layouts of the running episodes (they change at resets with random_layout)

    def update(self, sparse, shaped, done, events):
        """Feed the outputs of one step() (tensors [N], [N,2], [N], [N,2]); call it after EVERY step."""
        self.cumulative_sparse_rewards_by_agent += self.env.sparse_by_agent(events, self._lid)
        if self.env.random_layout:
            self._lid = self.env.layout_ids()
        self.cumulative_shaped_rewards_by_agent += shaped
        self.event_counts += (events.unsqueeze(-1) >> self._bits) & 1
        self.ep_length += 1
        d = done != 0
        finished = None
        if bool(d.any()):
            idx = torch.nonzero(d).squeeze(1)
            finished = {
                "env_index": idx,
                "ep_game_stats": self.event_counts[idx].clone(),
                "ep_sparse_r_by_agent": self.cumulative_sparse_rewards_by_agent[idx].clone(),
                "ep_shaped_r_by_agent": self.cumulative_shaped_rewards_by_agent[idx].clone(),
                "ep_sparse_r": self.cumulative_sparse_rewards_by_agent[idx].sum(1),
                "ep_shaped_r": self.cumulative_shaped_rewards_by_agent[idx].sum(1),
                "ep_length": self.ep_length[idx].clone(),
            }
            self.event_counts[idx] = 0
            self.cumulative_sparse_rewards_by_agent[idx] = 0
            self.cumulative_shaped_rewards_by_agent[idx] = 0
            self.ep_length[idx] = 0
        return finished
