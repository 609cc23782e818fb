"""The fused reverse-SDE loop: N steps of (score net -> predictor update -> noise) as one replayed CUDA graph.

Replaces the Python ``pc_sampling`` loop of the reference (diffusion/diffusion_continuous.py:231-258), which
issues ~700 kernel launches per step.  Three things make one captured step replayable N times:

* every per-step scalar lives in a device table indexed by a device-side step counter (``ldt_sde_step``,
  ``ldt_advance_step``);
* in unconditional sampling every sample shares the step's time value, so TimeEmbedding and all 25 adaLN
  projections are batch-invariant: they are computed for ALL N timesteps up front as one GEMM
  ([N, t_dim] x [t_dim, 6*hidden*blocks + 2*hidden]) and each step just selects its row (``ldt_select_row``);
* the per-step noise is generated inside the update kernel with the Philox counter layout of
  ``torch.randn_like`` on the CUDA default generator, and the generator offset is advanced afterwards by what
  the reference would have consumed, so the stream position seen by later torch calls matches.
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import ops
from ._lib import PRED_CORRECTOR
from .sde import _PRED_CODES, torch_randn_launch_geometry


_warned: set = set()
_verified = weakref.WeakSet()   # score_fn function objects whose probe passed (one probe per function, not per sample() call)


def _warn_once(key: str, msg: str) -> None:
    if key not in _warned:
        _warned.add(key)
        import warnings
        warnings.warn(msg, RuntimeWarning, stacklevel=3)


def find_score_module(score_fn, sde, probe=None):
    """Return the ldt_b200.Score behind a Trainer.score_fn closure (trainer/Latent_SDE_Trainer.py:57-61), or None.

    The fused loop hard-wires ``params = model(x, t, label, condition); score = -params / sqrt(SDE.var(t))``.  It is
    taken only when (1) the callable is a bound method of an object whose ``.model`` is our Score and whose ``.SDE`` is
    this SDE object (all three reference trainers), and (2) with ``probe = (t, x, label, condition)`` one real call of
    ``score_fn`` returns exactly that pair, bit for bit -- a subclass that overrides ``score_fn`` (rescaled score,
    guidance, another parameterisation) fails the probe and gets the generic per-step path with a one-time warning
    instead of being silently replaced.  The probe draws no random numbers."""
    from .score import Score
    owner = getattr(score_fn, "__self__", None)
    model = getattr(owner, "model", None) if owner is not None else None
    if not isinstance(model, Score):
        return None
    if getattr(owner, "SDE", None) is not sde:
        _warn_once("sde", "ldt_b200: score_fn is bound to an ldt_b200.Score but to another SDE object; sampling runs "
                          "the generic per-step path, not the fused CUDA-graph loop")
        return None
    func = getattr(score_fn, "__func__", None)
    if probe is not None and func is not None and func in _verified:
        probe = None    # this very function object already reproduced the hard-wired pair on a real call
    if probe is not None:
        t, x, label, condition = probe
        try:
            got = score_fn(t, x, label=label, condition=condition)
            tx = t.to(x)
            params = model(x, tx, label=label, condition=condition)
            score = -params / torch.sqrt(sde.var(tx))[:, None, None]
            ok = (isinstance(got, (tuple, list)) and len(got) == 2 and torch.equal(got[1], params)
                  and torch.equal(got[0], score))
        except Exception:   # a closure with another signature: not ours to fuse
            ok = False
        if ok and func is not None:
            try:
                _verified.add(func)
            except TypeError:
                pass
        if not ok:
            _warn_once("probe", "ldt_b200: score_fn does not compute (-model(x, t) / sqrt(SDE.var(t)), model(x, t)); sampling "
                                "runs the generic per-step path, not the fused CUDA-graph loop")
            return None
    return model


def modulation_table(score, P, timesteps: torch.Tensor, chunk: int = 256) -> torch.Tensor:
    """AdaLN rows for every timestep: f32 [N, 6*hidden*blocks + 2*hidden]."""
    N = timesteps.shape[0]
    dev = timesteps.device
    half = (score.t_dim // 4) // 2
    mod_len = score.mod_len
    table = torch.empty((N, mod_len), dtype=torch.float32, device=dev)
    w0, b0, w1, b1 = P["te"]
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        R = e - s
        c = torch.empty((R, score.t_dim), dtype=torch.float32, device=dev)
        sc = torch.empty((R, score.t_dim), dtype=torch.bfloat16, device=dev)
        scratch = torch.empty((R, score.t_dim + 2 * half), dtype=torch.float32, device=dev)
        ops.time_embedding(timesteps[s:e].contiguous(), P["freq"], w0, b0, w1, b1, None, c, sc, scratch)
        if score.precision == "tf32":
            Q = score.packed_tf32()
            ops.gemm(ops.round_pad_tf32(c, silu=True), Q["w_ada"], Q["b_ada"], table[s:e], ops.EPI_BIAS_F32)
        elif score.precision == "fp32":
            Q = score.packed_tf32()
            ops.gemm(ops.split_tf32(c, silu=True), Q["w_ada"], Q["b_ada"], table[s:e], ops.EPI_BIAS_F32, split_operands=True)
        else:
            ops.gemm(sc, P["w_ada"], P["b_ada"], table[s:e], ops.EPI_BIAS_F32)
    return table


def time_embedding_table(score, P, timesteps: torch.Tensor) -> torch.Tensor:
    """TimeEmbedding(t_i) for every timestep: f32 [N, t_dim] (model/layers.py:38-41)."""
    N, dev = timesteps.shape[0], timesteps.device
    half = (score.t_dim // 4) // 2
    w0, b0, w1, b1 = P["te"]
    table = torch.empty((N, score.t_dim), dtype=torch.float32, device=dev)
    sc = torch.empty((N, score.t_dim), dtype=torch.bfloat16, device=dev)
    scratch = torch.empty((N, score.t_dim + 2 * half), dtype=torch.float32, device=dev)
    ops.time_embedding(timesteps.contiguous(), P["freq"], w0, b0, w1, b1, None, table, sc, scratch)
    return table


class StepGraph:
    """One captured sampler step, replayable; owns the loop state buffers."""

    def __init__(self, score, sde, B, N, predictor, time_eps, probability_flow, device, use_graph=True,
                 per_sample_c=False, cross_attention=False, corrector_steps=0, snr=0.0):
        self.score, self.B, self.N = score, B, N   # (the plan cache holds the only strong reference to a plan)
        # AncestralCorrector (:212-229): corrector_steps extra (score evaluation + update) pairs per step, each with
        # its own randn_like draw, so one step consumes 1 + corrector_steps Philox launches' worth of offsets
        self.corrector_steps = corrector_steps
        self.ccoef = sde.corrector_coefficients(N, time_eps, snr, device) if corrector_steps > 0 else None
        P = score.packed()
        self.P = P
        self.coef, self.timesteps = sde.step_coefficients(predictor, N, time_eps, probability_flow, device)
        # Conditional sampling (score.py:134-135,148-149).  per_sample_c: c = t_emb(t_i) + extra[b] differs per sample,
        # so the adaLN GEMM runs every step on [B, t_dim] (1.6 % of the step's MACs at B = 64) from a [N, t_dim] table
        # of time embeddings.  cross_attention: even blocks attend to the condition tokens, whose K/V are projected
        # once per run (the reference re-projects them every step).  Both live in buffers owned by this plan, so the
        # captured graph stays valid across runs with new conditions.
        self.per_sample_c, self.cross_attention = per_sample_c, cross_attention
        self.extra = torch.zeros((B, score.t_dim), dtype=torch.float32, device=device) if per_sample_c else None
        self.kv_cond = None
        if cross_attention:
            self.kv_cond = [torch.empty((B * score.z_scale, 2 * score.hidden_size), dtype=torch.bfloat16, device=device)
                            for _ in range((score.num_blocks + 1) // 2)]
        if per_sample_c:
            self.table = time_embedding_table(score, P, self.timesteps)
            self.ws = score._workspace(B, B, device)
        else:
            self.table = modulation_table(score, P, self.timesteps)
            self.ws = score._workspace(B, 1, device)
        self.code = _PRED_CODES[predictor]
        T, D = score.z_scale, score.z_dim
        self.x = torch.empty((B, T, D), dtype=torch.float32, device=device)
        self.x_mean = torch.empty_like(self.x)
        self.params = torch.empty_like(self.x)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        self.mod_cur = torch.empty((1, self.table.shape[1]), dtype=torch.float32, device=device)
        self.rng_grid, self.offset_per_step = torch_randn_launch_geometry(self.x.numel(), device)
        # Philox {seed, base offset} live in device memory (read by ldt_sde_step at run time), so the step graph is
        # captured once per plan and replayed from any generator position
        self.rng_state = torch.zeros(2, dtype=torch.int64, device=device)
        self.captures = 0
        self.graph = None
        self.use_graph = use_graph
        # LDT_C_LOOP=1 (or use_c_loop = True): the whole N-step loop as ONE call of the C entry point ldt_sample_loop, which
        # captures and replays the step graph itself -- the path a non-Python host takes.  Same kernels, same bits.
        import os
        self.use_c_loop = os.environ.get("LDT_C_LOOP", "0") == "1"

    def set_condition(self, cond_tokens, extra) -> None:
        """Load this run's condition into the plan's buffers (pointers captured by the graph stay the same)."""
        if self.per_sample_c:
            self.extra.copy_(extra.float().expand(self.B, self.score.t_dim))
        if self.cross_attention:
            for dst, kv in zip(self.kv_cond, self.score.project_condition_tokens(self.P, cond_tokens)):
                dst.copy_(kv)

    def _step_body(self):
        B, T, D = self.x.shape
        if self.per_sample_c:
            ws = self.ws
            ops.cond_silu(self.table, self.step, self.extra, None, ws.sc)
            ops.gemm(ws.sc, self.P["w_ada"], self.P["b_ada"], ws.mod, ops.EPI_BIAS_F32)
            mod, mod_stride = ws.mod, ws.mod_len
        else:
            ops.select_row(self.table, self.step, self.mod_cur)
            mod, mod_stride = self.mod_cur, 0
        self.score.run_tokens(self.P, self.ws, self.x.view(B * T, D), mod, mod_stride, self.params.view(B * T, D),
                              self.kv_cond)
        draws = 1 + self.corrector_steps
        ops.sde_step(self.code, self.x, self.params, None, self.coef, self.step, 0, 0,
                     draws * self.offset_per_step, self.rng_grid, self.x, self.x_mean, rng_state=self.rng_state)
        for j in range(self.corrector_steps):
            self.score.run_tokens(self.P, self.ws, self.x.view(B * T, D), mod, mod_stride, self.params.view(B * T, D),
                                  self.kv_cond)
            ops.sde_step(PRED_CORRECTOR, self.x, self.params, None, self.ccoef, self.step, 0,
                         (1 + j) * self.offset_per_step, draws * self.offset_per_step, self.rng_grid,
                         self.x, self.x_mean, rng_state=self.rng_state)
        ops.advance_step(self.step)

    def run(self, x0: torch.Tensor, seed: int, offset: int, record_every: int | None = None) -> list:
        """Run all N steps from x0 (copied into the loop buffer) with Philox (seed, offset).  With record_every = s
        returns the x_mean snapshots after steps s, 2s, ... (the print_steps trajectory, :239-257)."""
        def i64(v):   # unsigned 64-bit pattern as the int64 torch stores
            v &= (1 << 64) - 1
            return v - (1 << 64) if v >= (1 << 63) else v

        self.x.copy_(x0)
        self.step.zero_()
        self.rng_state.copy_(torch.tensor([i64(seed), i64(offset)], dtype=torch.int64))
        snaps = []
        if self.use_c_loop and not (self.per_sample_c or self.cross_attention or self.corrector_steps or record_every):
            plan = self.score.c_plan(self.P, self.ws)
            if plan is not None:
                from . import _lib
                args = _lib.SampleArgs(score=C.pointer(plan), predictor=self.code, num_steps=self.N, use_graph=int(self.use_graph),
                                       mod_table=self.table.data_ptr(), mod_len=self.table.shape[1], mod_cur=self.mod_cur.data_ptr(),
                                       coef=self.coef.data_ptr(), step=self.step.data_ptr(), rng_state=self.rng_state.data_ptr(),
                                       offset_per_step=self.offset_per_step, rng_grid=self.rng_grid, x=self.x.data_ptr(),
                                       x_mean=self.x_mean.data_ptr(), params=self.params.data_ptr())
                per_step = 4 + 6 * self.score.num_blocks + 3     # token pass + select_row + sde_step + advance_step
                cur = torch.cuda.current_stream(self.x.device)
                side = cur if cur.cuda_stream != 0 else torch.cuda.Stream(self.x.device)   # the legacy stream cannot capture
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    ops.sample_loop(args, self.N * per_step)
                cur.wait_stream(side)
                self.launches_per_step = per_step
                return snaps
        if not self.use_graph:
            for i in range(self.N):
                self._step_body()
                if record_every and (i + 1) % record_every == 0:
                    snaps.append(self.x_mean.clone())
            return snaps
        if self.graph is None:
            self._step_body()  # warm-up outside capture (lazy module loading, attribute setup)
            self.x.copy_(x0)
            self.step.zero_()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                self._step_body()
            self.launches_per_step = ops.launch_count() - n0
            ops.add_launches(-self.launches_per_step)  # captured, not executed
            self.graph = g
            self.captures += 1
            # the capture itself does not execute; state is still (x0, step 0)
        for i in range(self.N):
            self.graph.replay()
            if record_every and (i + 1) % record_every == 0:
                snaps.append(self.x_mean.clone())
        ops.add_launches(self.N * self.launches_per_step)
        return snaps


_graph_cache: dict = {}


def fused_sample_loop(score, sde, x0, N, predictor, time_eps, probability_flow, denoise, use_graph=True,
                      cond_tokens=None, extra=None, print_steps=None, corrector_steps=0, snr=0.0):
    """x0 [B, z_scale, z_dim] on the device -> latent after N reverse steps (x_mean if denoise else x).

    cond_tokens [B, hidden, z_scale] (condition[0], cross-attended by the even blocks) and extra [B, t_dim]
    (condition[1] or the label embedding, added to the time embedding) select conditional sampling."""
    device = x0.device
    B = x0.shape[0]
    per_sample_c, cross = extra is not None, cond_tokens is not None
    if score.precision in ("tf32", "fp32") and (per_sample_c or cross):
        raise NotImplementedError("precision='tf32' / 'fp32' sample conditionally through the per-step path only")
    key = (score.precision, B, N, predictor, float(time_eps), bool(probability_flow), device, use_graph,
           per_sample_c, cross, corrector_steps, float(snr), score._fingerprint())
    sg = _graph_cache.get(key)
    # the plan belongs to these two objects: weak references, not id() (an id can be reused after garbage collection)
    if sg is not None and (sg.score_ref() is not score or sg.sde_ref() is not sde):
        sg = None
    if sg is None:
        _graph_cache.clear()  # one live plan: the buffers are large (modulation table ~0.6 GB at N=1000)
        sg = StepGraph(score, sde, B, N, predictor, time_eps, probability_flow, device, use_graph,
                       per_sample_c=per_sample_c, cross_attention=cross, corrector_steps=corrector_steps, snr=snr)
        sg.score_ref, sg.sde_ref = weakref.ref(score), weakref.ref(sde)
        _graph_cache[key] = sg
    if per_sample_c or cross:
        sg.set_condition(cond_tokens, extra)
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    record_every = (N - 1) // (print_steps - 2) if print_steps is not None else None
    snaps = sg.run(x0, seed, offset, record_every)
    # the reference draws one randn_like per predictor / corrector update from this generator (:160,204,224); leave it
    # where it would be
    gen.set_offset(offset + N * (1 + corrector_steps) * sg.offset_per_step)
    final = (sg.x_mean if denoise else sg.x).clone()
    if print_steps is not None:
        return [x0] + snaps + [final]
    return final
