"""Batched APG-MPC: P independent ``sde4mbrl/apg.py`` solves advanced in lock step on the device.

This is BASELINE config 4 (4096 independent hexacopter MPC solves) and SURVEY 8(f) rank 1.  The
reference vmaps ``apg_mpc`` over problems, which turns every ``lax.cond`` into a select and
evaluates all branches; here the same semantics are obtained with per-problem masks:

    candidates (maxls+1 step sizes per problem)  -> ONE value-only launch over (maxls+1)*P problems
    Armijo selection + box prox + momentum point -> ONE value-only launch over 2P problems (x and v)
    ynext = where(cost_x <= cost_v, x, v)        -> ONE value+gradient launch over P problems
    state update (masked by not_done)
For latency-bound sizes (2*P*N <= 8192 samples) the last two launches are merged: value+gradient for x AND v, then
point, cost and gradient of the better one are copied -- ~15 % more arithmetic, one sequential rollout less per
iteration (a single hexacopter solve: 3.8 -> 2.9 ms).  Smaller still (Q*P*N <= 4096 samples, Q = 2(maxls+1) [+ maxls with
a prox]) the iteration is SPECULATIVE: the x_k / v_k of every step size are evaluated with their gradients in the one
launch that also yields the line search's costs, and the accepted step's results are gathered afterwards -- one
sequential rollout per iteration.

All arithmetic (cost evaluations and the APG bookkeeping) runs in ``libsde4b200.so``; torch only
owns the buffers.  Arrays are problem-minor: ``[H, n_opt, P]``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .apg import APGState
from .engine import as_f32, as_key, _ptr, _stream

# speculative iterations while (points per problem) * P * N stays at or below this many samples (scripts/apg_spec.py;
# refitted with scripts/apg_modes2.py on the warp-specialised kernel: 14 points x 384 problems = 5 376 samples 2.81 ms
# against 3.18 ms merged, 6 720 samples 2.93 vs 3.21 ms, 7 168 samples 3.11 vs 3.14 ms, 8 960 samples 3.45 vs 3.33 ms)
SPECULATIVE_MAX_SAMPLES = 7168


class BatchedAPG:
    def __init__(self, engine, nn_params, stage, optimizer_params, time_steps, num_particles, lb=None, ub=None,
                 terminal=None, lanes=0, merge_xv_grad=None, progressive_ls=None, speculative=None):
        self.eng = engine
        self.pk = engine.packed(nn_params).private_copy()  # this solver's own device block (see set_params)
        self._ws = None
        self.stage = stage
        self.terminal = terminal
        self.op = optimizer_params
        assert "linesearch" in optimizer_params, "the batched solver implements the line-search variant (apg.py:174-187)"
        self.ts = np.asarray(time_steps, np.float32)
        self.H = int(self.ts.shape[0])
        self.N = int(num_particles)
        self.n_opt = engine.n_u + stage.desc.n_slack
        # lanes per sample of the three cost launches of an iteration: one int for all (0 = the library
        # picks from the launch size) or {"cand": .., "xv": .., "y": ..}
        self.lanes = lanes if isinstance(lanes, dict) else {"cand": lanes, "xv": lanes, "y": lanes}
        self.lb, self.ub = lb, ub
        self.merge_xv_grad = merge_xv_grad  # None: decided from the launch size (see step)
        self.progressive_ls = progressive_ls  # None: decided from the launch size (see _alloc)
        self.speculative = speculative  # None: decided from the launch size (see _alloc)
        self._modes_given = merge_xv_grad is not None or progressive_ls is not None  # an explicit two-launch request wins
        self.P = None
        self._graph = None

    # ------------------------------------------------------------------
    def _alloc(self, P, dev):
        H, n, K = self.H, self.n_opt, self.op["linesearch"]["maxls"] + 1
        ls_ = self.op["linesearch"]
        self.P, self.K = P, K
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        i = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
        self.t = {n_: i(P) for n_ in L.APG_STATE_INT}
        self.t.update({n_: f(P) for n_ in L.APG_STATE_F32})
        self.t.update({n_: f(H, n, P) for n_ in L.APG_STATE_VEC})
        self.cand, self.svals, self.cand_cost = f(H, n, K * P), f(K, P), f(K * P)
        self.xv, self.xv_cost = f(H, n, 2 * P), f(2 * P)
        if self.merge_xv_grad is None:
            # scripts/apg_modes.py / apg_modes2.py: better up to 4 k samples (3.9 vs 4.7 ms at 2048 x 1), worse from 8 k
            # (6.4 vs 5.9 ms at 4096 x 1) and clearly worse from 32 k
            self.merge_xv_grad = 2 * P * self.N < 8192
        if self.progressive_ls is None:
            # worth it when the launches are throughput bound and every warp belongs to one problem
            # (scripts/apg_modes.py: better from 16 k samples per candidate on)
            self.progressive_ls = self.N % 32 == 0 and P * self.N >= 16384 and self.op["linesearch"]["maxls"] > 1
        # speculative iteration: all points an iteration can end at in ONE value+gradient launch (one sequential
        # rollout per iteration instead of two); pays while that launch is still latency bound
        self.Q = 2 * K + (ls_["maxls"] if self.lb is not None else 0)
        if self.speculative is None:
            self.speculative = (not self._modes_given) and self.Q * P * self.N <= SPECULATIVE_MAX_SAMPLES
        if self.speculative:
            self.merge_xv_grad, self.progressive_ls = True, False
            self.spec, self.spec_cost, self.spec_grad = f(H, n, self.Q * P), f(self.Q * P), f(H, n, self.Q * P)
        self.ls_active = i(P)
        self.xvmask = torch.ones(2 * P, dtype=torch.int32, device=dev)  # [x needs evaluation | v always]
        self.xv_grad = f(H, n, 2 * P) if self.merge_xv_grad else None
        self.y, self.y_cost, self.y_grad = f(H, n, P), f(P), f(H, n, P)
        self.cost_tmp = f(P)
        self.sel_step, self.sel_ls, self.pick = f(P), i(P), i(P)
        st = L.ApgState()
        for name in L.APG_STATE_INT + L.APG_STATE_F32 + L.APG_STATE_VEC:
            setattr(st, name, self.t[name].data_ptr())
        self.st = st
        ls = self.op["linesearch"]
        prm = L.ApgParams()
        prm.P, prm.H, prm.n_opt, prm.maxls = P, H, n, ls["maxls"]
        prm.reset_increase = int(ls["reset_option"] == "increase")
        prm.has_prox = int(self.lb is not None)
        ms = self.op.get("moment_scale", None)
        prm.use_moment_scale = int(ms is not None)
        prm.moment_scale = float(ms) if ms is not None else 0.0
        prm.max_no_improvement_iter = int(self.op["max_no_improvement_iter"])
        prm.coef, prm.decrease_factor = float(ls["coef"]), float(ls["decrease_factor"])
        prm.increase_factor, prm.max_stepsize = float(ls.get("increase_factor", 1.0)), float(ls["max_stepsize"])
        prm.atol, prm.rtol = float(self.op["atol"]), float(self.op["rtol"])
        if self.lb is not None:
            lbv = np.broadcast_to(np.asarray(self.lb, np.float32), (n,))
            ubv = np.broadcast_to(np.asarray(self.ub, np.float32), (n,))
            for j in range(n):
                prm.lb[j], prm.ub[j] = float(lbv[j]), float(ubv[j])
        self.prm = prm
        assert L.load().sde4_apg_scratch_floats(C.byref(prm)) <= self.cand.numel()

    def set_params(self, nn_params):
        """New parameter values for the next solves.  The packed block is copied INTO this solver's device buffer, so a
        captured CUDA graph (which holds that buffer's address) stays valid."""
        pk = self.eng.packed(nn_params)
        if pk.fingerprint != self.pk.fingerprint:
            self.pk.assign(pk)
        return self

    def _workspace(self):
        """Scratch of the cost launches, owned by this solver: the CUDA graph of an iteration bakes its address in, so
        it must not be the engine's shared grow-only buffer (a later, larger call elsewhere would replace that one and
        leave the graph pointing at freed memory)."""
        if self._ws is None:
            lib, d, ns = L.load(), C.byref(self.eng.desc), self.stage.desc.n_slack
            P, K = self.P, self.K
            need = 1 << 16
            for probs, grad in ((K * P, False), (2 * P, True), (P, True), ((self.Q * P) if self.speculative else P, True)):
                need = max(need, lib.sde4_cost_workspace_bytes(d, probs, self.N, self.H, ns, int(grad), 0))
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.y0.device if getattr(self, "y0", None) is not None
                                   else torch.device("cuda", torch.cuda.current_device()))
        return self._ws

    def _pm_view(self, buf):
        """[H, n, Q] problem-minor buffer -> [Q, H, n] strided view the cost launcher accepts."""
        return buf.permute(2, 0, 1)

    def _cost(self, buf, cost_out, grad_buf=None, data_mod=0, which="y", mask=None):
        self.eng.cost_grad(self.pk, self.y0, self._pm_view(buf), self.ts, self.xref, self.stage.desc,
                           terminal=None if self.terminal is None else self.terminal.desc, key=self.keys, N=self.N,
                           want_grad=grad_buf is not None, grad_out=None if grad_buf is None else self._pm_view(grad_buf),
                           lanes=self.lanes.get(which, 0), data_mod=data_mod, cost_out=cost_out, problem_mask=mask,
                           workspace=self._workspace())

    # ------------------------------------------------------------------
    def init(self, x0, y0, xref, keys, momentum=None, stepsize=None):
        """``init_apg`` for every problem (apg.py:53-113; like ``apg_mpc`` the warm start is NOT re-projected,
        sde_mpc_design.py:236-239).  x0 [P,H,n_opt] (or [H,n_opt] shared), y0 [P,n_x], xref [P,H+1,n_x] or
        [H+1,n_x], keys [P,2] or [2]."""
        y0 = as_f32(y0).reshape(-1, self.eng.n_x).contiguous()
        P, dev = y0.shape[0], y0.device
        if self.P != P:
            self._alloc(P, dev)
            self._graph = None
            self._ws = None
        # persistent input buffers: a captured CUDA graph keeps pointing at them across init() calls
        xref_t, keys_t = as_f32(xref).contiguous(), as_key(keys).reshape(-1, 2)
        if getattr(self, "y0", None) is None or self.y0.shape != y0.shape or self.xref.shape != xref_t.shape \
                or self.keys.shape != keys_t.shape:
            self.y0, self.xref, self.keys = y0.clone(), xref_t.clone(), keys_t.clone()
            self._graph = None
        else:
            self.y0.copy_(y0)
            self.xref.copy_(xref_t)
            self.keys.copy_(keys_t)
        x0 = as_f32(x0)
        if x0.dim() == 2:
            x0 = x0.unsqueeze(0).expand(P, -1, -1)
        self.y.copy_(x0.permute(1, 2, 0))
        self._cost(self.y, self.y_cost, grad_buf=self.y_grad)
        ls = self.op["linesearch"]
        lib = L.load()
        # (both tensors stay referenced until the launch is enqueued: as temporaries inside the call expression the first
        #  one was freed before the second was made, the caching allocator handed out the same block again, and the
        #  kernel read the step size as the momentum)
        mom_t = None if momentum is None else as_f32(momentum).reshape(-1).contiguous()
        step_t = None if stepsize is None else as_f32(stepsize).reshape(-1).contiguous()
        for t_ in (mom_t, step_t):
            assert t_ is None or t_.numel() == P, "momentum / stepsize: one value per problem"
        L.check(lib.sde4_apg_init(C.byref(self.prm), C.byref(self.st), _ptr(self.y), _ptr(self.y_cost), _ptr(self.y_grad),
                                  _ptr(mom_t), _ptr(step_t),
                                  C.c_float(ls["init_stepsize"]), C.c_float(self.op.get("beta_init", 0.25)),
                                  _ptr(self.cand), _stream()))
        return self

    def step(self):
        """One masked ``one_step_apg`` for all problems (apg.py:164-274)."""
        lib, prm, st, P = L.load(), self.prm, self.st, self.P
        if self.speculative:
            return self._step_speculative()
        reuse_x = not self.merge_xv_grad  # cost(x) = cost(accepted candidate) unless the prox moved it
        xmask = self.xvmask[:P] if reuse_x else None
        L.check(lib.sde4_apg_candidates(C.byref(prm), C.byref(st), _ptr(self.cand), _ptr(self.svals), _ptr(xmask), _stream()))
        # the cost of the LAST candidate is never looked at (armijo_line_search tests k = 0..maxls-1 and falls back
        # to candidate maxls unseen, apg.py:362-379): evaluate the first maxls candidates only
        if self.progressive_ls:
            # throughput-bound sizes whose problems fill whole warps: one launch per step size, problems that have
            # accepted a candidate (or are finished) are skipped -- ~1.5 evaluations per problem instead of maxls
            for k in range(max(prm.maxls, 1)):
                self._cost(self.cand[:, :, k * P:(k + 1) * P], self.cand_cost[k * P:(k + 1) * P], which="cand",
                           mask=self.ls_active if k > 0 else None)
                L.check(lib.sde4_apg_ls_step(C.byref(prm), C.byref(st), _ptr(self.svals), _ptr(self.cand_cost), k,
                                             _ptr(self.ls_active), _stream()))
        else:
            n_eval = max(prm.maxls, 1) * P
            self._cost(self.cand[:, :, :n_eval], self.cand_cost[:n_eval], data_mod=P, which="cand")
        L.check(lib.sde4_apg_select(C.byref(prm), C.byref(st), _ptr(self.cand), _ptr(self.svals), _ptr(self.cand_cost),
                                    _ptr(self.xv), _ptr(self.sel_step), _ptr(self.sel_ls), _ptr(xmask),
                                    _ptr(self.xv_cost if reuse_x else None), _stream()))
        if self.merge_xv_grad:
            # latency-bound sizes: ONE value+gradient launch over the 2P problems x and v; pick copies point, cost
            # and gradient of the better one (one sequential rollout less per iteration, ~15 % more arithmetic)
            self._cost(self.xv, self.xv_cost, grad_buf=self.xv_grad, data_mod=P, which="xv")
            L.check(lib.sde4_apg_pick(C.byref(prm), _ptr(self.xv), _ptr(self.xv_cost), _ptr(self.y), _ptr(self.y_cost),
                                      _ptr(self.pick), _ptr(self.xv_grad), _ptr(self.y_grad), _stream()))
        else:
            # throughput-bound sizes: the reference's order -- value-only x / v, then the gradient at the chosen point
            self._cost(self.xv, self.xv_cost, data_mod=P, which="xv", mask=self.xvmask)  # x skipped where its cost is known
            L.check(lib.sde4_apg_pick(C.byref(prm), _ptr(self.xv), _ptr(self.xv_cost), _ptr(self.y), _ptr(self.y_cost),
                                      _ptr(self.pick), _ptr(None), _ptr(None), _stream()))
            self._cost(self.y, self.cost_tmp, grad_buf=self.y_grad)
        L.check(lib.sde4_apg_update(C.byref(prm), C.byref(st), _ptr(self.xv), _ptr(self.y), _ptr(self.y_cost),
                                    _ptr(self.y_grad), _ptr(self.pick), _ptr(self.sel_step), _ptr(self.sel_ls),
                                    _ptr(self.cand), _stream()))  # cand is dead here: |grad|^2 partials scratch

    def _step_speculative(self):
        """The same iteration with ONE cost launch: candidates -> every x_k = prox(cand_k), v_k and (with a prox) the raw
        candidates, value + gradient over Q*P problems -> Armijo selection on the raw costs -> gather the accepted
        step's x / v cost and gradient -> pick -> update."""
        lib, prm, st, P, K = L.load(), self.prm, self.st, self.P, self.K
        L.check(lib.sde4_apg_candidates(C.byref(prm), C.byref(st), _ptr(self.cand), _ptr(self.svals), _ptr(None), _stream()))
        L.check(lib.sde4_apg_spec_points(C.byref(prm), C.byref(st), _ptr(self.cand), _ptr(self.spec), _stream()))
        self._cost(self.spec, self.spec_cost, grad_buf=self.spec_grad, data_mod=P, which="xv")
        raw_cost = self.spec_cost[2 * K * P:] if self.lb is not None else self.spec_cost
        L.check(lib.sde4_apg_select(C.byref(prm), C.byref(st), _ptr(self.cand), _ptr(self.svals), _ptr(raw_cost),
                                    _ptr(self.xv), _ptr(self.sel_step), _ptr(self.sel_ls), _ptr(None), _ptr(None), _stream()))
        L.check(lib.sde4_apg_spec_gather(C.byref(prm), _ptr(self.sel_ls), _ptr(self.spec_cost), _ptr(self.spec_grad),
                                         _ptr(self.xv_cost), _ptr(self.xv_grad), _stream()))
        L.check(lib.sde4_apg_pick(C.byref(prm), _ptr(self.xv), _ptr(self.xv_cost), _ptr(self.y), _ptr(self.y_cost),
                                  _ptr(self.pick), _ptr(self.xv_grad), _ptr(self.y_grad), _stream()))
        L.check(lib.sde4_apg_update(C.byref(prm), C.byref(st), _ptr(self.xv), _ptr(self.y), _ptr(self.y_cost),
                                    _ptr(self.y_grad), _ptr(self.pick), _ptr(self.sel_step), _ptr(self.sel_ls),
                                    _ptr(self.cand), _stream()))

    def _capture(self):
        """Capture one masked iteration (6 bookkeeping + 2-3 cost + 2-3 reduce launches) in a CUDA graph.
        A throw-away eager iteration first makes every lazy set-up (workspace growth, func attributes)
        happen outside the capture; the solver state is restored afterwards."""
        names = L.APG_STATE_INT + L.APG_STATE_F32 + L.APG_STATE_VEC
        snap = {n_: self.t[n_].clone() for n_ in names}
        self.step()
        torch.cuda.synchronize()
        for n_ in names:
            self.t[n_].copy_(snap[n_])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step()
        for n_ in names:  # the capture itself does not execute, but keep the state explicit
            self.t[n_].copy_(snap[n_])
        self._graph = g

    def run(self, max_iter=None, check_every=0, use_graph=True):
        """``apg`` (apg.py:115-161): at most ``max_iter`` masked iterations.  ``check_every`` > 0 polls the
        not_done flags every that many iterations to exit early (one device->host sync each time).
        ``use_graph`` replays a captured CUDA graph per iteration (launch-bound otherwise for small P*N)."""
        max_iter = self.op["max_iter"] if max_iter is None else max_iter
        if use_graph and getattr(self, "_graph", None) is None:
            self._capture()
        for it in range(max_iter):
            if use_graph:
                self._graph.replay()
            else:
                self.step()
            if check_every and (it + 1) % check_every == 0 and not bool(self.t["not_done"].any()):
                break
        return self

    # ------------------------------------------------------------------
    def x_opt(self):
        """[P, H, n_opt] view of the best iterates."""
        return self.t["x_opt"].permute(2, 0, 1)

    def state_of(self, p):
        """``APGState`` of problem p (host scalars, device vectors) for parity checks / warm starts."""
        g = lambda n_: self.t[n_][p].item()
        v = lambda n_: self.t[n_][:, :, p]
        return APGState(num_iter_no_improv=int(g("num_iter_no_improv")), num_steps=int(g("num_steps")),
                        momentum=np.float32(g("momentum")), avg_momentum=np.float32(g("avg_momentum")),
                        avg_linesearch=np.float32(g("avg_linesearch")), avg_stepsize=np.float32(g("avg_stepsize")),
                        stepsize=np.float32(g("stepsize")), opt_cost=np.float32(g("opt_cost")), x_opt=v("x_opt"),
                        curr_cost=np.float32(g("curr_cost")), init_cost=np.float32(g("init_cost")),
                        grad_sqr=np.float32(g("grad_sqr")), not_done=bool(g("not_done")), xk=v("xk"), yk=v("yk"),
                        grad_yk=v("grad_yk"))
