"""Fused multi-tensor Adam with the EWMA-generator update in the same pass.

`FusedAdam` is a `torch.optim.Optimizer` (so `LambdaLR` and the learners' `opt_gen` / `opt_disc` surface work
unchanged) with torch.optim.Adam's update rule (reference utils/backprop_utils.py:109-120), executed by ONE
kernel launch over all parameters (`glb_adam_ewma_multi`) instead of the foreach kernel train; when a dict of
lagged ("EWMA generator") tensors is attached, `lagged = p*(1-beta) + lagged*beta` (reference
progan/learner.py:909-916) rides along in the same pass over the parameters.
"""
import torch

from . import _kernels as K


def _capturing():
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class FusedAdam(torch.optim.Optimizer):
    """CUDA-graph safe: the bias corrections live in a device-side `hyper` vector advanced by a one-thread kernel, the
    pointer tables are staged through pinned memory that stays untouched after a capture (`prepare_capture()`), and
    nothing in `step()` synchronises the host."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super(FusedAdam, self).__init__(params, defaults)
        self._ewma = None          # (named lagged tensors aligned with extra params, beta)
        self._ewma_started = False
        self._nsteps = 0
        self._hyper = {}           # (group index, cohort) -> {'t': device float[4] = lr, 1-b1^t, 1-b2^t, t ; 'lr': value on the device}
        self._tab = {}             # (group index, cohort) -> cached device tables of the eager path
        self._capture_slots = None

    def attach_ewma(self, named_params, lagged: dict, beta: float):
        """named_params: iterable of (name, param) whose lagged copies live in `lagged[name]` (same shapes/layout)."""
        self._ewma = (list(named_params), lagged, float(beta))

    def _tables(self, group, gi=0):
        """-> {cohort: (rows, sizes)}; rows = (p, g, m, v, lagged) pointers.  torch.optim.Adam keeps a step count per
        parameter (a parameter first seen later lags behind), so rows are grouped by the step they joined at."""
        by = {}
        in_opt = set()
        lag_of = {}
        if self._ewma is not None:
            named, lagged, _ = self._ewma
            lag_of = {id(p): lagged[n] for n, p in named}
        for p in group['params']:
            if p.grad is None:
                continue
            st = self.state[p]
            if not st:
                st['step'] = 0
                st['cohort'] = self._nsteps
                st['group'] = gi
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['step'] += 1
            g = p.grad
            if g.stride() != p.stride():
                g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                p.grad = g
            lag = lag_of.get(id(p))
            rows, sizes = by.setdefault(st['cohort'], ([], []))
            rows.append((p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                         lag.data_ptr() if lag is not None else 0))
            sizes.append(p.numel())
            in_opt.add(id(p))
        if self._ewma is not None:
            named, lagged, _ = self._ewma
            extra = [(n, p) for n, p in named if id(p) not in in_opt]
            if extra:
                key = next(iter(by)) if by else self._nsteps
                rows, sizes = by.setdefault(key, ([], []))
                for n, p in extra:
                    rows.append((p.data_ptr(), 0, 0, 0, lagged[n].data_ptr()))
                    sizes.append(p.numel())
        return by

    # -- device-side state ---------------------------------------------------------------------------------------
    def _hyper_for(self, gi, cohort, dev, lr):
        """One (lr, bias corrections, step count) vector per (param group, cohort): groups may differ in lr / betas and
        every vector is advanced exactly once per step()."""
        h = self._hyper.get((gi, cohort))
        if h is None:
            if _capturing():
                raise RuntimeError("FusedAdam: a new (group, cohort) appeared inside a CUDA-graph capture; run one eager "
                                   "step() with these parameters first")
            h = {'t': torch.zeros(4, dtype=torch.float32, device=dev), 'lr': None}
            self._hyper[(gi, cohort)] = h
        if h['lr'] != lr:
            if _capturing():
                raise RuntimeError("FusedAdam: the learning rate changed inside a CUDA-graph capture; call push_lr() first")
            h['t'][0:1].fill_(lr)
            h['lr'] = lr
        return h['t']

    def push_lr(self):
        """Write a changed learning rate (LambdaLR edits param_groups on the host) to the device copies; called by
        the learner before replaying a captured step."""
        for gi, group in enumerate(self.param_groups):
            for (hg, _cohort), h in self._hyper.items():
                if hg == gi and h['lr'] != group['lr']:
                    h['t'][0:1].fill_(group['lr'])
                    h['lr'] = group['lr']

    def prepare_capture(self):
        """Allocate the device tables a captured step() will point its kernel at.  Their CONTENT (pointers known on the
        host while capturing) is uploaded by finish_capture() after the capture has ended: a host-to-device copy issued
        inside a capture would leave capture-mode events in torch's pinned-memory allocator."""
        n = sum(len(g['params']) for g in self.param_groups) + (len(self._ewma[0]) if self._ewma else 0) + 8
        dev = self.param_groups[0]['params'][0].device
        cohorts = {st.get('cohort') for st in self.state.values() if st}
        slots = max(1, len(self.param_groups)) * (max(1, len(cohorts)) + 1)     # one table per (group, cohort) + one spare
        self._capture_slots = [torch.empty(6 * n, dtype=torch.int64, device=dev) for _ in range(slots)]
        self._pending_uploads = []

    def finish_capture(self):
        for devbuf, flat in getattr(self, '_pending_uploads', []):
            devbuf[:len(flat)].copy_(torch.tensor(flat, dtype=torch.int64))
        self._pending_uploads = []

    def _upload(self, key, rows, sizes, dev):
        n = len(rows)
        flat = [v for r in rows for v in r] + list(sizes)
        if _capturing():
            if not self._capture_slots:
                raise RuntimeError("FusedAdam.step() under CUDA-graph capture needs prepare_capture() first")
            devbuf = self._capture_slots.pop()
            self._pending_uploads.append((devbuf, flat))
            return devbuf[:5 * n], devbuf[5 * n:6 * n]
        ent = self._tab.get(key)
        if ent is not None and ent[0] == flat:
            return ent[1], ent[2]
        devbuf = torch.tensor(flat, dtype=torch.int64).to(dev, non_blocking=True)
        self._tab[key] = (flat, devbuf[:5 * n], devbuf[5 * n:6 * n])
        return devbuf[:5 * n], devbuf[5 * n:6 * n]

    # -- torch.optim.Adam-compatible state dicts (checkpoints) -----------------------------
    _TORCH_ADAM_GROUP_DEFAULTS = dict(amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False,
                                      fused=None, decoupled_weight_decay=False)

    def _true_step(self, st):
        """Steps taken by the parameter behind state `st`: the device-side count of its cohort when it exists (under
        CUDA-graph replay the host-side counter does not move), else the host-side count."""
        h = self._hyper.get((st.get('group', 0), st.get('cohort')))
        if h is not None:
            return int(round(float(h['t'][3])))       # device -> host read; checkpoints are not on the hot path
        return int(st.get('step', 0))

    def state_dict(self):
        """The layout torch.optim.Adam writes and reads (reference progan/learner.py:1271-1272, 1396-1397): per parameter
        `step` (a float32 scalar tensor), `exp_avg`, `exp_avg_sq`; the moments are handed out contiguous (they live in the
        parameter's channels_last memory here)."""
        sd = super(FusedAdam, self).state_dict()
        steps = {id(st): self._true_step(st) for st in self.state.values()}
        by_index = {}
        index = 0
        for group in self.param_groups:
            for p in group['params']:
                if p in self.state and self.state[p]:
                    by_index[index] = steps[id(self.state[p])]
                index += 1
        state = {}
        for idx, st in sd['state'].items():
            state[idx] = {'step': torch.tensor(float(by_index[idx]), dtype=torch.float32),
                          'exp_avg': st['exp_avg'].detach().clone(memory_format=torch.contiguous_format),
                          'exp_avg_sq': st['exp_avg_sq'].detach().clone(memory_format=torch.contiguous_format)}
        groups = [{**self._TORCH_ADAM_GROUP_DEFAULTS, **g} for g in sd['param_groups']]
        return {'state': state, 'param_groups': groups}

    def load_state_dict(self, state_dict):
        """Accepts torch.optim.Adam's state dict (a reference checkpoint) or this class's own."""
        super(FusedAdam, self).load_state_dict(state_dict)
        self._hyper, self._tab, self._capture_slots = {}, {}, None
        steps = [int(round(float(st['step']))) for st in self.state.values() if st]
        self._nsteps = max(steps) if steps else 0
        for p, st in self.state.items():
            if not st:
                continue
            t = int(round(float(st['step'])))
            st['step'] = t
            st['cohort'] = self._nsteps - t            # parameters that joined later have taken fewer steps
            for key in ('exp_avg', 'exp_avg_sq'):      # the kernel walks p, g, m, v with one linear index
                if st[key].stride() != p.stride() or st[key].device != p.device or st[key].dtype != p.dtype:
                    st[key] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(st[key])
            gi = next(i for i, g in enumerate(self.param_groups) if any(q is p for q in g['params']))
            st['group'] = gi
            if (gi, st['cohort']) not in self._hyper:
                b1, b2 = self.param_groups[gi]['betas']
                self._hyper[(gi, st['cohort'])] = {
                    't': torch.tensor([0.0, 1.0 - b1 ** t, 1.0 - b2 ** t, float(t)], dtype=torch.float32, device=p.device),
                    'lr': None}
        for st in self.state.values():
            st.pop('amsgrad', None)

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        for gi, group in enumerate(self.param_groups):
            by = self._tables(group, gi)
            if not by:
                continue
            dev = group['params'][0].device
            b1, b2 = group['betas']
            mode, beta = 0, 0.0
            if self._ewma is not None:
                beta = self._ewma[2]
                mode = 1 if self._ewma_started else 2
            for cohort, (rows, sizes) in by.items():
                hyper = self._hyper_for(gi, cohort, dev, group['lr'])
                K.adam_hyper_advance(hyper, b1, b2)
                table, szs = self._upload((gi, cohort), rows, sizes, dev)
                K.adam_ewma_multi(table, szs, len(rows), max(sizes), hyper, b1, b2, group['eps'],
                                  group['weight_decay'], beta, mode)
        self._nsteps += 1
        if self._ewma is not None:
            self._ewma_started = True
        K.weights_updated()          # parameters were rewritten through raw pointers: drop cached weight re-layouts
        return None
