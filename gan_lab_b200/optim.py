"""Fused multi-tensor Adam with the EWMA-generator update in the same pass.

`FusedAdam` is a `torch.optim.Optimizer` (so `LambdaLR` and the learners' `opt_gen` / `opt_disc` surface work
unchanged) with torch.optim.Adam's update rule (reference utils/backprop_utils.py:109-120), executed by ONE
kernel launch over all parameters (`glb_adam_ewma_multi`) instead of the foreach kernel train; when a dict of
lagged ("EWMA generator") tensors is attached, `lagged = p*(1-beta) + lagged*beta` (reference
progan/learner.py:909-916) rides along in the same pass over the parameters.
"""
import torch

from . import _kernels as K


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super(FusedAdam, self).__init__(params, defaults)
        self._ewma = None          # (named lagged tensors aligned with extra params, beta)
        self._ewma_started = False
        self._hyper_host = None
        self._hyper_dev = None

    def attach_ewma(self, named_params, lagged: dict, beta: float):
        """named_params: iterable of (name, param) whose lagged copies live in `lagged[name]` (same shapes/layout)."""
        self._ewma = (list(named_params), lagged, float(beta))

    def _tables(self, group):
        """-> {step: (rows, sizes)}; rows = (p, g, m, v, lagged) pointers.  torch.optim.Adam keeps a step count
        per parameter (a parameter that had no gradient in some step lags behind), so rows are grouped by it."""
        by_step = {}
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
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['step'] += 1
            g = p.grad
            if g.stride() != p.stride():
                g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                p.grad = g
            lag = lag_of.get(id(p))
            rows, sizes = by_step.setdefault(st['step'], ([], []))
            rows.append((p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                         lag.data_ptr() if lag is not None else 0))
            sizes.append(p.numel())
            in_opt.add(id(p))
        if self._ewma is not None:
            named, lagged, _ = self._ewma
            extra = [(n, p) for n, p in named if id(p) not in in_opt]
            if extra:
                key = next(iter(by_step)) if by_step else 1
                rows, sizes = by_step.setdefault(key, ([], []))
                for n, p in extra:
                    rows.append((p.data_ptr(), 0, 0, 0, lagged[n].data_ptr()))
                    sizes.append(p.numel())
        return by_step

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        for group in self.param_groups:
            by_step = self._tables(group)
            if not by_step:
                continue
            dev = group['params'][0].device
            b1, b2 = group['betas']
            mode, beta = 0, 0.0
            if self._ewma is not None:
                beta = self._ewma[2]
                mode = 1 if self._ewma_started else 2
            for step, (rows, sizes) in by_step.items():
                table = torch.tensor(rows, dtype=torch.int64).to(dev, non_blocking=True)
                szs = torch.tensor(sizes, dtype=torch.int64).to(dev, non_blocking=True)
                hyper = torch.tensor([group['lr'], 1.0 - b1 ** step, 1.0 - b2 ** step, 0.0],
                                     dtype=torch.float32).to(dev, non_blocking=True)
                K.adam_ewma_multi(table, szs, len(rows), max(sizes), hyper, b1, b2, group['eps'],
                                  group['weight_decay'], beta, mode)
        if self._ewma is not None:
            self._ewma_started = True
        K.weights_updated()          # parameters were rewritten through raw pointers: drop cached weight re-layouts
        return None
