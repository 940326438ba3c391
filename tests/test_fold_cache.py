"""Host logic of the folded-convolution operand cache (gan_lab_b200/_kernels.py::_folded_weights), on the CPU with the C-ABI
call recorded instead of executed: which operands a call asks the library to write, what a later call finds, and that an
optimiser step drops everything (the parameters are rewritten through raw pointers, no torch version bump)."""
import torch

import gan_lab_b200._kernels as K


def _patched(monkeypatch):
    calls = []
    monkeypatch.setattr(K, "_call", lambda name, *args: calls.append((name, args)))
    monkeypatch.setattr(K, "_stream", lambda: 0)
    K.weights_updated()
    return calls


def test_forward_only_then_backward_on_demand(monkeypatch):
    calls = _patched(monkeypatch)
    w = torch.randn(8, 4, 3, 3).contiguous(memory_format=torch.channels_last)
    fwd, bwd = K.upconv_weights(w, "fwd")                      # a no_grad forward: only the forward operand
    assert fwd is not None and bwd is None and tuple(fwd.shape) == (32, 2, 2, 4)
    name, args = calls[-1]
    assert name == "glb_upconv_weights" and args[1] == fwd.data_ptr() and args[2] is None and args[3:5] == (8, 4)
    f2, b2 = K.upconv_weights(w, "fwd")
    assert f2 is fwd and len(calls) == 1                       # cached
    f3, b3 = K.upconv_weights(w, "bwd")                        # the data gradient asks for the missing one only
    assert f3 is fwd and b3 is not None and tuple(b3.shape) == (4, 16, 8)
    name, args = calls[-1]
    assert len(calls) == 2 and args[1] is None and args[2] == b3.data_ptr()
    assert K.upconv_weights(w, "bwd")[1] is b3 and len(calls) == 2


def test_both_operands_in_one_call_and_downconv_argument_order(monkeypatch):
    calls = _patched(monkeypatch)
    w = torch.randn(8, 4, 3, 3).contiguous(memory_format=torch.channels_last)
    fwd, bwd = K.downconv_weights(w, "fwd", both=True)         # the forward of a layer whose input requires grad
    assert tuple(fwd.shape) == (8, 16, 4) and tuple(bwd.shape) == (16, 2, 2, 8)      # wt [Co,16,Ci], wp [4*Ci,2,2,Co]
    name, args = calls[-1]
    assert name == "glb_downconv_weights" and len(calls) == 1
    assert args[1] == bwd.data_ptr() and args[2] == fwd.data_ptr()                  # C signature: (w, wp, wt, Co, Ci, stream)
    assert K.downconv_weights(w, "bwd")[1] is bwd and len(calls) == 1
    up_f, up_b = K.upconv_weights(w, "fwd", both=True)         # same weight tensor, other role: its own entry
    assert len(calls) == 2 and tuple(up_f.shape) == (32, 2, 2, 4) and tuple(up_b.shape) == (4, 16, 8)


def test_optimiser_step_drops_the_operands(monkeypatch):
    calls = _patched(monkeypatch)
    w = torch.randn(8, 4, 3, 3).contiguous(memory_format=torch.channels_last)
    K.upconv_weights(w, "fwd", both=True)
    K.weights_updated()                                        # what FusedAdam.step() and the learners call after graph replays
    K.upconv_weights(w, "fwd", both=True)
    assert len(calls) == 2
    w.add_(1.0)                                                # an in-place torch update bumps the version: stale entry ignored
    K.upconv_weights(w, "fwd", both=True)
    assert len(calls) == 3


def test_data_parallel_packs_small_gradients_into_one_message(monkeypatch):
    """parallel.DataParallel, in-place NCCL mode, with the collectives replaced by a recorder that halves its argument (an
    average with an all-zero peer): gradients below `pack_below` elements travel as ONE flat message and are copied back,
    the others are reduced where they are -- every gradient ends up halved, with one collective per large tensor plus one."""
    import contextlib
    import torch.distributed as dist
    from gan_lab_b200.parallel import DataParallel
    seen = []

    class _Work(object):
        def wait(self):
            return True

    @contextlib.contextmanager
    def fake_group(device=None, async_ops=False):
        yield _Work()

    def fake_all_reduce(t, op=None, async_op=False):
        seen.append(t.numel())
        t.mul_(0.5)
        return _Work()

    monkeypatch.setattr(dist, "_coalescing_manager", fake_group)
    monkeypatch.setattr(dist, "all_reduce", fake_all_reduce)
    dp = DataParallel(world_size=2, bucket_bytes=1 << 30, inplace=True)
    dp.pack_below = 100
    m = torch.nn.ModuleList([torch.nn.Linear(4, 3), torch.nn.Linear(30, 20), torch.nn.Linear(3, 2)])     # 12+3, 600+20, 6+2 elements
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    dp.allreduce_grads(m)
    assert all(torch.equal(p.grad, torch.full_like(p, 0.5)) for p in m.parameters())
    assert sorted(seen) == [12 + 3 + 20 + 6 + 2, 600]           # one packed message for the five small tensors, one in place
