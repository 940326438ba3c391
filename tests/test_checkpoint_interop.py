"""Checkpoint interchange with the reference, host logic on CPU.

  * FusedAdam's state dict is torch.optim.Adam's (either loads the other's and continues identically);
  * a file written by this package's save_model() is read by the UNMODIFIED reference's load_model() -- only where the
    reference tree exists (the build container); the opposite direction is pinned by the committed reference `.tar` files in
    tests/test_host_wiring.py::test_learner_resume_from_reference_checkpoint.
"""
import contextlib
import copy
import io

import pytest
import torch

import gan_lab_b200._growth as growth
from gan_lab_b200.optim import FusedAdam
from gan_lab_b200.utils.latent_utils import set_random_source
from oracle import kernel_contracts
from oracle.reference_loader import reference_available

import parity_cases as PC


@pytest.fixture(autouse=True)
def cpu_double(monkeypatch):
    kernel_contracts.install(monkeypatch)
    monkeypatch.setattr(growth, "FMAP_MAX", 32)
    yield
    set_random_source(None)


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.randn(8, 4, 3, 3, generator=g).contiguous(memory_format=torch.channels_last), torch.randn(8, generator=g),
          torch.randn(5, 7, generator=g)]
    return [torch.nn.Parameter(p) for p in ps]


def _step(opt, params, seed, skip_last=False):
    g = torch.Generator().manual_seed(seed)
    for i, p in enumerate(params):
        p.grad = None if (skip_last and i == len(params) - 1) else torch.randn(p.shape, generator=g)
    opt.step()


def test_fused_adam_state_dict_is_torch_adams():
    kw = dict(lr=2e-3, betas=(0., .99), eps=1e-8, weight_decay=0)
    pa, pb = _params(0), _params(0)
    mine, ref = FusedAdam(pa, **kw), torch.optim.Adam(pb, **kw)
    _step(mine, pa, 1, skip_last=True); _step(ref, pb, 1, skip_last=True)     # the last parameter joins one step late
    for s in (2, 3):
        _step(mine, pa, s); _step(ref, pb, s)
    sd_m, sd_r = mine.state_dict(), ref.state_dict()
    assert sd_m["state"].keys() == sd_r["state"].keys()
    for k, st in sd_r["state"].items():
        assert float(sd_m["state"][k]["step"]) == float(st["step"]), k
        for key in ("exp_avg", "exp_avg_sq"):
            torch.testing.assert_close(sd_m["state"][k][key], st[key], rtol=1e-6, atol=1e-9)
    assert [float(sd_m["state"][k]["step"]) for k in sorted(sd_m["state"])] == [3.0, 3.0, 2.0]
    for key in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize"):
        assert sd_m["param_groups"][0][key] == sd_r["param_groups"][0][key], key

    # cross-load: torch's state into a fresh FusedAdam, FusedAdam's state into a fresh torch Adam; both carry on alike
    pc, pd = [torch.nn.Parameter(p.detach().clone()) for p in pb], [torch.nn.Parameter(p.detach().clone()) for p in pa]
    mine2, ref2 = FusedAdam(pc, **kw), torch.optim.Adam(pd, **kw)
    mine2.load_state_dict(copy.deepcopy(sd_r)); ref2.load_state_dict(copy.deepcopy(sd_m))   # (torch hands out live state)
    for p, q in zip(pc, mine2.param_groups[0]["params"]):
        assert mine2.state[q]["exp_avg"].stride() == p.stride()
    for s in (4, 5):
        _step(mine2, pc, s); _step(ref2, pd, s); _step(mine, pa, s); _step(ref, pb, s)
    for a, b, c, d in zip(pa, pb, pc, pd):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(c, b, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(d, b, rtol=1e-6, atol=1e-7)


@pytest.mark.skipif(not reference_available(), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("fname,model", PC.GROW_CASES + PC.GROW_TAPER_CASES)
def test_reference_loads_our_checkpoint(golden, fname, model, tmp_path):
    """This package trains through a resolution increase and saves mid-fade-in; the unmodified reference's load_model()
    rebuilds its own networks from the file: same parameters, same phase, and its train() runs on from there."""
    import numpy as np
    from oracle import make_golden as MG
    from oracle.reference_loader import load_reference, make_config
    g = golden(fname)
    L = PC._grow_learner(g, "cpu", model)
    PC._load(L.gen_model, g["g_sd0"]); PC._load(L.disc_model, g["d_sd0"]); PC._load(L.gen_model_lagged, g["g_sd0"])
    L.train(PC.ReplayLoader(g["served"], L.batch_size, "cpu"), num_main_iters=6)
    path = tmp_path / "ours.tar"
    L.save_model(path)

    ref = load_reference()
    MG._patch_small(ref, fmap_base=g.get("fmap_base", 8192))
    orig_load = torch.load
    torch.load = lambda *a, **k: orig_load(*a, **{"weights_only": False, **k})     # torch >= 2.6 default; see make_golden
    try:
        over = dict(bs_dict=dict(g["bs_dict"]), nimg_transition=g["nimg_transition"], res_dataset=g["data_res"],
                    lr_fctr_dict=dict(g["lr_fctr_dict"]))
        with contextlib.redirect_stdout(io.StringIO()):
            if model == "StyleGAN":
                R, _ = MG._build_style_learner(ref, g["res"], g["init_res"], 4, **over)
            else:
                R = ref.progan_learner.ProGANLearner(make_config("ProGAN", res=g["res"], init_res=g["init_res"], batch_size=4,
                                                                 len_latent=g["len_latent"], **over))
            R.load_model(path, dev_of_saved_model="cpu")
        G = L.gen_model
        assert (R.gen_model.curr_res, R.gen_model.fade_in_phase, R.gen_model.alpha) == (G.curr_res, G.fade_in_phase, G.alpha)
        assert (R.curr_img_num, R.curr_phase_num, R.batch_size, R.loss, R.gradient_penalty) == \
               (L.curr_img_num, L.curr_phase_num, L.batch_size, L.loss, L.gradient_penalty)
        for mine, theirs in ((L.gen_model, R.gen_model), (L.disc_model, R.disc_model), (L.gen_model_lagged, R.gen_model_lagged)):
            sm, st = mine.state_dict(), theirs.state_dict()
            assert list(sm.keys()) == list(st.keys())
            for k in sm:
                assert torch.equal(sm[k], st[k]), k
        assert type(R.config).__name__ == "LearnerConfigCopy" and type(R.config).__module__ == "_int"
        assert type(R.lagged_params).__module__ == "indexed"
        assert isinstance(R.nl, torch.nn.LeakyReLU) and type(R.nl).__module__.startswith("torch.")
        if model == "StyleGAN":
            assert torch.equal(R.gen_model.w_ewma, G.w_ewma)
        # ... and the reference trains on from it
        images = torch.randint(0, 256, (8, 8, 8, 3), dtype=torch.uint8).numpy()
        ds = MG.PILBoxDataset(images, 4)
        from torch.utils.data import BatchSampler, DataLoader, SequentialSampler
        dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=4, drop_last=True))
        torch.manual_seed(0); np.random.seed(0)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            R.train(dl, num_main_iters=2)
        assert R.curr_img_num == L.curr_img_num + 8
        assert all(torch.isfinite(p).all() for p in R.gen_model.parameters())
    finally:
        torch.load = orig_load
        MG._unpatch(ref)


def test_resnet_checkpoint_roundtrip_and_reference_load(golden, tmp_path):
    """GANLearner (ResNet GAN): save_model() -> load_model() keeps networks (incl. BatchNorm buffers), Adam state and
    bookkeeping, so the resumed learner's next iteration equals the uninterrupted one; where the reference tree exists, its
    own load_model() reads the same file."""
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    from gan_lab_b200.utils.latent_utils import TapeSource
    g = golden("resnet_train_res64.pt")
    bs = g["bs"]

    def run(L, n, tape):
        ds = TensorDataset(g["data"])
        dl = DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=bs, drop_last=True))
        set_random_source(TapeSource(tape, "cpu"))
        if not L.not_trained_yet:
            L.train_dataiter = iter(dl)         # a resumed learner restarts its loader (reference :497-498); do the same here
        L.train(dl, num_main_iters=n)

    L, _ = PC._resnet_learner(g, "cpu", num_disc_iters=g["num_disc_iters"], lr_base=g["lr"])
    PC._load(L.gen_model, g["g_sd0"]); PC._load(L.disc_model, g["d_sd0"])
    run(L, 1, g["tape"])
    path = tmp_path / "resnet.tar"
    L.save_model(path)
    L2, _ = PC._resnet_learner(g, "cpu")
    L2.load_model(path, dev_of_saved_model="cpu")
    for a, b in ((L.gen_model, L2.gen_model), (L.disc_model, L2.disc_model)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
    sd1, sd2 = L.opt_disc.state_dict(), L2.opt_disc.state_dict()
    assert sd1["state"].keys() == sd2["state"].keys() and sd1["state"]
    for k, st in sd1["state"].items():
        assert float(st["step"]) == float(sd2["state"][k]["step"]) == float(g["num_disc_iters"])
        assert torch.equal(st["exp_avg_sq"], sd2["state"][k]["exp_avg_sq"])
    assert (L2.config.num_disc_iters, L2.curr_img_num, L2.batch_size, L2.pretrained_model) == \
           (g["num_disc_iters"], L.curr_img_num, bs, True)
    # both carry on from the same point with the same draws: identical parameters afterwards
    run(L, 1, g["tape"]); run(L2, 1, g["tape"])
    for (k, a), (_, b) in zip(L.gen_model.state_dict().items(), L2.gen_model.state_dict().items()):
        torch.testing.assert_close(a, b, rtol=0, atol=0, msg=k)
    for (k, a), (_, b) in zip(L.disc_model.state_dict().items(), L2.disc_model.state_dict().items()):
        torch.testing.assert_close(a, b, rtol=0, atol=0, msg=k)

    if not reference_available():
        return
    from oracle import make_golden as MG
    from oracle.reference_loader import load_reference
    ref = load_reference()
    orig_load = torch.load
    torch.load = lambda *a, **k: orig_load(*a, **{"weights_only": False, **k})
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            R, _ = MG._resnet_learner(ref, g["res"], bs)
            R.load_model(path, dev_of_saved_model="cpu")
        assert type(R.gen_model).__module__ == "resnetgan.architectures"
        L3, _ = PC._resnet_learner(g, "cpu")
        L3.load_model(path, dev_of_saved_model="cpu")
        for mine, theirs in ((L3.gen_model, R.gen_model), (L3.disc_model, R.disc_model)):
            sm, st = mine.state_dict(), theirs.state_dict()
            assert list(sm.keys()) == list(st.keys())
            for k in sm:
                assert torch.equal(sm[k], st[k]), k
        assert float(next(iter(R.opt_disc.state_dict()["state"].values()))["step"]) == float(g["num_disc_iters"])
    finally:
        torch.load = orig_load
