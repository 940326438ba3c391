"""Host-side logic above the C-ABI, checked on CPU (no GPU in the build container).

The CUDA launchers in gan_lab_b200._kernels are swapped for their CPU contracts (oracle/kernel_contracts.py)
-- a TEST DOUBLE, installed only here -- so that what is exercised is the product's own Python: the autograd
Functions (incl. the closed {fprop, dgrad, wgrad} double backward that R1 / WGAN-GP needs), the drop-in modules and
their state_dict layout, the fused architecture forwards, the learners' loop, FusedAdam's pointer tables and the
EWMA bookkeeping.  Expected values are the fixtures produced by the UNMODIFIED reference (tests/golden).
The same case bodies run against the real kernels in tests/test_gpu_parity.py.
"""
import pytest

import gan_lab_b200._growth as growth
from gan_lab_b200.utils.latent_utils import set_random_source
from oracle import kernel_contracts

import parity_cases as PC

DEV = "cpu"


@pytest.fixture(autouse=True)
def cpu_double(monkeypatch):
    kernel_contracts.install(monkeypatch)
    monkeypatch.setattr(growth, "FMAP_MAX", 32)
    yield
    set_random_source(None)


@pytest.mark.parametrize("name", PC.CONV_CASES)
def test_conv2d_ex_module(golden, name):
    PC.case_conv2d_ex_module(golden, DEV, name)


@pytest.mark.parametrize("name", PC.LINEAR_CASES)
def test_linear_ex_module(golden, name):
    PC.case_linear_ex_module(golden, DEV, name)


def test_small_modules(golden):
    PC.case_small_modules(golden, DEV)


@pytest.mark.parametrize("name", PC.MBSTD_CASES)
def test_mbstd_module(golden, name):
    PC.case_mbstd_module(golden, DEV, name)


def test_style_epilogue_op(golden):
    PC.case_style_epilogue_op(golden, DEV)


@pytest.mark.parametrize("fname", PC.STYLE_NETS + PC.STYLE_NETS_TAPER)
def test_style_nets_modules(golden, fname):
    PC.case_style_nets_modules(golden, DEV, fname)


@pytest.mark.parametrize("fname", PC.PRO_NETS + PC.PRO_NETS_TAPER)
def test_pro_nets_modules(golden, fname):
    PC.case_pro_nets_modules(golden, DEV, fname)


@pytest.mark.parametrize("fname,model", PC.TRAIN_CASES)
def test_learner_train(golden, fname, model):
    PC.case_learner_train(golden, DEV, fname, model)


@pytest.mark.parametrize("fname,model", PC.GROW_CASES + PC.GROW_TAPER_CASES)
def test_learner_grow(golden, fname, model):
    PC.case_learner_grow(golden, DEV, fname, model)


@pytest.mark.parametrize("fname,model", PC.GROW_CASES + PC.GROW_TAPER_CASES)
def test_learner_grow_device_alpha(golden, fname, model):
    PC.case_learner_grow(golden, DEV, fname, model, device_alpha=True)


@pytest.mark.parametrize("fname,model", PC.RESUME_CASES)
def test_learner_resume_from_reference_checkpoint(golden, fname, model):
    from conftest import GOLDEN
    PC.case_learner_resume(golden, DEV, fname, model, GOLDEN)


@pytest.mark.parametrize("fname,model", PC.GROW_CASES + PC.GROW_TAPER_CASES)
def test_checkpoint_roundtrip(golden, fname, model, tmp_path):
    PC.case_checkpoint_roundtrip(golden, DEV, fname, model, tmp_path)


@pytest.mark.parametrize("fname,model", PC.METRICS_CASES)
def test_compute_metrics(golden, fname, model, tmp_path):
    PC.case_compute_metrics(golden, DEV, fname, model, tmp_path)


@pytest.mark.parametrize("name", PC.TRAIN_VARIANTS)
def test_train_variant(golden, name, monkeypatch):
    PC.case_train_variant(golden, DEV, name, monkeypatch)


@pytest.mark.parametrize("gp", ["r1", "r2"])
def test_shared_penalty_forward(gp):
    PC.case_shared_penalty_forward(DEV, gp)


@pytest.mark.parametrize("fname", PC.RESNET_NETS)
def test_resnet_nets_modules(golden, fname):
    PC.case_resnet_nets_modules(golden, DEV, fname)


@pytest.mark.parametrize("fname", PC.RESNET_NETS)
def test_resnet_nets_modules_fp64(golden, fname, monkeypatch):
    """Same case with the CPU doubles in fp64: no ReLU mask bit can flip, so every gradient must agree with the reference
    down to the fixture's own fp32 noise (this is what pins the host wiring of the ResNet path exactly)."""
    import torch
    import gan_lab_b200._kernels as K
    monkeypatch.setattr(K, "_chk", lambda *ts: None)
    PC.case_resnet_nets_modules(golden, DEV, fname, dtype=torch.float64, grad_tol=2e-5)


def test_resnet_train(golden):
    PC.case_resnet_train(golden, DEV)


def test_resnet_train_variant(golden):
    PC.case_resnet_train(golden, DEV, "resnet_train_res32_variant.pt")


def test_style_generator_eval_mode(golden):
    PC.case_style_eval(golden, DEV)


def test_train_runs_validation_metrics_where_the_reference_does(capsys):
    """train() with validation loaders: discriminator metrics after the D step and generator metrics after the G step of the
    first and every num_iters_valid-th iteration (reference progan/learner.py:820-832, 918-930)."""
    import torch
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    from gan_lab_b200.config import default_config
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    torch.manual_seed(0)
    cfg = default_config("StyleGAN", res=8, batch_size=4, dev=DEV, len_latent=32, len_dlatent=32, cutoff_trunc_trick=1,
                         gen_metrics=["generator loss", "fake realness"], disc_metrics=["discriminator loss"], num_iters_valid=3)
    L = StyleGANLearner(cfg)

    def loader(t):
        ds = TensorDataset(t)
        return DataLoader(ds, batch_sampler=BatchSampler(SequentialSampler(ds), batch_size=4, drop_last=True))

    calls = []
    orig = L.compute_metrics
    L.compute_metrics = lambda **kw: calls.append((kw["metrics_type"], L.curr_img_num)) or orig(**kw)
    L.train(loader(torch.rand(16, 3, 8, 8) * 2 - 1), valid_dl=loader(torch.rand(8, 3, 8, 8) * 2 - 1),
            z_valid_dl=loader(torch.randn(8, 32)), num_main_iters=4)
    assert calls == [("Discriminator", 0), ("Generator", 4), ("Discriminator", 8), ("Generator", 12)]
    out = capsys.readouterr().out
    assert out.count("Discriminator Validation Metrics:") == 2 and out.count("generator loss:") == 2
    assert (L.gen_metrics_num, L.disc_metrics_num) == (2, 2)


def test_batch_size_changes_with_resolution_through_the_device_loader():
    """bs_dict gives 8 samples per batch at 4x4 and 4 at 8x8: at the resolution increase train() re-sizes the loader's batches,
    asks it for the new resolution, recomputes the EWMA beta and the transition length, and counts images accordingly.  (The
    reference itself cannot run this under torch 2.11 -- BatchSampler caches its size -- so there is no fixture)"""
    import torch
    from gan_lab_b200.config import default_config
    from gan_lab_b200.data import DeviceImageLoader
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    torch.manual_seed(0)
    bs_dict = {r: 4 for r in (16, 32, 64, 128, 256, 512, 1024)}
    bs_dict.update({4: 8, 8: 4})
    cfg = default_config("StyleGAN", res=8, init_res=4, batch_size=8, dev=DEV, len_latent=32, len_dlatent=32, cutoff_trunc_trick=1,
                         bs_dict=bs_dict, nimg_transition=12)          # not a multiple of 8 -> rounded up to 16 at 4x4, 12 at 8x8
    L = StyleGANLearner(cfg)
    images = torch.randint(0, 256, (32, 8, 8, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
    dl = DeviceImageLoader(images, batch_size=L.batch_size, res=4, shuffle=False, device=DEV)
    seen = []
    L.train(dl, num_main_iters=7, step_callback=lambda i, ld, lg: seen.append(
        (L.gen_model.curr_res, L.batch_size, dl.res, dl.batch_sampler.batch_size, L.curr_img_num, round(float(L.gen_model.alpha), 6))))
    assert seen[0][:4] == (4, 8, 4, 8) and seen[1][:4] == (4, 8, 4, 8)
    assert [s[4] for s in seen] == [8, 16, 20, 24, 28, 32, 36]            # 2 x 8 images, then 4 per iteration
    assert all(s[:4] == (8, 4, 8, 4) for s in seen[2:])
    assert [s[5] for s in seen[2:5]] == [0.0, 0.5, 1.0]                  # delta_alpha = 4 / (12 - 4)
    assert L.nimg_transition_lst[:3] == [16, 12, float("inf")] or L.nimg_transition_lst[:2] == [16, 12]
    assert abs(L.beta - 0.5 ** (4 / 10000.)) < 1e-15


@pytest.mark.parametrize("gp,bs", [("r1", 4), ("r2", 8), ("r1", 2)])
def test_batched_d_passes(gp, bs):
    PC.case_batched_d_passes(DEV, gp, bs)


def test_resnet_resume_from_reference_checkpoint(golden):
    from conftest import GOLDEN
    PC.case_resnet_resume(golden, DEV, GOLDEN)


def test_full_size_state_dict_names_and_shapes(golden, monkeypatch):
    """The drop-in networks at their REAL sizes (512-channel plan, 4x4 ... 1024x1024, cfg3's mid-fade-in 256x256, both ResNets)
    carry exactly the reference's parameter / buffer names, order and logical shapes -- what lets its checkpoints load."""
    import math
    from gan_lab_b200.config import default_config
    from gan_lab_b200.progan.learner import ProGANLearner
    from gan_lab_b200.resnetgan.learner import GANLearner
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    monkeypatch.setattr(growth, "FMAP_MAX", 512)
    shapes = lambda m: [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    table = golden("state_dict_shapes.pt")
    assert len(table) == 10
    for (model, res, init_res), want in table.items():
        if model == "ResNet GAN":
            L = GANLearner(default_config("ResNet GAN", res=res, batch_size=8, dev=DEV))
        else:
            cls = StyleGANLearner if model == "StyleGAN" else ProGANLearner
            L = cls(default_config(model, res=res, init_res=init_res, batch_size=8, dev=DEV, use_ewma_gen=False,
                                   cutoff_trunc_trick=(min(4, int(math.log2(res)) - 2) or None)))
            if init_res != res:
                L.gen_model.increase_scale(); L.disc_model.increase_scale()
        assert shapes(L.gen_model) == want["g"], (model, res)
        assert shapes(L.disc_model) == want["d"], (model, res)
        del L


def test_fresh_networks_under_a_seed_equal_the_references(golden, monkeypatch):
    """Learner(config) under torch.manual_seed(s) starts from the reference's initial weights bit for bit (same initialisers drawing
    in the same order, incl. the blocks increase_scale() adds): sha256 of every parameter / buffer vs the reference's."""
    import hashlib
    import numpy as np
    import torch
    from gan_lab_b200.config import default_config
    from gan_lab_b200.progan.learner import ProGANLearner
    from gan_lab_b200.resnetgan.learner import GANLearner
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    dig = lambda t: hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()
    table = golden("init_digests.pt")
    for name, want in table.items():
        if name.startswith("ResNet"):
            torch.manual_seed(want["seed"])
            L = GANLearner(default_config("ResNet GAN", res=want["res"], batch_size=4, dev=DEV))
        else:
            monkeypatch.setattr(growth, "FMAP_MAX", want["fmap_max"]); monkeypatch.setattr(growth, "FMAP_BASE", want["fmap_base"])
            torch.manual_seed(want["seed"]); np.random.seed(want["seed"])
            L = (StyleGANLearner if name == "StyleGAN" else ProGANLearner)(default_config(name, dev=DEV, **want["kw"]))
            L.gen_model.increase_scale(); L.disc_model.increase_scale()
        for net, ref in ((L.gen_model, want["g"]), (L.disc_model, want["d"])):
            sd = net.state_dict()
            assert list(sd.keys()) == list(ref.keys()), name
            for k, v in sd.items():
                assert dig(v) == ref[k], (name, k)


def test_stylemixing_grid(tmp_path):
    """make_stylemixing_grid: first row = source B, first column = source A, cell (a, b) = A with B's styles from the group's
    stage on (stylegan/learner.py:306-431 without matplotlib): checked cell by cell against the evaluation-mode generator."""
    import numpy as np
    import torch
    from gan_lab_b200.config import default_config
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    torch.manual_seed(3)
    cfg = default_config("StyleGAN", res=32, batch_size=4, dev=DEV, len_latent=32, len_dlatent=32, cutoff_trunc_trick=2,
                         use_noise=False)
    L = StyleGANLearner(cfg)
    G = L.gen_model
    G.train(); G(torch.randn(4, 32)); G.eval()                      # one training forward creates w_ewma
    zb, zc, zf = torch.randn(3, 32), torch.randn(2, 32), torch.randn(1, 32)
    grid = L.make_stylemixing_grid(zb, zs_coarse=zc, zs_fine=zf, time_average=False, save_path=tmp_path / "mix.png")
    assert grid.shape == ((1 + 3) * 32, (1 + 3) * 32, 3) and grid.dtype == np.uint8
    assert (grid[:32, :32] == 255).all()

    def u8(x):
        return ((x[0].float() * .5 + .5).clamp(0, 1) * 255).round().to(torch.uint8).permute(1, 2, 0).numpy()

    def same(a, b):         # the grid evaluates whole rows as one batch: fp32 rounding may move a pixel by one uint8 step
        return int(np.abs(a.astype(int) - b.astype(int)).max()) <= 1

    with torch.no_grad():
        assert same(grid[:32, 64:96], u8(G(zb[1:2])))                                   # source B, column 2
        assert same(grid[64:96, :32], u8(G(zc[1:2])))                                   # source A (coarse row 2)
        assert same(grid[32:64, 96:128], u8(G(zc[0:1], x_mixing=zb[2:3], style_mixing_stage=1)))
        assert same(grid[96:128, 32:64], u8(G(zf[0:1], x_mixing=zb[0:1], style_mixing_stage=8)))
        assert not same(grid[32:64, 96:128], u8(G(zc[0:1])))                            # mixing did something
    from PIL import Image
    assert Image.open(tmp_path / "mix.png").size == (128, 128)
    assert not G.training


def test_batchnorm_eval_mode_and_resnet_metrics(golden, tmp_path):
    """BatchNorm2d in evaluation mode (running statistics; issued as a diagonal 1x1 convolution) against nn.BatchNorm2d, and with it
    the ResNet learner's validation metrics and image grid (the reference's compute_metrics runs its generator in eval mode)."""
    import torch
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    from gan_lab_b200 import ops
    from gan_lab_b200.utils import custom_layers as CL
    torch.manual_seed(0)
    ref = torch.nn.BatchNorm2d(12)
    with torch.no_grad():
        ref.weight.uniform_(.5, 1.5); ref.bias.normal_(); ref.running_mean.normal_(); ref.running_var.uniform_(.5, 2.)
    mine = CL.BatchNorm2d(12)
    mine.load_state_dict(ref.state_dict())
    ref.eval(); mine.eval()
    x = torch.randn(3, 12, 5, 7)
    torch.testing.assert_close(mine(x).contiguous(), ref(x), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mine(x, act=ops.ACT_LRELU, slope=0.0).contiguous(), torch.relu(ref(x)), rtol=1e-5, atol=1e-6)
    assert int(mine.num_batches_tracked) == 0                                   # eval mode leaves the buffers alone

    g = golden("resnet_nets_res32.pt")
    L, cfg = PC._resnet_learner(g, DEV, save_samples_dir=tmp_path, img_grid_sz=2)
    PC._load(L.gen_model, g["g_sd"]); PC._load(L.disc_model, g["d_sd"])
    zds = TensorDataset(torch.randn(6, cfg.len_latent)); xds = TensorDataset(torch.rand(6, 3, 32, 32) * 2 - 1)
    z_dl = DataLoader(zds, batch_sampler=BatchSampler(SequentialSampler(zds), batch_size=4, drop_last=False))
    x_dl = DataLoader(xds, batch_sampler=BatchSampler(SequentialSampler(xds), batch_size=4, drop_last=False))
    lines = L.compute_metrics(["fake realness", "generator loss", "image grid"], "Generator", z_dl)
    assert len(lines) == 2 and set(L.last_metrics) == {"fake realness", "generator loss"}
    assert abs(L.last_metrics["generator loss"] + L.last_metrics["fake realness"]) < 1e-5      # wgan: loss = -mean(D(G(z)))
    lines = L.compute_metrics(["real realness", "discriminator loss"], "Discriminator", z_dl, x_dl)
    assert len(lines) == 2 and all(torch.isfinite(torch.tensor(v)) for v in L.last_metrics.values())
    from PIL import Image
    assert Image.open(tmp_path / "resnetgan" / "image_grid" / "original" / "0.png").size == (64, 64)
    assert L.gen_model.training and L.disc_model.training


def test_resnet_compute_metrics_vs_reference(golden):
    """ResNet learner compute_metrics() vs the reference's raw values: eval-mode generator = BatchNorm on the running statistics."""
    import torch
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    g = golden("resnet_metrics_res32.pt")
    L, cfg = PC._resnet_learner(g, DEV, num_disc_iters=2)
    PC._load(L.gen_model, g["g_sd"]); PC._load(L.disc_model, g["d_sd"])
    bs = g["bs"]
    zds, xds = TensorDataset(g["z_valid"]), TensorDataset(g["x_valid"])
    z_dl = DataLoader(zds, batch_sampler=BatchSampler(SequentialSampler(zds), batch_size=bs, drop_last=False))
    x_dl = DataLoader(xds, batch_sampler=BatchSampler(SequentialSampler(xds), batch_size=bs, drop_last=False))
    lines = L.compute_metrics(g["gen_metrics"], "Generator", z_dl)
    for name, want in zip(g["gen_metrics"], g["raw_g"]):
        assert abs(L.last_metrics[name] - want) < 2e-4 * max(1.0, abs(want)), (name, L.last_metrics[name], want)
    assert [l.split(":")[0] for l in lines] == [l.split(":")[0] for l in g["vals_g"]]
    lines = L.compute_metrics(g["disc_metrics"], "Discriminator", z_dl, x_dl)
    for name, want in zip(g["disc_metrics"], g["raw_d"]):
        assert abs(L.last_metrics[name] - want) < 2e-4 * max(1.0, abs(want)), (name, L.last_metrics[name], want)
    assert [l.split(":")[0] for l in lines] == [l.split(":")[0] for l in g["vals_d"]]
