"""GPU parity of the rows SURVEY.md 8f adds around the train step (progressive-growing schedule, checkpoints, input
pipeline).  Kept in a file that sorts last so that the kernel / nets / train parity of tests/test_gpu_parity.py reports first under `pytest -x`."""
import pytest
import torch

import gan_lab_b200._growth as growth
import gan_lab_b200._kernels as K
from gan_lab_b200.utils.latent_utils import set_random_source

import parity_cases as PC

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def small_fmaps(monkeypatch):
    monkeypatch.setattr(growth, "FMAP_MAX", 32)
    K.set_conv_impl("fp32")
    yield
    set_random_source(None)


@pytest.mark.parametrize("fname,model", PC.GROW_CASES)
def test_learner_grow(golden, fname, model):
    PC.case_learner_grow(golden, DEV, fname, model)


@pytest.mark.parametrize("fname,model", PC.GROW_CASES)
def test_learner_grow_device_alpha(golden, fname, model):
    PC.case_learner_grow(golden, DEV, fname, model, device_alpha=True)


@pytest.mark.parametrize("fname,model", PC.RESUME_CASES)
def test_learner_resume_from_reference_checkpoint(golden, fname, model):
    from conftest import GOLDEN
    PC.case_learner_resume(golden, DEV, fname, model, GOLDEN)


@pytest.mark.parametrize("fname,model", PC.GROW_CASES)
def test_checkpoint_roundtrip(golden, fname, model, tmp_path):
    PC.case_checkpoint_roundtrip(golden, DEV, fname, model, tmp_path)


@pytest.mark.parametrize("fname,model", PC.METRICS_CASES)
def test_compute_metrics(golden, fname, model, tmp_path):
    PC.case_compute_metrics(golden, DEV, fname, model, tmp_path)


def test_fade_in_phase_replays_as_cuda_graph_with_device_alpha():
    """A fade-in phase with a MOVING alpha captured once (alpha lives in a device vector the blend kernels read): the graph
    key does not change with alpha, and the replayed graph really reads it -- with the vector forced to alpha = 1 the branch
    being faded out gets exactly zero gradient, with alpha = 0 the new branch does, whatever latents the replay draws."""
    import math
    from gan_lab_b200.config import default_config
    from gan_lab_b200.stylegan.learner import StyleGANLearner
    torch.manual_seed(0)
    cfg = default_config("StyleGAN", res=16, init_res=8, batch_size=4, dev=DEV, len_latent=32, len_dlatent=32,
                         cutoff_trunc_trick=2)
    L = StyleGANLearner(cfg)
    L.gen_model.increase_scale(); L.disc_model.increase_scale()
    L.gen_model.to(cfg.dev); L.disc_model.to(cfg.dev)
    L.gen_model.alpha = 0.2
    L.beta = L.get_smoothing_ewma_beta(10.)
    L._sync_lagged_structure()
    L._set_optimizer()
    L.gen_model.train(); L.disc_model.train()
    L._init_lagged(); L._attach_ewma()
    L.delta_alpha = 0.05
    L.enable_cuda_graphs(True, warmup_iters=2, device_alpha=True)
    x = torch.rand(4, 3, 16, 16, device=DEV) * 2 - 1
    keys = []
    for i in range(5):
        ld, lg = L.main_iteration(x)
        keys.append(L._graph_key())
        L.gen_model.alpha += L.delta_alpha
    torch.cuda.synchronize()
    assert L._graph is not None and len(set(keys)) == 1                 # captured once, replayed while alpha moved
    assert abs(float(L.state.alpha_dev.coef[0]) - 0.4) < 1e-6          # the value pushed before the last replay
    assert math.isfinite(float(ld)) and math.isfinite(float(lg))
    G = L.gen_model
    for alpha, dead, alive in ((1.0, G.prev_torgb, G.torgb), (0.0, G.torgb, G.prev_torgb)):
        L.state.alpha_dev.coef.copy_(torch.tensor([alpha, 1.0 - alpha]))
        L._graph["gd"].replay(); L._graph["gg"].replay()
        torch.cuda.synchronize()
        assert float(dead.conv2d.weight.grad.abs().max()) == 0.0, alpha
        assert float(alive.conv2d.weight.grad.abs().max()) > 0.0, alpha


# Configuration switches the headline fixtures leave at their defaults (host wiring also pinned on the CPU doubles,
# tests/test_host_wiring.py::test_train_variant).
@pytest.mark.parametrize("name", PC.TRAIN_VARIANTS)
def test_train_variant(golden, name, monkeypatch):
    PC.case_train_variant(golden, DEV, name, monkeypatch)


@pytest.mark.parametrize("gp,bs", [("r1", 4), ("r2", 8), ("r1", 2)])
def test_batched_d_passes(gp, bs):
    PC.case_batched_d_passes(DEV, gp, bs)


def test_resnet_resume_from_reference_checkpoint(golden):
    from conftest import GOLDEN
    PC.case_resnet_resume(golden, DEV, GOLDEN)


def test_resnet_steps_replay_as_cuda_graphs(golden):
    """ResNet GAN: one generator step and one discriminator step captured and replayed (opt-in): Adam's device-side step count,
    BatchNorm's num_batches_tracked and the parameters advance per replay, fresh latents per replay, everything finite."""
    import math
    g = golden("resnet_train_res64.pt")
    L, _ = PC._resnet_learner(g, DEV, num_disc_iters=2)
    L.gen_model.train(); L.disc_model.train()
    L.enable_cuda_graphs(True, warmup_iters=2)
    xs = [torch.rand(g["bs"], 3, g["res"], g["res"], device=DEV) * 2 - 1 for _ in range(2)]
    for _ in range(2):
        L.main_iteration(xs)
    assert L._graph is None
    ld, lg = L.main_iteration(xs)                       # capture + first replay
    assert L._graph is not None
    torch.cuda.synchronize()
    nbt = next(b for n, b in L.gen_model.named_buffers() if n.endswith("num_batches_tracked"))
    n0, t0 = int(nbt), float(next(iter(L.opt_disc._hyper.values()))["t"][3])
    w0 = next(L.disc_model.parameters()).detach().clone()
    l1 = (float(ld), float(lg))
    ld, lg = L.main_iteration(xs)
    torch.cuda.synchronize()
    l2 = (float(ld), float(lg))
    assert int(nbt) == n0 + 3                           # 1 generator step + 2 discriminator steps run the generator forward
    assert float(next(iter(L.opt_disc._hyper.values()))["t"][3]) == t0 + 2
    assert not torch.equal(w0, next(L.disc_model.parameters()).detach())
    assert all(math.isfinite(v) for v in l1 + l2) and l1 != l2
    for p in list(L.gen_model.parameters()) + list(L.disc_model.parameters()):
        assert torch.isfinite(p).all()


# ---- tapering channel counts (32, 32, 16, 8)


@pytest.mark.parametrize("fname", PC.STYLE_NETS_TAPER)
def test_style_nets_modules_taper(golden, fname):
    PC.case_style_nets_modules(golden, DEV, fname)


@pytest.mark.parametrize("fname", PC.PRO_NETS_TAPER)
def test_pro_nets_modules_taper(golden, fname):
    PC.case_pro_nets_modules(golden, DEV, fname)


@pytest.mark.parametrize("fname,model", PC.GROW_TAPER_CASES)
def test_learner_grow_taper(golden, fname, model):
    PC.case_learner_grow(golden, DEV, fname, model, device_alpha=True)


def test_resnet_train_variant(golden):
    PC.case_resnet_train(golden, DEV, "resnet_train_res32_variant.pt")


def test_batchnorm_eval_mode_vs_torch():
    from gan_lab_b200 import ops
    from gan_lab_b200.utils import custom_layers as CL
    torch.manual_seed(0)
    ref = torch.nn.BatchNorm2d(64).to(DEV)
    with torch.no_grad():
        ref.weight.uniform_(.5, 1.5); ref.bias.normal_(); ref.running_mean.normal_(); ref.running_var.uniform_(.5, 2.)
    mine = CL.BatchNorm2d(64).to(DEV)
    mine.load_state_dict(ref.state_dict())
    ref.eval(); mine.eval()
    x = torch.randn(4, 64, 16, 16, device=DEV)
    K.set_conv_impl("tf32")                 # the eval path must pick the exact kernel by itself
    try:
        torch.testing.assert_close(mine(x).contiguous(), ref(x), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(mine(x, act=ops.ACT_LRELU, slope=0.0).contiguous(), torch.relu(ref(x)), rtol=1e-5, atol=1e-5)
        assert K.get_conv_impl() == "tf32"
    finally:
        K.set_conv_impl("fp32")


def test_resnet_compute_metrics_vs_reference(golden):
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    g = golden("resnet_metrics_res32.pt")
    L, cfg = PC._resnet_learner(g, DEV, num_disc_iters=2)
    PC._load(L.gen_model, g["g_sd"]); PC._load(L.disc_model, g["d_sd"])
    zds, xds = TensorDataset(g["z_valid"]), TensorDataset(g["x_valid"])
    z_dl = DataLoader(zds, batch_sampler=BatchSampler(SequentialSampler(zds), batch_size=g["bs"], drop_last=False))
    x_dl = DataLoader(xds, batch_sampler=BatchSampler(SequentialSampler(xds), batch_size=g["bs"], drop_last=False))
    L.compute_metrics(g["gen_metrics"], "Generator", z_dl)
    for name, want in zip(g["gen_metrics"], g["raw_g"]):
        assert abs(L.last_metrics[name] - want) < 2e-4 * max(1.0, abs(want)), (name, L.last_metrics[name], want)
    L.compute_metrics(g["disc_metrics"], "Discriminator", z_dl, x_dl)
    for name, want in zip(g["disc_metrics"], g["raw_d"]):
        assert abs(L.last_metrics[name] - want) < 2e-4 * max(1.0, abs(want)), (name, L.last_metrics[name], want)


def test_resnet_nets_with_folded_resampling_match_the_two_kernel_sequence(monkeypatch):
    """TF32 path of the full-width ResNet nets (64 x 64, channels 512 ... 64): the upsample folded into the generator's 3x3
    convolutions (glb_upconv_*), the average pool folded into the discriminator's (glb_downconv_*) and the 1x1 skip convolution
    run before its upsample, against the literal two-kernel sequences (GLB_UPCONV=0 / GLB_DOWNCONV=0) on the same weights and
    inputs; eval-mode nets (running statistics), so no batch statistic amplifies the TF32 rounding difference of the two orders."""
    from gan_lab_b200.config import default_config
    from gan_lab_b200.resnetgan.learner import GANLearner
    torch.manual_seed(0)
    cfg = default_config("ResNet GAN", res=64, batch_size=4, dev=DEV, len_latent=128)
    L = GANLearner(cfg)
    G, D = L.gen_model.to(DEV), L.disc_model.to(DEV)
    G.eval(); D.eval()
    z = torch.randn(4, 128, device=DEV)
    x = torch.rand(4, 3, 64, 64, device=DEV) * 2 - 1
    K.set_conv_impl("tf32")
    try:
        outs = []
        for fold in ("1", "0"):
            monkeypatch.setenv("GLB_UPCONV", fold); monkeypatch.setenv("GLB_DOWNCONV", fold)
            K.weights_updated()
            n0 = K.launch_count()
            with torch.no_grad():
                img = G(z)
                d = D(x)
            outs.append((img.float().clone(), d.float().clone(), K.launch_count() - n0))
        (ia, da, na), (ib, db, nb) = outs
        assert not torch.equal(ia, ib) and not torch.equal(da, db)  # the folded kernels really ran (same values to TF32 rounding, not bitwise)
        assert float((ia - ib).abs().max()) < 5e-3 * float(ib.abs().max())
        assert float((da - db).abs().max()) < 5e-3 * max(1.0, float(db.abs().max()))
    finally:
        K.set_conv_impl("fp32")
