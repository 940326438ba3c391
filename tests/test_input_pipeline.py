"""Real-image input pipeline: Pillow BOX resize + ToTensor + Normalize.

CPU (`-m "not gpu"`): the numpy restatement (oracle/pil_box.py) and the library's host-side coefficient helper against
Pillow itself; the loader's host logic on the CPU contract.  GPU: the sm_100a kernel, bit-exact against both.
"""
import numpy as np
import pytest
import torch
from PIL import Image

from gan_lab_b200 import _kernels as K
from oracle import kernel_contracts, pil_box

SIZES = [(8, 8, 4, 4), (64, 64, 8, 8), (256, 256, 32, 32), (100, 60, 32, 32), (33, 47, 16, 8), (16, 16, 32, 32),
         (7, 9, 7, 9), (12, 12, 5, 7), (256, 256, 4, 4), (10, 10, 3, 3), (64, 48, 48, 64), (1, 1, 1, 1), (5, 3, 1, 1)]
MEAN, STD = (.5, .5, .5), (.5, .5, .5)


def _images(n, h, w, seed=0):
    rng = np.random.default_rng(seed + h * 131 + w)
    a = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    a[0, : max(1, h // 2)] = 255          # saturated and zero regions: the clip8 paths
    if n > 1:
        a[1] = 0
    return a


def _pil_chain(images, out_h, out_w, mean, std, flip=None):
    """The reference's transform chain, literally (data_config.py:312-342)."""
    from torchvision import transforms as T
    chain = T.Compose([T.Resize((out_h, out_w), interpolation=Image.BOX), T.ToTensor(), T.Normalize(mean=list(mean), std=list(std))])
    outs = []
    for i, a in enumerate(images):
        im = Image.fromarray(a)
        if flip is not None and flip[i]:
            im = T.Resize((out_h, out_w), interpolation=Image.BOX)(im).transpose(Image.FLIP_LEFT_RIGHT)
            outs.append(T.Compose([T.ToTensor(), T.Normalize(mean=list(mean), std=list(std))])(im))
        else:
            outs.append(chain(im))
    return torch.stack(outs)


@pytest.mark.parametrize("h,w,oh,ow", SIZES)
def test_oracle_matches_pillow(h, w, oh, ow):
    imgs = _images(3, h, w)
    for a in imgs:
        ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.BOX))
        assert np.array_equal(pil_box.box_resize_u8(a, oh, ow), ref)
    flip = np.array([True, False, True])
    mean, std = (.485, .456, .406), (.229, .224, .225)
    assert torch.equal(pil_box.input_pipeline(imgs, None, (oh, ow), mean, std, flip), _pil_chain(imgs, oh, ow, mean, std, flip))


@pytest.mark.parametrize("n_in,n_out", [(8, 4), (1024, 128), (1024, 4), (100, 32), (47, 8), (16, 32), (9, 9), (12, 5), (10, 3),
                                        (1, 1), (4096, 512)])
def test_library_coefficient_tables_are_pillows(n_in, n_out):
    """glb_box_resize_tables is a host function of the product library (no GPU involved): same tables as the restatement."""
    bounds, kk = K.box_resize_tables(n_in, n_out)
    b_ref, k_ref = pil_box.coeffs(n_in, n_out)
    assert np.array_equal(bounds.numpy(), b_ref) and np.array_equal(kk.numpy(), k_ref)


def test_device_loader_host_logic(monkeypatch):
    """Epoch order, sharding, batch-size / resolution changes mid-epoch and mirror flags, on the CPU contract of the kernel."""
    from gan_lab_b200.data import DeviceImageLoader
    kernel_contracts.install(monkeypatch)
    imgs = torch.from_numpy(_images(10, 16, 16))
    dl = DeviceImageLoader(imgs, batch_size=4, res=4, shuffle=False, device="cpu")
    assert len(dl.dataset) == 10 and len(dl) == 2
    it = iter(dl)
    (x0,) = next(it)
    assert x0.shape == (4, 3, 4, 4) and torch.equal(x0, _pil_chain(imgs[:4].numpy(), 4, 4, MEAN, STD))
    dl.set_resolution(8); dl.batch_sampler.batch_size = 2          # what train() does at a resolution increase
    (x1,) = next(it)
    assert x1.shape == (2, 3, 8, 8) and torch.equal(x1, _pil_chain(imgs[4:6].numpy(), 8, 8, MEAN, STD))
    assert sum(1 for _ in it) == 2                                  # samples 6..9; nothing dropped at batch size 2
    # two ranks see disjoint strided shards of the same shuffled order; mirror flags flip whole samples
    shards = []
    for r in range(2):
        d = DeviceImageLoader(imgs, batch_size=1, res=16, shuffle=True, mirror=True, device="cpu", seed=3, rank=r, world_size=2)
        shards.append(torch.cat([x for (x,) in d]))
    full = _pil_chain(imgs.numpy(), 16, 16, MEAN, STD)
    seen = []
    for s in shards:
        assert s.shape[0] == 5
        for x in s:
            hits = [i for i in range(10) if torch.equal(x, full[i]) or torch.equal(x, full[i].flip(-1))]
            assert len(hits) >= 1
            seen.append(hits[0])
    assert sorted(seen) == list(range(10))


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("h,w,oh,ow", SIZES + [(1024, 1024, 128, 128), (1024, 1024, 8, 8), (300, 500, 64, 64)])
def test_kernel_bit_exact_vs_pillow(h, w, oh, ow):
    imgs = _images(5, h, w)
    src = torch.from_numpy(imgs).cuda()
    idx = torch.tensor([4, 0, 2, 2, 1])
    flip = torch.tensor([1, 0, 0, 1, 0], dtype=torch.uint8)
    mean, std = (.485, .456, .406), (.229, .224, .225)
    got = K.u8_box_resize_normalize(src, idx, (oh, ow), mean, std, flip).cpu()
    want = pil_box.input_pipeline(imgs, idx.numpy(), (oh, ow), mean, std, flip.numpy())
    assert got.shape == want.shape
    assert torch.equal(got, want), float((got - want).abs().max())
    if h * w <= 300 * 500:
        assert torch.equal(got, _pil_chain(imgs[idx.numpy()], oh, ow, mean, std, flip.numpy().astype(bool)))
    got2 = K.u8_box_resize_normalize(src, None, (oh, ow), MEAN, STD).cpu()
    assert torch.equal(got2, pil_box.input_pipeline(imgs, None, (oh, ow), MEAN, STD))


@pytest.mark.gpu
def test_kernel_unaligned_views_and_pinned_loader():
    """Source tensors whose first byte is not 16-byte aligned (a slice of a larger buffer) and the pinned-host loader path."""
    from gan_lab_b200.data import DeviceImageLoader
    imgs = _images(6, 37, 53)
    flat = torch.zeros(imgs.size + 64, dtype=torch.uint8, device="cuda")
    for off in (0, 1, 7, 15):
        view = flat[off:off + imgs.size].view(6, 37, 53, 3)
        view.copy_(torch.from_numpy(imgs))
        got = K.u8_box_resize_normalize(view, None, (16, 16), MEAN, STD).cpu()
        assert torch.equal(got, pil_box.input_pipeline(imgs, None, (16, 16), MEAN, STD)), off
    dl = DeviceImageLoader(torch.from_numpy(imgs), batch_size=3, res=8, shuffle=False, device="cuda")
    xs = torch.cat([x for (x,) in dl]).cpu()
    assert torch.equal(xs, pil_box.input_pipeline(imgs, None, (8, 8), MEAN, STD))
    dl2 = DeviceImageLoader(torch.from_numpy(imgs).cuda(), batch_size=3, res=8, shuffle=False, device="cuda")
    assert torch.equal(torch.cat([x for (x,) in dl2]).cpu(), xs)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,oh,ow", [(40, 300, 20, 150), (24, 520, 24, 260), (16, 257, 16, 257), (512, 512, 256, 256)])
def test_kernel_bit_exact_several_column_tiles(h, w, oh, ow):
    imgs = _images(3, h, w)
    got = K.u8_box_resize_normalize(torch.from_numpy(imgs).cuda(), None, (oh, ow), MEAN, STD).cpu()
    assert torch.equal(got, pil_box.input_pipeline(imgs, None, (oh, ow), MEAN, STD))


def test_no_cpu_path():
    """The product path has no CPU fallback: a host tensor is refused loudly (the CPU restatement lives under oracle/ only)."""
    from gan_lab_b200._lib import GlbError
    imgs = torch.from_numpy(_images(2, 8, 8))
    with pytest.raises(GlbError):
        K.u8_box_resize_normalize(imgs, None, (4, 4), MEAN, STD)
    from gan_lab_b200.data import DeviceImageLoader
    with pytest.raises(GlbError):
        next(iter(DeviceImageLoader(imgs, batch_size=2, res=4, shuffle=False, device="cpu")))
    with pytest.raises(ValueError):
        DeviceImageLoader(imgs.float(), batch_size=2, res=4)
