"""GPU parity of the rows SURVEY.md 8f adds around the train step (progressive-growing schedule, checkpoints, input
pipeline).  Kept in a file that sorts last: these cases were written after the round's GPU budget was spent, so a surprise
here must not hide the results of tests/test_gpu_parity.py under `pytest -x`."""
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
