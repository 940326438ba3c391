"""Whole-step parity of the path bench.py times: cfg2 (StyleGAN 128x128, batch 8) at its REAL widths (512 / 256 / 128
channels), tensor-core (TF32) convolutions, run eagerly and as the captured-and-replayed CUDA graph pair, against one main
iteration of the UNMODIFIED reference (tests/golden/style_cfg2_fullwidth_step.pt, made by oracle/make_golden.py from
progan/learner.py:734-916 + stylegan/architectures.py:411-528 on CPU fp32).

Errors are per tensor: the estimated relative L2 error of the WHOLE tensor (random projections) and the max-norm error on a
strided sample (oracle/summaries.py).  Bounds are multiples of two measured yardsticks, see the comment above FP32_MULT and
DESIGN.md section 2 ("The benchmarked path, pinned") for the measured values.
"""
import json
import os

import pytest
import torch

import parity_cases as PC

DEV = "cuda"


def _dump(tag, out):
    d = os.environ.get("GLB_DUMP_PARITY")
    if d:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, f"cfg2_fullwidth_{tag}.json"), "w") as f:
            json.dump(out, f, indent=1)


# Bounds.  fp32 path: multiples of the reference's OWN noise floor at this size (fixture key `self_noise`: the unmodified
# reference re-run on the CPU with every weight moved by one fp32 ulp; worst tensor per class).  TF32 path: multiples of what the
# unmodified reference ITSELF deviates by when it runs on the B200 through stock PyTorch with its default cuDNN-TF32 convolutions
# (tests/golden/cfg2_reference_on_b200_deviation.json, measured by tests/calibrate_tf32_bounds.py on the same fixture): at random
# initialisation leaky-ReLU masks within rounding of zero and mean-removing InstanceNorms amplify a 5e-4 operand rounding to
# 2-8 % (median over tensors) / up to 80 % (cancellation-dominated bias gradients of the R1 penalty) in ANY TF32 implementation.
FP32_LOSS, FP32_MULT = 1e-4, 6.0     # measured: <= 3.6x the floor (split-K / atomics order varies run to run)
TF32_MULT = 2.0                       # measured: 0.9-1.3x the yardstick
BF16_MULT = 12.0         # bf16 operands carry 8 mantissa bits: 4x the TF32 rounding step; measured 2-10x the yardstick (opt-in mode,
                         # not the default: image 2 %, logits 3 %, gradients 9-19 % (median over tensors) at random initialisation)


def _tf32_yardstick():
    with open(os.path.join(os.path.dirname(__file__), "golden", "cfg2_reference_on_b200_deviation.json")) as f:
        return json.load(f)["reference_cudnn_tf32"]


@pytest.mark.gpu
@pytest.mark.parametrize("impl,graph", [("fp32", False), ("tf32", False), ("tf32", True), ("bf16", False), ("bf16", True)])
def test_cfg2_fullwidth_step_vs_reference(golden, impl, graph):
    out = PC.case_cfg2_fullwidth_step(golden, DEV, impl, graph)
    _dump(f"{impl}_{'graph' if graph else 'eager'}", out)
    msg = {k: v for k, v in out.items() if not k.endswith("per_tensor")}
    if impl == "fp32":
        floor = golden("style_cfg2_fullwidth_step.pt")["self_noise"]
        mult = FP32_MULT
        assert out["loss_d"] < FP32_LOSS and out["g_alone_loss"] < FP32_LOSS, msg
        assert out["loss_g"] < max(FP32_LOSS, mult * floor["loss_g"]), msg          # behind the discriminator's first Adam step
        assert out["gp_value"] < max(10 * FP32_LOSS, mult * floor["gp_value"]), msg
        assert out["g_alone_img_l2"] < 10 * FP32_LOSS and out["g_alone_logits"] < 10 * FP32_LOSS, msg
        assert out["w_ewma"] < 10 * FP32_LOSS, msg
        flips = 0.03
    else:
        floor = _tf32_yardstick()
        mult = TF32_MULT if impl == "tf32" else BF16_MULT
        base = 1e-3          # never tighter than this: the yardstick's own small entries are single draws
        assert out["loss_d"] < max(base, mult * floor["loss_d"]) and out["g_alone_loss"] < max(base, mult * floor["g_alone_loss"]), msg
        assert out["loss_g"] < max(3 * base, 4 * mult * floor["loss_g"]), msg      # behind the discriminator's first Adam step
        assert out["gp_value"] < max(5 * base, 4 * mult * floor["gp_value"]), msg
        assert out["g_alone_img_l2"] < mult * floor["g_alone_img"]["l2"], msg
        assert out["g_alone_logits"] < mult * floor["g_alone_logits"], msg
        assert out["w_ewma"] < 5e-3, msg
        flips = max(0.03, mult * floor["p1_flip_fraction"])
    for key in ("gp_grads", "g_alone_grads", "d_grads", "g_grads"):
        fl = floor[key]
        assert out[key + "_l2"] < mult * fl["l2"], (key, out[key + "_l2"], fl, msg)
        assert out[key + "_smax"] < mult * fl["smax"], (key, out[key + "_smax"], fl, msg)
    assert out["p1_max_abs_diff_over_lr"] <= 2.001, msg
    assert out["p1_flip_fraction"] < flips, msg
    assert out["lagged_rel_err"] < 5e-7, msg


def test_cfg2_fullwidth_fixture_is_reproducible_from_its_seeds(golden):
    """No GPU needed: the drop-in learner built on the CPU under the fixture's seed has the reference's initial weights bit for
    bit at full width (sha256 per tensor) and regenerates the same real batch; the summaries carry every parameter."""
    g = golden("style_cfg2_fullwidth_step.pt")
    L, data = PC.cfg2_fullwidth_learner(g, "cpu")
    names = {"g." + n for n, _ in L.gen_model.named_parameters()} | {"d." + n for n, _ in L.disc_model.named_parameters()}
    assert set(g["p1"].keys()) == names
    trained = {n for n in names if not n.startswith(("g.prev_torgb", "d.prev_fromrgb"))}
    assert set(g["grads"].keys()) == trained
    assert len(g["losses"]) == 2 and g["gp_alone"] > 0
