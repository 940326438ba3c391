"""Calibration of the TF32 whole-step bounds (TEST INFRASTRUCTURE; run on a GPU box, not collected by pytest):

    python tests/calibrate_tf32_bounds.py > gpurun_out/r2_reference_on_b200_deviation.json

Runs the UNMODIFIED reference (baseline/_ref, the pip-installed copy that travels with the gpurun snapshot) through stock
PyTorch on the B200 -- cuDNN / cuBLAS, once with TF32 disabled and once with PyTorch's defaults (cuDNN convolutions in TF32) --
on the inputs of tests/golden/style_cfg2_fullwidth_step.pt (same seeds, the CPU run's taped random draws replayed) and reports
how far each run is from the reference's own CPU fp32 run, in the metrics tests/test_cfg2_fullwidth.py uses.  The TF32 line is
the yardstick for this repo's tcgen05 kind::tf32 path: the reference's own GPU default deviates that much from its CPU path.
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import make_golden as MG          # noqa: E402
from oracle import summaries as S             # noqa: E402
from oracle.reference_loader import load_reference, reference_available   # noqa: E402


def deviation(run, g):
    gd = lambda d, tag: {k: v for k, v in d.items() if k.startswith(tag)}
    per = lambda a, b: {k: S.compare_summaries(a[k], b[k]) for k in b}
    out = dict(
        loss_d=abs(run["losses"][0] - g["losses"][0]) / max(1.0, abs(g["losses"][0])),
        loss_g=abs(run["losses"][1] - g["losses"][1]) / max(1.0, abs(g["losses"][1])),
        gp_value=abs(run["gp_alone"] - g["gp_alone"]) / abs(g["gp_alone"]),
        g_alone_loss=abs(run["g_alone"]["loss"] - g["g_alone"]["loss"]) / max(1.0, abs(g["g_alone"]["loss"])),
        g_alone_img=S.compare_summaries(run["g_alone"]["img"], g["g_alone"]["img"]),
        g_alone_logits=float((run["g_alone"]["logits"] - g["g_alone"]["logits"]).abs().max() / g["g_alone"]["logits"].abs().max()),
        gp_grads=S.class_stats(per(run["gp_grads"], g["gp_grads"])),
        g_alone_grads=S.class_stats(per(run["g_alone"]["grads"], g["g_alone"]["grads"])),
        d_grads=S.class_stats(per(gd(run["grads"], "d."), gd(g["grads"], "d."))),
        g_grads=S.class_stats(per(gd(run["grads"], "g."), gd(g["grads"], "g."))),
        p1_flip_fraction=MG._flip_fraction(run["p1"], g["p1"], g["lr"]))
    return out


def main():
    assert reference_available() and torch.cuda.is_available()
    ref = load_reference()
    g = torch.load(ROOT / "tests" / "golden" / "style_cfg2_fullwidth_step.pt", weights_only=False)
    res = {"what": "unmodified reference on the B200 through stock PyTorch vs its own CPU fp32 run (the fixture)",
           "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}
    for mode in ("fp32", "tf32"):
        torch.backends.cudnn.allow_tf32 = (mode == "tf32")           # PyTorch's default is True: cuDNN convolutions in TF32
        torch.backends.cuda.matmul.allow_tf32 = False                 # PyTorch's default
        run = MG._cfg2_fullwidth_run(ref, g["res"], g["bs"], g["seed"], False, dev="cuda",
                                     replay={"g_alone": g["g_alone"]["tape"], "train": g["tape"]})
        assert run["g_digests"] == g["g_digests"] and run["d_digests"] == g["d_digests"] and run["data_digest"] == g["data_digest"]
        res["reference_cudnn_" + mode] = deviation(run, g)
    res["reference_cpu_one_ulp_weight_perturbation"] = g["self_noise"]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
