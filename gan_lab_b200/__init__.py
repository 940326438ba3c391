"""gan_lab_b200: B200-native (sm_100a) implementation of gan-lab's generator/discriminator training step.

Host code mirrors the reference's layer and Learner API (sidward14/gan-lab, `gan_lab/`); every op runs a
hand-written CUDA kernel through the C-ABI in include/ganlab_b200.h.  There is no CPU, Triton or PyTorch
fallback: without the built shared library (or without a CUDA device) the ops raise.
"""
from ._lib import LIB, GlbError  # noqa: F401
from ._kernels import set_conv_impl, get_conv_impl, launch_count  # noqa: F401

__version__ = "0.1.0"
