"""TEST INFRASTRUCTURE ONLY.

`oracle/` is a CPU restatement (plain PyTorch, fp32 or fp64) of the reference's
generator/discriminator training-step algorithm.  It is the *checker* for the
CUDA path, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.
"""
