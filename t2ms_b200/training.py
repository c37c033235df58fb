"""Training step of the DiT (train.py:66-87).  Placeholder until the backward kernels land."""
from __future__ import annotations


def dit_forward_autograd(model, x, t, text):
    raise NotImplementedError(
        "t2ms_b200: the DiT backward kernels are not built yet; call the model under torch.no_grad() / .eval() "
        "for generation (there is deliberately no PyTorch fallback)")
