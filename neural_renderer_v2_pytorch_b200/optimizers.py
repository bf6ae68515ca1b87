"""``Adam`` with a parameter-wise learning-rate factor ``param.lr`` (reference
``neural_renderer_torch/optimizers.py:9-37``, which subclasses *chainer*'s Adam and therefore cannot
run in the torch package).  Same update, written as a ``torch.optim.Optimizer``:

    m += (1 - beta1) (g - m);  v += (1 - beta2) (g^2 - v);  v = max(v, 0)
    param -= lr_t * factor * m / (sqrt(v) + eps),   lr_t = alpha sqrt(1 - beta2^t) / (1 - beta1^t)

(chainer's bias-corrected step size), ``factor = param.lr`` when the attribute exists; a parameter
without gradient, or with ``factor == 0``, is left alone (optimizers.py:15-19).  Outside the
accelerated path."""
import math

import torch


class Adam(torch.optim.Optimizer):
    def __init__(self, params, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8):
        super().__init__(params, dict(alpha=alpha, beta1=beta1, beta2=beta2, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["beta1"], group["beta2"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                factor = getattr(p, "lr", 1.0)
                if factor == 0:
                    continue
                st = self.state[p]
                if not st:
                    st["t"] = 0
                    st["m"] = torch.zeros_like(p)
                    st["v"] = torch.zeros_like(p)
                st["t"] += 1
                g, m, v = p.grad, st["m"], st["v"]
                m.add_(g - m, alpha=1 - b1)
                v.add_(g * g - v, alpha=1 - b2).clamp_(min=0)
                lr_t = group["alpha"] * math.sqrt(1 - b2 ** st["t"]) / (1 - b1 ** st["t"])
                p.addcdiv_(m, v.sqrt().add_(group["eps"]), value=-lr_t * factor)
        return loss
