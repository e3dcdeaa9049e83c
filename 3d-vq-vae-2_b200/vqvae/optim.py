"""Adam(amsgrad=True) of the reference's configure_optimizers (vqvae/model.py:91-93) on the library's step kernel
(vq3d_adam_amsgrad_step): same update rule and state names as torch.optim.Adam, one launch per parameter tensor."""
import torch

from . import _ops


class FusedAdamAMSGrad(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        o = _ops.default()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                    st["max_exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                o.adam_amsgrad_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], st["max_exp_avg_sq"], group["lr"], b1, b2, group["eps"], st["step"])
        return loss
