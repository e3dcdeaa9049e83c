"""Adam(amsgrad=True) of the reference's configure_optimizers (vqvae/model.py:91-93) on the library's step kernel
(vq3d_adam_amsgrad_step): same update rule and state names as torch.optim.Adam.

The 7.5 M parameters of the Full model live in ~7 000 tensors (every Fixup scalar is one), so a per-tensor step is
thousands of 10 us launches.  With `flatten=True` (default) the optimizer moves the parameters into ONE flat fp32 buffer
(each `p.data` becomes a view of it), keeps the gradients and the three moment buffers flat as well (`p.grad` and the
per-parameter state entries are views), and a step is ONE launch; the flat gradient is also the buffer that the
data-parallel all-reduce sends (`vqvae.parallel.allreduce_gradients`)."""
import torch

from . import _ops


class FusedAdamAMSGrad(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, flatten=True):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.flat_param = self.flat_grad = None
        self._step = 0
        if flatten and len(self.param_groups) == 1:
            self._flatten()

    @torch.no_grad()
    def _flatten(self):
        ps = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not ps or any((not p.is_cuda) or p.dtype != torch.float32 for p in ps) or len({p.device for p in ps}) != 1:
            return
        n = sum(p.numel() for p in ps)
        dev = ps[0].device
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self._m, self._v, self._vmax = (torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(3))
        self._step_state = torch.zeros(3, dtype=torch.float64, device=dev)       # step, 1 - b1^step, 1 - b2^step (device side: graph replay)
        self._slices = []
        off = 0
        for p in ps:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view(p.shape)
            if p.grad is not None:
                self.flat_grad[off:off + k].copy_(p.grad.reshape(-1))
            gv = self.flat_grad[off:off + k].view(p.shape)
            p.grad = gv
            st = self.state[p]
            st["step"] = 0
            st["exp_avg"] = self._m[off:off + k].view(p.shape)
            st["exp_avg_sq"] = self._v[off:off + k].view(p.shape)
            st["max_exp_avg_sq"] = self._vmax[off:off + k].view(p.shape)
            self._slices.append((p, off, k, gv))
            off += k

    def current_step(self) -> int:
        """Optimisation steps applied so far.  In flat mode the authoritative counter is the device-side one (a CUDA-graph
        replay advances it without the host seeing it)."""
        if self.flat_grad is None:
            return max((int(st.get("step", 0)) for st in self.state.values()), default=0)
        return int(round(float(self._step_state[0])))

    def state_dict(self):
        """torch.optim.Adam's layout (per-parameter `step`, `exp_avg`, `exp_avg_sq`, `max_exp_avg_sq`), with the step
        counters brought up to date first (in flat mode they are refreshed lazily)."""
        if self.flat_grad is not None:
            self._step = self.current_step()
            for p, _, _, _ in self._slices:
                self.state[p]["step"] = self._step
        return super().state_dict()

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """Restore a `state_dict()` (or a torch.optim.Adam(amsgrad=True) one of the same parameters).  In flat mode the
        loaded moments are copied INTO the flat buffers and the per-parameter entries become views of them again, and the
        step counter / bias corrections on the device are rebuilt, so the next step continues the saved trajectory."""
        super().load_state_dict(state_dict)
        if self.flat_grad is None:
            return
        step = 0
        for p, off, k, _ in self._slices:
            st = self.state[p]
            for name, flat in (("exp_avg", self._m), ("exp_avg_sq", self._v), ("max_exp_avg_sq", self._vmax)):
                view = flat[off:off + k].view(p.shape)
                loaded = st.get(name)
                if loaded is None:
                    view.zero_()
                elif loaded.data_ptr() != view.data_ptr():
                    view.copy_(loaded.to(view.device, torch.float32))
                st[name] = view
            s = st.get("step", 0)
            step = max(step, int(s.item() if torch.is_tensor(s) else s))
            st["step"] = step
        b1, b2 = self.param_groups[0]["betas"]
        self._step = step
        self._step_state.copy_(torch.tensor([float(step), 1.0 - b1 ** step, 1.0 - b2 ** step], dtype=torch.float64))
        for p, _, _, _ in self._slices:
            self.state[p]["step"] = step

    def zero_grad(self, set_to_none: bool = True):
        if self.flat_grad is None:
            return super().zero_grad(set_to_none=set_to_none)
        self.flat_grad.zero_()              # the .grad views stay in place: autograd accumulates into them

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        o = _ops.default()
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        if self.flat_grad is not None:
            check_params = self._step % 100 == 0
            for p, off, k, gv in self._slices:
                if p.grad is not gv:                    # a gradient that was replaced (e.g. Module.zero_grad(set_to_none=True)) is folded back in
                    if p.grad is None:
                        gv.zero_()
                    else:
                        gv.copy_(p.grad)
                    p.grad = gv
                if check_params and p.data_ptr() != self.flat_param.data_ptr() + 4 * off:
                    raise RuntimeError("FusedAdamAMSGrad: a parameter was re-allocated after the optimizer flattened it "
                                       "(move the model to its device before configure_optimizers())")
            self._step += 1
            if check_params:                            # per-parameter step counters (torch.optim.Adam's state layout) follow lazily
                for p, _, _, _ in self._slices:
                    self.state[p]["step"] = self._step
            o.adam_amsgrad_step_dev(self.flat_param, self.flat_grad, self._m, self._v, self._vmax, group["lr"], b1, b2, group["eps"], self._step_state)
            return loss
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
