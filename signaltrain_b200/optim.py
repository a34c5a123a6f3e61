"""Adam for the SignalTrain model, one multi-tensor CUDA launch per step.  Drop-in for the
`torch.optim.Adam(model.parameters(), lr=..., weight_decay=0)` the reference builds at train.py:228:
same `param_groups` / `state_dict()` layout (per-parameter `step`, `exp_avg`, `exp_avg_sq`)."""
import torch


class Adam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if weight_decay != 0:
            raise NotImplementedError("signaltrain_b200.optim.Adam: weight_decay must be 0 (train.py:228)")
        self._model = model
        params = model.ordered_parameters()
        # the group carries every key torch.optim.Adam's does, so a checkpoint's optimizer entry loads into either class
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                                      capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False))
        self._step = 0

    def load_state_dict(self, state_dict):
        """torch's loader restores exp_avg / exp_avg_sq / step per parameter; the fused step reads its counter from
        `_step`, so it is taken from the restored state (bias correction must continue where the checkpoint stopped)."""
        super().load_state_dict(state_dict)
        params = self._model.ordered_parameters()
        steps = [float(self.state[p]["step"]) for p in params if "step" in self.state.get(p, {})]
        self._step = int(max(steps)) if steps else 0

    def _state_lists(self):
        params = self._model.ordered_parameters()
        m, v = [], []
        for p in params:
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            m.append(st["exp_avg"])
            v.append(st["exp_avg_sq"])
        return params, m, v

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0, max_norm=0.0):
        if closure is not None:
            raise NotImplementedError("closure is not supported")
        params, m, v = self._state_lists()
        grads = [p.grad for p in params]
        if any(g is None for g in grads):
            raise RuntimeError("signaltrain_b200.optim.Adam.step: every parameter needs a gradient")
        group = self.param_groups[0]
        self._step = int(self.state[params[0]]["step"].item()) + 1 if self._step == 0 else self._step + 1
        eng = self._model.mpaec._engine_for_device(params[0].device)
        hp = eng.adam_hp(group["lr"], self._step, group["betas"], group["eps"], grad_scale, max_norm)
        eng.adam_step([p.data for p in params], [g.contiguous() for g in grads], m, v, hp)
        for p in params:
            self.state[p]["step"] = torch.tensor(float(self._step))
        return None
