"""Drop-in for sound_synthesis/engine/ema.py::EMA (the 'on-GPU EMA' item).

Same constructor / update / state_dict / load_state_dict / modify_to_inference / modify_to_train surface.  The reference keeps the shadow model
on the CPU and, every `update_interval` iterations, clones the whole live state_dict to the host, blends it tensor by tensor and reloads it
(1.5 GB down the PCIe bus per update for the 383 M-parameter denoiser).  Here the shadow lives where the model lives (`device=None`, the default)
and an update is two fused multi-tensor launches (`torch._foreach_mul_` / `_foreach_add_`) over the floating-point entries, in place; integer
buffers are copied.  `device=torch.device('cpu')` reproduces the reference's placement.  This is optimizer-side code like torch.optim.AdamW: it is
not on the sampling / training hot path and uses stock torch kernels.
"""
from __future__ import annotations

import copy

import torch


class EMA(object):
    def __init__(self, model, decay=0.99, update_interval=1, device=None):
        self.decay = decay
        self.update_iterval = update_interval  # (sic) the reference's attribute name
        self.model = model
        src = self._source()
        self.device = device if device is not None else next(src.parameters()).device
        with torch.no_grad():
            self.ema_model = copy.deepcopy(src)
        self.ema_model.to(self.device)
        for p in self.ema_model.parameters():
            p.requires_grad_(False)
        self.cur_state_dict = {k: v.clone().to(self.device) for k, v in src.state_dict().items()}

    def _source(self):
        m = self.model
        return m.get_ema_model() if hasattr(m, "get_ema_model") and callable(m.get_ema_model) else m

    @torch.no_grad()
    def update(self, iteration):
        if (iteration + 1) % self.update_iterval != 0:
            return
        cur, ema = self._source().state_dict(), self.ema_model.state_dict()
        fe, fc = [], []
        for k, e in ema.items():
            c = cur[k].to(self.device, non_blocking=True)
            if e.is_floating_point():
                fe.append(e)
                fc.append(c if c.dtype == e.dtype else c.to(e.dtype))
            else:
                e.copy_(c)  # the reference's `e * decay + c * (1 - decay)` on integer buffers is not meaningful; keep them in sync
        if fe:
            torch._foreach_mul_(fe, self.decay)                 # state_dict() tensors alias the shadow model's storage: in place
            torch._foreach_add_(fe, fc, alpha=1.0 - self.decay)

    def state_dict(self):
        return self.ema_model.state_dict()

    def load_state_dict(self, state_dict, strict=True):
        self.ema_model.load_state_dict({k: v.clone().to(self.device) for k, v in state_dict.items()}, strict=strict)

    @torch.no_grad()
    def modify_to_inference(self):
        """Swap the shadow weights into the live model (validation / sampling), remembering the live ones."""
        src = self._source()
        self.cur_state_dict = {k: v.clone().to(self.device) for k, v in src.state_dict().items()}
        dev = next(src.parameters()).device
        src.load_state_dict({k: v.to(dev) for k, v in self.ema_model.state_dict().items()})

    @torch.no_grad()
    def modify_to_train(self):
        src = self._source()
        dev = next(src.parameters()).device
        src.load_state_dict({k: v.clone().to(dev) for k, v in self.cur_state_dict.items()})
