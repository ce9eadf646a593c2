"""The reference's OWN host code driven over the drop-in DALLE (CPU suite, build container only: skipped where /root/reference is absent).

`Solver.sample` (reference engine/solver_spec.py:191-262) is imported unmodified and run against the drop-in `DALLE`; only I/O is stubbed
(logger) and -- because this container has no GPU -- the three compute entry points of the drop-in (`transformer.sample`, `decode_to_img`,
`get_tokens`) are bound to the CPU oracle for the duration of the test.  What is under test is the host-side contract: the call
`model.sample(batch=batch, step=...)`, the key names and tensor layouts Solver.sample consumes, the files it writes from them, the
eval()/train() toggling, and the EMA swap around the call.  (tests/test_gpu_dropin_api.py runs the same sequence on the real kernels.)"""
import math
import os
import sys
import types

import pytest
import torch

from oracle import diffsound_oracle as O
from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference is not present on this machine")


def _reference_solver_class():
    ref_harness.install_shims()
    if "torch._six" not in sys.modules:  # removed from torch long ago; the reference's lr_scheduler.py:8 imports `inf` from it
        six = types.ModuleType("torch._six")
        six.inf = math.inf
        sys.modules["torch._six"] = six
    from sound_synthesis.engine.solver_spec import Solver
    return Solver


class _Logger:
    def __init__(self):
        self.lines, self.images = [], []

    def log_info(self, msg, **kw):
        self.lines.append(str(msg))

    def add_images(self, tag, img_tensor, global_step=None, dataformats="NCHW"):
        self.images.append((tag, tuple(img_tensor.shape)))


def test_reference_solver_sample_runs_over_the_dropin_dalle(tmp_path):
    import _pkg
    _pkg.load()
    from diffsound_b200.utils import builders
    from diffsound_b200.utils.misc import instantiate_from_config, retarget_config
    Solver = _reference_solver_class()
    K, D, NL, NH, CD, B = 32, 64, 1, 1, 64, 2
    dd = dict(builders.DDCONFIG, ch=32)  # a narrow decoder keeps the CPU oracle fast; resolution / grid are the real ones
    torch.manual_seed(0)
    dalle = instantiate_from_config(retarget_config(builders.dalle_config(K, D, NL, NH, CD, ddconfig=dd))).train()
    tsd = {k: v.detach() for k, v in dalle.transformer.state_dict().items()}
    csd = {k: v.detach() for k, v in dalle.state_dict().items() if k.startswith("content_codec.")}
    calls = []

    def oracle_sample(condition_token, condition_mask, condition_embed, content_token=None, filter_ratio=0.5, **kw):
        calls.append((filter_ratio, None if content_token is None else tuple(content_token.shape), kw.get("step")))
        gen = torch.Generator().manual_seed(int(filter_ratio * 100))
        start = int(100 * filter_ratio)
        x_init = None
        if start:
            x_init = O.q_sample_ids({k: tsd[k] for k in tsd if k.startswith("log_")}, content_token,
                                    torch.full((B,), start - 1, dtype=torch.long), torch.rand(B, K + 1, 265, generator=gen), T=100, num_classes=K + 1)
        tok = O.sample(tsd, condition_embed, gen, n_layer=NL, n_head=NH, spatial=(5, 53), truncation=None,
                       steps=list(range((start or 100) - 1, -1, -1))[-6:], x_init=x_init)   # the last 6 steps keep the test short
        return {"content_token": tok.clamp(max=K - 1)}

    dalle.transformer.sample = oracle_sample
    dalle.decode_to_img = lambda index, zshape, stage="first": O.decode_to_img(csd, index)
    dalle.get_tokens = lambda spec: (torch.zeros(B, 256, 5, 53), torch.randint(0, K, (B, 265), generator=torch.Generator().manual_seed(1)))

    s = Solver.__new__(Solver)  # __init__ builds optimizers / dataloaders / tensorboard: out of scope, stubbed by direct attribute set-up
    s.logger, s.ema, s.model, s.debug = _Logger(), None, dalle, False
    s.args = types.SimpleNamespace(amp=False)
    s.last_iter, s.last_epoch = 41, 3
    s.image_dir = str(tmp_path)
    s.dataloader = {"train_iterations": 10}
    g = torch.Generator().manual_seed(2)
    batch = {"image": torch.rand(B, 1, 80, 848, generator=g) * 2 - 1, "text": ["a dog barks", "rain"], "condition_embed": torch.randn(B, 77, CD, generator=g)}
    s.sample(batch, phase="train", step_type="iteration")

    assert [c[0] for c in calls] == [0, 0.5, 1.0] and all(c[2] == 41 for c in calls) and calls[1][1] == (B, 265)
    want_imgs = {"input_image", "reconstruction_image", "cond1_cont1_fr0_image", "cond1_cont1_fr0.5_image", "cond1_cont1_fr1.0_image"}
    assert {t.split("/")[-1] for t, _ in s.logger.images} == want_imgs
    assert all(shape == (B, 1, 80, 848) for _, shape in s.logger.images)
    for k in want_imgs:  # Solver.sample wrote one image grid per key and the captions as text
        files = os.listdir(os.path.join(str(tmp_path), "train", k))
        assert any(f.endswith(".jpg") for f in files), (k, files)
    assert any(f.endswith(".txt") for f in os.listdir(os.path.join(str(tmp_path), "train", "condition")))
    assert dalle.training


def test_reference_ema_wrapper_swaps_weights_around_sample():
    """The reference's own EMA class (engine/ema.py) over the drop-in: deep copy, update, modify_to_inference / modify_to_train round trip."""
    import _pkg
    _pkg.load()
    from diffsound_b200.utils import builders
    from diffsound_b200.utils.misc import instantiate_from_config, retarget_config
    ref_harness.install_shims()
    from sound_synthesis.engine.ema import EMA
    torch.manual_seed(0)
    dalle = instantiate_from_config(retarget_config(builders.dalle_config(32, 64, 1, 1, 64, ddconfig=dict(builders.DDCONFIG, ch=32))))
    ema = EMA(dalle, decay=0.5, update_interval=1, device=torch.device("cpu"))
    assert ema.ema_model is not dalle.transformer and ema.ema_model.transformer.engine.m is ema.ema_model.transformer
    w0 = dalle.transformer.transformer.to_logits[1].weight.detach().clone()
    with torch.no_grad():
        dalle.transformer.transformer.to_logits[1].weight.add_(1.0)
    ema.update(0)
    ema.modify_to_inference()
    assert torch.allclose(dalle.transformer.transformer.to_logits[1].weight, w0 + 0.5)
    assert not dalle.transformer.transformer.engine.packed  # load_state_dict marks the packed copies stale
    ema.modify_to_train()
    assert torch.allclose(dalle.transformer.transformer.to_logits[1].weight, w0 + 1.0)
