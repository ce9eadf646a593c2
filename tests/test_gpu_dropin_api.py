"""Boundary holes closed in round 2: DALLE.sample / reconstruct, DiffusionTransformer.sample_uniform_only, stale packed weights after
in-place parameter updates, EMA deep copies after a first sample, out-of-range token ids.  All through the drop-in nn.Module API on the B200."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dalle(K=64, D=128, NL=2, NH=2, CD=64, precision=None):
    import _pkg
    _pkg.load()
    from diffsound_b200.utils import builders
    return builders.build_dalle(K=K, D=D, NL=NL, NH=NH, CD=CD, precision=precision, seed=0)


def test_dalle_sample_follows_the_solver_call_sequence():
    """Solver.sample (reference engine/solver_spec.py:191-262) calls model.sample(batch=batch, step=last_iter) under no_grad and treats every
    4-D tensor with 1 or 3 channels as an image grid; everything else is written as text.  Same call, same consumer logic, real kernels."""
    dalle = _dalle().train()
    g = torch.Generator().manual_seed(0)
    B = 2
    batch = {"image": torch.rand(B, 1, 80, 848, generator=g) * 2 - 1, "text": ["a dog barks", "rain on a roof"],
             "condition_embed": torch.randn(B, 77, 64, generator=g)}
    with torch.no_grad():
        samples = dalle.sample(batch=batch, step=7)
    assert list(samples.keys()) == ["condition", "input_image", "reconstruction_image", "cond1_cont1_fr0_image", "cond1_cont1_fr0.5_image", "cond1_cont1_fr1.0_image"]
    assert samples["condition"] == batch["text"] and samples["input_image"] is batch["image"]
    images = {k: v for k, v in samples.items() if torch.is_tensor(v) and v.dim() == 4 and v.shape[1] in (1, 3)}
    assert set(images) == set(samples) - {"condition"}
    for k, v in images.items():
        assert v.shape == (B, 1, 80, 848) and torch.isfinite(v).all(), k
    assert dalle.training  # sample() leaves the model in train mode, as the reference does (:335)
    # the reconstruction is decode(encode(mel)) and equals reconstruct()
    assert torch.equal(dalle.reconstruct(batch["image"]), samples["reconstruction_image"])
    quant_z, tok = dalle.get_tokens(batch["image"].cuda())
    assert torch.equal(dalle.decode_to_img(tok, quant_z.shape), samples["reconstruction_image"])
    # logits requested: one-hot probabilities of the final grid, (B, K+1, L)
    out = dalle.sample(batch=batch, filter_ratio=[0], return_logits=True)
    assert out["logits"].shape == (B, 65, 265) and float(out["logits"].sum(1).min()) == pytest.approx(1.0)


def test_sample_uniform_only_matches_stagewise_reference_flow():
    dalle = _dalle()
    tr = dalle.transformer
    cond = torch.randn(2, 77, 64, device="cuda")
    torch.manual_seed(11)
    a = tr.sample_uniform_only(None, None, cond, filter_ratio=0, batch_size=2)["content_token"]
    # restated reference flow (diffusion_transformer.py:700-727): CPU randint start grid, then 100 x p_sample through the stage methods
    torch.manual_seed(11)
    x = torch.randint(0, tr.num_classes - 2, (2, tr.shape)).cuda()
    from diffsound_b200.modeling.transformers.diffusion_transformer import index_to_log_onehot
    log_z = index_to_log_onehot(x, tr.num_classes)
    for ti in range(99, -1, -1):
        log_z = tr.p_sample(log_z, cond, torch.full((2,), ti, device="cuda", dtype=torch.long))
    assert torch.equal(a, log_z.argmax(1))
    b = tr.sample_uniform_only(None, None, cond, content_token=a, filter_ratio=0.5, batch_size=2)["content_token"]
    assert b.shape == a.shape and int(b.max()) < tr.num_classes - 1


def test_inplace_parameter_updates_invalidate_packed_weights():
    """optimizer.step() / p.data.copy_() do not go through load_state_dict: sampling afterwards must use the NEW weights."""
    dalle = _dalle()
    tr = dalle.transformer
    cond = torch.randn(2, 77, 64, device="cuda")
    torch.manual_seed(3)
    before = tr.sample(None, None, cond, filter_ratio=0, batch_size=2)["content_token"]
    with torch.no_grad():
        for p in tr.transformer.parameters():
            p.add_(torch.randn_like(p) * 0.05)  # in place, like an optimizer step
    torch.manual_seed(3)
    after = tr.sample(None, None, cond, filter_ratio=0, batch_size=2)["content_token"]
    fresh = _dalle().transformer
    fresh.load_state_dict(tr.state_dict())
    torch.manual_seed(3)
    want = fresh.sample(None, None, cond, filter_ratio=0, batch_size=2)["content_token"]
    assert torch.equal(after, want), "sample() after an in-place update used stale packed weights"
    assert not torch.equal(before, after)
    # a real optimizer step through the training path, then sample again
    opt = torch.optim.SGD(tr.transformer.parameters(), lr=0.5)
    batch = {"content_token": torch.randint(0, 64, (2, 265)), "condition_embed": cond}
    loss = dalle(batch=batch, return_loss=True)["loss"]
    loss.backward()
    opt.step()
    torch.manual_seed(3)
    stepped = tr.sample(None, None, cond, filter_ratio=0, batch_size=2)["content_token"]
    fresh.load_state_dict(tr.state_dict())
    torch.manual_seed(3)
    assert torch.equal(stepped, fresh.sample(None, None, cond, filter_ratio=0, batch_size=2)["content_token"])


def test_ema_deepcopy_after_first_sample_and_forward():
    """The reference EMA deep-copies the model (engine/ema.py:19); that must work after CUDA graphs / workspaces exist."""
    from diffsound_b200.engine_utils.ema import EMA
    dalle = _dalle().train()
    cond = torch.randn(2, 77, 64, device="cuda")
    dalle.transformer.sample(None, None, cond, filter_ratio=0, batch_size=2)            # captures a CUDA graph
    dalle(batch={"content_token": torch.randint(0, 64, (2, 265)), "condition_embed": cond}, return_loss=True)["loss"].backward()  # train graphs
    ema = EMA(dalle, decay=0.5)
    shadow = ema.ema_model
    assert shadow is not dalle.transformer and shadow._graphs == {} and shadow.transformer.engine.m is shadow.transformer
    cp = copy.deepcopy(dalle.transformer)
    torch.manual_seed(5)
    a = cp.sample(None, None, cond, filter_ratio=0, batch_size=2)["content_token"]
    torch.manual_seed(5)
    b = dalle.transformer.sample(None, None, cond, filter_ratio=0, batch_size=2)["content_token"]
    assert torch.equal(a, b)
    ema.update(0)
    ema.modify_to_inference()
    dalle.transformer.sample(None, None, cond, filter_ratio=0, batch_size=2)
    ema.modify_to_train()


def test_out_of_range_token_ids_raise_like_the_reference():
    dalle = _dalle()
    tr = dalle.transformer
    cond = torch.randn(1, 77, 64, device="cuda")
    bad = torch.full((1, 265), 70, dtype=torch.long, device="cuda")  # num_embed = 65
    with pytest.raises(IndexError, match="out of range"):
        tr.transformer(bad, cond, torch.zeros(1, dtype=torch.long, device="cuda"))
    with pytest.raises(AssertionError):
        dalle(batch={"content_token": torch.full((1, 265), 99, dtype=torch.long), "condition_embed": cond}, return_loss=True)
    tr.transformer(torch.zeros(1, 265, dtype=torch.long, device="cuda"), cond, torch.zeros(1, dtype=torch.long, device="cuda"))  # flag was reset
