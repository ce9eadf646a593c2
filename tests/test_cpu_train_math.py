"""A13 on the CPU: the per-column loss / gradient math that train.cu runs per warp (csrc/train_loss_math.cuh), compiled here for the
host with one serial lane, against torch autograd through the oracle restatement of _train_loss."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import diffsound_oracle as O
from tests.helpers import ROOT, portable_uniform


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lossmath") / "loss_math_host.so")
    src = os.path.join(ROOT, "tests", "native", "loss_math_host.cpp")
    inc = os.path.join(ROOT, "text-to-sound-synthesis_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", inc, src, "-o", out])
    lib = ctypes.CDLL(out)
    lib.loss_columns_host.restype = ctypes.c_int
    return lib


def _run(lib, logits_blk, x0, xt, t, g_main, g_aux, sched8, K, L, T, mw):
    B = x0.shape[0]
    dz = np.zeros((B, L, K), np.float32); prob = np.zeros((B, K + 1, L), np.float32)
    col = np.zeros((B, L, 2), np.float32); hits = np.zeros((B, L, 2), np.int32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    arrs = [np.ascontiguousarray(a) for a in (logits_blk, x0, xt, t, g_main, g_aux, sched8)]
    rc = lib.loss_columns_host(*[P(a) for a in arrs], B, K, L, T, ctypes.c_float(mw[0]), ctypes.c_float(mw[1]), P(dz), P(prob), P(col), P(hits))
    assert rc == 0
    return dz, prob, col, hits


@pytest.mark.parametrize("K,scale,mw", [(32, 2.0, (1.0, 1.0)), (256, 6.0, (1.0, 1.0)), (256, 30.0, (0.7, 1.3)), (512, 1.0, (1.0, 1.0))])
def test_column_loss_and_logit_gradient_match_autograd(host_lib, K, scale, mw):
    B, L, T = 5, 40, 100
    sched = O.schedule_buffers(T, K + 1)
    names = ["log_at", "log_bt", "log_ct", "log_1_min_ct", "log_cumprod_at", "log_cumprod_bt", "log_cumprod_ct", "log_1_min_cumprod_ct"]
    sched8 = np.stack([np.concatenate([sched[n].numpy(), np.zeros(T + 1 - len(sched[n]), np.float32)]) for n in names]).astype(np.float32)
    x0 = (portable_uniform(1, (B, L)) * K).long().clamp(max=K - 1)
    t = torch.tensor([57, 0, 99, 1, 20])
    pt = torch.tensor([0.013, 0.004, 0.01, 0.02, 0.01])
    x_t = O.q_sample_ids(sched, x0, t, portable_uniform(2, (B, K + 1, L)), T=T, num_classes=K + 1)
    out = ((portable_uniform(3, (B, K, L)) - 0.5) * scale)
    out[:, :, :5] += 12.0 * torch.nn.functional.one_hot(x0[:, :5], K).permute(0, 2, 1)  # some confident-and-right columns
    out.requires_grad_(True)
    aux_w, adaptive = 5e-4, True
    ref = O.train_loss_from_logits(sched, out, x0, x_t, t, pt, T=T, aux_weight=aux_w, adaptive_aux=adaptive, mask_weight=mw)
    ref["loss"].backward()
    g_main = (1.0 / (pt * B * L)).float()
    g_aux = g_main * aux_w * (t.float() / T + 1.0)
    dz, prob, col, hits = _run(host_lib, out.detach().permute(0, 2, 1).contiguous().numpy(), x0.numpy(), x_t.numpy(), t.numpy(),
                               g_main.numpy(), g_aux.numpy(), sched8, K, L, T, mw)
    assert np.abs(prob - ref["log_model_prob"].detach().numpy()).max() < 2e-5
    kl_loss = torch.from_numpy(col[..., 0]).sum(-1)
    assert torch.allclose(kl_loss, ref["kl_loss"].detach(), rtol=2e-5, atol=1e-5)
    assert torch.allclose(torch.from_numpy(col[..., 1]).sum(-1), ref["kl_aux_loss"].detach(), rtol=2e-5, atol=1e-5)
    loss = float(((kl_loss.double() * g_main.double()) + torch.from_numpy(col[..., 1]).sum(-1).double() * g_aux.double()).sum())
    assert abs(loss - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    gref = out.grad.permute(0, 2, 1).numpy()
    assert np.abs(dz - gref).max() <= 2e-4 * np.abs(gref).max(), (np.abs(dz - gref).max(), np.abs(gref).max())
    assert np.array_equal(hits[..., 0], (ref["x0_recon"] == x0).numpy().astype(np.int32))
    assert np.array_equal(hits[..., 1], (ref["xt_1_recon"] == x_t).numpy().astype(np.int32))
