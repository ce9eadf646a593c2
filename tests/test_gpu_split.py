"""Split-fp16 ("f16x3") building blocks of the parity-grade denoiser mode: every kernel against an fp64 reference on the B200.

The reference computes its nn.Linear / attention in fp32 (transformer_utils.py:43-58, :91-109, :248-253); these kernels reach
fp32-class accuracy on the fp16 tensor pipe by carrying every operand as an fp16 (hi | lo) pair and running three MMA passes."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    from tests import gpu_common
    return gpu_common


def _pair_value(p, C):
    return p[:, :C].double() + p[:, C:2 * C].double()


def test_split_f16_reconstructs_fp32(G):
    x = torch.randn(300, 192, device="cuda") * torch.logspace(-3, 2, 192, device="cuda")
    p = G.ops.split_f16(x)
    assert p.shape == (300, 384) and p.dtype == torch.float16
    rec = _pair_value(p, 192)
    # 22 significand bits, with an absolute floor of half an fp16 subnormal quantum (2^-25) for the lo half of small values
    assert bool(((rec - x.double()).abs() <= torch.maximum(3e-7 * x.double().abs(), torch.tensor(3.0e-8, dtype=torch.float64, device="cuda"))).all())
    ps = G.ops.split_f16(x * 1e-4, 2.0 ** 12)  # power-of-two prescale keeps small values out of fp16's subnormals
    xs = (x * 1e-4).double()
    assert bool(((_pair_value(ps, 192) / 4096 - xs).abs() <= torch.maximum(3e-7 * xs.abs(), torch.tensor(3.0e-8 / 4096, dtype=torch.float64, device="cuda"))).all())


@pytest.mark.parametrize("M,N,K", [(4240, 1024, 1024), (530, 3072, 1024), (265, 1024, 4096), (154, 2048, 512), (100, 256, 1024)])
def test_gemm_f16x3_matches_fp64(G, M, N, K):
    """out = A W^T + b (+ residual) at fp32-class accuracy; also the (hi | lo) pair output and the GELU2 epilogue."""
    ops = G.ops
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.02
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda")
    ap, wp = ops.split_f16(a), ops.split_f16(w, 2.0 ** 17)
    ref = _pair_value(ap, K) @ (_pair_value(wp, K) / 2.0 ** 17).T + b.double()
    out = ops.gemm_f16x3(ap, wp, b, alpha=2.0 ** -17)
    scale = float(ref.abs().max())
    tol = 3e-6 * max(1.0, K / 1024)  # fp32 accumulation of 3K products in the tensor core (not round-to-nearest per add): error grows with K
    print(f"gemm_f16x3 M={M} N={N} K={K}: rel err {float((out.double() - ref).abs().max()) / scale:.2e} (tol {tol:.1e})")
    assert float((out.double() - ref).abs().max()) / scale < tol
    out_r = ops.gemm_f16x3(ap, wp, b, residual=r, alpha=2.0 ** -17)
    assert float((out_r.double() - (ref + r.double())).abs().max()) / scale < tol
    pair = ops.gemm_f16x3(ap, wp, b, alpha=2.0 ** -17, split_out=True)
    assert pair.shape == (M, 2 * N) and pair.dtype == torch.float16
    assert float((_pair_value(pair, N) - ref).abs().max()) / scale < tol
    g = ops.gemm_f16x3(ap, wp, b, alpha=2.0 ** -17, gelu=True)
    ref_g = ref * torch.sigmoid(1.702 * ref)
    assert float((g.double() - ref_g).abs().max()) / scale < 1e-5  # __expf in the epilogue


def test_layernorm_split_output(G):
    ops = G.ops
    x = torch.randn(2, 265, 1024, device="cuda") * 3
    gam, bet = torch.randn(1024, device="cuda"), torch.randn(1024, device="cuda")
    ref = ops.layernorm(x, gam, bet)
    pr = ops.layernorm(x, gam, bet, split=True)
    assert pr.shape == (2, 265, 2048)
    assert float((_pair_value(pr.view(-1, 2048), 1024) - ref.view(-1, 1024).double()).abs().max()) < 2e-6
    tab = torch.randn(100, 2048, device="cuda") * 0.1
    t = torch.tensor([3, 77], device="cuda")
    ref = ops.ada_layernorm(x, tab, t)
    pr = ops.ada_layernorm(x, tab, t, split=True)
    assert float((_pair_value(pr.view(-1, 2048), 1024) - ref.view(-1, 1024).double()).abs().max()) < 2e-6


@pytest.mark.parametrize("B,H,Lq,Lk", [(2, 3, 265, 265), (2, 16, 265, 77), (1, 2, 77, 77), (3, 1, 128, 288), (1, 1, 9, 33),
                                       (10, 16, 265, 265), (19, 16, 265, 77), (37, 5, 140, 265)])  # more (batch, head) units than CTAs: multi-head slices
def test_attention_tc_split_matches_fp64(G, B, H, Lq, Lk):
    ops = G.ops
    D = H * 64
    q = torch.randn(B * Lq, D, device="cuda") * 1.5
    k = torch.randn(B * Lk, D, device="cuda") * 1.5
    v = torch.randn(B * Lk, D, device="cuda")
    qp, kp, vp = ops.split_f16(q), ops.split_f16(k), ops.split_f16(v)
    out = torch.full((B * Lq, 2 * D), float("nan"), dtype=torch.float16, device="cuda")
    ops.attention_tc_split(qp[:, :D], kp[:, :D], vp[:, :D], out[:, :D], q_lo=D, k_lo=D, v_lo=D, o_lo=D, B=B, H=H, Lq=Lq, Lk=Lk, scale=0.125)
    qd = _pair_value(qp, D).view(B, Lq, H, 64).permute(0, 2, 1, 3)
    kd = _pair_value(kp, D).view(B, Lk, H, 64).permute(0, 2, 1, 3)
    vd = _pair_value(vp, D).view(B, Lk, H, 64).permute(0, 2, 1, 3)
    ref = (torch.softmax(qd @ kd.transpose(-1, -2) * 0.125, -1) @ vd).permute(0, 2, 1, 3).reshape(B * Lq, D)
    got = _pair_value(out, D)
    assert torch.isfinite(got).all()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"attention_tc_split B={B} H={H} Lq={Lq} Lk={Lk}: rel err {err:.2e}")
    assert err < 3e-6


def test_attention_tc_split_inside_packed_buffers(G):
    """The engine's calling convention: hi halves are column views of wider buffers (qkv = [Qh Kh Vh | Ql Kl Vl])."""
    ops = G.ops
    B, H, L = 2, 16, 265
    D = H * 64
    x = torch.randn(B * L, 3 * D, device="cuda")
    qkv = ops.split_f16(x)  # (M, 6D)
    att = torch.zeros(B * L, 2 * D, dtype=torch.float16, device="cuda")
    ops.attention_tc_split(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:3 * D], att[:, :D], q_lo=3 * D, k_lo=3 * D, v_lo=3 * D, o_lo=D, B=B, H=H, Lq=L, Lk=L,
                           scale=0.125)
    full = _pair_value(qkv, 3 * D)
    sp = lambda m: m.view(B, L, H, 64).permute(0, 2, 1, 3)
    ref = (torch.softmax(sp(full[:, :D]) @ sp(full[:, D:2 * D]).transpose(-1, -2) * 0.125, -1) @ sp(full[:, 2 * D:])).permute(0, 2, 1, 3).reshape(B * L, D)
    assert float((_pair_value(att, D) - ref).abs().max() / ref.abs().max()) < 3e-6


def test_split_mode_survives_large_dynamic_range(G):
    """Range stress the fp16 containers: activations x64 and weights / 64 (and the reverse) must give the same fp32-class answer -- the
    weight prescale and the fp32 accumulator carry the exponent, not the fp16 operands."""
    ops = G.ops
    M, N, K = 530, 1024, 1024
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.02
    ref = a.double() @ w.double().T
    for sa, sw in ((64.0, 1 / 64.0), (1 / 64.0, 64.0), (512.0, 1.0)):
        wq = w * sw
        s = 13 - math.frexp(float(wq.abs().max()))[1]
        out = ops.gemm_f16x3(ops.split_f16(a * sa), ops.split_f16(wq, 2.0 ** s), alpha=2.0 ** -s)
        err = float((out.double() / (sa * sw) - ref).abs().max() / ref.abs().max())
        assert err < 3e-6, (sa, sw, err)


def test_gemm_two_operands_dual_output_and_column_groups(G):
    """The epilogue / producer features the MelGAN state-buffer path relies on, each against fp64:
    (1) A2: one GEMM over two activation buffers (shortcut(x) + conv1x1(y)), in place into rows of [raw pair | LeakyReLU pair];
    (2) out_col_group: polyphase ConvTranspose1d columns (phase * Cout + c) scattered into (r*T, 4*Cout) state rows;  (3) batch + pad rows."""
    import math
    ops = G.ops
    from diffsound_b200.packing import PackedConv
    B, T, C, P = 2, 300, 64, 9
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, T, C, device="cuda", generator=g)
    y = torch.randn(B, T, C, device="cuda", generator=g)
    ws, w1 = torch.randn(C, C, device="cuda", generator=g) * 0.1, torch.randn(C, C, device="cuda", generator=g) * 0.1
    bias = torch.randn(C, device="cuda", generator=g)
    S = torch.zeros(B, T + 2 * P, 4 * C, dtype=torch.float16, device="cuda")
    S[:, P:P + T, :2 * C] = ops.split_f16(x.view(-1, C)).view(B, T, 2 * C)
    Y = ops.split_f16(y.view(-1, C)).view(B, T, 2 * C).contiguous()
    cv = PackedConv([ws, w1], bias)
    xv = (S[:, P:P + T, :C].double() + S[:, P:P + T, C:2 * C].double())
    yv = (Y[..., :C].double() + Y[..., C:].double())
    wsv = (cv.w[:, :C].double() + cv.w[:, cv.Kp:cv.Kp + C].double()) * cv.alpha
    w1v = (cv.w[:, 2 * cv.Kp:2 * cv.Kp + C].double() + cv.w[:, 3 * cv.Kp:3 * cv.Kp + C].double()) * cv.alpha
    ref = xv @ wsv.T + yv @ w1v.T + bias.double()
    ld = 4 * C
    ops.gemm_desc(A=S.data_ptr(), A2=Y.data_ptr(), W=cv.w.data_ptr(), out=S.data_ptr() + 2 * P * ld, M=T, N=C, K=cv.Kp, batch=B,
                  taps=cv.taps([(P, 0, C, 0), (0, 0, C, 1)]), a_rows=T + 2 * P, a_cols=ld, lda=ld, a_batch_stride=(T + 2 * P) * ld, lda2=2 * C, a2_rows=T,
                  a2_cols=2 * C, a2_batch_stride=T * 2 * C, ldw=cv.w.shape[1], w_cols=cv.w.shape[1], ldo=ld, out_batch_stride=(T + 2 * P) * ld, bias=cv.bias,
                  flags=ops.OUT_F16_SPLIT | ops.DUAL_LRELU, alpha=cv.alpha, split_off=C, dual_off=2 * C)
    raw = S[:, P:P + T, :C].double() + S[:, P:P + T, C:2 * C].double()
    act = S[:, P:P + T, 2 * C:3 * C].double() + S[:, P:P + T, 3 * C:].double()
    sc = float(ref.abs().max())
    assert float((raw - ref).abs().max()) / sc < 3e-6
    assert float((act - torch.nn.functional.leaky_relu(ref, 0.2)).abs().max()) / sc < 3e-6
    assert float(S[:, :P].abs().max()) == 0.0 and float(S[:, P + T:].abs().max()) == 0.0  # pad rows untouched
    # (2) column groups: N = r * Cout logical columns -> r consecutive state rows of 4*Cout halves
    r_, Cout = 4, 32
    wq = torch.randn(r_ * Cout, C, device="cuda", generator=g) * 0.1
    cq = PackedConv([wq], torch.zeros(r_ * Cout, device="cuda"))
    Sn = torch.zeros(B, r_ * T + 2 * P, 4 * Cout, dtype=torch.float16, device="cuda")
    Ain = ops.split_f16(x.view(-1, C)).view(B, T, 2 * C).contiguous()
    ops.gemm_desc(A=Ain.data_ptr(), W=cq.w.data_ptr(), out=Sn.data_ptr() + 2 * P * 4 * Cout, M=T, N=r_ * Cout, K=cq.Kp, batch=B, taps=cq.taps([(0, 0, C, 0)]),
                  a_rows=T, a_cols=2 * C, lda=2 * C, a_batch_stride=T * 2 * C, ldw=cq.w.shape[1], w_cols=cq.w.shape[1], ldo=r_ * 4 * Cout,
                  out_batch_stride=(r_ * T + 2 * P) * 4 * Cout, bias=cq.bias, flags=ops.OUT_F16_SPLIT | ops.DUAL_LRELU, alpha=cq.alpha, split_off=Cout,
                  dual_off=2 * Cout, out_col_group=Cout, out_col_group_stride=4 * Cout)
    wqv = (cq.w[:, :C].double() + cq.w[:, cq.Kp:cq.Kp + C].double()) * cq.alpha
    refq = ((Ain[..., :C].double() + Ain[..., C:].double()) @ wqv.T).view(B, T, r_, Cout).reshape(B, r_ * T, Cout)
    got = Sn[:, P:P + r_ * T, :Cout].double() + Sn[:, P:P + r_ * T, Cout:2 * Cout].double()
    assert float((got - refq).abs().max()) / float(refq.abs().max()) < 3e-6
    gact = Sn[:, P:P + r_ * T, 2 * Cout:3 * Cout].double() + Sn[:, P:P + r_ * T, 3 * Cout:].double()
    assert float((gact - torch.nn.functional.leaky_relu(refq, 0.2)).abs().max()) / float(refq.abs().max()) < 3e-6


@pytest.mark.parametrize("C,T,B", [(64, 30000, 2), (32, 1500, 2), (32, 40000, 3), (24, 333, 1)])
def test_resident_w_conv_kernel_matches_generic_and_fp64(G, C, T, B):
    """dsb_gemm_ex(resident_w=1): the narrow-channel MelGAN form (all W boxes resident in shared memory, shared A boxes, N <= 128) must give the
    SAME bits as the generic kernel on the same tap list, and fp32-class agreement with fp64 -- for a dilated 3-tap conv (LeakyReLU + pair output,
    32-channel rows folded to [hi | lo] . [Wh | Wh], [Wl | 0]), the in-place two-operand ResnetBlock tail (dual output), and a 7-tap N = 1 conv
    with tanh into fp32 (reference vocoder/modules.py:72-85, :121-126)."""
    ops = G.ops
    from diffsound_b200.packing import PackedConv
    P, d = 9, 3
    Cs = (C + 7) // 8 * 8
    ld = 4 * Cs
    fold = Cs == 32
    g = torch.Generator(device="cuda").manual_seed(C + T)
    x = torch.randn(B, T + 2 * P, C, device="cuda", generator=g)
    S0 = torch.zeros(B, T + 2 * P, ld, dtype=torch.float16, device="cuda")
    pr = ops.split_f16(x.view(-1, C)).view(B, T + 2 * P, 2 * C)
    S0[..., 2 * Cs:2 * Cs + C], S0[..., 3 * Cs:3 * Cs + C] = pr[..., :C], pr[..., C:]  # act pair (pad rows filled too: the "reflected" halo)
    S0[..., :C], S0[..., Cs:Cs + C] = pr[..., :C], pr[..., C:]                          # raw pair
    xv = pr[..., :C].double() + pr[..., C:].double()
    wd = [torch.randn(C, C, device="cuda", generator=g) * 0.1 for _ in range(3)]
    bias = torch.randn(C, device="cuda", generator=g)
    g1 = PackedConv(wd, bias, fold=fold)

    def wval(cv, j, cin):
        return (cv.w[:, j * 2 * cv.Kp:j * 2 * cv.Kp + cin].double() + cv.w[:, j * 2 * cv.Kp + cv.Kp:j * 2 * cv.Kp + cv.Kp + cin].double()) * cv.alpha

    spatial = [(P + (j - 1) * d, 2 * Cs, 3 * Cs, 0) for j in range(3)]
    t64 = g1.taps64(spatial)
    assert g1.resident_ok(len(t64))
    outs = []
    for res in (0, 1):
        Y = torch.zeros(B, T, 2 * Cs, dtype=torch.float16, device="cuda")
        ops.gemm_desc(A=S0.data_ptr(), W=g1.w.data_ptr(), out=Y.data_ptr(), M=T, N=C, K=64, batch=B, taps=t64, a_rows=T + 2 * P, a_cols=ld, lda=ld,
                      a_batch_stride=(T + 2 * P) * ld, ldw=g1.w.shape[1], w_cols=g1.w.shape[1], ldo=2 * Cs, out_batch_stride=T * 2 * Cs, bias=g1.bias,
                      flags=ops.OUT_F16_SPLIT | ops.LRELU, alpha=g1.alpha, split_off=Cs, resident_w=res, block_n=0 if res else 128)  # block_n: the generic kernel, not the fused pair form
        outs.append(Y)
    assert torch.equal(outs[0], outs[1])
    ref = sum(xv[:, P + (j - 1) * d:P + (j - 1) * d + T] @ wval(g1, j, C).T for j in range(3)) + bias.double()
    ref = torch.nn.functional.leaky_relu(ref, 0.2)
    got = outs[1][..., :C].double() + outs[1][..., Cs:Cs + C].double()
    assert float((got - ref).abs().max()) / float(ref.abs().max()) < 3e-6
    # in-place two-operand tail with dual output
    ws, w1 = torch.randn(C, C, device="cuda", generator=g) * 0.1, torch.randn(C, C, device="cuda", generator=g) * 0.1
    g2 = PackedConv([ws, w1], bias, fold=fold)
    t2 = g2.taps64([(P, 0, Cs, 0), (0, 0, Cs, 1)])
    Yin = outs[1]
    res_states = []
    for res in (0, 1):
        S = S0.clone()
        ops.gemm_desc(A=S.data_ptr(), A2=Yin.data_ptr(), W=g2.w.data_ptr(), out=S.data_ptr() + 2 * P * ld, M=T, N=C, K=64, batch=B, taps=t2,
                      a_rows=T + 2 * P, a_cols=ld, lda=ld, a_batch_stride=(T + 2 * P) * ld, lda2=2 * Cs, a2_rows=T, a2_cols=2 * Cs, a2_batch_stride=T * 2 * Cs,
                      ldw=g2.w.shape[1], w_cols=g2.w.shape[1], ldo=ld, out_batch_stride=(T + 2 * P) * ld, bias=g2.bias,
                      flags=ops.OUT_F16_SPLIT | ops.DUAL_LRELU, alpha=g2.alpha, split_off=Cs, dual_off=2 * Cs, resident_w=res)
        res_states.append(S)
    assert torch.equal(res_states[0], res_states[1])
    ref2 = xv[:, P:P + T] @ wval(g2, 0, C).T + got @ wval(g2, 1, C).T + bias.double()
    S = res_states[1]
    raw = S[:, P:P + T, :C].double() + S[:, P:P + T, Cs:Cs + C].double()
    act = S[:, P:P + T, 2 * Cs:2 * Cs + C].double() + S[:, P:P + T, 3 * Cs:3 * Cs + C].double()
    sc = float(ref2.abs().max())
    assert float((raw - ref2).abs().max()) / sc < 3e-6
    assert float((act - torch.nn.functional.leaky_relu(ref2, 0.2)).abs().max()) / sc < 3e-6
    assert torch.equal(S[:, :P], S0[:, :P]) and torch.equal(S[:, P + T:], S0[:, P + T:])
    # 7 taps, N = 1, tanh, fp32 output
    wl = [torch.randn(1, C, device="cuda", generator=g) * 0.05 for _ in range(7)]
    cl = PackedConv(wl, torch.randn(1, device="cuda", generator=g) * 0.1, fold=fold)
    t7 = cl.taps64([(P - 3 + j, 2 * Cs, 3 * Cs, 0) for j in range(7)])
    wavs = []
    for res in (0, 1):
        wav = torch.zeros(B, T, 1, dtype=torch.float32, device="cuda")
        ops.gemm_desc(A=S0.data_ptr(), W=cl.w.data_ptr(), out=wav.data_ptr(), M=T, N=1, K=64, batch=B, taps=t7, a_rows=T + 2 * P, a_cols=ld, lda=ld,
                      a_batch_stride=(T + 2 * P) * ld, ldw=cl.w.shape[1], w_cols=cl.w.shape[1], ldo=1, out_batch_stride=T, bias=cl.bias, flags=ops.TANH,
                      alpha=cl.alpha, resident_w=res, block_n=0 if res else 128)
        wavs.append(wav)
    assert torch.equal(wavs[0], wavs[1])
    ref7 = torch.tanh(sum(xv[:, P - 3 + j:P - 3 + j + T] @ wval(cl, j, C).T for j in range(7)) + cl.bias.double())
    assert float((wavs[1].double() - ref7).abs().max()) < 1e-5  # absolute, after tanh: pre-activations reach ~3 over 7 x C products


@pytest.mark.parametrize("B,T", [(2, 5000), (1, 515), (3, 217088 // 8)])
def test_conv_out_pair_matches_fp64(G, B, T):
    """dsb_conv_out_pair: MelGAN's Conv1d(32 -> 1, k=7) + tanh on the FMA pipe, off the activated (hi | lo) pair columns of a state buffer
    (reference vocoder/modules.py:121-126), vs fp64; also against the tensor-core form of the same layer (dsb_gemm_ex, N = 1)."""
    ops = G.ops
    P, C = 9, 32
    g = torch.Generator(device="cuda").manual_seed(T)
    x = torch.randn(B, T + 2 * P, C, device="cuda", generator=g) * 40.0
    S = torch.zeros(B, T + 2 * P, 4 * C, dtype=torch.float16, device="cuda")
    pr = ops.split_f16(x.view(-1, C)).view(B, T + 2 * P, 2 * C)
    S[..., 2 * C:] = pr
    S[..., :2 * C] = 7.0  # raw columns: must not be read
    xv = pr[..., :C].double() + pr[..., C:].double()
    w = torch.randn(7, C, device="cuda", generator=g) * 0.01
    bias = torch.randn(1, device="cuda", generator=g) * 0.1
    scale = 2.0 ** -3
    out = ops.conv_out_pair(S, T, P - 3, 2 * C, w, bias, scale)
    ref = torch.tanh(sum(xv[:, P - 3 + j:P - 3 + j + T] @ w[j].double() for j in range(7)) * scale + bias.double())
    assert out.shape == (B, T)
    assert float((out.double() - ref).abs().max()) < 2e-6
