"""The tcgen05 3xTF32 dense-layer kernel against fp64 ground truth: its error must sit at the fp32 level (the library
fp32 sgemm is measured alongside for scale), for every shape the networks use, plus ragged M, the two-operand
(concatenation-free) form, the additive input, bias and ReLU, and the autograd wrapper."""
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _err(got, ref64):
    return float((got.double() - ref64).abs().max() / ref64.abs().max())


@pytest.mark.parametrize("M,N,K1,K2", [(32768, 128, 128, 0), (1000, 384, 128, 0), (4097, 128, 384, 0), (777, 128, 128, 128),
                                       (128, 128, 32, 0), (5, 256, 64, 32)])
@pytest.mark.parametrize("relu", [False, True])
def test_gemm_matches_fp64(M, N, K1, K2, relu):
    from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K1)
    x = torch.randn(M, K1, device="cuda", generator=g) * 3
    x2 = torch.randn(M, K2, device="cuda", generator=g) if K2 else None
    W = torch.randn(N, K1 + K2, device="cuda", generator=g) * 0.2
    bias = torch.randn(N, device="cuda", generator=g)
    add = torch.randn(M, N, device="cuda", generator=g)
    assert ops._tc_ok(x, W, x2)
    got = ops.linear(x, W, bias, relu, x2=x2, add=add)
    xin = x if x2 is None else torch.cat([x, x2], 1)
    ref = xin.double() @ W.double().t() + bias.double() + add.double()
    lib = torch.addmm(bias, xin, W.t()) + add
    if relu:
        ref, lib = torch.relu(ref), torch.relu(lib)
    e_tc, e_lib = _err(got, ref), _err(lib, ref)
    assert e_tc < 2.5e-6, (e_tc, e_lib)
    print(f"M={M} N={N} K={K1 + K2}: tcgen05 3xTF32 err {e_tc:.2e}, library fp32 err {e_lib:.2e}")


def test_gemm_strided_weight_slice_and_bias_only():
    """The semantic layer feeds W[:, 4:] (row stride 388) and a precomputed additive term instead of a bias."""
    from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
    g = torch.Generator(device="cuda").manual_seed(0)
    Wfull = torch.randn(128, 388, device="cuda", generator=g) * 0.1
    x = torch.randn(3000, 384, device="cuda", generator=g)
    add = torch.randn(3000, 128, device="cuda", generator=g)
    W = Wfull[:, 4:]
    assert ops._tc_ok(x, W, None)
    got = ops.linear(x, W, None, False, add=add)
    ref = x.double() @ W.double().t() + add.double()
    assert _err(got, ref) < 2e-6


def test_linear_autograd_matches_library():
    from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
    g = torch.Generator(device="cuda").manual_seed(1)
    M, E = 2050, 128
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x, x2, W, b, add = mk(M, E), mk(M, E), mk(E, 2 * E) * 0.1, mk(E) * 0.1, mk(M, E)
    pre = torch.addmm(b, torch.cat([x, x2], 1), W.t()) + add
    add = add + (pre.abs() < 1e-3) * 1e-2        # keep every pre-activation away from the ReLU knife edge (mask parity)
    a = [t.clone().requires_grad_(True) for t in (x, x2, W, b, add)]
    out = ops.linear(a[0], a[2], a[3], True, x2=a[1], add=a[4])
    r = [t.clone().requires_grad_(True) for t in (x, x2, W, b, add)]
    ref = torch.relu(torch.addmm(r[3], torch.cat([r[0], r[1]], 1), r[2].t()) + r[4])
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    dout = mk(M, E)
    out.backward(dout)
    ref.backward(dout)
    for u, v in zip(a, r):
        torch.testing.assert_close(u.grad, v.grad, rtol=1e-4, atol=1e-5 * max(1.0, float(v.grad.abs().max())))


def test_gemm_throughput_note():
    """Not a pass/fail perf gate: records the measured rate so a regression is visible in the test log."""
    from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
    M, N, K = 32768, 384, 128
    x = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    b = torch.randn(N, device="cuda")
    for fn, name in ((lambda: ops.linear(x, W, b), "tcgen05 3xTF32"), (lambda: torch.addmm(b, x, W.t()), "library fp32")):
        for _ in range(3):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            fn()
        e.record()
        e.synchronize()
        ms = s.elapsed_time(e) / 20
        print(f"{name}: {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")


@pytest.mark.parametrize("R,N,K,ldx_extra", [(5000, 128, 128, 0), (2048, 384, 128, 0), (70001, 128, 256, 4), (33, 128, 128, 0), (4097, 256, 384, 0)])
def test_wgrad_tf32x3_matches_fp64(R, N, K, ldx_extra):
    """dW = dY^T X (split-K tensor-core kernel) against an fp64 reference; also deterministic run to run."""
    import ctypes
    from distributed_multi_agent_reinforcement_learning_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(R + N)
    dy = torch.randn(R, N, device="cuda", generator=g)
    xfull = torch.randn(R, K + ldx_extra, device="cuda", generator=g)
    x = xfull[:, :K]
    ref = (dy.double().t() @ x.double())
    ws = torch.empty(int(L.marl_wgrad_workspace_bytes(R, N, K)), dtype=torch.uint8, device="cuda")
    outs = []
    for _ in range(2):
        dW = torch.full((N, K + 8), 7.0, device="cuda")
        db = torch.full((N,), 3.0, device="cuda")
        _lib.check(L.marl_wgrad_tf32x3(R, N, K, dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dW.data_ptr(), dW.stride(0),
                                       db.data_ptr(), 0, ws.data_ptr(), _lib.stream_ptr()), "marl_wgrad_tf32x3")
        torch.cuda.synchronize()
        outs.append(dW.clone())
        bref = dy.double().sum(0)
        assert (db.double() - bref).abs().max().item() <= 2e-6 * max(1.0, bref.abs().max().item()) * (R ** 0.5)
    assert torch.equal(outs[0], outs[1])
    assert (outs[0][:, K:] == 7.0).all()                       # nothing written outside the tile columns
    err = (outs[0][:, :K].double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    fp32 = ((dy.t() @ x).double() - ref).abs().max().item()    # what the library sgemm achieves
    assert err <= max(5e-6 * scale, 4.0 * fp32), (err, fp32, scale)      # fp32-level: a few ulp of the accumulated magnitude
    # accumulate mode
    _lib.check(L.marl_wgrad_tf32x3(R, N, K, dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dW.data_ptr(), dW.stride(0), None, 1,
                                   ws.data_ptr(), _lib.stream_ptr()), "marl_wgrad_tf32x3")
    torch.testing.assert_close(dW[:, :K], 2 * outs[0][:, :K], rtol=1e-6, atol=1e-6 * scale)


@pytest.mark.parametrize("M,N,K1,K2,relu,use_add", [(1000, 128, 128, 0, True, False), (4097, 384, 128, 0, False, False), (333, 128, 128, 128, True, False),
                                                    (70000, 128, 384, 0, False, True), (260, 512, 256, 0, True, True)])
def test_rowgemm_matches_fp64(M, N, K1, K2, relu, use_add):
    """Persistent row-tile GEMM (csrc/rowgemm_tf32x3.cu) through policy_ops.linear, forward and both gradients, vs fp64."""
    from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
    g = torch.Generator(device="cuda").manual_seed(M)
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x, x2 = mk(M, K1).requires_grad_(True), (mk(M, K2).requires_grad_(True) if K2 else None)
    W, b = (mk(N, K1 + K2) / 8).requires_grad_(True), mk(N).requires_grad_(True)
    add = mk(M, N) if use_add else None
    if relu:                                      # keep pre-activations away from the ReLU knife edge (mask parity in backward)
        xin0 = x.detach() if x2 is None else torch.cat([x.detach(), x2.detach()], 1)
        pre = torch.addmm(b.detach(), xin0, W.detach().t()) + (add if add is not None else 0)
        fix = (pre.abs() < 1e-3) * 1e-2
        add = fix if add is None else add + fix
    assert ops._rowgemm_ok(M, N, K1, K2)
    out = ops.linear(x, W, b, relu=relu, x2=x2, add=add)
    xin = x.double() if x2 is None else torch.cat([x.double(), x2.double()], 1)
    ref = xin @ W.double().t() + b.double() + (add.double() if add is not None else 0)
    ref = torch.relu(ref) if relu else ref
    scale = ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() <= 4e-6 * scale      # 3-4 output tiles share one accumulator per tile
    dout = mk(M, N)
    out.backward(dout)
    gx, gW, gb = x.grad.clone(), W.grad.clone(), b.grad.clone()
    gx2 = x2.grad.clone() if x2 is not None else None
    for t in (x, W, b) + ((x2,) if x2 is not None else ()):
        t.grad = None
    ref.backward(dout.double())
    for mine, theirs in ((gx, x.grad), (gW, W.grad), (gb, b.grad)) + (((gx2, x2.grad),) if x2 is not None else ()):
        assert (mine.double() - theirs.double()).abs().max().item() <= 5e-6 * theirs.abs().max().item() + 1e-7


@pytest.mark.parametrize("R,N,K", [(1000, 128, 4), (70001, 128, 4), (333, 96, 8)])
def test_skinny_wgrad_and_relu_bwd(R, N, K):
    """Weight / bias gradient of a 4- or 8-wide input layer in one pass over dY, and the fused ReLU backward."""
    from distributed_multi_agent_reinforcement_learning_b200 import policy_ops as ops
    g = torch.Generator(device="cuda").manual_seed(R)
    p = torch.randn(R, K, device="cuda", generator=g)
    W = torch.randn(N, K + 3, device="cuda", generator=g).requires_grad_(True)
    b = torch.randn(N, device="cuda", generator=g).requires_grad_(True)
    dy = torch.randn(R, N, device="cuda", generator=g)
    out = ops.skinny_linear(p, W[:, :K], b)
    torch.testing.assert_close(out, torch.addmm(b, p, W[:, :K].t()), rtol=0, atol=0)
    out.backward(dy)
    ref_w, ref_b = dy.double().t() @ p.double(), dy.double().sum(0)
    scale = float(ref_w.abs().max())
    assert float((W.grad[:, :K].double() - ref_w).abs().max()) <= 2e-6 * max(1.0, scale) * (R ** 0.5) / 10
    assert float((b.grad.double() - ref_b).abs().max()) <= 2e-6 * max(1.0, float(ref_b.abs().max())) * (R ** 0.5) / 10
    assert not W.grad[:, K:].any()
    g1, g2 = W.grad.clone(), None
    W.grad = None
    ops.skinny_linear(p, W[:, :K], b).backward(dy)
    assert torch.equal(W.grad, g1)                                  # deterministic
    y = torch.randn(R, N, device="cuda", generator=g)
    torch.testing.assert_close(ops.relu_bwd(dy, y), dy * (y > 0), rtol=0, atol=0)
    q = torch.randn(R, 9, device="cuda", generator=g)                # the actor head's shape: [E, 9] = feat^T dlogits
    got, ref = ops.skinny_outer(dy, q), dy.double().t() @ q.double()
    assert float((got.double() - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max())) * (R ** 0.5) / 10
