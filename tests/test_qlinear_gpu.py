"""GPU: the module-level path (QuantLinear.forward = fused act-quant + tcgen05 GEMM through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import qqq_oracle as O

pytestmark = pytest.mark.gpu


def _module(p, K, N, gs, dev="cuda:0"):
    import qqq_b200

    ql = qqq_b200.QuantLinear(4, gs, K, N, bias=False)
    ql.B.copy_(torch.from_numpy(p["B"]))
    ql.s_channel.copy_(torch.from_numpy(p["s2"]))
    if gs != -1:
        ql.s_group.copy_(torch.from_numpy(p["s3"]))
    return ql.to(dev)


@pytest.mark.parametrize("M,K,N,gs", [(1, 4096, 4096, -1), (9, 512, 256, 128), (300, 1024, 384, -1)])
def test_forward_bit_exact_vs_oracle(M, K, N, gs):
    """Includes BASELINE config 1 (M=1, K=N=4096, per-channel) on the GPU path."""
    p = O.make_problem(M, K, N, gs, seed=4)
    ql = _module(p, K, N, gs)
    y = ql(torch.from_numpy(p["x"]).cuda()).cpu().numpy()
    A8, s1 = O.dynamic_quant(p["x"], cuda_semantics=True)
    ref = O.qqq_gemm_oracle(A8, p["B"], s1, p["s2"], p["s3"])
    assert np.array_equal(y.view(np.uint16), ref.view(np.uint16))


def test_forward_keeps_leading_dims_and_bias():
    p = O.make_problem(6, 256, 128, -1, seed=2)
    ql = _module(p, 256, 128, -1)
    ql.bias = torch.full((128,), 0.5, dtype=torch.half, device="cuda:0")
    x = torch.from_numpy(p["x"]).cuda().reshape(2, 3, 256)
    y = ql(x)
    assert y.shape == (2, 3, 128)
    ql.bias = None
    y0 = ql(x)
    assert torch.equal(y, y0 + 0.5)


@pytest.mark.parametrize("M,K,N,gs", [(1, 4096, 1024, 128), (33, 1024, 320, -1), (300, 2048, 384, 128), (1024, 512, 1152, -1),
                                      (32, 4096, 4096, 128), (2, 256, 64, -1)])
def test_bias_in_the_epilogue_equals_the_reference_two_step_result(M, K, N, gs):
    """`D + bias` (qlinear_marlin.py:286-288) folded into the GEMM epilogue: identical bits to the oracle's D followed by an
    fp16 add, over whole-tile and stream-K (finisher) schedules, ragged M and N % 128 == 64."""
    p = O.make_problem(M, K, N, gs, seed=M + N)
    ql = _module(p, K, N, gs)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(N)).half()
    ql.bias = bias.cuda()
    y = ql(torch.from_numpy(p["x"]).cuda())
    A8, s1 = O.dynamic_quant(p["x"], cuda_semantics=True)
    want = torch.from_numpy(O.qqq_gemm_oracle(A8, p["B"], s1, p["s2"], p["s3"])) + bias
    assert torch.equal(y.cpu().view(torch.int16), want.view(torch.int16))
    assert int(ql.workspace.abs().sum()) == 0


@pytest.mark.parametrize("gs", [-1, 128])
def test_merged_linears_bit_identical_to_separate(gs):
    import qqq_b200

    K = 1024
    ps = [O.make_problem(40, K, n, gs, seed=s) for n, s in ((256, 1), (128, 2), (384, 3))]
    x = torch.from_numpy(ps[0]["x"]).cuda()
    mods = [_module(p, K, p["B"].shape[1] // 2, gs) for p in ps]
    merged = qqq_b200.merge_quant_linears(mods)
    y = merged(x)
    parts = torch.split(y, merged.split_sizes, dim=-1)
    for m, part in zip(mods, parts):
        assert torch.equal(m(x), part)


def test_graph_capture_replays_the_same_bits():
    from qqq_b200 import graph

    p = O.make_problem(32, 1024, 512, 128, seed=6)
    ql = _module(p, 1024, 512, 128)
    x = torch.from_numpy(p["x"]).cuda()
    eager = ql(x).clone()
    g = graph.capture(lambda t: ql(t), x)
    out = g(x).clone()
    out2 = g(x * 1).clone()
    assert torch.equal(eager, out) and torch.equal(eager, out2)


def test_fork_join_branches_give_the_serial_bits_in_a_graph():
    """q/k/v-style linears on forked streams (graph.fork_join) inside a captured graph: same bits as one after the other,
    replay after replay, with the shared quantisation issued before the fork."""
    import qqq_b200
    from qqq_b200 import graph, ops

    K = 1024
    ps = [O.make_problem(48, K, n, gs, seed=s) for n, gs, s in ((256, -1, 1), (128, -1, 2), (384, -1, 3))]
    mods = [_module(p, K, p["B"].shape[1] // 2, -1) for p in ps]
    tail = _module(O.make_problem(48, 256, 128, -1, seed=9), 256, 128, -1)
    x = torch.from_numpy(ps[0]["x"]).cuda()

    def serial(t):
        qa = qqq_b200.QuantizedActivation(*ops.dynamic_quant(t))
        ys = [m(qa) for m in mods]
        return torch.cat(ys + [tail(ys[0])], dim=-1)

    def forked(t):
        qa = qqq_b200.QuantizedActivation(*ops.dynamic_quant(t))
        ys = graph.fork_join([(lambda m=m: m(qa)) for m in mods])
        return torch.cat(ys + [tail(ys[0])], dim=-1)

    want = serial(x).clone()
    assert torch.equal(forked(x), want)  # eager, on real side streams
    g = graph.capture(forked, x)
    for _ in range(3):
        assert torch.equal(g(x), want)
    x2 = x * 0.5
    assert torch.equal(g(x2), serial(x2))
    assert all(int(m.workspace.abs().sum()) == 0 for m in mods)


def test_pipelined_runner_overlaps_copies_without_mixing_steps():
    """graph.PipelinedRunner: H2D of step i+1 / D2H of step i-1 on their own streams around two alternating graphs — every
    step must deliver the result of ITS OWN input."""
    from qqq_b200 import graph

    p = O.make_problem(64, 1024, 512, -1, seed=12)
    ql = _module(p, 1024, 512, -1)
    xs = [(torch.from_numpy(p["x"]) * (0.25 * (i + 1))).half().pin_memory() for i in range(6)]
    want = [ql(x.cuda()).cpu() for x in xs]
    outs = [torch.empty(64, 512, dtype=torch.float16).pin_memory() for _ in xs]
    r = graph.PipelinedRunner(lambda t: ql(t), xs[0].cuda())
    for x, o in zip(xs, outs):
        r.step(x, o)
    r.drain()
    torch.cuda.synchronize()
    for w, o in zip(want, outs):
        assert torch.equal(w, o)
