"""GPU numerics of the fused tcgen05 MLP-chain kernel (through the C ABI).

Two references, both plain PyTorch fp32:
  * the kernel's own numeric model (operands and inter-layer activations rounded to bf16, fp32 accumulation):
    must agree to MODEL_RTOL -- this checks descriptors / layouts / pooling exactly;
  * the reference's fp32 arithmetic (what pn_kit's Conv2d stacks compute): must agree to BF16_RTOL of the
    output scale -- the stated bf16-vs-fp32 tolerance of north_star.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

MODEL_RTOL = 2e-3   # vs the bf16-operand / fp32-accumulate model (differences: accumulation order, 1-ulp bf16 ties)
BF16_RTOL = 3e-2    # vs pure fp32, relative to the output's max magnitude


@pytest.fixture(scope="module")
def mlp():
    import __graft_entry__  # noqa: F401
    from pcc_b200 import mlp_ops
    return mlp_ops


def make_layers(dims, relu, seed):
    g = torch.Generator().manual_seed(seed)
    layers = []
    for (ci, co), r in zip(zip(dims[:-1], dims[1:]), relu):
        bound = 1.0 / ci ** 0.5
        w = ((torch.rand((co, ci), generator=g) * 2 - 1) * bound).cuda()
        b = ((torch.rand((co,), generator=g) * 2 - 1) * bound).cuda()
        layers.append((w, b, r))
    return layers


def ref_chain(x, layers, group, model_bf16, first_fp32=None):
    """model_bf16: the kernel's numeric model -- operands, bias and inter-layer activations rounded to bf16, fp32
    accumulation; a first layer with <= 7 input channels (multi-layer chain, one fp32 input) runs in plain fp32."""
    rnd = (lambda t: t.bfloat16().float()) if model_bf16 else (lambda t: t)
    h = x.double() if not model_bf16 else x
    if first_fp32 is None:
        first_fp32 = len(layers) >= 2 and layers[0][0].shape[1] <= 7
    for i, (w, b, r) in enumerate(layers):
        if model_bf16 and i == 0 and first_fp32:
            h = h @ w.t() + b
        elif model_bf16:
            h = rnd(h) @ rnd(w).t() + rnd(b)  # the bias is folded into the MMA as a bf16 column
        else:
            h = h @ w.double().t() + b.double()
        if r:
            h = torch.relu(h)
    h = h.float()
    if group > 1:
        h = h.view(-1, group, h.shape[1]).max(dim=1)[0]
    return h


CASES = [
    # dims, relu, rows, group
    ([16, 32], [False], 128, 0),
    ([3, 32], [True], 256, 0),
    ([64, 128], [True], 384, 0),
    ([40, 200], [False], 300, 0),                       # ragged K, two M tiles, tail rows
    ([3, 32, 64, 128], [True, True, True], 4096, 16),   # pn_kit.SetAbstraction body (AE.py:16)
    ([3, 32, 64, 128], [True, True, True], 1000, 0),
    ([131, 128, 256], [True, True], 1024, 0),           # first half of pn_kit.PointNet (AE.py:17)
    ([144, 128, 64, 32, 3], [True, True, True, False], 640, 0),  # pn_kit.MLP decoder tail (AE.py:27)
    ([6, 64, 64, 128], [True, True, True], 2048, 32),   # PointNet++ SA1-like (nsample 32)
    ([19, 64, 128], [True, True], 2048, 64),
    ([19, 64, 128], [True, False], 1024, 128),
    ([35, 64, 16], [True, False], 1024, 256),           # max over 256 points spanning two tiles
    ([35, 64, 16], [True, False], 2048, 512),
    ([20, 100, 48], [True, True], 512, 0),              # Cout not a multiple of 16: ones channel inside a TMEM chunk
    ([64, 128, 200], [True, False], 512, 0),            # ragged last Cout, vector fp32 store path
    ([30, 300, 8], [True, False], 384, 0),              # Cout > 256: two MMA N chunks
    ([30, 64, 300], [True, True], 1024, 64),            # pooled last layer with three M tiles
]


@pytest.mark.parametrize("dims,relu,rows,group", CASES)
def test_fused_chain_numerics(mlp, dims, relu, rows, group):
    torch.manual_seed(rows + len(dims))
    layers = make_layers(dims, relu, seed=sum(dims))
    x = (torch.rand(rows, dims[0], device="cuda") - 0.5) * 2
    y = mlp.fused_chain(x, layers, group)
    torch.cuda.synchronize()
    ym = ref_chain(x, layers, group, model_bf16=True)
    yf = ref_chain(x, layers, group, model_bf16=False)
    assert y.shape == ym.shape
    scale = yf.abs().max().item() + 1e-12
    err_model = (y - ym).abs().max().item() / scale
    err_fp32 = (y - yf).abs().max().item() / scale
    assert err_model < MODEL_RTOL, f"vs bf16 model: {err_model}"
    assert err_fp32 < BF16_RTOL, f"vs fp32: {err_fp32}"


def test_row_stride_and_repeat_launches(mlp):
    layers = make_layers([3, 32, 64, 128], [True, True, True], seed=1)
    big = torch.rand(512, 8, device="cuda")
    y1 = mlp.fused_chain(big[:, :3].contiguous(), layers, 16)
    for _ in range(3):
        y2 = mlp.fused_chain(big[:, :3], layers, 16)  # a strided view: row stride 8, 3 channels
    assert torch.equal(y1, y2)


def test_input_segments_bf16_and_broadcast(mlp):
    """cat((xyz, feat)) of AE.py:39 and cat((features, tiled latent)) of AE.py:50-51 as input segments."""
    torch.manual_seed(3)
    layers = make_layers([144, 128, 64, 32, 3], [True, True, True, False], seed=9)
    k = 128
    lin = (torch.rand(5 * k, 128, device="cuda") - 0.5).bfloat16()
    lat = torch.randint(-3, 4, (5, 16), device="cuda").float()
    y = mlp.fused_chain([(lin, 1), (lat, k)], layers)
    x = torch.cat((lin.float(), lat.repeat_interleave(k, dim=0)), dim=1)
    ym = ref_chain(x, layers, 0, model_bf16=True)
    assert (y - ym).abs().max().item() / ym.abs().max().item() < MODEL_RTOL
    # bf16 output + group max, fp32 segment that is not 8-aligned comes second
    layers2 = make_layers([131, 128, 256], [True, True], seed=10)
    feat = (torch.rand(512, 128, device="cuda") - 0.5).bfloat16()
    xyz = torch.rand(512, 3, device="cuda") - 0.5
    yb = mlp.fused_chain([(feat, 1), (xyz, 1)], layers2, group=0, out_dtype=torch.bfloat16)
    ym2 = ref_chain(torch.cat((feat.float(), xyz), dim=1), layers2, 0, model_bf16=True)
    assert yb.dtype == torch.bfloat16
    assert (yb.float() - ym2).abs().max().item() / ym2.abs().max().item() < 1e-2  # one extra bf16 rounding of the output


@pytest.mark.parametrize("patches", [149, 301, 2048])
def test_decoder_chain_two_tiles_in_flight(mlp, patches):
    """More tiles than SMs: the decoder chain runs as dec_chain2_kernel (two independent tile chains per CTA, X2 / X3 written over
    X1's slabs).  Against the bf16-operand model, and -- same MMAs, same epilogue arithmetic -- bit-identical to the one-tile
    kernel, which a fresh process selects with PCC_DEC_SLOTS=1 (odd tile counts: the two slots of a CTA get different numbers of
    tiles, some none)."""
    import subprocess, sys, os, tempfile
    layers = make_layers([144, 128, 64, 32, 3], [True, True, True, False], seed=9)
    k = 128
    g = torch.Generator(device="cuda").manual_seed(patches)
    lin = (torch.rand(patches * k, 128, device="cuda", generator=g) - 0.5).bfloat16()
    lat = torch.randint(-3, 4, (patches, 16), device="cuda", generator=g).float()
    y = mlp.fused_chain([(lin, 1), (lat, k)], layers)
    x = torch.cat((lin.float(), lat.repeat_interleave(k, dim=0)), dim=1)
    ym = ref_chain(x, layers, 0, model_bf16=True)
    assert (y - ym).abs().max().item() / ym.abs().max().item() < MODEL_RTOL
    if patches == 301:
        with tempfile.TemporaryDirectory() as d:
            torch.save({"lin": lin.cpu(), "lat": lat.cpu(), "layers": [(w.cpu(), b.cpu(), r) for w, b, r in layers]}, os.path.join(d, "in.pt"))
            code = ("import sys, torch; sys.path[:0] = [%r, %r]; import __graft_entry__; from pcc_b200 import mlp_ops; "
                    "t = torch.load(%r); layers = [(w.cuda(), b.cuda(), r) for w, b, r in t['layers']]; "
                    "y = mlp_ops.fused_chain([(t['lin'].cuda(), 1), (t['lat'].cuda(), 128)], layers); torch.save(y.cpu(), %r)"
                    % (os.getcwd(), os.path.join(os.getcwd(), "point-cloud-compression_b200"), os.path.join(d, "in.pt"), os.path.join(d, "out.pt")))
            subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, PCC_DEC_SLOTS="1"), timeout=300)
            assert torch.equal(torch.load(os.path.join(d, "out.pt")), y.cpu())


def test_wide_fp32_segment_single_layer_and_run_chain_planner(mlp):
    """PointnetSAModule SA3-like: [gathered features fp32 256ch | xyz 3ch] -> 256 -> 256 -> 512 -> 1024, max over 128.
    The planner runs the layers that fit as fused launches and the rest on the streamed tcgen05 GEMM."""
    torch.manual_seed(5)
    rows = 128 * 16
    feat = torch.rand(rows, 256, device="cuda") - 0.5
    xyz = torch.rand(rows, 3, device="cuda") - 0.5
    layers1 = make_layers([259, 256], [True], seed=21)
    y1 = mlp.fused_chain([(feat, 1), (xyz, 1)], layers1, 0, torch.bfloat16)
    ym1 = ref_chain(torch.cat((feat, xyz), 1), layers1, 0, model_bf16=True)
    assert not torch.isnan(y1.float()).any()
    assert (y1.float() - ym1).abs().max().item() / ym1.abs().max().item() < 1e-2
    layers = make_layers([259, 256, 256, 512, 1024], [True] * 4, seed=22)
    y = mlp.run_chain([(feat, 1), (xyz, 1)], layers, group=128)
    yf = ref_chain(torch.cat((feat, xyz), 1), layers, 128, model_bf16=False)
    assert y.shape == (16, 1024)
    assert (y - yf).abs().max().item() / yf.abs().max().item() < BF16_RTOL


def test_unsupported_chain_is_an_error_not_a_fallback(mlp):
    layers = make_layers([256, 512, 1024], [True, True], seed=2)
    x = torch.rand(128, 256, device="cuda")
    with pytest.raises((ValueError, RuntimeError)):
        mlp.fused_chain(x, layers, 0)


@pytest.mark.parametrize("patches,cout", [(1, 16), (5, 16), (300, 16), (3, 7)])
def test_pn_tail_fused_kernel(mlp, patches, cout):
    """pn_kit.PointNet tail 256 -> 512 (ReLU) -> d, max over the patch's 256 points, in one launch (csrc/pn_tail.cu)."""
    torch.manual_seed(patches)
    layers = make_layers([256, 512, cout], [True, False], seed=31)
    x = ((torch.rand(patches * 256, 256, device="cuda") - 0.3)).bfloat16()
    assert mlp.pn_tail_supported(x, layers, 256)
    y = mlp.pn_tail(x, layers)
    torch.cuda.synchronize()
    rnd = lambda t: t.bfloat16().float()  # noqa: E731
    (w2, b2, _), (w3, b3, _) = layers
    h = torch.relu(x.float() @ rnd(w2).t() + b2)                   # fp32 bias in the epilogue of the streamed layer
    ym = (rnd(h) @ rnd(w3).t() + rnd(b3)).view(patches, 256, cout).max(dim=1)[0]
    yf = (torch.relu(x.double() @ w2.double().t() + b2.double()) @ w3.double().t() + b3.double()).view(patches, 256, cout).max(dim=1)[0]
    scale = yf.abs().max().item() + 1e-12
    assert y.shape == ym.shape
    assert (y - ym).abs().max().item() / scale < MODEL_RTOL
    assert (y - yf.float()).abs().max().item() / scale < BF16_RTOL


# ---- streamed tensor-core GEMM layer (csrc/gemm_ws.cu) and the fused grouping pass ------------------------------------
LINEAR_CASES = [
    # rows, cin, cout, relu, group
    (1000, 64, 128, True, 0),        # tail rows (M % 128 != 0), one K slab
    (4096, 131, 128, True, 0),       # PointNet++ SA2 first layer: cin padded 131 -> 192
    (5000, 256, 256, False, 0),      # no ReLU, BN = 256, tail rows
    (2048, 1024, 512, True, 0),      # deep K: ring wraps many times
    (2048, 128, 128, True, 32),      # pooled, one warp per output row
    (4096, 128, 256, True, 64),      # pooled over two warps (atomicMax path)
    (8192, 512, 1024, True, 128),    # PointNet++ SA3 last layer (PPPF_AE.py:33), pooled over the tile
    (4096, 259, 256, True, 0),       # SA3 first layer
    (1024, 64, 384, True, 256),      # N = 3 x 128 (BN = 128), group spanning two tiles
    (300 * 128, 1024, 2048, True, 0),  # many tiles per CTA: accumulator double buffering, staging reuse
]


@pytest.mark.parametrize("rows,cin,cout,relu,group", LINEAR_CASES)
def test_streamed_linear_numerics(mlp, rows, cin, cout, relu, group):
    g = torch.Generator(device="cuda").manual_seed(rows + cin + cout)
    kp = (cin + 63) // 64 * 64
    x = torch.zeros(rows, kp, device="cuda", dtype=torch.bfloat16)
    x[:, :cin] = (torch.rand(rows, cin, device="cuda", generator=g) - 0.5) * 2
    w = (torch.rand(cout, cin, device="cuda", generator=g) - 0.5) * (2.0 / cin ** 0.5)
    b = (torch.rand(cout, device="cuda", generator=g) - 0.5) * 0.2
    y = mlp.linear(x, w, b, relu, group)
    torch.cuda.synchronize()
    ref = x[:, :cin].float() @ w.to(torch.bfloat16).float().t() + b          # bf16 operands, fp32 accumulation
    if relu:
        ref = torch.relu(ref)
    if group > 1:
        ref = ref.view(rows // group, group, cout).max(dim=1)[0]
        assert y.dtype == torch.float32 and y.shape == ref.shape
        tol = 2e-3   # fp32 out: only the accumulation order differs
    else:
        assert y.dtype == torch.bfloat16 and y.shape == ref.shape
        tol = 6e-3   # + one bf16 rounding of the output (2^-8 relative)
    scale = ref.abs().max().item() + 1e-12
    assert (y.float() - ref).abs().max().item() / scale < tol


def test_streamed_linear_rejects_unsupported_shapes(mlp):
    x = torch.zeros(256, 64, device="cuda", dtype=torch.bfloat16)
    w, b = torch.zeros(100, 64, device="cuda"), torch.zeros(100, device="cuda")
    assert mlp.linear(x, w, b, True).shape == (256, 100)   # cout is padded to the 128-column granule and sliced
    with pytest.raises(ValueError):
        mlp.linear(x, w, b, True, group=48)             # a pooling group the kernel does not take
    w, b = torch.zeros(128, 64, device="cuda"), torch.zeros(128, device="cuda")
    with pytest.raises(ValueError):
        mlp.linear(x, w, b, False, group=64)            # pooling over > 32 rows without the ReLU
    with pytest.raises(ValueError):
        mlp.linear(x[:, :48], w[:, :48], b, True)       # K not padded to 64


def test_gather_concat_bf16_matches_torch(mlp):
    g = torch.Generator(device="cuda").manual_seed(5)
    B, N, C, M = 3, 500, 128, 777
    feat = torch.rand(B, N, C, device="cuda", generator=g)
    xyz = torch.rand(B, N, 3, device="cuda", generator=g)
    idx = torch.randint(0, N, (B, M), device="cuda", generator=g)
    idx[0, :5] = -1                                     # ball-query padding reads point 0 (pointnet_sa_module.py:27)
    out = mlp.gather_concat_bf16(feat, xyz, idx, 192)
    ci = idx.clamp(min=0)
    ref = torch.zeros(B, M, 192, device="cuda")
    ref[:, :, :C] = torch.gather(feat, 1, ci[:, :, None].expand(-1, -1, C))
    ref[:, :, C:C + 3] = torch.gather(xyz, 1, ci[:, :, None].expand(-1, -1, 3))
    assert torch.equal(out.view(B, M, 192), ref.to(torch.bfloat16))
    out2 = mlp.gather_concat_bf16(None, xyz, idx, 64)   # xyz only
    assert torch.equal(out2.view(B, M, 64)[:, :, :3], ref[:, :, C:C + 3].to(torch.bfloat16)) and not out2.view(B, M, 64)[:, :, 3:].any()


@pytest.mark.parametrize("BS,P", [(64, 256), (5, 128), (3, 16), (300, 256)])
def test_sa_chain_indexed_is_bit_identical_to_the_grouped_route(BS, P):
    """pcc_sa_chain_indexed (patch + byte kNN table in, neighbours gathered and recentred inside the kernel) against the general
    route (pcc_knn_f32 writes the recentred [BS, P, 16, 3] tensor, pcc_mlp_chain reads it): the same fp32 subtraction feeds the
    same kernel body, so the features must be bit-identical, fp32 and bf16 outputs alike (pn_kit.py:190-207)."""
    import __graft_entry__  # noqa: F401
    from pcc_b200 import mlp_ops, ops
    torch.manual_seed(BS)
    x = (torch.rand(BS, P, 3, device="cuda") - 0.5) * 1.3
    layers = [(torch.randn(32, 3, device="cuda") * 0.5, torch.randn(32, device="cuda") * 0.1, True),
              (torch.randn(64, 32, device="cuda") * 0.2, torch.randn(64, device="cuda") * 0.1, True),
              (torch.randn(128, 64, device="cuda") * 0.2, torch.randn(128, device="cuda") * 0.1, True)]
    assert mlp_ops.sa_indexed_supported(P, 16, layers)
    _, _, grouped = ops.knn(x, x, 16, return_nn=True, centre_sub=True, nn_only=True)
    idx8 = ops.knn_patch_u8(x, 16)
    for dt in (torch.float32, torch.bfloat16):
        want = mlp_ops.fused_chain(grouped.reshape(BS * P * 16, 3), layers, group=16, out_dtype=dt)
        got = mlp_ops.sa_chain_indexed(x, idx8, layers, out_dtype=dt)
        assert got.shape == want.shape == (BS * P, 128) and torch.equal(got, want)
    assert not mlp_ops.sa_indexed_supported(P, 8, layers) and not mlp_ops.sa_indexed_supported(300, 16, layers)


@pytest.mark.parametrize("M,Na,Nb", [(64, 128, 128), (8192, 128, 64), (100_000, 64, 32), (1_000_003 // 8 * 8, 128, 64), (2048, 1024, 256),
                                     (5000, 16, 144), (333 * 8, 264, 200)])
def test_wgrad_matches_fp32_contraction(M, Na, Nb):
    """pcc_wgrad_bf16 (MN-major tensor-core contraction over the rows + the ones-block column sums) against dy^T x and dy.sum(0)
    computed in fp64 from the same bf16 operands: fp32 accumulation, so the error is a few ulp of the largest partial sums."""
    import __graft_entry__  # noqa: F401
    from pcc_b200 import mlp_ops
    torch.manual_seed(M % 1000 + Na)
    dy = (torch.randn(M, Na, device="cuda") * 0.5).to(torch.bfloat16)
    x = torch.relu(torch.randn(M, Nb, device="cuda")).to(torch.bfloat16)
    dw, db = mlp_ops.wgrad(dy, x)
    want_w = dy.double().t() @ x.double()
    want_b = dy.double().sum(0)
    scale = float((dy.double().abs().t() @ x.double().abs()).max())
    assert dw.shape == (Na, Nb) and float((dw.double() - want_w).abs().max()) < 2e-5 * scale
    assert float((db.double() - want_b).abs().max()) < 2e-5 * float(dy.double().abs().sum(0).max())
    # strided operands (a column window of a wider activation) and no bias
    if Na >= 64 and (Na // 2) % 8 == 0:
        dw2, none = mlp_ops.wgrad(dy[:, :Na // 2], x, want_bias=False)
        assert none is None and float((dw2.double() - want_w[:Na // 2]).abs().max()) < 2e-5 * scale


@pytest.mark.parametrize("n_patches", [1, 3, 148, 300])
def test_pointnet_fused_matches_the_two_launch_route(n_patches):
    """pcc_pointnet_fused_bf16 (131 -> 128 -> 256 -> 512 -> 16 + max over the patch in one kernel, nothing in HBM in between)
    against the round-1 route (warp-specialised front chain -> [M, 256] bf16 in HBM -> fused tail) and against an fp32 torch body:
    the two kernels round at the same points except the second layer's bias (fp32 in the fused epilogue, bf16 in the front
    chain's packed weights)."""
    import __graft_entry__  # noqa: F401
    from pcc_b200 import mlp_ops
    torch.manual_seed(n_patches)
    M = n_patches * 256
    feat = torch.relu(torch.randn(M, 128, device="cuda")).to(torch.bfloat16)
    xyz = (torch.rand(M, 3, device="cuda") - 0.5) * 1.2
    dims = [131, 128, 256, 512, 16]
    layers = [(torch.randn(o, i, device="cuda") / i ** 0.5, torch.randn(o, device="cuda") * 0.1, l < 3)
              for l, (i, o) in enumerate(zip(dims[:-1], dims[1:]))]
    assert mlp_ops.pointnet_fused_supported(feat, xyz, layers, 256)
    got = mlp_ops.pointnet_fused(feat, xyz, layers)
    two = mlp_ops.run_chain([(feat, 1), (xyz, 1)], layers, group=256)
    h = torch.cat((feat.float(), xyz), dim=1)
    for w, b, relu in layers:
        h = h @ w.t() + b
        h = torch.relu(h) if relu else h
    want = h.view(n_patches, 256, 16).max(dim=1)[0]
    scale = float(want.abs().max())
    assert got.shape == two.shape == (n_patches, 16)
    assert float((got - two).abs().max()) < 4e-3 * scale
    assert float((got - want).abs().max()) < 3e-2 * scale
