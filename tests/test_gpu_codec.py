"""GPU parity of the batched codec driver (compress -> decompress -> eval) against the CPU restatement of the
reference flow (oracle/torch_modules.py, itself validated against the reference's AE on CPU)."""
import numpy as np
import pytest
import torch

from tools import synth

pytestmark = pytest.mark.gpu

# The device network bodies run on the tensor cores with bf16 operands / fp32 accumulation (north_star: "within a
# bf16/fp32 tolerance"); the CPU reference flow is fp32.  The latent lives in [-3.4, 3.4] before rounding.
# Measured on B200 with the seeded weights: latent 6e-4, decoded coordinates 2.5e-4, Chamfer 2e-4 relative, PSNR 0.01 dB.
LATENT_ATOL = 5e-3      # absolute, on the pre-rounding latent
SYMBOL_MATCH_MIN = 0.98  # fraction of rounded symbols identical to the fp32 flow (the rest sit near a .5 boundary)
REC_ATOL = 2e-3         # absolute, decoded coordinates (unit cube) for identical symbols
CHAMFER_RTOL = 2e-3     # end-to-end metric of the bf16 network vs the fp32 flow
PSNR_ATOL_DB = 0.05


@pytest.fixture(scope="module")
def setup():
    import __graft_entry__  # noqa: F401
    import pcc_b200
    from pcc_b200.codec import PatchCodec
    from pcc_b200.modules import AE
    sd = synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11)
    ae = AE(256, 128, 16, 7)
    ae.load_state_dict(sd)
    ae = ae.cuda().eval()
    return pcc_b200, PatchCodec(ae), sd


def test_roundtrip_matches_cpu_reference_flow(setup):
    from oracle import torch_modules as tm
    pcc, codec, sd = setup
    clouds = synth.modelnet_like(2, 8192, seed=31)
    start = torch.tensor([5, 9], dtype=torch.int64)
    x = torch.from_numpy(clouds).cuda()
    c = codec.compress(x, start.cuda())
    rec = codec.decompress(c["latent_q"], c["centres"], 8192, c["center"], c["longest"])
    met = codec.evaluate(rec, x).cpu().numpy()
    for b in range(2):
        ref = tm.compress_decompress_eval(sd, clouds[b], int(start[b]), threads=8)
        assert np.array_equal(c["centres"][b].cpu().numpy(), ref["centres"])  # FPS + quantisation: exact
        lat = c["latent"][b].cpu().numpy()
        assert np.abs(lat - ref["latent"]).max() < LATENT_ATOL
        lq, rq = c["latent_q"][b].cpu().numpy(), ref["latent_q"]
        near_half = np.abs(np.abs(ref["latent"] - np.floor(ref["latent"])) - 0.5) < LATENT_ATOL
        assert np.array_equal(lq[~near_half], rq[~near_half])
        assert (lq == rq).mean() >= SYMBOL_MATCH_MIN
        # decoder parity on identical symbols
        rec_b = codec.decompress(torch.from_numpy(rq)[None].cuda(), c["centres"][b:b + 1], 8192, c["center"][b:b + 1],
                                 c["longest"][b:b + 1])
        assert np.abs(rec_b[0].cpu().numpy() - ref["rec"]).max() < REC_ATOL
        if np.array_equal(lq, rq):
            assert abs(met[b, 0] - ref["chamfer"]) <= CHAMFER_RTOL * ref["chamfer"]
            assert abs(met[b, 1] - ref["d1_psnr"]) < PSNR_ATOL_DB


def test_ae_forward_signature_and_shapes(setup):
    pcc, codec, sd = setup
    x = torch.rand(5, 256, 3, device="cuda") - 0.5
    new_xyz, latent, latent_q = codec.ae(x)  # AE.py:34-55
    assert new_xyz.shape == (5, 128, 3) and latent.shape == (5, 16) and latent_q.shape == (5, 16)
    assert torch.equal(latent_q, latent.round())
    assert float(latent.abs().max()) <= 3.4 + 1e-6  # sigmoid spread (L - 0.2) / 2


def test_centre_modes_coded_and_reference(setup):
    """compress.py:96-101 with the octree coder on the device: the stream equals the oracle's (pinned to the reference's
    pn_kit.encode_sampled_np), 'reference' centres equal the reference decoder's output, 'coded' centres are what the
    inverse decoder recovers, and decompress works from the coded bytes alone."""
    from oracle import oracle as orc
    from pcc_b200.codec import PatchCodec
    pcc, codec, sd = setup
    clouds = synth.modelnet_like(3, 8192, seed=33)
    x = torch.from_numpy(clouds).cuda()
    start = torch.tensor([1, 2, 3], dtype=torch.int64).cuda()
    pc, _, _, _ = pcc.ops.normalize(x)
    _, fps_xyz = pcc.ops.fps(pc, 64, start, 1e10, return_xyz=True)
    codes, total, depths = orc.encode_sampled_np(fps_xyz.cpu().numpy(), 1, 8192, 0.25)
    for mode in ("coded", "reference"):
        c = PatchCodec(codec.ae, centre_mode=mode).compress(x, start)
        o = c["octree"]
        assert o["depth"].cpu().tolist() == depths
        for b, code in enumerate(codes):
            assert np.array_equal(o["bits"][b, :len(code)].cpu().numpy(), code)
            assert np.array_equal(o["bytes"][b, :(len(code) + 7) // 8].cpu().numpy(), orc.bits_to_bytes(code))
        if mode == "reference":
            ref = np.stack([orc.octree_decode_ref(code) for code in codes])
            assert np.array_equal(c["centres"].cpu().numpy(), ref)
            assert len(np.unique(ref[0], axis=0)) <= 8
        else:
            dec, count, _ = pcc.ops.octree_decode(o["bits"], o["nbits"], mode=1, cap=64)
            assert torch.equal(dec, c["centres"]) and count.cpu().tolist() == [64, 64, 64]
            want = np.stack([orc.octree_stream_centres(fps_xyz[b].cpu().numpy(), depths[b], 64) for b in range(3)])
            assert np.array_equal(c["centres"].cpu().numpy(), want)
            from oracle import torch_modules as tm   # the CPU flow in the same mode: same stream, same centres
            ref = tm.compress_decompress_eval(sd, clouds[0], 1, threads=8, centre_mode="coded")
            assert np.array_equal(ref["centres"], want[0]) and np.array_equal(ref["octree"]["bits"], codes[0])
            assert np.abs(c["latent"][0].cpu().numpy() - ref["latent"]).max() < LATENT_ATOL
        rec = codec.decompress(c["latent_q"], c["centres"], 8192, c["center"], c["longest"])
        met = codec.evaluate(rec, x).cpu().numpy()
        assert rec.shape == (3, 8192, 3) and np.isfinite(met).all()


def test_roundtrip_sweep_matches_per_batch_calls(setup):
    """The double-buffered host sweep (upload of batch s + 1 under batch s) returns what per-batch calls return."""
    from pcc_b200.codec import PatchCodec
    pcc, codec, sd = setup
    codec = PatchCodec(codec.ae, centre_mode="coded")
    host = [torch.from_numpy(synth.modelnet_like(2, 8192, seed=70 + i)).pin_memory() for i in range(5)]
    start = torch.zeros(2, dtype=torch.int64, device="cuda")
    got = {}

    def sink(s, lat, cen, met, octree):
        got[s] = (lat.cpu(), cen.cpu(), met.cpu(), octree["bytes"].cpu(), octree["nbits"].cpu())

    assert codec.roundtrip_sweep(iter(host), start, sink) == 5
    torch.cuda.synchronize()
    eager = dict(got)
    got.clear()
    assert codec.roundtrip_sweep(iter(host), start, sink, graphed=True) == 5      # CUDA-graph replay: same results
    torch.cuda.synchronize()
    for s in range(5):
        assert all(torch.equal(a, b) for a, b in zip(got[s], eager[s]))
    graphed = dict(got)
    got.clear()
    for n in (2, 3):                       # one graph per staging buffer, each on its own compute stream
        got.clear()
        assert codec.roundtrip_sweep(iter(host), start, sink, graphed=True, streams=n) == 5
        torch.cuda.synchronize()
        for s in range(5):
            assert all(torch.equal(a, b) for a, b in zip(got[s], graphed[s]))
    for s, h in enumerate(host):
        lat, cen, met, _, octree = codec.roundtrip(h.cuda(), start, return_octree=True)
        assert torch.equal(got[s][0], lat.cpu()) and torch.equal(got[s][1], cen.cpu())
        assert torch.equal(got[s][2], met.cpu())
        assert torch.equal(got[s][3], octree["bytes"].cpu()) and torch.equal(got[s][4], octree["nbits"].cpu())


def test_evaluate_all_matches_oracle_metrics(setup):
    from oracle import oracle as orc
    pcc, codec, sd = setup
    x = synth.modelnet_like(2, 8192, seed=91)
    y = synth.decompressed_like(x, seed=92)
    m = {k: v.cpu().numpy() for k, v in codec.evaluate_all(torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda()).items()}
    for b in range(2):
        psnr1, mse1 = orc.d1_psnr(x[b], y[b])
        psnr2, mse2 = orc.p2plane_psnr(x[b], y[b])
        assert abs(m["d1_psnr"][b] - psnr1) < 1e-3 and abs(m["d2_psnr"][b] - psnr2) < 1e-3     # dB
        assert abs(m["uc"][b] - orc.calc_uc(x[b], y[b])) <= 1e-9 * m["uc"][b]
        mn, mx = x[b].min(), x[b].max()
        cham = orc.chamfer(((y[b] - mn) / (mx - mn))[None], ((x[b] - mn) / (mx - mn))[None])[0]
        assert abs(m["chamfer"][b] - cham) <= 1e-5 * cham


def test_entropy_stage_matches_oracle_and_round_trips(setup):
    """compress.py:131-136 / decompress.py:88-93 on the device: the CDFs equal pn_kit.pmf_to_cdf + torchac's normalisation (CPU
    statement), the byte streams equal the CPU restatement of the coder on the same CDFs, and decoding returns the latents."""
    from oracle import oracle as orc
    from pcc_b200.modules import ConditionalProbabilityModel
    from pcc_b200 import torchac_compat
    pcc, codec, sd = setup
    torch.manual_seed(5)
    prob = ConditionalProbabilityModel(7, 16).cuda().eval()
    x = torch.from_numpy(synth.modelnet_like(3, 8192, seed=95)).cuda()
    c = codec.compress(x, torch.zeros(3, dtype=torch.int64, device="cuda"))
    data, nbytes = codec.encode_latents(prob, c["latent_q"], c["centres"])
    with torch.no_grad():
        pmf = prob(c["centres"])
    cdf_dev = pcc.ops.pmf_to_cdf_u16(pmf).cpu().numpy()
    cdf_cpu = orc.pmf_to_cdf_u16(pmf.cpu().numpy())
    assert np.array_equal(cdf_dev, cdf_cpu)
    sym = (c["latent_q"].cpu().numpy().astype(np.int16) + 3).reshape(3, -1)
    for b in range(3):
        want = orc.range_encode(cdf_cpu[b].reshape(-1, 8), sym[b])
        got = data[b, :int(nbytes[b])].cpu().numpy().tobytes()
        assert got == want
        assert np.array_equal(orc.range_decode(cdf_cpu[b].reshape(-1, 8), got), sym[b])
    back = codec.decode_latents(prob, c["centres"], data, nbytes)
    assert torch.equal(back, c["latent_q"])
    # the torchac-named entry points (one stream for the whole tensor, CPU tensors in, as compress.py:134-136 calls them)
    cdf_float = torch.cat([torch.zeros_like(pmf[..., :1]), pmf.cumsum(-1)], -1).clamp(max=1.0)[:1].cpu()
    s16 = torch.from_numpy(sym[:1].reshape(1, 64, 16))
    stream = torchac_compat.encode_float_cdf(cdf_float, s16, check_input_bounds=True)
    assert isinstance(stream, bytes) and torch.equal(torchac_compat.decode_float_cdf(cdf_float, stream), s16)


@pytest.mark.parametrize("mode", ["coded", "reference"])
def test_file_formats_round_trip(setup, tmp_path, mode):
    """compress.py:138-151 / decompress.py:72-116 file formats: .p.bin / .s.bin / .c.bin written for a batch decode back to the
    reconstruction the in-memory path produces; the .s.bin bytes are pn_kit.binary_array_to_byte_array of the reference coder's
    stream and the .c.bin is (center, longest) as 4 float32."""
    from oracle import oracle as orc
    from pcc_b200.codec import PatchCodec
    from pcc_b200.modules import ConditionalProbabilityModel
    pcc, codec, sd = setup
    torch.manual_seed(7)
    prob = ConditionalProbabilityModel(7, 16).cuda().eval()
    codec = PatchCodec(codec.ae, centre_mode=mode)
    x = torch.from_numpy(synth.modelnet_like(3, 8192, seed=97)).cuda()
    start = torch.tensor([4, 5, 6], dtype=torch.int64, device="cuda")
    names = ["a", "b", "c"]
    bits = codec.compress_to_files(x, names, str(tmp_path), prob, start)
    c = codec.compress(x, start)
    want = codec.decompress(c["latent_q"], c["centres"], 8192, c["center"], c["longest"])
    got = codec.decompress_from_files(names, str(tmp_path), prob, S=64)
    assert torch.equal(got, want)
    pc, _, _, _ = pcc.ops.normalize(x)
    _, fps_xyz = pcc.ops.fps(pc, 64, start, 1e10, return_xyz=True)
    codes, _, _ = orc.encode_sampled_np(fps_xyz.cpu().numpy(), 1, 8192, 0.25)
    for b, n in enumerate(names):
        assert (tmp_path / (n + ".s.bin")).read_bytes() == orc.bits_to_bytes(codes[b]).tobytes()
        cs = np.fromfile(tmp_path / (n + ".c.bin"), dtype=np.float32)
        assert cs.shape == (4,) and np.array_equal(cs[:3], c["center"][b].cpu().numpy()) and cs[3] == float(c["longest"][b])
        assert bits[b] == 8 * sum((tmp_path / (n + e)).stat().st_size for e in (".p.bin", ".s.bin", ".c.bin"))
    # the evaluation table of eval.py:189-219 (same columns, same rounding)
    df = codec.evaluate_to_csv(names, got, x, bits, str(tmp_path / "eval.csv"))
    assert list(df.columns) == ["filename", "p2pointPSNR", "p2planePSNR", "chamfer_distance", "n_points_input",
                                "n_points_output", "bpp", "uniformity coefficient"]
    assert (tmp_path / "eval.csv").exists() and len(df) == 3 and df["n_points_output"][0] == 8192
    psnr1, _ = orc.d1_psnr(x[0].cpu().numpy(), got[0].cpu().numpy())
    assert abs(df["p2pointPSNR"][0] - round(psnr1, 3)) <= 2e-3 and df["bpp"][0] == bits[0] / 8192


def test_streams_coded_in_a_batch_decode_one_cloud_at_a_time(setup, tmp_path):
    """the probability model runs on kernels whose per-row results do not depend on the batch size, so .p.bin
    streams written by compress_to_files with B = 3 decode with B = 1 (the reference's own batch size), in any order, to
    exactly the batched reconstruction -- also while a Trainer elsewhere in the process had switched TF32 matmuls on."""
    from pcc_b200.codec import PatchCodec
    from pcc_b200.modules import ConditionalProbabilityModel
    pcc, codec, sd = setup
    torch.manual_seed(9)
    prob = ConditionalProbabilityModel(7, 16).cuda().eval()
    codec = PatchCodec(codec.ae, centre_mode="coded")
    x = torch.from_numpy(synth.modelnet_like(3, 8192, seed=131)).cuda()
    start = torch.tensor([40, 50, 60], dtype=torch.int64, device="cuda")
    names = ["p", "q", "r"]
    codec.compress_to_files(x, names, str(tmp_path), prob, start)
    want = codec.decompress_from_files(names, str(tmp_path), prob, S=64)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for b in (2, 0, 1):
            got = codec.decompress_from_files([names[b]], str(tmp_path), prob, S=64)
            assert torch.equal(got[0], want[b])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    c = codec.compress(x, start)
    assert torch.equal(want, codec.decompress(c["latent_q"], c["centres"], 8192, c["center"], c["longest"]))


def test_graphed_sweep_guards(setup):
    """ADVICE r1 hazards of the captured sweep: the graphs read the cache's own copy of start_idx (the caller may free or
    change its tensor), a batch of another shape is refused, and a weight update re-captures instead of replaying stale
    packed weights."""
    import copy
    from pcc_b200.codec import PatchCodec
    pcc, codec, sd = setup
    ae = copy.deepcopy(codec.ae)
    codec = PatchCodec(ae, centre_mode="coded")
    host = [torch.from_numpy(synth.modelnet_like(2, 8192, seed=170 + i)).pin_memory() for i in range(3)]
    got = {}

    def sink(s, lat, cen, met, octree):
        got[s] = (lat.cpu(), cen.cpu(), met.cpu())

    def eager(start):
        return [tuple(t.cpu() for t in codec.roundtrip(h.cuda(), start)[:3]) for h in host]

    start = torch.tensor([3, 4], dtype=torch.int64, device="cuda")
    codec.roundtrip_sweep(iter(host), start, sink, graphed=True)
    torch.cuda.synchronize()
    assert all(all(torch.equal(a, b) for a, b in zip(got[s], e)) for s, e in enumerate(eager(start)))
    # a different start tensor (the first one is gone): same captured graphs, new values
    del start
    start2 = torch.tensor([100, 2000], dtype=torch.int64, device="cuda")
    graphs_before = codec._sweep_graphs["graphs"]
    codec.roundtrip_sweep(iter(host), start2, sink, graphed=True)
    torch.cuda.synchronize()
    assert codec._sweep_graphs["graphs"] is graphs_before
    assert all(all(torch.equal(a, b) for a, b in zip(got[s], e)) for s, e in enumerate(eager(start2)))
    # ragged batch: refused, not silently computed on the stale staging buffer
    with pytest.raises(ValueError):
        codec.roundtrip_sweep(iter(host + [host[0][:1]]), start2, sink, graphed=True)
    torch.cuda.synchronize()
    # weight update: re-capture
    with torch.no_grad():
        ae.pn.mlp_Modules[3][0].bias.add_(0.25)
    codec.roundtrip_sweep(iter(host), start2, sink, graphed=True)
    torch.cuda.synchronize()
    assert codec._sweep_graphs["graphs"] is not graphs_before
    assert all(all(torch.equal(a, b) for a, b in zip(got[s], e)) for s, e in enumerate(eager(start2)))


def test_evaluate_sweep_equals_chunk_by_chunk_evaluate(setup):
    """PatchCodec.evaluate_sweep (chunks alternating over two streams, inputs produced on the caller's stream) returns the rows of
    evaluate() chunk by chunk, bit for bit -- ragged last chunk, and an empty sweep."""
    _, codec, _ = setup
    x = torch.from_numpy(synth.modelnet_like(7, 8192, seed=31)).cuda()
    y = torch.from_numpy(synth.decompressed_like(x.cpu().numpy(), seed=32)).cuda()

    def chunks():
        for i in range(0, 7, 3):
            yield y[i:i + 3] * 1.5, x[i:i + 3] * 1.5

    got = codec.evaluate_sweep(chunks())
    want = torch.cat([codec.evaluate(a, b) for a, b in chunks()])
    assert got.shape == (7, 3) and torch.equal(got, want)
    assert codec.evaluate_sweep(iter(())).shape == (0, 3)

