"""Transformer / conformer encoder (rows a20-a23): oracle vs golden (CPU), CUDA path vs golden + oracle (GPU)."""
import copy
import json

import pytest
import torch as th

from conftest import FLOAT_TOL, golden_names, import_reference, load_golden, rel_err
from oracle.encoder import EncoderOracle, digit_shift

DEV = "cuda:0"


def _sd(g):
    return {k[2:]: v for k, v in g.items() if k.startswith("p.")}


@pytest.mark.parametrize("name", golden_names("enc_"))
def test_oracle_encoder_golden(name):
    cfg, g = load_golden(name)
    y, yl = EncoderOracle(cfg, _sd(g))(g["x"], g["lens"].clone())
    assert th.equal(yl, g["ylens"])                         # subsampled lengths: integers, exact (Q17)
    assert rel_err(y, g["y"]) < 1e-5


@pytest.mark.parametrize("name", golden_names("enc_"))
def test_state_dict_layout_matches_golden(name):
    from aps_b200.asr.transformer import TransformerEncoder
    cfg, g = load_golden(name)
    net = TransformerEncoder(**copy.deepcopy(cfg))
    want = {k: tuple(v.shape) for k, v in _sd(g).items()}
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == want
    net.load_state_dict(_sd(g), strict=True)


def test_training_mode_and_cpu_are_refused():
    from aps_b200.asr.transformer import TransformerEncoder
    cfg, g = load_golden("enc_0")
    net = TransformerEncoder(**copy.deepcopy(cfg))
    with pytest.raises(RuntimeError, match="inference forward only"):
        net.train()(g["x"], None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.eval()(g["x"], None)


@pytest.mark.reference
def test_digit_shift_and_oracle_vs_live_reference():
    import_reference()
    from aps.asr.transformer.encoder import TransformerEncoder
    from aps.asr.transformer.utils import digit_shift as ref_shift
    t = th.randn(7, 2, 3, 13)
    assert th.equal(digit_shift(t), ref_shift(t))
    th.manual_seed(0)
    ak = dict(att_dim=64, nhead=4, feedforward_dim=128, att_dropout=0.1, ffn_dropout=0.1, kernel_size=15, pre_norm=False)
    for pose, pk in (("rel", dict(lradius=20, rradius=20)), ("abs", {}), ("xl", {})):
        cfg = dict(arch="cfmr", input_size=80, output_proj=-1, num_layers=2, proj="conv2d",
                   proj_kwargs=dict(conv_channels=32, num_layers=2), pose=pose, pose_kwargs=pk, arch_kwargs=dict(ak))
        enc = TransformerEncoder(**copy.deepcopy(cfg)).eval()
        x, lens = th.randn(3, 67, 80), th.tensor([67, 58, 37])
        with th.no_grad():
            y, yl = enc(x, lens.clone())
        o, ol = EncoderOracle(cfg, enc.state_dict())(x, lens.clone())
        assert th.equal(yl, ol) and rel_err(o, y) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names("enc_"))
def test_encoder_golden_gpu(name):
    from aps_b200.asr.transformer import TransformerEncoder
    cfg, g = load_golden(name)
    net = TransformerEncoder(**copy.deepcopy(cfg))
    net.load_state_dict(_sd(g), strict=True)
    net = net.to(DEV).eval()
    y, yl = net(g["x"].to(DEV), g["lens"].to(DEV))
    assert th.equal(yl.cpu(), g["ylens"])
    assert y.shape == g["y"].shape
    assert rel_err(y, g["y"]) < FLOAT_TOL
    # no lengths: no key padding mask
    y2, _ = net(g["x"].to(DEV), None)
    o2, _ = EncoderOracle(cfg, _sd(g))(g["x"], None)
    assert rel_err(y2, o2) < FLOAT_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["enc_0", "enc_3", "enc_4"])
def test_encoder_tensor_core_engine_and_graph_replay(name, monkeypatch):
    """Same goldens through the tcgen05 3xTF32 GEMM engine, and three calls with one shape so that the third
    is a CUDA-graph replay (must reproduce the eager result)."""
    from aps_b200 import ops
    from aps_b200.asr.transformer import TransformerEncoder
    monkeypatch.setattr(ops, "GEMM_ENGINE", "tc")
    cfg, g = load_golden(name)
    net = TransformerEncoder(**copy.deepcopy(cfg))
    net.load_state_dict(_sd(g), strict=True)
    net = net.to(DEV).eval()
    outs = [net(g["x"].to(DEV), g["lens"].to(DEV))[0] for _ in range(3)]
    assert rel_err(outs[0], g["y"]) < FLOAT_TOL
    assert rel_err(outs[2], outs[0]) < 1e-6 and len(net._graphs) == 1 and isinstance(list(net._graphs.values())[0], tuple)
    # a different batch through the replayed graph
    x2 = th.flip(g["x"], [0])
    l2 = th.flip(g["lens"], [0])
    y2, _ = net(x2.to(DEV), l2.to(DEV))
    o2, _ = EncoderOracle(cfg, _sd(g))(x2, l2.clone())
    assert rel_err(y2, o2) < FLOAT_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("bn", ["", "64", "128", "256"])
def test_tensor_core_gemm_vs_fp64(monkeypatch, bn):
    """tcgen05 3xTF32 GEMM (every tile width): error against fp64 must stay at the fp32 level (well below TF32's 1e-3);
    ragged M / N, more tiles than SMs, unaligned rows (N = 257 takes the transposed scalar epilogue)."""
    from aps_b200 import ops
    monkeypatch.setattr(ops, "GEMM_ENGINE", "tc")
    monkeypatch.setenv("APS_B200_TC_BN", bn)
    th.manual_seed(2)
    for (M, K, N) in ((128, 32, 64), (3200, 256, 2048), (3200, 2048, 256), (333, 96, 200), (700, 2304, 256), (300, 64, 257),
                      (40000, 64, 320)):
        x, w, b = th.randn(M, K, device=DEV), th.randn(N, K, device=DEV) / K**0.5, th.randn(N, device=DEV)
        r = th.randn(M, N, device=DEV)
        ref = x.double() @ w.double().t() + b.double()
        y = ops.linear(x, w, b)
        # the TMEM accumulator rounds toward zero: the error grows ~K/8 * 2^-24 (1.2e-5 at K = 2048), far below TF32's 1e-3
        assert float((y.double() - ref).abs().max() / ref.abs().max()) < 3e-5, (M, K, N)
        y2 = ops.linear(x, w, b, act="swish", alpha=0.5, residual=r)
        ref2 = 0.5 * ref * th.sigmoid(ref) + r.double()
        assert float((y2.double() - ref2).abs().max() / ref2.abs().max()) < 3e-5, (M, K, N)


@pytest.mark.gpu
def test_tensor_core_epilogues(monkeypatch):
    """GLU (aligned and unaligned output rows), PReLU + post affine, strided input rows on the tensor-core engine."""
    import torch.nn.functional as F
    from aps_b200 import ops
    monkeypatch.setattr(ops, "GEMM_ENGINE", "tc")
    th.manual_seed(4)
    x, w, b = th.randn(300, 256, device=DEV), th.randn(512, 256, device=DEV) / 16, th.randn(512, device=DEV)
    ref = F.linear(x.double(), w.double(), b.double())
    for half in (256, 129):
        n = 2 * half
        wi = th.stack([w[:half], w[half:n]], 1).reshape(n, 256).contiguous()
        bi = th.stack([b[:half], b[half:n]], 1).reshape(n).contiguous()
        assert rel_err(ops.linear(x, wi, bi, act="glu"), F.glu(ref[:, :n], -1)) < 1e-5
    sl, ps, pt = th.rand(512, device=DEV), th.rand(512, device=DEV) + 0.5, th.randn(512, device=DEV)
    refp = th.where(ref >= 0, ref, ref * sl.double()) * ps.double() + pt.double()
    assert rel_err(ops.linear(x, w, b, act="prelu", slope=sl, post=(ps, pt)), refp) < 1e-5
    for act, fn in (("relu", F.relu), ("tanh", th.tanh), ("sigmoid", th.sigmoid), ("gelu", F.gelu)):
        assert rel_err(ops.linear(x, w, b, act=act), fn(ref)) < 1e-5
    big = th.randn(640, 300, device=DEV)
    assert rel_err(ops.linear(big[:, 20:148], w[:, :128].contiguous()), F.linear(big[:, 20:148].double(), w[:, :128].double())) < 1e-5


@pytest.mark.gpu
def test_tensor_core_conv_path_vs_torch(monkeypatch):
    """conv2d as an implicit GEMM on the tensor-core engine (in-kernel im2col gather + TF32 split)."""
    import torch.nn.functional as F
    from aps_b200 import ops
    monkeypatch.setattr(ops, "GEMM_ENGINE", "tc")
    th.manual_seed(3)
    for (B, H, W, Ci, Co, k, s_, p_, d_) in ((3, 40, 21, 32, 64, (3, 3), (2, 2), (1, 1), (1, 1)),
                                             (1, 18, 9, 64, 32, (5, 2), (2, 1), (2, 0), (1, 1)),
                                             (2, 31, 17, 32, 288, (3, 3), (2, 1), (0, 1), (1, 2)),
                                             (4, 100, 20, 256, 256, (3, 3), (2, 2), (1, 1), (1, 1))):
        x, w, b = th.randn(B, Ci, H, W), th.randn(Co, Ci, *k) * 0.1, th.randn(Co)
        ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), stride=s_, padding=p_, dilation=d_)).permute(0, 2, 3, 1)
        got = ops.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous().to(DEV), w.permute(0, 2, 3, 1).contiguous().to(DEV),
                              b.to(DEV), stride=s_, padding=p_, dilation=d_, act="relu")
        assert got.shape == ref.shape and rel_err(got, ref) < 3e-5


@pytest.mark.gpu
def test_tma_fed_conv_vs_fp64_and_gather(monkeypatch):
    """Convolutions whose A tiles come in as strided TMA boxes (MODE 4 of the engine: whole output rows per 128-row tile,
    zero padding = the TMA's out-of-bounds fill, lo tile derived in shared memory) against fp64, and against the
    gather-fed kernel on the same input: the conformer front's geometries (padding 1, odd height), no padding, stride 1 and
    (2, 1), dilation, 5 x 2 taps, a last tile shorter than the others, one output row per image, 64- / 128- / 256-wide
    tiles, rows longer than a tile (column chunks), the lo companion output, and a geometry that must fall back."""
    import torch.nn.functional as F
    from aps_b200 import ops
    monkeypatch.setattr(ops, "GEMM_ENGINE", "tc")
    th.manual_seed(11)
    for (B, H, W, Ci, Co, k, s_, p_, d_) in (
            (3, 39, 40, 256, 256, (3, 3), (2, 2), (1, 1), (1, 1)),      # front conv2 (OW 20: 6 rows per tile), odd height
            (5, 100, 20, 256, 256, (3, 3), (2, 2), (1, 1), (1, 1)),     # front conv3 (OW 10, 12 rows per tile, short last tile)
            (3, 38, 39, 32, 64, (3, 3), (2, 2), (0, 0), (1, 1)),        # no padding, odd width
            (4, 4, 250, 64, 128, (3, 3), (2, 2), (0, 0), (1, 1)),       # one output row per image (OH = 1, OW = 124)
            (2, 30, 25, 32, 320, (5, 2), (2, 1), (2, 1), (1, 1)),       # 5 x 2 taps, stride (2, 1), two column blocks
            (2, 24, 31, 64, 48, (3, 3), (1, 1), (1, 1), (1, 2)),        # stride 1, dilation along w
            (6, 9, 40, 32, 64, (3, 3), (3, 4), (2, 0), (2, 1)),         # stride (3, 4), dilation along h
            (2, 10, 480, 32, 48, (3, 3), (2, 2), (0, 0), (1, 1)),       # OW = 239: two chunks of 120 columns per output row
            (2, 33, 251, 64, 128, (5, 2), (2, 1), (2, 1), (1, 1)),      # DCCRN encoder geometry (OW = 252: chunks of 126)
            (2, 6, 140, 32, 36, (3, 3), (1, 1), (1, 0), (1, 1))):       # OW = 138 in chunks of 69: 54 % fill -> gather fallback
        x, w, b = th.randn(B, Ci, H, W), th.randn(Co, Ci, *k) * 0.1, th.randn(Co)
        ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), stride=s_, padding=p_, dilation=d_)).permute(0, 2, 3, 1)
        xg, wg, bg = x.permute(0, 2, 3, 1).contiguous().to(DEV), w.permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV)
        got, lo = ops.conv2d_nhwc(xg, wg, bg, stride=s_, padding=p_, dilation=d_, act="relu", want_lo=True)
        assert got.shape == ref.shape and rel_err(got, ref) < 3e-5, (B, H, W, Ci, Co, k)
        assert th.equal(lo, ops.lo_companion(got))
        monkeypatch.setenv("APS_B200_NO_CONV_TMA", "1")
        old = ops.conv2d_nhwc(xg, wg, bg, stride=s_, padding=p_, dilation=d_, act="relu")
        monkeypatch.delenv("APS_B200_NO_CONV_TMA")
        assert rel_err(got, old) < 1e-5, (B, H, W, Ci, Co, k)


@pytest.mark.gpu
@pytest.mark.parametrize("classes", [True, False])
def test_tensor_core_conv_transpose_vs_torch(monkeypatch, classes):
    """conv_transpose2d as an implicit gather GEMM; with stride_h > 1 the rows are walked class-major (oh % stride_h)
    and the all-zero taps of a tile are skipped — both orders must give the reference result."""
    import torch.nn.functional as F
    from aps_b200 import ops
    monkeypatch.setattr(ops, "GEMM_ENGINE", "tc")
    if not classes:
        monkeypatch.setenv("APS_B200_TC_NO_CLASSES", "1")
    th.manual_seed(5)
    for (B, H, W, Ci, Co, k, s_, p_, op_) in ((2, 9, 30, 64, 32, (3, 3), (2, 1), (1, 1), (0, 0)),
                                              (2, 4, 25, 256, 128, (3, 3), (2, 1), (0, 1), (1, 0)),
                                              (1, 17, 11, 32, 64, (3, 3), (2, 2), (1, 1), (1, 1)),
                                              (3, 17, 30, 64, 4, (3, 3), (2, 1), (1, 1), (0, 0)),
                                              (2, 7, 13, 32, 36, (3, 3), (1, 1), (1, 1), (0, 0)),
                                              (2, 6, 9, 32, 40, (5, 3), (3, 1), (2, 1), (2, 0)),
                                              (5, 40, 60, 32, 32, (3, 3), (2, 1), (0, 1), (1, 0))):
        x, w, b = th.randn(B, Ci, H, W), th.randn(Ci, Co, *k) * 0.1, th.randn(Co)
        ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=s_, padding=p_, output_padding=op_).permute(0, 2, 3, 1)
        got = ops.conv_transpose2d_nhwc(x.permute(0, 2, 3, 1).contiguous().to(DEV),
                                        w.transpose(0, 1).permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV), stride=s_,
                                        padding=p_, output_padding=op_)
        assert got.shape == ref.shape and rel_err(got, ref) < 3e-5


@pytest.mark.gpu
def test_tensor_core_conv_transpose_fused_skip(monkeypatch):
    """DCCRN "cat" skip connection read in place by the transposed-conv gather == conv_transpose2d(cat_complex(x, skip))."""
    import torch.nn.functional as F
    from aps_b200 import ops
    monkeypatch.setattr(ops, "GEMM_ENGINE", "tc")
    th.manual_seed(6)
    for (B, H, W, C2, Co) in ((2, 9, 30, 64, 32), (3, 4, 25, 128, 128), (2, 17, 40, 64, 4)):
        x, sk = th.randn(B, H, W, C2), th.randn(B, H, W, C2)
        w, b = th.randn(Co, 3, 3, 2 * C2) * 0.05, th.randn(Co)
        cat = ops.cat_complex(x, sk)                                            # [re_x | re_s | im_x | im_s]
        ref = F.conv_transpose2d(cat.permute(0, 3, 1, 2).double(), w.permute(3, 0, 1, 2).double(), b.double(),
                                 stride=(2, 1), padding=(1, 1)).permute(0, 2, 3, 1)
        got = ops.conv_transpose2d_nhwc(x.to(DEV), w.to(DEV), b.to(DEV), stride=(2, 1), padding=(1, 1), skip=sk.to(DEV))
        assert got.shape == ref.shape and rel_err(got, ref) < 3e-5
        # a shape the engine does not take (C2 = 16) goes through the materialised concat
    x, sk = th.randn(1, 5, 7, 16), th.randn(1, 5, 7, 16)
    w, b = th.randn(6, 3, 3, 32) * 0.1, th.randn(6)
    ref = F.conv_transpose2d(ops.cat_complex(x, sk).permute(0, 3, 1, 2), w.permute(3, 0, 1, 2), b, stride=(2, 1),
                             padding=(1, 1)).permute(0, 2, 3, 1)
    got = ops.conv_transpose2d_nhwc(x.to(DEV), w.to(DEV), b.to(DEV), stride=(2, 1), padding=(1, 1), skip=sk.to(DEV))
    assert rel_err(got, ref) < 1e-5


@pytest.mark.gpu
def test_narrow_output_conv_transpose_vs_torch():
    """Transposed convolution with <= 8 output channels (last DCCRN decoder layer) on its dedicated kernel: plain and
    with the cat-skip tensor read in place, strides 1 / 2 / 3, ragged channel counts, LeakyReLU epilogue."""
    import torch.nn.functional as F
    from aps_b200 import ops
    th.manual_seed(8)
    for (B, H, W, Cx, Co, k, s_, p_, op_, skip) in ((2, 17, 30, 32, 4, (3, 3), (2, 1), (1, 1), (0, 0), True),
                                                    (3, 9, 11, 16, 6, (3, 3), (2, 1), (1, 1), (0, 0), True),
                                                    (1, 5, 7, 40, 8, (3, 2), (2, 2), (1, 0), (1, 1), True),
                                                    (2, 6, 9, 12, 1, (5, 3), (3, 1), (2, 1), (2, 0), False),
                                                    (2, 8, 8, 4, 3, (1, 1), (1, 1), (0, 0), (0, 0), False),
                                                    (1, 33, 70, 64, 2, (3, 3), (2, 1), (0, 1), (1, 0), False)):
        x = th.randn(B, H, W, Cx)
        sk = th.randn(B, H, W, Cx) if skip else None
        Cin = 2 * Cx if skip else Cx
        w, b = th.randn(Co, *k, Cin) * 0.1, th.randn(Co)
        full = ops.cat_complex(x, sk) if skip else x
        ref = F.leaky_relu(F.conv_transpose2d(full.permute(0, 3, 1, 2).double(), w.permute(3, 0, 1, 2).double(), b.double(),
                                              stride=s_, padding=p_, output_padding=op_), 0.01).permute(0, 2, 3, 1)
        got = ops.conv_transpose2d_nhwc(x.to(DEV), w.to(DEV), b.to(DEV), stride=s_, padding=p_, output_padding=op_,
                                        act="leaky_relu", leaky=0.01, skip=sk.to(DEV) if skip else None)
        assert got.shape == ref.shape and rel_err(got, ref) < 2e-6


@pytest.mark.gpu
def test_dense_kernels_vs_torch(monkeypatch):
    """Exact-fp32 (SIMT) GEMM epilogues / implicit conv / LayerNorm / depthwise conv against plain fp32 torch on the
    CPU (the tensor-core engine has its own tests above)."""
    import torch.nn.functional as F
    from aps_b200 import ops
    monkeypatch.setattr(ops, "GEMM_ENGINE", "simt")
    th.manual_seed(1)
    for (M, K, N) in ((37, 257, 50), (300, 256, 512), (1000, 96, 64), (5, 8, 6), (129, 2304, 256)):
        x, w, b, r = th.randn(M, K), th.randn(N, K) / K**0.5, th.randn(N), th.randn(M, N)
        ref = F.linear(x, w, b)
        xg, wg, bg, rg = x.to(DEV), w.to(DEV), b.to(DEV), r.to(DEV)
        assert rel_err(ops.linear(xg, wg, bg), ref) < 1e-5
        assert rel_err(ops.linear(xg, wg, bg, act="swish", alpha=0.5, residual=rg), 0.5 * ref * th.sigmoid(ref) + r) < 1e-5
        assert rel_err(ops.linear(xg, wg, None, act="relu"), F.relu(F.linear(x, w))) < 1e-5
        assert rel_err(ops.linear(xg, wg, bg, act="gelu"), F.gelu(ref)) < 1e-5
        assert rel_err(ops.linear(xg, wg, bg, act="tanh"), th.tanh(ref)) < 1e-5
        slope = th.rand(N)
        assert rel_err(ops.linear(xg, wg, bg, act="prelu", slope=slope.to(DEV)), th.where(ref >= 0, ref, ref * slope)) < 1e-5
        if N % 2 == 0:
            wi = th.stack([w[:N // 2], w[N // 2:]], 1).reshape(N, K)
            bi = th.stack([b[:N // 2], b[N // 2:]], 1).reshape(N)
            assert rel_err(ops.linear(xg, wi.to(DEV), bi.to(DEV), act="glu"), F.glu(ref, -1)) < 1e-5
    # strided input rows (a column slice of a wider matrix)
    big = th.randn(64, 300)
    assert rel_err(ops.linear(big.to(DEV)[:, 20:148], wg[:, :128].contiguous()), F.linear(big[:, 20:148], w[:, :128])) < 1e-5
    # implicit-GEMM convolutions (NHWC) incl. Cin = 1 and dilation / asymmetric stride
    for (B, H, W, Ci, Co, k, s, p, d) in ((2, 33, 20, 1, 16, (3, 3), (2, 2), (1, 1), (1, 1)),
                                          (2, 21, 40, 1, 256, (3, 3), (2, 2), (1, 1), (1, 1)),    # thin kernel, padded rows
                                          (2, 9, 37, 2, 32, (3, 3), (2, 1), (0, 1), (1, 1)),      # thin kernel, 2 channels
                                          (1, 5, 2500, 1, 16, (3, 3), (1, 2), (1, 0), (1, 1)),    # rows too wide to stage
                                          (3, 17, 11, 8, 24, (3, 3), (2, 2), (1, 1), (1, 1)),
                                          (2, 20, 9, 4, 6, (5, 2), (2, 1), (2, 0), (1, 1)),
                                          (1, 40, 1, 12, 10, (3, 1), (1, 1), (2, 0), (2, 1))):
        x, w, b = th.randn(B, Ci, H, W), th.randn(Co, Ci, *k) * 0.2, th.randn(Co)
        ref = F.relu(F.conv2d(x, w, b, stride=s, padding=p, dilation=d)).permute(0, 2, 3, 1)
        got = ops.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous().to(DEV), w.permute(0, 2, 3, 1).contiguous().to(DEV),
                              b.to(DEV), stride=s, padding=p, dilation=d, act="relu")
        assert got.shape == ref.shape and rel_err(got, ref) < 1e-5
    # LayerNorm(alpha * x + res)
    x, r, g_, b_ = th.randn(77, 256), th.randn(77, 256), th.rand(256) + 0.5, th.randn(256)
    ref = F.layer_norm(0.5 * x + r, (256,), g_, b_)
    assert rel_err(ops.layernorm(x.to(DEV), g_.to(DEV), b_.to(DEV), residual=r.to(DEV), alpha=0.5), ref) < 1e-5
    # depthwise conv over time on batch-major rows
    N, T, D, K = 3, 29, 40, 15
    x, w, b = th.randn(N, T, D), th.randn(D, 1, K), th.randn(D)
    ref = F.conv1d(x.transpose(1, 2), w, b, padding=7, groups=D).transpose(1, 2)
    got = ops.dwconv1d(x.reshape(N * T, D).to(DEV), N, T, w.view(D, K).t().contiguous().to(DEV), b.to(DEV), left_pad=7)
    assert rel_err(got.view(N, T, D), ref) < 1e-5
    ref = F.conv1d(x.transpose(1, 2), w[..., :3], b, padding=4, dilation=4, groups=D).transpose(1, 2)
    got = ops.dwconv1d(x.reshape(N * T, D).to(DEV), N, T, w.view(D, K)[:, :3].t().contiguous().to(DEV), b.to(DEV),
                       dilation=4, left_pad=4)
    assert rel_err(got.view(N, T, D), ref) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_c4_conformer_full_size_subset_vs_oracle(engine, monkeypatch):
    """BASELINE config[3]: conformer 12L d=256 h=4 rel-pos, conv2d x3 front, B=64 on 80-d fbank [64, 398, 80].
    Parity of sampled utterances against the CPU oracle + batch-shard invariance (bit identical rows)."""
    from aps_b200 import ops
    from aps_b200.asr.transformer import TransformerEncoder
    monkeypatch.setattr(ops, "GEMM_ENGINE", engine)
    cfg = dict(arch="cfmr", input_size=80, output_proj=-1, num_layers=12, proj="conv2d",
               proj_kwargs=dict(conv_channels=256, num_layers=3), pose="rel",
               pose_kwargs=dict(dropout=0.1, lradius=256, rradius=256),
               arch_kwargs=dict(att_dim=256, nhead=4, feedforward_dim=2048, att_dropout=0.1, ffn_dropout=0.1,
                                kernel_size=15, pre_norm=False))
    th.manual_seed(0)
    net = TransformerEncoder(**copy.deepcopy(cfg)).eval()
    x = th.randn(64, 398, 80)
    lens = th.full((64,), 398, dtype=th.int64)
    lens[1::2] = 301
    dev_net = copy.deepcopy(net).to(DEV)
    y, yl = dev_net(x.to(DEV), lens.to(DEV))
    assert y.shape == (64, 50, 256) and yl.tolist()[:2] == [50, 38]
    rows = [0, 1]
    # the key padding mask spans max(len): run the oracle on rows that include a full-length utterance
    o, ol = EncoderOracle(cfg, net.state_dict())(x[rows], lens[rows].clone())
    assert th.equal(ol, yl[rows].cpu())
    assert rel_err(y[rows], o) < FLOAT_TOL
    alone, _ = dev_net(x[rows].to(DEV), lens[rows].to(DEV))
    assert th.equal(alone, y[rows])


@pytest.mark.gpu
@pytest.mark.parametrize("bn", ["", "64", "128", "256"])
def test_tensor_core_tma_fed_linear_and_lo_companion(monkeypatch, bn):
    """MODE 3 of the GEMM engine: the activation is fed RAW (the tensor core truncates it to TF32) next to its lo
    companion, both by TMA.  Error against fp64 stays at the fp32 level; the companion a kernel writes is bit-identical
    to ops.lo_companion of its output."""
    from aps_b200 import ops
    monkeypatch.setattr(ops, "GEMM_ENGINE", "tc")
    monkeypatch.setenv("APS_B200_TC_BN", bn)
    th.manual_seed(3)
    for (M, K, N) in ((128, 32, 64), (3200, 256, 2048), (3200, 2048, 256), (333, 96, 200), (700, 2304, 256), (3200, 256, 768)):
        x, w, b = th.randn(M, K, device=DEV), th.randn(N, K, device=DEV) / K**0.5, th.randn(N, device=DEV)
        r = th.randn(M, N, device=DEV)
        xl = ops.lo_companion(x)
        hi = (x.view(th.int32) & -8192).view(th.float32)
        assert float((hi.double() + xl.double() - x.double()).abs().max() / x.abs().max()) < 2.0**-20
        ref = x.double() @ w.double().t() + b.double()
        y, yl = ops.linear(x, w, b, x_lo=xl, want_lo=True)
        assert float((y.double() - ref).abs().max() / ref.abs().max()) < 3e-5, (M, K, N)
        assert th.equal(yl, ops.lo_companion(y))
        y2 = ops.linear(x, w, b, act="swish", alpha=0.5, residual=r, x_lo=xl)
        ref2 = 0.5 * ref * th.sigmoid(ref) + r.double()
        assert float((y2.double() - ref2).abs().max() / ref2.abs().max()) < 3e-5, (M, K, N)
        if N % 2 == 0:
            y3 = ops.linear(x, w, b, act="glu", x_lo=xl)
            ref3 = ref[:, 0::2] * th.sigmoid(ref[:, 1::2])
            assert float((y3.double() - ref3).abs().max() / ref3.abs().max()) < 3e-5, (M, K, N)


@pytest.mark.gpu
@pytest.mark.parametrize("ksplit", [2, 5, 8])
def test_split_k_linear_reduced_by_layernorm2(monkeypatch, ksplit):
    """Skinny GEMM (d_ff -> d_model) cut into K slices whose raw partial sums are reduced by the LayerNorm kernel:
    LN(alpha * (x @ W.T + b) + res) against fp64; rows are bit-identical whatever the batch (slices depend on K only)."""
    from aps_b200 import ops
    monkeypatch.setattr(ops, "GEMM_ENGINE", "tc")
    th.manual_seed(4)
    for (M, K, N) in ((3200, 2048, 256), (3200, 2560, 256), (200, 1024, 128)):
        x, w, b = th.randn(M, K, device=DEV), th.randn(N, K, device=DEV) / K**0.5, th.randn(N, device=DEV)
        r, g, be = th.randn(M, N, device=DEV), th.rand(N, device=DEV) + 0.5, th.randn(N, device=DEV)
        xl = ops.lo_companion(x)
        parts = ops.linear(x, w, None, x_lo=xl, ksplit=ksplit)
        assert parts.shape == (ksplit, M, N)
        v = 0.5 * (x.double() @ w.double().t() + b.double()) + r.double()
        y, yl = ops.layernorm2(parts, g, be, 1e-5, bias=b, residual=r, alpha=0.5)
        ref = th.nn.functional.layer_norm(v, (N,), g.double(), be.double(), 1e-5)
        assert float((y.double() - ref).abs().max() / ref.abs().max()) < 3e-5, (M, K, N)
        assert th.equal(yl, ops.lo_companion(y))
        y0, _ = ops.layernorm2(parts, None, None, 0.0, bias=b, residual=r, alpha=0.5, normalize=False)
        assert float((y0.double() - v).abs().max() / v.abs().max()) < 3e-5
        # a sub-batch gives bit-identical rows
        sub = ops.linear(x[:128].contiguous(), w, None, x_lo=xl[:128].contiguous(), ksplit=ksplit)
        assert th.equal(sub, parts[:, :128])


@pytest.mark.gpu
def test_lo_companions_of_attention_and_depthwise_conv():
    from aps_b200 import ops
    th.manual_seed(5)
    N, L, H, E = 4, 50, 4, 256
    qkv = th.randn(N * L, 3 * E, device=DEV)
    pos = th.randn(2 * L - 1, E // H, device=DEV)
    ctx, lo = ops.mhsa(qkv, N, L, H, mode=1, pos=pos, kpm_fill=-3.4e38, want_lo=True)
    assert th.equal(ctx, ops.mhsa(qkv, N, L, H, mode=1, pos=pos, kpm_fill=-3.4e38)) and th.equal(lo, ops.lo_companion(ctx))
    x = th.randn(N * L, E, device=DEV)
    w, b = th.randn(15, E, device=DEV), th.randn(E, device=DEV)
    c, cl = ops.dwconv1d(x, N, L, w, b, left_pad=7, act="swish", want_lo=True)
    assert th.equal(c, ops.dwconv1d(x, N, L, w, b, left_pad=7, act="swish")) and th.equal(cl, ops.lo_companion(c))
    y, yl = ops.conv2d_nhwc(th.randn(4, 40, 24, 32, device=DEV), th.randn(64, 3, 3, 32, device=DEV) / 17, None, stride=(2, 2),
                            padding=(1, 1), act="relu", want_lo=True)       # 960 output rows: tensor-core engine
    assert yl is not None and th.equal(yl, ops.lo_companion(y))


@pytest.mark.gpu
@pytest.mark.parametrize("L", [50, 64, 100, 131])
def test_register_tiled_attention_matches_the_simple_kernel(monkeypatch, L):
    """mhsa_tiled_kernel (head dim 64) against mhsa_kernel (one warp per query) for the three position modes, key padding
    and additive masks, and against an fp64 torch evaluation of the abs mode."""
    from aps_b200 import ops
    th.manual_seed(6)
    N, H, E = 3, 4, 256
    qkv = th.randn(N * L, 3 * E, device=DEV)
    lens = th.tensor([L, L - 7, max(1, L // 2)], device=DEV)
    kpm = (th.arange(L, device=DEV)[None, :] >= lens[:, None]).to(th.uint8).contiguous()
    amask = th.zeros(L, L, device=DEV).masked_fill(th.rand(L, L, device=DEV) < 0.1, float("-inf"))
    amask.fill_diagonal_(0.0)
    pos1 = th.randn(2 * L - 1, E // H, device=DEV)
    pos2 = th.randn(2 * L - 1, E, device=DEV)
    u, v = th.randn(H, E // H, device=DEV), th.randn(H, E // H, device=DEV)
    cases = [dict(mode=0, kpm=kpm, kpm_fill=float("-inf")), dict(mode=0, attn_mask=amask, kpm_fill=float("-inf")),
             dict(mode=1, pos=pos1, kpm=kpm, kpm_fill=-3.4028234663852886e38),
             dict(mode=2, pos=pos2, rel_u=u, rel_v=v, kpm=kpm, kpm_fill=-3.4028234663852886e38, qpos_is_value=True, attn_mask=amask)]
    for kw in cases:
        monkeypatch.delenv("APS_B200_MHSA", raising=False)
        fast = ops.mhsa(qkv, N, L, H, **kw)
        monkeypatch.setenv("APS_B200_MHSA", "simple")
        slow = ops.mhsa(qkv, N, L, H, **kw)
        assert rel_err(fast, slow) < 2e-5, kw["mode"]
    monkeypatch.delenv("APS_B200_MHSA", raising=False)
    q, k, vv = [t.view(N, L, H, E // H).permute(0, 2, 1, 3).double() for t in qkv.view(N, L, 3, E).unbind(2)]
    sc = q @ k.transpose(-1, -2) / (E // H)**0.5
    sc = sc.masked_fill(kpm.bool()[:, None, None, :], float("-inf"))
    ref = (th.softmax(sc, -1) @ vv).permute(0, 2, 1, 3).reshape(N * L, E)
    assert rel_err(ops.mhsa(qkv, N, L, H, mode=0, kpm=kpm, kpm_fill=float("-inf")), ref) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("arch,pose", [("cfmr", "rel"), ("xfmr", "abs"), ("cfmr", "xl")])
def test_batch_decoding_prep_equals_one_utterance_at_a_time(arch, pose):
    """Row f4: the ragged decoding batch of aps/asr/ctc.py:58-84 in ONE pass.  Every utterance must come out as if it had
    been run alone (the reference loops for exactly that reason); "xl" depends on the padded length and takes the loop."""
    from aps_b200.asr.decoding import batch_decoding_prep
    from aps_b200.asr.transformer import TransformerEncoder
    from aps_b200.transform import AsrTransform
    th.manual_seed(7)
    tf = AsrTransform(feats="fbank-log-cmvn", frame_len=400, frame_hop=160, window="hamm", pre_emphasis=0.97,
                      num_mels=80).to(DEV).eval()
    ak = dict(att_dim=128, nhead=2, feedforward_dim=1024, att_dropout=0.1, ffn_dropout=0.1, pre_norm=False)
    if arch == "cfmr":
        ak["kernel_size"] = 15
    cfg = dict(arch=arch, input_size=80, output_proj=-1, num_layers=2, proj="conv2d",
               proj_kwargs=dict(conv_channels=32, num_layers=3), pose=pose,
               pose_kwargs=dict(lradius=32, rradius=32) if pose == "rel" else {}, arch_kwargs=ak)
    enc = TransformerEncoder(**copy.deepcopy(cfg)).eval()
    with th.no_grad():
        for name, buf in enc.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.2 * th.randn(buf.shape))
            if name.endswith("running_var"):
                buf.copy_(0.5 + th.rand(buf.shape))
    dev_enc = copy.deepcopy(enc).to(DEV)
    wavs = [0.1 * th.randn(n) for n in (32000, 24321, 18000, 40000, 14001, 32000)]
    out, ln = batch_decoding_prep(tf, dev_enc, [w.to(DEV) for w in wavs])
    assert out.shape[0] == len(wavs) and out.shape[1] == int(ln.max())
    for i, w in enumerate(wavs):
        f, _ = tf(w[None].to(DEV), None)
        y, _ = dev_enc(f, None)
        assert int(ln[i]) == y.shape[1]
        assert rel_err(out[i, :y.shape[1]], y[0]) < FLOAT_TOL, (arch, pose, i)
        assert float(out[i, y.shape[1]:].abs().max()) == 0.0 if y.shape[1] < out.shape[1] else True
    # ... and against the CPU oracle of the reference, utterance 1 alone
    from oracle import transform as OT
    f1, _ = OT.AsrFeatures(OT.AsrFeatCfg())(wavs[1][None], None)
    o1, _ = EncoderOracle(cfg, enc.state_dict())(f1, None)
    assert rel_err(out[1, :o1.shape[1]], o1[0]) < FLOAT_TOL
    # time-major variant
    out_t, _ = batch_decoding_prep(tf, dev_enc, [w.to(DEV) for w in wavs], batch_first=False)
    assert th.equal(out_t.transpose(0, 1), out)
