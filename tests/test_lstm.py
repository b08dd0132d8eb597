"""LSTM recurrence kernel (csrc/lstm.cu, aps_b200_lstm_fwd) against torch.nn.LSTM on the CPU in fp32/fp64 — the
module the reference's DCCRN bottleneck uses (aps/sse/bss/dccrn.py:29-34)."""
import copy

import pytest
import torch as th

from aps_b200 import _lib

gpu = pytest.mark.gpu


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_lstm_symbol_exported():
    assert "aps_b200_lstm_fwd" in _lib.exported_symbols()


def test_oracle_lstm_matches_torch_lstm():
    """The oracle's LSTM restatement (oracle/dccrn.py:_lstmp, used by every DCCRN parity test) against torch.nn.LSTM,
    the module the reference instantiates (aps/sse/bss/dccrn.py:29-34)."""
    from oracle import dccrn as OD
    th.manual_seed(0)
    mod = th.nn.LSTM(12, 8, num_layers=2, batch_first=True).eval()
    proj = th.nn.Linear(8, 12, bias=False)
    sd = {"p.lstm." + k: v.detach() for k, v in mod.state_dict().items()}
    sd["p.proj.weight"] = proj.weight.detach()
    x = th.randn(3, 5, 4, 3)
    with th.no_grad():
        want = proj(mod(x.view(3, 5, 12))[0]).view(3, 5, 4, -1)
        got = OD._lstmp(sd, "p.", x, 2)
    assert rel_err(got, want) < 1e-6


def test_lstm_host_checks_without_gpu():
    """Argument errors surface before any device work; CPU tensors are refused (no CPU fallback)."""
    from aps_b200 import ops
    mod = th.nn.LSTM(8, 8, batch_first=True).eval()
    with pytest.raises(RuntimeError):
        ops.lstm(th.zeros(2, 3, 8), mod)
    with pytest.raises(RuntimeError):
        ops.lstm_multi([th.zeros(2, 3, 8)] * 2, [mod, th.nn.LSTM(8, 12, batch_first=True)])


@gpu
@pytest.mark.parametrize("rows,frames,feats,hidden,layers,bidir", [
    (5, 7, 12, 32, 1, False),          # ragged rows, K handled by the SIMT projection
    (3, 1, 16, 8, 1, False),           # single frame: no recurrent product at all
    (40, 20, 64, 40, 2, False),        # hidden not a multiple of the 16-unit / 32-k tiles (DCCRN goldens use 40)
    (33, 25, 32, 64, 2, True),         # bidirectional, two layers
    (70, 50, 256, 512, 2, False),      # the DCCRN bottleneck shape (recurrence on the tensor-core engine: rows >= 64, H % 32 == 0)
    (130, 30, 48, 96, 1, False),       # ... two 128-row blocks per module, the second mostly padding
    (6, 9, 10, 6, 2, False),           # hidden % 4 != 0: runs zero padded to 8 units (no library fallback)
    (12, 5, 20, 30, 2, True),          # ... bidirectional: the next layer's input columns are padded per direction
])
def test_lstm_matches_torch(rows, frames, feats, hidden, layers, bidir):
    from aps_b200 import ops
    th.manual_seed(rows * 1000 + hidden)
    mod = th.nn.LSTM(feats, hidden, num_layers=layers, bidirectional=bidir, batch_first=True).eval()
    x = th.randn(rows, frames, feats)
    with th.no_grad():
        want, _ = mod.double()(x.double())
        mod.float()
        got = ops.lstm(x.cuda(), mod.cuda())
    assert got.shape == want.shape
    assert rel_err(got.cpu(), want) < 2e-5


@gpu
def test_lstm_rows_are_independent():
    """Rows of a batch never mix: a row computed alone equals the same row inside a large batch, bit for bit."""
    from aps_b200 import ops
    th.manual_seed(3)
    mod = th.nn.LSTM(32, 48, num_layers=2, batch_first=True).eval().cuda()
    x = th.randn(37, 11, 32, device="cuda")
    with th.no_grad():
        full = ops.lstm(x, mod)
        one = ops.lstm(x[20:21].contiguous(), mod)
    # 1 row: the projection runs on the SIMT kernel, 37*11 rows on the tensor-core engine -> compare at tolerance
    assert rel_err(one, full[20:21]) < 1e-5


@gpu
def test_lstm_refuses_unsupported():
    from aps_b200 import ops
    with pytest.raises(RuntimeError, match="batch_first"):
        ops.lstm(th.zeros(2, 3, 8, device="cuda"), th.nn.LSTM(8, 8).cuda())


@gpu
@pytest.mark.parametrize("bidir", [False, True])
def test_lstm_multi_matches_single(bidir):
    """Two modules advanced together (one launch per frame for both) == each module on its own, bit for bit."""
    from aps_b200 import ops
    th.manual_seed(11)
    mods = [th.nn.LSTM(24, 40, num_layers=2, bidirectional=bidir, batch_first=True).eval().cuda() for _ in range(2)]
    xs = [th.randn(9, 13, 24, device="cuda") for _ in range(2)]
    with th.no_grad():
        both = ops.lstm_multi(xs, mods)
        for x, m, y in zip(xs, mods, both):
            assert th.equal(ops.lstm(x, m), y)
            want, _ = copy.deepcopy(m).cpu().double()(x.cpu().double())
            assert rel_err(y.cpu(), want) < 2e-5


@gpu
def test_lstm_tensor_core_recurrence_matches_simt(monkeypatch):
    """Three modules advanced together on the grouped tensor-core recurrence (aps_b200_lstm_group_tc_fwd) against the
    fp32-FMA recurrence kernel and fp64 torch: 200 frames, so a rounding difference would have time to grow."""
    from aps_b200 import ops
    th.manual_seed(21)
    mods = [th.nn.LSTM(40, 64, num_layers=2, batch_first=True).eval().cuda() for _ in range(3)]
    xs = [th.randn(100, 200, 40, device="cuda") for _ in range(3)]
    with th.no_grad():
        tc = ops.lstm_multi(xs, mods)
        monkeypatch.setenv("APS_B200_LSTM", "simt")
        simt = ops.lstm_multi(xs, mods)
        for x, m, a, b in zip(xs, mods, tc, simt):
            want, _ = copy.deepcopy(m).cpu().double()(x.cpu().double())
            assert rel_err(a.cpu(), want) < 2e-5 and rel_err(b.cpu(), want) < 2e-5
            assert rel_err(a, b) < 2e-5
