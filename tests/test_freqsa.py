"""Frequency-domain spectral-approximation tasks (row f1: aps/task/sse.py:207-455, "sse@freq_linear_sa",
"sse@freq_mel_sa") against values of the live reference (tests/golden/freqsa_0.npz, oracle/gen_golden.py).

The CPU test drives the task shells with the oracle's STFT as context (host logic: reference magnitude, masking,
distances, permutations); the GPU test uses the real EnhTransform context (fused polar STFT kernel)."""
import pytest
import torch as th
import torch.nn as nn

from conftest import load_golden, rel_err
from oracle import transform as O

FLOAT_TOL = 1e-4


class _OracleCtx:
    """forward_stft context on the CPU: the oracle's dense-DFT STFT with the golden's framing"""

    def __init__(self, kw):
        self.K, self.w = O.dft_kernel(kw["frame_len"], O.window(kw["window"], kw["frame_len"]))
        self.hop, self.center = kw["frame_hop"], kw["center"]

    def __call__(self, wav, return_polar=False):
        return O.stft_dense(wav, self.K, self.w, self.hop, center=self.center, polar=return_polar)


class _Stub(nn.Module):
    """A 'network' that returns the golden masks (the reference task only needs `enh_transform.ctx` and a call)"""

    def __init__(self, enh, masks):
        super().__init__()
        self.enh_transform = enh
        self.masks = masks

    def forward(self, mix):
        return self.masks


class _EnhShim:
    def __init__(self, ctx):
        self._ctx = ctx

    def ctx(self, name="forward_stft"):
        assert name == "forward_stft"
        return self._ctx


def _run(kw, g, enh, dev):
    from aps_b200.task import LinearFreqSaTask, MelFreqSaTask
    masks = [g["mask0"].to(dev), g["mask1"].to(dev)]
    egs = {"mix": g["mix"].to(dev), "ref": [g["ref0"].to(dev), g["ref1"].to(dev)]}
    worst = 0.0
    for name, (kind, cfg) in kw["cfgs"].items():
        task = (LinearFreqSaTask if kind == "linear" else MelFreqSaTask)(_Stub(enh, masks), **cfg).to(dev)
        with th.no_grad():
            loss = task(egs)["loss"]
        want = g["loss_" + name]
        err = abs(float(loss) - float(want)) / abs(float(want))
        assert err < FLOAT_TOL, f"{name}: {float(loss)} vs {float(want)}"
        worst = max(worst, err)
    return worst


def test_freqsa_host_logic_vs_reference():
    kw, g = load_golden("freqsa_0")
    _run(kw, g, _EnhShim(_OracleCtx(kw["enh"])), th.device("cpu"))


def test_freqsa_argument_errors():
    from aps_b200.task import LinearFreqSaTask
    kw, g = load_golden("freqsa_0")
    stub = _Stub(_EnhShim(_OracleCtx(kw["enh"])), [])
    with pytest.raises(ValueError):
        LinearFreqSaTask(stub, masking=False, truncated=1.0)            # sse.py:236-238
    task = LinearFreqSaTask(_Stub(_EnhShim(_OracleCtx(kw["enh"])), [g["mask0"]]), num_spks=2)
    with pytest.raises(RuntimeError):                                     # one output, two references
        task({"mix": g["mix"], "ref": [g["ref0"], g["ref1"]]})


@pytest.mark.gpu
def test_freqsa_gpu_vs_reference():
    from aps_b200.transform import EnhTransform
    kw, g = load_golden("freqsa_0")
    dev = th.device("cuda", 0)
    _run(kw, g, EnhTransform(**kw["enh"]).to(dev), dev)
