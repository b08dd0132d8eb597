"""Spectral-approximation and complex mapping / masking tasks (row f1: aps/task/sse.py:207-841, "sse@freq_linear_sa",
"sse@freq_mel_sa", "sse@time_linear_sa", "sse@time_mel_sa", "sse@complex_mapping", "sse@complex_masking") against values
of the live reference (tests/golden/freqsa_0.npz, timesa_0.npz, oracle/gen_golden.py).

The CPU test drives the task shells with the oracle's STFT as context (host logic: reference magnitude, masking,
distances, permutations); the GPU test uses the real EnhTransform context (fused polar STFT kernel)."""
import pytest
import torch as th
import torch.nn as nn

from conftest import load_golden, rel_err
from oracle import transform as O

FLOAT_TOL = 1e-4


class _OracleCtx(nn.Module):
    """forward_stft context on the CPU: the oracle's dense-DFT STFT with the golden's framing"""

    def __init__(self, kw):
        super().__init__()
        self.K, self.w = O.dft_kernel(kw.get("frame_len", 512), O.window(kw.get("window", "sqrthann"), kw.get("frame_len", 512)))
        self.hop, self.center = kw.get("frame_hop", 256), kw.get("center", False)

    def forward(self, wav, return_polar=False):
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


# ---- time-domain SA + complex mapping / masking ---------------------------------------------------------------------
def _run_timesa(kw, g, dev, oracle_ctx):
    from aps_b200 import task as T
    table = {"time_linear": (T.LinearTimeSaTask, "est"), "time_mel": (T.MelTimeSaTask, "est"),
             "complex_mapping": (T.ComplexMappingTask, "spec"), "complex_masking": (T.ComplexMaskingTask, "spec")}
    if oracle_ctx:
        enh = _EnhShim(_OracleCtx(kw["enh"]))
    else:
        from aps_b200.transform import EnhTransform
        enh = EnhTransform(**kw["enh"]).to(dev)
    egs = lambda: {"mix": g["mix"].clone().to(dev), "ref": [g["ref0"].clone().to(dev), g["ref1"].clone().to(dev)]}

    def check(task, name):
        if oracle_ctx and isinstance(task, T.sse.TimeSaTask):
            task.ctx = _OracleCtx(name[1])                      # the task's own framing arguments
        task = task.to(dev)
        with th.no_grad():
            loss = task(egs())["loss"]
        want = g["loss_" + name[0]]
        assert abs(float(loss) - float(want)) / abs(float(want)) < FLOAT_TOL, f"{name[0]}: {float(loss)} vs {float(want)}"

    for name, (kind, cfg) in kw["cfgs"].items():
        cls, key = table[kind]
        out = [g[key + "0"].to(dev), g[key + "1"].to(dev)]
        check(cls(_Stub(enh, out), **cfg), (name, cfg))
    task = T.LinearTimeSaTask(_Stub(enh, [g["est0"].to(dev), g["est1"].to(dev)]))
    task.pre_emphasis = 0.97
    before = g["ref0"].clone()
    check(task, ("tlin_preemph", {}))
    assert th.equal(before, g["ref0"])                          # unlike the reference, the inputs are left alone


def test_wa_task_vs_reference():
    """sse@wa is plain tensor arithmetic (no kernel of ours involved), so it is checked on the CPU."""
    from aps_b200.task import WaTask
    kw, g = load_golden("timesa_0")
    for objf in ("L1", "L2"):
        task = WaTask(_Stub(None, [g["est0"], g["est1"]]), objf=objf)
        loss = task({"mix": g["mix"], "ref": [g["ref0"], g["ref1"]]})["loss"]
        assert rel_err(loss, g["loss_wa_" + objf]) < 1e-5


def test_timesa_host_logic_vs_reference():
    kw, g = load_golden("timesa_0")
    _run_timesa(kw, g, th.device("cpu"), oracle_ctx=True)


def test_complex_masking_compressed_path_raises_like_the_reference():
    """sse.py:783-796 divides [N, F, T, 2] by [N, F, T]: a broadcast error for ordinary shapes, reproduced."""
    from aps_b200.task import ComplexMaskingTask
    kw, g = load_golden("timesa_0")
    task = ComplexMaskingTask(_Stub(_EnhShim(_OracleCtx(kw["enh"])), [g["spec0"], g["spec1"]]), compress_masks=True)
    with pytest.raises(RuntimeError):
        task({"mix": g["mix"], "ref": [g["ref0"], g["ref1"]]})


@pytest.mark.gpu
def test_timesa_gpu_vs_reference():
    kw, g = load_golden("timesa_0")
    _run_timesa(kw, g, th.device("cuda", 0), oracle_ctx=False)
