"""Time-domain Conv-TasNet (`sse@time_tcn`): oracle vs golden on the CPU, CUDA path vs
golden and vs the oracle at the reference's default size on the GPU."""
import pytest
import torch as th

from conftest import FLOAT_TOL, HAS_REFERENCE, golden_names, import_reference, load_golden, rel_err
from oracle import tcn as OT

DEV = "cuda:0"


def _sd(g):
    return {k[2:]: v for k, v in g.items() if k.startswith("p.")}


def _oracle(cfg, sd, mix):
    n = cfg["net"]
    return OT.time_tcn_forward(sd, mix, n["L"], n["R"], n["X"], n["num_spks"], n["norm"], n["non_linear"],
                               n.get("causal", False), n.get("skip_residual", False))


@pytest.mark.parametrize("name", golden_names("timetcn_"))
def test_oracle_timetcn_golden(name):
    cfg, g = load_golden(name)
    out = th.stack(_oracle(cfg, _sd(g), g["mix"]))
    ref = g["wav"] if g["wav"].dim() == 3 else g["wav"][None]
    assert rel_err(out, ref) < 1e-5


@pytest.mark.reference
def test_oracle_timetcn_vs_live_reference():
    import_reference()
    from aps.sse.bss.tcn import TimeConvTasNet
    th.manual_seed(5)
    kw = dict(L=20, N=32, X=2, R=2, B=16, H=24, P=3, norm="IN", num_spks=2, non_linear="sigmoid")
    net = TimeConvTasNet(**kw).eval()
    mix = 0.1 * th.randn(2, 2003)
    with th.no_grad():
        ref = net(mix)
    out = OT.time_tcn_forward(net.state_dict(), mix, 20, 2, 2, 2, "IN", "sigmoid")
    assert rel_err(th.stack(out), th.stack(ref)) < 1e-6


@pytest.mark.parametrize("name", golden_names("timetcn_"))
def test_timetcn_state_dict_loads_strict(name):
    from aps_b200.sse.bss import TimeConvTasNet
    cfg, g = load_golden(name)
    net = TimeConvTasNet(**cfg["net"])
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in _sd(g).items()}
    net.load_state_dict(_sd(g), strict=True)


def test_timetcn_argument_errors():
    from aps_b200.sse.bss import TimeConvTasNet
    with pytest.raises(ValueError, match="Unsupported nonlinear"):
        TimeConvTasNet(non_linear="tanh")
    with pytest.raises(RuntimeError, match="mixture_consistency"):
        TimeConvTasNet(mixture_consistency="fix")
    net = TimeConvTasNet(L=20, N=32, X=2, R=1, B=16, H=24).eval()
    with pytest.raises(RuntimeError, match="Expects 2D tensor"):
        net(th.zeros(100))
    with pytest.raises(RuntimeError, match="CUDA"):
        net(th.zeros(2, 400))


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names("timetcn_"))
def test_timetcn_golden_gpu(name):
    from aps_b200.sse.bss import TimeConvTasNet
    cfg, g = load_golden(name)
    net = TimeConvTasNet(**cfg["net"])
    net.load_state_dict(_sd(g), strict=True)
    net = net.to(DEV).eval()
    stack = lambda v: th.stack(v) if isinstance(v, list) else v
    wav = stack(net(g["mix"].to(DEV)))
    assert wav.shape == g["wav"].shape and rel_err(wav, g["wav"]) < FLOAT_TOL
    one = stack(net.infer(g["mix"][1].to(DEV)))
    assert one.shape == g["one"].shape and rel_err(one, g["one"]) < FLOAT_TOL


@pytest.mark.gpu
def test_timetcn_default_size_vs_oracle():
    """The reference's default network (L=20, N=256, X=8, R=4, B=256, H=512: the published TCN recipes) on 8 x 1 s at 8 kHz;
    two utterances against the CPU oracle, and batch-shard invariance."""
    from aps_b200.sse.bss import TimeConvTasNet
    th.manual_seed(21)
    net = TimeConvTasNet().eval()
    with th.no_grad():
        for name, buf in net.named_buffers():
            if name.endswith("running_var"):
                buf.copy_(0.5 + th.rand(buf.shape))
    mix = 0.1 * th.randn(8, 8000)
    dev_net = net.to(DEV)
    out = dev_net(mix.to(DEV))
    assert len(out) == 2 and out[0].shape == (8, 8000)
    rows = [0, 7]
    sd = {k: v.cpu() for k, v in dev_net.state_dict().items()}
    ref = OT.time_tcn_forward(sd, mix[rows])
    for s in range(2):
        assert rel_err(out[s][rows], ref[s]) < FLOAT_TOL
    alone = dev_net(mix[rows].to(DEV))
    assert rel_err(alone[0], out[0][rows]) < 1e-5
