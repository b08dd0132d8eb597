#!/usr/bin/env python
"""bench.py — the measurement contract of this repo.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one synthetic batch.  DEFAULT workload (`asr_encoder`): the chain
BASELINE.json's metric names — "frames/sec through transform+encoder" — i.e. the reference's
`ASREncoderBase._training_prep` (aps/asr/ctc.py:113-134): waveform [64, 64000] per GPU -> AsrTransform
(fbank-log-cmvn, BASELINE configs[1]'s transform) -> TransformerEncoder (conformer 12L, d=256, 4 heads, rel-pos,
conv2d x3 front, BASELINE configs[3]) through the PUBLIC modules of this package.  Weak scaling: every rank owns its own
64 utterances (batch shard, no data-path collective); one NCCL all-reduce of {max elapsed, sum frames} after the region.

One JSON line on rank 0:
  value        frames/s (10 ms transform frames), inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e          the same through the same public calls from PINNED HOST waveforms: H2D of the step's waveforms and D2H of
               the step's encoder output inside the timed region (three streams, double buffered)
  roofline     the dominant kernel of the step.  Tensor-bound workloads: tc_gemm_kernel (tcgen05 3xTF32) — every launch of
               one eager step is bracketed by CUDA events (ops.PROFILE): achieved = MMA work (3 TF32 passes x algorithmic
               FLOPs) / summed launch durations, peak = dense TF32 = MEASURED_PEAKS bf16_tflops / 2
  roofline_f1  second leg: the fused STFT->fbank kernel against the measured HBM copy bandwidth
  cpu_baseline the STOCK reference (oracle/_ref, staged by oracle/build_ref.sh; kind "reference") — or the oracle port when
               that tree is absent (kind "port") — on the host cores, bounded sample, rank 0 at N = 1 only
`--impl reference` times that CPU implementation alone on the same config (same builder code, reference classes).
Secondary workloads (`--workload fbank|encoder|mvdr_tcn|stft_istft|dccrn`) use the same runner and schema.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

import torch as th

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, SECONDS, HOP, NFFT, MELS = 16000, 4, 160, 512, 80
S = SR * SECONDS
T_ASR = (S - NFFT) // HOP + 1            # 397 frames per utterance (librosa mode, a3)
ASR_CFG = dict(feats="fbank-log-cmvn", frame_len=400, frame_hop=HOP, window="hamm", pre_emphasis=0.97,
               num_mels=MELS, stft_mode="librosa")
ENC_CFG = dict(arch="cfmr", input_size=80, output_proj=-1, num_layers=12, proj="conv2d",
               proj_kwargs=dict(conv_channels=256, num_layers=3), pose="rel",
               pose_kwargs=dict(dropout=0.1, lradius=256, rradius=256),
               arch_kwargs=dict(att_dim=256, nhead=4, feedforward_dim=2048, att_dropout=0.1, ffn_dropout=0.1,
                                kernel_size=15, pre_norm=False))
DCCRN_CFG = dict(cplx=True, K="3,3;3,3;3,3;3,3;3,3;3,3;3,3", S="2,1;2,1;2,1;2,1;2,1;2,1;2,1", P="1,1,1,1,1,0,0",
                 O="0,0,0,0,0,0,1", C="16,32,64,64,128,128,256", num_spks=2, rnn_resize=512, non_linear="sigmoid",
                 connection="cat")                                  # tests/python/test_nnet_sse.py:216-228
F1_DRAM_TRAFFIC_B256 = 65502976 + 6131456     # ncu --set full, B=256 launch (profiles/r01_fbank_s2.txt); see roofline_f1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--workload", default="asr_encoder",
                    choices=["asr_encoder", "fbank", "encoder", "mvdr_tcn", "stft_istft", "dccrn"],
                    help="asr_encoder (default) = the BASELINE metric: waveform -> AsrTransform -> conformer encoder, "
                         "B=64 per GPU; the others are the remaining BASELINE configs with the same JSON schema")
    return ap.parse_args()


def peaks():
    hbm, tf, src = 6650.0, 1590.0, "fallback (B200_PROFILING.md)"
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            hbm, tf, src = float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return hbm, tf, src


# ------------------------------------------------------------------------------------------ implementations
def namespace(impl: str):
    """The classes a workload is built from: this package ('ours') or the stock reference staged in oracle/_ref
    ('reference').  The SAME builder code runs on both — the surfaces are drop-in compatible."""
    ns = types.SimpleNamespace(impl=impl)
    if impl == "ours":
        from aps_b200.asr.filter import MvdrBeamformer
        from aps_b200.asr.transformer import TransformerEncoder
        from aps_b200.cplx import ComplexTensor
        from aps_b200.sse.bss import DCCRN, FreqConvTasNet
        from aps_b200.task import SisnrTask
        from aps_b200.transform import AsrTransform, EnhTransform
        from aps_b200.transform.utils import STFT, iSTFT
    else:
        ref = os.path.join(ROOT, "oracle", "_ref")
        if not os.path.isdir(os.path.join(ref, "aps")):
            return None
        if ref not in sys.path:
            sys.path.insert(0, ref)
        import warnings
        warnings.filterwarnings("ignore")
        from aps.asr.filter.mvdr import MvdrBeamformer
        from aps.asr.transformer.encoder import TransformerEncoder
        from aps.cplx import ComplexTensor
        from aps.sse.bss.dccrn import DCCRN
        from aps.sse.bss.tcn import FreqConvTasNet
        from aps.task.sse import SisnrTask
        from aps.transform import AsrTransform, EnhTransform
        from aps.transform.utils import STFT, iSTFT
    for k, v in dict(locals()).items():
        if k not in ("ns", "impl", "ref", "warnings"):
            setattr(ns, k, v)
    return ns


class Workload:
    """batch: utterances per GPU and step; frames: metric units per step and GPU; rotate: distinct resident input
    batches (so that a step never finds its input in the 126 MB L2)."""
    name = desc = ""
    batch, frames, rotate = 0, 0, 4
    tensor_bound = True

    def inputs(self, gen, batch):                       # -> tuple of CPU tensors
        raise NotImplementedError

    def build(self, ns, dev):                           # -> step(*inputs on dev) -> tensor or tuple of tensors
        raise NotImplementedError

    def hbm_bytes(self):                                # algorithmic bytes per step (HBM-bound workloads)
        return None


class AsrEncoder(Workload):
    name = "asr_encoder"
    batch, frames, rotate = 64, 64 * T_ASR, 8
    desc = ("waveform [64, 64000] -> AsrTransform fbank-log-cmvn (400/160, hamm, preemph 0.97, librosa; configs[1]) -> "
            "conformer encoder 12L d=256 h=4 ffn 2048 rel-pos, conv2d x3 front (configs[3]), B=64 x 4 s per GPU; "
            "the aps/asr/ctc.py:113-134 chain; frames = 10 ms transform frames (397 per utterance)")

    def inputs(self, gen, batch):
        return (0.1 * th.randn(batch, S, generator=gen), th.full((batch,), S, dtype=th.int64))

    def build(self, ns, dev):
        import copy
        th.manual_seed(0)
        tf = ns.AsrTransform(**ASR_CFG).to(dev).eval()
        enc = ns.TransformerEncoder(**copy.deepcopy(ENC_CFG)).to(dev).eval()
        self.modules = (tf, enc)

        def step(wav, lens):
            lens = lens.cpu() if ns.impl == "ours" else lens          # the loader's lengths are host tensors
            feats, nfr = tf(wav, lens)
            out, _ = enc(feats, nfr)
            return out
        return step


class Fbank(Workload):
    name = "fbank"
    batch, frames, rotate = 256, 256 * T_ASR, 4
    tensor_bound = False
    desc = "AsrTransform fbank-log-cmvn (400/160, hamm, preemph 0.97, librosa), B=256 x 4 s @ 16 kHz per GPU (configs[1])"

    def inputs(self, gen, batch):
        return (0.1 * th.randn(batch, S, generator=gen), th.full((batch,), S, dtype=th.int64))

    def build(self, ns, dev):
        tf = ns.AsrTransform(**ASR_CFG).to(dev).eval()
        self.modules = (tf,)
        return lambda wav, lens: tf(wav, lens.cpu() if ns.impl == "ours" else lens)[0]

    def hbm_bytes(self):
        return self.batch * S * 4 + self.batch * T_ASR * MELS * 4


class Encoder(Workload):
    name = "encoder"
    batch, frames, rotate = 64, 64 * 398, 4
    desc = "conformer encoder 12L d=256 h=4 rel-pos, conv2d x3 front, forward on [64, 398, 80] fbank (configs[3] alone)"

    def inputs(self, gen, batch):
        return (th.randn(batch, 398, 80, generator=gen),)

    def build(self, ns, dev):
        import copy
        th.manual_seed(0)
        enc = ns.TransformerEncoder(**copy.deepcopy(ENC_CFG)).to(dev).eval()
        self.modules = (enc,)
        return lambda x: enc(x, None)[0]


class MvdrTcn(Workload):
    name = "mvdr_tcn"
    batch, frames, rotate = 64, 64 * (S // HOP), 3
    desc = ("4-ch STFT (512/256 sqrthann) + ref-channel log-spectrogram-cmvn + freq-TCN sigmoid mask + MVDR (att 512), "
            "B=64 x 4 ch x 4 s (configs[2]); frames = samples / 160")

    def inputs(self, gen, batch):
        return (0.1 * th.randn(batch, 4, S, generator=gen),)

    def build(self, ns, dev):
        th.manual_seed(0)
        enh = ns.EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, window="sqrthann")
        tcn = ns.FreqConvTasNet(enh_transform=enh, in_features=257, num_bins=257, num_spks=1,
                                non_linear="sigmoid").to(dev).eval()
        mvdr = ns.MvdrBeamformer(257, att_dim=512).to(dev).eval()
        self.modules = (tcn, mvdr)

        def step(wav):
            packed, _ = tcn.enh_transform.encode(wav, None)
            feats = tcn.enh_transform(packed)
            mask = tcn.mask_predict(feats)                               # N x F x T (aps/sse/bss/tcn.py:458-469)
            y = mvdr(mask.transpose(1, 2), ns.ComplexTensor(packed[..., 0], packed[..., 1]))
            return (y.real, y.imag)
        return step


class StftIstft(Workload):
    name = "stft_istft"
    batch, frames, rotate = 128, 128 * (S // HOP), 4
    tensor_bound = False
    desc = "STFT -> iSTFT round trip 512/256 sqrthann center, B=128 x 4 s (the transform pair of configs[4]); frames = samples / 160"

    def inputs(self, gen, batch):
        return (0.1 * th.randn(batch, S, generator=gen),)

    def build(self, ns, dev):
        kw = dict(frame_len=512, frame_hop=256, window="sqrthann", center=True)
        f, g = ns.STFT(**kw).to(dev), ns.iSTFT(**kw).to(dev)
        self.modules = (f, g)
        return lambda x: g(f(x))

    def hbm_bytes(self):
        return 2 * 3080.0 * self.batch * 251


class Dccrn(Workload):
    name = "dccrn"
    batch, frames, rotate = 128, 128 * (S // HOP), 3
    desc = "DCCRN (C=16..256, cat, 2 spk, complex LSTM 512) forward + PIT Si-SNR on B=128 x 4 s (configs[4]); frames = samples / 160"

    def inputs(self, gen, batch):
        x = th.rand(batch, S, generator=gen)
        return (x, 0.5 * x, 0.5 * x.flip(-1))

    def build(self, ns, dev):
        th.manual_seed(0)
        enh = ns.EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, center=True)
        net = ns.DCCRN(enh_transform=enh, **DCCRN_CFG).to(dev).eval()
        task = ns.SisnrTask(net, num_spks=2)
        self.modules = (task,)
        return lambda mix, r0, r1: task({"mix": mix, "ref": [r0, r1]})["loss"].reshape(1)


WORKLOADS = {w.name: w for w in (AsrEncoder, Fbank, Encoder, MvdrTcn, StftIstft, Dccrn)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def _port_step(wl):
    """Fallback when oracle/_ref is absent: the oracle port (asr_encoder / fbank / encoder only)."""
    import copy

    from oracle.encoder import EncoderOracle
    from oracle.transform import AsrFeatCfg, AsrFeatures
    if wl.name not in ("asr_encoder", "fbank", "encoder"):
        return None
    feats = AsrFeatures(AsrFeatCfg(**ASR_CFG))
    enc = None
    if wl.name != "fbank":
        from aps_b200.asr.transformer import TransformerEncoder   # parameter container only: no kernel is launched
        th.manual_seed(0)
        cfg = copy.deepcopy(ENC_CFG)
        sd = TransformerEncoder(**copy.deepcopy(ENC_CFG)).state_dict()
        enc = EncoderOracle(cfg, sd)
    if wl.name == "encoder":
        return lambda x: enc.forward(x, None)[0]
    if wl.name == "fbank":
        return lambda wav, lens: feats(wav, lens)[0]

    def step(wav, lens):
        f, n = feats(wav, lens)
        return enc.forward(f, n)[0]
    return step


def cpu_arm(wl):
    """-> (step callable on CPU tensors, kind, cores)"""
    cores = os.cpu_count() or 1
    th.set_num_threads(cores)
    ns = namespace("reference")
    if ns is not None:
        return wl.build(ns, th.device("cpu")), "reference", cores
    return _port_step(wl), "port", cores


def cpu_baseline(wl, seconds: float):
    """Bounded sample of the same workload on the host cores: passes of a reduced batch for about `seconds`."""
    step, kind, cores = cpu_arm(wl)
    if step is None:
        return {"value": None, "unit": "frames/s", "cores": cores, "kind": kind,
                "sample": "no CPU implementation of this workload without oracle/_ref"}
    b = max(1, wl.batch // 8)
    inp = wl.inputs(th.Generator().manual_seed(0), b)
    with th.no_grad():
        t0 = time.perf_counter()
        step(*[t.clone() for t in inp])                                  # warm-up (also sizes the sample)
        first = time.perf_counter() - t0
        n, t0 = 0, time.perf_counter()
        while True:
            step(*[t.clone() for t in inp])
            n += 1
            el = time.perf_counter() - t0
            if el >= seconds or el + first > 2.5 * seconds or n >= 200:
                break
    frames = wl.frames * b / wl.batch
    impl = "stock reference modules (oracle/_ref)" if kind == "reference" else "oracle port"
    return {"value": frames * n / el, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"{n} passes of B={b} of the same workload ({el:.1f} s wall, {impl}, torch CPU fp32, {cores} threads)"}


def config_of(wl, world):
    return {"workload": wl.desc, "frames_per_step_per_gpu": wl.frames, "batch_per_gpu": wl.batch,
            "parallelism": f"dp{world} (batch shard per utterance, weights replicated, no data-path collective)",
            "l2": f"{wl.rotate} distinct input batches rotate; activations + split weights of a step exceed the 126 MB L2"}


def run_reference(args, wl, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the same config, all host threads, rank 0 only."""
    if rank != 0:
        return
    step, kind, cores = cpu_arm(wl)
    if step is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not staged and the oracle has no port of this workload"}))
        return
    # the default workload runs the FULL per-GPU batch every step (same config); the heavy secondary ones a bounded sample
    b = wl.batch if wl.name in ("asr_encoder", "fbank", "encoder", "stft_istft") else max(1, wl.batch // 8)
    inp = wl.inputs(th.Generator().manual_seed(0), b)
    with th.no_grad():
        for _ in range(args.warmup):
            step(*[t.clone() for t in inp])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(*[t.clone() for t in inp])
        el = time.perf_counter() - t0
    val = wl.frames * (b / wl.batch) * args.steps / el
    sample = (f"each step = B={b} of the same workload on {cores} host threads (torch CPU fp32, "
              f"{'stock reference modules from oracle/_ref' if kind == 'reference' else 'oracle port'})")
    print(json.dumps({
        "impl": "reference", "metric": "frames/sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(wl, world),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def _as_tuple(o):
    return tuple(o) if isinstance(o, (tuple, list)) else (o,)


def _set_graphs(wl, flag):
    for m in wl.modules:
        for sub in m.modules():
            if hasattr(sub, "use_graphs"):
                sub.use_graphs = flag


def kernel_leg(wl, step, batches, dev, peak_tf, peak_hbm, peak_src, ms_per_step):
    """Roofline of the dominant kernel, measured live.  One eager step (no CUDA graph) records every tensor-core GEMM
    launch as a replayable C-ABI call (ops.PROFILE); the recorded calls alone — same buffers, same order — are then
    captured into a CUDA graph and the replay is timed with CUDA events: the device time of all tc_gemm_kernel launches
    of one step, free of host gaps (event pairs around eager launches of 10-30 us kernels measure the Python call).
    Also returns the C-ABI launch calls per step."""
    from aps_b200 import _lib, ops
    _set_graphs(wl, False)
    with th.no_grad():
        step(*batches[0])
        th.cuda.synchronize(dev)
        c0 = _lib.CALLS
        ops.PROFILE = []
        step(*batches[1 % len(batches)])
        th.cuda.synchronize(dev)
        recs, ops.PROFILE = ops.PROFILE, None
        calls = _lib.CALLS - c0
    _set_graphs(wl, True)
    if wl.tensor_bound:
        n = len(recs)
        if not n:
            return None, calls
        flops = sum(r[1] for r in recs)
        for r in recs:                                   # warm (tensor maps, attributes) outside the capture
            r[2]()
        th.cuda.synchronize(dev)
        graph = th.cuda.CUDAGraph()
        with th.cuda.graph(graph):
            for r in recs:
                r[2]()
        ms = 1e30
        for _ in range(6):
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record(th.cuda.current_stream(dev))
            graph.replay()
            e1.record(th.cuda.current_stream(dev))
            th.cuda.synchronize(dev)
            ms = min(ms, e0.elapsed_time(e1))
        mma_tf = 3.0 * flops / (ms * 1e-3) / 1e12
        peak = peak_tf / 2.0
        return {"bound": "tensor", "achieved": mma_tf, "peak": peak, "unit": "TFLOP/s", "frac": mma_tf / peak,
                "traffic": None,
                "kernel": f"tc_gemm_kernel<BN,MODE> (tcgen05.mma kind::tf32, 3 passes hi*hi + hi*lo + lo*hi): all {n} launches of one step",
                "launches_per_step": n, "kernel_ms_per_step": ms, "share_of_step": ms / ms_per_step,
                "algorithmic_flops_per_step": flops, "fp32_equivalent_tflops": flops / (ms * 1e-3) / 1e12,
                "how": "achieved = 3 x algorithmic FLOPs (the TF32 MMA passes that fp32 parity needs) / device time of the step's "
                       "GEMM launches replayed alone from a CUDA graph (CUDA events, best of 6); peak = dense TF32 = bf16_tflops / 2",
                "peak_source": peak_src + " bf16_tflops (burst) / 2"}, calls
    nbytes = wl.hbm_bytes()
    ach = nbytes / (ms_per_step * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak_hbm, "unit": "GB/s", "frac": ach / peak_hbm, "traffic": None,
            "kernel": "whole step (HBM-bound kernels only)", "algorithmic_bytes_per_step": nbytes,
            "peak_source": peak_src + " hbm_gbs"}, calls


def f1_leg(wl, batches, dev, peak_hbm, peak_src):
    """Second leg: the fused STFT -> |X| -> mel -> log -> CMVN kernel alone, launched back to back through the shell
    function that wraps the C-ABI call (no NaN-guard sync), on this workload's own waveform batches."""
    from aps_b200.transform.asr import _match_tail, fused_wave_features
    tf = wl.modules[0]
    layers = list(tf.transform)
    tail = _match_tail(layers, 1)
    assert tail is not None and tail[5] == len(layers), "the whole chain must map onto the fused kernel"
    wavs = [b[0] for b in batches]
    B = wavs[0].shape[0]
    reps = 40
    for i in range(5):
        fused_wave_features(layers[0], wavs[i % len(wavs)], tail, rescale=False, utt_preemph=0.0)
    th.cuda.synchronize(dev)
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record(th.cuda.current_stream(dev))
    for i in range(reps):
        fused_wave_features(layers[0], wavs[i % len(wavs)], tail, rescale=False, utt_preemph=0.0)
    e1.record(th.cuda.current_stream(dev))
    th.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    nbytes = B * S * 4 + B * T_ASR * MELS * 4
    ach = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak_hbm, "unit": "GB/s", "frac": ach / peak_hbm,
            "traffic": F1_DRAM_TRAFFIC_B256 * B // 256,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the `ncu --set full` capture of this kernel at "
                              "B=256 (profiles/), scaled by B/256",
            "kernel": "frontend_kernel<256,0,...> (fused framing+preemph+window+rFFT+|X|+mel+log+cmvn)",
            "algorithmic_bytes_per_launch": nbytes, "kernel_ms": ms, "batch": B, "peak_source": peak_src + " hbm_gbs"}


def run_ours(args, wl, rank, world, local):
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    import torch.distributed as dist

    from aps_b200 import _lib
    assert th.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()                                          # fail loudly if the CUDA library is missing
    step = wl.build(namespace("ours"), dev)
    gen = th.Generator().manual_seed(1234 + rank)
    host = [wl.inputs(gen, wl.batch) for _ in range(wl.rotate)]
    # integer tensors (lengths) stay on the host, as the loader delivers them
    batches = [tuple(t.to(dev) if t.is_floating_point() else t for t in b) for b in host]
    stream = th.cuda.current_stream(dev)
    R = wl.rotate

    with th.no_grad():
        for i in range(args.warmup):
            out = step(*batches[i % R])
        th.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        # ---- device-resident timing: K steps, CUDA events on the launching stream -----------------------------
        th.cuda.synchronize(dev)
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            out = step(*batches[i % R])
        e1.record(stream)
        th.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)

        # ---- end to end from pinned host buffers ----------------------------------------------------------------
        # H2D of batch i+1 and D2H of result i-1 overlap the compute of batch i (three streams, double buffered);
        # every step moves its own inputs in and its own result out inside the timed region.
        outs = _as_tuple(out)
        fl = [j for j, t in enumerate(host[0]) if t.is_floating_point()]
        h_in = [[host[b][j].pin_memory() for j in fl] for b in range(2)]
        d_in = [[th.empty_like(host[0][j], device=dev) for j in fl] for _ in range(2)]
        h_out = [[th.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs] for _ in range(2)]
        h2d = sum(t.numel() * t.element_size() for t in h_in[0])
        d2h = sum(t.numel() * t.element_size() for t in h_out[0])
        s_in, s_out = th.cuda.Stream(dev), th.cuda.Stream(dev)
        ev_in = [th.cuda.Event() for _ in range(2)]
        ev_used = [th.cuda.Event() for _ in range(2)]

        def upload(i):
            b = i % 2
            with th.cuda.stream(s_in):
                s_in.wait_event(ev_used[b])              # the step that last read d_in[b] is done
                for dst, src in zip(d_in[b], h_in[b]):
                    dst.copy_(src, non_blocking=True)
                ev_in[b].record(s_in)

        def e2e_run(k):
            for b in range(2):
                ev_used[b].record(stream)
            upload(0)
            for i in range(k):
                b = i % 2
                if i + 1 < k:
                    upload(i + 1)
                stream.wait_event(ev_in[b])
                args_ = list(host[b])
                for j, t in zip(fl, d_in[b]):
                    args_[j] = t
                res = _as_tuple(step(*args_))            # the public calls
                ev_used[b].record(stream)
                with th.cuda.stream(s_out):
                    s_out.wait_event(ev_used[b])
                    for dst, src in zip(h_out[b], res):
                        src.record_stream(s_out)
                        dst.copy_(src, non_blocking=True)
            s_out.synchronize()

        e2e_run(3)
        th.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        k2 = max(6, min(args.steps, 20))
        t0 = time.perf_counter()
        e2e_run(k2)
        th.cuda.synchronize(dev)
        e2e_ms = 1e3 * (time.perf_counter() - t0)        # host wall clock brackets all three streams
        clocks = sampler.stop() if sampler is not None else None

    # ---- aggregate over ranks: max time, sum frames -------------------------------------------------------------
    stats = th.tensor([ms, e2e_ms, float(wl.frames * args.steps), float(wl.frames * k2)], dtype=th.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms, frames, e2e_frames = float(mx[0]), float(mx[1]), float(sm[2]), float(sm[3])
    else:
        frames, e2e_frames = float(stats[2]), float(stats[3])
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak_hbm, peak_tf, peak_src = peaks()
    roof, calls = kernel_leg(wl, step, batches, dev, peak_tf, peak_hbm, peak_src, ms / args.steps)
    roof_f1 = f1_leg(wl, batches, dev, peak_hbm, peak_src) if wl.name in ("asr_encoder", "fbank") else None
    if wl.name == "fbank":
        roof, roof_f1 = roof_f1, None
    if world > 1:
        dist.barrier()
    cpu = cpu_baseline(wl, args.cpu_seconds) if world == 1 else None
    line = {
        "metric": "frames/sec", "value": frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(wl, world),
        "roofline": roof, "roofline_f1": roof_f1, "cpu_baseline": cpu,
        "e2e": {"value": e2e_frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": k2, "ms_per_step": e2e_ms / k2,
                "path": "public module calls on pinned-host batches: H2D of the inputs | step | D2H of the result, "
                        "three streams, double buffered"},
        "gpu_launches": calls * (args.steps + args.warmup + k2 + 3),
        "gpu_launches_note": f"{calls} C-ABI launch calls per step counted on one eager step (each is >= 1 kernel; the encoder "
                             "replays them from a CUDA graph)",
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]()
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)
    return run_ours(args, wl, rank, world, local)


if __name__ == "__main__":
    main()
