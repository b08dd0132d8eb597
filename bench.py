#!/usr/bin/env python
"""bench.py — the measurement contract of this repo.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one synthetic batch.  Workload at every N (weak scaling,
per-GPU work fixed): BASELINE.json configs[1] — AsrTransform STFT -> 80-mel fbank -> log -> per-frame
CMVN on B = 256 x 4 s @ 16 kHz per GPU ("fbank-log-cmvn", frame 400 / hop 160, hamming,
pre-emphasis 0.97, stft_mode librosa => 397 frames per utterance).  The batch is sharded per
utterance: rank r owns its own 256 utterances, there is no data-path collective; one NCCL all-reduce
carries {frames, max elapsed} after the timed region.

Prints ONE JSON line on rank 0.  `value` = frames/s with inputs resident in HBM (CUDA events on the
launching stream, max over ranks); `e2e` = the same metric through the public AsrTransform call with
pinned HOST buffers, H2D of the waveforms and D2H of the features inside the timed region;
`roofline` = algorithmic HBM bytes of the fused kernel / its event-timed duration against the measured
copy bandwidth in MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle (a port of the reference's
dense-DFT algorithm, oracle/transform.py) on a bounded sample with all host threads.

`--impl reference` times that CPU port alone (the reference itself is PyTorch-on-CPU code that cannot
travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch as th

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, SECONDS, BATCH, HOP, NFFT, MELS = 16000, 4, 256, 160, 512, 80
S = SR * SECONDS
T = (S - NFFT) // HOP + 1            # 397 (librosa mode)
CFG = dict(feats="fbank-log-cmvn", frame_len=400, frame_hop=HOP, window="hamm", pre_emphasis=0.97,
           num_mels=MELS, stft_mode="librosa")
WORKLOAD = "AsrTransform fbank-log-cmvn (400/160, hamm, preemph 0.97, librosa), B=256 x 4 s @ 16 kHz per GPU"
ALG_BYTES_PER_STEP = BATCH * S * 4 + BATCH * T * MELS * 4   # read every sample once + write 80 floats/frame
F1_DRAM_TRAFFIC_BYTES = 65502976 + 6131456                  # measured per launch (ncu, profiles/r01_fbank_s2.txt)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--workload", default="fbank", choices=["fbank", "encoder", "mvdr_tcn", "stft_istft", "dccrn"],
                    help="fbank = BASELINE configs[1] (the headline, default); the others are the remaining "
                         "single-GPU configs, reported with the same JSON schema (N = 1 only)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def cpu_port_frames_per_s(seconds: float, batch: int = 32):
    """The CPU oracle (port of the reference algorithm) on a bounded sample of the same workload."""
    from oracle.transform import AsrFeatCfg, AsrFeatures
    cores = os.cpu_count() or 1
    th.set_num_threads(cores)
    f = AsrFeatures(AsrFeatCfg(**CFG))
    g = th.Generator().manual_seed(0)
    x = 0.1 * th.randn(batch, S, generator=g)
    lens = th.full((batch,), S, dtype=th.int64)
    with th.no_grad():
        f(x, lens)                                       # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            f(x, lens)
            n += 1
            el = time.perf_counter() - t0
            if el >= seconds or n >= 200:
                break
    return batch * T * n / el, cores, f"{n} passes of B={batch} x 4 s ({el:.1f} s wall, torch CPU fp32, {cores} threads)"


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU algorithm (oracle port) on this box's host cores."""
    if rank != 0:
        return
    from oracle.transform import AsrFeatCfg, AsrFeatures
    cores = os.cpu_count() or 1
    th.set_num_threads(cores)
    batch = 32
    f = AsrFeatures(AsrFeatCfg(**CFG))
    x = 0.1 * th.randn(batch, S, generator=th.Generator().manual_seed(0))
    lens = th.full((batch,), S, dtype=th.int64)
    with th.no_grad():
        for _ in range(args.warmup):
            f(x, lens)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            f(x, lens)
        el = time.perf_counter() - t0
    val = batch * T * args.steps / el
    sample = f"each step = B={batch} x 4 s of the same workload on {cores} host threads (torch CPU fp32)"
    print(json.dumps({
        "impl": "reference", "metric": "frames/sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def _time_steps(fn, steps, warmup, dev):
    for i in range(warmup):
        fn(i)
    th.cuda.synchronize(dev)
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record(th.cuda.current_stream(dev))
    for i in range(steps):
        fn(i)
    e1.record(th.cuda.current_stream(dev))
    th.cuda.synchronize(dev)
    return e0.elapsed_time(e1)


def run_extra(args):
    """Secondary workloads (BASELINE configs[2..3] + the STFT/iSTFT pair), one GPU, inputs resident."""
    import copy
    dev = th.device("cuda", 0)
    th.cuda.set_device(dev)
    gen = th.Generator(device=dev).manual_seed(1234)
    peak_bw, peak_src = peaks()
    tf_peak = 1590.0
    try:
        tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    sampler = ClockSampler(0)
    R = 3
    if args.workload == "encoder":
        from aps_b200.asr.transformer import TransformerEncoder
        cfg = dict(arch="cfmr", input_size=80, output_proj=-1, num_layers=12, proj="conv2d",
                   proj_kwargs=dict(conv_channels=256, num_layers=3), pose="rel",
                   pose_kwargs=dict(dropout=0.1, lradius=256, rradius=256),
                   arch_kwargs=dict(att_dim=256, nhead=4, feedforward_dim=2048, att_dropout=0.1, ffn_dropout=0.1,
                                    kernel_size=15, pre_norm=False))
        th.manual_seed(0)
        net = TransformerEncoder(**copy.deepcopy(cfg)).to(dev).eval()
        B, T = 64, 398
        xs = [th.randn(B, T, 80, device=dev, generator=gen) for _ in range(R)]
        ms = _time_steps(lambda i: net(xs[i % R], None), args.steps, args.warmup, dev)
        frames, launches = B * T, 12 * 14 + 5
        flops = 6.18e9 * B                                   # SURVEY.md §8d without the vocabulary projection
        ach = flops / (ms / args.steps * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                "traffic": None, "kernel": "tc_gemm_kernel<BN> (tcgen05 3xTF32, fp32-equivalent FLOPs counted once) + "
                                           "mhsa / layernorm / dwconv kernels, CUDA-graph replay",
                "algorithmic_flops_per_step": flops, "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)"}
        wl = "conformer encoder 12L d=256 h=4 rel-pos, conv2d x3 front, forward on [64, 398, 80] fbank (configs[3])"
    elif args.workload == "mvdr_tcn":
        from aps_b200.asr.filter import MvdrBeamformer
        from aps_b200.cplx import ComplexTensor
        from aps_b200.sse.bss import FreqConvTasNet
        from aps_b200.transform import EnhTransform
        B, C = 64, 4
        enh = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, window="sqrthann")
        tcn = FreqConvTasNet(enh_transform=enh, in_features=257, num_bins=257, num_spks=1, non_linear="sigmoid").to(dev).eval()
        mvdr = MvdrBeamformer(257, att_dim=512).to(dev).eval()
        xs = [0.1 * th.randn(B, C, S, device=dev, generator=gen) for _ in range(R)]

        def step(i):
            packed, _ = tcn.enh_transform.encode(xs[i % R], None)
            mask = tcn.mask_predict(tcn.enh_transform(packed))
            return mvdr(mask.transpose(1, 2), ComplexTensor(packed[..., 0], packed[..., 1]))

        ms = _time_steps(step, args.steps, args.warmup, dev)
        frames, launches = B * (S // HOP), 1 + 1 + 56 + 5
        bytes_ = B * (C * S * 4 + C * 257 * 249 * 8 + 2056 * 249 + 8.96e6)   # STFT + feats + MVDR
        ach = bytes_ / (ms / args.steps * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak_bw, "unit": "GB/s", "frac": ach / peak_bw, "traffic": None,
                "kernel": "whole step (4-ch STFT, ref-channel feats, freq-TCN GEMMs, covar, solve, beamform); "
                          "the TCN GEMMs (2.43 GFLOP/utt, fp32 SIMT) dominate the time",
                "algorithmic_bytes_per_step": bytes_, "peak_source": peak_src}
        wl = "4-ch STFT + freq-TCN sigmoid mask + MVDR, B=64 x 4 s (configs[2]); value in 10 ms frames/s"
    elif args.workload == "dccrn":
        from aps_b200.sse.bss import DCCRN
        from aps_b200.task import SisnrTask
        from aps_b200.transform import EnhTransform
        B = 128
        enh = EnhTransform(feats="spectrogram-log-cmvn", frame_len=512, frame_hop=256, center=True)
        th.manual_seed(0)
        net = DCCRN(enh_transform=enh, cplx=True, K="3,3;3,3;3,3;3,3;3,3;3,3;3,3", S="2,1;2,1;2,1;2,1;2,1;2,1;2,1",
                    P="1,1,1,1,1,0,0", O="0,0,0,0,0,0,1", C="16,32,64,64,128,128,256", num_spks=2, rnn_resize=512,
                    non_linear="sigmoid", connection="cat").to(dev).eval()       # tests/python/test_nnet_sse.py:216-228
        task = SisnrTask(net, num_spks=2)
        xs = [th.rand(B, S, device=dev, generator=gen) for _ in range(R)]
        refs = [[0.5 * x, 0.5 * x.flip(-1)] for x in xs]
        with th.no_grad():
            ms = _time_steps(lambda i: task({"mix": xs[i % R], "ref": refs[i % R]})["loss"], args.steps, args.warmup, dev)
        Tf = S // 256 + 1                                           # STFT frames (512/256, center)
        # FLOPs of the complex (transposed) convolutions as the reference computes them (4 real convs each)
        flops, Fq, Cs = 0.0, 257, [1, 16, 32, 64, 64, 128, 128, 256]
        fqs = [Fq]
        for i, pd in enumerate([1, 1, 1, 1, 1, 0, 0]):
            Fq = (Fq + 2 * pd - 3) // 2 + 1
            fqs.append(Fq)
            flops += 2.0 * B * Fq * Tf * (2 * Cs[i + 1]) * (9 * 2 * Cs[i])
        dec_c = [512, 256, 256, 128, 128, 64, 32]                 # "cat" inputs; outputs 128,128,64,64,32,16,num_spks
        dec_o = [128, 128, 64, 64, 32, 16, 2]
        for i in range(7):
            flops += 2.0 * B * fqs[7 - i] * Tf * (2 * dec_c[i]) * (9 * 2 * dec_o[i])
        frames, launches = B * (S // HOP), 1 + 7 + 7 + 2 * 3 + 2 * (S // 256 + 1) + 2 + 6   # + one LSTM launch per STFT frame and layer
        ach = flops / (ms / args.steps * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                "traffic": None, "kernel": "tc_gemm_kernel<BN, conv / conv_transpose> (tcgen05 3xTF32 implicit GEMMs on stacked "
                                           "re/im channels; FLOPs counted as the reference computes them, incl. the zero taps "
                                           "of the transposed convolutions that the kernel skips) + fused LSTM recurrence (csrc/lstm.cu) "
                                           "(exact fp32, ~35 % of the step), narrow-output transposed conv, STFT, iSTFT x2, cmask, fused Si-SNR",
                "algorithmic_flops_per_step": flops, "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)"}
        wl = ("DCCRN (C=16..256, cat, 2 spk) forward + PIT Si-SNR on B=128 x 4 s (configs[4]); value in 10 ms frames/s")
    else:
        from aps_b200.transform.utils import STFT, iSTFT
        B = 128
        kw = dict(frame_len=512, frame_hop=256, window="sqrthann", center=True)
        f, g = STFT(**kw).to(dev), iSTFT(**kw).to(dev)
        xs = [0.1 * th.randn(B, S, device=dev, generator=gen) for _ in range(R)]
        ms = _time_steps(lambda i: g(f(xs[i % R])), args.steps, args.warmup, dev)
        frames, launches = B * (S // HOP), 2
        bytes_ = 2 * 3080.0 * B * 251
        ach = bytes_ / (ms / args.steps * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak_bw, "unit": "GB/s", "frac": ach / peak_bw, "traffic": None,
                "kernel": "frontend_kernel<256,1,1> + istft_kernel<256>", "algorithmic_bytes_per_step": bytes_,
                "peak_source": peak_src}
        wl = "STFT -> iSTFT round trip 512/256 center, B=128 x 4 s (the transform pair of configs[4]); 10 ms frames/s"
    clocks = sampler.stop()
    print(json.dumps({
        "metric": "frames/sec", "value": frames * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "l2": f"{R} distinct resident input batches rotate"},
        "roofline": roof, "cpu_baseline": None, "e2e": None, "gpu_launches": launches * (args.steps + args.warmup),
        "clocks": clocks}), flush=True)


def main():
    args = parse()
    if args.workload != "fbank":
        return run_extra(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    import torch.distributed as dist
    from aps_b200 import _lib
    from aps_b200.transform import AsrTransform
    assert th.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()                                          # fail loudly if the CUDA library is missing
    transform = AsrTransform(**CFG).to(dev).eval()
    lens = th.full((BATCH,), S, dtype=th.int64)          # host-side lengths (the loader's egs["src_len"])

    # ---- inputs: R distinct resident batches so a step never finds its input in the 126 MB L2 ------------
    R = 4
    gen = th.Generator(device=dev).manual_seed(1234 + rank)
    wavs = [0.1 * th.randn(BATCH, S, device=dev, generator=gen) for _ in range(R)]
    launches = 0

    # `value` times the fused kernel back to back through the shell function that wraps the C ABI call
    # (no NaN-guard host sync between launches); the public AsrTransform.forward, which adds the reference's
    # check_valid() host sync per call, is what `e2e` times.
    from aps_b200.transform.asr import _match_tail, fused_wave_features
    layers = list(transform.transform)
    tail = _match_tail(layers, 1)
    assert tail is not None and tail[5] == len(layers), "the whole chain must map onto the fused kernel"

    def step(i):
        return fused_wave_features(layers[0], wavs[i % R], tail, rescale=False, utt_preemph=0.0)

    stream = th.cuda.current_stream(dev)
    for i in range(args.warmup):
        out = step(i)
    th.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    # ---- device-resident timing: K steps, CUDA events on the launching stream ---------------------------
    th.cuda.synchronize(dev)
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        out = step(i)
        launches += 1
    e1.record(stream)
    th.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    frames = BATCH * out.shape[1] * args.steps
    assert out.shape == (BATCH, T, MELS)

    # ---- end to end through the public call with pinned host buffers --------------------------------------
    # Three streams, double buffered: H2D of batch i+1 and D2H of batch i-1 overlap the kernel + NaN guard of
    # batch i (the guard's host sync only waits for the compute stream).  Every step still moves its own
    # 65.5 MB in and 32.5 MB out inside the timed region.
    h_in = [th.empty(BATCH, S, pin_memory=True).copy_(w) for w in wavs[:2]]
    h_out = [th.empty(BATCH, T, MELS, pin_memory=True) for _ in range(2)]
    d_in = [th.empty(BATCH, S, device=dev) for _ in range(2)]
    s_in, s_out = th.cuda.Stream(dev), th.cuda.Stream(dev)
    ev_in = [th.cuda.Event() for _ in range(2)]
    ev_used = [th.cuda.Event() for _ in range(2)]
    ev_out = [th.cuda.Event() for _ in range(2)]

    def upload(i):
        b = i % 2
        with th.cuda.stream(s_in):
            s_in.wait_event(ev_used[b])                  # the kernel that last read d_in[b] is done
            d_in[b].copy_(h_in[b], non_blocking=True)
            ev_in[b].record(s_in)

    def e2e_run(k):
        for b in range(2):
            ev_used[b].record(stream)
            ev_out[b].record(s_out)
        upload(0)
        for i in range(k):
            b = i % 2
            if i + 1 < k:
                upload(i + 1)
            stream.wait_event(ev_in[b])
            feats, nf = transform(d_in[b], lens)         # public call: kernel + check_valid (host sync on this stream)
            ev_used[b].record(stream)
            feats.record_stream(s_out)
            with th.cuda.stream(s_out):
                s_out.wait_event(ev_used[b])
                h_out[b].copy_(feats, non_blocking=True)
                ev_out[b].record(s_out)
        s_out.synchronize()

    e2e_run(3)
    th.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    k2 = max(6, min(args.steps, 20))
    t0 = time.perf_counter()
    e2e_run(k2)
    th.cuda.synchronize(dev)
    e2e_ms = 1e3 * (time.perf_counter() - t0)            # host wall clock brackets all three streams
    launches += k2
    clocks = sampler.stop() if sampler is not None else None

    # ---- aggregate over ranks: max time, sum frames -----------------------------------------------------------
    stats = th.tensor([ms, e2e_ms, float(frames), float(BATCH * T * k2)], dtype=th.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms, frames, e2e_frames = float(mx[0]), float(mx[1]), float(sm[2]), float(sm[3])
    else:
        e2e_frames = float(stats[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = frames / (ms * 1e-3)
    peak, peak_src = peaks()
    kern_ms = ms / args.steps                            # one kernel per step
    achieved = ALG_BYTES_PER_STEP / (kern_ms * 1e-3) / 1e9
    cpu_v, cores, sample = cpu_port_frames_per_s(args.cpu_seconds) if world >= 1 else (None, 0, "")
    line = {
        "metric": "frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": BATCH * T, "parallelism": f"dp{world} (batch shard, no data-path collective)",
                   "l2": f"{R} distinct resident input batches rotate (4 x 98 MB in+out > 126 MB L2), no flush inside the timed region"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": F1_DRAM_TRAFFIC_BYTES,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this "
                                       "kernel on this workload (profiles/r01_fbank_s2.txt: 65.50 MB read + 6.13 MB written; "
                                       "the rest of the 32.5 MB output is still in L2 when the kernel ends)",
                     "kernel": "frontend_kernel<256,0,5,FULL> (fused framing+preemph+window+rFFT+|X|+mel+log+cmvn)",
                     "algorithmic_bytes_per_launch": ALG_BYTES_PER_STEP, "peak_source": peak_src,
                     "kernel_ms": kern_ms},
        "cpu_baseline": {"value": cpu_v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": e2e_frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": BATCH * S * 4,
                "d2h_bytes_per_step": BATCH * T * MELS * 4, "steps": k2,
                "path": "AsrTransform.forward on pinned-host batches: H2D | fused kernel + NaN guard | D2H on three streams, double buffered"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
