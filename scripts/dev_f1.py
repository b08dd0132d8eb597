#!/usr/bin/env python
"""Kernel-only timing of the F1 feature kernel on the bench workload (development aid; bench.py is the contract)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import bench
from aps_b200.transform import AsrTransform
from aps_b200.transform.asr import _match_tail, fused_wave_features

dev = th.device("cuda", 0)
tr = AsrTransform(**bench.CFG).to(dev).eval()
layers = list(tr.transform)
tail = _match_tail(layers, 1)
gen = th.Generator(device=dev).manual_seed(1234)
wavs = [0.1 * th.randn(bench.BATCH, bench.S, device=dev, generator=gen) for _ in range(4)]
for i in range(5):
    out = fused_wave_features(layers[0], wavs[i % 4], tail, rescale=False, utt_preemph=0.0)
th.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        out = fused_wave_features(layers[0], wavs[i % 4], tail, rescale=False, utt_preemph=0.0)
    e1.record(); th.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 50 * 1e3)
print(f"F1 {os.environ.get('APS_B200_LIB', 'default')}: {best:.1f} us/launch  checksum {float(out.double().sum()):.6f}")
