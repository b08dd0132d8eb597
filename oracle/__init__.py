"""CPU oracle for the APS hot path — TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithms of the reference (funcwj/aps)
for the path named in BASELINE.json.  It exists to CHECK the CUDA product in
`aps_b200/` and to serve as the timed CPU baseline in `bench.py`.

Rules (enforced by tests/test_layout.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py` (its `cpu_baseline`
    leg and `--impl reference`) may import anything from here;
  * nothing under `aps_b200/` imports `oracle`;
  * the product has no CPU fallback: it raises if the CUDA library is missing.

Pinning: every function in here is validated against the LIVE reference imported
from /root/reference (through the stand-in packages in `oracle/ref_shims/`) by
`tests/test_oracle_vs_reference.py` when that tree is present, and against the
committed fixtures in `tests/golden/` (generated from the live reference by
`oracle/gen_golden.py`) everywhere else.  The one part that is *parity unpinned*
is the mel filterbank VALUES: they come from librosa==0.8.1 (requirements.txt:4),
which is neither vendored in the reference nor installed here; `oracle.transform.
mel_filterbank` restates librosa's published HTK formula.
"""
