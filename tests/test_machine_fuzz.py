"""CPU: the state machine of the CUDA path (csrc/machine.h, compiled for the host by tests/host_machine_harness.cpp, the
oracle's alignment functions in place of the kernels) against the oracle on the randomised workloads of
tests/test_gpu_fuzz.py (other seeds).  300 further seeds were run once by hand (all equal); 40 stay in the suite."""
import numpy as np
import pytest

import oracle_lib as ol
import synth
from test_gpu_fuzz import _case


@pytest.mark.parametrize("k", range(100, 140))
def test_host_machine_matches_oracle(k):
    cfg, knobs = _case(k)
    fd = ol.FlatData(synth.generate(synth.SynthConfig(**cfg)))
    best = np.array([fd.cands["score"][s["cand_begin"]:s["cand_begin"] + s["cand_count"]].max() if s["cand_count"] else 0.0
                     for s in fd.segs])
    if (best < 0).any():
        pytest.skip("synthetic scores went negative (outside the reference's defined behaviour)")
    want = ol.oracle_predict(fd)
    got = ol.host_machine_predict(fd, spec_k=knobs["lookahead"] if knobs["lookahead"] >= 0 else 3)
    got = got[0] if isinstance(got, tuple) else got
    assert ol.results_equal(want, got) == [], cfg
