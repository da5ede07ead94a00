"""CPU check of the generated ACS schedule: the numpy interpreter runs the same instruction list
the CUDA kernel compiles (gen_viterbi_acs.py) plus a model of the event logic, against the oracle."""
import numpy as np
import pytest

import viterbi_model as VM
from oracle import port as O


# (the numpy interpreter is slow; rate 7/8 - 672 byte times per block - runs for the newest schedule only: every rate and
# schedule is also covered, on the kernels' own source, by tests/test_viterbi_host_emul_cpu.py)
@pytest.mark.parametrize("rate,m,ber,schedule", [(0, 4, 0.03, "swar"), (0, 4, 0.03, "h16"), (0, 4, 0.03, "h16b"), (2, 2, 0.02, "swar"),
                                                 (2, 2, 0.02, "h16"), (2, 2, 0.02, "h16b"), (4, 6, 0.004, "h16b")])
def test_schedule_model_matches_oracle(rate, m, ber, schedule):
    k, n = O.RATE_KN[rate]
    data = np.random.default_rng(rate).integers(0, 256, 96 * k, dtype=np.uint8)
    rx = O.flip_bits(O.conv_encode(data, m, rate), m, ber, 3)
    ref = O.Viterbi(m, rate).work(rx)
    out, _ = VM.decode_stream(rx, m, rate, chunk_bytes=10 ** 9, warm=0, schedule=schedule)
    assert np.array_equal(out, ref)
    out, fix = VM.decode_stream(rx, m, rate, chunk_bytes=48, warm=1, schedule=schedule)
    assert np.array_equal(out, ref) and fix > 0  # boundary verification + repair is exact


@pytest.mark.parametrize("mod,hdr", [("gen_viterbi_acs", "viterbi_acs_gen.cuh"), ("gen_viterbi_acs_h16", "viterbi_acs_h16_gen.cuh"),
                                     ("gen_viterbi_acs_h16b", "viterbi_acs_h16b_gen.cuh")])
def test_generated_header_is_current(mod, hdr):
    import importlib
    import os
    G = importlib.import_module(mod)
    path = os.path.join(os.path.dirname(G.__file__), hdr)
    before = open(path).read()
    G.main()
    assert open(path).read() == before, "%s is stale: run gr_dvbt_b200/csrc/%s.py" % (hdr, mod)
