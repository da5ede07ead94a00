"""The bench line's contract, checked on the committed line of the final build (profiles/r02_bench_rx_p34_final.json,
`python bench.py` on one B200) and on the N = 8 line (profiles/r02_bench_rx_8gpu_p31.json): the keys the driver reads, the
roofline and cpu_baseline objects, e2e with its copy roof, and the internal consistency of the numbers (value = units /
time, frac = achieved / peak, every parity flag true).  No GPU needed: it guards the contract against drift in bench.py's
output code, whose field names these two files were written with."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = ["profiles/r02_bench_rx_p34_final.json", "profiles/r02_bench_rx_8gpu_p31.json"]


@pytest.mark.parametrize("path", LINES)
def test_committed_bench_line_keeps_the_contract(path):
    d = json.load(open(os.path.join(ROOT, path)))
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == base["metric"]
    for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity_check"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None      # BASELINE.md publishes no number
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["gpu_launches"] > 0
    cfg = d["config"]
    assert "workload" in cfg and "configs[1]" in cfg["workload"] and "model" not in cfg
    # value = whole-job units / time of the timed region
    units = cfg["samples_per_step"] / 1e6 * d["n_gpus"]
    assert abs(d["value"] - units / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "h2d_roof_gbs", "frac_of_h2d_roof"):
        assert k in e, k
    assert e["h2d_bytes_per_step"] == cfg["samples_per_capture"] * 8 and e["d2h_bytes_per_step"] > 0
    assert 0.5 < e["frac_of_h2d_roof"] < 1.05 and e["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "alu" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] is None or r["traffic"] > r["hbm"]["algorithmic_bytes"]      # the write-through ring: DRAM bytes >> algorithmic bytes
    assert abs(r["hbm"]["frac"] - r["hbm"]["achieved"] / r["hbm"]["peak"]) < 1e-9
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and 0 < c["value"] < d["value"]
    cl = d["clocks"]
    assert cl["sm_mhz"] and cl["sm_max_mhz"] and not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # parity: the headline, every other configuration, both Viterbi sweeps
    assert d["parity_check"] is True
    assert set(d["per_config"]) == {"configs[0]", "configs[2]", "configs[3]"}
    assert all(v["parity_check"] is True for v in d["per_config"].values())
    vs = d["viterbi_sweep"]
    assert len(vs["cases"]) == 45 and all(x["parity"] for x in vs["cases"])
    assert len(vs["soft_cases"]) == 10 and all(x["parity"] for x in vs["soft_cases"])


def test_viterbi_workload_line_is_assembled(monkeypatch, capsys):
    """`bench.py --workload viterbi`: the line's assembly code runs end to end on stand-ins (no GPU here) and keeps the contract"""
    import io
    import sys
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench

    class FakeWorkload:
        nbytes_in, nbytes_out, alg_bytes, info_bits = 1000, 896, 1872, (896 - 24) * 8
        h2d, d2h = 1000, 872

        def __init__(self, mbit): self.kernel_ms = []
        def setup_gpu(self, seed): pass
        def step_resident(self, i): self.kernel_ms.append(0.5); return 872
        def step_e2e(self, i): return 872
        def check(self): return True
        def units_per_step(self): return self.info_bits / 1e6
        def describe(self): return {"workload": "viterbi_decoder stage of configs[1] (stand-in)"}
        def cpu_sample(self, hint): return 1.0, 0.1, "port"

    class FakeLib:
        n = 0
        def dvbt_b200_kernel_launches(self): FakeLib.n += 7; return FakeLib.n

    def timed(stepfn, steps, warm, w):
        for i in range(warm + steps):
            stepfn(i)
        w.last_i = warm + steps - 1
        return 2.0 * steps

    monkeypatch.setattr(bench, "ViterbiWorkload", FakeWorkload)
    monkeypatch.setattr(bench, "all_ranks_ok", lambda flag, device: bool(flag))
    args = type("A", (), dict(mbit=1.0, steps=4, warmup=3))()
    out = io.StringIO()
    rc = bench.main_viterbi(args, out, "metric", "Mbit/s (Viterbi decoded bits)", None, FakeLib(), timed, lambda: None, 6556.2, "measured", None, {})
    assert rc == 0
    d = json.loads(out.getvalue())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity_check"):
        assert k in d, k
    assert d["ms_per_step"] == 2.0 and abs(d["value"] - FakeWorkload.info_bits / 1e6 / 2e-3) < 1e-9
    assert d["roofline"]["bound"] == "alu" and abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert d["e2e"]["h2d_bytes_per_step"] == 1000 and d["cpu_baseline"]["kind"] == "port"
