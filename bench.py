#!/usr/bin/env python3
"""bench.py — throughput of the B200 DVB-T receive hot path (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload W]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of
synthetic input that is already resident in HBM (`value`), and the same pass through the
C ABI with HOST buffers, copies inside the timed region (`e2e`).  `--impl reference` times
the reference's own CPU code (oracle/_ref when it was built, else the oracle port) on all
host cores on a bounded sample of the same workload.

Workloads
  viterbi : config 5 of BASELINE.json at rate 7/8 / QAM64 — the Viterbi stage of configs[1]
            (2k/QAM64/7-8), input in the viterbi_decoder block's own format.
  rx      : configs[1], full receive chain from baseband samples (when built).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


RANK = env_int("RANK", 0)
LOCAL_RANK = env_int("LOCAL_RANK", 0)
WORLD = env_int("WORLD_SIZE", 1)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# workload: viterbi (rate 7/8, QAM64)
# ---------------------------------------------------------------------------------------------
class ViterbiWorkload:
    name = "viterbi"
    RATE, M, CON = 4, 6, 2
    NBUF = 2  # distinct input buffers cycled per step: 2 x 84 MB > 126 MB L2

    def __init__(self, mbit_per_step):
        from oracle import port as O
        self.O = O
        k, n = O.RATE_KN[self.RATE]
        self.k, self.n = k, n
        nblocks = max(8, int(mbit_per_step * 1e6 / 8 / (96 * k)))
        self.nblocks = nblocks
        self.nbytes_out = nblocks * 96 * k
        self.nbytes_in = nblocks * 768 * n // self.M
        self.info_bits = (self.nbytes_out - 24) * 8
        # algorithmic bytes per info bit: n/(k*m) in + 1/8 out
        self.alg_bytes = self.nbytes_in + (self.nbytes_out - 24)

    def describe(self):
        return {"workload": "viterbi_decoder stage of configs[1] (2k/QAM64/rate-7/8): config 5 microbench, rate 7/8, m=6, "
                            "one stream of %d x 768-blocks per GPU per step" % self.nblocks,
                "input_bytes_per_step": self.nbytes_in, "l2_policy": "%d input buffers (%.0f MB total > 126 MB L2) cycled" % (self.NBUF, self.NBUF * self.nbytes_in / 1e6),
                "parallelism": "independent streams per GPU, no data-path collective"}

    def make_inputs(self, seed):
        O = self.O
        bufs = []
        for b in range(self.NBUF):
            data = np.random.default_rng(seed * 16 + b).integers(0, 256, self.nbytes_out, dtype=np.uint8)
            bufs.append((data, O.conv_encode(data, self.M, self.RATE)))
        return bufs

    # ---- GPU arm
    def setup_gpu(self, seed):
        import torch
        import gr_dvbt_b200 as g
        self.torch = torch
        self.g = g
        self.dec = g.viterbi_decoder(self.CON, g.NH, self.RATE)
        self.host = self.make_inputs(seed)
        self.d_in = [torch.from_numpy(rx).cuda() for _, rx in self.host]
        self.d_out = torch.zeros(self.nbytes_out, dtype=torch.uint8, device="cuda")
        self.pin_in = [torch.from_numpy(rx).pin_memory() for _, rx in self.host]
        self.pin_out = torch.zeros(self.nbytes_out, dtype=torch.uint8).pin_memory()
        self.kernel_ms = []

    def step_resident(self, i):
        b = i % self.NBUF
        n = self.dec.decode_dev(self.d_in[b].data_ptr(), self.nbytes_in, self.nbytes_in, 1, self.d_out.data_ptr(), self.nbytes_out)
        st = self.dec.last_stats()
        self.kernel_ms.append(st["acs_kernel_ms"])
        assert st["repaired"] == 0
        return n

    def step_e2e(self, i):
        b = i % self.NBUF
        import ctypes as C
        n_out = C.c_size_t(0)
        self.g.capi.check(self.g.capi.lib().dvbt_b200_viterbi_decode_host(self.dec._h, self.pin_in[b].data_ptr(), self.nbytes_in, self.nbytes_in, 1,
                                                                          self.pin_out.data_ptr(), self.nbytes_out, C.byref(n_out)))
        return int(n_out.value)

    def check(self):
        got = self.d_out[: self.nbytes_out - 24].cpu().numpy()
        b = (self.last_i) % self.NBUF
        return bool(np.array_equal(got, self.host[b][0][: len(got)]))

    def units_per_step(self):
        return self.info_bits / 1e6  # Mbit

    h2d = property(lambda self: self.nbytes_in)
    d2h = property(lambda self: self.nbytes_out - 24)

    # ---- CPU arm (one process; the reference keeps process-global decoder state)
    def cpu_sample(self, seconds_hint):
        """returns (Mbit decoded, seconds, kind) for ONE core"""
        O = self.O
        nblocks = 150
        data = np.random.default_rng(5).integers(0, 256, nblocks * 96 * self.k, dtype=np.uint8)
        rx = O.conv_encode(data, self.M, self.RATE)
        from oracle import refchain as R
        if R.available():
            t = time.time()
            out, _ = R.rx_viterbi(rx, self.CON, self.RATE, None, blocks_per_call=16)
            dt = time.time() - t
            kind = "reference"
        else:
            v = O.Viterbi(self.M, self.RATE)
            t = time.time()
            out = v.work(rx)
            dt = time.time() - t
            kind = "port"
        assert np.array_equal(out, data[: len(out)])
        return len(out) * 8 / 1e6, dt, kind


def cpu_worker(args):
    mbit, reps = args
    w = ViterbiWorkload(mbit)
    tot_bits, tot_t, kind = 0.0, 0.0, "port"
    for _ in range(reps):
        b, t, kind = w.cpu_sample(0)
        tot_bits += b
        tot_t += t
    return tot_bits, tot_t, kind


def run_cpu_all_cores(reps):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t = time.time()
    with ctx.Pool(cores) as pool:
        res = pool.map(cpu_worker, [(1.0, reps)] * cores)
    wall = time.time() - t
    bits = sum(r[0] for r in res)
    per_core = [r[0] / r[1] for r in res]
    return bits / max(r[1] for r in res), cores, res[0][2], float(np.mean(per_core)), wall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="viterbi")
    ap.add_argument("--mbit", type=float, default=640.0, help="decoded Mbit per GPU per step (viterbi workload)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    metric = "RX Msamples/s (baseband) & Viterbi Mbit/s @1/2/4/8 GPU vs SSE2 CPU; HBM GB/s %peak"

    if a.impl == "reference":
        if RANK != 0:
            return 0
        w = ViterbiWorkload(a.mbit)
        vals = []
        for i in range(a.warmup + a.steps):
            agg, cores, kind, per_core, wall = run_cpu_all_cores(1)
            if i >= a.warmup:
                vals.append((agg, wall))
        v = float(np.mean([x[0] for x in vals]))
        line = {"metric": metric, "value": v, "unit": "Mbit/s (Viterbi decoded bits, %d processes)" % cores, "impl": "reference", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean([x[1] for x in vals]) * 1e3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": w.describe(),
                "cpu_baseline": {"value": v, "unit": "Mbit/s", "cores": cores, "kind": kind,
                                 "sample": "150 x 768-blocks (%.2f Mbit) of the same rate-7/8 m=6 stream per process per step, one process per core" % (150 * 96 * 7 * 8 / 1e6)},
                "e2e": {"value": v, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(LOCAL_RANK)
    import gr_dvbt_b200 as g
    g.capi.check(g.capi.lib().dvbt_b200_set_device(LOCAL_RANK))
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL_RANK))
        # the only collective on this path: the configuration
        cfg = torch.tensor([a.mbit, a.steps, a.warmup], dtype=torch.float64, device="cuda")
        dist.broadcast(cfg, 0)
        a.mbit, a.steps, a.warmup = float(cfg[0]), int(cfg[1]), int(cfg[2])

    w = ViterbiWorkload(a.mbit)
    w.setup_gpu(seed=1 + RANK)
    lib = g.capi.lib()

    def barrier():
        if WORLD > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(stepfn, steps, warm):
        for i in range(warm):
            stepfn(i)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            stepfn(warm + i)
            w.last_i = warm + i
        e1.record()
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1)
        wall_ms = (time.perf_counter() - t0) * 1e3
        # every step synchronises its own stream inside the C ABI, so host wall time between the
        # barriers brackets the device work; take the larger of the two clocks
        ms = max(dev_ms, wall_ms)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if WORLD > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0])

    sampler = ClockSampler(LOCAL_RANK)
    l0 = lib.dvbt_b200_kernel_launches()
    w.kernel_ms = []
    sampler.start()
    ms = timed(w.step_resident, a.steps, a.warmup)
    clocks = sampler.stop()
    launches = (lib.dvbt_b200_kernel_launches() - l0) * a.steps // (a.steps + a.warmup)
    ok = w.check()
    kms = float(np.mean(w.kernel_ms[a.warmup:]))
    ms_e2e = timed(w.step_e2e, a.steps, a.warmup)

    if RANK == 0:
        peak, peak_src = load_peaks()
        units = w.units_per_step() * WORLD
        value = units / (ms / a.steps / 1e3)
        e2e = units / (ms_e2e / a.steps / 1e3)
        achieved = w.alg_bytes / (kms / 1e3) / 1e9
        cb_bits, cb_t, cb_kind = w.cpu_sample(0)
        line = {"metric": metric, "value": value, "unit": "Mbit/s (Viterbi decoded bits)", "n_gpus": WORLD, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic (seeded random TS bytes, K=7 encoded, punctured 7/8, error free)", "config": w.describe(),
                "parity_check": ok, "gpu_launches": int(launches), "clocks": clocks,
                "e2e": {"value": e2e, "unit": "Mbit/s", "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h, "ms_per_step": ms_e2e / a.steps,
                        "api": "dvbt_b200_viterbi_decode_host on pinned host buffers"},
                "roofline": {"kernel": "vit_acs_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": peak_src, "avg_launch_ms": kms,
                             "note": "ALU/shared-memory bound kernel (64 add-compare-select per decoded bit): HBM fraction is reported as the metric demands; "
                                     "ACS rate = %.1f T state-updates/s" % (w.info_bits * 64 / (kms / 1e3) / 1e12)},
                "cpu_baseline": {"value": cb_bits / cb_t, "unit": "Mbit/s", "cores": 1, "kind": cb_kind,
                                 "sample": "150 x 768-blocks (%.2f Mbit) of the same rate-7/8 m=6 stream, one thread" % cb_bits}}
        print(json.dumps(line))
    if WORLD > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
