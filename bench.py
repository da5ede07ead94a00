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


# ---- multi-GPU plumbing (no data-path collective; the configuration is the only broadcast) ----
def broadcast_config(values, device):
    """rank 0's list of numbers -> every rank (one torch.distributed broadcast)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, 0)
    return [float(v) for v in t.cpu()]


def max_over_ranks(ms, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def streams_of_rank(nstreams, world, rank):
    """stream s is decoded on GPU s mod G (config 4 of BASELINE.json)"""
    return [s for s in range(nstreams) if s % world == rank]


def seed_of_rank(rank, base=1):
    return env_int("BENCH_SEED", base) + rank


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        if os.environ.get("BENCH_SAMPLER", "1") == "0":   # diagnostic switch: no nvidia-smi process during the timed region
            return
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



# ---------------------------------------------------------------------------------------------
# workload: rx — configs[1] of BASELINE.json: 2k / QAM64 / rate 7/8 full receive chain from the
# 10 Msps capture (resampler 64/70, multiply_const, acquisition, FFT, demod, demap, inner
# deinterleavers, Viterbi, outer deinterleaver, RS, descrambler)
# ---------------------------------------------------------------------------------------------
class RxWorkload:
    name = "rx"
    CON, CR, TM = 2, 4, 0
    GAIN = 0.0022097087
    SUPERFRAMES_BASE = 4  # generated once with the reference TX blocks, then tiled
    # independent captures decoded concurrently per GPU (one handle = one CUDA stream + one host thread each): the
    # single-block control kernels of one capture overlap the wide kernels of the others.  Bounded by the host cores per rank.
    NCONC = max(1, env_int("BENCH_STREAMS", min(4, max(1, (os.cpu_count() or 1) // max(1, WORLD)))))

    def __init__(self, tiles):
        self.tiles = int(tiles)

    def describe(self):
        return {"workload": "configs[1]: 2k/QAM64/rate-7/8 RX, synthetic 10 Msps baseband capture -> TS (full flowgraph "
                            "apps/dvbt_rx_demo_2k_QAM64_rate78.grc: resampler 64/70, multiply_const, ofdm_sym_acquisition, FFT, "
                            "demod_reference_signals, dvbt_demap, symbol/bit deinterleavers, viterbi_decoder, convolutional_deinterleaver, "
                            "reed_solomon_dec, energy_descramble)",
                "captures_per_step": self.NCONC, "samples_per_capture": self.nfile,
                "samples_per_step": self.nfile * self.NCONC, "ofdm_symbols_per_step": self.nsym * self.NCONC,
                "input_bytes_per_step": self.nfile * 8 * self.NCONC,
                "l2_policy": "%d distinct resident captures of %.0f MB per step > 126 MB L2" % (self.NCONC, self.nfile * 8 / 1e6),
                "parallelism": "%d independent captures in flight per GPU (one stream each), independent captures per GPU, "
                               "no data-path collective" % self.NCONC}

    def build_capture(self, seed):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle import refchain as R
        from dvbt_testlib import tx_frequency_domain, ofdm_modulate, to_capture_rate
        if not R.available():
            raise RuntimeError("bench rx workload needs oracle/_ref (reference TX blocks) to synthesise the capture")
        nbase = 272 * self.SUPERFRAMES_BASE
        tx = tx_frequency_domain(self.CON, self.CR, self.TM, nbase, seed)
        X0 = tx["X"][:nbase]
        x0 = ofdm_modulate(X0, self.TM, gain=1.0)           # one block of whole superframes, 64/7 Msps
        x = np.tile(x0, self.tiles)
        self.nsym = nbase * self.tiles
        cap = to_capture_rate(np.concatenate([np.zeros(300, np.complex64), x]))
        self.nfile = len(cap)
        self.ts_src = tx["ts"]
        return cap

    def setup_gpu(self, seed):
        import torch
        import gr_dvbt_b200 as g
        self.torch, self.g = torch, g
        self.rx = g.rx_chain(self.CON, g.NH, self.CR, g.G1_32, self.TM)
        cap = self.build_capture(seed)
        self.d_in = torch.from_numpy(cap).cuda()
        self.pin_in = torch.from_numpy(cap).pin_memory()
        self.ts_cap = self.nsym * 1512
        self.d_ts = torch.zeros(self.ts_cap, dtype=torch.uint8, device="cuda")
        self.pin_ts = torch.zeros(self.ts_cap, dtype=torch.uint8).pin_memory()
        self.kernel_ms = []
        self.stage_ms = []
        self.ts_bytes = 0
        # e2e: two handles (two CUDA streams) so that the H2D copy of one capture overlaps the kernels of the other
        self.rx2 = [self.rx, g.rx_chain(self.CON, g.NH, self.CR, g.G1_32, self.TM)]
        self.pin_ts2 = [self.pin_ts, torch.zeros(self.ts_cap, dtype=torch.uint8).pin_memory()]

    def e2e_pipelined(self, steps):
        """`steps` host-buffer calls spread over two worker threads (one handle each); returns wall ms"""
        import ctypes as C
        lib = self.g.capi.lib()

        def worker(k, count):
            n = C.c_size_t(0)
            for _ in range(count):
                self.g.capi.check(lib.dvbt_b200_rx_run_file_host(self.rx2[k]._h, self.pin_in.data_ptr(), self.nfile, self.GAIN,
                                                                 self.pin_ts2[k].data_ptr(), self.ts_cap, C.byref(n)))
        counts = [steps - steps // 2, steps // 2]
        th = [threading.Thread(target=worker, args=(k, counts[k])) for k in range(2)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return (time.perf_counter() - t0) * 1e3

    def resident_pair(self, steps):
        """`steps` batches of TWO independent captures, one per handle (= per CUDA stream), each driven by its own host
        thread: the single-block control kernels of one capture (acq_compose, demod_scan, acq_finish ...) overlap the wide
        kernels of the other.  Returns wall ms between start and join (every call synchronises its stream)."""
        ns = self.NCONC
        if not hasattr(self, "d_in2"):
            g = self.g
            while len(self.rx2) < ns:
                self.rx2.append(g.rx_chain(self.CON, g.NH, self.CR, g.G1_32, self.TM))
            self.d_in2 = [self.d_in] + [self.d_in.clone() for _ in range(ns - 1)]     # resident captures: ns x 402 MB > L2
            self.d_ts2 = [self.d_ts] + [self.torch.zeros_like(self.d_ts) for _ in range(ns - 1)]
            self.torch.cuda.synchronize()
        self.pair_bytes = [0] * ns

        def worker(k):
            for _ in range(steps):
                self.pair_bytes[k] = self.rx2[k].run_file_dev(self.d_in2[k].data_ptr(), self.nfile, self.GAIN, self.d_ts2[k].data_ptr(), self.ts_cap)
        th = [threading.Thread(target=worker, args=(k,)) for k in range(ns)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return (time.perf_counter() - t0) * 1e3

    def in_flight_sweep(self, legs, steps):
        """informational: the resident leg with other numbers of captures in flight (same kernels, more handles) and with
        launch-geometry knobs of the library that are read at every launch; legs = [(label, captures in flight, env)]"""
        res = {}
        keep = self.NCONC
        try:
            for label, ns, env in legs:
                for k in ("DVBT_B200_VIT_SM_DIV",):
                    os.environ.pop(k, None)
                os.environ.update(env)
                g = self.g
                while len(self.rx2) < ns:
                    self.rx2.append(g.rx_chain(self.CON, g.NH, self.CR, g.G1_32, self.TM))
                while len(self.d_in2) < ns:
                    self.d_in2.append(self.d_in.clone())
                    self.d_ts2.append(self.torch.zeros_like(self.d_ts))
                self.torch.cuda.synchronize()
                self.NCONC = ns
                self.resident_pair(2)
                ms = self.resident_pair(steps)
                same = all(b == self.ts_bytes for b in self.pair_bytes) and all(self.torch.equal(self.d_ts2[0][: self.ts_bytes], t[: self.ts_bytes]) for t in self.d_ts2[1:ns])
                res[label] = {"captures_in_flight": ns, "env": env, "ms_per_capture": ms / steps / ns, "value": self.nfile / 1e6 * ns / (ms / steps / 1e3), "outputs_identical": bool(same)}
        except Exception as e:   # informational leg: never fatal
            res["error"] = repr(e)[:200]
        finally:
            self.NCONC = keep
            os.environ.pop("DVBT_B200_VIT_SM_DIV", None)
        return res

    def step_resident(self, i):
        n = self.rx.run_file_dev(self.d_in.data_ptr(), self.nfile, self.GAIN, self.d_ts.data_ptr(), self.ts_cap)
        inf = self.rx.info()
        self.kernel_ms.append(inf["ms_viterbi_acs"])
        self.stage_ms.append({k: v for k, v in inf.items() if k.startswith("ms_")})
        self.ts_bytes = n
        self.info = inf
        return n

    def step_e2e(self, i):
        import ctypes as C
        n = C.c_size_t(0)
        self.g.capi.check(self.g.capi.lib().dvbt_b200_rx_run_file_host(self.rx._h, self.pin_in.data_ptr(), self.nfile, self.GAIN,
                                                                      self.pin_ts.data_ptr(), self.ts_cap, C.byref(n)))
        return int(n.value)

    def noisy_leg(self, snr_db, steps, seed):
        """same capture + complex AWGN at `snr_db` (signal power over the non-silent part, noise over the full 10 MHz band):
        device-resident throughput with the Viterbi traceback and RS correction doing real work"""
        torch = self.torch
        x = self.d_in
        p_sig = float((x[1000:].abs() ** 2).mean())
        gen = torch.Generator(device="cuda").manual_seed(seed)
        sigma = (p_sig / (10.0 ** (snr_db / 10.0)) / 2.0) ** 0.5
        noise = torch.randn(x.shape[0], 2, device="cuda", generator=gen, dtype=torch.float32) * sigma
        d_noisy = (torch.view_as_real(x) + noise).contiguous()
        del noise
        for _ in range(2):
            self.rx.run_file_dev(d_noisy.data_ptr(), self.nfile, self.GAIN, self.d_ts.data_ptr(), self.ts_cap)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            n = self.rx.run_file_dev(d_noisy.data_ptr(), self.nfile, self.GAIN, self.d_ts.data_ptr(), self.ts_cap)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        inf = self.rx.info()
        ts = self.d_ts[:n].cpu().numpy().reshape(-1, 188)
        # the capture is one 4-superframe block tiled: only the first tile maps 1:1 onto the source TS (from packet 1328 on)
        m = min(len(ts), 3900)
        src = self.ts_src[: len(self.ts_src) // 188 * 188].reshape(-1, 188)
        good = int((ts[:m] == src[1328:1328 + m]).all(axis=1).sum()) if m and inf["acq_lost_at"] == -1 else 0
        return {"snr_db": snr_db, "ms_per_step": ms, "value": self.nfile / 1e6 / (ms / 1e3), "ms_viterbi_acs": inf["ms_viterbi_acs"],
                "viterbi_repaired_chunks": inf["viterbi_repaired"], "ts_packets": int(len(ts)), "first_tile_packets_checked": int(m),
                "first_tile_packets_equal_to_source": good}

    def check(self):
        ts = self.d_ts[: min(self.ts_bytes, 188 * 3000)].cpu().numpy()
        ref = self.ts_src[1328 * 188: 1328 * 188 + len(ts)]
        n = min(len(ts), len(ref))
        return bool(n > 188 * 100 and np.array_equal(ts[:n], ref[:n]) and self.info["acq_lost_at"] == -1)

    def units_per_step(self):
        return self.nfile / 1e6  # Msamples of the 10 Msps capture

    h2d = property(lambda self: self.nfile * 8)
    d2h = property(lambda self: int(self.ts_bytes))

    @property
    def viterbi_bits(self):
        return self.info["viterbi_bytes"] * 8

    @property
    def alg_bytes(self):
        # Viterbi stage in the reference I/O format: n/(k*m) B in + 1/8 B out per decoded bit
        return self.viterbi_bits * (8.0 / (7 * 6) + 0.125)


def ncu_traffic(workload, tiles):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json), or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(workload)
    except (OSError, ValueError):
        return None, None
    if not t or (workload == "rx" and t.get("tiles") != tiles):
        return None, None
    return float(t["dram_bytes_read"] + t["dram_bytes_write"]), {k: t[k] for k in ("source", "alu_pipe_active_pct", "issue_active_pct", "dram_throughput_pct") if k in t}


def cpu_worker(args):
    mbit, reps = args
    w = ViterbiWorkload(mbit)
    tot_bits, tot_t, kind = 0.0, 0.0, "port"
    for _ in range(reps):
        b, t, kind = w.cpu_sample(0)
        tot_bits += b
        tot_t += t
    return tot_bits, tot_t, kind


# ---- reference RX chain on the CPU (the reference's own blocks via oracle/_ref; numpy/scipy stand in for the
# stock GNU Radio resampler and FFT, which are not part of gr-dvbt)
_CPU_RX = {}


def cpu_rx_prepare(seed=7, nsym=1904):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import refchain as R
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate, to_capture_rate
    tx = tx_frequency_domain(RxWorkload.CON, RxWorkload.CR, RxWorkload.TM, nsym, seed)
    x = ofdm_modulate(tx["X"][:nsym], RxWorkload.TM, gain=1.0, offset=300)
    _CPU_RX["cap"] = to_capture_rate(x)
    _CPU_RX["ts"] = tx["ts"]


def cpu_rx_chain(_=None):
    """one pass of the reference flowgraph over the prepared capture; returns (Msamples, seconds, per-stage seconds)"""
    from scipy.signal import resample_poly
    from oracle import refchain as R
    cap = _CPU_RX["cap"]
    con, cr, tm = RxWorkload.CON, RxWorkload.CR, RxWorkload.TM
    N, P, K, cp = R.mode_dims(tm)
    st = {}
    t0 = time.time()
    x = (resample_poly(cap, 32, 35, window=("kaiser", 7.0)) * np.float32(RxWorkload.GAIN)).astype(np.complex64)
    st["resample(scipy)"] = time.time() - t0; t = time.time()
    sym, cons, _tags = R.rx_acquisition(x, tm)
    st["ofdm_sym_acquisition"] = time.time() - t; t = time.time()
    Xf = np.fft.fftshift(np.fft.fft(sym, axis=1), axes=1).astype(np.complex64)
    st["fft(numpy)"] = time.time() - t; t = time.time()
    Y, tags = R.rx_demod(Xf, con, cr, tm)
    st["demod_reference_signals"] = time.time() - t; t = time.time()
    dm = R.rx_demap(Y, con, tm)
    st["dvbt_demap"] = time.time() - t; t = time.time()
    sd, bd = R.rx_deinterleave(dm, tags, con, tm)
    st["inner_deinterleavers"] = time.time() - t; t = time.time()
    sf = [tg for tg in tags if tg[1] == "superframe_start"][0][0]
    vo, vtags = R.rx_viterbi(bd, con, cr, sf * P)
    st["viterbi_decoder"] = time.time() - t; t = time.time()
    cd, rd, ts = R.rx_outer(vo, vtags)
    st["outer(deint+rs+descramble)"] = time.time() - t
    dt = time.time() - t0
    ok = len(ts) > 0 and np.array_equal(ts, _CPU_RX["ts"][1328 * 188: 1328 * 188 + len(ts)])
    return len(cap) / 1e6, dt, st, ok, len(vo) * 8 / 1e6


def run_cpu_all_cores(reps, workload):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t = time.time()
    if workload == "rx":
        with ctx.Pool(cores) as pool:
            res = pool.map(cpu_rx_chain, range(cores))
        wall = time.time() - t
        units = sum(r[0] for r in res)
        assert all(r[3] for r in res), "reference chain did not reproduce the transmitted TS"
        return units / max(r[1] for r in res), cores, "reference", float(np.mean([r[0] / r[1] for r in res])), wall
    with ctx.Pool(cores) as pool:
        res = pool.map(cpu_worker, [(1.0, reps)] * cores)
    wall = time.time() - t
    bits = sum(r[0] for r in res)
    per_core = [r[0] / r[1] for r in res]
    return bits / max(r[1] for r in res), cores, res[0][2], float(np.mean(per_core)), wall


def acs_variants_child(mbit, out):
    """Child process of the A/B leg below: the Viterbi stage of configs[1] (rate 7/8, QAM64, error free and with bit
    errors) decoded by the default ACS schedule and by the opt-in ones, ACS kernel time from the library's CUDA events.
    Runs in its own process so that a fault in an opt-in kernel cannot touch the measured run."""
    import torch
    import gr_dvbt_b200 as g
    from oracle import port as O
    g.capi.check(g.capi.lib().dvbt_b200_set_device(0))
    w = ViterbiWorkload(mbit)
    host = w.make_inputs(12345)
    data, rx = host[0]
    noisy = O.flip_bits(rx[: 400 * 768 * w.n // w.M], w.M, 0.004, 9)
    noisy_ref = O.Viterbi(w.M, w.RATE).work(noisy[: 40 * 768 * w.n // w.M])
    d_in = [torch.from_numpy(r).cuda() for _, r in host]
    d_out = torch.zeros(w.nbytes_out, dtype=torch.uint8, device="cuda")
    res = {"workload": "rate 7/8, m=6, %d x 768-blocks (%.1f Mbit) per launch, %d input buffers cycled" % (w.nblocks, w.info_bits / 1e6, len(d_in))}
    noisy_out = {}
    knobs = ("DVBT_B200_VIT_ACS", "DVBT_B200_VIT_TPSM", "DVBT_B200_VIT_BD")
    variants = [("h16", {"DVBT_B200_VIT_ACS": "h16"}), ("h16b", {"DVBT_B200_VIT_ACS": "h16b"}),
                ("h16b_512", {"DVBT_B200_VIT_ACS": "h16b", "DVBT_B200_VIT_TPSM": "512", "DVBT_B200_VIT_BD": "512"})]
    for variant, env in variants:
        for k in knobs:
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            dec = g.viterbi_decoder(w.CON, g.NH, w.RATE)
            ms = []
            for i in range(3 + 10):
                dec.decode_dev(d_in[i % len(d_in)].data_ptr(), w.nbytes_in, w.nbytes_in, 1, d_out.data_ptr(), w.nbytes_out)
                st = dec.last_stats()
                if i >= 3:
                    ms.append(st["acs_kernel_ms"])
            got = d_out[: w.nbytes_out - 24].cpu().numpy()
            clean_ok = bool(np.array_equal(got, host[(3 + 10 - 1) % len(d_in)][0][: len(got)]))
            nz = dec.decode(noisy)[0]
            noisy_out[variant] = nz
            res[variant] = {"acs_kernel_ms": float(np.mean(ms)), "mbit_per_s_kernel": w.info_bits / 1e6 / (float(np.mean(ms)) / 1e3),
                            "error_free_input_decodes_to_source": clean_ok, "repaired_chunks": st["repaired"],
                            "noisy_input_equals_oracle_prefix": bool(np.array_equal(nz[: len(noisy_ref)], noisy_ref))}
        except Exception as e:   # an opt-in variant must never cost the bench its line
            res[variant] = {"error": repr(e)[:300]}
    if all(v in noisy_out for v in ("h16", "h16b")):
        res["noisy_input_same_bytes_both_schedules"] = bool(np.array_equal(noisy_out["h16"], noisy_out["h16b"]))
    print(json.dumps(res), file=out, flush=True)
    return 0


def acs_variants_leg(mbit):
    """A/B of the ACS schedules (default h16 vs opt-in h16b) in a child process; informational, never fatal."""
    import subprocess
    try:
        env = dict(os.environ)
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "DVBT_B200_VIT_ACS"):
            env.pop(k, None)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--acs-ab-child", "--mbit", "%g" % mbit],
                           capture_output=True, text=True, timeout=240, env=env)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if lines:
            return json.loads(lines[-1])
        return {"error": "child printed nothing", "returncode": r.returncode, "stderr_tail": r.stderr[-400:]}
    except Exception as e:
        return {"error": repr(e)[:300]}


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries loaded later print there too (NCCL's version banner, for
    one), so file descriptor 1 is pointed at stderr for the rest of the run and the JSON line goes to a private copy of
    the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="rx", choices=["rx", "viterbi"])
    ap.add_argument("--mbit", type=float, default=640.0, help="decoded Mbit per GPU per step (viterbi workload)")
    ap.add_argument("--tiles", type=int, default=20, help="rx workload: capture = tiles x 4 superframes (1088 OFDM symbols each); 20 tiles = 50.3 M samples (SURVEY §8d config 2: >= 50 M)")
    ap.add_argument("--acs-ab-child", action="store_true", help=argparse.SUPPRESS)
    a = ap.parse_args()
    if a.acs_ab_child:
        return acs_variants_child(a.mbit, out)
    a.warmup = max(a.warmup, 3)
    metric = "RX Msamples/s (baseband) & Viterbi Mbit/s @1/2/4/8 GPU vs SSE2 CPU; HBM GB/s %peak"
    rx = a.workload == "rx"
    unit = "Msamples/s (10 Msps-domain complex64 baseband, whole RX chain to TS)" if rx else "Mbit/s (Viterbi decoded bits)"

    if a.impl == "reference":
        if RANK != 0:
            return 0
        if rx:
            from oracle import refchain as R
            if not R.available():
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference sources compiled verbatim) was not built on this box"}), file=out, flush=True)
                return 0
            cpu_rx_prepare()
            w = RxWorkload(a.tiles)
            w.nsym, w.nfile = 1904, len(_CPU_RX["cap"])
            sample = "one capture of 1904 OFDM symbols (%.2f Msamples at 10 Msps) per process per step, one process per core (the reference keeps process-global Viterbi state)" % (len(_CPU_RX["cap"]) / 1e6)
        else:
            w = ViterbiWorkload(a.mbit)
            sample = "150 x 768-blocks (%.2f Mbit) of the same rate-7/8 m=6 stream per process per step, one process per core" % (150 * 96 * 7 * 8 / 1e6)
        vals = []
        for i in range(a.warmup + a.steps):
            agg, cores, kind, per_core, wall = run_cpu_all_cores(1, a.workload)
            if i >= a.warmup:
                vals.append((agg, wall))
        v = float(np.mean([x[0] for x in vals]))
        line = {"metric": metric, "value": v, "unit": unit, "impl": "reference", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean([x[1] for x in vals]) * 1e3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic", "config": w.describe(),
                "cpu_baseline": {"value": v, "unit": unit.split(" (")[0], "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": v, "unit": unit.split(" (")[0], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=out, flush=True)
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(LOCAL_RANK)
    import gr_dvbt_b200 as g
    g.capi.check(g.capi.lib().dvbt_b200_set_device(LOCAL_RANK))
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL_RANK))
        # the only collective on this path: the configuration
        cfg = broadcast_config([a.mbit, a.steps, a.warmup, a.tiles], "cuda")
        a.mbit, a.steps, a.warmup, a.tiles = float(cfg[0]), int(cfg[1]), int(cfg[2]), int(cfg[3])

    w = RxWorkload(a.tiles) if rx else ViterbiWorkload(a.mbit)
    w.setup_gpu(seed=seed_of_rank(RANK))
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
        per_step = []
        for i in range(steps):
            ts = time.perf_counter()
            stepfn(warm + i)
            per_step.append((time.perf_counter() - ts) * 1e3)
            w.last_i = warm + i
        e1.record()
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1)
        wall_ms = (time.perf_counter() - t0) * 1e3
        # every step synchronises its own stream inside the C ABI, so host wall time between the
        # barriers brackets the device work; take the larger of the two clocks
        ms = max(dev_ms, wall_ms)
        if os.environ.get("BENCH_VERBOSE"):
            sys.stderr.write("[bench rank %d] per-step wall ms: %s\n" % (RANK, " ".join("%.2f" % v for v in per_step)))
        ms = max_over_ranks(ms, "cuda")
        barrier()
        return ms

    sampler = ClockSampler(LOCAL_RANK)
    l0 = lib.dvbt_b200_kernel_launches()
    w.kernel_ms = []
    sampler.start()
    ms = timed(w.step_resident, a.steps, a.warmup)
    clocks = sampler.stop()
    if rx:
        if os.environ.get("BENCH_VERBOSE"):
            keys = [k for k in w.stage_ms[-1]]
            sys.stderr.write("[bench rank %d] stages: %s | info: %s\n" % (RANK, " ".join("%s=%.3f" % (k[3:], float(np.mean([s_[k] for s_ in w.stage_ms[a.warmup:]]))) for k in keys),
                                                                        json.dumps({k: v for k, v in w.info.items() if not k.startswith("ms_")})))
        sys.stderr.write("[bench rank %d] resident: %.3f ms/step (max over ranks), sum of this rank's stage times %.3f ms\n"
                         % (RANK, ms / a.steps, float(np.mean([sum(v for k, v in s_.items() if k not in ("ms_viterbi_acs", "ms_fft", "ms_equalise")) for s_ in w.stage_ms[a.warmup:]]))))
    launches = (lib.dvbt_b200_kernel_launches() - l0) // (a.steps + a.warmup)   # kernels of this library per step
    ok = w.check()
    kms = float(np.mean(w.kernel_ms[a.warmup:]))
    ms_pair = None
    if rx:
        # the headline step: NCONC captures in flight per GPU
        w.resident_pair(a.warmup)
        barrier()
        l1 = lib.dvbt_b200_kernel_launches()
        sampler.start()
        ms_pair = max_over_ranks(w.resident_pair(a.steps), "cuda")
        clocks_pair = sampler.stop()
        if clocks_pair.get("samples"):
            clocks = clocks_pair
        launches = (lib.dvbt_b200_kernel_launches() - l1) // a.steps
        barrier()
        same = bool(all(b == w.ts_bytes for b in w.pair_bytes) and all(w.torch.equal(w.d_ts2[0][: w.ts_bytes], t[: w.ts_bytes]) for t in w.d_ts2[1:]))
        if os.environ.get("BENCH_VERBOSE"):
            sys.stderr.write("[bench rank %d] %d concurrent captures: %.3f ms per batch, outputs identical: %s\n" % (RANK, w.NCONC, ms_pair / a.steps, same))
        ok = ok and same
    if os.environ.get("BENCH_QUICK"):   # tuning runs: resident legs only, no JSON line
        if RANK == 0:
            sys.stderr.write("[bench quick] one capture %.3f ms, %d concurrent %.3f ms per capture, acs %.3f ms, parity %s\n"
                             % (ms / a.steps, w.NCONC if rx else 1, (ms_pair / a.steps / w.NCONC) if rx else 0.0, kms, ok))
        if WORLD > 1:
            dist.destroy_process_group()
        return 0
    noisy = w.noisy_leg(27.0, max(3, a.steps // 4), 4242 + RANK) if rx and RANK == 0 else None
    ms_e2e = timed(w.step_e2e, a.steps, a.warmup)
    ms_e2e_pipe = None
    if rx:
        w.e2e_pipelined(2)
        barrier()
        ms_e2e_pipe = max_over_ranks(w.e2e_pipelined(a.steps), "cuda")
        barrier()

    if RANK == 0:
        peak, peak_src = load_peaks()
        units = w.units_per_step() * WORLD
        value = units / (ms / a.steps / 1e3)
        e2e = units / (ms_e2e / a.steps / 1e3)
        single = {"ms_per_capture": ms / a.steps, "value": value}
        if rx:
            ms = ms_pair
            value = units * w.NCONC / (ms / a.steps / 1e3)
        achieved = w.alg_bytes / (kms / 1e3) / 1e9
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": WORLD, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32" if rx else "u8",
                "config": w.describe(), "parity_check": ok, "gpu_launches": int(launches), "clocks": clocks}
        if rx:
            line["chain_info"] = {k: v for k, v in w.info.items() if not k.startswith("ms_")}
        if rx:
            vbits = w.viterbi_bits
            stage = {k: float(np.mean([s[k] for s in w.stage_ms[a.warmup:]])) for k in w.stage_ms[-1]}
            line["data"] = "synthetic (seeded random TS -> reference TX blocks -> IFFT/CP -> 35/32 resampler -> 10 Msps capture, no added noise)"
            line["viterbi_mbit_per_s"] = vbits * w.NCONC * WORLD / (ms / a.steps / 1e3) / 1e6
            line["realtime_factor"] = value / WORLD / 10.0
            line["stage_ms"] = stage
            single["note"] = ("one capture at a time on one handle (one stream): the stage_ms / roofline kernel times below are measured "
                              "in this leg with CUDA events; `value` is the same chain with %d captures in flight" % w.NCONC)
            line["one_capture_at_a_time"] = single
            line["awgn"] = noisy   # informational: not the metric's configuration (BASELINE configs[1] is noise free)
            if WORLD == 1 and not os.environ.get("BENCH_NO_SWEEP") and (os.cpu_count() or 1) >= 8:
                # informational: more captures in flight than the headline's NCONC (single-GPU value per count)
                nc = w.NCONC
                line["in_flight_sweep"] = w.in_flight_sweep(
                    [("%d_again" % nc, nc, {}), ("%d_acs_on_half_the_sms" % nc, nc, {"DVBT_B200_VIT_SM_DIV": "2"}),
                     ("%d_acs_on_a_quarter_of_the_sms" % nc, nc, {"DVBT_B200_VIT_SM_DIV": "4"}), ("6", 6, {}), ("8", 8, {}),
                     ("8_acs_on_a_quarter_of_the_sms", 8, {"DVBT_B200_VIT_SM_DIV": "4"})], max(3, a.steps // 4))
            if WORLD == 1 and not os.environ.get("BENCH_NO_ACS_AB"):
                # informational: the opt-in ACS schedule beside the default one on the Viterbi stage alone (child process)
                line["acs_variants"] = acs_variants_leg(vbits / 1e6)
            e2e_pipe = units / (ms_e2e_pipe / a.steps / 1e3)
            line["e2e"] = {"value": e2e_pipe, "unit": "Msamples/s", "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h,
                           "ms_per_step": ms_e2e_pipe / a.steps,
                           "api": "dvbt_b200_rx_run_file_host on pinned host buffers, two handles driven by two host threads "
                                  "(copy of one capture overlaps the kernels of the other); every step copies its capture H2D and its TS D2H",
                           "single_handle": {"value": e2e, "ms_per_step": ms_e2e / a.steps}}
            info_bits = vbits
            cpu_rx_prepare()
            cu, ct, cst, cok, cvit = cpu_rx_chain()
            line["cpu_baseline"] = {"value": cu / ct, "unit": "Msamples/s", "cores": 1, "kind": "reference",
                                    "sample": "one capture of 1904 OFDM symbols (%.2f Msamples at 10 Msps), one thread; the reference's own blocks "
                                              "(oracle/_ref) with scipy/numpy standing in for the stock GNU Radio resampler and FFT" % cu,
                                    "viterbi_mbit_per_s": cvit / cst["viterbi_decoder"], "ts_ok": cok,
                                    "stage_share": {k: round(v / ct, 3) for k, v in cst.items()}}
        else:
            line["data"] = "synthetic (seeded random TS bytes, K=7 encoded, punctured 7/8, error free)"
            line["e2e"] = {"value": e2e, "unit": "Mbit/s", "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h, "ms_per_step": ms_e2e / a.steps,
                           "api": "dvbt_b200_viterbi_decode_host on pinned host buffers"}
            info_bits = w.info_bits
            cb_bits, cb_t, cb_kind = w.cpu_sample(0)
            line["cpu_baseline"] = {"value": cb_bits / cb_t, "unit": "Mbit/s", "cores": 1, "kind": cb_kind,
                                    "sample": "150 x 768-blocks (%.2f Mbit) of the same rate-7/8 m=6 stream, one thread" % cb_bits}
        traffic, traffic_src = ncu_traffic(a.workload, a.tiles)
        line["roofline"] = {"kernel": "vit_acs_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": traffic, "ncu": traffic_src, "algorithmic_bytes": w.alg_bytes,
                            "peak_source": peak_src, "avg_launch_ms": kms,
                            "note": "dominant kernel of the step; bound by the integer ALU pipe (64 add-compare-select per decoded bit as halfword "
                                    "VIADDMNMX.U16x2 + IMAD; ncu: see the `ncu` object), so the HBM fraction is small by nature; DRAM traffic above the "
                                    "algorithmic bytes is the survivor-row write-through to the global ring (deliberate: it frees shared memory "
                                    "for 3x the resident warps); ACS rate = %.1f T state-updates/s" % (info_bits * 64 / (kms / 1e3) / 1e12)}
        if rx:
            # the HBM-bound kernels of the step against the same measured peak (algorithmic bytes of SURVEY §8d, CUDA-event
            # times of the one-capture-at-a-time leg); the ACS kernel above is the dominant one but bound by the ALU pipe
            inf = w.info
            nout = (w.nfile - 1) * 32 // 35 + 1
            others = [("resample_multi_kernel<2> (rational_resampler 64/70 + multiply_const)", "ms_resample", 8.0 * w.nfile + 8.0 * nout),
                      ("acq_fftd_kernel<2048> (derotation + CP removal + forward FFT)", "ms_fft", inf["acq_symbols"] * (8.0 * (2048 + 64) + 8.0 * 2048)),
                      ("demod_equalise_kernel (channel estimate + equalise + demap)", "ms_equalise", inf["symbols_parsed"] * (8.0 * 2048 + 1512.0))]
            line["roofline_other"] = [{"kernel": nm, "bound": "hbm", "achieved": by / (stage[key] / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": by / (stage[key] / 1e3) / 1e9 / peak, "algorithmic_bytes": by, "avg_launch_ms": stage[key]}
                                      for nm, key, by in others if stage.get(key, 0) > 0]
        print(json.dumps(line), file=out, flush=True)
    if WORLD > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
