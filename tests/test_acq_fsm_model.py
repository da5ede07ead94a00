"""CPU model check of the acquisition tables' detector (gr_dvbt_b200/csrc/acq.cu, acq_pass1_kernel / acq_pass2_kernel).

The reference's peak_detect_process (lib/ofdm_sym_acquisition_impl.cc:72-146) updates its running average the same way
whatever state it is in, so for a window of 16 lambda values and a start average the average sequence is a straight
float recurrence, and the state machine only sees it through the tests "v > avg * rise" and "v > avg * fall".  The CUDA
tables compute those tests first (two 16-bit masks) and replay the state machine on the masks.  This file restates
both formulations in float32 numpy - the reference loop line by line, the mask replay as the kernel does it - and
requires identical verdicts (number of peaks > 0, peak of peaks, final average bits) on adversarial windows: ties,
plateaus, several peaks, infinities, NaNs, zeros, sign changes.  It pins the algorithm, not the CUDA build (the GPU
suite compares the kernels' output with the oracle)."""
import numpy as np
import pytest

F = np.float32
RISE, FALL, ALPHA = F(0.8), F(0.9), F(0.9)   # ofdm_sym_acquisition_impl.cc:448
ONE_MINUS = F(1.0) - ALPHA


def reference_detect(d, avg):
    """peak_detect_process :72-146, float32 operation by operation"""
    avg = F(avg)
    state, peak_index, i = 0, 0, 0
    peak_val = F(-np.inf)
    peaks = []
    n = len(d)
    with np.errstate(all="ignore"):
        while i < n:
            if state == 0:
                if d[i] > avg * RISE:
                    state = 1
                else:
                    avg = ALPHA * d[i] + ONE_MINUS * avg
                    i += 1
            else:
                if d[i] > peak_val:
                    peak_val = d[i]
                    peak_index = i
                    avg = ALPHA * d[i] + ONE_MINUS * avg
                    i += 1
                elif d[i] > avg * FALL:
                    avg = ALPHA * d[i] + ONE_MINUS * avg
                    i += 1
                else:
                    peaks.append(peak_index)
                    state = 0
                    peak_val = F(-np.inf)
    best = -1
    if peaks:
        mx, best = d[peaks[0]], peaks[0]
        for k in peaks[1:]:
            if d[k] > mx:
                mx, best = d[k], k
    return len(peaks), best, F(avg)


def mask_replay_detect(d, avg):
    """acq_pass2_kernel: threshold masks from the average recurrence, then the state machine on the masks"""
    avg = F(avg)
    rise = fall = 0
    with np.errstate(all="ignore"):
        for i, v in enumerate(d):
            rise |= int(v > avg * RISE) << i
            fall |= int(v > avg * FALL) << i
            avg = ALPHA * v + ONE_MINUS * avg
    state, npeaks, peak_index, best = 0, 0, 0, 0
    peak_val, peak_at, best_val = F(-np.inf), F(0), F(0)
    for i, v in enumerate(d):
        while True:
            if state == 0:
                if (rise >> i) & 1:
                    state = 1
                    continue
                break
            if v > peak_val:
                peak_val, peak_at, peak_index = v, v, i
                break
            if (fall >> i) & 1:
                break
            if npeaks == 0 or peak_at > best_val:
                best_val, best = peak_at, peak_index
            npeaks += 1
            state = 0
            peak_val = F(-np.inf)
    return npeaks, (best if npeaks else -1), F(avg)


def windows(rng, count):
    out = []
    for k in range(count):
        kind = k % 8
        base = -np.abs(rng.normal(0.05, 0.05, 16))          # lambda is <= ~0 away from the peak
        if kind == 0:
            w = base
        elif kind == 1:                                      # one clean peak
            w = base.copy(); w[rng.integers(0, 16)] = -1e-4
        elif kind == 2:                                      # several peaks, some equal
            w = base.copy(); idx = rng.choice(16, 3, replace=False); w[idx] = rng.choice([-1e-4, -2e-4, -1e-4], 3)
        elif kind == 3:                                      # plateaus and exact ties
            w = np.repeat(rng.choice([-0.05, -0.01, -0.001, 0.0], 8), 2)
        elif kind == 4:                                      # sign changes, zeros
            w = rng.normal(0, 0.05, 16); w[rng.integers(0, 16)] = 0.0
        elif kind == 5:                                      # non-finite values
            w = base.copy(); w[rng.integers(0, 16)] = rng.choice([np.inf, -np.inf, np.nan])
        elif kind == 6:                                      # monotone ramps (no falling edge: the peak is never recorded)
            w = np.sort(base) if k % 16 < 8 else np.sort(base)[::-1]
        else:                                                # large dynamic range
            w = base * rng.choice([1e-20, 1.0, 1e20], 16)
        out.append(w.astype(np.float32))
    return out


@pytest.mark.parametrize("seed", [0, 1])
def test_mask_replay_equals_reference_detector(seed):
    rng = np.random.default_rng(seed)
    starts = [F(0), F(-0.05), F(-0.5), F(0.03), F(-1e-30), F(np.inf), F(np.nan)]
    n = 0
    for w in windows(rng, 1600):
        for a in (starts[n % len(starts)], F(rng.normal(-0.05, 0.05))):
            r = reference_detect(w, a)
            m = mask_replay_detect(w, a)
            assert (r[0] > 0) == (m[0] > 0) and r[0] == m[0], (w, a, r, m)
            assert r[1] == m[1], (w, a, r, m)
            assert r[2].tobytes() == m[2].tobytes() or (np.isnan(r[2]) and np.isnan(m[2])), (w, a, r, m)
            n += 1
    assert n == 3200
