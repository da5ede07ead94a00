"""Tuning sweep of the one-lane ACS kernel on the bench's rx workload: resident threads per SM and shared-memory
ring depth (DVBT_B200_VIT_TPSM / DVBT_B200_VIT_DEPTH are read at every launch).  Prints ms of the ACS kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench


def main():
    w = bench.RxWorkload(16)
    w.setup_gpu(seed=1)
    snr = float(os.environ.get("SWEEP_SNR", "0"))
    x = w.d_in
    if snr > 0:
        p_sig = float((x[1000:].abs() ** 2).mean())
        sigma = (p_sig / (10.0 ** (snr / 10.0)) / 2.0) ** 0.5
        x = torch.view_as_complex((torch.view_as_real(x) + torch.randn(x.shape[0], 2, device="cuda") * sigma).contiguous())
    ref = None
    configs = [(128, 24, 128, 0), (256, 0, 256, 0), (256, 0, 128, 0), (320, 0, 320, 0), (384, 0, 384, 0), (384, 0, 128, 0), (384, 0, 384, 4), (448, 0, 448, 0),
               (512, 0, 512, 0), (512, 0, 256, 0), (512, 0, 512, 4), (512, 3, 512, 0), (384, 4, 384, 0), (256, 6, 256, 0)]
    if os.environ.get("SWEEP_ONLY"):
        configs = [tuple(int(v) for v in c.split(",")) for c in os.environ["SWEEP_ONLY"].split(";")]
    for cfg in configs:
        tpsm, depth, bd = cfg[:3]
        os.environ["DVBT_B200_VIT_FMAADD"] = str(cfg[3] if len(cfg) > 3 else 0)
        os.environ["DVBT_B200_VIT_TPSM"] = str(tpsm)
        os.environ["DVBT_B200_VIT_DEPTH"] = str(depth)
        os.environ["DVBT_B200_VIT_BD"] = str(bd)
        ms = []
        for i in range(int(os.environ.get("SWEEP_REPS", "6"))):
            n = w.rx.run_file_dev(x.data_ptr(), w.nfile, w.GAIN, w.d_ts.data_ptr(), w.ts_cap)
            ms.append(w.rx.info()["ms_viterbi_acs"])
        ts = w.d_ts[:n].cpu().numpy()
        if ref is None:
            ref = ts.copy()
        print("tpsm %4d depth %2d bd %3d fma %s: acs %.3f ms (min %.3f)  same_ts=%s repaired=%d" % (tpsm, depth, bd, os.environ["DVBT_B200_VIT_FMAADD"], float(np.median(ms[2:])), min(ms), bool(np.array_equal(ts, ref)), w.rx.info()["viterbi_repaired"]), flush=True)


if __name__ == "__main__":
    main()
