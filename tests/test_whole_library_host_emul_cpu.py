"""The WHOLE library on the CPU: every .cu file of gr_dvbt_b200/csrc - kernels and host orchestration - compiled for the
host on the stand-in CUDA runtime of tests/emul/ (one host thread per CUDA thread, kernel launches rewritten from <<< >>>)
into tests/emul/_build/libdvbt_b200_emul.so, which exports the C ABI of include/dvbt_b200.h.  The GPU parity tests that
use host buffers are then run against it unchanged, through the same ctypes layer.

Test infrastructure only: the product never loads this library (gr_dvbt_b200.capi.lib() loads libdvbt_b200.so or raises);
the fixture below swaps the library object in for the duration of a test.  What it proves: the source is right wherever
the hardware is not involved.  What it cannot prove: anything about the GPU (semantics of PRMT / VIADDMNMX, CUDA's
sincosf / atan2f, FMA contraction in the float kernels, launch limits, timing) - that is what `pytest -m gpu` is for."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from gr_dvbt_b200 import capi
from oracle import refchain as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import build_vit_emul  # noqa: E402

# a kernel that is not warp-converged would dead-lock the lock-step emulation: never hang the suite (the host threads sit
# inside a C call, so only the thread method of pytest-timeout can end the run)
pytestmark = pytest.mark.timeout(900, method="thread")

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")


@pytest.fixture()
def emulated_library(monkeypatch):
    lib = capi.declare(C.CDLL(build_vit_emul.build_all()))
    monkeypatch.setattr(capi, "_lib", lib)
    monkeypatch.setenv("DVBT_B200_VIT_LANES", "1")   # warp primitives with partial member masks (two-lane ACS) are not emulated
    return lib


def test_fused_chain_on_the_reference_fixture(emulated_library):
    import test_golden_gpu as T
    T.test_chain_stage_by_stage_against_golden()


def test_blocks_on_the_reference_golden_vectors(emulated_library):
    import test_golden_gpu as T
    T.test_blocks_against_golden()


def test_smoke_chain_against_the_oracle(emulated_library):
    import smoke_chain
    assert smoke_chain.run() >= 1504


@needs_ref
def test_apps_test_ts_known_answer(emulated_library):
    """BASELINE configs[0]: the reference's own apps/test.ts comes back from packet 504 on"""
    import test_zz_apps_test_ts_gpu as T
    T.test_cuda_chain_reproduces_test_ts_from_packet_504()


@needs_ref
def test_chain_stage_by_stage_qam64_rate78(emulated_library):
    """BASELINE configs[1] mode (2k / QAM64 / 7/8) against the reference chain, stage by stage"""
    import test_rx_chain_gpu as T
    T.test_chain_matches_reference_stage_by_stage(R.QAM64, R.C7_8, R.T2k, 330, 1328)


@needs_ref
def test_chain_with_noise_and_both_rs_builds(emulated_library):
    import test_rx_chain_gpu as T
    T.test_chain_with_noise_uses_rs_and_matches_both_reference_builds()


@needs_ref
def test_baseband_chain_round_trip(emulated_library):
    """time-domain loopback: acquisition + fused FFT + the rest of the chain"""
    import test_acq_gpu as T
    T.test_baseband_chain_round_trip(R.QAM16, R.C1_2, R.T2k, 420, 504)


def test_viterbi_streaming_work_with_tags(emulated_library):
    import test_viterbi_gpu as T
    T.test_streaming_general_work_with_tags(4, 6)


@needs_ref
def test_demod_streaming_and_lock_loss(emulated_library):
    import test_demod_gpu as T
    T.test_streaming_calls_carry_state()


@pytest.mark.parametrize("schedule", ["h16b", "swar"])
def test_other_acs_schedules_through_the_whole_library(emulated_library, monkeypatch, schedule):
    """the opt-in h16b schedule (and the byte-SWAR baseline) through run_decode(): batch decode with many chunk
    boundaries, the repair path, the split survivor ring, streaming work() with tags, and the fused chain"""
    import test_golden_gpu as TG
    import test_viterbi_gpu as T
    monkeypatch.setenv("DVBT_B200_VIT_ACS", schedule)
    T.test_batch_matches_oracle(4, 6, 0.01)
    T.test_batch_matches_oracle(0, 4, 0.01)
    T.test_repair_path_is_exact(4, 6, 0.006)
    T.test_split_survivor_ring_is_exact(4, 6, 0.008, 2)
    T.test_streaming_general_work_with_tags(0, 4)
    TG.test_chain_stage_by_stage_against_golden()


@needs_ref
def test_config3_8k_qam16_rate12(emulated_library):
    import test_zz_apps_test_ts_gpu as T
    T.test_cuda_chain_config3_8k_qam16_rate12_stage_by_stage()


# ---- round 2: the stream interface, re-synchronisation, the transmit chain and the soft-decision mode ------------------------
@needs_ref
def test_stream_in_pieces_from_post_fft_symbols(emulated_library):
    """a capture in uneven pieces (some of one symbol) == the one-shot run: every block's call-to-call state is carried"""
    import test_stream_resync_gpu as T
    T.test_pieces_give_the_one_shot_transport_stream(*T.STREAM_CASES[2])


def test_stream_edge_cases(emulated_library):
    import test_stream_edge_cases_gpu as T
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    tx = tx_frequency_domain(con, cr, tm, 420, 17)
    cap = (con, cr, tm, tx, ofdm_modulate(tx["X"], tm, offset=333, cfo_bins=0.0, seed=2))
    T.test_empty_and_one_sample_pieces_change_nothing(cap)
    T.test_level_change_in_mid_stream_is_refused(cap)


@needs_ref
def test_transmit_chain_against_the_reference_tx_blocks(emulated_library):
    import test_tx_chain_gpu as T
    T.test_tx_stages_match_reference_blocks(*T.MODES[0])


def test_soft_decision_viterbi(emulated_library):
    """+-1 values == the hard decoder; arbitrary values == oracle/port's scalar restatement; the repair path on soft codes"""
    import test_soft_decision_gpu as T
    T.test_plus_minus_one_is_the_hard_decoder(4)
    T.test_soft_values_match_the_scalar_restatement(4, 0.38)
    T.test_soft_values_match_the_scalar_restatement(0, 0.7)
    T.test_soft_repair_path_is_exact()
    T.test_extreme_values_do_not_overflow_the_metrics()


def test_soft_decision_chain(emulated_library):
    """noise-free: the hard chain's TS; signs == the reference demapper; values == the documented rule; the chain's Viterbi
    output == the scalar soft decoder on the tapped values; pieces == one shot (rates 1/2, 3/4, 5/6, 7/8 between them)"""
    import test_soft_chain_gpu as T
    T.test_noise_free_soft_chain_gives_the_hard_chain_ts(R.QAM64, R.C7_8, R.T2k)
    T.test_soft_values_follow_the_documented_rule_and_the_decoder_is_exact(R.QAM16, R.C1_2)
    T.test_soft_values_follow_the_documented_rule_and_the_decoder_is_exact(R.QAM64, R.C3_4)
    T.test_soft_stream_in_pieces_equals_one_shot()
