"""CPU tests of the boundary: the library loads, exports every symbol include/dvbt_b200.h declares,
and fails loudly (no fallback) without a CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dvbt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dvbt_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import gr_dvbt_b200.capi as capi
    lib = capi.lib()
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_cites_the_reference_for_every_block():
    text = open(os.path.join(ROOT, "include", "dvbt_b200.h")).read()
    for f in ("viterbi_decoder_impl.cc", "ofdm_sym_acquisition_impl.cc", "demod_reference_signals_impl.cc", "dvbt_demap_impl.cc",
              "reed_solomon_dec_impl.cc"):
        assert f in text


def test_no_device_means_error_not_fallback():
    import gr_dvbt_b200 as g
    lib = g.capi.lib()
    if lib.dvbt_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(g.DvbtError) as e:
        g.viterbi_decoder(g.QAM16, g.NH, g.C1_2)
    assert e.value.code == -19 and "no CPU fallback" in str(e.value)
    with pytest.raises(g.DvbtError):
        g.rx_chain(g.QAM64, g.NH, g.C7_8, g.G1_32, g.T2k)


def test_argument_validation_precedes_device_use():
    import gr_dvbt_b200.capi as capi
    lib = capi.lib()
    h = C.c_void_p()
    bad = capi.ViterbiParams(7, 0, 0, 768, 0, -1)
    assert lib.dvbt_b200_viterbi_create(C.byref(bad), C.byref(h)) == -22
    assert b"constellation" in lib.dvbt_b200_last_error()
    bad_rs = capi.RsdecParams(2, 8, 0x11D, 255, 223, 16, 0, 8)
    assert lib.dvbt_b200_rsdec_create(C.byref(bad_rs), C.byref(h)) == -22


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gr_dvbt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")) and "shim_harness" not in f:
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "libdvbt_oracle" not in src, f
