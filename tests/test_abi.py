"""The C-ABI library loads on a CPU-only box and exports every symbol include/ake_b200.h declares;
plan construction (pure host code) reproduces the reference's state_dict layout and error behaviour."""
import ctypes as C

import pytest

import audio_key_estimation_b200 as ake
from audio_key_estimation_b200 import _lib
from audio_key_estimation_b200._lib import PcnConfig
from conftest import golden_state_dict


def _cfg(**kw):
    d = dict(pitches=288, pitch_classes=12, num_layers=2, kernel_size=7, conv_layers=3, n_filters=4, head_layers=2,
             time_pool_size=2, genre=0, max_pool=0, frames=5, loc_window_size=10)
    d.update(kw)
    return PcnConfig(**d)


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = _lib.header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_lib._SIGNATURES)
    assert lib.ake_abi_version() == 2


@pytest.mark.parametrize("genre", [False, True])
def test_plan_tensor_table_matches_reference_state_dict(genre):
    lib = _lib.lib()
    h = C.c_void_p()
    cfg = _cfg(genre=int(genre))
    assert lib.ake_pcn_create(C.byref(cfg), C.byref(h)) == 0
    try:
        sd = {k: v for k, v in golden_state_dict(genre).items() if not k.endswith("num_batches_tracked")}
        n = lib.ake_pcn_num_tensors(h)
        names = [lib.ake_pcn_tensor_name(h, i).decode() for i in range(n)]
        assert names == list(sd)
        shape4 = (C.c_int64 * 4)()
        for i, k in enumerate(names):
            nd = lib.ake_pcn_tensor_shape(h, i, C.byref(shape4))
            assert tuple(shape4[:nd]) == tuple(sd[k].shape), k
        assert lib.ake_pcn_param_floats(h) == sum(v.numel() for v in sd.values())
        n_learn = sum(v.numel() for k, v in sd.items() if "running_" not in k)
        assert n_learn == (167031 if genre else 162902)
        out = PcnConfig()
        assert lib.ake_pcn_get_config(h, C.byref(out)) == 0 and out.pitches == 288 and out.genre == int(genre)
        assert lib.ake_pcn_workspace_bytes(h, 4, 151, 0) > 0
        assert lib.ake_pcn_workspace_bytes(h, 4, 151, 1) > 0  # train mode keeps pre-BN buffers, eval mode the tensor-core staging planes
    finally:
        lib.ake_pcn_destroy(h)


def test_non_default_architecture_tensor_tables_match_the_reference():
    """with every built architecture switch the plan's tensor table == the state_dict of the unmodified reference
    (names, order, shapes; tests/golden/variants.npz from oracle/make_golden_variants.py), so strict loading works."""
    import json
    from conftest import load_golden
    meta = json.loads(bytes(load_golden("variants.npz")["meta"]).decode())
    lib = _lib.lib()
    for group in ("variants", "tables"):
        for name, entry in meta[group].items():
            opt = dict(entry["opt"])
            cfg = _cfg(**{k: int(v) for k, v in opt.items()})
            h = C.c_void_p()
            assert lib.ake_pcn_create(C.byref(cfg), C.byref(h)) == 0, (name, lib.ake_last_error())
            try:
                want = [(k, tuple(shape)) for k, shape in entry["tensors"] if not k.endswith("num_batches_tracked")]
                shape4 = (C.c_int64 * 4)()
                got = []
                for i in range(lib.ake_pcn_num_tensors(h)):
                    nd = lib.ake_pcn_tensor_shape(h, i, C.byref(shape4))
                    got.append((lib.ake_pcn_tensor_name(h, i).decode(), tuple(shape4[:nd])))
                assert got == want, name
                assert lib.ake_pcn_workspace_bytes(h, 2, 70, 0) > 0 and lib.ake_pcn_workspace_bytes(h, 2, 70, 1) > 0
                assert lib.ake_pcn_workspace_bytes(h, 2, 70, 2) == 0   # the backward pass covers the default architecture only
            finally:
                lib.ake_pcn_destroy(h)
            net = ake.PitchClassNet(288, 12, int(opt.get("num_layers", 2)), 7, opt=ake.default_opt(**opt))
            assert [k for k in net.state_dict()] == [k for k, _ in entry["tensors"]]


@pytest.mark.parametrize("flag", ["only_semitones"])
def test_unsupported_architecture_switches_fail_loudly(flag):
    lib = _lib.lib()
    h = C.c_void_p()
    cfg = _cfg(**{flag: 1})
    assert lib.ake_pcn_create(C.byref(cfg), C.byref(h)) == _lib.AKE_ERR_UNSUPPORTED
    assert flag.encode() in lib.ake_last_error()
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.AKE_ERR_UNSUPPORTED)


def test_invalid_arguments():
    lib = _lib.lib()
    h = C.c_void_p()
    for cfg in (_cfg(pitch_classes=10), _cfg(pitches=100), _cfg(pitches=0)):
        assert lib.ake_pcn_create(C.byref(cfg), C.byref(h)) == _lib.AKE_ERR_INVALID
    assert lib.ake_pcn_create(None, C.byref(h)) == _lib.AKE_ERR_INVALID
    # librosa 0.9.2 ParameterError cases: hop not a multiple of 2^7, top filter beyond Nyquist
    assert lib.ake_cqt_create(44100.0, 8820, 288, 36, 0.0, 1.0, 0.01, C.byref(h)) == _lib.AKE_ERR_INVALID
    assert b"multiple of 2^7" in lib.ake_last_error()
    assert lib.ake_cqt_create(8000.0, 1280, 288, 36, 0.0, 1.0, 0.01, C.byref(h)) == _lib.AKE_ERR_INVALID
    assert lib.ake_cqt_create(48000.0, 9600, 288, 36, 0.0, 1.0, 1.5, C.byref(h)) == _lib.AKE_ERR_INVALID
    # compute entry points refuse null pointers before touching CUDA
    cfg = _cfg()
    assert lib.ake_pcn_create(C.byref(cfg), C.byref(h)) == 0
    assert lib.ake_pcn_forward_f32(h, None, 1, 64, None, 0, None, None, None, None, None, 0, None) == _lib.AKE_ERR_INVALID
    lib.ake_pcn_destroy(h)
    assert lib.ake_launch_count(1) >= 0 and lib.ake_launch_count(0) == 0
