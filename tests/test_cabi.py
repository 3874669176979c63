"""CPU suite, part 2: the C-ABI library builds for sm_100a, loads, and exports what include/regnet_b200.h
declares.  No compute calls (no GPU here); argument validation that happens before any CUDA call is exercised."""
import ctypes
import subprocess

from regnet_for_3d_grasping_b200 import _lib


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/regnet_b200.h but not exported"
    assert set(declared) == set(_lib._SIGNATURES.keys()), "ctypes table and header disagree"


def test_library_is_sm100a_with_blackwell_instructions(lib_path):
    elf = subprocess.check_output(["cuobjdump", "-lelf", lib_path], text=True)
    assert "sm_100a" in elf
    sass = subprocess.check_output(["cuobjdump", "-sass", lib_path], text=True)
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UCGABAR", "REDUX"):
        assert mnemonic in sass, f"{mnemonic} missing: the tcgen05/TMA/cluster paths were not compiled in"


def test_abi_version_and_error_channel(lib_path):
    lib = _lib.load()
    assert lib.regnet_abi_version() == 2
    # argument errors are detected before any CUDA call, so they work without a GPU
    rc = lib.regnet_farthest_point_sample(ctypes.c_void_p(16), 0, 0, 0, 1, 10, 11, ctypes.c_void_p(16), None, None)
    assert rc == 1 and b"num_points" in lib.regnet_last_error()
    rc = lib.regnet_farthest_point_sample(ctypes.c_void_p(16), 0, 0, 0, 1, 10, 0, ctypes.c_void_p(16), None, None)
    assert rc == 1 and b"num_centroids" in lib.regnet_last_error()
    rc = lib.regnet_point_search(ctypes.c_void_p(16), 0, 0, 0, ctypes.c_void_p(16), 0, 0, 0, 1, 8, 8, 2, None, None, None)
    assert rc == 1 and b"3 neighbours" in lib.regnet_last_error()
    rc = lib.regnet_point_search(ctypes.c_void_p(16), 0, 0, 0, ctypes.c_void_p(16), 0, 0, 0, 1, 8, 2, 3, None, None, None)
    assert rc == 1 and b"at least 3" in lib.regnet_last_error()
    rc = lib.regnet_ball_query(ctypes.c_void_p(16), 0, 0, 0, ctypes.c_void_p(16), 0, 0, 0, 1, 8, 8, 0.1, 0, None, None, None, None)
    assert rc == 1 and b"num_neighbours" in lib.regnet_last_error()
    rc = lib.regnet_conv1x1_train(ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 8, 12, ctypes.c_void_p(16), ctypes.c_void_p(16),
                                  8, 8, ctypes.c_void_p(16), None, 3, None)
    assert rc == 1 and b"multiple of 8" in lib.regnet_last_error()
    rc = lib.regnet_conv1x1_train(ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 8, 16, ctypes.c_void_p(16), ctypes.c_void_p(16),
                                  8, 8, ctypes.c_void_p(16), None, 2, None)
    assert rc == 1 and b"passes" in lib.regnet_last_error()
    cfg = _lib.ScoreNetConfig()
    cfg.batch, cfg.num_points = 1, 100
    for i, m in enumerate((200, 50, 10)):
        cfg.num_centroids[i], cfg.radius[i], cfg.num_neighbours[i] = m, 0.1, 64
    h = ctypes.c_void_p()
    rc = lib.regnet_scorenet_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == 1 and b"num_centroids[0]" in lib.regnet_last_error()


def test_shims_refuse_cpu_tensors(lib_path):
    import pytest
    import torch
    from regnet_for_3d_grasping_b200 import pn2_ext
    x = torch.rand(1, 3, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        pn2_ext.farthest_point_sample(x, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        pn2_ext.ball_query(x, x, 0.1, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        pn2_ext.point_search(x, x, 3)
    for name in ("furthest_point_sample", "three_nn", "three_interpolate", "group_points_forward",
                 "group_points_backward", "interpolate_backward"):
        assert callable(getattr(pn2_ext, name))
