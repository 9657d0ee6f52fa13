"""CPU: the C-ABI library loads, exports every symbol include/uvs_b200.h declares, and the ctypes
mirror of its structs has the C layout.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "uvs_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import uvs_b200
    return uvs_b200._lib


def test_every_declared_symbol_is_exported(lib):
    text = open(HEADER).read()
    declared = set(re.findall(r"^int\s+(uvs_\w+)\s*\(", text, flags=re.M))
    assert len(declared) >= 15
    handle = lib.load()
    for name in declared:
        assert hasattr(handle, name), name
    assert declared == set(lib.EXPORTED)
    assert handle.uvs_version() == 4


def test_struct_layout_matches_c(lib, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "uvs_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(uvs_filter_params), sizeof(uvs_noise_params),'
                   'sizeof(uvs_closed_loop_cfg), sizeof(uvs_closed_loop_io), offsetof(uvs_closed_loop_cfg, sweep),'
                   'offsetof(uvs_closed_loop_cfg, P0), offsetof(uvs_closed_loop_io, p_layout), sizeof(uvs_twist_cfg),'
                   'offsetof(uvs_twist_cfg, err_threshold), offsetof(uvs_closed_loop_cfg, servo_gain),'
                   'offsetof(uvs_closed_loop_io, flags_log), offsetof(uvs_twist_cfg, stale_fk),'
                   'offsetof(uvs_closed_loop_cfg, servo_substeps));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(lib.FilterParams), C.sizeof(lib.NoiseParams), C.sizeof(lib.ClosedLoopCfg),
            C.sizeof(lib.ClosedLoopIO), lib.ClosedLoopCfg.sweep.offset, lib.ClosedLoopCfg.P0.offset,
            lib.ClosedLoopIO.p_layout.offset, C.sizeof(lib.TwistCfg), lib.TwistCfg.err_threshold.offset,
            lib.ClosedLoopCfg.servo_gain.offset, lib.ClosedLoopIO.flags_log.offset, lib.TwistCfg.stale_fk.offset,
            lib.ClosedLoopCfg.servo_substeps.offset]
    assert got == want


def test_errors_are_reported_not_raised_across_the_boundary(lib):
    handle = lib.load()
    rc = handle.uvs_ibvs(2, 1, None, None, None, 0.2, None, None, None)      # nf = 2 is not a scene
    assert rc == -1
    assert "nf" in lib.last_error()
    p = lib.FilterParams()
    p.method, p.nf = 1, 4                                                    # ANALYTICAL is not a filter
    rc = handle.uvs_filter_step(C.byref(p), 1, 8, 8, 0, 8, 8, 1.0, None, None, None, None)
    assert rc == -1 and "method" in lib.last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "uncalibrated-visual-servoing_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("oracle/noise_model.py", ""), fn


def test_host_layer_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import uvs_b200
    with pytest.raises(lib.UvsError):
        uvs_b200.engine.fp64_peak()


def test_twist_loop_validates_before_touching_the_device(lib):
    """uvs_twist_loop / uvs_closed_loop reject bad scalar arguments with UVS_ERR_INVALID (no CUDA call is needed to say so)."""
    handle = lib.load()
    cfg, tw, io = lib.ClosedLoopCfg(), lib.TwistCfg(), lib.ClosedLoopIO()
    cfg.filt.method, cfg.filt.nf, cfg.filt.joseph, cfg.T, cfg.dt = 2, 4, 1, 10, 0.05
    io.stats, io.status = 8, 8                                   # never dereferenced on the error paths below
    tw.x0_mode, tw.focal_x0, tw.focal = 7, 1.0, 1.0
    assert handle.uvs_twist_loop(C.byref(cfg), C.byref(tw), 0, 1, C.byref(io), None) == -1
    assert "x0_mode" in lib.last_error()
    tw.x0_mode, tw.dp_zero_mask = 1, 64
    assert handle.uvs_twist_loop(C.byref(cfg), C.byref(tw), 0, 1, C.byref(io), None) == -1
    tw.dp_zero_mask, tw.focal_x0 = 0, 0.0
    assert handle.uvs_twist_loop(C.byref(cfg), C.byref(tw), 0, 1, C.byref(io), None) == -1 and "focal" in lib.last_error()
    tw.focal_x0 = 1024.0
    cfg.servo_gain = -0.5
    assert handle.uvs_twist_loop(C.byref(cfg), C.byref(tw), 0, 1, C.byref(io), None) == -1
    assert handle.uvs_closed_loop(C.byref(cfg), 0, 1, C.byref(io), None) == -1 and "servo_gain" in lib.last_error()
    cfg.servo_gain = 0.0
    cfg.filt.mode = 2                                            # the twist loop has no block-diagonal family
    assert handle.uvs_twist_loop(C.byref(cfg), C.byref(tw), 0, 1, C.byref(io), None) == -3
    cfg.filt.mode, cfg.filt.method = 0, 1                        # ANALYTICAL is a twist-loop method, not a closed-loop one
    assert handle.uvs_closed_loop(C.byref(cfg), 0, 1, C.byref(io), None) == -1 and "method" in lib.last_error()
    assert handle.uvs_twist_loop(C.byref(cfg), C.byref(tw), 0, 0, C.byref(io), None) == 0      # empty batch: no-op
    # range checks that used to fail silently inside the kernels (NaN weights, frozen filters): ADVICE r01
    cfg.filt.method = 5
    for field, bad, word in (("dt", 0.0, "dt"), ("sweep_param", 3, "sweep_param")):
        old = getattr(cfg, field)
        setattr(cfg, field, bad)
        assert handle.uvs_closed_loop(C.byref(cfg), 0, 1, C.byref(io), None) == -1 and word in lib.last_error()
        setattr(cfg, field, old)
    cfg.filt.kernel_bw = 0.0
    assert handle.uvs_closed_loop(C.byref(cfg), 0, 1, C.byref(io), None) == -1 and "kernel_bw" in lib.last_error()
    cfg.filt.kernel_bw, cfg.filt.method, cfg.filt.fpi_epoch_max = 10.0, 3, 0
    assert handle.uvs_closed_loop(C.byref(cfg), 0, 1, C.byref(io), None) == -1 and "fpi_epoch_max" in lib.last_error()
    cfg.filt.fpi_epoch_max, cfg.sweep_len, cfg.sweep_epoch = 10, 12, 0
    assert handle.uvs_closed_loop(C.byref(cfg), 0, 1, C.byref(io), None) == -1 and "sweep_epoch" in lib.last_error()
    assert handle.uvs_reduce_stats(1, 0, 1, 1, 513, 8, 8, 8, 8, None) == -1 and "n_levels" in lib.last_error()
