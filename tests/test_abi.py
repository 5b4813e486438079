"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest
import torch

from util import CFG, ROOT


def _declared_symbols():
    import glob
    syms = set()
    for path in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        txt = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        syms |= set(re.findall(r"\b(tvb_[a-z0-9_]+)\s*\(", txt))
    return sorted(syms)


def test_library_exports_every_declared_symbol():
    from usdm_b200 import engine
    lib = engine.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"libtvb.so does not export {s}"
        assert s in engine._SIGNATURES, f"{s} has no ctypes signature in usdm_b200/engine.py"
    assert set(engine._SIGNATURES) == set(syms)
    assert b"sm_100a" in lib.tvb_version()


def test_only_sm100a_code_in_library():
    import subprocess, shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from usdm_b200 import engine
    out = subprocess.run([cuobjdump, "-lelf", engine.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from usdm_b200 import engine
    lib = engine.load_library()
    cfg = engine.TvbConfig(80, 10000, 1280, 1024, 4096, 16, 24, 31, 16, 2, 1e-4)
    h = C.c_void_p()
    rc = lib.tvb_create(C.byref(cfg), 0, C.byref(h))
    assert rc == engine.ERR_CUDA
    assert b"no CPU path" in lib.tvb_last_error()
    with pytest.raises(RuntimeError):
        engine.check(rc)


def test_create_rejects_unsupported_config():
    from usdm_b200 import engine
    lib = engine.load_library()
    cfg = engine.TvbConfig(80, 10000, 1280, 1000, 4096, 16, 24, 31, 16, 2, 1e-4)   # head_dim != 64
    h = C.c_void_p()
    rc = lib.tvb_create(C.byref(cfg), 0, C.byref(h))
    assert rc == engine.ERR_INVALID
    with pytest.raises(ValueError):
        engine.check(rc)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_voicebox_has_no_cpu_path():
    import usdm_b200
    small = dict(CFG, num_hidden_layers=2)
    m = usdm_b200.Voicebox(**small)
    x = torch.zeros(1, 8, dtype=torch.long)
    y = torch.zeros(1, 80, 8)
    with pytest.raises(RuntimeError):
        m.estimator(x, y, y, torch.tensor([0.5]), torch.tensor([8]))
    with pytest.raises(RuntimeError):
        m.generate(x, y, torch.tensor([8]), 2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "usdm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references the oracle"
                assert "/root/reference" not in src


def test_pair_gemm_tile_width_rule():
    """Host logic of the CTA-pair GEMM's tile width (tvb.cu: pair_gemm_bn_rule) on a 148-SM device.  Full waves must stay on
    256-column tiles (narrow tiles there cost configs[1] 11 %: QKV +29 %, FFN-in +53 %); launches below one wave
    of the 74 CTA pairs take 128-column tiles when those need fewer (waves x width)."""
    from usdm_b200 import engine
    lib = engine.load_library()
    bn = lambda rows, n: lib.tvb_debug_pair_gemm_bn(148, rows, n)
    pad = lambda n_seq, tokens: n_seq * (-(-tokens // 256) * 256)        # every sequence is padded to whole 256-row blocks
    # BASELINE configs[1] (2 x 16 sequences of 1134 tokens), configs[2] (2 x 8 x 2856), and a 64-utterance ragged batch: always 256
    for rows in (pad(32, 1134), pad(16, 2856), pad(32, 862), pad(128, 700)):
        for n in (256, 1024, 2048, 3072, 4096):
            assert bn(rows, n) == 256, (rows, n)
    # one 5 s utterance + 272-frame prompt under CFG (2 x 703 tokens -> 1536 rows): N = 1024 is 24 tiles on 74 pairs -> narrow tiles;
    # QKV (72 tiles, a tie) stays wide
    rows = pad(2, 703)
    assert rows == 1536
    assert bn(rows, 1024) == 128 and bn(rows, 256) == 128
    assert bn(rows, 3072) == 256
    assert bn(rows, 4096) == 256                  # 96 tiles > 74 pairs: a full wave exists
    # N that only 128-column tiles can cover
    assert bn(rows, 384) == 128
    # the choice only depends on (rows, N): monotone hand-over to wide tiles as the batch grows
    widths = [bn(256 * k, 1024) for k in range(1, 40)]
    first_wide = widths.index(256)
    assert all(w == 128 for w in widths[:first_wide]) and all(w == 256 for w in widths[first_wide:])
    assert lib.tvb_debug_pair_gemm_bn(148, 100, 1024) < 0 and lib.tvb_debug_pair_gemm_bn(148, 256, 100) < 0
