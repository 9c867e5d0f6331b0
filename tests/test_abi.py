"""The C-ABI library loads on a CPU-only box and exports exactly what include/x3d_b200.h declares;
the product path refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "x3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(x3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from x3d_tf_b200 import _lib
    from x3d_tf_b200 import build as b
    b.build()
    names = _declared_functions()
    assert len(names) >= 10
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/x3d_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding table and header disagree"
    assert _lib.lib().x3d_version() == 100


def test_struct_layouts_match_header(tmp_path):
    """Every field of the argument structs sits where a C compiler puts it: a probe that includes
    include/x3d_b200.h is compiled with gcc and its offsetof / sizeof are compared with ctypes."""
    import subprocess
    from x3d_tf_b200 import _lib
    structs = {"x3d_pw_args": _lib.PwArgs, "x3d_pw_tc_args": _lib.PwTcArgs}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "x3d_b200.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == ctypes.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"
    assert _lib.PwArgs.rows_per_clip.offset == 80 and _lib.PwTcArgs.rows_per_clip.offset == 88


def test_native_crc32c_matches_python():
    from x3d_tf_b200 import _lib, tf_bundle
    h = _lib.lib()
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 64, 1000, 4099):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        tf_bundle.set_native_crc32c(None)
        want = tf_bundle.crc32c(data)
        assert h.x3d_crc32c(data, n, 0) == want
        assert h.x3d_crc32c(data[n // 2:], n - n // 2, h.x3d_crc32c(data[:n // 2], n // 2, 0)) == want
    assert h.x3d_crc32c(b"123456789", 9, 0) == 0xE3069283


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call, with a message."""
    from x3d_tf_b200 import _lib
    h = _lib.lib()
    assert h.x3d_dw3x3x3_fwd(None, None, None, None, None, 1, 1, 1, 1, 8, 1, 0, 0, 0, None) == -1
    assert b"null" in h.x3d_last_error()
    assert h.x3d_dw_partial_blocks(16, 56, 56, 56, 1, 1) == 49        # 7x7 tiles of 8x8
    assert h.x3d_dw_partial_blocks(16, 56, 56, 54, 1, 1) == 0     # C not a multiple of 8
    assert h.x3d_softmax_viewmean_fwd(1, 1, 7, 400, 2, None) == -1 and b"multiple" in h.x3d_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from x3d_tf_b200 import model as M
    from x3d_tf_b200.config import get_config
    M.reset_block_counters()
    m = M.X3D(get_config("X3D_XS"))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        m(np.zeros((10, 4, 32, 32, 3), np.float32))
    import x3d_tf_b200.model as mod
    src = open(mod.__file__).read() + open(os.path.join(os.path.dirname(mod.__file__), "ops.py")).read()
    assert "oracle" not in src.replace("oracle/", ""), "product code must not import the oracle"


def test_model_variables_and_summary(summaries, capsys):
    from x3d_tf_b200 import model as M
    from x3d_tf_b200.arch import build_arch, variable_shapes
    from x3d_tf_b200.config import get_config
    for variant, shape in (("X3D_M", (16, 224, 224, 3)), ("X3D_XL", (16, 312, 312, 3))):
        M.reset_block_counters()
        cfg = get_config(variant)
        m = M.X3D(cfg)
        assert {k: v.shape for k, v in m.named_variables().items()} == dict(variable_shapes(build_arch(cfg)))
        text = m.summary(shape)
        g = summaries[variant]
        for row in g["layers"]:
            assert str(row["params"]) in text
        assert f"Total params: {g['total']:,}" in text and f"Trainable params: {g['trainable']:,}" in text
        assert m.stages[-1]._inner_channels == build_arch(cfg).conv5_channels
    # the reference's process-global block counter: a second X3D-L flips SE placement
    M.reset_block_counters()
    a = M.X3D(get_config("X3D_L"))
    b = M.X3D(get_config("X3D_L"))
    assert a.stages[0].blocks[0].bottleneck.has_se and not b.stages[0].blocks[0].bottleneck.has_se
