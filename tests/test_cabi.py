"""CPU tests of the drop-in boundary: the C-ABI library builds, loads without a GPU / driver and
exports every symbol include/os2s.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "os2s.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(os2s_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for need in ["os2s_logmel_forward", "os2s_conv1d_fwd", "os2s_conv1d_dgrad", "os2s_conv1d_wgrad",
                 "os2s_bn_stats", "os2s_bn_apply_fwd", "os2s_bn_bwd", "os2s_fc_fwd", "os2s_fc_bwd",
                 "os2s_ctc_loss_fwd_bwd", "os2s_ctc_greedy", "os2s_opt_step", "os2s_last_error"]:
        assert need in syms


def test_library_builds_loads_and_exports_every_declared_symbol():
    from openseq2seq_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in _declared_symbols():
        assert hasattr(lib, name), "missing export: " + name
    lib.os2s_last_error.restype = ctypes.c_char_p
    assert lib.os2s_version() >= 100
    assert lib.os2s_opt_chunk_elems() > 0


def test_argument_validation_needs_no_gpu():
    from openseq2seq_b200 import _lib as L
    lib = L.load()
    assert lib.os2s_conv1d_fwd(None, None, None, 1, 1, 64, 64, 1, 1, 0, 0, None, None) == -1
    assert b"null pointer" in lib.os2s_last_error()
    assert lib.os2s_fc_fwd(None, None, None, None, 1, 1, 1, None) == -1


def test_opt_hparams_struct_layout_matches_header(tmp_path):
    """The ctypes mirror of os2s_opt_hparams has the layout the C compiler gives the header's struct
    (checked by compiling a probe against include/os2s.h with gcc)."""
    import subprocess
    from openseq2seq_b200.engine import OptHParams
    fields = [f[0] for f in OptHParams._fields_]
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "os2s.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(os2s_opt_hparams));\n' +
                   "".join('  printf("%%zu\\n", offsetof(os2s_opt_hparams, %s));\n' % f for f in fields) +
                   "  return 0;\n}\n")
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert ctypes.sizeof(OptHParams) == out[0]
    assert [getattr(OptHParams, f).offset for f in fields] == out[1:]
