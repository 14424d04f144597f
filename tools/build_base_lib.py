"""Build libos2s_b200 from the csrc/ of a git revision into openseq2seq_b200/lib/libos2s_b200_base.so, for
same-box A/B runs of two builds (OS2S_LIB_PATH selects the library):  python tools/build_base_lib.py [rev]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openseq2seq_b200 import build as B  # noqa: E402

rev = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
tmp = tempfile.mkdtemp(prefix="os2s_base_")
subprocess.check_call("git archive %s openseq2seq_b200/csrc include | tar -x -C %s" % (rev, tmp), shell=True, cwd=ROOT)
B.CSRC = os.path.join(tmp, "openseq2seq_b200", "csrc")
B.HERE = os.path.join(tmp, "openseq2seq_b200")
B.LIB_DIR = os.path.join(ROOT, "openseq2seq_b200", "lib")
B.LIB_PATH = os.path.join(B.LIB_DIR, "libos2s_b200_base.so")
print(B.build(force=True))
