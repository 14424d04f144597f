"""Drop-in import surface: `open_seq2seq` (the reference's package name, plugin classes and config
system re-implemented over JasperEngine) and a minimal `tensorflow` stand-in exporting exactly the
symbols the speech2text configs touch.

    import openseq2seq_b200.compat as compat; compat.install()

appends this directory to sys.path (appended, not prepended: a real TensorFlow or a real
open_seq2seq checkout, if present, wins).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install():
    if _HERE not in sys.path:
        sys.path.append(_HERE)
    return _HERE
