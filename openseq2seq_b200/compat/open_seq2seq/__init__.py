"""`open_seq2seq` import surface re-implemented over the B200 engine (speech2text / Jasper path).

Class and function names, constructor signatures and params-dict schemas follow the reference so
that example_configs/speech2text/jasper*.py load unchanged; the bodies are new (no TF graph).
Anything outside the Jasper speech-to-text training path raises NotImplementedError when used.
"""
