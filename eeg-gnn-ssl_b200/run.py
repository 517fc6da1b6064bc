"""Launcher: run an UNMODIFIED script of the reference (``train.py`` / ``train_ssl.py``) on the CUDA-backed
``model`` package.

    cd /path/to/eeg-gnn-ssl && python /path/to/repo/eeg-gnn-ssl_b200/run.py train.py --task detection ...

``python train.py`` puts the script's own directory first on ``sys.path``, so a ``PYTHONPATH`` entry cannot make
``from model.model import ...`` (train.py:19, train_ssl.py:18) resolve anywhere but the reference's ``model/``.
This launcher fixes the order -- [this directory, the script's directory, the rest] -- and runs the script with
``runpy`` as ``__main__``; nothing of the reference is edited.  ``--dropin-check`` imports the script without
running its ``__main__`` block and prints where the classes it bound came from (used by tests/test_cpu_dropin.py).
"""
import json
import os
import runpy
import sys


def main(argv):
    check = False
    if argv and argv[0] == "--dropin-check":
        check, argv = True, argv[1:]
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    here = os.path.dirname(os.path.abspath(__file__))
    ref_root = os.path.dirname(script)
    os.environ.setdefault("DCGRU_REFERENCE_ROOT", ref_root)
    rest = [p for p in sys.path if os.path.abspath(p or ".") not in (here, ref_root)]
    sys.path[:] = [here, ref_root] + rest
    sys.argv = [script] + argv[1:]
    if not check:
        runpy.run_path(script, run_name="__main__")
        return
    g = runpy.run_path(script, run_name="__dropin_check__")
    out = {}
    for name in ("DCRNNModel_classification", "DCRNNModel_nextTimePred", "DenseCNN", "LSTMModel", "CNN_LSTM"):
        if name in g:
            out[name] = os.path.abspath(sys.modules[g[name].__module__].__file__)
    print("DROPIN " + json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1:])
