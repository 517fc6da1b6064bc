"""Drop-in ``model`` package: ``model.cell`` / ``model.model`` are the CUDA-backed files in this directory.

The reference's ``train.py:20-22`` also imports its baseline models (``model.densecnn``, ``model.lstm``,
``model.cnnlstm``), which are not part of the DCGRU path and are not re-implemented here: when the reference
checkout is known -- ``DCGRU_REFERENCE_ROOT`` or any ``sys.path`` entry that holds a ``model/lstm.py`` -- its
``model/`` directory is appended to this package's search path, so those sub-modules resolve to the reference's
own files while ``model.cell`` / ``model.model`` keep resolving here (this directory comes first)."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))


def _reference_model_dir():
    roots = [_os.environ.get("DCGRU_REFERENCE_ROOT")] + list(_sys.path)
    for r in roots:
        if not r:
            continue
        d = _os.path.join(_os.path.abspath(r), "model")
        if d != _here and _os.path.isfile(_os.path.join(d, "lstm.py")) and _os.path.isfile(_os.path.join(d, "cell.py")):
            return d
    return None


_ref = _reference_model_dir()
if _ref is not None and _ref not in __path__:
    __path__.append(_ref)
