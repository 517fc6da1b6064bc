"""Importable alias of the ``eeg-gnn-ssl_b200/`` source directory (a hyphen cannot be imported):
``import eeg_gnn_ssl_b200.model.cell`` resolves to ``eeg-gnn-ssl_b200/model/cell.py``."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "eeg-gnn-ssl_b200"))
__version__ = "0.1.0"
