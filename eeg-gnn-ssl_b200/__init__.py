"""dcgru_b200 -- B200-native diffusion-convolutional GRU (forward + backward) behind the Python class
surface of tsy935/eeg-gnn-ssl's ``model/cell.py`` and ``model/model.py``.

    from eeg_gnn_ssl_b200.model.cell import DCGRUCell
    from eeg_gnn_ssl_b200.model.model import DCRNNEncoder, DCGRUDecoder, \
        DCRNNModel_classification, DCRNNModel_nextTimePred

or put ``eeg-gnn-ssl_b200/`` first on ``sys.path`` so the reference's own
``from model.model import ...`` lines (train.py:9, train_ssl.py:5) resolve here unchanged
(see INTEGRATION.md).  The arithmetic lives in ``csrc/`` (CUDA, sm_100a) behind the C ABI of
``include/dcgru_b200.h``; there is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"
