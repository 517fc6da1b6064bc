"""Drop-in for the reference's ``model/cell.py``: same class names, constructor signatures,
attribute names, parameter creation order and ``state_dict`` keys, with the arithmetic
running in libdcgru_b200 (CUDA, sm_100a).

* ``DiffusionGraphConv`` is the parameter container ``dconv_gate`` / ``dconv_candidate`` that
  checkpoints and ``utils.build_finetune_model`` (utils.py:166-176) address by name.
* ``DCGRUCell.forward`` is one GRU step = the persistent layer kernel run for T = 1.
"""
import os
import sys

import torch
import torch.nn as nn

_pkg_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if __package__ in (None, "", "model"):          # imported as top-level ``model.cell`` (drop-in mode)
    sys.path.insert(0, os.path.dirname(_pkg_dir))
    import eeg_gnn_ssl_b200.ops as ops           # noqa: E402
else:
    from .. import ops


def num_supports_of(filter_type):
    """model/cell.py:151-158: only 'dual_random_walk' has two supports."""
    return 2 if filter_type == "dual_random_walk" else 1


class DiffusionGraphConv(nn.Module):
    """weight (C*M, out) with row index c*M + m, biases (out,) -- model/cell.py:18-48."""

    def __init__(self, num_supports, input_dim, hid_dim, num_nodes, max_diffusion_step, output_dim,
                 bias_start=0.0, filter_type='laplacian'):
        super().__init__()
        self._input_size = input_dim + hid_dim
        self._num_nodes = num_nodes
        self._max_diffusion_step = max_diffusion_step
        self._filter_type = filter_type
        self._num_matrices = num_supports * max_diffusion_step + 1
        # same tensors, same init calls, same RNG consumption as the reference
        self.weight = nn.Parameter(torch.empty(self._input_size * self._num_matrices, output_dim))
        self.biases = nn.Parameter(torch.empty(output_dim))
        nn.init.xavier_normal_(self.weight.data, gain=1.414)
        nn.init.constant_(self.biases.data, val=bias_start)

    def forward(self, supports, inputs, state, output_size, bias_start=0.0):
        raise NotImplementedError(
            "dcgru_b200: the diffusion convolution runs fused inside DCGRUCell / DCRNNEncoder / "
            "DCGRUDecoder kernels; DiffusionGraphConv only holds the parameters")


class DCGRUCell(nn.Module):
    def __init__(self, input_dim, num_units, max_diffusion_step, num_nodes, filter_type="laplacian",
                 nonlinearity='tanh', use_gc_for_ru=True):
        super().__init__()
        # model/cell.py:146 -- anything that is not the string 'tanh' selects relu
        self._activation = torch.tanh if nonlinearity == 'tanh' else torch.relu
        self._act_name = "tanh" if nonlinearity == 'tanh' else "relu"
        self._input_dim = input_dim
        self._num_nodes = num_nodes
        self._num_units = num_units
        self._max_diffusion_step = max_diffusion_step
        self._use_gc_for_ru = use_gc_for_ru
        self._num_supports = num_supports_of(filter_type)
        if not use_gc_for_ru:
            raise NotImplementedError("use_gc_for_ru=False has no implementation in the reference either "
                                      "(model/cell.py:218-219 is `pass`)")
        self.dconv_gate = DiffusionGraphConv(
            num_supports=self._num_supports, input_dim=input_dim, hid_dim=num_units, num_nodes=num_nodes,
            max_diffusion_step=max_diffusion_step, output_dim=num_units * 2, filter_type=filter_type)
        self.dconv_candidate = DiffusionGraphConv(
            num_supports=self._num_supports, input_dim=input_dim, hid_dim=num_units, num_nodes=num_nodes,
            max_diffusion_step=max_diffusion_step, output_dim=num_units, filter_type=filter_type)

    @property
    def output_size(self):
        return self._num_nodes * self._num_units

    def desc(self, num_supports=None):
        return ops.make_desc(self._num_nodes, self._input_dim, self._num_units, self._max_diffusion_step,
                             self._num_supports if num_supports is None else num_supports, self._act_name)

    def flat_params(self):
        return (self.dconv_gate.weight, self.dconv_gate.biases,
                self.dconv_candidate.weight, self.dconv_candidate.biases)

    def check_supports(self, supports):
        if self._max_diffusion_step > 0 and len(supports) != self._num_supports:
            raise ValueError(f"expected {self._num_supports} supports for this filter_type, got {len(supports)}")

    def forward(self, supports, inputs, state):
        """inputs (B, N*input_dim), state (B, N*num_units) -> (output, new_state), both (B, N*num_units)."""
        self.check_supports(supports)
        b = inputs.shape[0]
        p = ops.graph_poly(list(supports), b, self._num_nodes, self._max_diffusion_step)
        h_seq, _ = ops.encoder_layer(inputs.reshape(1, b, -1), state, p, *self.flat_params(), self.desc())
        out = h_seq[0]
        return out, out

    def init_hidden(self, batch_size):
        # CPU zeros, like the reference (model/cell.py:223-225); the caller moves them
        return torch.zeros(batch_size, self._num_nodes * self._num_units)
