"""Drop-in for the DCRNN parts of the reference's ``model/model.py``: ``DCRNNEncoder``,
``DCGRUDecoder`` and the two task models that call them, same signatures and ``state_dict`` keys.

Encoder: one persistent CUDA kernel per layer over all T steps (model/model.py:90-99).
Decoder: one persistent CUDA kernel for all To steps x L cells + projection (model/model.py:182-202).
"""
import random

import numpy as np

import torch
import torch.nn as nn

try:
    from .cell import DCGRUCell, ops
except ImportError:                      # drop-in mode: this file is ``model/model.py`` on sys.path
    from model.cell import DCGRUCell, ops


class DCRNNEncoder(nn.Module):
    def __init__(self, input_dim, max_diffusion_step, hid_dim, num_nodes, num_rnn_layers,
                 dcgru_activation=None, filter_type='laplacian', device=None):
        super().__init__()
        self.hid_dim = hid_dim
        self.num_rnn_layers = num_rnn_layers
        self._device = device
        self._num_nodes = num_nodes
        self._max_diffusion_step = max_diffusion_step
        cells = []
        for layer in range(num_rnn_layers):              # cell 0 sees the input features, the rest hid_dim
            cells.append(DCGRUCell(input_dim=input_dim if layer == 0 else hid_dim, num_units=hid_dim,
                                   max_diffusion_step=max_diffusion_step, num_nodes=num_nodes,
                                   nonlinearity=dcgru_activation, filter_type=filter_type))
        self.encoding_cells = nn.ModuleList(cells)

    def forward(self, inputs, initial_hidden_state, supports):
        """inputs (T,B,N,Fin), initial_hidden_state (L,B,N*H) ->
        (output_hidden (L,B,N*H), top-layer sequence (T,B,N*H))."""
        t_len, b = inputs.shape[0], inputs.shape[1]
        cur = inputs.flatten(2) if inputs.dim() == 4 else inputs     # a view: no transpose copy
        self.encoding_cells[0].check_supports(supports)
        p = ops.graph_poly(list(supports), b, self._num_nodes, self._max_diffusion_step)
        last = []
        for layer, cell in enumerate(self.encoding_cells):
            cur, h_last = ops.encoder_layer(cur, initial_hidden_state[layer], p, *cell.flat_params(),
                                            cell.desc())
            last.append(h_last)
        return torch.stack(last, dim=0), cur

    def forward_head(self, inputs, initial_hidden_state, supports, sel_t, drop_mask, fc):
        """The classification model's use of the encoder: all layers, then the fused head on the top layer's
        sequence (inputs / initial_hidden_state as in forward; sel_t (B) int32 = seq_len-1) -> logits (B,C)."""
        b = inputs.shape[1]
        cur = inputs.flatten(2) if inputs.dim() == 4 else inputs
        self.encoding_cells[0].check_supports(supports)
        p = ops.graph_poly(list(supports), b, self._num_nodes, self._max_diffusion_step)
        cells = list(self.encoding_cells)
        for layer, cell in enumerate(cells[:-1]):
            cur, _ = ops.encoder_layer(cur, initial_hidden_state[layer], p, *cell.flat_params(), cell.desc())
        top = cells[-1]
        return ops.encoder_top_head(cur, initial_hidden_state[len(cells) - 1], p, *top.flat_params(), fc.weight, fc.bias,
                                    top.desc(), sel_t, drop_mask)

    def init_hidden(self, batch_size):
        return torch.stack([c.init_hidden(batch_size) for c in self.encoding_cells], dim=0)


class DCGRUDecoder(nn.Module):
    def __init__(self, input_dim, max_diffusion_step, num_nodes, hid_dim, output_dim, num_rnn_layers,
                 dcgru_activation=None, filter_type='laplacian', device=None, dropout=0.0):
        super().__init__()
        self.input_dim = input_dim
        self.hid_dim = hid_dim
        self.num_nodes = num_nodes
        self.output_dim = output_dim
        self.num_rnn_layers = num_rnn_layers
        self._device = device
        self._max_diffusion_step = max_diffusion_step
        # creation order matters for seed-for-seed init parity and the key layout of checkpoints:
        # the shared upper cell first, then cell 0, then the projection (model/model.py:126-146)
        shared = DCGRUCell(input_dim=hid_dim, num_units=hid_dim, max_diffusion_step=max_diffusion_step,
                           num_nodes=num_nodes, nonlinearity=dcgru_activation, filter_type=filter_type)
        first = DCGRUCell(input_dim=input_dim, num_units=hid_dim, max_diffusion_step=max_diffusion_step,
                          num_nodes=num_nodes, nonlinearity=dcgru_activation, filter_type=filter_type)
        self.decoding_cells = nn.ModuleList([first] + [shared] * (num_rnn_layers - 1))
        self.projection_layer = nn.Linear(self.hid_dim, self.output_dim)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, inputs, initial_hidden_state, supports, teacher_forcing_ratio=None):
        """inputs (To,B,N,Fo) -> outputs (To,B,N*Fo)."""
        if self.input_dim != self.output_dim:
            raise ValueError("the decoder feeds its projection back as input: input_dim must equal output_dim")
        to_len, b = inputs.shape[0], inputs.shape[1]
        targets = inputs.reshape(to_len, b, -1)
        # one python draw per step for the whole batch, in the reference's order (model/model.py:198-202)
        mask = 0
        if teacher_forcing_ratio is not None:
            for t in range(to_len):
                if random.random() < teacher_forcing_ratio:
                    mask |= 1 << t
        # nn.Dropout before the projection (model/model.py:192): masks drawn by torch, applied in-kernel
        drop = None
        if self.training and self.dropout.p > 0:
            ones = torch.ones((b, self.num_nodes, self.hid_dim), device=targets.device)
            drop = torch.stack([self.dropout(ones) for _ in range(to_len)], dim=0)
        cells = list(self.decoding_cells)
        cells[0].check_supports(supports)
        uniq, index = [], []
        for c in cells:
            for i, u in enumerate(uniq):
                if u is c:
                    index.append(i)
                    break
            else:
                index.append(len(uniq))
                uniq.append(c)
        flat = [t for c in uniq for t in c.flat_params()]
        p = ops.graph_poly(list(supports), b, self.num_nodes, self._max_diffusion_step)
        return ops.decoder(targets, initial_hidden_state, p, self.projection_layer.weight,
                           self.projection_layer.bias, drop, cells[0].desc(), self.num_rnn_layers, to_len,
                           mask, index, flat)


# ---------------------------------------------------------------------------------------------------
# task models (callers of the hot path; thin torch glue)
# ---------------------------------------------------------------------------------------------------
class DCRNNModel_classification(nn.Module):
    def __init__(self, args, num_classes, device=None):
        super().__init__()
        self.num_nodes = args.num_nodes
        self.num_rnn_layers = args.num_rnn_layers
        self.rnn_units = args.rnn_units
        self._device = device
        self.num_classes = num_classes
        self.encoder = DCRNNEncoder(input_dim=args.input_dim, max_diffusion_step=args.max_diffusion_step,
                                    hid_dim=args.rnn_units, num_nodes=args.num_nodes,
                                    num_rnn_layers=args.num_rnn_layers,
                                    dcgru_activation=args.dcgru_activation, filter_type=args.filter_type)
        self.fc = nn.Linear(args.rnn_units, num_classes)
        self.dropout = nn.Dropout(args.dropout)
        self.relu = nn.ReLU()

    def forward(self, input_seq, seq_lengths, supports):
        """input_seq (B,T,N,Fin), seq_lengths (B,) -> pooled logits (B, num_classes)."""
        b = input_seq.shape[0]
        # zeros created on the device (init_hidden() keeps the reference's CPU-tensor contract for callers that use it)
        h0 = torch.zeros(self.num_rnn_layers, b, self.num_nodes * self.rnn_units, device=input_seq.device)
        # layers below the top one as usual; the top layer and the head (last relevant step -> dropout -> ReLU -> fc ->
        # max over nodes, model/model.py:257-270) run as one fused operator (csrc/head.cu): the head's gradient reaches
        # the BPTT kernel as one (B,N*H) slab instead of a dense (T,B,N*H) tensor
        drop = None
        if self.training and self.dropout.p > 0:     # mask drawn by torch (same RNG stream as nn.Dropout), applied in-kernel
            drop = self.dropout(torch.ones((b, self.num_nodes, self.rnn_units), device=input_seq.device))
        sel = (seq_lengths.to(input_seq.device) - 1).to(torch.int32)
        return self.encoder.forward_head(input_seq.transpose(0, 1), h0, supports, sel, drop, self.fc)


class DCRNNModel_nextTimePred(nn.Module):
    def __init__(self, args, device=None):
        super().__init__()
        self.num_nodes = args.num_nodes
        self.num_rnn_layers = args.num_rnn_layers
        self.rnn_units = args.rnn_units
        self._device = device
        self.output_dim = args.output_dim
        self.cl_decay_steps = args.cl_decay_steps
        self.use_curriculum_learning = bool(args.use_curriculum_learning)
        self.encoder = DCRNNEncoder(input_dim=args.input_dim, max_diffusion_step=args.max_diffusion_step,
                                    hid_dim=args.rnn_units, num_nodes=args.num_nodes,
                                    num_rnn_layers=args.num_rnn_layers,
                                    dcgru_activation=args.dcgru_activation, filter_type=args.filter_type)
        self.decoder = DCGRUDecoder(input_dim=args.output_dim, max_diffusion_step=args.max_diffusion_step,
                                    num_nodes=args.num_nodes, hid_dim=args.rnn_units,
                                    output_dim=args.output_dim, num_rnn_layers=args.num_rnn_layers,
                                    dcgru_activation=args.dcgru_activation, filter_type=args.filter_type,
                                    device=device, dropout=args.dropout)

    def forward(self, encoder_inputs, decoder_inputs, supports, batches_seen=None):
        """(B,T,N,Fin), (B,To,N,Fo) -> predictions (B,To,N,Fo)."""
        b, to_len, n, _ = decoder_inputs.shape
        h0 = torch.zeros(self.num_rnn_layers, b, self.num_nodes * self.rnn_units, device=encoder_inputs.device)
        context, _ = self.encoder(encoder_inputs.transpose(0, 1), h0, supports)
        ratio = None
        if self.training and self.use_curriculum_learning and batches_seen is not None:
            # inverse-sigmoid scheduled sampling (utils.py:385-390)
            # np.exp like the reference: saturates to inf (ratio 0) instead of raising OverflowError in long runs
            with np.errstate(over="ignore"):
                ratio = float(self.cl_decay_steps / (self.cl_decay_steps + np.exp(batches_seen / self.cl_decay_steps)))
        out = self.decoder(decoder_inputs.transpose(0, 1), context, supports, teacher_forcing_ratio=ratio)
        return out.reshape(to_len, b, n, -1).transpose(0, 1)
