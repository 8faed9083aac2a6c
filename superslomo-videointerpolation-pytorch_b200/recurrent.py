"""Bidirectional convolutional LSTM / GRU bottleneck of the recurrent configuration (SSMR,
configs/superslomo_recurrent.ini:97,105 `BOTTLENECK=CLSTM`), so that `FullModel(cfg)` runs both of the
reference's configurations.  Like the rest of the U-Nets this is plain torch/cuDNN and OUT OF SCOPE of
the B200 rebuild; it exists because the window loop of the synthesis path (SURVEY.md section 8 rows
a6/a7, config C4) is driven by it: the windows of one sample are coupled through this module, which is
why work is sharded over samples and timesteps, never over windows (sharding.py).

Parameter names and shapes follow the reference's modules so its checkpoints load unchanged:
  scripts/models/CLSTM/convlstm.py:11-57, 180-204   `forward_net|reverse_net.cell_list.<l>.conv.{weight,bias}`
  scripts/models/CLSTM/convgru.py:12-52, 159-185    `...cell_list.<l>.{update_gate,reset_gate,out_gate}.{weight,bias}`
Each direction has hidden_channels // 2 channels; the two outputs are concatenated on the channel axis with
the reverse direction flipped back in time (convlstm.py:196-202).

What is done differently from the reference's step loop: a gate convolution over cat([x_t, h]) is linear in
its two halves, so the input half (`weight[:, :C_in]` and the bias) runs ONCE over all T windows as one batched
convolution and only the hidden half (`weight[:, C_in:]`) stays in the sequential loop -- T-fold fewer launches
for half of the FLOPs; no torch.cat / torch.stack of the inputs per step.  The hidden state is created on the
input's device (the reference hard-codes .cuda(), convlstm.py:54).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _Cell(nn.Module):
    """Holds the reference-shaped gate convolutions of one layer of one direction."""

    def __init__(self, kind, in_channels, hidden_channels, kernel_size):
        super().__init__()
        self.kind, self.cin, self.hid = kind, in_channels, hidden_channels
        pad = (kernel_size[0] // 2, kernel_size[1] // 2)
        self.pad = pad
        mk = lambda cout: nn.Conv2d(in_channels + hidden_channels, cout, kernel_size, padding=pad, bias=True)
        if kind == "lstm":
            self.conv = mk(4 * hidden_channels)                        # i, f, o, g   convlstm.py:37-38
        else:
            self.update_gate, self.reset_gate, self.out_gate = mk(hidden_channels), mk(hidden_channels), mk(hidden_channels)

    def _halves(self, conv):
        return conv.weight[:, :self.cin], conv.weight[:, self.cin:], conv.bias

    def _run_joint(self, x):
        """The reference's step loop as it stands (one convolution over cat([x_t, h]) per step): same summation
        order inside the convolution as the reference, used to pin the module against it (split_input=False)."""
        B, T, hid = x.shape[0], x.shape[1], self.hid
        h = x.new_zeros((B, hid) + tuple(x.shape[3:]))
        c = torch.zeros_like(h)
        outs = []
        for t in range(T):
            if self.kind == "lstm":
                i, f, o, cand = torch.split(self.conv(torch.cat([x[:, t], h], dim=1)), hid, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(cand)
                h = torch.sigmoid(o) * torch.tanh(c)
            else:
                xin = torch.cat([x[:, t], h], dim=1)
                update, reset = torch.sigmoid(self.update_gate(xin)), torch.sigmoid(self.reset_gate(xin))
                cand = torch.tanh(self.out_gate(torch.cat([x[:, t], h * reset], dim=1)))
                h = h * (1 - update) + cand * update
            outs.append(h)
        return torch.stack(outs, dim=1)

    def run(self, x, split_input=True):
        """x: B x T x C x H x W -> hidden states B x T x hid x H x W (zero initial state)."""
        if not split_input:
            return self._run_joint(x)
        B, T = x.shape[0], x.shape[1]
        flat = x.reshape(B * T, *x.shape[2:])
        hid, pad = self.hid, self.pad
        h = x.new_zeros((B, hid) + tuple(x.shape[3:]))
        outs = []
        if self.kind == "lstm":
            wx, wh, b = self._halves(self.conv)
            gx = F.conv2d(flat, wx, b, padding=pad).view(B, T, 4 * hid, *x.shape[3:])
            c = torch.zeros_like(h)
            for t in range(T):
                g = gx[:, t] + F.conv2d(h, wh, None, padding=pad)
                i, f, o, cand = torch.split(g, hid, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(cand)    # convlstm.py:44-51
                h = torch.sigmoid(o) * torch.tanh(c)
                outs.append(h)
        else:
            uwx, uwh, ub = self._halves(self.update_gate)
            rwx, rwh, rb = self._halves(self.reset_gate)
            owx, owh, ob = self._halves(self.out_gate)
            # update and reset gates read the same input: one convolution for both, over all T
            ur_x = F.conv2d(flat, torch.cat([uwx, rwx], 0), torch.cat([ub, rb], 0), padding=pad).view(B, T, 2 * hid, *x.shape[3:])
            o_x = F.conv2d(flat, owx, ob, padding=pad).view(B, T, hid, *x.shape[3:])
            urwh = torch.cat([uwh, rwh], 0)
            for t in range(T):
                ur = torch.sigmoid(ur_x[:, t] + F.conv2d(h, urwh, None, padding=pad))
                update, reset = ur[:, :hid], ur[:, hid:]
                cand = torch.tanh(o_x[:, t] + F.conv2d(h * reset, owh, None, padding=pad))   # convgru.py:41-47
                h = h * (1 - update) + cand * update
                outs.append(h)
        return torch.stack(outs, dim=1)


class _Direction(nn.Module):
    """`num_layers` stacked cells run over the sequence (ConvLSTM / ConvGRU of the reference)."""

    def __init__(self, kind, in_channels, hidden_channels, kernel_size, num_layers):
        super().__init__()
        self.cell_list = nn.ModuleList(
            [_Cell(kind, in_channels if l == 0 else hidden_channels, hidden_channels, kernel_size) for l in range(num_layers)])

    def forward(self, x, split_input=True):
        for cell in self.cell_list:
            x = cell.run(x, split_input)
        return x


class BiConvRecurrent(nn.Module):
    """ConvBLSTM (kind="lstm") / ConvBGRU (kind="gru") with the reference's state_dict layout.
    forward(x): B x T x C x H x W -> B x T x hidden_channels x H x W."""

    def __init__(self, kind, in_channels=512, hidden_channels=512, kernel_size=(3, 3), num_layers=2):
        super().__init__()
        assert kind in ("lstm", "gru")
        self.forward_net = _Direction(kind, in_channels, hidden_channels // 2, kernel_size, num_layers)
        self.reverse_net = _Direction(kind, in_channels, hidden_channels // 2, kernel_size, num_layers)
        # True: input half of the gate convolutions batched over the windows (fast); False: the reference's joint
        # convolution per step (same rounding as the reference, for the golden comparison)
        self.split_input = True

    def forward(self, x):
        fwd = self.forward_net(x, self.split_input)
        rev = self.reverse_net(x.flip(1), self.split_input).flip(1)
        return torch.cat([fwd, rev], dim=2)
