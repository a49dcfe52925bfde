"""Drop-in for the four DGL heads of reference models/fusion_modules.py (same class names,
constructor signatures, parameter names and init order; `forward(x, y) -> (x_out, y_out, out)`).
The forwards are autograd Functions over the C-ABI linear / gated kernels, so the reference's own
two-backward `train_epoch` works on them; the fused training step (step.DGLStep) bypasses them
and calls gdl_dgl_head_linear directly.
"""
import torch.nn as nn

from . import autograd as _ag


class SumFusion_DGL(nn.Module):
    """reference fusion_modules.py:16-30."""

    def __init__(self, input_dim=512, output_dim=100):
        super().__init__()
        self.fc_x = nn.Linear(input_dim, output_dim)
        self.fc_y = nn.Linear(input_dim, output_dim)

    def forward(self, x, y):
        outx = _ag.linear(x, self.fc_x.weight, self.fc_x.bias)
        outy = _ag.linear(y, self.fc_y.weight, self.fc_y.bias)
        output = _ag.linear(x.detach(), self.fc_x.weight, self.fc_x.bias) + \
            _ag.linear(y.detach(), self.fc_y.weight, self.fc_y.bias)
        return outx, outy, output


class ConcatFusion_DGL(nn.Module):
    """reference fusion_modules.py:45-59.  fc_auxi exists (state_dict, init RNG) but is never used."""

    def __init__(self, input_dim=512 * 2, output_dim=100):
        super().__init__()
        self.fc_out = nn.Linear(input_dim, output_dim)
        self.fc_auxi = nn.Linear(input_dim, output_dim)

    def forward(self, x, y):
        D = x.shape[1]
        W, b = self.fc_out.weight, self.fc_out.bias
        # cat(x,0) W^T = x Wx^T and cat(0,y) W^T = y Wy^T with W = [Wx | Wy]
        x_part = _ag.linear(x, W, None, col0=0, cols=D)
        y_part = _ag.linear(y, W, None, col0=D, cols=y.shape[1])
        xd = _ag.linear(x.detach(), W, None, col0=0, cols=D)
        yd = _ag.linear(y.detach(), W, None, col0=D, cols=y.shape[1])
        output = _ag.add_bias(xd + yd, b)
        return _ag.add_bias(x_part, b), _ag.add_bias(y_part, b), output


class FiLM_DGL(nn.Module):
    """reference fusion_modules.py:126-178 (an outer-product head despite its name)."""

    def __init__(self, input_dim=512, dim=512, output_dim=100, x_film=True):
        super().__init__()
        self.fc = nn.Linear(dim * dim, dim)
        self.fc_out = nn.Linear(dim, output_dim)
        self.x_film = x_film

    def forward(self, x, y):
        out = _ag.linear(_ag.outer_linear(x.detach(), y.detach(), self.fc.weight, self.fc.bias),
                         self.fc_out.weight, self.fc_out.bias)
        z_x = _ag.linear(_ag.outer_linear(x, x, self.fc.weight, self.fc.bias),
                         self.fc_out.weight, self.fc_out.bias)
        z_y = _ag.linear(_ag.outer_linear(y, y, self.fc.weight, self.fc.bias),
                         self.fc_out.weight, self.fc_out.bias)
        return z_x, z_y, out


class GatedFusion_DGL(nn.Module):
    """reference fusion_modules.py:213-250."""

    def __init__(self, input_dim=512, dim=512, output_dim=100, x_gate=True):
        super().__init__()
        self.fc_x = nn.Linear(input_dim, dim)
        self.fc_y = nn.Linear(input_dim, dim)
        self.fc_out = nn.Linear(dim, output_dim)
        self.x_gate = x_gate
        self.sigmoid = nn.Sigmoid()

    def forward(self, x, y):
        out_x = _ag.linear(x, self.fc_x.weight, self.fc_x.bias)
        out_y = _ag.linear(y, self.fc_y.weight, self.fc_y.bias)
        if self.x_gate:
            m = _ag.gate(out_x.detach(), out_y.detach())
        else:
            m = _ag.gate(out_y.detach(), out_x.detach())
        output = _ag.linear(m, self.fc_out.weight, self.fc_out.bias)
        ox = _ag.linear(_ag.gate(out_x, out_x), self.fc_out.weight, self.fc_out.bias)
        oy = _ag.linear(_ag.gate(out_y, out_y), self.fc_out.weight, self.fc_out.bias)
        return ox, oy, output


# BASELINE.json / the reference README call these "*_AUXI"; no such classes exist upstream
# (SURVEY.md "naming trap") — harmless aliases.
ConcatFusion_AUXI = ConcatFusion_DGL
SumFusion_AUXI = SumFusion_DGL
FiLM_AUXI = FiLM_DGL
GatedFusion_AUXI = GatedFusion_DGL
