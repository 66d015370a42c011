"""Mirror of reference training/util.py:13-43 (TransformerWithToken_layoutganpp): learned token +
torch-style post-norm encoder stack, executed by the sm_100a kernels."""
import torch
import torch.nn as nn

from .. import functional as Fn
from .detr_transformer import TransformerEncoderStack


class TransformerWithToken_layoutganpp(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, num_layers):
        super().__init__()
        self.token = nn.Parameter(torch.randn(1, 1, d_model))
        self.register_buffer("token_mask", torch.zeros(1, 1, dtype=torch.bool))
        self.core = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=dim_feedforward),
            num_layers=num_layers)
        self.d_model = d_model

    def forward(self, x, B, L, src_key_padding_mask):
        """x bf16 [B*L, d] batch-major -> bf16 [B*(L+1), d]; row b*(L+1) is the token output."""
        d = self.d_model
        tok = Fn.to_bf16_padded(self.token.view(1, d)).view(1, 1, d).expand(B, 1, d)
        x = torch.cat([tok, x.view(B, L, d)], dim=1).reshape(B * (L + 1), d)
        mask = torch.cat([self.token_mask.expand(B, -1), src_key_padding_mask], dim=1)
        return TransformerEncoderStack.run(self.core, x, B, L + 1, mask)
