"""LayoutNet — the layout feature extractor behind the layout-FID metric (mirror of the reference's
training/networks_layoutnet.py:16-86; same constructor, `state_dict` keys and `extract_features` / `forward`
signatures, so the released LayoutNet checkpoints load unchanged), executed by the sm_100a kernels.

Activations are bf16 `[B*L, 256]` batch-major rows like everywhere else in this package; the returned features and
head outputs are fp32.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as Fn
from .. import kernels as K
from .detr_transformer import TransformerEncoderStack
from .util import TransformerWithToken_layoutganpp


class LayoutNet(nn.Module):
    def __init__(self, num_label):
        super().__init__()
        d_model, nhead, num_layers, max_bbox = 256, 4, 4, 50
        self.d_model = d_model
        # encoder
        self.emb_label = nn.Embedding(num_label, d_model)
        self.fc_bbox = nn.Linear(4, d_model)
        self.enc_fc_in = nn.Linear(d_model * 2, d_model)
        self.enc_transformer = TransformerWithToken_layoutganpp(d_model=d_model, dim_feedforward=d_model // 2, nhead=nhead,
                                                                num_layers=num_layers)
        self.fc_out_disc = nn.Linear(d_model, 1)
        # decoder
        self.pos_token = nn.Parameter(torch.rand(max_bbox, 1, d_model))
        self.dec_fc_in = nn.Linear(d_model * 2, d_model)
        te = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=d_model // 2)
        self.dec_transformer = nn.TransformerEncoder(te, num_layers=num_layers, enable_nested_tensor=False)
        self.fc_out_cls = nn.Linear(d_model, num_label)
        self.fc_out_bbox = nn.Linear(d_model, 4)

    @staticmethod
    def _remap(label, label_idx_replace, label_idx_replace_2):
        """Label-space remapping of reference :48-60, out of place (the reference edits the caller's tensor in place)."""
        if label_idx_replace:
            lut = torch.tensor([2, 2, 2, 2, 2, 4, 7, 3], device=label.device)          # :49-52 applied in order
            return torch.where(label < 8, lut[label.clamp(0, 7)], label)
        if label_idx_replace_2:
            lut = torch.tensor([3, 2, 4, 3, 2], device=label.device)                    # :54-59 applied in order
            return torch.where(label < 5, lut[label.clamp(0, 4)], label)
        return label

    def _token_features(self, bbox, label, padding_mask):
        if bbox.device.type != "cuda":
            raise RuntimeError("layoutdetr_b200.LayoutNet runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, N, _ = bbox.shape
        b = Fn.linear(Fn.to_bf16_padded(bbox.reshape(B * N, 4).float()), self.fc_bbox.weight, self.fc_bbox.bias)
        l = F.embedding(label, self.emb_label.weight).to(torch.bfloat16).reshape(B * N, -1)
        x = Fn.linear(torch.cat([b, l], dim=-1), self.enc_fc_in.weight, self.enc_fc_in.bias, act=K.ACT_RELU)
        x = self.enc_transformer(x, B, N, padding_mask)                                  # [B*(N+1), 256]
        return x.view(B, N + 1, self.d_model)[:, 0, :]

    def extract_features(self, bbox, label, padding_mask, label_idx_replace=False, label_idx_replace_2=False):
        """[B, N, 4] boxes, [B, N] labels, [B, N] bool (True = padding) -> fp32 [B, 256] (the token output, reference :46-65)."""
        label = self._remap(label, label_idx_replace, label_idx_replace_2)
        return Fn.to_f32(self._token_features(bbox, label, padding_mask).contiguous())

    def forward(self, bbox, label, padding_mask):
        """-> (logit_disc [B], logit_cls [M, L], bbox_pred [M, 4]) over the M valid slots (reference :67-86)."""
        from .networks_detr import valid_index
        B, N, _ = bbox.shape
        H = self.d_model
        x0 = self._token_features(bbox, label, padding_mask).contiguous()
        logit_disc = Fn.linear_f32(x0, self.fc_out_disc.weight, self.fc_out_disc.bias).squeeze(-1)
        t = Fn.to_bf16_padded(self.pos_token[:N].reshape(N, H))
        x = torch.cat([x0.unsqueeze(1).expand(-1, N, -1), t.unsqueeze(0).expand(B, -1, -1)], dim=-1).reshape(B * N, 2 * H)
        x = Fn.linear(x, self.dec_fc_in.weight, self.dec_fc_in.bias, act=K.ACT_RELU)
        x = TransformerEncoderStack.run(self.dec_transformer, x, B, N, padding_mask)
        valid, _ = valid_index(padding_mask)
        x = x.index_select(0, valid)
        logit_cls = Fn.linear_f32(x, self.fc_out_cls.weight, self.fc_out_cls.bias)
        bbox_pred = Fn.linear_f32(x, self.fc_out_bbox.weight, self.fc_out_bbox.bias).sigmoid()
        return logit_disc, logit_cls, bbox_pred
