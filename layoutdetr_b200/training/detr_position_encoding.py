"""2-D sine position embedding (mirror of reference training/detr_position_encoding.py:38-58).

The embedding depends only on the padding mask.  For the all-False masks the LayoutDETR path
produces (equal-size backgrounds, detr_util/misc.py:320-342) it is a constant of (h, w), so it is
computed once per feature-map size on the host with the reference's formula and cached on the device
(SURVEY.md §8a-5)."""
import math

import torch
import torch.nn as nn


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.scale = 2 * math.pi if scale is None else scale
        self._cache = {}

    def from_mask(self, mask):
        """mask: bool [B, h, w] (True = padding) -> fp32 [B, h*w, 2*num_pos_feats] (token-major)."""
        not_mask = ~mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            eps = 1e-6
            y_embed = y_embed / (y_embed[:, -1:, :] + eps) * self.scale
            x_embed = x_embed / (x_embed[:, :, -1:] + eps) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).flatten(3)
        pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).flatten(3)
        pos = torch.cat((pos_y, pos_x), dim=3)           # [B, h, w, 2F]  (channels-last == token-major)
        return pos.flatten(1, 2)

    def for_size(self, h, w, device):
        """Cached [h*w, 2F] fp32 embedding for an unpadded h x w feature map."""
        key = (h, w, str(device))
        if key not in self._cache:
            mask = torch.zeros((1, h, w), dtype=torch.bool)
            self._cache[key] = self.from_mask(mask)[0].to(device).contiguous()
        return self._cache[key]

    def forward(self, h, w, device):
        return self.for_size(h, w, device)
