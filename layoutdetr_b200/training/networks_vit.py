"""ViT-B/16 image encoder on the sm_100a kernels — mirror of the reference's training/networks_vit.py `VisionTransformer`
(:139-225; BLIP's ViT re-based on `nn.TransformerEncoder`): same constructor signature, `forward(x, mask)` contract and
state_dict keys (`patch_embed.proj.*`, `cls_token`, `pos_embed`, `token_mask`, `transformer.layers.{i}.*`, `transformer.norm.*`,
`norm.*`), so checkpoints of the reference class load by name.

    patch_embed   Conv2d(3, 768, 16, stride 16)          -> one GEMM over the non-overlapping 16 x 16 x 3 patches
    mask_embed    MaxPool2d(16, 16) of the [B, 1, H, W] mask -> per-patch validity -> key-padding mask (+ the token's False)
    tokens        [cls_token ; patches] + pos_embed       -> bf16 rows [B * (N + 1), 768]
    transformer   12 x post-norm nn.TransformerEncoderLayer(768, 12 heads, 3072, GELU, LayerNorm eps 1e-5) + LayerNorm(1e-6)
    norm          LayerNorm(1e-6) of the class token      -> [B, 768]                          (reference forward, :203-221)

`ViTBackbone` (BASELINE configs[3]: 1024^2 background) exposes the patch tokens through the backbone
interface `Generator` / `Discriminator` use — (features [B*h*w, 768], sine position embedding, h, w) with `num_channels = 768`
feeding `input_proj` — as an alternative to ResNet-50; the reference defines the class but never wires it.
The image encoder has 257 (256^2) or 4097 (1024^2) keys per query: above the fused attention kernel's 256-key limit, so attention
runs as batched tcgen05 GEMMs with the masked-softmax kernel in between (functional.AttentionFn, unfused branch).
"""
from functools import partial

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import kernels as K
from .detr_position_encoding import PositionEmbeddingSine
from .detr_transformer import TransformerEncoderStack


class PatchEmbed(nn.Module):
    """timm's PatchEmbed as used by the reference (:182): holder of `proj` = Conv2d(in, embed, patch, stride=patch)."""

    def __init__(self, img_size=(224, 224), patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = tuple(img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size[0] // patch_size, img_size[1] // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, image):
        """fp32 NCHW [B, 3, H, W] -> bf16 [B * N, embed_dim] (row = b * N + patch, patches in raster order)."""
        B, C, H, W = image.shape
        assert (H, W) == self.img_size, "input %dx%d does not match the model's %dx%d" % (H, W, *self.img_size)
        rows = K.nchw_to_nhwc(image.float(), torch.bfloat16)
        p = self.patch_size[0]
        return Fn.conv2d(rows, self.proj.weight, None, self.proj.bias, None, B, H, W, p, 0, K.ACT_NONE)


class MaskEmbed(nn.Module):
    """Per-patch validity of a [B, 1, H, W] mask: any non-zero pixel keeps the patch (reference :27-46, MaxPool2d(16, 16))."""

    def __init__(self, img_height=224, img_width=224, patch_size=16, flatten=True):
        super().__init__()
        self.img_size = [img_height, img_width]
        self.patch_size = [patch_size, patch_size]
        self.grid_size = [img_height // patch_size, img_width // patch_size]
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.MaxPool2d(kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1] and C == 1
        x = self.proj(x.float())                 # [B, 1, h, w]: a reduction over 256 pixels per patch of a 1-channel mask (glue)
        if self.flatten:
            x = x.flatten(2).squeeze(1)
        return x.to(torch.bool)


class VisionTransformer(nn.Module):
    def __init__(self, img_height=224, img_width=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=True, qk_scale=None, representation_size=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_layer=None,
                 use_grad_checkpointing=False, ckpt_layer=0):
        super().__init__()
        self.num_features = self.embed_dim = embed_dim
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.patch_embed = PatchEmbed(img_size=(img_height, img_width), patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        self.mask_embed = MaskEmbed(img_height=img_height, img_width=img_width, patch_size=patch_size)
        self.register_buffer('token_mask', torch.zeros(1, 1, dtype=torch.bool))
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        te = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=num_heads, dim_feedforward=int(embed_dim * mlp_ratio),
                                        dropout=drop_rate, activation='gelu')
        self.transformer = nn.TransformerEncoder(te, num_layers=depth, norm=norm_layer(embed_dim), enable_nested_tensor=False)
        self.norm = norm_layer(embed_dim)
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def tokens(self, x, mask=None):
        """-> (bf16 [B * (N + 1), C] output of the encoder stack incl. its final LayerNorm, B, N + 1); row b * (N + 1) is the class
        token.  mask: [B, 1, H, W], non-zero = image content (None: everything valid)."""
        if not x.is_cuda:
            raise RuntimeError("layoutdetr_b200.VisionTransformer runs on CUDA (sm_100a) only; there is no CPU fallback")
        B = x.shape[0]
        C = self.embed_dim
        patches = self.patch_embed(x)                                                   # [B * N, C] bf16
        N = patches.shape[0] // B
        if mask is None:
            key_pad = torch.zeros((B, N + 1), dtype=torch.bool, device=x.device)
        else:
            key_pad = torch.cat([self.token_mask.expand(B, -1), ~self.mask_embed(mask)], dim=1)
        cls = Fn.to_bf16_padded(self.cls_token.view(1, C)).view(1, 1, C).expand(B, 1, C)
        tok = torch.cat([cls, patches.view(B, N, C)], dim=1).reshape(B * (N + 1), C)
        tok = Fn.add_bcast(tok, self.pos_embed[0, :N + 1].contiguous())               # + pos_embed, broadcast over the batch
        tok = Fn.dropout(tok, self.pos_drop.p if self.training else 0.0)
        out = TransformerEncoderStack.run(self.transformer, tok, B, N + 1, key_pad)
        return out, B, N + 1

    def forward(self, x, mask=None):
        """-> fp32 [B, C]: LayerNorm of the class token (reference :203-221)."""
        out, B, L = self.tokens(x, mask)
        cls = out.view(B, L, self.embed_dim)[:, 0, :].contiguous()
        y = Fn.layernorm(cls, self.norm.weight, self.norm.bias, self.norm.eps)
        return Fn.to_f32(y)


class ViTBackbone(nn.Module):
    """ViT-B/16 behind the backbone interface of networks_detr (`Joiner.forward`): image -> (patch tokens bf16 [B*h*w, 768],
    sine position embedding fp32 [h*w, 256], h, w).  The token grid is h = H / 16, w = W / 16 (S = 4096 at 1024^2)."""

    def __init__(self, img_height=256, img_width=256, hidden_dim=256):
        super().__init__()
        self.body = VisionTransformer(img_height=img_height, img_width=img_width)
        self.position = PositionEmbeddingSine(num_pos_feats=hidden_dim // 2, normalize=True)
        self.num_channels = self.body.embed_dim

    def forward(self, image):
        out, B, L = self.body.tokens(image, None)
        h, w = self.body.patch_embed.grid_size
        feat = out.view(B, L, -1)[:, 1:, :].reshape(B * (L - 1), -1)
        return feat, self.position.for_size(h, w, image.device), h, w


def build_vit_backbone(img_height=256, img_width=256, hidden_dim=256):
    return ViTBackbone(img_height, img_width, hidden_dim)
