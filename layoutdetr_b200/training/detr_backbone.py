"""ResNet-50 + FrozenBatchNorm2d backbone on the sm_100a conv path (im2col + tcgen05 GEMM with the
frozen-BN scale/shift, residual add and ReLU fused into the GEMM epilogue).

Mirror of the reference's training/detr_backbone.py (Backbone :98, BackboneBase :68, Joiner :117,
FrozenBatchNorm2d :29) and of torchvision.models.resnet50 (v1.5: stride on the 3x3 conv), with the
same state_dict keys (`backbone.0.body.layer{1..4}.{i}.conv{1,2,3}.weight`, `.bn{1,2,3}.*`,
`.downsample.{0,1}.*`).  Activations are channels-last bf16 `[B*H*W, C]` end to end.
"""
import torch
import torch.nn as nn

from .. import engine as E
from .. import functional as Fn
from .. import kernels as K
from .detr_position_encoding import PositionEmbeddingSine


class FrozenBatchNorm2d(nn.Module):
    """Fixed statistics / affine: y = x * scale + shift with scale = w * rsqrt(rv + 1e-5)."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, *args):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, *args)

    @staticmethod
    def _fold(weight, bias, running_mean, running_var):
        scale = weight.float() * (running_var.float() + 1e-5).rsqrt()
        shift = bias.float() - running_mean.float() * scale
        return torch.stack([scale, shift]).contiguous()

    def folded(self):
        ss = E.derived((self.weight, self.bias, self.running_mean, self.running_var), "fbn", FrozenBatchNorm2d._fold)
        return ss[0], ss[1]


def _conv(cin, cout, k, stride=1, pad=0):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=pad, bias=False)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 1)
        self.bn1 = FrozenBatchNorm2d(planes)
        self.conv2 = _conv(planes, planes, 3, stride, 1)
        self.bn2 = FrozenBatchNorm2d(planes)
        self.conv3 = _conv(planes, planes * 4, 1)
        self.bn3 = FrozenBatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x, B, H, W):
        s1, b1 = self.bn1.folded()
        s2, b2 = self.bn2.folded()
        s3, b3 = self.bn3.folded()
        out = Fn.conv2d(x, self.conv1.weight, s1, b1, None, B, H, W, 1, 0, K.ACT_RELU)
        out = Fn.conv2d(out, self.conv2.weight, s2, b2, None, B, H, W, self.stride, 1, K.ACT_RELU)
        Ho, Wo = K.conv_out_size(H, 3, self.stride, 1), K.conv_out_size(W, 3, self.stride, 1)
        if self.downsample is not None:
            sd, bd = self.downsample[1].folded()
            identity = Fn.conv2d(x, self.downsample[0].weight, sd, bd, None, B, H, W, self.stride, 0, K.ACT_NONE)
        else:
            identity = x
        out = Fn.conv2d(out, self.conv3.weight, s3, b3, identity, B, Ho, Wo, 1, 0, K.ACT_RELU)
        return out, Ho, Wo


class ResNet50Body(nn.Module):
    """conv1 .. layer4 of torchvision's resnet50 (what IntermediateLayerGetter keeps, detr_backbone.py:79)."""

    def __init__(self):
        super().__init__()
        self.conv1 = _conv(3, 64, 7, 2, 3)
        self.bn1 = FrozenBatchNorm2d(64)
        self.inplanes = 64
        self.layer1 = self._make_layer(64, 3, 1)
        self.layer2 = self._make_layer(128, 4, 2)
        self.layer3 = self._make_layer(256, 6, 2)
        self.layer4 = self._make_layer(512, 3, 2)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * 4:
            downsample = nn.Sequential(_conv(self.inplanes, planes * 4, 1, stride), FrozenBatchNorm2d(planes * 4))
        layers = [Bottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * 4
        for _ in range(1, blocks):
            layers.append(Bottleneck(self.inplanes, planes, 1, None))
        return nn.Sequential(*layers)

    def forward(self, image):
        """image: fp32 NCHW [B, 3, H, W] -> (bf16 [B*h*w, 2048], h, w)."""
        B, C, H, W = image.shape
        x = K.nchw_to_nhwc(image, torch.bfloat16)
        s, b = self.bn1.folded()
        x = Fn.conv2d(x, self.conv1.weight, s, b, None, B, H, W, 2, 3, K.ACT_RELU)
        H, W = K.conv_out_size(H, 7, 2, 3), K.conv_out_size(W, 7, 2, 3)
        x = Fn.maxpool3s2(x, B, H, W)
        H, W = K.conv_out_size(H, 3, 2, 1), K.conv_out_size(W, 3, 2, 1)
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                x, H, W = blk(x, B, H, W)
        return x, H, W


class Backbone(nn.Module):
    """`backbone.0` of the reference Joiner: holds `.body`."""

    def __init__(self, name="resnet50", train_backbone=True, return_interm_layers=None, dilation=False):
        super().__init__()
        assert name == "resnet50" and not dilation
        self.body = ResNet50Body()
        for pname, p in self.body.named_parameters():
            if not train_backbone or ("layer2" not in pname and "layer3" not in pname and "layer4" not in pname):
                p.requires_grad_(False)
        self.num_channels = 2048

    def forward(self, image):
        return self.body(image)


class Joiner(nn.Sequential):
    def __init__(self, backbone, position_embedding):
        super().__init__(backbone, position_embedding)
        self.num_channels = backbone.num_channels

    def forward(self, image):
        """-> (features bf16 [B*h*w, 2048], pos fp32 [h*w, 256], h, w).  The padding mask of the reference's
        NestedTensor is all-False for the equal-size batches this path sees, so it is not materialised."""
        feat, h, w = self[0](image)
        pos = self[1].for_size(h, w, feat.device)
        return feat, pos, h, w


def build_backbone():
    backbone = Backbone("resnet50", train_backbone=True, return_interm_layers=None, dilation=False)
    return Joiner(backbone, PositionEmbeddingSine(num_pos_feats=128, normalize=True))
