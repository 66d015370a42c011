"""StyleGAN2 background decoder (`Discriminator.bg_decoder`) on the sm_100a kernels.

Mirror of the part of the reference's training/networks_stylegan2.py that LayoutDETR instantiates
(Decoder :972, DecoderMappingNetwork :903, SynthesisNetwork :465, SynthesisBlock :361,
SynthesisLayer :272, ToRGBLayer :336, FullyConnectedLayer :92, modulated_conv2d :30) for the
configuration networks_detr.py:261 builds: architecture 'skip', use_noise=False, num_fp16_res=0,
conv_clamp=None, fused_modconv_default=False (=> activations are scaled by the styles before and by the
demodulation coefficients after a convolution that shares its weights across the batch).

Parameter / buffer names equal the reference's (`mapping.fc{i}.weight`, `synthesis.b{res}.conv0.affine.weight`,
`synthesis.b{res}.torgb.weight`, `resample_filter`, `w_avg`, ...).  Activations are channels-last bf16
`[B*H*W, C]`; every convolution is a tcgen05 GEMM (3x3 via im2col, stride-2 transposed 3x3 via
GEMM + col2im), the FIR resampling is the upfirdn2d kernel and demodulation + bias + leaky-ReLU is one
fused elementwise kernel.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import functional as Fn
from .. import kernels as K

SQRT2 = float(np.sqrt(2.0))


def setup_filter(f):
    """Normalised separable->2-D low-pass filter (reference torch_utils/ops/upfirdn2d.py:70-117, defaults)."""
    f = torch.as_tensor(f, dtype=torch.float32)
    if f.ndim == 1:
        f = f.ger(f)
    return f / f.sum()


class FullyConnectedLayer(nn.Module):
    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.activation = activation
        self.weight = nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = float(lr_multiplier / np.sqrt(in_features))
        self.bias_gain = float(lr_multiplier)

    def forward(self, x, out_f32=False):
        """x bf16 [M, in] -> bf16 (or fp32) [M, out]."""
        if self.activation == 'linear':
            return Fn.scaled_linear(x, self.weight, self.bias, self.weight_gain, self.bias_gain, K.ACT_NONE, 1.0, out_f32)
        assert self.activation == 'lrelu'
        return Fn.scaled_linear(x, self.weight, self.bias, self.weight_gain, self.bias_gain, K.ACT_LRELU, SQRT2, False)

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


def _demod_coefs(weight, styles):
    """dcoefs[b, o] = rsqrt(sum_{i,kh,kw} (W[o,i,kh,kw] * styles[b,i])^2 + 1e-8) (networks_stylegan2.py:58-63),
    evaluated as (styles^2) @ (sum_k W^2)^T in fp32 — a [B,Cin]x[Cin,Cout] product, negligible next to the conv."""
    wsq = weight.float().square().sum(dim=[2, 3])                 # [Cout, Cin]
    return (styles.square() @ wsq.t() + 1e-8).rsqrt()


class SynthesisLayer(nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=False,
                 activation='lrelu', resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False):
        super().__init__()
        assert not use_noise and conv_clamp is None and activation == 'lrelu' and kernel_size == 3
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.w_dim = w_dim
        self.resolution = resolution
        self.up = up
        self.use_noise = use_noise
        self.activation = activation
        self.conv_clamp = conv_clamp
        self.register_buffer('resample_filter', setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.act_gain = SQRT2
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w, B, gain=1):
        """x bf16 [B*Hin*Hin, Cin], w bf16 [B, w_dim] -> bf16 [B*res*res, Cout]."""
        Hin = self.resolution // self.up
        styles = self.affine(w, out_f32=True)                                           # [B, Cin] fp32
        dcoefs = _demod_coefs(self.weight, styles)                                      # [B, Cout] fp32
        xs = Fn.scale_channels(x, styles, B, Hin * Hin, self.in_channels)
        if self.up == 1:
            y = Fn.conv2d(xs, self.weight, None, None, None, B, Hin, Hin, 1, self.padding, K.ACT_NONE)
        else:
            y = Fn.conv_transpose_up2(xs, self.weight, B, Hin, Hin)                     # [B*(2H+1)^2, Cout]
            y = Fn.upfirdn_nhwc(y, self.resample_filter, B, 2 * Hin + 1, 2 * Hin + 1, pad=(1, 1, 1, 1), gain=float(self.up ** 2))
        R = self.resolution
        return Fn.demod_bias_act(y, dcoefs, self.bias, B, R * R, self.out_channels, K.ACT_LRELU, self.act_gain * gain)

    def extra_repr(self):
        return f'in_channels={self.in_channels:d}, out_channels={self.out_channels:d}, w_dim={self.w_dim:d}, resolution={self.resolution:d}, up={self.up}'


class ToRGBLayer(nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        assert kernel_size == 1 and conv_clamp is None
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.w_dim = w_dim
        self.conv_clamp = conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = float(1 / np.sqrt(in_channels * (kernel_size ** 2)))

    def forward(self, x, w, B, pixels):
        styles = self.affine(w, out_f32=True) * self.weight_gain
        xs = Fn.scale_channels(x, styles, B, pixels, self.in_channels)
        return Fn.linear_f32(xs, self.weight, self.bias)   # [B*pixels, 3] fp32


class SynthesisBlock(nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip',
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, use_fp16=False, fp16_channels_last=False,
                 fused_modconv_default=False, **layer_kwargs):
        super().__init__()
        assert architecture == 'skip' and not use_fp16
        self.in_channels = in_channels
        self.w_dim = w_dim
        self.resolution = resolution
        self.img_channels = img_channels
        self.is_last = is_last
        self.architecture = architecture
        self.register_buffer('resample_filter', setup_filter(resample_filter))
        self.num_conv = 0
        self.num_torgb = 0
        if in_channels == 0:
            self.const = nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        if in_channels != 0:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, up=2,
                                        resample_filter=resample_filter, conv_clamp=conv_clamp, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution,
                                    conv_clamp=conv_clamp, **layer_kwargs)
        self.num_conv += 1
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp)
        self.num_torgb += 1
        self.out_channels = out_channels

    def forward(self, x, img, ws, B):
        """x bf16 [B*(res/2)^2, Cin] or None; img fp32 [B*(res/2)^2, 3] or None; ws: list of bf16 [B, w_dim]."""
        R = self.resolution
        w_iter = iter(ws)
        if self.in_channels == 0:
            c = Fn.to_bf16_padded(self.const.permute(1, 2, 0).reshape(R * R, self.out_channels))
            x = c.unsqueeze(0).expand(B, -1, -1).reshape(B * R * R, self.out_channels)
            x = self.conv1(x, next(w_iter), B)
        else:
            x = self.conv0(x, next(w_iter), B)
            x = self.conv1(x, next(w_iter), B)
        if img is not None:
            # upsample2d(img, f): up=2, pad (2,1,2,1), gain 4 (torch_utils/ops/upfirdn2d.py:314-348)
            img = Fn.upfirdn_nhwc(img, self.resample_filter, B, R // 2, R // 2, up=2, pad=(2, 1, 2, 1), gain=4.0)
        y = self.torgb(x, next(w_iter), B, R * R)
        img = img + y if img is not None else y
        return x, img


class SynthesisNetwork(nn.Module):
    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=4, **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.w_dim = w_dim
        self.img_resolution = img_resolution
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.img_channels = img_channels
        self.num_fp16_res = num_fp16_res
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        channels_dict = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        self.num_ws = 0
        for res in self.block_resolutions:
            in_channels = channels_dict[res // 2] if res > 4 else 0
            out_channels = channels_dict[res]
            is_last = (res == self.img_resolution)
            block = SynthesisBlock(in_channels, out_channels, w_dim=w_dim, resolution=res, img_channels=img_channels,
                                   is_last=is_last, use_fp16=False, **block_kwargs)
            self.num_ws += block.num_conv
            if is_last:
                self.num_ws += block.num_torgb
            setattr(self, f'b{res}', block)

    def forward(self, w, B):
        """w bf16 [B, w_dim] (the mapping output; the reference broadcasts it to all num_ws layers)."""
        x = img = None
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            x, img = block(x, img, [w] * (block.num_conv + block.num_torgb), B)
        R = self.img_resolution
        return img.view(B, R, R, self.img_channels).permute(0, 3, 1, 2).contiguous()


class DecoderMappingNetwork(nn.Module):
    def __init__(self, z_dim, w_dim, num_ws, num_layers=8, layer_features=None, activation='lrelu', lr_multiplier=0.01,
                 w_avg_beta=0.998):
        super().__init__()
        self.z_dim = z_dim
        self.w_dim = w_dim
        self.num_ws = num_ws
        self.num_layers = num_layers
        self.w_avg_beta = w_avg_beta
        if layer_features is None:
            layer_features = w_dim
        features_list = [z_dim] + [layer_features] * (num_layers - 1) + [w_dim]
        for idx in range(num_layers):
            setattr(self, f'fc{idx}', FullyConnectedLayer(features_list[idx], features_list[idx + 1], activation=activation,
                                                          lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        assert truncation_psi == 1, "truncation is not on the LayoutDETR path"
        x = z
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if update_emas and self.w_avg_beta is not None:
            self.w_avg.copy_(x.detach().float().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        return x


class Decoder(nn.Module):
    def __init__(self, z_dim, w_dim, img_resolution, img_channels, use_noise, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim = z_dim
        self.w_dim = w_dim
        self.img_resolution = img_resolution
        self.img_channels = img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels,
                                          use_noise=use_noise, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = DecoderMappingNetwork(z_dim=z_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)

    def forward(self, z, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        """z: bf16/fp32 [B, z_dim] -> fp32 image [B, img_channels, R, R]."""
        B = z.shape[0]
        z = Fn.to_bf16_padded(z) if z.dtype != torch.bfloat16 else z
        w = self.mapping(z, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(w, B)
