"""GPU parity of the StyleGAN2 background decoder and the drop-in ops vs the CPU oracle / reference goldens."""
import numpy as np
import pytest
import torch

from helpers import golden

pytestmark = pytest.mark.gpu


def _decoder():
    from layoutdetr_b200.synthetic import synth_tensor
    from layoutdetr_b200.training.networks_stylegan2 import Decoder
    dec = Decoder(z_dim=256, w_dim=512, channel_max=512, channel_base=8192, img_channels=3, img_resolution=256,
                  use_noise=False, num_fp16_res=0, conv_clamp=None, fused_modconv_default=False).eval()
    sd = {k: synth_tensor("bg_decoder." + k, v.shape).to(v.dtype) for k, v in dec.state_dict().items()}
    dec.load_state_dict(sd)
    return dec


def test_bg_decoder_blocks_vs_oracle():
    """Walk the synthesis blocks with identical inputs on both sides and report the per-block error."""
    from oracle import layoutdetr_oracle as O
    dec = _decoder()
    sd = {"bg_decoder." + k: v.detach().clone() for k, v in dec.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    B = 2
    z = torch.randn((B, 256), generator=g)
    with torch.no_grad():
        # oracle, keeping intermediates
        x = z
        for i in range(8):
            x = O.sg2_fc(sd, "bg_decoder.mapping.fc%d" % i, x, "lrelu", 0.01)
        w_ref = x
        dec = dec.cuda()
        from layoutdetr_b200 import functional as Fn
        w = dec.mapping(Fn.to_bf16_padded(z.cuda()))
        e_w = float((w.float().cpu() - w_ref).abs().max() / w_ref.abs().max())
        print("mapping w rel err %.3e (|w|max %.3f)" % (e_w, float(w_ref.abs().max())))
        assert e_w < 3e-2
        img_ref = O.sg2_decoder(sd, "bg_decoder", z)
        img = dec(z.cuda()).float().cpu()
        err = float((img - img_ref).abs().max() / img_ref.std())
        rms = float((img - img_ref).pow(2).mean().sqrt() / img_ref.std())
        print("bg_decoder image: max err / std = %.3e, rms err / std = %.3e" % (err, rms))
        # per-block comparison with the ORACLE's w fed to both sides (isolates the synthesis path)
        wq = w_ref.to(torch.bfloat16)
        feat = None
        x_dev = None
        res = 4
        while res <= 256:
            bp = "bg_decoder.synthesis.b%d" % res
            blk = getattr(dec.synthesis, "b%d" % res)
            f = sd[bp + ".resample_filter"]
            if res == 4:
                feat = sd[bp + ".const"].unsqueeze(0).repeat(B, 1, 1, 1)
            else:
                feat = O.bias_act(O.sg2_modconv(sd, bp + ".conv0", feat, wq.float(), 2, f), sd[bp + ".conv0.bias"], act="lrelu")
            feat = O.bias_act(O.sg2_modconv(sd, bp + ".conv1", feat, wq.float(), 1, f), sd[bp + ".conv1.bias"], act="lrelu")
            if res == 4:
                c = Fn.to_bf16_padded(blk.const.permute(1, 2, 0).reshape(16, blk.out_channels))
                x_dev = c.unsqueeze(0).expand(B, -1, -1).reshape(B * 16, blk.out_channels)
                x_dev = blk.conv1(x_dev, wq.cuda(), B)
            else:
                x_dev = blk.conv0(x_dev, wq.cuda(), B)
                x_dev = blk.conv1(x_dev, wq.cuda(), B)
            ours = x_dev.float().cpu().view(B, res, res, -1).permute(0, 3, 1, 2)
            e = float((ours - feat).abs().max() / feat.std())
            r = float((ours - feat).pow(2).mean().sqrt() / feat.std())
            print("block %3d: feat max err/std %.3e rms/std %.3e (std %.3f)" % (res, e, r, float(feat.std())))
            assert r < 3e-2, "block %d diverges" % res
            # continue from the oracle's features (bf16-rounded) so errors do not compound across blocks
            x_dev = feat.permute(0, 2, 3, 1).reshape(B * res * res, -1).to(torch.bfloat16).cuda().contiguous()
            res *= 2
        assert rms < 3e-2 and err < 0.25, (err, rms)


def test_ops_match_reference_goldens():
    from layoutdetr_b200.torch_utils.ops import bias_act, upfirdn2d
    ops = golden("ops_ref.pt")
    for c in ops["bias_act"]:
        x = c["x"].cuda().requires_grad_(True)
        b = c["b"].cuda().requires_grad_(True)
        y = bias_act.bias_act(x, b, dim=c["dim"], act=c["act"], gain=c["gain"], clamp=c["clamp"])
        torch.testing.assert_close(y.detach().cpu(), c["y"], atol=2e-5, rtol=2e-5, msg=lambda m: "bias_act %s: %s" % (c["act"], m))
        dx, db = torch.autograd.grad(y, [x, b], c["dy"].cuda())
        torch.testing.assert_close(dx.cpu(), c["dx"], atol=2e-5, rtol=2e-4, msg=lambda m: "bias_act dx %s: %s" % (c["act"], m))
        torch.testing.assert_close(db.cpu(), c["db"], atol=1e-4, rtol=2e-4, msg=lambda m: "bias_act db %s: %s" % (c["act"], m))
    for c in ops["upfirdn2d"]:
        x = c["x"].cuda().requires_grad_(True)
        y = upfirdn2d.upfirdn2d(x, c["f"].cuda(), up=c["up"], down=c["down"], padding=c["padding"], flip_filter=c["flip_filter"], gain=c["gain"])
        torch.testing.assert_close(y.detach().cpu(), c["y"], atol=1e-5, rtol=1e-5)
        (dx,) = torch.autograd.grad(y, [x], c["dy"].cuda())
        torch.testing.assert_close(dx.cpu(), c["dx"], atol=1e-5, rtol=1e-5)
        # channels-last input goes through the strided path
        xcl = c["x"].cuda().contiguous(memory_format=torch.channels_last)
        ycl = upfirdn2d.upfirdn2d(xcl, c["f"].cuda(), up=c["up"], down=c["down"], padding=c["padding"], flip_filter=c["flip_filter"], gain=c["gain"])
        torch.testing.assert_close(ycl.cpu(), c["y"], atol=1e-5, rtol=1e-5)


def test_fma_and_conv2d_resample():
    from layoutdetr_b200.torch_utils.ops import fma, conv2d_resample, upfirdn2d
    from oracle import layoutdetr_oracle as O
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    a = torch.randn((2, 8, 6, 6), generator=g); b = torch.randn((2, 8, 1, 1), generator=g); c = torch.randn((1, 1, 6, 6), generator=g)
    y = fma.fma(a.cuda(), b.cuda(), c.cuda())
    torch.testing.assert_close(y.cpu(), a * b + c, atol=1e-6, rtol=1e-6)
    x = torch.randn((2, 16, 8, 8), generator=g); w = torch.randn((24, 16, 3, 3), generator=g) * 0.1
    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    y = conv2d_resample.conv2d_resample(x.cuda(), w.cuda(), f.cuda(), up=2, padding=1, flip_weight=False)
    ref = O.upfirdn2d(F.conv_transpose2d(x, w.transpose(0, 1), stride=2), f, padding=(1, 1, 1, 1), gain=4)
    assert float((y.cpu() - ref).abs().max() / ref.std()) < 3e-2
    y = conv2d_resample.conv2d_resample(x.cuda(), w.cuda(), padding=1)
    ref = F.conv2d(x, w, padding=1)
    assert float((y.cpu() - ref).abs().max() / ref.std()) < 3e-2
