"""TEST INFRASTRUCTURE — CPU restatement of one training iteration's losses and backward
(reference training/loss.py:75-218 with the default weights of train.py:262-277), built on
oracle/layoutdetr_oracle.py.  Used as the checker in tests and as the timed CPU baseline in bench.py."""
import torch
import torch.nn.functional as F

from . import layoutdetr_oracle as O

W = dict(Dreal_bbox_cls=50.0, Dreal_bbox_rec=500.0, Dreal_text_rec=0.1, Dreal_text_len_rec=2.0, Dreal_im_rec=0.5,
         Ggen_bbox_rec=100.0, Ggen_bbox_gIoU=4.0, Ggen_overlapping=7.0, Ggen_alignment=17.0, Ggen_z_rec=5.0,
         Ggen_bbox_cls=50.0, Ggen_text_rec=1.0, Ggen_text_len_rec=1.0)


def _leafs(sd, frozen_prefix="text_encoder."):
    out = {}
    for k, v in sd.items():
        if v.is_floating_point() and not k.startswith(frozen_prefix) and not any(t in k for t in ("running_", "resample_filter", "w_avg")) \
                and ".bn" not in k and "downsample.1" not in k:
            out[k] = v.detach().clone().requires_grad_(True)
        else:
            out[k] = v.detach()
    # tied LM head
    for k in list(out):
        if k.endswith("cls.predictions.decoder.weight"):
            out[k] = out[k.replace("cls.predictions.decoder.weight", "bert.embeddings.word_embeddings.weight")]
    return out


def gmain_loss(sdG, sdD, tok, inp):
    keep = ~inp["padding_mask"]
    bbox_fake, loss_z, cls_logits, loss_lm, loss_text_len, _ = O.generator_forward(
        sdG, tok, inp["z"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"], inp["background"], reconst=True)
    lg, lgu = O.discriminator_forward(sdD, tok, bbox_fake, inp["bbox_class"], inp["bbox_text"], inp["padding_mask"], inp["background"])
    terms = [F.softplus(-lg), F.softplus(-lgu),
             F.mse_loss(bbox_fake[keep], inp["bbox_real"][keep]) * W["Ggen_bbox_rec"],
             O.generalized_iou_loss(bbox_fake[keep], inp["bbox_real"][keep]) * W["Ggen_bbox_gIoU"],
             O.compute_overlap(bbox_fake, keep) * W["Ggen_overlapping"],
             O.compute_alignment(bbox_fake, keep) * W["Ggen_alignment"],
             loss_z * W["Ggen_z_rec"], F.cross_entropy(cls_logits, inp["bbox_class"][keep]) * W["Ggen_bbox_cls"],
             loss_lm * W["Ggen_text_rec"], loss_text_len * W["Ggen_text_len_rec"]]
    return sum(terms).mean()


def dmain_loss(sdG, sdD, tok, inp):
    keep = ~inp["padding_mask"]
    with torch.no_grad():
        bbox_fake = O.generator_forward(sdG, tok, inp["z"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"], inp["background"])
    lg, lgu = O.discriminator_forward(sdD, tok, bbox_fake, inp["bbox_class"], inp["bbox_text"], inp["padding_mask"], inp["background"])
    loss_fake = (F.softplus(lg) + F.softplus(lgu)).mean()
    (rl, rlu, bbox_rec, cls_logits, loss_lm, loss_text_len, bg_rec, bbox_rec_u, cls_logits_u) = O.discriminator_forward(
        sdD, tok, inp["bbox_real"], inp["bbox_class"], inp["bbox_text"], inp["padding_mask"], inp["background"], reconst=True)
    tgt = inp["bbox_class"][keep]
    terms = [F.softplus(-rl), F.softplus(-rlu), F.mse_loss(bbox_rec, inp["bbox_real"][keep]) * W["Dreal_bbox_rec"],
             F.cross_entropy(cls_logits, tgt) * W["Dreal_bbox_cls"], loss_lm * W["Dreal_text_rec"],
             loss_text_len * W["Dreal_text_len_rec"], F.mse_loss(bg_rec, inp["background"]) * W["Dreal_im_rec"],
             F.mse_loss(bbox_rec_u, inp["bbox_real"][keep]) * W["Dreal_bbox_rec"], F.cross_entropy(cls_logits_u, tgt) * W["Dreal_bbox_cls"]]
    return loss_fake, sum(terms).mean()


def iteration(sdG, sdD, tok, inp):
    """Forward + backward of Gmain and Dmain (no optimizer): returns (loss_G, loss_D_fake, loss_D_real)."""
    g = _leafs(sdG)
    lG = gmain_loss(g, sdD, tok, inp)
    lG.backward()
    d = _leafs(sdD)
    lf, lr = dmain_loss(sdG, d, tok, inp)
    lf.backward()
    lr.backward()
    return float(lG), float(lf), float(lr)


def phase_gradients(sdG, sdD, tok, inp, phase):
    """{state_dict key: gradient} of one phase ('Gmain' | 'Dmain'), as reference loss.accumulate_gradients leaves them in `.grad`
    (training/loss.py:88-218; Dmain = the fake-sample and the real-sample backward accumulated)."""
    if phase == "Gmain":
        leaves = _leafs(sdG)
        gmain_loss(leaves, sdD, tok, inp).backward()
    else:
        leaves = _leafs(sdD)
        lf, lr = dmain_loss(sdG, leaves, tok, inp)
        lf.backward()
        lr.backward()
    return {k: v.grad for k, v in leaves.items() if v.requires_grad and v.grad is not None}
