"""Training objective — mirror of the reference's training/loss.py (StyleGAN2Loss.accumulate_gradients :75-218):
same constructor arguments, phase names, loss terms and weights; G/D are the sm_100a modules.

Path-length (Greg) and R1 (Dreg) regularisers need double backward through the kernels; they are disabled
in the reference's default configuration (`--pl-weight 0 --gamma 0`, train.py) and raise here if enabled.
"""
import torch
import torch.nn.functional as F

from ..metrics.metric_layoutnet import generalized_iou_loss, layout_overlap_alignment
from ..torch_utils import training_stats
from .. import functional as Fn
from ..lanes import LANES
from .networks_detr import valid_index


def _after_backward():
    """Backward kernels run on the lanes (streams) of their forward and add into parameter gradients as a side effect
    autograd does not see: the issuing stream waits for them here (lanes.py)."""
    if LANES.active(2):
        LANES.join_children()


class Loss:
    def accumulate_gradients(self, phase, bbox_real, bbox_class, bbox_text, bbox_patch, padding_mask, background, real_c, gen_z, gen_c, gain, cur_nimg):
        raise NotImplementedError()


class StyleGAN2Loss(Loss):
    def __init__(self, device, G, D, augment_pipe=None, r1_gamma=0.0, style_mixing_prob=0, pl_weight=0.0, pl_batch_shrink=2, pl_decay=0.01, pl_no_weight_grad=False, blur_init_sigma=0, blur_fade_kimg=0,
                 Dreal_bbox_cls_weight=50.0, Dreal_bbox_rec_weight=500.0, Dreal_text_rec_weight=0.1, Dreal_text_len_rec_weight=2.0, Dreal_im_rec_weight=0.5,
                 Ggen_bbox_rec_weight=100.0, Ggen_bbox_gIoU_weight=4.0, Ggen_overlapping_weight=7.0, Ggen_alignment_weight=17.0,
                 Ggen_z_rec_weight=5.0, Ggen_bbox_cls_weight=50.0, Ggen_text_rec_weight=1.0, Ggen_text_len_rec_weight=1.0):
        super().__init__()
        self.device = device
        self.G = G
        self.D = D
        self.augment_pipe = augment_pipe
        self.r1_gamma = r1_gamma
        self.pl_weight = pl_weight
        self.w = dict(Dreal_bbox_cls=Dreal_bbox_cls_weight, Dreal_bbox_rec=Dreal_bbox_rec_weight, Dreal_text_rec=Dreal_text_rec_weight,
                      Dreal_text_len_rec=Dreal_text_len_rec_weight, Dreal_im_rec=Dreal_im_rec_weight,
                      Ggen_bbox_rec=Ggen_bbox_rec_weight, Ggen_bbox_gIoU=Ggen_bbox_gIoU_weight, Ggen_overlapping=Ggen_overlapping_weight,
                      Ggen_alignment=Ggen_alignment_weight, Ggen_z_rec=Ggen_z_rec_weight, Ggen_bbox_cls=Ggen_bbox_cls_weight,
                      Ggen_text_rec=Ggen_text_rec_weight, Ggen_text_len_rec=Ggen_text_len_rec_weight)
        self.last = {}

    def run_G(self, z, bbox_class, bbox_real, bbox_text, bbox_patch, padding_mask, background, c, reconst=False, update_emas=False):
        return self.G(z, bbox_class, bbox_real, bbox_text, bbox_patch, padding_mask, background, c, reconst)

    def run_D(self, bbox, bbox_class, bbox_text, bbox_patch, padding_mask, background, c, reconst=False, blur_sigma=0, update_emas=False):
        return self.D(bbox, bbox_class, bbox_text, bbox_patch, padding_mask, background, c, reconst)

    def accumulate_gradients(self, phase, bbox_real, bbox_class, bbox_text, bbox_patch, padding_mask, background, real_c, gen_z, gen_c, gain, cur_nimg,
                             before_backward=None):
        """Phases as in the reference, plus 'Dgen' / 'Dreal': the two halves of 'Dmain' (fake-sample pass :146-157 and
        real-sample pass :161-218), which the lane scheduler issues on different streams.  `before_backward` is called
        right before each `.backward()` (stream ordering hook)."""
        assert phase in ['Gmain', 'Greg', 'Gboth', 'Dmain', 'Dreg', 'Dboth', 'Dgen', 'Dreal']
        pre_bwd = before_backward if before_backward is not None else (lambda: None)
        if self.pl_weight == 0:
            phase = {'Greg': 'none', 'Gboth': 'Gmain'}.get(phase, phase)
        if self.r1_gamma == 0:
            phase = {'Dreg': 'none', 'Dboth': 'Dmain'}.get(phase, phase)
        if phase in ('Greg', 'Gboth', 'Dreg', 'Dboth'):
            raise NotImplementedError("path-length / R1 regularisation needs double backward, which the sm_100a kernels do not "
                                      "provide; the reference's defaults (--pl-weight 0 --gamma 0) never reach it")
        W = self.w
        report = training_stats.report
        keep = ~padding_mask
        vidx, _ = valid_index(padding_mask)
        sel = lambda t: t.reshape(-1, *t.shape[2:]).index_select(0, vidx)      # == t[keep], without a device sync

        if phase == 'Gmain':
            bbox_fake, loss_z, cls_logits, loss_lm, loss_text_len = self.run_G(gen_z, bbox_class, bbox_real, bbox_text, bbox_patch, padding_mask, background, gen_c, reconst=True)
            gen_logits, gen_logits_uncond = self.run_D(bbox_fake, bbox_class, bbox_text, bbox_patch, padding_mask, background, gen_c)
            report('Loss/scores/fake', gen_logits)
            overlapping, alignment = layout_overlap_alignment(bbox_fake, keep)      # compute_overlap + compute_alignment, one launch
            terms = dict(
                loss_Ggen=F.softplus(-gen_logits),
                loss_Ggen_uncond=F.softplus(-gen_logits_uncond),
                loss_Ggen_bbox_rec=F.mse_loss(sel(bbox_fake), sel(bbox_real)) * W['Ggen_bbox_rec'],
                loss_Ggen_bbox_gIoU=generalized_iou_loss(sel(bbox_fake), sel(bbox_real)) * W['Ggen_bbox_gIoU'],
                loss_Ggen_overlapping=overlapping * W['Ggen_overlapping'],
                loss_Ggen_alignment=alignment * W['Ggen_alignment'],
                loss_Ggen_z_rec=loss_z * W['Ggen_z_rec'],
                loss_Ggen_bbox_cls=Fn.cross_entropy(cls_logits, sel(bbox_class)) * W['Ggen_bbox_cls'],
                loss_Ggen_text_rec=loss_lm * W['Ggen_text_rec'],
                loss_Ggen_text_len_rec=loss_text_len * W['Ggen_text_len_rec'],
            )
            for k, v in terms.items():
                report('Loss/G/' + k, v)
            self.last['Gmain'] = {k: v.detach() for k, v in terms.items()}
            loss = sum(terms.values()).mean().mul(gain)
            pre_bwd()
            loss.backward()
            _after_backward()

        if phase in ('Dmain', 'Dgen'):
            with torch.no_grad():
                bbox_fake = self.run_G(gen_z, bbox_class, bbox_real, bbox_text, bbox_patch, padding_mask, background, gen_c, update_emas=True)
            gen_logits, gen_logits_uncond = self.run_D(bbox_fake, bbox_class, bbox_text, bbox_patch, padding_mask, background, gen_c, update_emas=True)
            report('Loss/scores/fake', gen_logits)
            loss_Dgen = F.softplus(gen_logits)
            loss_Dgen_uncond = F.softplus(gen_logits_uncond)
            report('Loss/D/loss_Dgen', loss_Dgen)
            report('Loss/D/loss_Dgen_uncond', loss_Dgen_uncond)
            self.last.setdefault('Dmain', {}).update(loss_Dgen=loss_Dgen.detach(), loss_Dgen_uncond=loss_Dgen_uncond.detach())
            loss = (loss_Dgen + loss_Dgen_uncond).mean().mul(gain)
            pre_bwd()
            loss.backward()
            _after_backward()

        if phase in ('Dmain', 'Dreal'):
            bbox_real_tmp = bbox_real.detach()
            (real_logits, real_logits_uncond, bbox_rec, cls_logits, loss_lm, loss_text_len, bg_rec, bbox_rec_uncond,
             cls_logits_uncond) = self.run_D(bbox_real_tmp, bbox_class, bbox_text, bbox_patch, padding_mask, background, real_c, reconst=True)
            report('Loss/scores/real', real_logits)
            cls_tgt = sel(bbox_class)
            terms = dict(
                loss_Dreal=F.softplus(-real_logits),
                loss_Dreal_uncond=F.softplus(-real_logits_uncond),
                loss_Dreal_bbox_rec=F.mse_loss(bbox_rec, sel(bbox_real_tmp)) * W['Dreal_bbox_rec'],
                loss_Dreal_bbox_cls=Fn.cross_entropy(cls_logits, cls_tgt) * W['Dreal_bbox_cls'],
                loss_Dreal_text_rec=loss_lm * W['Dreal_text_rec'],
                loss_Dreal_text_len_rec=loss_text_len * W['Dreal_text_len_rec'],
                loss_Dreal_bg_rec=F.mse_loss(bg_rec, background) * W['Dreal_im_rec'],
                loss_Dreal_bbox_rec_uncond=F.mse_loss(bbox_rec_uncond, sel(bbox_real_tmp)) * W['Dreal_bbox_rec'],
                loss_Dreal_bbox_cls_uncond=Fn.cross_entropy(cls_logits_uncond, cls_tgt) * W['Dreal_bbox_cls'],
            )
            for k, v in terms.items():
                report('Loss/D/' + k, v)
            self.last.setdefault('Dmain', {}).update({k: v.detach() for k, v in terms.items()})
            loss = sum(terms.values()).mean().mul(gain)
            pre_bwd()
            loss.backward()
            _after_backward()
