"""CPU: with the overlay first on sys.path, `training.networks_detr` / `torch_utils.ops.*` resolve to the sm_100a
implementations while untouched modules still come from the reference checkout (only where it exists)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LAYOUTDETR_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "training")), reason="reference checkout not present on this box")
def test_overlay_resolution():
    code = r'''
import sys, types
for n in ("seaborn", "skimage", "skimage.transform", "pytorch_fid", "pytorch_fid.fid_score"):
    m = types.ModuleType(n); sys.modules[n] = m
sys.modules["pytorch_fid.fid_score"].calculate_frechet_distance = None
import training.networks_detr as nd, torch_utils.ops.bias_act as ba, torch_utils.ops.upfirdn2d as uf
import torch_utils.ops.conv2d_gradfix as cg, torch_utils.misc as misc, dnnlib
assert nd.Generator.__module__ == "layoutdetr_b200.training.networks_detr", nd.Generator.__module__
assert "layoutdetr_b200" in ba.bias_act.__module__ and "layoutdetr_b200" in uf.upfirdn2d.__module__
import training.networks_layoutnet as ln, training.dataset_layoutganpp as dsl
assert ln.LayoutNet.__module__ == "layoutdetr_b200.training.networks_layoutnet" and "layoutdetr_b200" in dsl.LayoutDataset.__module__
import metrics.layout_frechet_inception_distance as lfid, metrics.overlap50k_alignment50k_layoutwise_iou50k_layoutwise_docsim50k as oa
import metrics.metric_utils_layout as mul
assert lfid.compute_layout_fid.__module__ == "layoutdetr_b200.metrics.sweep_entry" and mul.__file__.startswith(%r)
assert oa.compute_overlap_alignment_laywise_IoU_layerwise_DocSim.__module__ == "layoutdetr_b200.metrics.sweep_entry"
assert hasattr(cg, "no_weight_gradients") and misc.__file__.startswith(%r) and dnnlib.__file__.startswith(%r)
import training.training_loop as tl, training.loss as tloss, metrics.metric_layoutnet as mln, inspect
assert tl.training_loop.__module__ == "layoutdetr_b200.training.training_loop" and tloss.StyleGAN2Loss.__module__ == "layoutdetr_b200.training.loss"
assert mln.generalized_iou_loss.__module__ == "layoutdetr_b200.metrics.metric_layoutnet" and hasattr(mln, "compute_iou_for_layout")
import importlib.util
spec = importlib.util.spec_from_file_location("_ref_tl", %r + "/training/training_loop.py")
ref_src = open(spec.origin).read()
import ast
ref_fn = [n for n in ast.parse(ref_src).body if isinstance(n, ast.FunctionDef) and n.name == "training_loop"][0]
ref_args = [a.arg for a in ref_fn.args.args]
ours = list(inspect.signature(tl.training_loop).parameters)
assert ours[:len(ref_args)] == ref_args, (ours, ref_args)          # same keyword arguments, same order (+ extensions at the end)
print("OVERLAY_OK")
''' % (REF, REF, REF, REF)
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
        probe = f.name
    env = dict(os.environ, LAYOUTDETR_REFERENCE=REF, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "layoutdetr_b200.dropin.run", probe], env=env, capture_output=True, text=True,
                       cwd=ROOT, timeout=300)
    os.unlink(probe)
    assert "OVERLAY_OK" in r.stdout, r.stdout + r.stderr
