"""Loader throughput on the host: samples/s of LayoutDataset in full mode (every key the reference
loader produces — nine 1024^2 patches_orig + masks per sample) vs lean mode (only what the training path reads), on a
synthetic zip with the real on-disk geometry (1024 x 1024 pages, 8 elements).  CPU only; prints one JSON line.

    python tools/loader_bench.py [--samples 6] [--page 1024] [--background-size 256]
"""
import argparse, io, json, os, sys, tempfile, time, zipfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import PIL.Image

from layoutdetr_b200.training.dataset_layoutganpp import LayoutDataset


def make_zip(path, samples, page, elems=8, seed=0):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:page, 0:page].astype(np.float32)

    def img(h, w, c=3):
        a = np.stack([127 + 120 * np.sin(xx[:h, :w] / (7 + 40 * rng.rand())) * np.cos(yy[:h, :w] / (5 + 40 * rng.rand())) for _ in range(c)], -1)
        return np.clip(a + rng.randint(0, 8, a.shape), 0, 255).astype(np.uint8)

    meta = []
    with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as z:
        def put(name, arr):
            buf = io.BytesIO()
            PIL.Image.fromarray(arr).save(buf, format="PNG", compress_level=1)
            z.writestr(name, buf.getvalue())
        for s in range(samples):
            base = "p%03d" % s
            for i in range(elems):
                put("%s_%d_patch.png" % (base, i), img(120, 400))
                put("%s_%d_patch_orig.png" % (base, i), img(page, page))
                put("%s_%d_patch_mask.png" % (base, i), (img(page, page, 1)[:, :, 0] > 127).astype(np.uint8) * 255)
            put(base + "_background_orig.png", img(page, page))
            meta.append([base, dict(bboxes=rng.uniform(0.1, 0.6, (elems, 4)).tolist(), labels=[int(v) for v in rng.randint(0, 8, elems)],
                                    texts=["element %d" % i for i in range(elems)], page_label=None,
                                    attr=dict(name=base, width=page, height=page, num_bbox_labels=8))])
        z.writestr("non_image.json", json.dumps(dict(samples=meta)))


def rate(ds, n):
    t0 = time.perf_counter()
    nbytes = 0
    for i in range(n):
        s, _ = ds[i % len(ds)]
        nbytes += sum(v.nbytes for v in s.values() if isinstance(v, np.ndarray))
    dt = time.perf_counter() - t0
    return n / dt, nbytes / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--page", type=int, default=1024)
    ap.add_argument("--background-size", type=int, default=256)
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "a", "b", "train.zip")
        os.makedirs(os.path.dirname(path))
        make_zip(path, args.samples, args.page)
        full = LayoutDataset(path=path, background_size=args.background_size)
        lean = LayoutDataset(path=path, background_size=args.background_size, lean=True)
        r_full, b_full = rate(full, args.samples)
        r_lean, b_lean = rate(lean, args.samples * 4)
    print(json.dumps(dict(metric="LayoutDataset samples/s per host core (one worker)", full=r_full, lean=r_lean, unit="samples/s",
                          bytes_per_sample_full=b_full, bytes_per_sample_lean=b_lean, page=args.page, elements=8,
                          background_size=args.background_size,
                          note="full = every key of the reference loader (its arithmetic, bit-identical on the fixture); lean = hot-path keys only")))


if __name__ == "__main__":
    main()
