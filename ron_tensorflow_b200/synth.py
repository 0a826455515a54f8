"""Seeded synthetic VOC-shaped inputs for the detection hot path.

NumPy only, so the same arrays can feed the CUDA path, the oracle and the golden
generator.  Seeds follow ``20260000 + config * 1000 + image_index``.
"""
import numpy as np

SEED_BASE = 20260000


def image_seed(config, image_index):
    return SEED_BASE + config * 1000 + image_index


def make_gt(seed, num_gt, num_classes=21):
    """Ground truth of one image: boxes f32[G,4] (ymin,xmin,ymax,xmax) in [0,1], labels i64[G].

    Centres U(0.1,0.9)^2, sides log-uniform in [0.05,0.6], corners clipped to [0,1],
    minimum side 0.02 (what preprocessing hands to RONNet.bboxes_encode:
    reference preprocessing/tf_image.py:416-419 clips GT to the crop)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.uniform(0.1, 0.9, size=(num_gt, 2))
    s = np.exp(rng.uniform(np.log(0.05), np.log(0.6), size=(num_gt, 2)))
    lo = np.clip(c - s / 2, 0., 1.)
    hi = np.clip(c + s / 2, 0., 1.)
    hi = np.maximum(hi, lo + 0.02)
    over = hi > 1.
    lo = np.where(over, lo - (hi - 1.), lo)
    hi = np.minimum(hi, 1.)
    boxes = np.concatenate([lo, hi], 1).astype(np.float32)
    labels = rng.integers(1, num_classes, size=num_gt).astype(np.int64)
    return boxes, labels


def make_gt_batch(config, batch, g_lo, g_hi, num_classes=21, g_max=None, first_image=0):
    """Padded batch: boxes f32[B,Gmax,4], labels i64[B,Gmax], counts i32[B]."""
    counts = np.zeros(batch, np.int32)
    items = []
    for b in range(batch):
        seed = image_seed(config, first_image + b)
        rng = np.random.Generator(np.random.PCG64(seed ^ 0x5bd1e995))
        g = int(rng.integers(g_lo, g_hi + 1))
        counts[b] = g
        items.append(make_gt(seed, g, num_classes))
    g_max = int(g_max or counts.max())
    boxes = np.zeros((batch, g_max, 4), np.float32)
    labels = np.zeros((batch, g_max), np.int64)
    for b, (bx, lb) in enumerate(items):
        boxes[b, :counts[b]] = bx
        labels[b, :counts[b]] = lb
    return boxes, labels, counts


def make_predictions(seed, batch, num_anchors, num_classes=21, hot=300, dense=False):
    """Network-output stand-ins for ``batch`` images.

    loc f32[B,N,4] ~ N(0,0.5) clipped to [-3,3] (so exp stays finite); class scores
    f32[B,N,C] = softmax of N(0,1) logits with +4 on background and +7 on one random
    class for ``hot`` anchors per image; objectness f32[B,N] = sigmoid(N(-4,1.5)), 0.9 on
    the hot anchors.  ``dense=True`` drops the background bias (top-k stress)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    B, N, C = batch, num_anchors, num_classes
    loc = np.clip(rng.normal(0., 0.5, size=(B, N, 4)), -3., 3.).astype(np.float32)
    logits = rng.normal(0., 1., size=(B, N, C)).astype(np.float32)
    if not dense:
        logits[:, :, 0] += np.float32(4.)
    obj = rng.normal(-4., 1.5, size=(B, N)).astype(np.float32)
    obj = (1. / (1. + np.exp(-obj.astype(np.float64)))).astype(np.float32)
    hot = min(hot, N)
    for b in range(B):
        idx = rng.choice(N, size=hot, replace=False)
        cls = rng.integers(1, C, size=hot)
        logits[b, idx, cls] += np.float32(7.)
        obj[b, idx] = np.float32(0.9)
    m = logits.max(-1, keepdims=True)
    e = np.exp((logits - m).astype(np.float32))
    pred = (e / e.sum(-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return loc, pred, obj


def split_layers(x, layer_sizes, feat_shapes=None, anchors_per_cell=None):
    """[B,N,...] -> list of per-layer [B,n_l,...] (or [B,H,W,A,...] when shapes are given),
    the list-of-layers form the reference's net emits (nets/ron_vgg_320.py:486-508)."""
    out = []
    o = 0
    for i, n in enumerate(layer_sizes):
        t = np.ascontiguousarray(x[:, o:o + n])
        if feat_shapes is not None:
            H, W = feat_shapes[i]
            t = t.reshape((x.shape[0], H, W, anchors_per_cell[i]) + x.shape[2:])
        out.append(t)
        o += n
    return out


def make_loss_inputs(seed, batch, num_anchors=21250, num_classes=21):
    """Synthetic network outputs for the loss-mask stage (reference nets/ron_vgg_320.py:635-771): class logits
    [B,N,C], localisations [B,N,4], objectness logits [B,N,2] and objectness scores [B,N] = sigmoid(N(-3,2))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    logits = rng.normal(0, 1, (batch, num_anchors, num_classes)).astype(np.float32)
    loc = rng.normal(0, 0.5, (batch, num_anchors, 4)).astype(np.float32)
    obj_logits = rng.normal(0, 1, (batch, num_anchors, 2)).astype(np.float32)
    obj_pred = (1. / (1. + np.exp(-rng.normal(-3, 2, (batch, num_anchors))))).astype(np.float32)
    return logits, loc, obj_logits, obj_pred


VOC_CLASSES = ('aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow', 'diningtable',
               'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train', 'tvmonitor')


def make_voc_eval_case(seed, n_images=40, width=500, height=375):
    """A small synthetic PASCAL-VOC evaluation problem (reference datasets/voc_eval.py): per image a list of objects
    {name, difficult, bbox [xmin, ymin, xmax, ymax] 1-based ints} and, per class index 1..20, per image, a float32
    [k,5] array of detections (x1, y1, x2, y2, score) -- jittered copies of the objects plus random boxes.  The
    scores of one class are all different to 3 decimals (the result files keep 3, and np.argsort leaves the order
    of equal scores unspecified)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ids = ['%06d' % (i + 1) for i in range(n_images)]
    annots = []
    for _ in ids:
        objs = []
        for _ in range(int(rng.integers(1, 7))):
            w, h = int(rng.integers(30, 250)), int(rng.integers(30, 200))
            x0, y0 = int(rng.integers(1, width - w)), int(rng.integers(1, height - h))
            objs.append(dict(name=VOC_CLASSES[int(rng.integers(0, 20))], difficult=int(rng.random() < 0.15),
                             bbox=[x0, y0, x0 + w, y0 + h]))
        annots.append(objs)
    all_boxes = [[np.zeros((0, 5), np.float32) for _ in ids] for _ in range(21)]
    for c in range(1, 21):
        dets = []
        for i, objs in enumerate(annots):
            for o in objs:
                if o['name'] != VOC_CLASSES[c - 1]:
                    continue
                for _ in range(int(rng.integers(0, 4))):                 # 0-3 detections near the object
                    j = rng.normal(0, 12, 4)
                    b = np.array(o['bbox'], np.float64) - 1 + j
                    dets.append((i, b))
            for _ in range(int(rng.integers(0, 3))):                     # and a few anywhere
                w, h = rng.uniform(20, 300), rng.uniform(20, 250)
                x0, y0 = rng.uniform(0, width - 20), rng.uniform(0, height - 20)
                dets.append((i, np.array([x0, y0, x0 + w, y0 + h])))
        scores = rng.permutation(999)[:len(dets)] + 1                    # distinct k / 1000
        per_img = [[] for _ in ids]
        for (i, b), sc in zip(dets, scores):
            per_img[i].append(np.concatenate([np.round(b, 1), [sc / 1000.]]))
        for i in range(n_images):
            if per_img[i]:
                all_boxes[c][i] = np.array(per_img[i], np.float32)
    return ids, annots, all_boxes


def write_voc_tree(voc_root, ids, annots, set_type='test', with_size=None, with_images=False):
    """Annotations/*.xml + ImageSets/Main/<set>.txt in the layout datasets/voc_eval.py:37-44 reads; with_size =
    (height, width, depth) adds the <size> element and with_images tiny stand-in JPEGImages/*.jpg files, which
    datasets/pascalvoc_to_tfrecords.py needs."""
    import os
    base = os.path.join(voc_root, 'VOC2007')
    os.makedirs(os.path.join(base, 'Annotations'), exist_ok=True)
    os.makedirs(os.path.join(base, 'ImageSets', 'Main'), exist_ok=True)
    with open(os.path.join(base, 'ImageSets', 'Main', set_type + '.txt'), 'wt') as f:
        f.write(''.join(i + '\n' for i in ids))
    for i, objs in zip(ids, annots):
        xml = ['<annotation><filename>%s.jpg</filename>' % i]
        if with_size:
            xml.append('<size><width>%d</width><height>%d</height><depth>%d</depth></size>' % (with_size[1], with_size[0], with_size[2]))
        if with_images:
            os.makedirs(os.path.join(base, 'JPEGImages'), exist_ok=True)
            with open(os.path.join(base, 'JPEGImages', i + '.jpg'), 'wb') as f:
                f.write(b'\xff\xd8 synthetic ' + i.encode('ascii') + b' \xff\xd9')
        for o in objs:
            xml.append('<object><name>%s</name><pose>Unspecified</pose><truncated>0</truncated><difficult>%d</difficult>'
                       '<bndbox><xmin>%d</xmin><ymin>%d</ymin><xmax>%d</xmax><ymax>%d</ymax></bndbox></object>'
                       % ((o['name'], o['difficult']) + tuple(o['bbox'])))
        xml.append('</annotation>')
        with open(os.path.join(base, 'Annotations', i + '.xml'), 'wt') as f:
            f.write(''.join(xml))
