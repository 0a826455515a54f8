"""datasets/voc_eval.py: result-file format and the PASCAL VOC evaluator.  The
fixture is the output of the UNMODIFIED reference class on a synthetic VOC tree (pure NumPy;
tests/golden/make_golden.py: gen_voc_eval)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import ron_oracle as O
from ron_tensorflow_b200 import synth
from _util import need_cuda, eq


def _tree(tmp_path, g):
    seed, n = int(g['cfg'][0]), int(g['cfg'][1])
    ids, annots, all_boxes = synth.make_voc_eval_case(seed, n)
    synth.write_voc_tree(str(tmp_path / 'voc'), ids, annots)
    return ids, annots, all_boxes


def _evaluator(tmp_path):
    from ron_tensorflow_b200.datasets import voc_eval
    return voc_eval.DetectorEvalPascal(str(tmp_path / 'voc'), str(tmp_path / 'devkit'), 'test',
                                       output_dir=str(tmp_path / 'output_{}'))


def test_result_files_match_reference_format(golden, tmp_path):
    """write_voc_results_file is host-side text formatting: byte-identical files, no GPU needed."""
    g = golden('voc_eval')
    ids, annots, all_boxes = _tree(tmp_path, g)
    ev = _evaluator(tmp_path)
    ev.write_voc_results_file(all_boxes)
    blob = b''.join(open(ev.get_voc_results_file_template(c), 'rb').read() for c in synth.VOC_CLASSES)
    if hashlib.sha256(blob).digest() != g['files_sha'].tobytes():
        person = open(ev.get_voc_results_file_template('person'), 'rb').read()
        assert person == g['person_file'].tobytes()
        pytest.fail('result files differ from the reference (the person file matches)')
    assert open(ev.get_voc_results_file_template('person'), 'rb').read() == g['person_file'].tobytes()


def _oracle_eval(ev, cls, use07):
    names, recs = ev._annotations()
    index = {n: i for i, n in enumerate(names)}
    gtb, gtd, off = [], [], [0]
    for n in names:
        R = [o for o in recs[n] if o['name'] == cls]
        gtb.extend(o['bbox'] for o in R); gtd.extend(o['difficult'] for o in R); off.append(len(gtb))
    lines = [l.strip().split(' ') for l in open(ev.get_voc_results_file_template(cls))]
    conf = np.array([float(x[1]) for x in lines])
    BB = np.array([[float(z) for z in x[2:]] for x in lines])
    img = np.array([index[x[0]] for x in lines])
    order = np.argsort(-conf)
    tp, fp = O.voc_match(img[order], BB[order], off, gtb, gtd, 0.5)
    tp, fp = np.cumsum(tp), np.cumsum(fp)
    rec = tp / float(np.sum(~np.asarray(gtd, bool)))
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, O.voc_ap(rec, prec, use07)


def test_oracle_matches_reference_voc_eval(golden, tmp_path):
    g = golden('voc_eval')
    _, _, all_boxes = _tree(tmp_path, g)
    ev = _evaluator(tmp_path)
    ev.write_voc_results_file(all_boxes)
    for ci, cls in enumerate(synth.VOC_CLASSES):
        rec, prec, ap07 = _oracle_eval(ev, cls, True)
        _, _, ap12 = _oracle_eval(ev, cls, False)
        assert ap07 == g['ap07'][ci] and ap12 == g['ap12'][ci], cls
        assert rec.shape[0] == g['n_dets'][ci]
        if ci in (6, 14):
            assert np.array_equal(rec, g['rec_%d' % ci]) and np.array_equal(prec, g['prec_%d' % ci])


@pytest.mark.gpu
def test_cuda_matches_reference_voc_eval(golden, tmp_path):
    need_cuda()
    g = golden('voc_eval')
    _, _, all_boxes = _tree(tmp_path, g)
    ev = _evaluator(tmp_path)
    ev.write_voc_results_file(all_boxes)
    for ci, cls in enumerate(synth.VOC_CLASSES):
        fn = ev.get_voc_results_file_template(cls)
        rec, prec, ap07 = ev.voc_eval(fn, cls, None, ovthresh=0.5, use_07_metric=True)
        _, _, ap12 = ev.voc_eval(fn, cls, None, ovthresh=0.5, use_07_metric=False)
        assert ap07 == g['ap07'][ci] and ap12 == g['ap12'][ci], cls       # float64, bit for bit
        if ci in (6, 14):
            eq(rec, g['rec_%d' % ci], 'recall'); eq(prec, g['prec_%d' % ci], 'precision')
    aps = ev.do_python_eval(use_07=True)
    assert np.array_equal(np.asarray(aps), g['ap07'])
    assert os.path.exists(os.path.join(ev.output_dir, 'person_pr.pkl'))


@pytest.mark.gpu
def test_cuda_voc_match_vs_oracle_large_and_degenerate():
    """2 000 images, up to 60 ground-truth boxes each (two lane rounds), empty images, degenerate boxes (NaN overlap)."""
    need_cuda()
    from ron_tensorflow_b200 import core
    rng = np.random.Generator(np.random.PCG64(12))
    n_img = 2000
    gt_cnt = rng.integers(0, 61, n_img); gt_cnt[::7] = 0
    det_cnt = rng.integers(0, 30, n_img)
    gt_off = np.concatenate([[0], np.cumsum(gt_cnt)]); det_off = np.concatenate([[0], np.cumsum(det_cnt)])
    def boxes(n):
        c = rng.uniform(50, 450, (n, 2)); s = rng.uniform(10, 200, (n, 2))
        return np.round(np.concatenate([c - s / 2, c + s / 2], 1), 1)
    gtb = np.floor(boxes(gt_off[-1])); detb = boxes(det_off[-1])
    # detections copied from ground truth (exact hits, duplicates -> FP after the first), degenerate pairs
    for i in range(0, n_img, 3):
        if gt_cnt[i] and det_cnt[i] >= 2:
            detb[det_off[i]] = gtb[gt_off[i]]; detb[det_off[i] + 1] = gtb[gt_off[i]]
    gtb[5] = [10, 10, 10, 10]
    if det_cnt[np.searchsorted(gt_off, 5, 'right') - 1]:
        detb[det_off[np.searchsorted(gt_off, 5, 'right') - 1]] = [10, 10, 10, 10]
    diff = (rng.random(gt_off[-1]) < 0.2).astype(np.uint8)
    img = np.repeat(np.arange(n_img), det_cnt)
    rtp, rfp = O.voc_match(img, detb, gt_off, gtb, diff, 0.5)
    tp, fp = core.voc_match(detb, det_off, gtb, gt_off, diff, 0.5)
    eq(tp.cpu().numpy().astype(np.float64), rtp, 'tp'); eq(fp.cpu().numpy().astype(np.float64), rfp, 'fp')
    assert rtp.sum() > 100 and rfp.sum() > 100
