"""Drop-in for the reference's ``datasets/voc_eval.py`` (offline PASCAL VOC evaluation, SURVEY.md section 8f
rank 4): ``DetectorEvalPascal`` with the same constructor, file layout, result-file format and methods.  The
per-class matching of detections to ground truth (:249-281) runs on the GPU (``ronk_voc_match``, float64 like
the reference); parsing, the cumulative sums and the AP formulas (:126-155, :283-290) are the host-side NumPy
float64 steps of the reference, in the same order.

reference map: __init__ :28-44, parse_rec :55-73, get_voc_results_file_template :76-83,
write_voc_results_file :86-100, do_python_eval :103-123, voc_ap :126-155, voc_eval :158-296.
Differences: the annotation cache (:194-212, a pickle under cachedir) is kept in memory instead, and
nothing is printed except the AP summary of do_python_eval.
"""
import os
import pickle
import xml.etree.ElementTree as ET

import numpy as np

from .. import core

VOC_CLASSES = ('aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow', 'diningtable',
               'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train', 'tvmonitor')


class DetectorEvalPascal(object):
    def __init__(self, voc_root, devkit_root, set_type='test', output_dir='output_{}'):
        self._set_type = set_type
        output_dir = output_dir.format(set_type)
        if not os.path.isdir(output_dir):
            os.mkdir(output_dir)
        self._voc_root = voc_root
        self._output_dir = output_dir
        self._annopath = os.path.join(voc_root, 'VOC2007', 'Annotations', '%s.xml')
        self._imgpath = os.path.join(voc_root, 'VOC2007', 'JPEGImages', '%s.jpg')
        self._imgsetpath = os.path.join(voc_root, 'VOC2007', 'ImageSets', 'Main', '{:s}.txt')
        self._devkit_path = os.path.join(devkit_root, 'VOC2007')
        rootpath = os.path.join(voc_root, 'VOC2007')
        with open(os.path.join(rootpath, 'ImageSets', 'Main', set_type + '.txt')) as f:
            self._image_ids = [(rootpath, line.strip()) for line in f]
        self._recs = None

    @property
    def image_ids(self):
        return self._image_ids

    @property
    def output_dir(self):
        return self._output_dir

    def evaluate_detections(self, box_list):
        self.write_voc_results_file(box_list)
        self.do_python_eval()

    def parse_rec(self, filename):
        """reference :55-73: the objects of one annotation file; boxes are shifted to 0-based."""
        objects = []
        for obj in ET.parse(filename).findall('object'):
            bbox = obj.find('bndbox')
            objects.append({'name': obj.find('name').text, 'pose': obj.find('pose').text,
                            'truncated': int(obj.find('truncated').text), 'difficult': int(obj.find('difficult').text),
                            'bbox': [int(bbox.find(k).text) - 1 for k in ('xmin', 'ymin', 'xmax', 'ymax')]})
        return objects

    def get_voc_results_file_template(self, cls):
        filedir = os.path.join(self._devkit_path, 'results')
        if not os.path.exists(filedir):
            os.makedirs(filedir)
        return os.path.join(filedir, 'det_' + self._set_type + '_%s.txt' % cls)

    def write_voc_results_file(self, all_boxes):
        """reference :86-100: all_boxes[class 1..20][image] = [k,5] (x1, y1, x2, y2, score), 0-based pixels;
        one line per detection: image id, score (3 decimals), 1-based corners (1 decimal)."""
        for cls_ind, cls in enumerate(VOC_CLASSES):
            with open(self.get_voc_results_file_template(cls), 'wt') as f:
                for im_ind, index in enumerate(self._image_ids):
                    dets = all_boxes[cls_ind + 1][im_ind]
                    if isinstance(dets, (list, tuple)) and len(dets) == 0:
                        continue
                    dets = np.asarray(dets.cpu() if hasattr(dets, 'cpu') else dets)
                    for k in range(dets.shape[0]):
                        f.write('{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n'.format(
                            index[1], dets[k, -1], dets[k, 0] + 1, dets[k, 1] + 1, dets[k, 2] + 1, dets[k, 3] + 1))

    def do_python_eval(self, use_07=True):
        cachedir = os.path.join(self._devkit_path, 'annotations_cache')
        aps = []
        for cls in VOC_CLASSES:
            rec, prec, ap = self.voc_eval(self.get_voc_results_file_template(cls), cls, cachedir, ovthresh=0.5,
                                          use_07_metric=use_07)
            aps.append(ap)
            print('AP for {} = {:.4f}'.format(cls, ap))
            with open(os.path.join(self._output_dir, cls + '_pr.pkl'), 'wb') as f:
                pickle.dump({'rec': rec, 'prec': prec, 'ap': ap}, f)
        print('Mean AP = {:.4f}'.format(np.mean(aps)))
        return aps

    def voc_ap(self, rec, prec, use_07_metric=True):
        """reference :126-155."""
        if use_07_metric:
            ap = 0.
            for t in np.arange(0., 1.1, 0.1):
                p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
                ap = ap + p / 11.
            return ap
        mrec = np.concatenate(([0.], rec, [1.]))
        mpre = np.concatenate(([0.], prec, [0.]))
        for i in range(mpre.size - 1, 0, -1):
            mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
        i = np.where(mrec[1:] != mrec[:-1])[0]
        return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])

    def _annotations(self):
        if self._recs is None:
            with open(self._imgsetpath.format(self._set_type), 'r') as f:
                names = [x.strip() for x in f.readlines()]
            self._recs = (names, {n: self.parse_rec(self._annopath % n) for n in names})
        return self._recs

    def voc_eval(self, detpath, classname, cachedir=None, ovthresh=0.5, use_07_metric=True):
        """reference :158-296: (rec, prec, ap) of one class from its result file; (-1., -1., -1.) when the file
        has no detection."""
        imagenames, recs = self._annotations()
        index = {n: i for i, n in enumerate(imagenames)}
        gt_boxes, gt_diff, gt_off = [], [], [0]
        for n in imagenames:                                           # :215-226
            R = [o for o in recs[n] if o['name'] == classname]
            gt_boxes.extend(o['bbox'] for o in R)
            gt_diff.extend(o['difficult'] for o in R)
            gt_off.append(len(gt_boxes))
        npos = int(np.sum(~np.asarray(gt_diff, bool)))
        with open(detpath.format(classname), 'r') as f:
            lines = f.readlines()
        if not any(lines):
            return -1., -1., -1.
        split = [x.strip().split(' ') for x in lines]
        confidence = np.array([float(x[1]) for x in split])
        BB = np.array([[float(z) for z in x[2:]] for x in split])
        img = np.array([index[x[0]] for x in split], np.int64)
        sorted_ind = np.argsort(-confidence)                           # :242 (the order of equal scores is NumPy's)
        BB, img = BB[sorted_ind, :], img[sorted_ind]
        group = np.argsort(img, kind='stable')                         # by image, confidence order kept inside
        det_off = np.concatenate([[0], np.cumsum(np.bincount(img, minlength=len(imagenames)))])
        tp_g, fp_g = core.voc_match(BB[group], det_off, np.asarray(gt_boxes, np.float64).reshape(-1, 4), gt_off,
                                    np.asarray(gt_diff, np.uint8), ovthresh)
        tp = np.zeros(len(img)); fp = np.zeros(len(img))
        tp[group] = tp_g.cpu().numpy()
        fp[group] = fp_g.cpu().numpy()
        fp = np.cumsum(fp)                                             # :283-290
        tp = np.cumsum(tp)
        rec = tp / float(npos)
        prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
        return rec, prec, self.voc_ap(rec, prec, use_07_metric)
