"""Drop-in for the hot-path half of the reference's ``nets/ssd_vgg_300.py`` (reference
:67-124 parameters, :180-231 wrappers, :306-380 anchors)."""
from . import ssd_vgg_512
from .ssd_vgg_512 import SSDParams, ssd_anchor_one_layer, ssd_anchors_all_layers  # noqa: F401  (same rule)


class SSDNet(ssd_vgg_512.SSDNet):
    """reference: nets/ssd_vgg_300.py:83-231.  Same pipeline as SSD-512 plus the clip after NMS
    (:225-229)."""
    default_params = SSDParams(
        img_shape=(300, 300),
        num_classes=21,
        no_annotation_label=21,
        feat_layers=['block4', 'block7', 'block8', 'block9', 'block10', 'block11'],
        feat_shapes=[(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)],
        anchor_size_bounds=[0.15, 0.90],
        anchor_sizes=[(21., 45.), (45., 99.), (99., 153.), (153., 207.), (207., 261.), (261., 315.)],
        anchor_ratios=[[2, .5], [2, .5, 3, 1. / 3], [2, .5, 3, 1. / 3], [2, .5, 3, 1. / 3], [2, .5], [2, .5]],
        anchor_steps=[8, 16, 32, 64, 100, 300],
        anchor_offset=0.5,
        normalizations=[20, -1, -1, -1, -1, -1],
        prior_scaling=[0.1, 0.1, 0.2, 0.2])
    _clip_after_nms = True
