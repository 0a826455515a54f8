"""ron_tensorflow_b200 -- B200-native (sm_100a) implementation of RON_Tensorflow's detection
hot path: all-layer joint GT->anchor matching + target encoding, and the inference
post-process (decode, objectness gate, class-wise top-k, NMS, VOC TP/FP), behind the
reference's own Python entry points (``nets.ron_vgg_320``, ``nets.ssd_common``,
``nets.ssd_vgg_300/512``, ``tf_extended``).  Python host code -> ctypes -> libronk.so
(hand-written CUDA).  No CPU fallback."""
__version__ = '0.1.0'
