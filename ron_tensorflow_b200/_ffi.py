"""ctypes binding of libronk.so (include/ronk.h).  No fallback: a missing library or a
failing call raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libronk.so')

RONK_OK, RONK_EINVAL, RONK_ECUDA, RONK_ENOMEM, RONK_ELIMIT = 0, -1, -2, -3, -4
KIND_RON, KIND_SSD = 0, 1
NMS_MIN, NMS_UNION = 0, 1
MATCH_NO_IGNORE_BETWEEN, MATCH_NO_GT_MAX_FIRST = 1, 2
SELECT_LOC_DECODED, SELECT_NO_SAMPLING, SELECT_TEST_REBUILD = 1, 2, 4

c_void_p, c_int, c_float, c_double, c_size_t, c_longlong = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t, ctypes.c_longlong)
P = ctypes.POINTER

# name -> (restype, argtypes); the same list is what tests/test_abi.py checks against include/ronk.h
SIGNATURES = {
    'ronk_version': (c_int, []),
    'ronk_last_error': (ctypes.c_char_p, []),
    'ronk_launch_count': (c_longlong, []),
    'ronk_stream_capture_id': (c_int, [c_void_p, P(ctypes.c_ulonglong)]),
    'ronk_anchors_create': (c_int, [c_int, c_int, c_int, c_int, P(c_int), P(c_double), P(c_int), P(c_double),
                                    P(c_int), P(c_double), c_double, P(c_int), P(c_void_p)]),
    'ronk_anchors_create_flat': (c_int, [c_int, c_int, c_int, P(c_float), P(c_int), P(c_void_p)]),
    'ronk_anchors_destroy': (None, [c_void_p]),
    'ronk_anchors_num': (c_int, [c_void_p]),
    'ronk_anchors_num_layers': (c_int, [c_void_p]),
    'ronk_anchors_layer_info': (c_int, [c_void_p, c_int, P(c_int), P(c_int), P(c_int), P(c_int)]),
    'ronk_anchors_table': (c_void_p, [c_void_p, c_int]),
    'ronk_anchors_layer_hw': (c_int, [c_void_p, c_int, P(c_float), P(c_float)]),
    'ronk_encode_workspace_bytes': (c_size_t, [c_int, c_int]),
    'ronk_encode_workspace_init': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'ronk_match_encode': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float,
                                  P(c_float), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    'ronk_decode': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, P(c_float), c_void_p, c_void_p]),
    'ronk_select_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int]),
    'ronk_decode_select_topk': (c_int, [c_void_p, P(c_void_p), P(c_void_p), P(c_void_p), c_int, c_int, c_float,
                                        c_float, P(c_float), c_float, P(c_float), c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    'ronk_select_topk_flagged': (c_int, [c_void_p, P(c_void_p), P(c_void_p), P(c_void_p), c_int, c_int, c_float,
                                         c_float, P(c_float), c_float, P(c_float), c_int, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p]),
    'ronk_nms_batch_tiered': (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    'ronk_sort_topk': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ronk_clip': (c_int, [P(c_float), c_void_p, c_longlong, c_void_p, c_void_p]),
    'ronk_nms_workspace_bytes': (c_size_t, [c_int, c_int]),
    'ronk_nms_batch': (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p]),
    'ronk_areas': (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    'ronk_pairwise': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'ronk_overlap_ref': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'ronk_select_mask': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                 c_void_p]),
    'ronk_select_all_classes': (c_int, [c_void_p, c_longlong, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    'ronk_gather_i64': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ronk_dual_max_match_workspace_bytes': (c_size_t, [c_int]),
    'ronk_dual_max_match': (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    'ronk_flaten_predict': (c_int, [P(c_void_p), P(c_void_p), P(c_int), c_int, c_int, c_float, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    'ronk_filter_boxes_mask': (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p]),
    'ronk_minsize_mask': (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p]),
    'ronk_filter_min_count': (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    'ronk_filter_min_write': (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    'ronk_rowmax_mask': (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    'ronk_compact_workspace_bytes': (c_size_t, [c_int]),
    'ronk_compact_indices': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ronk_gather_rows': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'ronk_bboxes_resize': (c_int, [P(c_float), c_void_p, c_longlong, c_void_p, c_void_p]),
    'ronk_class_columns': (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    'ronk_keep_by_class': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_float, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    'ronk_group_by_label': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ronk_mark_positions': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ronk_np_select_mask': (c_int, [c_void_p, c_longlong, c_int, c_int, c_float, c_void_p, c_void_p]),
    'ronk_np_select_gather': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    'ronk_np_clip': (c_int, [P(c_float), c_void_p, c_longlong, c_void_p, c_void_p]),
    'ronk_np_nms': (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    'ronk_voc_match': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p,
                               c_void_p]),
    'ronk_tpfp_records_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'ronk_tpfp_records_append': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int,
                                         c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ronk_sort_rows_workspace_bytes': (c_size_t, [c_int, c_int]),
    'ronk_sort_rows': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ronk_average_precision_workspace_bytes': (c_size_t, [c_longlong, c_int]),
    'ronk_average_precision_records': (c_int, [c_void_p, c_longlong, c_void_p, c_int, P(c_double), c_int, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ronk_sparse_rows_packet_bytes': (c_size_t, [c_int]),
    'ronk_sparse_rows_pack': (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    'ronk_host_rows_apply': (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    'ronk_sparse_labels_packet_bytes': (c_size_t, [c_int]),
    'ronk_sparse_labels_pack': (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    'ronk_host_targets_apply': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]),
    'ronk_loss_workspace_bytes': (c_size_t, []),
    'ronk_loss_workspace_init': (c_int, [c_void_p, c_void_p]),
    'ronk_loss_masks': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_float, c_float, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_float, c_void_p, c_void_p,
                                c_void_p]),
    'ronk_smooth_l1': (c_int, [c_void_p, c_void_p, c_longlong, c_float, c_float, c_double, c_void_p, c_void_p]),
    'ronk_smooth_l1_backward': (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_float, c_float, c_double, c_void_p, c_void_p]),
    'ronk_localization_loss_backward': (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_double, c_float, c_void_p, c_void_p,
                                                c_void_p, c_void_p]),
    'ronk_localization_loss': (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_double, c_float, c_void_p, c_void_p,
                                       c_void_p]),
    'ronk_tpfp_match': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """Load libronk.so (once).  Raises RuntimeError when the CUDA extension was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('%s is missing: build the CUDA extension first '
                               '(python -m ron_tensorflow_b200.build); there is no CPU fallback' % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.ronk_version() != 100:
            raise RuntimeError('libronk.so version mismatch: rebuild with python -m ron_tensorflow_b200.build')
        _lib = l
    return _lib


def last_error():
    return lib().ronk_last_error().decode('utf-8', 'replace')


def check(rc):
    """Map RONK_E* to the exceptions the reference's callers would see: ValueError for bad
    arguments (graph-build time errors in the reference), RuntimeError for runtime failures."""
    if rc == RONK_OK:
        return
    msg = last_error()
    if rc in (RONK_EINVAL, RONK_ELIMIT):
        raise ValueError(msg)
    if rc == RONK_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def farr(values):
    return (c_float * len(values))(*[float(v) for v in values])


def iarr(values):
    return (c_int * len(values))(*[int(v) for v in values])


def darr(values):
    return (c_double * len(values))(*[float(v) for v in values])
