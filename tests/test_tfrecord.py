"""Pascal VOC TFRecords: tests/golden/voc_synth_000.tfrecord was written by the
reference's own datasets/pascalvoc_to_tfrecords.py (unmodified, over the shim's tf.train.Example /
TFRecordWriter) from a synthetic VOC tree; the reader must return what that converter put in, must agree with
the official protobuf runtime on the wire format, and its ground truth must drive the encode kernel."""
import os
import struct

import numpy as np
import pytest

from ron_tensorflow_b200 import synth
from ron_tensorflow_b200.datasets import pascalvoc_tfrecord as R
from _util import need_cuda, eq

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, 'golden', 'voc_synth_000.tfrecord')
SEED, N = 515, 12


def test_crc32c_known_answers():
    assert R.crc32c(b'123456789') == 0xe3069283            # the CRC-32C check value
    assert R.crc32c(b'') == 0 and R.crc32c(bytes(32)) == 0x8a9136aa      # RFC 3720 B.4: 32 bytes of zeros
    assert R.masked_crc32c(b'123456789') == ((((0xe3069283 >> 15) | (0xe3069283 << 17)) + 0xa282ead8) & 0xffffffff)


def test_reader_returns_what_the_reference_converter_wrote():
    ids, annots, _ = synth.make_voc_eval_case(SEED, N)
    recs = list(R.read_voc_tfrecords(FIXTURE, with_image=True))
    assert len(recs) == N
    label_of = {c: i + 1 for i, c in enumerate(synth.VOC_CLASSES)}          # pascalvoc_common.py:24-46
    for i, objs, r in zip(ids, annots, recs):                               # sorted file names == id order
        assert r['shape'].tolist() == [375, 500, 3]
        # pascalvoc_to_tfrecords.py:118-122: float(text) / shape, python doubles stored as float32
        want = np.array([[o['bbox'][1] / 375, o['bbox'][0] / 500, o['bbox'][3] / 375, o['bbox'][2] / 500] for o in objs],
                        np.float64).astype(np.float32).reshape(-1, 4)
        assert np.array_equal(r['object/bbox'], want)
        assert r['object/label'].tolist() == [label_of[o['name']] for o in objs]
        assert r['object/difficult'].tolist() == [o['difficult'] for o in objs]
        assert r['object/truncated'].tolist() == [0] * len(objs)
        assert r['image'] == b'\xff\xd8 synthetic ' + i.encode('ascii') + b' \xff\xd9' and r['format'] == b'JPEG'
    boxes, labels, diff, counts = R.gt_batch(recs)
    assert boxes.shape == (N, counts.max(), 4) and labels.dtype == np.int64 and counts.dtype == np.int32
    assert all(np.array_equal(boxes[b, :counts[b]], recs[b]['object/bbox']) and not boxes[b, counts[b]:].any() for b in range(N))


def _official_example_class():
    """tf.train.Example rebuilt with the official protobuf runtime from its public definition (example.proto,
    feature.proto)."""
    pytest.importorskip('google.protobuf')
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name='ronk_example.proto', package='ronk_tf', syntax='proto3')
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name, *fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = '.ronk_tf.' + tname
        return m
    msg('BytesList', ('value', 1, T.TYPE_BYTES, T.LABEL_REPEATED, None))
    msg('FloatList', ('value', 1, T.TYPE_FLOAT, T.LABEL_REPEATED, None))
    msg('Int64List', ('value', 1, T.TYPE_INT64, T.LABEL_REPEATED, None))
    feat = msg('Feature', ('bytes_list', 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, 'BytesList'),
               ('float_list', 2, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, 'FloatList'),
               ('int64_list', 3, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, 'Int64List'))
    feat.oneof_decl.add(name='kind')
    for f in feat.field:
        f.oneof_index = 0
    feats = msg('Features', ('feature', 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, 'Features.FeatureEntry'))
    entry = feats.nested_type.add(name='FeatureEntry')
    entry.options.map_entry = True
    entry.field.add(name='key', number=1, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    entry.field.add(name='value', number=2, type=T.TYPE_MESSAGE, label=T.LABEL_OPTIONAL, type_name='.ronk_tf.Feature')
    msg('Example', ('features', 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, 'Features'))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName('ronk_tf.Example'))


def test_wire_format_agrees_with_the_official_protobuf_runtime():
    Example = _official_example_class()
    n = 0
    for payload in R.read_records(FIXTURE):
        ex = Example.FromString(payload)                       # what the shim's writer produced is a valid Example
        mine = R.parse_example(payload)
        assert set(mine) == set(ex.features.feature)
        for k, f in ex.features.feature.items():
            kind = f.WhichOneof('kind')
            if kind == 'bytes_list':
                assert mine[k] == list(f.bytes_list.value)
            elif kind == 'float_list':
                assert np.array_equal(mine[k], np.array(f.float_list.value, np.float32))
            else:
                assert np.array_equal(mine[k], np.array(f.int64_list.value, np.int64))
        n += 1
    assert n == N
    # and the other direction: bytes from the official serializer (negative int64, empty lists, many values)
    ex = Example()
    ex.features.feature['a'].int64_list.value.extend([0, 1, -1, 2 ** 62, -2 ** 63, 300])
    ex.features.feature['b'].float_list.value.extend([0.5, -1.25, 3e38])
    ex.features.feature['c'].bytes_list.value.extend([b'', b'xyz', bytes(range(256))])
    ex.features.feature['d'].float_list.SetInParent()
    got = R.parse_example(ex.SerializeToString())
    assert got['a'].tolist() == [0, 1, -1, 2 ** 62, -2 ** 63, 300]
    assert np.array_equal(got['b'], np.array([0.5, -1.25, 3e38], np.float32))
    assert got['c'] == [b'', b'xyz', bytes(range(256))] and got['d'].shape == (0,)


def test_corruption_is_detected(tmp_path):
    data = bytearray(open(FIXTURE, 'rb').read())
    data[40] ^= 0x01
    bad = tmp_path / 'bad.tfrecord'
    bad.write_bytes(bytes(data))
    with pytest.raises(ValueError):
        list(R.read_records(str(bad)))
    (tmp_path / 'short.tfrecord').write_bytes(bytes(data[:20]))
    with pytest.raises(ValueError):
        list(R.read_records(str(tmp_path / 'short.tfrecord')))
    assert len(list(R.read_records(str(bad), check_crc=False))) == N


@pytest.mark.gpu
def test_tfrecord_ground_truth_drives_the_encode_kernel():
    """TFRecord -> padded GT batch -> match + encode on the GPU, against the oracle image by image."""
    need_cuda()
    from oracle import ron_oracle as O
    from ron_tensorflow_b200.nets import ron_vgg_320
    recs = list(R.read_voc_tfrecords(FIXTURE))
    boxes, labels, diff, counts = R.gt_batch(recs)
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    t = net.bboxes_encode_batch(labels, boxes, counts, anchors, positive_threshold=0.56, ignore_threshold=0.3)
    enc, cor, inside = O.encode_anchor_tables(O.anchors_all_layers(O.RON320), O.RON320.img_shape, O.RON320.allowed_borders)
    for b in range(len(recs)):
        r = O.encode_image(labels[b, :counts[b]], boxes[b, :counts[b]], enc, cor, inside, 0.56, 0.3)
        eq(t['labels'][b], r['labels'], 'labels'); eq(t['loc'][b], r['loc'], 'loc'); eq(t['scores'][b], r['scores'], 'scores')


def test_parser_fuzz_against_the_official_runtime():
    """Random Examples (all three list kinds, empty and long lists, extreme values) serialised by the official
    protobuf runtime parse to the same content; a record file written from them reads back in order."""
    hyp = pytest.importorskip('hypothesis')
    from hypothesis import given, settings, strategies as st
    Example = _official_example_class()
    i64 = st.integers(min_value=-2 ** 63, max_value=2 ** 63 - 1)
    f32 = st.floats(width=32, allow_nan=False)
    feature = st.one_of(st.tuples(st.just('i'), st.lists(i64, max_size=40)), st.tuples(st.just('f'), st.lists(f32, max_size=40)),
                        st.tuples(st.just('b'), st.lists(st.binary(max_size=30), max_size=6)))

    @settings(max_examples=150, deadline=None)
    @given(st.dictionaries(st.text(min_size=1, max_size=12), feature, max_size=8))
    def check(feats):
        ex = Example()
        for k, (kind, vals) in feats.items():
            f = ex.features.feature[k]
            if kind == 'i':
                f.int64_list.value.extend(vals); f.int64_list.SetInParent()
            elif kind == 'f':
                f.float_list.value.extend(vals); f.float_list.SetInParent()
            else:
                f.bytes_list.value.extend(vals); f.bytes_list.SetInParent()
        got = R.parse_example(ex.SerializeToString())
        assert set(got) == set(feats)
        for k, (kind, vals) in feats.items():
            if kind == 'i':
                assert got[k].tolist() == vals
            elif kind == 'f':
                assert np.array_equal(np.asarray(got[k], np.float32), np.array(vals, np.float32))
            else:
                assert got[k] == vals
    check()
