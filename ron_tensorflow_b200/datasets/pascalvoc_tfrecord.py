"""Reader for the Pascal VOC TFRecords of the reference: the files its
``datasets/pascalvoc_to_tfrecords.py`` writes and its ``datasets/pascalvoc_common.py:96-121`` decodes --
turned into the ground-truth tensors the match + encode path takes.  Pure Python + NumPy host code (the
reference reads them with TF-slim queue runners on the CPU too); no TensorFlow, no protobuf package.

Formats, both public and stable:
* TFRecord framing: uint64 length, uint32 masked CRC-32C of the length, payload, uint32 masked CRC-32C of
  the payload (little endian; mask(c) = ((c >> 15) | (c << 17)) + 0xa282ead8).
* payload = ``tf.train.Example``: Example{features=1}, Features{map<string, Feature> feature=1} (entries:
  key=1, value=2), Feature{bytes_list=1 | float_list=2 | int64_list=3}, each list {repeated value=1},
  floats / int64s packed or not.

Schema (pascalvoc_common.py:96-109, pascalvoc_to_tfrecords.py:154-168): image/shape int64[3] (h, w, c),
image/object/bbox/{ymin,xmin,ymax,xmax} float (already divided by the image size), .../label int64 (1..20),
.../difficult, .../truncated int64, image/encoded + image/format bytes.
"""
import struct

import numpy as np

__all__ = ['crc32c', 'masked_crc32c', 'read_records', 'parse_example', 'read_voc_tfrecords', 'gt_batch']


def _make_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (0x82f63b78 if c & 1 else 0)
        tab.append(c)
    return tab


_TABLE = _make_table()


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), the checksum of the TFRecord format."""
    c = crc ^ 0xffffffff
    for b in data:
        c = _TABLE[(c ^ b) & 0xff] ^ (c >> 8)
    return c ^ 0xffffffff


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff


def read_records(path, check_crc=True):
    """Yields the payload of every record of one TFRecord file."""
    with open(path, 'rb') as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) != 12:
                raise ValueError('%s: truncated record header' % path)
            (length,), (lcrc,) = struct.unpack('<Q', head[:8]), struct.unpack('<I', head[8:])
            if check_crc and masked_crc32c(head[:8]) != lcrc:
                raise ValueError('%s: corrupted record length' % path)
            data = f.read(length)
            tail = f.read(4)
            if len(data) != length or len(tail) != 4:
                raise ValueError('%s: truncated record' % path)
            if check_crc and masked_crc32c(data) != struct.unpack('<I', tail)[0]:
                raise ValueError('%s: corrupted record payload' % path)
            yield data


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7f) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """(field number, wire type, value) of one message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield field, wt, v


def _feature(buf):
    """Feature -> list of bytes | float32 array | int64 array."""
    for field, wt, v in _fields(buf):
        if wt != 2:
            continue
        if field == 1:                                           # BytesList
            return [bytes(x) for f, w, x in _fields(v) if f == 1 and w == 2]
        if field == 2:                                           # FloatList: packed (wire type 2) or repeated fixed32
            vals = []
            for f, w, x in _fields(v):
                if f == 1:
                    vals.append(np.frombuffer(bytes(x), '<f4'))
            return np.concatenate(vals).astype(np.float32) if vals else np.zeros(0, np.float32)
        if field == 3:                                           # Int64List: packed varints or repeated varints
            vals = []
            for f, w, x in _fields(v):
                if f != 1:
                    continue
                if w == 0:
                    vals.append(x)
                else:
                    p = 0
                    while p < len(x):
                        u, p = _varint(x, p)
                        vals.append(u)
            a = np.array(vals, np.uint64) if vals else np.zeros(0, np.uint64)
            return a.astype(np.int64)                            # two's complement for negative values
    return []


def parse_example(payload):
    """Serialized tf.train.Example -> {feature name: list of bytes | float32 array | int64 array}."""
    out = {}
    buf = memoryview(payload)
    for field, wt, feats in _fields(buf):
        if field != 1 or wt != 2:
            continue
        for f, w, entry in _fields(feats):
            if f != 1 or w != 2:
                continue
            key, val = None, None
            for ef, ew, ev in _fields(entry):
                if ef == 1 and ew == 2:
                    key = bytes(ev).decode('utf-8')
                elif ef == 2 and ew == 2:
                    val = _feature(ev)
            if key is not None:
                out[key] = val if val is not None else []
    return out


def read_voc_tfrecords(paths, with_image=False):
    """Yields one dict per image, the items of pascalvoc_common.py:110-118: shape int64[3], object/bbox float32
    [G,4] (ymin, xmin, ymax, xmax), object/label, object/difficult, object/truncated int64 [G] (and image, format)."""
    for path in ([paths] if isinstance(paths, str) else paths):
        for payload in read_records(path):
            ex = parse_example(payload)
            box = [np.asarray(ex.get('image/object/bbox/' + k, np.zeros(0, np.float32)), np.float32)
                   for k in ('ymin', 'xmin', 'ymax', 'xmax')]
            rec = {'shape': np.asarray(ex['image/shape'], np.int64),
                   'object/bbox': np.stack(box, -1) if box[0].size else np.zeros((0, 4), np.float32)}
            for k in ('label', 'difficult', 'truncated'):
                rec['object/' + k] = np.asarray(ex.get('image/object/bbox/' + k, np.zeros(0, np.int64)), np.int64)
            if with_image:
                rec['image'] = (ex.get('image/encoded') or [b''])[0]
                rec['format'] = (ex.get('image/format') or [b'jpeg'])[0]
            yield rec


def gt_batch(records, g_max=None):
    """Padded ground truth of a list of records, as RONNet.bboxes_encode_batch / bboxes_matching_batch take it:
    boxes float32 [B,Gmax,4], labels int64 [B,Gmax], difficults int64 [B,Gmax], counts int32 [B]."""
    records = list(records)
    counts = np.array([r['object/label'].shape[0] for r in records], np.int32)
    g_max = int(g_max or max(int(counts.max()) if counts.size else 1, 1))
    B = len(records)
    boxes = np.zeros((B, g_max, 4), np.float32)
    labels = np.zeros((B, g_max), np.int64)
    diff = np.zeros((B, g_max), np.int64)
    for b, r in enumerate(records):
        g = min(int(counts[b]), g_max)
        boxes[b, :g], labels[b, :g], diff[b, :g] = r['object/bbox'][:g], r['object/label'][:g], r['object/difficult'][:g]
    return boxes, labels, diff, np.minimum(counts, g_max).astype(np.int32)
