"""Wire codec for the reference's model protobuf (lopq/lopq/lopq_model_pb2.py:21, schema `lopq_model.proto`):

    message Vector { repeated float values = 1 [packed = true]; }
    message Matrix { repeated float values = 1 [packed = true]; repeated uint32 shape = 2; }
    message LOPQModelParams { optional uint32 D = 1, V = 2, M = 3, num_subquantizers = 4;
                              repeated Matrix Cs = 5, Rs = 6; repeated Vector mus = 7; repeated Matrix subs = 8; }

Written against the protobuf encoding itself (varints, length-delimited fields, little-endian packed float32), so models
move in and out without the generated module (which modern protobuf runtimes refuse to import) and without a Python
loop over values: the float payloads go through NumPy buffers.  Serialisation is byte-identical to the protobuf
runtime's (fields in number order, `shape` unpacked as proto2 does); the parser also accepts packed `shape`.
"""
import numpy as np


def _varint(n):
    n = int(n)
    if n < 0:
        raise ValueError("negative varint")
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = result = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _floats(a):
    return np.ascontiguousarray(np.asarray(a), dtype="<f4").tobytes()


def encode_vector(a):
    p = _floats(np.ravel(a))
    return _ld(1, p) if p else b""


def encode_matrix(a):
    a = np.asarray(a)
    p = _floats(a)                                   # C order, as np.nditer(a, order='C') (model.py:764)
    out = _ld(1, p) if p else b""
    for s in a.shape:
        out += _varint((2 << 3) | 0) + _varint(s)
    return out


def encode_model(D, V, M, num_subquantizers, Cs=None, Rs=None, mus=None, subs=None):
    """export_proto (model.py:748-786): Rs / mus are flattened over (split, cluster), subs over (split, sub-quantizer)."""
    out = b""
    for f, v in ((1, D), (2, V), (3, M), (4, num_subquantizers)):
        out += _varint((f << 3) | 0) + _varint(v)
    if Cs is not None:
        for C in Cs:
            out += _ld(5, encode_matrix(C))
    if Rs is not None:
        for half in Rs:
            for R in half:
                out += _ld(6, encode_matrix(R))
    if mus is not None:
        for half in mus:
            for mu in half:
                out += _ld(7, encode_vector(mu))
    if subs is not None:
        for half in subs:
            for sub in half:
                out += _ld(8, encode_matrix(sub))
    return out


def _fields(buf):
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _read_varint(buf, pos)
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            if pos + ln > n:
                raise ValueError("truncated field %d" % field)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield field, wt, val


def _decode_array(buf, with_shape):
    chunks, shape = [], []
    for field, wt, val in _fields(buf):
        if field == 1 and wt == 2:
            chunks.append(np.frombuffer(val, dtype="<f4"))
        elif field == 1 and wt == 5:
            chunks.append(np.frombuffer(val, dtype="<f4"))
        elif field == 2 and wt == 0:
            shape.append(val)
        elif field == 2 and wt == 2:                 # packed shape
            p = 0
            while p < len(val):
                v, p = _read_varint(val, p)
                shape.append(v)
    values = np.concatenate(chunks).astype(np.float64) if chunks else np.zeros(0)
    if with_shape:
        return values.reshape(shape)                  # np.reshape(C.values, C.shape), model.py:804
    return values


def decode_model(buf):
    """Parsed LOPQModelParams as a dict: scalars D, V, M, num_subquantizers and the lists Cs, Rs, mus, subs of float64
    arrays (float32 values, as the protobuf runtime hands them to NumPy)."""
    buf = memoryview(bytes(buf))
    out = {"D": 0, "V": 0, "M": 0, "num_subquantizers": 0, "Cs": [], "Rs": [], "mus": [], "subs": []}
    names = {1: "D", 2: "V", 3: "M", 4: "num_subquantizers"}
    for field, wt, val in _fields(buf):
        if field in names and wt == 0:
            out[names[field]] = int(val)
        elif field == 5 and wt == 2:
            out["Cs"].append(_decode_array(val, True))
        elif field == 6 and wt == 2:
            out["Rs"].append(_decode_array(val, True))
        elif field == 7 and wt == 2:
            out["mus"].append(_decode_array(val, False))
        elif field == 8 and wt == 2:
            out["subs"].append(_decode_array(val, True))
    return out
