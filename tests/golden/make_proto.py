"""Generates tests/golden/ref_model_proto.pb: a small model serialised by the protobuf RUNTIME with the reference's own
schema (the serialized FileDescriptorProto inside /root/reference/lopq/lopq/lopq_model_pb2.py:21; the generated module
itself cannot be imported by protobuf >= 4), filled in the order and with the value conversion of the reference's
LOPQModel.export_proto (model.py:748-786).  Run here (needs /root/reference); the fixture is committed."""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF_PB2 = "/root/reference/lopq/lopq/lopq_model_pb2.py"


def reference_message_class():
    from google.protobuf import descriptor_pool, message_factory
    src = open(REF_PB2).read()
    m = re.search(r"serialized_pb=_b\('(.*?)'\)\n", src, re.S)
    raw = eval("b'" + m.group(1) + "'")
    pool = descriptor_pool.DescriptorPool()
    pool.AddSerializedFile(raw)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("com.flickr.vision.lopq.LOPQModelParams"))


def small_params(seed=5, D=8, V=3, M=4, K=16):
    from tests.util import random_model_params
    return random_model_params(D, V, M, K, seed)


def serialize_like_reference(params, V, M, K):
    """model.py:748-786 with the runtime message class."""
    from itertools import chain
    Cs, Rs, mus, subs = params
    msg = reference_message_class()()
    msg.D = 2 * Cs[0].shape[1]
    msg.V, msg.M, msg.num_subquantizers = V, M, K
    for C in Cs:
        mm = msg.Cs.add(); mm.values.extend(map(float, np.nditer(C, order="C"))); mm.shape.extend(C.shape)
    for R in chain(*Rs):
        mm = msg.Rs.add(); mm.values.extend(map(float, np.nditer(R, order="C"))); mm.shape.extend(R.shape)
    for mu in chain(*mus):
        msg.mus.add().values.extend(map(float, np.nditer(mu, order="C")))
    for sub in chain(*subs):
        mm = msg.subs.add(); mm.values.extend(map(float, np.nditer(sub, order="C"))); mm.shape.extend(sub.shape)
    return msg.SerializeToString()


if __name__ == "__main__":
    buf = serialize_like_reference(small_params(), 3, 4, 16)
    with open(os.path.join(HERE, "ref_model_proto.pb"), "wb") as f:
        f.write(buf)
    print("wrote %d bytes" % len(buf))
