"""Hand-assembles tests/golden/h5py_like_keras.h5 from the HDF5 File Format Specification (version 1.1 structures, what
libhdf5 1.8/1.10 emits for h5py's default libver='earliest'), WITHOUT using depgan_b200.h5lite's writer.

Purpose ("harden h5lite against what h5py actually writes"): `load_weights` (EG:383, EU:402) has to read files
written by Keras 2.x `model.save` (TG:892, TU:622), and no h5py / libhdf5 exists here to produce one.  This script builds
a small full-model file containing the constructs h5lite's own writer never produces, so the reader is tested against
an independent encoder:

  * group B-trees with several leaves AND an internal (level 1) node; symbol nodes left half full, as after libhdf5's
    node splits; symbol-table entries of groups carry cache type 1 + scratch-pad (B-tree / heap addresses)
  * object headers whose messages overflow the first chunk into a continuation block (message 0x0010), NIL padding
    messages, modification-time (0x0012) and old / new fill-value messages (0x0004 / 0x0005)
  * `layer_names` split into `layer_names0` / `layer_names1` (Keras' 64 512-byte object-header limit work-around)
  * variable-length string attributes (`model_config`, `training_config`: Python str -> H5T_VARIABLE) living in a global
    heap collection (GCOL)
  * compact data layout (bias vectors stored inside the object header), contiguous layout elsewhere
  * TF-scope suffixes (`conv2d_a_1/kernel:0`, `bn_a_3/gamma:0`) and auto-numbered layers
  * a trailing `/optimizer_weights` group with nested `training/Adam/...` datasets (one int64 scalar)
  * local heaps with a free block, names not in insertion order

run from the repo root:  python tests/golden/make_h5py_like.py
"""
import json
import struct
from pathlib import Path

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 2  # group B-tree ranks stored in the superblock (libhdf5 defaults: 4 and 16)


def pad8(b):
    return b + b"\0" * (-len(b) % 8)


class Asm:
    def __init__(self):
        self.b = bytearray(b"\0" * 96)  # superblock v0 (56 bytes) + root symbol-table entry (40 bytes)

    def put(self, data, align=8):
        while len(self.b) % align:
            self.b.append(0)
        addr = len(self.b)
        self.b += data
        return addr

    # ---- messages -------------------------------------------------------------------------------------
    @staticmethod
    def msg(mtype, data, flags=0):
        data = pad8(data)
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    @staticmethod
    def space(shape):
        # dataspace message version 1: version, rank, flags, reserved(5), dims
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)

    @staticmethod
    def dt_f32():
        # class 1 (float) version 1; bit field: little endian, implied-1 mantissa normalisation, sign bit 31
        return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def dt_i64():
        return struct.pack("<BBBBI", 0x10, 0x08, 0x00, 0x00, 8) + struct.pack("<HH", 0, 64)

    @staticmethod
    def dt_str(n):
        return struct.pack("<BBBBI", 0x13, 0x00, 0x00, 0x00, n)

    @staticmethod
    def dt_vlen_str(utf8=False):
        # class 9 version 1: type = string (1), padding = null terminated (0), charset
        return struct.pack("<BBBBI", 0x19, 0x01, 0x01 if utf8 else 0x00, 0x00, 16) + Asm.dt_str(1)

    def attr(self, name, dt, sp, data):
        nm = name.encode() + b"\0"
        body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(sp)) + pad8(nm) + pad8(dt) + pad8(sp) + data
        return self.msg(0x000C, body)

    def attr_fixed_strings(self, name, strings):
        width = max([len(s) for s in strings] + [1])
        data = b"".join(s.encode().ljust(width, b"\0") for s in strings)
        return self.attr(name, self.dt_str(width), self.space((len(strings),)), data)

    def attr_scalar_str(self, name, s):
        return self.attr(name, self.dt_str(len(s)), self.space(()), s.encode())

    # ---- global heap (variable-length data) -----------------------------------------------------------
    def gcol(self, blobs):
        """One global heap collection holding `blobs`; returns (address, [object index of each blob])."""
        body = bytearray()
        idx = []
        for i, blob in enumerate(blobs, start=1):
            body += struct.pack("<HH4xQ", i, 1, len(blob)) + pad8(blob)
            idx.append(i)
        size = 16 + len(body) + 16
        size = max(4096, (size + 4095) & ~4095)
        free = size - 16 - len(body)
        body += struct.pack("<HH4xQ", 0, 0, free) + b"\0" * (free - 16)  # object 0 = the free space
        addr = self.put(b"GCOL" + struct.pack("<B3xQ", 1, size) + bytes(body))
        return addr, idx

    def attr_vlen_str(self, name, text, gaddr, gidx, utf8=False):
        data = struct.pack("<IQI", len(text.encode()), gaddr, gidx)
        return self.attr(name, self.dt_vlen_str(utf8), self.space(()), data)

    # ---- object headers -------------------------------------------------------------------------------
    def ohdr(self, msgs, first_chunk=None):
        """Version-1 object header.  With first_chunk (bytes) smaller than the messages, the rest goes into a
        continuation block referenced by a 0x0010 message, and the first chunk is filled up with a NIL message."""
        total = sum(len(m) for m in msgs)
        if first_chunk is None or total <= first_chunk:
            body = b"".join(msgs)
            return self.put(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)
        cont_msg_len = 8 + 16
        head, tail, used = [], [], cont_msg_len
        for m in msgs:
            if not tail and used + len(m) <= first_chunk:
                head.append(m)
                used += len(m)
            else:
                tail.append(m)
        nil = first_chunk - used
        extra = []
        if nil >= 8:
            extra.append(self.msg(0x0000, b"\0" * (nil - 8)))
        else:
            first_chunk = used
        tail_block = b"".join(tail) + self.msg(0x0000, b"\0" * 8)  # libhdf5 leaves NIL space at the end of a chunk too
        cont_addr = self.put(tail_block)
        cont = self.msg(0x0010, struct.pack("<QQ", cont_addr, len(tail_block)))
        body = b"".join(head) + cont + b"".join(extra)
        assert len(body) == first_chunk, (len(body), first_chunk)
        nmsgs = len(head) + 1 + len(extra) + len(tail) + 1
        return self.put(struct.pack("<BxHII4x", 1, nmsgs, 1, len(body)) + body)

    def dataset(self, arr, compact=False, mtime=1556000000):
        arr = np.asarray(arr, order="C")  # (ascontiguousarray would turn the 0-d iteration counter into a vector)
        dt = self.dt_f32() if arr.dtype == np.float32 else self.dt_i64()
        raw = arr.astype("<f4" if arr.dtype == np.float32 else "<i8").tobytes()
        msgs = [self.msg(0x0001, self.space(arr.shape)), self.msg(0x0003, dt, flags=1),
                self.msg(0x0004, struct.pack("<I", 0)),                       # old fill value: size 0
                self.msg(0x0005, struct.pack("<BBBB", 2, 2, 2, 0))]           # fill value v2: late alloc, undefined
        if compact:
            msgs.append(self.msg(0x0008, struct.pack("<BBH", 3, 0, len(raw)) + raw))
        else:
            addr = self.put(raw, align=1) if raw else UNDEF  # libhdf5 does not align raw data
            msgs.append(self.msg(0x0008, struct.pack("<BBQQ", 3, 1, addr, len(raw))))
        msgs.append(self.msg(0x0012, struct.pack("<B3xI", 1, mtime)))
        return self.ohdr(msgs)

    # ---- groups ---------------------------------------------------------------------------------------
    def group(self, children, attrs=(), first_chunk=None, fill=5):
        """children: {name: (object header address, (btree, heap) or None)}; `fill` names per symbol node."""
        names = sorted(children, key=lambda s: s.encode())
        # local heap: offset 0 = "", then the names in REVERSE order (offsets need not be sorted), then a free block
        seg = bytearray(8)
        offs = {}
        for n in reversed(names):
            offs[n] = len(seg)
            seg += pad8(n.encode() + b"\0")
        free_off = len(seg)
        seg += struct.pack("<QQ", 1, 32) + b"\0" * 16  # free block: next = 1 (none), size 32
        seg_addr = self.put(bytes(seg))
        heap_addr = self.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, seg_addr))
        # leaves
        leaves = []  # (address, offset of the largest name)
        for i in range(0, max(len(names), 1), fill):
            chunk = names[i:i + fill]
            node = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(chunk)))
            for n in chunk:
                addr, grp = children[n]
                if grp:
                    node += struct.pack("<QQII", offs[n], addr, 1, 0) + struct.pack("<QQ", *grp)
                else:
                    node += struct.pack("<QQII16x", offs[n], addr, 0, 0)
            node += b"\0" * (8 + 40 * 2 * LEAF_K - len(node))
            leaves.append((self.put(bytes(node)), offs[chunk[-1]] if chunk else 0))

        def tree(kids, level):
            """kids: [(child address, key = heap offset of the largest name below)] -> [(node address, key)]"""
            cap = 2 * INTERNAL_K
            out = []
            per = cap if len(kids) <= cap else (cap + 1) // 2 + 1  # split nodes are left partly filled
            for i in range(0, len(kids), per):
                part = kids[i:i + per]
                node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, level, len(part), UNDEF, UNDEF))
                node += struct.pack("<Q", 0)  # key 0: the empty string at heap offset 0 (libhdf5 writes the left bound)
                for addr, key in part:
                    node += struct.pack("<QQ", addr, key)
                node += b"\0" * (24 + 8 + 16 * cap - len(node))
                out.append((self.put(bytes(node)), part[-1][1]))
            return out

        level, nodes = 0, tree(leaves, 0)
        while len(nodes) > 1:
            level += 1
            nodes = tree(nodes, level)
        btree = nodes[0][0]
        msgs = [self.msg(0x0011, struct.pack("<QQ", btree, heap_addr))] + list(attrs)
        return self.ohdr(msgs, first_chunk), (btree, heap_addr), level

    def finish(self, root_addr, root_grp):
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.b), UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", *root_grp)
        assert len(sb) == 96
        self.b[:96] = sb
        return bytes(self.b)


def content():
    """The logical content of the fixture (also what tests/test_h5lite.py expects to read back)."""
    rng = np.random.default_rng(20190530)
    layers = []  # (layer name, tf scope or None, [(weight, array)])
    layers.append(("input_gen_chn_0", None, []))
    for i, (name, scope, cin, cout) in enumerate([("conv2d_gen_0", "conv2d_gen_0_1", 1, 4), ("conv2d_gen_1", "conv2d_gen_1_1", 4, 4),
                                                  ("conv2d_gen_2", "conv2d_gen_2_3", 4, 6)]):
        layers.append((name, scope, [("kernel", rng.standard_normal((3, 3, cin, cout)).astype(np.float32)),
                                     ("bias", rng.standard_normal((cout,)).astype(np.float32))]))
        bn = name.replace("conv2d", "bn")
        layers.append((bn, bn + scope[len(name):], [(w, rng.standard_normal((cout,)).astype(np.float32))
                                                    for w in ("gamma", "beta", "moving_mean", "moving_variance")]))
        layers.append((name.replace("conv2d", "relu"), None, []))
        layers.append(("do_gen_%d" % i, None, []))
    layers.append(("dense_7", "dense_7", [("kernel", rng.standard_normal((6, 1)).astype(np.float32)),
                                          ("bias", np.array([0.25], np.float32))]))
    for i in range(12):  # weight-less layers: enough names for several symbol nodes and an internal B-tree node
        layers.append(("%s_%d" % (("add_noiseZ", "mul_noiseZ", "maxpool2d_gen", "concat_gen")[i % 4], i), None, []))
    layers.append(("non_lin_segment", None, []))
    model_config = json.dumps({"class_name": "Model", "config": {"name": "Gen_UNet2D", "layers": [
        {"name": n, "class_name": "Layer", "config": {"name": n}, "inbound_nodes": []} for n, _, _ in layers]}})
    training_config = json.dumps({"optimizer_config": {"class_name": "Adam", "config": {"lr": 1e-4, "beta_1": 0.9}},
                                  "loss": "categorical_crossentropy", "metrics": ["accuracy"]})
    opt = [("training/Adam/iterations:0", np.array(1234, np.int64)),
           ("training/Adam/Variable:0", rng.standard_normal((3, 3, 1, 4)).astype(np.float32)),
           ("training/Adam/Variable_1:0", rng.standard_normal((4,)).astype(np.float32))]
    return layers, model_config, training_config, opt


def build():
    layers, model_config, training_config, opt = content()
    a = Asm()
    layer_objs = {}
    for name, scope, weights in layers:
        wn = ["%s/%s:0" % (scope, w) for w, _ in weights]
        attrs = [a.attr_fixed_strings("weight_names", wn)]
        kids = {}
        if weights:
            ds = {"%s:0" % w: (a.dataset(arr, compact=(w in ("bias", "beta"))), None) for w, arr in weights}
            saddr, sgrp, _ = a.group(ds, fill=3)
            kids[scope] = (saddr, sgrp)
        addr, grp, _ = a.group(kids, attrs)
        layer_objs[name] = (addr, grp)
    names = [n for n, _, _ in layers]
    half = len(names) // 2 + 3
    mw_attrs = [a.attr_fixed_strings("layer_names0", names[:half]), a.attr_fixed_strings("layer_names1", names[half:]),
                a.attr_scalar_str("backend", "tensorflow"), a.attr_scalar_str("keras_version", "2.2.4")]
    mw_addr, mw_grp, mw_level = a.group(layer_objs, mw_attrs, first_chunk=200, fill=5)
    assert mw_level >= 1, "the fixture must contain an internal B-tree node"
    # /optimizer_weights/training/Adam/<name>
    adam = {n.split("/")[-1]: (a.dataset(arr), None) for n, arr in opt}
    adam_addr, adam_grp, _ = a.group(adam)
    tr_addr, tr_grp, _ = a.group({"Adam": (adam_addr, adam_grp)})
    ow_addr, ow_grp, _ = a.group({"training": (tr_addr, tr_grp)}, [a.attr_fixed_strings("weight_names", [n for n, _ in opt])])
    gaddr, gidx = a.gcol([model_config.encode(), training_config.encode()])
    root_attrs = [a.attr_scalar_str("keras_version", "2.2.4"), a.attr_scalar_str("backend", "tensorflow"),
                  a.attr_vlen_str("model_config", model_config, gaddr, gidx[0]),
                  a.attr_vlen_str("training_config", training_config, gaddr, gidx[1], utf8=True)]
    root_addr, root_grp, _ = a.group({"model_weights": (mw_addr, mw_grp), "optimizer_weights": (ow_addr, ow_grp)},
                                     root_attrs, first_chunk=120)
    return a.finish(root_addr, root_grp)


if __name__ == "__main__":
    out = Path(__file__).resolve().parent / "h5py_like_keras.h5"
    out.write_bytes(build())
    print("wrote", out, out.stat().st_size, "bytes")
