"""Library-free reader/writer for the HDF5 subset that Keras 2.x weight files use.

There is no h5py / libhdf5 in this environment, and the drop-in surface has to read the reference's
``netG_*.h5`` / ``trained_depuresnet_*.h5`` files unchanged (``load_weights`` EG:383, EU:402; written by
``model.save`` TG:892, TU:622).  Those files are "old-style" HDF5 as h5py emits with libver='earliest':

  superblock v0 -> root symbol-table entry -> object headers v1 (with continuation blocks)
  groups   = symbol-table message (B-tree v1 'TREE' + local heap 'HEAP' + symbol nodes 'SNOD')
  datasets = dataspace + datatype (IEEE float / fixed strings) + data layout v3 contiguous (or compact)
  attributes = attribute messages v1/v2/v3 (``layer_names``, ``weight_names`` as fixed-length byte strings;
               variable-length ``model_config`` / ``training_config`` are skipped)

Unsupported constructs (chunked / filtered datasets, new-style groups, dense attribute storage) raise a clear
error instead of returning garbage.  The writer emits the same structures so a real h5py can read the result.
"""
from __future__ import annotations

import re
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


# ==========================================================================================================
# reader
# ==========================================================================================================
class _Obj:
    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.msgs = f._read_header(addr)
        self._children = None

    # ---- classification ----
    def is_group(self):
        return any(t == 0x0011 for t, _ in self.msgs)

    def is_dataset(self):
        return any(t == 0x0008 for t, _ in self.msgs)

    # ---- groups ----
    def children(self):
        if self._children is None:
            if any(t in (0x0002, 0x0006) for t, _ in self.msgs) and not self.is_group():
                raise H5Error("new-style (link-message) groups are not supported")
            out = {}
            for t, d in self.msgs:
                if t == 0x0011:
                    btree, heap = struct.unpack_from("<QQ", d, 0)
                    self.f._walk_btree(btree, self.f._heap_data(heap), out)
            self._children = out
        return self._children

    def __contains__(self, name):
        return name.split("/")[0] in self.children() and (("/" not in name) or name.split("/", 1)[1] in self[name.split("/")[0]])

    def __getitem__(self, path):
        obj = self
        for part in [p for p in path.split("/") if p]:
            ch = obj.children()
            if part not in ch:
                raise KeyError(path)
            obj = _Obj(self.f, ch[part])
        return obj

    def keys(self):
        return list(self.children().keys())

    # ---- attributes ----
    def attrs(self):
        out = {}
        for t, d in self.msgs:
            if t == 0x0015:
                raise H5Error("dense attribute storage is not supported")
            if t != 0x000C:
                continue
            ver = d[0]
            nsz, tsz, ssz = struct.unpack_from("<HHH", d, 2)
            pos = 8
            if ver == 3:
                pos += 1
            pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
            name = d[pos:pos + nsz].split(b"\0")[0].decode("utf8", "replace")
            pos += pad(nsz)
            dt = d[pos:pos + tsz]
            pos += pad(tsz)
            sp = d[pos:pos + ssz]
            pos += pad(ssz)
            try:
                out[name] = self.f._decode(dt, sp, d[pos:])
            except H5Error:
                out[name] = None  # a datatype outside the subset: not needed for loading weights
        return out

    # ---- datasets ----
    def read(self):
        dt = sp = lay = None
        for t, d in self.msgs:
            if t == 0x0001:
                sp = d
            elif t == 0x0003:
                dt = d
            elif t == 0x0008:
                lay = d
            elif t == 0x000B:
                raise H5Error("filtered (compressed) datasets are not supported")
        if dt is None or sp is None or lay is None:
            raise H5Error("not a dataset")
        shape = self.f._shape(sp)
        dtype, _ = self.f._dtype(dt)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        ver = lay[0]
        if ver == 3:
            cls = lay[1]
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", lay, 2)
                raw = b"" if addr == UNDEF else self.f.buf[addr:addr + nbytes]
            elif cls == 0:
                (size,) = struct.unpack_from("<H", lay, 2)
                raw = lay[4:4 + size]
            else:
                raise H5Error("chunked datasets are not supported")
        elif ver in (1, 2):
            rank, cls = lay[1], lay[2]
            if cls == 1:
                (addr,) = struct.unpack_from("<Q", lay, 8)
                raw = self.f.buf[addr:addr + nbytes]
            elif cls == 0:
                (size,) = struct.unpack_from("<I", lay, 8 + 4 * rank)
                raw = lay[12 + 4 * rank:12 + 4 * rank + size]
            else:
                raise H5Error("chunked datasets are not supported")
        else:
            raise H5Error("unknown data layout version %d" % ver)
        if len(raw) < nbytes:
            if len(raw) == 0:
                return np.zeros(shape, dtype)
            raise H5Error("dataset data is truncated")
        return np.frombuffer(raw[:nbytes], dtype=dtype).reshape(shape).copy()


class File(_Obj):
    """Read-only view of an HDF5 file held in memory."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != SIG:
            raise H5Error("%s is not an HDF5 file" % path)
        ver = b[8]
        if ver not in (0, 1):
            raise H5Error("superblock version %d is not supported (expected the 'earliest' format Keras/h5py use)" % ver)
        if b[13] != 8 or b[14] != 8:
            raise H5Error("only 8-byte offsets/lengths are supported")
        pos = 24 + (4 if ver == 1 else 0)
        base, _, _, _ = struct.unpack_from("<QQQQ", b, pos)
        if base != 0:
            raise H5Error("non-zero base address is not supported")
        pos += 32
        _, root_addr = struct.unpack_from("<QQ", b, pos)
        super().__init__(self, root_addr)

    # ---- object headers ----
    def _read_header(self, addr):
        b = self.buf
        if b[addr:addr + 4] == b"OHDR":
            raise H5Error("version-2 object headers are not supported")
        if b[addr] != 1:
            raise H5Error("bad object header at %d" % addr)
        (nmsg,) = struct.unpack_from("<H", b, addr + 2)
        (size,) = struct.unpack_from("<I", b, addr + 8)
        msgs = []
        blocks = [(addr + 16, size)]
        while blocks and len(msgs) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(msgs) < nmsg:
                t, sz, flags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8:pos + 8 + sz]
                pos += 8 + sz
                if flags & 2:
                    raise H5Error("shared object header messages are not supported")
                if t == 0x0010:
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off, ln))
                msgs.append((t, data))
        return msgs

    # ---- groups ----
    def _heap_data(self, addr):
        b = self.buf
        if b[addr:addr + 4] != b"HEAP":
            raise H5Error("bad local heap")
        size, _, data = struct.unpack_from("<QQQ", b, addr + 8)
        return b[data:data + size]

    def _walk_btree(self, addr, heap, out):
        b = self.buf
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad B-tree node")
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 0:
            raise H5Error("unexpected B-tree node type")
        pos = addr + 24
        for i in range(used):
            (child,) = struct.unpack_from("<Q", b, pos + 8)
            pos += 16
            if level > 0:
                self._walk_btree(child, heap, out)
            else:
                if b[child:child + 4] != b"SNOD":
                    raise H5Error("bad symbol node")
                (n,) = struct.unpack_from("<H", b, child + 6)
                for k in range(n):
                    noff, oaddr = struct.unpack_from("<QQ", b, child + 8 + 40 * k)
                    out[heap[noff:heap.index(b"\0", noff)].decode("utf8")] = oaddr

    # ---- datatypes / dataspaces ----
    @staticmethod
    def _shape(sp):
        ver, rank, flags = sp[0], sp[1], sp[2]
        off = 8 if ver == 1 else 4
        if ver == 2 and sp[3] == 2:
            return (0,)
        return tuple(struct.unpack_from("<%dQ" % rank, sp, off)) if rank else ()

    @staticmethod
    def _dtype(dt):
        cls, bits0 = dt[0] & 0x0F, dt[1]
        (size,) = struct.unpack_from("<I", dt, 4)
        if cls == 1:
            if bits0 & 1:
                raise H5Error("big-endian floats are not supported")
            return np.dtype("<f%d" % size), None
        if cls == 0:
            return np.dtype("%s%s%d" % ("<", "i" if dt[1] & 8 else "u", size)), None
        if cls == 3:
            return np.dtype("S%d" % size), None
        if cls == 9 and (bits0 & 0x0F) == 1:
            return np.dtype("V16"), "vlen_str"  # (length u32, global heap address u64, object index u32)
        raise H5Error("datatype class %d is not supported" % cls)

    def _global_heap_object(self, addr, index):
        """Object `index` of the global heap collection at `addr` (variable-length data: Keras' model_config /
        training_config, which h5py stores as H5T_VARIABLE strings)."""
        b = self.buf
        if b[addr:addr + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        (size,) = struct.unpack_from("<Q", b, addr + 8)
        pos, end = addr + 16, addr + size
        while pos + 16 <= end:
            idx, _, osize = struct.unpack_from("<HH4xQ", b, pos)
            if idx == 0:
                break  # free space: no further objects
            if idx == index:
                return b[pos + 16:pos + 16 + osize]
            pos += 16 + ((osize + 7) & ~7)
        raise H5Error("global heap object %d not found" % index)

    def _decode(self, dt, sp, data):
        shape = self._shape(sp)
        dtype, kind = self._dtype(dt)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if kind == "vlen_str":
            vals = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", data, 16 * i)
                vals.append(b"" if ln == 0 else self._global_heap_object(gaddr, gidx)[:ln])
            return np.array(vals, dtype=object).reshape(shape) if shape else vals[0]
        arr = np.frombuffer(data[:n * dtype.itemsize], dtype=dtype)
        return arr.reshape(shape).copy() if shape else arr[0]


def _names_attr(attrs, key):
    """Keras splits oversized name lists into key0, key1, ... (saving.py save_attributes_to_hdf5_group)."""
    if key in attrs and attrs[key] is not None:
        vals = np.atleast_1d(attrs[key])
    else:
        vals, i = [], 0
        while "%s%d" % (key, i) in attrs:
            vals.extend(np.atleast_1d(attrs["%s%d" % (key, i)]))
            i += 1
    return [v.decode("utf8") if isinstance(v, (bytes, np.bytes_)) else str(v) for v in vals]


def read_keras_file(path):
    """-> (layer_names in file order, {layer: [(weight_name, ndarray), ...]})."""
    f = File(path)
    g = f["model_weights"] if "model_weights" in f.children() else f  # Keras falls back the same way
    attrs = g.attrs()
    layers = _names_attr(attrs, "layer_names")
    if not layers:
        raise H5Error("%s has no layer_names attribute (not a Keras weight file)" % path)
    out = {}
    for ln in layers:
        lg = g[ln]
        wn = _names_attr(lg.attrs(), "weight_names")
        out[ln] = [(w, lg[w].read()) for w in wn]
    return layers, out


def read_keras_configs(path):
    """-> (model_config, training_config) of a Keras full-model file as parsed JSON (None when absent): the attributes
    `model.save` writes next to /model_weights (TG:892, TU:622) and `keras.models.load_model` reads back."""
    import json
    attrs = File(path).attrs()
    out = []
    for key in ("model_config", "training_config"):
        v = attrs.get(key)
        if v is None:
            out.append(None)
            continue
        if isinstance(v, np.ndarray):
            v = v.ravel()[0]
        out.append(json.loads(v.decode("utf8") if isinstance(v, (bytes, np.bytes_)) else str(v)))
    return tuple(out)


def load_keras_weights(path, wanted):
    """wanted: [('layer/weight', shape)].  Resolves tensors through each layer group's ``weight_names``
    attribute (TF scopes of later folds are uniquified: ``conv2d_gen_0_1/kernel:0``), keyed by the h5 layer
    group name + the weight's base name.  Returns {'layer/weight': float32 array}."""
    layers, content = read_keras_file(path)
    have = {}
    for ln in layers:
        for wname, arr in content[ln]:
            base = wname.split("/")[-1].split(":")[0]
            have["%s/%s" % (ln, base)] = arr
    auto_dense = [ln for ln in layers if re.fullmatch(r"dense_\d+", ln) and content[ln]]
    out = {}
    for name, shape in wanted:
        key = name
        if key not in have and name.startswith("dense_1/") and len(auto_dense) == 1:
            key = auto_dense[0] + "/" + name.split("/")[1]  # auto-named critic Dense (TG:342)
        if key not in have:
            raise KeyError("%s: weight %s not found in the file" % (path, name))
        arr = np.asarray(have[key], np.float32)
        if tuple(arr.shape) != tuple(shape):
            raise ValueError("%s: %s has shape %s, the model expects %s" % (path, name, arr.shape, tuple(shape)))
        out[name] = arr
    return out


# ==========================================================================================================
# writer
# ==========================================================================================================
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(t, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", t, len(data), flags) + data


def _dt_float32():
    return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _dt_string(n):
    return struct.pack("<BBBBI", 0x13, 0x00, 0x00, 0x00, n)  # null-padded ASCII fixed-length string


def _space(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


class VLenStr:
    """Marks an attribute value to be stored as a scalar variable-length string (what h5py does with Python bytes / str:
    Keras' keras_version, backend, model_config, training_config)."""

    def __init__(self, text):
        self.raw = text.encode("utf8") if isinstance(text, str) else bytes(text)


def _dt_vlen_str():
    # class 9 (variable length) version 1, type = string, null-terminated, ASCII; base type: 1-byte fixed string
    return struct.pack("<BBBBI", 0x19, 0x01, 0x00, 0x00, 16) + _dt_string(1)


def _dt_int64():
    return struct.pack("<BBBBI", 0x10, 0x08, 0x00, 0x00, 8) + struct.pack("<HH", 0, 64)


def _attr_msg(name, value, writer=None):
    if isinstance(value, VLenStr):
        gaddr = writer.global_heap(value.raw)
        dt, sp, data = _dt_vlen_str(), _space(()), struct.pack("<IQI", len(value.raw), gaddr, 1)
    elif isinstance(value, (bytes, str)):
        raw = value.encode("utf8") if isinstance(value, str) else value
        dt, sp, data = _dt_string(max(1, len(raw))), _space(()), raw or b"\0"
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "S":
            dt, sp, data = _dt_string(arr.dtype.itemsize), _space(arr.shape), arr.tobytes()
        else:
            arr = arr.astype("<f4")
            dt, sp, data = _dt_float32(), _space(arr.shape), arr.tobytes()
    nm = name.encode("utf8") + b"\0"
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + data
    return _msg(0x000C, body)


class _Writer:
    K = 16

    def __init__(self):
        self.buf = bytearray(96)  # superblock placeholder

    def alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    def header(self, msgs):
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)

    def global_heap(self, blob):
        """A global heap collection holding one object (index 1) + the free-space object; returns its address."""
        body = struct.pack("<HH4xQ", 1, 1, len(blob)) + _pad8(blob)
        size = max(4096, (16 + len(body) + 16 + 4095) & ~4095)
        free = size - 16 - len(body)
        body += struct.pack("<HH4xQ", 0, 0, free) + b"\0" * (free - 16)
        return self.alloc(b"GCOL" + struct.pack("<B3xQ", 1, size) + body)

    def dataset(self, arr):
        arr = np.asarray(arr)
        if arr.dtype.kind in "iu":  # the optimizer's iteration counter
            arr = np.asarray(arr, "<i8", order="C")
            dtmsg = _dt_int64()
        else:
            arr = np.asarray(arr, "<f4", order="C")
            dtmsg = _dt_float32()
        addr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        msgs = [_msg(0x0001, _space(arr.shape)), _msg(0x0003, dtmsg, flags=1),
                _msg(0x0005, struct.pack("<BBBBI", 2, 2, 0, 1, 0)),
                _msg(0x0008, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes))]
        return self.header(msgs)

    def group(self, children, attrs=()):
        """children: {name: object header address}."""
        names = sorted(children, key=lambda s: s.encode("utf8"))
        heap = bytearray(8)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode("utf8") + b"\0")
        heap_data = self.alloc(bytes(heap) if len(heap) >= 8 else b"\0" * 8)
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data))
        K = self.K  # group leaf / internal K, also written into the superblock
        if len(names) > 4 * K * K:
            raise H5Error("too many links for a single-level group B-tree")
        chunks = [names[i:i + 2 * K] for i in range(0, len(names), 2 * K)]
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(chunks), UNDEF, UNDEF))
        tree += struct.pack("<Q", 0)  # key 0: the empty string at heap offset 0
        for ch in chunks:  # one symbol node per 2K names; key i+1 = largest name in child i
            snod = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(ch)))
            for n in ch:
                snod += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
            snod += b"\0" * (8 + 40 * 2 * K - len(snod))
            tree += struct.pack("<QQ", self.alloc(bytes(snod)), offs[ch[-1]])
        tree += b"\0" * (24 + 8 + 16 * 2 * K - len(tree))
        tree_addr = self.alloc(bytes(tree))
        msgs = [_msg(0x0011, struct.pack("<QQ", tree_addr, heap_addr))] + [_attr_msg(k, v, self) for k, v in attrs]
        return self.header(msgs), tree_addr, heap_addr

    def finish(self, root):
        addr, tree, heap = root
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.K, self.K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, addr, 1, 0) + struct.pack("<QQ", tree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def _fixed_strings(names):
    enc = [n.encode("utf8") for n in names]
    width = max([len(e) for e in enc] + [1])
    return np.array(enc, dtype="S%d" % width) if enc else np.zeros((0,), "S1")


def save_keras_weights(path, weights, order, extra_layers=(), tf_scope_suffix=""):
    """Writes ``/model_weights`` in the Keras 2.x layout: attr ``layer_names`` (order of first appearance in
    ``order``), per-layer groups with attr ``weight_names`` = ['<scope>/<weight>:0', ...] and nested datasets.
    ``extra_layers`` adds weight-less layers (Input/Activation/...) exactly as Keras lists them;
    ``tf_scope_suffix`` ('_1', ...) reproduces the uniquified TF scopes of folds 2-4."""
    w = _Writer()
    layers = []
    for name in order:
        ln = name.split("/")[0]
        if ln not in layers:
            layers.append(ln)
    layer_objs = {}
    for ln in layers:
        wnames = [n for n in order if n.split("/")[0] == ln]
        scope = ln + tf_scope_suffix
        ds = {n.split("/")[1] + ":0": w.dataset(weights[n]) for n in wnames}
        scope_grp = w.group(ds)[0]
        attr = _fixed_strings(["%s/%s:0" % (scope, n.split("/")[1]) for n in wnames])
        layer_objs[ln] = w.group({scope: scope_grp}, [("weight_names", attr)])[0]
    for ln in extra_layers:
        layer_objs[ln] = w.group({}, [("weight_names", _fixed_strings([]))])[0]
    all_layers = list(layers) + [l for l in extra_layers if l not in layers]
    mw = w.group(layer_objs, [("layer_names", _fixed_strings(all_layers)), ("backend", b"tensorflow"),
                              ("keras_version", b"2.2.4")])[0]
    root = w.group({"model_weights": mw}, [("keras_version", b"2.2.4"), ("backend", b"tensorflow")])
    with open(path, "wb") as fh:
        fh.write(w.finish(root))


def _nested_group(w, items):
    """items: [('a/b/c:0', array)] -> object header address of a group tree whose leaves are datasets."""
    here, sub = {}, {}
    for name, arr in items:
        head, _, rest = name.partition("/")
        if rest:
            sub.setdefault(head, []).append((rest, arr))
        else:
            here[head] = w.dataset(arr)
    for head, rest in sub.items():
        here[head] = _nested_group(w, rest)
    return here


def save_keras_model(path, weights, layer_names, layer_weights, model_config=None, training_config=None,
                     optimizer_weights=None, tf_scope_suffix=""):
    """Writes a Keras 2.x FULL-MODEL file (what `model.save` produces, TG:892 / TU:622):

      /                      attrs keras_version, backend, model_config [, training_config]   (variable-length strings)
      /model_weights         attrs layer_names (Keras' model.layers order, weight-less layers included; split into
                             layer_names0.. above the 64 512-byte object-header limit like Keras does), backend,
                             keras_version; one group per layer with attr weight_names and datasets <scope>/<weight>:0
      /optimizer_weights     attr weight_names + nested datasets (compiled models only)

    weights: {'layer/weight': array}; layer_weights: {layer: [weight names in Keras order]} (keras_config.describe)."""
    w = _Writer()
    layer_objs = {}
    for ln in layer_names:
        wn = list(layer_weights.get(ln, ()))
        scope = ln + tf_scope_suffix
        kids = {}
        if wn:
            ds = {n + ":0": w.dataset(weights["%s/%s" % (ln, n)]) for n in wn}
            kids[scope] = w.group(ds)[0]
        attr = _fixed_strings(["%s/%s:0" % (scope, n) for n in wn])
        layer_objs[ln] = w.group(kids, [("weight_names", attr)])[0]
    names = _fixed_strings(list(layer_names))
    limit = 64512  # HDF5_OBJECT_HEADER_LIMIT of keras/engine/saving.py
    if names.nbytes <= limit:
        name_attrs = [("layer_names", names)]
    else:
        parts = int(np.ceil(names.nbytes / float(limit)))
        name_attrs = [("layer_names%d" % i, c) for i, c in enumerate(np.array_split(names, parts))]
    mw = w.group(layer_objs, name_attrs + [("backend", VLenStr("tensorflow")), ("keras_version", VLenStr("2.2.4"))])[0]
    root = {"model_weights": mw}
    if optimizer_weights:
        tree = _nested_group(w, optimizer_weights)

        def build(node):
            return w.group({k: (build(v) if isinstance(v, dict) else v) for k, v in node.items()})[0]

        ow_names = _fixed_strings([n for n, _ in optimizer_weights])
        root["optimizer_weights"] = w.group({k: (build(v) if isinstance(v, dict) else v) for k, v in tree.items()},
                                            [("weight_names", ow_names)])[0]
    attrs = [("keras_version", VLenStr("2.2.4")), ("backend", VLenStr("tensorflow"))]
    if model_config is not None:
        attrs.append(("model_config", VLenStr(model_config)))
    if training_config is not None:
        attrs.append(("training_config", VLenStr(training_config)))
    with open(path, "wb") as fh:
        fh.write(w.finish(w.group(root, attrs)))
