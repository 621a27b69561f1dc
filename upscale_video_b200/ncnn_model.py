"""Model loading for the frame-upscale path.

Two on-disk formats are understood:

* the ncnn pair ``<name>.param`` (text graph) + ``<name>.bin`` (raw weight blob) that the
  reference hands to ``ncnn.Net.load_param/load_model`` in ``init_worker``
  (reference ``upscale/upscale_processing.py:70-71``; files ``models/*.param|.bin``);
* ``<name>.b2sr`` -- this repo's own single-file container (JSON graph + little-endian arrays,
  fp16 wherever that is lossless) produced by ``tools/convert_models.py`` from the ncnn pair.

Both decode to the same in-memory :class:`Graph`.  The graph is then *recognised* as one of the network
families the CUDA engine runs (``compact_desc``) and its parameters packed into the flat blob that
``b2sr_create`` (``include/b2sr.h``) takes.

ncnn text grammar (public ncnn format): line 1 magic ``7767517``; line 2 ``<layers> <blobs>``; then per layer
``type name n_in n_out in... out... key=value...`` where a key ``-233xx`` introduces an array
``count,v0,v1...`` for parameter id ``xx``.  ``.bin`` layout per Convolution: u32 tag (``0x01306B47`` -> fp16
weights padded to 4 bytes, ``0`` -> raw fp32), weights in OIHW order, then fp32 bias when ``5=1``; PReLU:
raw fp32 slopes with no tag.
"""
from __future__ import annotations

import json
import os
import struct
from dataclasses import dataclass, field

import numpy as np

NCNN_MAGIC = "7767517"
FP16_TAG = 0x01306B47
B2SR_MAGIC = b"B2SRv1\0\0"


@dataclass
class Layer:
    type: str
    name: str
    bottoms: list
    tops: list
    params: dict = field(default_factory=dict)  # int id -> int | float | list
    weights: dict = field(default_factory=dict)  # "weight" | "bias" | "slope" -> np.ndarray

    def p(self, key, default=0):
        return self.params.get(key, default)


@dataclass
class Graph:
    layers: list
    n_blobs: int = 0

    def convs(self):
        return [l for l in self.layers if l.type == "Convolution"]

    def nbytes(self):
        return sum(a.nbytes for l in self.layers for a in l.weights.values())


def _parse_value(text):
    try:
        return int(text)
    except ValueError:
        return float(text)


def parse_param(text: str) -> Graph:
    lines = [ln.strip() for ln in text.splitlines() if ln.strip()]
    if lines[0] != NCNN_MAGIC:
        raise ValueError("not an ncnn param file (magic %r)" % lines[0])
    n_layers, n_blobs = (int(v) for v in lines[1].split())
    layers = []
    for ln in lines[2:]:
        tok = ln.split()
        ltype, name, n_in, n_out = tok[0], tok[1], int(tok[2]), int(tok[3])
        bottoms = tok[4 : 4 + n_in]
        tops = tok[4 + n_in : 4 + n_in + n_out]
        params = {}
        for kv in tok[4 + n_in + n_out :]:
            k, v = kv.split("=", 1)
            k = int(k)
            if k <= -23300:  # array parameter: "-233xx=count,v0,v1,..."
                items = v.split(",")
                params[-k - 23300] = [_parse_value(x) for x in items[1 : 1 + int(items[0])]]
            else:
                params[k] = _parse_value(v)
        layers.append(Layer(ltype, name, bottoms, tops, params))
    if len(layers) != n_layers:
        raise ValueError("param file declares %d layers, found %d" % (n_layers, len(layers)))
    return Graph(layers, n_blobs)


def attach_bin(graph: Graph, blob: bytes) -> int:
    """Consume ``blob`` in layer order; returns the number of bytes used (must equal len(blob))."""
    off = 0
    for l in graph.layers:
        if l.type == "Convolution":
            count = int(l.p(6))
            (tag,) = struct.unpack_from("<I", blob, off)
            off += 4
            if tag == FP16_TAG:
                w = np.frombuffer(blob, "<f2", count, off).copy()
                off += (count * 2 + 3) & ~3
            elif tag == 0:
                w = np.frombuffer(blob, "<f4", count, off).copy()
                off += count * 4
            else:
                raise ValueError("unsupported ncnn weight tag 0x%08x in %s" % (tag, l.name))
            cout, k = int(l.p(0)), int(l.p(1))
            cin = count // (cout * k * k)
            l.weights["weight"] = w.reshape(cout, cin, k, k)
            if l.p(5):
                l.weights["bias"] = np.frombuffer(blob, "<f4", cout, off).copy()
                off += cout * 4
        elif l.type == "PReLU":
            n = int(l.p(0))
            l.weights["slope"] = np.frombuffer(blob, "<f4", n, off).copy()
            off += n * 4
    return off


def load_ncnn(param_path: str, bin_path: str) -> Graph:
    with open(param_path, "r") as f:
        g = parse_param(f.read())
    with open(bin_path, "rb") as f:
        blob = f.read()
    used = attach_bin(g, blob)
    if used != len(blob):
        raise ValueError("%s: consumed %d of %d bytes" % (bin_path, used, len(blob)))
    return g


# ----------------------------------------------------------------------------------------------
# .b2sr container
# ----------------------------------------------------------------------------------------------
def _fp16_lossless(a: np.ndarray) -> bool:
    return a.dtype == np.float32 and bool(np.array_equal(a.astype(np.float16).astype(np.float32), a))


def save_b2sr(graph: Graph, path: str) -> None:
    """Arrays keep the ncnn (OIHW) element order; weights are narrowed to fp16 only when that is exact."""
    meta, chunks, off = [], [], 0
    for l in graph.layers:
        entry = {"type": l.type, "name": l.name, "bottoms": l.bottoms, "tops": l.tops,
                 "params": {str(k): v for k, v in l.params.items()}, "arrays": {}}
        for key, arr in l.weights.items():
            if key == "weight" and _fp16_lossless(arr):
                arr = arr.astype(np.float16)
            raw = np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<")).tobytes()
            entry["arrays"][key] = {"dtype": arr.dtype.name, "shape": list(arr.shape), "offset": off,
                                    "nbytes": len(raw)}
            pad = (-len(raw)) % 16
            chunks.append(raw + b"\0" * pad)
            off += len(raw) + pad
        meta.append(entry)
    header = json.dumps({"n_blobs": graph.n_blobs, "layers": meta}, separators=(",", ":")).encode()
    header += b" " * ((-(len(B2SR_MAGIC) + 4 + len(header))) % 16)
    with open(path, "wb") as f:
        f.write(B2SR_MAGIC)
        f.write(struct.pack("<I", len(header)))
        f.write(header)
        for c in chunks:
            f.write(c)


def load_b2sr(path: str) -> Graph:
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != B2SR_MAGIC:
        raise ValueError("%s: not a b2sr model" % path)
    (hlen,) = struct.unpack_from("<I", data, 8)
    head = json.loads(data[12 : 12 + hlen].decode())
    base = 12 + hlen
    layers = []
    for e in head["layers"]:
        l = Layer(e["type"], e["name"], e["bottoms"], e["tops"], {int(k): v for k, v in e["params"].items()})
        for key, a in e["arrays"].items():
            arr = np.frombuffer(data, np.dtype(a["dtype"]).newbyteorder("<"), int(np.prod(a["shape"])),
                                base + a["offset"])
            l.weights[key] = arr.reshape(a["shape"]).astype(a["dtype"])
        layers.append(l)
    return Graph(layers, head.get("n_blobs", 0))


def packaged_model_dir() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")


def load_model(model_path: str, stem: str) -> Graph:
    """Resolve ``<model_path>/<stem>`` the way ``init_worker`` names models (reference
    ``upscale_processing.py:70-71``: ``str(scale) + model_file`` + ``.param``/``.bin``), preferring the
    ncnn pair, then a ``.b2sr`` beside it, then the converted copy packaged with this repo."""
    param = os.path.join(model_path, stem + ".param")
    binf = os.path.join(model_path, stem + ".bin")
    if os.path.exists(param) and os.path.exists(binf):
        return load_ncnn(param, binf)
    for d in (model_path, packaged_model_dir()):
        cand = os.path.join(d, stem + ".b2sr")
        if os.path.exists(cand):
            return load_b2sr(cand)
    raise FileNotFoundError("model %s(.param/.bin|.b2sr) not found under %s" % (stem, model_path))


# ----------------------------------------------------------------------------------------------
# Family recognition + blob packing for the C ABI
# ----------------------------------------------------------------------------------------------
FAMILY_COMPACT = 1  # SRVGGNetCompact: conv(cin->nf)+PReLU, n_mid x [conv(nf->nf)+PReLU], conv(nf->cin*r^2),
#                     PixelShuffle(r) + nearest-upsample(input, r)   (reference models/2x_Compact_Pretrain.param:3-42)


@dataclass
class CompactDesc:
    cin: int
    nf: int
    n_mid: int
    scale: int
    cout_last: int
    input_blob: str
    output_blob: str


def compact_desc(graph: Graph) -> CompactDesc | None:
    """Return the SRVGGNetCompact description of ``graph`` or None when it is some other topology
    (e.g. the RRDB ``4x_Valar_v1``).  The check is structural and strict: every layer must be accounted for."""
    ls = graph.layers
    if len(ls) < 8 or ls[0].type != "Input" or ls[1].type != "Split" or len(ls[1].tops) != 2:
        return None
    skip_blob, x = ls[1].tops[0], ls[1].tops[1]
    i, convs = 2, []
    while i + 1 < len(ls) and ls[i].type == "Convolution" and ls[i + 1].type == "PReLU":
        c, p = ls[i], ls[i + 1]
        if c.bottoms != [x] or p.bottoms != c.tops:
            return None
        convs.append(c)
        x = p.tops[0]
        i += 2
    if len(convs) < 2 or i + 3 >= len(ls):
        return None
    last, ps, interp, add = ls[i], ls[i + 1], ls[i + 2], ls[i + 3]
    if (last.type, ps.type, interp.type, add.type) != ("Convolution", "PixelShuffle", "Interp", "BinaryOp"):
        return None
    if i + 4 != len(ls) or last.bottoms != [x] or ps.bottoms != last.tops:
        return None
    if interp.bottoms != [skip_blob] or sorted(add.bottoms) != sorted([ps.tops[0], interp.tops[0]]):
        return None
    r = int(ps.p(0, 1))
    if ps.p(1, 0) != 0 or add.p(0, 0) != 0:  # PixelShuffle mode 0, BinaryOp op 0 (add)
        return None
    if r > 1 and (interp.p(0) != 1 or float(interp.p(1, 1.0)) != r or float(interp.p(2, 1.0)) != r):
        return None
    for c in convs + [last]:
        if (c.p(1), c.p(2, 1), c.p(3, 1), c.p(4, 0), c.p(5, 0), c.p(9, 0)) != (3, 1, 1, 1, 1, 0):
            return None
    nf = int(convs[0].p(0))
    cin = convs[0].weights["weight"].shape[1]
    if any(c.weights["weight"].shape != (nf, nf, 3, 3) for c in convs[1:]):
        return None
    cout_last = int(last.p(0))
    if cout_last != cin * r * r or last.weights["weight"].shape != (cout_last, nf, 3, 3):
        return None
    return CompactDesc(cin, nf, len(convs) - 1, r, cout_last, ls[0].tops[0], add.tops[0])


def pack_compact_blob(graph: Graph) -> tuple:
    """Flat fp32 blob for ``b2sr_create``: per conv in graph order ``weight[OIHW] | bias[O] | slope[O]``
    (slope omitted for the last conv).  fp32 keeps the ABI trivially typed; the engine narrows weights to
    fp16 itself and refuses (error) if that is not exact."""
    d = compact_desc(graph)
    if d is None:
        raise ValueError("graph is not an SRVGGNetCompact")
    parts = []
    ls = graph.layers
    for i, l in enumerate(ls):
        if l.type != "Convolution":
            continue
        parts.append(l.weights["weight"].astype(np.float32).ravel())
        parts.append(l.weights["bias"].astype(np.float32).ravel())
        if i + 1 < len(ls) and ls[i + 1].type == "PReLU":
            parts.append(ls[i + 1].weights["slope"].astype(np.float32).ravel())
    return d, np.ascontiguousarray(np.concatenate(parts))


# ----------------------------------------------------------------------------------------------
# Generic graphs (anything that is not an SRVGGNetCompact, today 4x_Valar_v1): flat op list for b2sr_create_graph
# ----------------------------------------------------------------------------------------------
OP_CONV, OP_PRELU, OP_PIXELSHUFFLE, OP_NEAREST, OP_ADD, OP_CONCAT = 1, 2, 3, 4, 5, 6


@dataclass
class GraphProgram:
    ops: list          # dicts with the fields of b2sr_graph_op
    n_slots: int
    in_slot: int
    out_slot: int
    scale: int
    weights: np.ndarray  # flat fp32 blob referenced by w_off / b_off
    input_blob: str = "input"
    output_blob: str = "output"


def _analyse_graph(graph: Graph, input_blob: str, output_blob: str, views: bool):
    """Passes 1 and 2 of :func:`compile_graph` (shared with :func:`compile_fused`): nodes with inferred shapes,
    Split aliases resolved, concat families found.  Returns a dict of the tables the later passes use."""
    alias = {}

    def val(name):  # resolve Split aliases
        while name in alias:
            name = alias[name]
        return name

    for l in graph.layers:
        if l.type == "Split":
            for t in l.tops:
                alias[t] = l.bottoms[0]
    wparts, woff = [], 0

    def push_w(arr):
        nonlocal woff
        a = np.ascontiguousarray(arr, np.float32).ravel()
        wparts.append(a)
        off = woff
        woff += a.size
        return off

    # ---- pass 1: nodes and shapes -------------------------------------------------------------------------
    chans, res = {}, {}
    nodes = []
    seen_input = False
    for l in graph.layers:
        t = l.type
        if t == "Input":
            if l.tops != [input_blob]:
                raise ValueError("graph input blob is %r, expected %r" % (l.tops, input_blob))
            chans[input_blob], res[input_blob] = 3, 1
            seen_input = True
            continue
        if t == "Split":
            continue
        ins = [val(b) for b in l.bottoms]
        for b in ins:
            if b not in chans:
                raise ValueError("layer %s reads undefined blob %s" % (l.name, b))
        op = {"type": 0, "nin": len(ins), "cin": 0, "cout": 0, "k": 0, "act": 0, "slope": 0.0,
              "coef": [1.0, 1.0], "plain": 0, "r": 1, "w_off": -1, "b_off": -1}
        c_out, r_out = chans[ins[0]], res[ins[0]]
        in_place_ok = False
        if t == "Convolution":
            k, pad = int(l.p(1)), int(l.p(4, 0))
            if (l.p(2, 1), l.p(3, 1)) != (1, 1) or k not in (1, 3) or pad != k // 2:
                raise ValueError("unsupported convolution %s (k=%d pad=%d)" % (l.name, k, pad))
            w = l.weights["weight"].astype(np.float32)
            act = int(l.p(9, 0))
            if act not in (0, 2):
                raise ValueError("unsupported convolution activation %d in %s" % (act, l.name))
            if int(w.shape[1]) != c_out:
                raise ValueError("convolution %s expects %d channels, gets %d" % (l.name, w.shape[1], c_out))
            op.update(type=OP_CONV, cin=int(w.shape[1]), cout=int(w.shape[0]), k=k, act=act,
                      slope=float(l.p(10, [0.0])[0]) if act == 2 else 0.0, w_off=push_w(w),
                      b_off=push_w(l.weights["bias"]) if l.p(5, 0) else -1)
            c_out = int(w.shape[0])
        elif t == "PReLU":
            op.update(type=OP_PRELU, w_off=push_w(l.weights["slope"]))
            in_place_ok = True
        elif t == "PixelShuffle":
            r = int(l.p(0, 1))
            if l.p(1, 0) != 0:
                raise ValueError("PixelShuffle mode %r" % l.p(1))
            op.update(type=OP_PIXELSHUFFLE, r=r)
            c_out, r_out = c_out // (r * r), r_out * r
        elif t == "Interp":
            sy, sx = float(l.p(1, 1.0)), float(l.p(2, 1.0))
            if l.p(0) != 1 or sy != sx or sy != int(sy):
                raise ValueError("unsupported Interp %s" % l.name)
            op.update(type=OP_NEAREST, r=int(sy))
            r_out *= int(sy)
        elif t == "BinaryOp":
            if l.p(0, 0) != 0 or len(ins) != 2:
                raise ValueError("unsupported BinaryOp %s" % l.name)
            op.update(type=OP_ADD, plain=1)
            in_place_ok = True
        elif t == "Eltwise":
            co = l.p(1, [1.0] * len(ins))
            if l.p(0, 0) != 1 or len(ins) != 2:
                raise ValueError("unsupported Eltwise %s" % l.name)
            op.update(type=OP_ADD, coef=[float(np.float32(co[0])), float(np.float32(co[1]))])
            in_place_ok = True
        elif t == "Concat":
            if l.p(0, 0) != 0 or len(ins) > 6:
                raise ValueError("unsupported Concat %s" % l.name)
            op.update(type=OP_CONCAT)
            c_out = sum(chans[b] for b in ins)
        else:
            raise ValueError("unsupported layer type %s" % t)
        out_name = l.tops[0]
        chans[out_name], res[out_name] = c_out, r_out
        nodes.append({"op": op, "ins": ins, "out": out_name, "in_place_ok": in_place_ok, "view": False})
    out_v = val(output_blob)
    if not seen_input or out_v not in chans:
        raise ValueError("graph has no blob %r" % (output_blob if seen_input else input_blob))

    # ---- pass 2: concat families ----------------------------------------------------------------------------
    member = {}    # value -> (family, channel offset)
    fam_ld = []    # family -> channels of the wide slot
    if views:
        concats = [n for n in nodes if n["op"]["type"] == OP_CONCAT]
        concat_outs = {n["out"] for n in concats}
        taken = set()
        for big in sorted(concats, key=lambda n: -len(n["ins"])):
            if id(big) in taken:
                continue
            vals = big["ins"]
            if (len(set(vals)) != len(vals) or any(v in member or v in concat_outs or v == input_blob for v in vals)
                    or big["out"] == out_v):
                continue
            fam = len(fam_ld)
            off = 0
            for v in vals:
                member[v] = (fam, off)
                off += chans[v]
            fam_ld.append(off)
            for n in concats:
                if id(n) not in taken and n["out"] != out_v and n["ins"] == vals[:len(n["ins"])]:
                    taken.add(id(n))
                    n["view"] = True
                    n["family"] = fam

    return {"nodes": nodes, "chans": chans, "res": res, "out_v": out_v, "member": member, "fam_ld": fam_ld,
            "weights": np.concatenate(wparts) if wparts else np.zeros(1, np.float32)}


def compile_graph(graph: Graph, input_blob: str = "input", output_blob: str = "output", views: bool = True) -> GraphProgram:
    """Lower an ncnn graph to the op list of ``b2sr_create_graph`` (include/b2sr.h).

    Three passes: (1) layers -> nodes with inferred shapes (Split layers become aliases); (2) Concat families: when
    values are only ever concatenated as prefixes of one list ([x], [x,x1], [x,x1,x2] ... -- every dense block of the
    Valar RRDB graph), the producers write straight into channel slices of one wide slot and the Concat layers
    disappear, their outputs being strided views (``views=False`` keeps them as copies); (3) slot assignment with
    recycling after a value's last use (the Valar graph has 2127 blobs but never more than a dozen alive).

    Layer semantics follow the ncnn layer definitions the reference's model files rely on (reference
    models/4x_Valar_v1.param:1-1208): Convolution keys 0=out-ch 1=kernel 4=pad 5=bias 6=weight-count 9=activation
    (2 = LeakyReLU, slope in -23310), PReLU 0=slopes, PixelShuffle 0=factor (mode 0), Interp 0=1 nearest with scales
    1/2, BinaryOp 0=0 add, Eltwise 0=1 sum with coefficients -23301, Concat 0=0 channel axis."""
    A = _analyse_graph(graph, input_blob, output_blob, views)
    nodes, chans, res, out_v, member, fam_ld = A["nodes"], A["chans"], A["res"], A["out_v"], A["member"], A["fam_ld"]

    # ---- pass 3: slots ------------------------------------------------------------------------------------------
    live = [n for n in nodes if not n["view"]]
    view_of = {n["out"]: n["family"] for n in nodes if n["view"]}
    last_use = {}
    for i, n in enumerate(live):
        for b in n["ins"]:
            last_use[b] = i
    last_use[out_v] = len(live) + 1
    fam_last = [-1] * len(fam_ld)
    for v, (f, _) in member.items():
        fam_last[f] = max(fam_last[f], last_use.get(v, -1))
    for v, f in view_of.items():
        fam_last[f] = max(fam_last[f], last_use.get(v, -1))
    free, n_slots = [], 0

    def new_slot():
        nonlocal n_slots
        if free:
            return free.pop()
        n_slots += 1
        return n_slots - 1

    in_slot = new_slot()
    loc = {input_blob: (in_slot, 0, 3, 0)}  # value -> (slot, channel offset, channels, pixel stride; 0 = dense)
    fam_slot = [None] * len(fam_ld)
    ops = []
    for i, n in enumerate(live):
        op, ins = n["op"], n["ins"]
        for b in ins:
            if b in view_of and b not in loc:
                f = view_of[b]
                loc[b] = (fam_slot[f], 0, chans[b], fam_ld[f])
        il = [loc[b] for b in ins]
        dying = [b for b in dict.fromkeys(ins) if last_use.get(b) == i and b not in member and b not in view_of]
        out_name = n["out"]
        if out_name in member:
            f, off = member[out_name]
            if fam_slot[f] is None:
                fam_slot[f] = new_slot()  # before the dying inputs are released: a strided output never aliases an input
            oloc = (fam_slot[f], off, chans[out_name], fam_ld[f])
            for b in dying:
                free.append(loc[b][0])
        elif n["in_place_ok"]:
            for b in dying:  # dense, same shape: the elementwise op may overwrite it
                free.append(loc[b][0])
            oloc = (new_slot(), 0, chans[out_name], 0)
        else:
            oloc = (new_slot(), 0, chans[out_name], 0)
            for b in dying:
                free.append(loc[b][0])
        for f in range(len(fam_ld)):
            if fam_last[f] == i and fam_slot[f] is not None:
                free.append(fam_slot[f])
        loc[out_name] = oloc
        op["in"] = [l[0] for l in il]
        op["in_off"] = [l[1] for l in il]
        op["in_c"] = [l[2] for l in il]
        op["in_ld"] = [l[3] for l in il]
        op["out"], op["out_off"], op["out_c"], op["out_ld"] = oloc
        op["in_res"], op["out_res"] = int(res[ins[0]]), int(res[out_name])
        ops.append(op)
    return GraphProgram(ops, n_slots, in_slot, loc[out_v][0], int(res[out_v]), A["weights"], input_blob, output_blob)


# ----------------------------------------------------------------------------------------------
# Fused lowering for the tcgen05 graph kernels (b2sr_create_fused): every convolution carries its bias, activation
# and the element-wise adds that follow it; activations live in fp16 (what the next convolution reads through TMA)
# and, where a later residual add needs the unrounded value, also in fp32.
# ----------------------------------------------------------------------------------------------
FOP_CONV, FOP_NEAREST = 1, 2
F16, F32 = 2, 4


@dataclass
class FusedProgram:
    ops: list      # dicts with the fields of b2sr_fused_op
    bufs: list     # dicts with the fields of b2sr_fused_buf: channels (pixel stride), dtype (F16 / F32), res
    scale: int
    weights: np.ndarray
    input_blob: str = "input"
    output_blob: str = "output"


def compile_fused(graph: Graph, input_blob: str = "input", output_blob: str = "output", fp32_chain: int = 3,
                  fuse_shortcuts: bool = True) -> FusedProgram | None:
    """Lower an ncnn graph to the op list of ``b2sr_create_fused`` (include/b2sr.h), or return None when the graph has
    a layer that does not fit the fused form (the caller then uses :func:`compile_graph`).

    A fused op is ``v = act(conv(x) + bias)`` followed by up to two residual terms ``v = v * cv + r * cr`` -- the
    BinaryOp / Eltwise layers that consume the convolution's value and nothing else does (reference
    models/4x_Valar_v1.param:11-12 ``Conv_4, Conv_6, Add_7``; :22 ``Add_19``; :57-58 ``Add_57, Add_60``).  The op is
    scheduled where its last add stood.  Concat layers must all be views of a wide buffer (the dense-block pattern
    :func:`_analyse_graph` recognises): a convolution then reads channels [0, cin) of that buffer and producers
    write their channel slice.  Values read by a convolution or an Interp are stored in fp16, values read by a
    residual add in fp32 (both when both happen), which is the arithmetic of the generic engine (fp32 storage,
    fp16 rounding when a convolution stages its input) without the fp32 round trips."""
    A = _analyse_graph(graph, input_blob, output_blob, True)
    nodes, chans, res, out_v, member, fam_ld = A["nodes"], A["chans"], A["res"], A["out_v"], A["member"], A["fam_ld"]
    consumers = {}
    for i, n in enumerate(nodes):
        for b in n["ins"]:
            consumers.setdefault(b, []).append(i)
    view_fam = {n["out"]: n["family"] for n in nodes if n["view"]}
    absorbed, sched = set(), []
    for i, n in enumerate(nodes):
        t = n["op"]["type"]
        if i in absorbed:
            continue
        if t == OP_CONV:
            v, resid, pos = n["out"], [], i
            while v != out_v and len(resid) < 2:
                cs = consumers.get(v, [])
                if len(cs) != 1 or cs[0] in absorbed:
                    break
                a = nodes[cs[0]]
                if a["op"]["type"] != OP_ADD or a["ins"][0] == a["ins"][1]:
                    break
                oi = 1 if a["ins"][0] == v else 0
                other = a["ins"][oi]
                if other == input_blob or other in view_fam:
                    break
                cv, cr = (1.0, 1.0) if a["op"]["plain"] else (a["op"]["coef"][1 - oi], a["op"]["coef"][oi])
                resid.append((other, float(cv), float(cr)))
                absorbed.add(cs[0])
                pos, v = cs[0], a["out"]
            sched.append((pos, {"kind": FOP_CONV, "node": n, "resid": resid, "out": v}))
        elif t == OP_NEAREST:
            sched.append((i, {"kind": FOP_NEAREST, "node": n, "resid": [], "out": n["out"]}))
        elif t == OP_CONCAT:
            if not n["view"]:
                return None
        else:
            return None  # PReLU / PixelShuffle / free-standing adds: not an RRDB-style graph
    sched.sort(key=lambda e: e[0])
    sched = [e[1] for e in sched]
    # 1x1 shortcut fusion: "v = act(conv3x3(X)) + conv1x1(X[:c])" with a bias-less, activation-less 1x1 convolution over a
    # prefix of the same input, consumed by nothing else, becomes one op (the kernel accumulates the 1x1 beside the 3x3)
    if fuse_shortcuts:
        by_out = {s_["out"]: s_ for s_ in sched}
        drop = set()
        for s_ in sched:
            if s_["kind"] != FOP_CONV or len(s_["resid"]) != 1:
                continue
            other, cv, cr = s_["resid"][0]
            b_ = by_out.get(other)
            n_, o_ = s_["node"], s_["node"]["op"]
            if b_ is None or b_["kind"] != FOP_CONV or b_["resid"] or id(b_) in drop:
                continue
            ob = b_["node"]["op"]
            src_a, src_b = n_["ins"][0], b_["node"]["ins"][0]
            same_prefix = (src_a in view_fam and src_b in member and member[src_b] == (view_fam[src_a], 0))
            if (ob["k"] != 1 or ob["act"] != 0 or ob["b_off"] >= 0 or o_["k"] != 3 or o_["cout"] != 32 or ob["cout"] != 32
                    or ob["cin"] > 64 or ob["cin"] % 16 or not same_prefix or len(consumers.get(other, [])) != 1 or s_["out"] == out_v):
                continue
            s_["sc"] = (int(ob["w_off"]), int(ob["cin"]), cv, cr)
            s_["resid"] = []
            drop.add(id(b_))
        sched = [s_ for s_ in sched if id(s_) not in drop]
    if not sched or sched[-1]["out"] != out_v or sched[-1]["kind"] != FOP_CONV or chans[out_v] != 3:
        return None
    made_by = {s["out"]: s for s in sched}

    # ---- which stored form(s) every value needs ------------------------------------------------------------
    need16, need32 = set(), set()
    for s in sched:
        src = s["node"]["ins"][0]
        if src == input_blob:
            if s["kind"] != FOP_CONV or s["node"]["op"]["cin"] != 3:
                return None
        elif src in view_fam:
            if s["kind"] != FOP_CONV:
                return None
        else:
            if src not in made_by:
                return None
            need16.add(src)
        for other, _, _ in s["resid"]:
            if other not in made_by:
                return None
    # A residual source is kept in fp32 only when it sits on a long chain of residual adds (the RRDB trunk: every block
    # output is the next block's residual, 70 deep): rounding those to fp16 would accumulate.  Short chains (the 1x1
    # shortcut -> x2 -> x4 of a dense block, depth 2) are read from the fp16 copy the convolutions use anyway, which
    # saves their fp32 round trip through HBM.
    rdepth = {}
    for s in reversed(sched):
        for other, _, _ in s["resid"]:
            rdepth[other] = max(rdepth.get(other, 0), 1 + rdepth.get(s["out"], 0))
    for s in sched:
        for other, _, _ in s["resid"]:
            (need32 if made_by[other]["kind"] == FOP_CONV and rdepth[other] >= fp32_chain else need16).add(other)
    for v in member:
        if v not in made_by:
            return None
        need16.add(v)
    if any(v in view_fam or v == input_blob for v in consumers if v not in made_by and v not in view_fam and v != input_blob):
        return None

    # ---- storages and their lifetimes over the schedule --------------------------------------------------------
    # storage id: ("fam", f) | ("h", value) | ("f", value);  key = (dtype, pixel stride in channels, resolution)
    def storages_written(s):
        v, out = s["out"], []
        if v in need16:
            out.append(("fam", member[v][0]) if v in member else ("h", v))
        if v in need32:
            out.append(("f", v))
        return out

    def storages_read(s):
        src, out = s["node"]["ins"][0], []
        if src in view_fam:
            out.append(("fam", view_fam[src]))
        elif src != input_blob:
            out.append(("fam", member[src][0]) if src in member else ("h", src))
        for other, _, _ in s["resid"]:
            if other in need32:
                out.append(("f", other))
            else:
                out.append(("fam", member[other][0]) if other in member else ("h", other))
        return out

    def key_of(st):
        if st[0] == "fam":
            v0 = next(v for v, (f, _) in member.items() if f == st[1])
            return (F16, fam_ld[st[1]], res[v0])
        return (F16 if st[0] == "h" else F32, chans[st[1]], res[st[1]])

    first, last = {}, {}
    for j, s in enumerate(sched):
        for st in storages_written(s):
            first.setdefault(st, j)
            last[st] = max(last.get(st, j), j)
        for st in storages_read(s):
            if st not in first:
                return None  # read before written
            last[st] = j
    bufs, free, where = [], {}, {}
    ops = []
    for j, s in enumerate(sched):
        for st in storages_written(s):  # outputs are placed before this op's dying inputs are released: never in place
            if st not in where:
                k = key_of(st)
                if free.get(k):
                    where[st] = free[k].pop()
                else:
                    bufs.append({"channels": k[1], "dtype": k[0], "res": k[2]})
                    where[st] = len(bufs) - 1
        n, v = s["node"], s["out"]
        o = n["op"]
        src = n["ins"][0]
        op = {"type": s["kind"], "res": int(res[v]), "in_buf": -1, "in_off": 0, "cin": int(chans[src]), "k": int(o["k"]),
              "cout": int(chans[v]), "act": int(o["act"]), "slope": float(o["slope"]), "w_off": int(o["w_off"]), "b_off": int(o["b_off"]),
              "nres": len(s["resid"]), "res_buf": [-1, -1], "res_off": [0, 0], "coef_v": [1.0, 1.0], "coef_r": [0.0, 0.0],
              "out16_buf": -1, "out16_off": 0, "out32_buf": -1, "out32_off": 0, "r": int(o["r"]), "final": int(v == out_v),
              "sc_cin": 0, "sc_coef_v": 1.0, "sc_coef_r": 0.0, "sc_w_off": -1}
        if s.get("sc"):
            op["sc_w_off"], op["sc_cin"] = s["sc"][0], s["sc"][1]
            op["sc_coef_v"], op["sc_coef_r"] = float(np.float32(s["sc"][2])), float(np.float32(s["sc"][3]))
        if src in view_fam:
            op["in_buf"] = where[("fam", view_fam[src])]
        elif src != input_blob:
            op["in_buf"], op["in_off"] = (where[("fam", member[src][0])], member[src][1]) if src in member else (where[("h", src)], 0)
        for q, (other, cv, cr) in enumerate(s["resid"]):
            if other in need32:
                op["res_buf"][q] = where[("f", other)]
            elif other in member:
                op["res_buf"][q], op["res_off"][q] = where[("fam", member[other][0])], member[other][1]
            else:
                op["res_buf"][q] = where[("h", other)]
            op["coef_v"][q], op["coef_r"][q] = float(np.float32(cv)), float(np.float32(cr))
        if v in need16:
            op["out16_buf"], op["out16_off"] = (where[("fam", member[v][0])], member[v][1]) if v in member else (where[("h", v)], 0)
        if v in need32:
            op["out32_buf"] = where[("f", v)]
        if not op["final"] and op["out16_buf"] < 0 and op["out32_buf"] < 0:
            return None  # dead value
        ops.append(op)
        for st, lj in last.items():
            if lj == j and st in where and first[st] <= j:
                free.setdefault(key_of(st), []).append(where[st])
    return FusedProgram(ops, bufs, int(res[out_v]), A["weights"], input_blob, output_blob)
