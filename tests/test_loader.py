"""CPU: model loading (upscale_video_b200/ncnn_model.py) -- ncnn text/bin grammar, .b2sr container, family
recognition and blob packing for b2sr_create.  The product loader is checked against the oracle's independent
reader on the packaged models."""
import os
import struct

import numpy as np
import pytest

from conftest import HURR, golden
from oracle import oracle
from upscale_video_b200 import ncnn_model as M


def _tiny_ncnn(tmp_path, nf=8, n_mid=2, r=2, fp16=True):
    rng = np.random.default_rng(3)
    lines, blob = [], b""
    names = iter(range(100))

    def conv(name, bottom, top, cin, cout):
        nonlocal blob
        w = (rng.integers(-64, 64, cout * cin * 9) / 64.0).astype(np.float32)
        if fp16:
            raw = w.astype("<f2").tobytes()
            blob += struct.pack("<I", 0x01306B47) + raw + b"\0" * ((-len(raw)) % 4)
        else:
            blob += struct.pack("<I", 0) + w.astype("<f4").tobytes()
        b = (rng.integers(-32, 32, cout) / 32.0).astype("<f4")
        blob += b.tobytes()
        lines.append("Convolution %s 1 1 %s %s 0=%d 1=3 4=1 5=1 6=%d" % (name, bottom, top, cout, cout * cin * 9))
        return w.reshape(cout, cin, 3, 3), b

    def prelu(name, bottom, top, n):
        nonlocal blob
        s = (rng.integers(0, 32, n) / 64.0).astype("<f4")
        blob += s.tobytes()
        lines.append("PReLU %s 1 1 %s %s 0=%d" % (name, bottom, top, n))

    lines.append("Input input 0 1 input")
    lines.append("Split splitncnn_0 1 2 input input_s0 input_s1")
    x = "input_s1"
    conv("c0", x, "b0", 3, nf)
    prelu("p0", "b0", "a0", nf)
    x = "a0"
    for i in range(n_mid):
        conv("c%d" % (i + 1), x, "b%d" % (i + 1), nf, nf)
        prelu("p%d" % (i + 1), "b%d" % (i + 1), "a%d" % (i + 1), nf)
        x = "a%d" % (i + 1)
    conv("clast", x, "y", nf, 3 * r * r)
    lines.append("PixelShuffle ps 1 1 y ys 0=%d" % r)
    lines.append("Interp up 1 1 input_s0 xs 0=1 1=%d.000000e+00 2=%d.000000e+00" % (r, r))
    lines.append("BinaryOp add 2 1 ys xs output")
    text = "7767517\n%d %d\n" % (len(lines), len(lines) + 2) + "\n".join(lines) + "\n"
    p = tmp_path / "t.param"
    p.write_text(text)
    (tmp_path / "t.bin").write_bytes(blob)
    return str(p), str(tmp_path / "t.bin")


@pytest.mark.parametrize("fp16", [True, False])
def test_parse_ncnn_pair_and_recognise(tmp_path, fp16):
    param, binf = _tiny_ncnn(tmp_path, fp16=fp16)
    g = M.load_ncnn(param, binf)
    d = M.compact_desc(g)
    assert (d.cin, d.nf, d.n_mid, d.scale, d.cout_last, d.input_blob, d.output_blob) == (3, 8, 2, 2, 12, "input", "output")
    d2, blob = M.pack_compact_blob(g)
    assert blob.dtype == np.float32 and blob.size == 8 * 3 * 9 + 16 + 2 * (8 * 8 * 9 + 16) + 12 * 8 * 9 + 12
    # the oracle's own reader sees the same numbers
    layers = oracle.read_ncnn(param, binf)
    convs = [L for L in layers if L["type"] == "Convolution"]
    assert np.array_equal(convs[0]["arrays"]["weight"], g.convs()[0].weights["weight"].astype(np.float32))
    assert np.array_equal(blob[:8 * 3 * 9], convs[0]["arrays"]["weight"].ravel())


def test_bin_must_be_fully_consumed(tmp_path):
    param, binf = _tiny_ncnn(tmp_path)
    with open(binf, "ab") as f:
        f.write(b"\0\0\0\0")
    with pytest.raises(ValueError):
        M.load_ncnn(param, binf)


def test_bad_magic():
    with pytest.raises(ValueError):
        M.parse_param("12345\n1 1\nInput input 0 1 input\n")


def test_b2sr_roundtrip(tmp_path):
    param, binf = _tiny_ncnn(tmp_path, fp16=False)
    g = M.load_ncnn(param, binf)
    out = str(tmp_path / "t.b2sr")
    M.save_b2sr(g, out)
    g2 = M.load_b2sr(out)
    assert [l.name for l in g.layers] == [l.name for l in g2.layers]
    for a, b in zip(g.layers, g2.layers):
        assert a.params == b.params
        for k in a.weights:
            assert np.array_equal(a.weights[k].astype(np.float32), b.weights[k].astype(np.float32))
    # fp32-stored weights that are fp16-exact are narrowed (lossless), like 4x_Compact_Pretrain
    assert g2.convs()[0].weights["weight"].dtype == np.float16


def test_non_compact_graph_is_rejected(tmp_path):
    param, binf = _tiny_ncnn(tmp_path)
    g = M.load_ncnn(param, binf)
    g.layers[-1].params[0] = 2  # BinaryOp mul instead of add
    assert M.compact_desc(g) is None
    with pytest.raises(ValueError):
        M.pack_compact_blob(g)


@pytest.mark.parametrize("stem,expect", [
    ("2x_Compact_Pretrain", (3, 64, 16, 2, 12)),
    ("4x_Compact_Pretrain", (3, 64, 16, 4, 48)),
    (HURR, (3, 24, 8, 1, 3)),
])
def test_packaged_models(model_dir, stem, expect):
    g = M.load_model(model_dir, stem)
    d = M.compact_desc(g)
    assert (d.cin, d.nf, d.n_mid, d.scale, d.cout_last) == expect
    _, blob = M.pack_compact_blob(g)
    # every parameter the reference ships for the Compact family is exactly representable in fp16
    assert np.array_equal(blob.astype(np.float16).astype(np.float32), blob)
    # product loader == oracle's independent reader
    layers = oracle.read_model(model_dir, stem)
    ow = [L["arrays"]["weight"] for L in layers if L["type"] == "Convolution"]
    pw = [l.weights["weight"].astype(np.float32) for l in g.convs()]
    assert len(ow) == len(pw) and all(np.array_equal(a, b) for a, b in zip(ow, pw))


def test_mac_count_matches_survey(model_dir):
    g = M.load_model(model_dir, "2x_Compact_Pretrain")
    assert sum(l.weights["weight"].size for l in g.convs()) == 598464


def _run_program_numpy(prog, x):
    """Execute a compiled op list (ncnn_model.compile_graph) on the CPU with the oracle's layer functions.  Slots are
    H x W x ld arrays; operands are channel slices of them (the strided views of include/b2sr.h b2sr_graph_op)."""
    slots = [None] * prog.n_slots
    slots[prog.in_slot] = np.ascontiguousarray(x, np.float64)
    w = prog.weights
    h0, w0 = x.shape[:2]
    for o in prog.ops:
        ins = [slots[s][:, :, off:off + c] for s, off, c in zip(o["in"][:o["nin"]], o["in_off"], o["in_c"])]
        for a, ld, s in zip(ins, o["in_ld"], o["in"]):
            assert a.shape == (h0 * o["in_res"], w0 * o["in_res"], a.shape[2]) and slots[s].shape[2] == (ld or a.shape[2])
        t = o["type"]
        if t == M.OP_CONV:
            wt = w[o["w_off"]:o["w_off"] + o["cout"] * o["cin"] * o["k"] ** 2].reshape(o["cout"], o["cin"], o["k"], o["k"])
            b = w[o["b_off"]:o["b_off"] + o["cout"]] if o["b_off"] >= 0 else None
            y = oracle.conv(ins[0], wt, b, o["k"] // 2, o["act"], np.array([o["slope"]], np.float32) if o["act"] == 2 else None, "f64")
        elif t == M.OP_PRELU:
            sl = w[o["w_off"]:o["w_off"] + ins[0].shape[2]].astype(np.float64)
            y = np.where(ins[0] < 0, ins[0] * sl, ins[0])
        elif t == M.OP_PIXELSHUFFLE:
            y = oracle.pixelshuffle(ins[0], o["r"])
        elif t == M.OP_NEAREST:
            y = oracle.nearest(ins[0], float(o["r"]), float(o["r"]))
        elif t == M.OP_ADD:
            y = ins[0] + ins[1] if o["plain"] else ins[0] * np.float64(np.float32(o["coef"][0])) + ins[1] * np.float64(np.float32(o["coef"][1]))
        elif t == M.OP_CONCAT:
            y = np.concatenate(ins, axis=2)
        assert y.shape == (h0 * o["out_res"], w0 * o["out_res"], o["out_c"])
        ld = o["out_ld"] or o["out_c"]
        if slots[o["out"]] is None or slots[o["out"]].shape != (y.shape[0], y.shape[1], ld) or not o["out_ld"]:
            # a fresh (or recycled) slot; poison it so that reading a slice nobody wrote shows up
            slots[o["out"]] = np.full((y.shape[0], y.shape[1], ld), np.nan)
        slots[o["out"]][:, :, o["out_off"]:o["out_off"] + o["out_c"]] = y
    out = slots[prog.out_slot]
    assert not np.isnan(out).any()
    return out


@pytest.mark.parametrize("stem", ["2x_Compact_Pretrain", "4x_Valar_v1"])
def test_compiled_program_equals_graph_interpreter(model_dir, stem):
    """compile_graph (Split aliasing, slot recycling, in-place elementwise ops) must not change the function: run the op
    list with the oracle's layers and compare with the oracle's own graph interpreter (bit-exact in f64)."""
    if not os.path.exists(os.path.join(model_dir, stem + ".b2sr")):
        pytest.skip("model not packaged")
    g = M.load_model(model_dir, stem)
    x = np.random.default_rng(1).random((12, 14, 3))
    ref = oracle.run_graph(oracle.read_model(model_dir, stem), x, "f64")
    for views in (True, False):
        prog = M.compile_graph(g, views=views)
        assert prog.n_slots <= 12 and prog.scale == int(stem[0])
        n_concat = sum(o["type"] == M.OP_CONCAT for o in prog.ops)
        if stem == "4x_Valar_v1":  # every dense-block Concat is a prefix chain -> all of them become views
            assert n_concat == (0 if views else 4 * 3 * 23)
        got = _run_program_numpy(prog, x)
        assert got.shape == ref.shape and np.array_equal(got, ref)


def test_fused_program_equals_graph_interpreter(model_dir):
    """compile_fused (adds folded into the convolution that feeds them, ops scheduled where their last add stood,
    fp16/fp32 buffer recycling, channel-slice views instead of Concat) must not change the function: the float64
    interpretation of the fused op list equals the oracle's graph interpreter bit for bit; with the device's storage
    types (fp16 activations, fp32 residual trunk) it stays within 1 LSB of the golden output."""
    import fused_emulator as FE
    stem = "4x_Valar_v1"
    if not os.path.exists(os.path.join(model_dir, stem + ".b2sr")):
        pytest.skip("model not packaged")
    g = M.load_model(model_dir, stem)
    prog = M.compile_fused(g)
    assert prog is not None and prog.scale == 4 and len(prog.bufs) <= 16
    convs = [o for o in prog.ops if o["type"] == M.FOP_CONV]
    n_sc = sum(o["sc_cin"] > 0 for o in convs)
    assert n_sc == 23 * 3 and all(o["sc_cin"] == 64 and o["cin"] == 96 for o in convs if o["sc_cin"])  # every 1x1 shortcut rides on x2's conv
    assert len(convs) + n_sc == len(g.convs()) and sum(o["type"] == M.FOP_NEAREST for o in prog.ops) == 2
    assert sum(o["nres"] for o in convs) + n_sc == 23 * 3 * 3 + 23 + 1  # every BinaryOp / Eltwise of the graph is folded
    plain = M.compile_fused(g, fuse_shortcuts=False, fp32_chain=1)  # the unfused, all-fp32-residual lowering stays available
    assert sum(o["type"] == M.FOP_CONV for o in plain.ops) == len(g.convs()) and not any(o["sc_cin"] for o in plain.ops)
    assert prog.ops[-1]["final"] == 1 and sum(o["final"] for o in prog.ops) == 1
    for o in convs:  # what the tcgen05 kernel needs from a view: 16-byte aligned channel offsets, <= 3 groups of 64
        assert o["in_off"] % 8 == 0 and o["out16_off"] % 8 == 0 and o["cin"] <= 192
        assert o["out16_buf"] != o["in_buf"] or o["out16_off"] >= o["in_off"] + o["cin"]  # never overwrites what it reads
        assert all(rb not in (o["out16_buf"], o["out32_buf"]) or prog.bufs[rb]["channels"] > o["cout"] for rb in o["res_buf"][:o["nres"]])
    x = np.random.default_rng(1).random((12, 14, 3))
    ref = oracle.run_graph(oracle.read_model(model_dir, stem), x, "f64")
    assert np.array_equal(FE.run_fused(prog, x, exact=True), ref)
    gold = golden("valar4x_crop")
    dev = FE.run_fused(prog, gold["x"], exact=False)
    u8 = oracle.saturate_u8(dev.astype(np.float32) * np.float32(255))
    d = np.abs(u8.astype(int) - gold["y"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.03


def test_compact_graphs_do_not_lower_to_fused(model_dir):
    """PReLU / PixelShuffle graphs are served by the Compact kernels; compile_fused must decline them."""
    assert M.compile_fused(M.load_model(model_dir, "2x_Compact_Pretrain")) is None


def test_valar_persistent_segments_plan(model_dir):
    """Host-only: how the fused 4x_Valar_v1 program is cut into persistent segments (b2sr_fused_describe_segments): one
    segment per RRDB (15 convolutions = 18 stages on a 148-SM part, 8 bands wide), dense-block slices and the fp32 trunk
    copies between the blocks in rings, the segment's input and output in frame buffers, every stage gated by the op that
    wrote the newest slice it reads and back-pressured by the last reader of the ring it writes."""
    import os
    from upscale_video_b200 import engine as E
    if not os.path.exists(os.path.join(model_dir, "4x_Valar_v1.b2sr")):
        pytest.skip("4x_Valar_v1 not packaged")
    prog = M.compile_fused(M.load_model(model_dir, "4x_Valar_v1"))
    segs = E.fused_segments(prog, sms=148)
    assert len(segs) == 23
    for k, s in enumerate(segs):
        assert (s["op_begin"], s["op_end"]) == (1 + 15 * k, 15 + 15 * k) and len(s["stages"]) == 18
        st = s["stages"]
        assert [x["op"] - s["op_begin"] for x in st] == [0, 1, 2, 3, 4, 4, 5, 6, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14]
        assert st[0]["in_inst"] == -1 and st[0]["gate_op"] == -1          # x1 of the first block reads the frame buffer
        assert st[1]["grp_ring"] == [0, 1, 0] and st[6]["grp_ring"] == [1, 0, 0]
        for x in st[1:]:
            assert x["gate_op"] == x["op"] - 1                            # a chain: each op waits for the one before it
        assert all(x["out16_inst"] == -1 and x["out32_inst"] == -1 and x["bp_op"] == -1 for x in st[16:])  # RRDB output: frames
        for x in st[:16]:
            assert x["out16_inst"] >= 0 and s["inst"][x["out16_inst"]]["last_reader"] == x["bp_op"]
    small = E.fused_segments(prog, sms=48)  # room for 6 stages of 8 bands: one dense block (x1..x4 + the two halves of x5) per segment
    assert len(small) == 69 and all(len(s["stages"]) == 6 and s["op_end"] - s["op_begin"] == 4 for s in small)
