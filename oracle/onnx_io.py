"""Hand-rolled ONNX (protobuf wire format) writer + reader.  TEST INFRASTRUCTURE ONLY.

The reference consumes `models/<model>/noiseN_scaleSx.onnx` (src/main.cpp:201-204,
src/tensorrt/img2img_build.cpp:87-92).  Neither the real files nor the `onnx` package exist in
this image, so the oracle emits ONNX files for its seeded synthetic models with this writer; the
product's C++ importer (csrc/onnx_reader.cpp) reads the same wire format, and
cv2.dnn.readNetFromONNX validates the emitted files (tests/test_onnx.py).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np

FLOAT, INT64 = 1, 7
A_FLOAT, A_INT, A_STRING, A_TENSOR, A_FLOATS, A_INTS = 1, 2, 3, 4, 6, 7


# ---- wire primitives ----------------------------------------------------------------------
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def _f_varint(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(v)


def _f_bytes(field: int, b: bytes) -> bytes:
    return _key(field, 2) + _varint(len(b)) + b


def _f_str(field: int, s: str) -> bytes:
    return _f_bytes(field, s.encode())


def _f_float(field: int, f: float) -> bytes:
    return _key(field, 5) + struct.pack("<f", f)


# ---- message builders ----------------------------------------------------------------------
def tensor(name: str, arr: np.ndarray) -> bytes:
    if arr.dtype == np.int64:
        dt = INT64
    else:
        arr = arr.astype(np.float32)
        dt = FLOAT
    body = _f_bytes(1, b"".join(_varint(int(d)) for d in arr.shape)) if arr.ndim else b""
    body += _f_varint(2, dt) + _f_str(8, name) + _f_bytes(9, np.ascontiguousarray(arr).tobytes())
    return body


def attr_i(name: str, v: int) -> bytes:
    return _f_str(1, name) + _f_varint(3, v) + _f_varint(20, A_INT)


def attr_f(name: str, v: float) -> bytes:
    return _f_str(1, name) + _f_float(2, v) + _f_varint(20, A_FLOAT)


def attr_ints(name: str, vs: Sequence[int]) -> bytes:
    return _f_str(1, name) + _f_bytes(8, b"".join(_varint(int(v)) for v in vs)) + _f_varint(20, A_INTS)


def node(op: str, inputs: Sequence[str], outputs: Sequence[str], name: str = "", attrs: Sequence[bytes] = ()) -> bytes:
    b = b"".join(_f_str(1, i) for i in inputs) + b"".join(_f_str(2, o) for o in outputs)
    b += _f_str(3, name or outputs[0]) + _f_str(4, op)
    b += b"".join(_f_bytes(5, a) for a in attrs)
    return b


def value_info(name: str, dims: Sequence) -> bytes:
    shape = b""
    for d in dims:
        dim = _f_str(2, d) if isinstance(d, str) else _f_varint(1, int(d))
        shape += _f_bytes(1, dim)
    ttype = _f_varint(1, FLOAT) + _f_bytes(2, shape)
    return _f_str(1, name) + _f_bytes(2, _f_bytes(1, ttype))


def model(nodes: Sequence[bytes], inits: Sequence[bytes], inputs: Sequence[bytes], outputs: Sequence[bytes],
          name: str = "w2x", opset: int = 13) -> bytes:
    g = b"".join(_f_bytes(1, n) for n in nodes) + _f_str(2, name)
    g += b"".join(_f_bytes(5, t) for t in inits)
    g += b"".join(_f_bytes(11, i) for i in inputs) + b"".join(_f_bytes(12, o) for o in outputs)
    return (_f_varint(1, 7) + _f_str(2, "w2x-b200-oracle") + _f_bytes(7, g)
            + _f_bytes(8, _f_str(1, "") + _f_varint(2, opset)))


# ---- graph emitter for the oracle's torch modules --------------------------------------------
class GraphBuilder:
    def __init__(self):
        self.nodes: List[bytes] = []
        self.inits: List[bytes] = []
        self.n = 0

    def _new(self, hint: str) -> str:
        self.n += 1
        return f"{hint}_{self.n}"

    def init(self, name: str, arr) -> str:
        self.inits.append(tensor(name, np.asarray(arr)))
        return name

    def conv(self, x: str, m, name: str) -> str:
        w = self.init(name + ".weight", m.weight.detach().numpy())
        ins = [x, w]
        if m.bias is not None:
            ins.append(self.init(name + ".bias", m.bias.detach().numpy()))
        y = self._new("conv")
        k = m.kernel_size
        self.nodes.append(node("Conv", ins, [y], name, [
            attr_ints("dilations", [1, 1]), attr_i("group", 1), attr_ints("kernel_shape", k),
            attr_ints("pads", [m.padding[0], m.padding[1], m.padding[0], m.padding[1]]),
            attr_ints("strides", m.stride)]))
        return y

    def conv_transpose(self, x: str, m, name: str) -> str:
        w = self.init(name + ".weight", m.weight.detach().numpy())
        ins = [x, w]
        if m.bias is not None:
            ins.append(self.init(name + ".bias", m.bias.detach().numpy()))
        y = self._new("convt")
        self.nodes.append(node("ConvTranspose", ins, [y], name, [
            attr_ints("dilations", [1, 1]), attr_i("group", 1), attr_ints("kernel_shape", m.kernel_size),
            attr_ints("pads", [m.padding[0], m.padding[1], m.padding[0], m.padding[1]]),
            attr_ints("strides", m.stride)]))
        return y

    def unary(self, op: str, x: str, attrs=()) -> str:
        y = self._new(op.lower())
        self.nodes.append(node(op, [x], [y], attrs=attrs))
        return y

    def lrelu(self, x: str, alpha: float = 0.1) -> str:
        if getattr(self, "mutate", "") == "alpha" and not getattr(self, "_alpha_done", False) and self.n > 40:
            self._alpha_done = True
            alpha = 0.2
        return self.unary("LeakyRelu", x, [attr_f("alpha", alpha)])

    def binary(self, op: str, a: str, b: str) -> str:
        y = self._new(op.lower())
        self.nodes.append(node(op, [a, b], [y]))
        return y

    def crop(self, x: str, p: int) -> str:
        """F.pad(x, (-p,)*4) as Slice over axes 2,3."""
        if getattr(self, "mutate", "") == "crop" and p == 16:
            p = 5
        if getattr(self, "mutate", "") == "pad":
            pads = self.init(self._new("pads"), np.array([0, 0, -p, -p, 0, 0, -p, -p], np.int64))
            y = self._new("pad")
            self.nodes.append(node("Pad", [x, pads], [y]))
            return y
        big = 2 ** 31 - 1
        s = self.init(self._new("starts"), np.array([p, p], np.int64))
        e = self.init(self._new("ends"), np.array([-p, -p], np.int64))
        a = self.init(self._new("axes"), np.array([2, 3], np.int64))
        y = self._new("slice")
        self.nodes.append(node("Slice", [x, s, e, a], [y]))
        return y

    def clip01(self, x: str) -> str:
        lo = self.init(self._new("min"), np.array(0.0, np.float32))
        hi = self.init(self._new("max"), np.array(1.0, np.float32))
        y = self._new("clip")
        self.nodes.append(node("Clip", [x, lo, hi], [y]))
        return y


def _emit_unetconv(g: GraphBuilder, x: str, m, name: str) -> str:
    x = g.lrelu(g.conv(x, m.conv[0], name + ".conv.0"))
    x = g.lrelu(g.conv(x, m.conv[2], name + ".conv.2"))
    if m.se is not None:
        z = g.unary("GlobalAveragePool", x)
        z = g.unary("Relu", g.conv(z, m.se.conv1, name + ".se.conv1"))
        z = g.unary("Sigmoid", g.conv(z, m.se.conv2, name + ".se.conv2"))
        x = g.binary("Mul", x, z)
    return x


def _emit_bottom(g: GraphBuilder, x: str, m, name: str) -> str:
    import torch.nn as nn
    return g.conv_transpose(x, m, name) if isinstance(m, nn.ConvTranspose2d) else g.conv(x, m, name)


def export_cunet(model_t, path: str | None = None, mutate: str = "") -> bytes:
    """Emit oracle.models.CUNet / UpCUNet as ONNX (NCHW, dynamic batch/H/W).  `mutate` writes a deliberately different graph with
    the same weight shapes (tests of the importer's topology checks): "alpha" (LeakyReLU slope 0.2 once), "crop" (a skip cropped
    by 5), "noclip", "pad" (crops as negative Pad nodes, which is also valid), "shuffle" (initializers in another order)."""
    g = GraphBuilder()
    g.mutate = mutate
    u1, u2 = model_t.unet1, model_t.unet2
    x = "x"
    # unet1
    x1 = _emit_unetconv(g, x, u1.conv1, "unet1.conv1")
    x2 = g.lrelu(g.conv(x1, u1.conv1_down, "unet1.conv1_down"))
    x2 = _emit_unetconv(g, x2, u1.conv2, "unet1.conv2")
    x2 = g.lrelu(g.conv_transpose(x2, u1.conv2_up, "unet1.conv2_up"))
    x1 = g.crop(x1, 4)
    x3 = g.lrelu(g.conv(g.binary("Add", x1, x2), u1.conv3, "unet1.conv3"))
    z1 = _emit_bottom(g, x3, u1.conv_bottom, "unet1.conv_bottom")
    # unet2
    x1 = _emit_unetconv(g, z1, u2.conv1, "unet2.conv1")
    x2 = g.lrelu(g.conv(x1, u2.conv1_down, "unet2.conv1_down"))
    x2 = _emit_unetconv(g, x2, u2.conv2, "unet2.conv2")
    x3 = g.lrelu(g.conv(x2, u2.conv2_down, "unet2.conv2_down"))
    x3 = _emit_unetconv(g, x3, u2.conv3, "unet2.conv3")
    x3 = g.lrelu(g.conv_transpose(x3, u2.conv3_up, "unet2.conv3_up"))
    x2 = g.crop(x2, 4)
    x4 = _emit_unetconv(g, g.binary("Add", x2, x3), u2.conv4, "unet2.conv4")
    x4 = g.lrelu(g.conv_transpose(x4, u2.conv4_up, "unet2.conv4_up"))
    x1 = g.crop(x1, 16)
    x5 = g.lrelu(g.conv(g.binary("Add", x1, x4), u2.conv5, "unet2.conv5"))
    z2 = _emit_bottom(g, x5, u2.conv_bottom, "unet2.conv_bottom")
    y = g.binary("Add", g.crop(z1, 20), z2)
    if mutate != "noclip":
        y = g.clip01(y)
    g.nodes.append(node("Identity", [y], ["y"]))
    if mutate == "shuffle":
        import random
        random.Random(1).shuffle(g.inits)
    blob = model(g.nodes, g.inits,
                 [value_info("x", ["batch", 3, "height", "width"])],
                 [value_info("y", ["batch", 3, "out_height", "out_width"])])
    if path:
        with open(path, "wb") as f:
            f.write(blob)
    return blob


# ---- reader (for tests: round-trips the writer; mirrors csrc/onnx_reader.cpp) ---------------------
def _read_varint(b: bytes, i: int) -> Tuple[int, int]:
    v = 0
    s = 0
    while True:
        c = b[i]
        i += 1
        v |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return v, i


def _fields(b: bytes):
    i = 0
    while i < len(b):
        k, i = _read_varint(b, i)
        f, wt = k >> 3, k & 7
        if wt == 0:
            v, i = _read_varint(b, i)
        elif wt == 2:
            n, i = _read_varint(b, i)
            v = b[i:i + n]
            i += n
        elif wt == 5:
            v = b[i:i + 4]
            i += 4
        elif wt == 1:
            v = b[i:i + 8]
            i += 8
        else:
            raise ValueError(f"wire type {wt}")
        yield f, wt, v


def _parse_tensor(b: bytes):
    dims: List[int] = []
    dt, name, raw = FLOAT, "", b""
    floats: List[float] = []
    for f, wt, v in _fields(b):
        if f == 1:
            if wt == 2:
                j = 0
                while j < len(v):
                    d, j = _read_varint(v, j)
                    dims.append(d)
            else:
                dims.append(v)
        elif f == 2:
            dt = v
        elif f == 8:
            name = v.decode()
        elif f == 9:
            raw = v
        elif f == 4:
            floats += list(struct.unpack(f"<{len(v)//4}f", v)) if wt == 2 else [struct.unpack("<f", v)[0]]
    if dt == FLOAT:
        arr = np.frombuffer(raw, np.float32) if raw else np.array(floats, np.float32)
    elif dt == INT64:
        arr = np.frombuffer(raw, np.int64)
    else:
        raise ValueError(f"dtype {dt}")
    return name, arr.reshape(dims)


def read_model(blob: bytes):
    """Returns (nodes, initializers): nodes = [(op_type, name, inputs, outputs)], initializers = {name: array}."""
    graph = None
    for f, wt, v in _fields(blob):
        if f == 7:
            graph = v
    assert graph is not None, "no graph"
    nodes = []
    inits: Dict[str, np.ndarray] = {}
    for f, wt, v in _fields(graph):
        if f == 1:
            ins, outs, name, op = [], [], "", ""
            for nf, nwt, nv in _fields(v):
                if nf == 1:
                    ins.append(nv.decode())
                elif nf == 2:
                    outs.append(nv.decode())
                elif nf == 3:
                    name = nv.decode()
                elif nf == 4:
                    op = nv.decode()
            nodes.append((op, name, ins, outs))
        elif f == 5:
            n, a = _parse_tensor(v)
            inits[n] = a
    return nodes, inits


# ---- SwinUNet emitter ---------------------------------------------------------------------------------------------------
# Weight-bearing nodes (Conv, LayerNormalization, MatMul + Add(bias), Add(relative position bias)) are emitted exactly as a
# constant-folded torch export carries them (MatMul weights as [in, out]; the relative-position bias pre-gathered to
# [1, heads, 36, 36]).  The data-movement nodes between them (roll / window partition / head split) are emitted as a
# best-effort Reshape/Transpose/Softmax skeleton: the real nunif exports are unobtainable here, no generic ONNX
# runtime in the image can execute window attention graphs, and the product's importer keys on the weight-bearing nodes only.
def _emit_linear(g: GraphBuilder, x: str, m, name: str) -> str:
    if getattr(g, "variant", "") == "decomposed":
        # older / un-folded exports: Linear as Gemm with the torch [out, in] weight and transB = 1
        w = g.init(name + ".weight", m.weight.detach().numpy())
        ins = [x, w]
        if m.bias is not None:
            ins.append(g.init(name + ".bias", m.bias.detach().numpy()))
        y = g._new("gemm")
        g.nodes.append(node("Gemm", ins, [y], name + ".matmul", [attr_f("alpha", 1.0), attr_f("beta", 1.0), attr_i("transB", 1)]))
        return y
    w = g.init(name + ".weight_t", m.weight.detach().numpy().T.copy())  # [in, out]
    y = g._new("matmul")
    g.nodes.append(node("MatMul", [x, w], [y], name + ".matmul"))
    if m.bias is not None:
        b = g.init(name + ".bias", m.bias.detach().numpy())
        y2 = g._new("addb")
        g.nodes.append(node("Add", [y, b], [y2], name + ".add"))
        y = y2
    return y


def _emit_layernorm(g: GraphBuilder, x: str, m, name: str) -> str:
    s = g.init(name + ".weight", m.weight.detach().numpy())
    b = g.init(name + ".bias", m.bias.detach().numpy())
    if getattr(g, "variant", "") == "decomposed":
        # opset < 17: ReduceMean -> Sub -> Pow -> ReduceMean -> Add(eps) -> Sqrt -> Div -> Mul(gamma) -> Add(beta)
        ax = [attr_ints("axes", [-1]), attr_i("keepdims", 1)]
        mean = g.unary("ReduceMean", x, ax)
        d = g.binary("Sub", x, mean)
        sq = g.binary("Pow", d, g.init(g._new("two"), np.array(2.0, np.float32)))
        var = g.unary("ReduceMean", sq, ax)
        std = g.unary("Sqrt", g.binary("Add", var, g.init(g._new("eps"), np.array(float(m.eps), np.float32))))
        y = g._new("ln")
        g.nodes.append(node("Add", [g.binary("Mul", g.binary("Div", d, std), s), b], [y], name))  # carries the LayerNorm's name
        return y
    y = g._new("ln")
    g.nodes.append(node("LayerNormalization", [x, s, b], [y], name, [attr_i("axis", -1), attr_f("epsilon", float(m.eps))]))
    return y


def _emit_swin_block(g: GraphBuilder, x: str, blk, name: str) -> str:
    h = _emit_layernorm(g, x, blk.norm1, name + ".norm1")
    qkv = _emit_linear(g, h, blk.attn.qkv, name + ".attn.qkv")
    q = g.unary("Transpose", g.unary("Reshape", qkv))   # skeleton: roll + window partition + head split
    scores = g.binary("MatMul", q, q)
    bias = blk.attn.get_relative_position_bias().detach().numpy()  # [1, heads, N, N]
    bname = g.init(name + ".attn.relative_position_bias", bias)
    sb = g._new("add")
    g.nodes.append(node("Add", [scores, bname], [sb], name + ".attn.bias_add"))
    scores = sb
    p = g.unary("Softmax", scores, [attr_i("axis", -1)])
    o = g.unary("Reshape", g.unary("Transpose", g.binary("MatMul", p, q)))
    o = _emit_linear(g, o, blk.attn.proj, name + ".attn.proj")
    x = g.binary("Add", x, o)
    h = _emit_layernorm(g, x, blk.norm2, name + ".norm2")
    h = _emit_linear(g, h, blk.mlp[0], name + ".mlp.0")
    h = g.binary("Mul", h, g.unary("Erf", h))  # skeleton of the erf-GELU decomposition
    h = _emit_linear(g, h, blk.mlp[3], name + ".mlp.3")
    return g.binary("Add", x, h)


def export_swin(model_t, path: str | None = None, variant: str = "") -> bytes:
    """Emit oracle.models.SwinUNet as ONNX (see the note above about the skeleton nodes).  variant "decomposed" writes what an
    older exporter would: LayerNorm as its primitive ops, Linear as Gemm(transB = 1), initializers in shuffled order, opset 13."""
    g = GraphBuilder()
    g.variant = variant
    x = g.lrelu(g.conv("x", model_t.patch[0], "patch.0"))
    x = g.lrelu(g.conv(x, model_t.patch[2], "patch.2"))
    x = g.unary("Transpose", g.crop(x, 6), [attr_ints("perm", [0, 2, 3, 1])])

    def blocks(x, stage, name):
        for i, blk in enumerate(stage.block):
            x = _emit_swin_block(g, x, blk, f"{name}.block.{i}")
        return x

    def down(x, m, name):
        x = g.unary("Transpose", x, [attr_ints("perm", [0, 3, 1, 2])])
        x = g.conv(x, m.conv, name + ".conv")
        return g.unary("Transpose", x, [attr_ints("perm", [0, 2, 3, 1])])

    def up(x, m, name):
        x = _emit_linear(g, x, m.proj, name + ".proj")
        x = g.unary("DepthToSpace", g.unary("Transpose", x, [attr_ints("perm", [0, 3, 1, 2])]), [attr_i("blocksize", 2)])
        return g.unary("Transpose", x, [attr_ints("perm", [0, 2, 3, 1])])

    x3 = blocks(x, model_t.swin1, "swin1")
    x4 = blocks(down(x3, model_t.down1, "down1"), model_t.swin2, "swin2")
    x5 = blocks(down(x4, model_t.down2, "down2"), model_t.swin3, "swin3")
    x = blocks(g.binary("Add", up(x5, model_t.up2, "up2"), x4), model_t.swin4, "swin4")
    x = blocks(g.binary("Add", up(x, model_t.up1, "up1"), x3), model_t.swin5, "swin5")
    if model_t.up0 is not None:
        x = up(x, model_t.up0, "up0")
    x = _emit_linear(g, x, model_t.to_image.proj, "to_image.proj")
    x = g.unary("Transpose", x, [attr_ints("perm", [0, 3, 1, 2])])
    if model_t.to_image.scale > 1:
        x = g.unary("DepthToSpace", x, [attr_i("blocksize", model_t.to_image.scale)])
    y = g.clip01(x)
    g.nodes.append(node("Identity", [y], ["y"]))
    if variant == "decomposed":
        import random
        random.Random(0).shuffle(g.inits)
    blob = model(g.nodes, g.inits, [value_info("x", ["batch", 3, "height", "width"])],
                 [value_info("y", ["batch", 3, "out_height", "out_width"])], name="w2x_swin_unet", opset=13 if variant == "decomposed" else 17)
    if path:
        with open(path, "wb") as f:
            f.write(blob)
    return blob
