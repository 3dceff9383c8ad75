"""oracle/np_oracle.py -- TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's per-node semantics.

An independent second oracle (SURVEY.md section 8c "Oracle plan (2)"): it interprets an OnnxStream model directory
op by op in float64 (optionally rounding every node output to fp16, the reference's m_use_fp16_arithmetic storage
rule) and is used to (a) arbitrate wherever oracle/_ref computed a result through oracle/xnn_shim.cpp rather than real
XNNPACK (softmax, dynamic matmul, transpose, f16 Conv/FC), (b) generate the golden vectors under tests/golden/.
Pinned by tests/test_cpu.py against oracle/_ref (the reference's own code run here) on every tiny architecture.

Each handler cites the reference branch it restates (src/onnxstream.cpp).  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline leg may import this module; the product never does.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import numpy as np

try:
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


def _parse_tensor(s: str):
    if not s:
        return None
    name, rest = s.split("(", 1)
    rest = rest[:-1]
    wtype, scale, zp = None, 0.0, 0
    if ":" in rest:
        wtype, shape = rest.split(":", 1)
        if wtype.startswith("uint8["):
            sc, z = wtype[6:-1].split(",")
            scale, zp, wtype = float(sc), int(z), "uint8"
    else:
        shape = rest
    dims = tuple(int(x) for x in shape.split(",") if x)
    return dict(name=name, wtype=wtype, shape=dims, scale=scale, zp=zp)


def parse_model(text: str):
    """src/onnxstream.cpp:2445-2616."""
    ops = []
    for line in text.replace("\r", "\n").split("\n"):
        if not line:
            continue
        sec = line.split("*")
        name, typ = sec[0].split(":")
        ins = [_parse_tensor(t) for t in sec[1][len("input:"):].split(";")]
        outs = [_parse_tensor(t) for t in sec[2][len("output:"):].split(";")]
        attrs = dict(kv.split(":") for kv in sec[3].split(";")) if len(sec) == 4 else {}
        ops.append(dict(name=name, type=typ, inputs=ins, outputs=outs, attrs=attrs))
    return ops


_NP = {"float32": np.float32, "float16": np.float16, "int64": np.int64, "uint8": np.uint8}


class NumpyOracle:
    def __init__(self, model_dir: Optional[str] = None, text: Optional[str] = None, blobs: Optional[Dict[str, np.ndarray]] = None, fp16: bool = False):
        self.dir = model_dir
        self.text = text if text is not None else open(os.path.join(model_dir, "model.txt")).read()
        self.blobs = blobs
        self.fp16 = fp16
        self.ops = parse_model(self.text)

    # -- weights (src/onnxstream.cpp:2662-2760, 2885-2909) --
    def _weight(self, t):
        fn = t["name"]
        shape = t["shape"]
        if fn.endswith("_nchw.bin"):
            fn = fn[:-len("_nchw.bin")] + "_nhwc.bin"
            o, i, kh, kw = shape
            shape = (o, kh, kw, i)
        if self.blobs is not None:
            a = np.asarray(self.blobs[fn]).reshape(shape)
        else:
            a = np.fromfile(os.path.join(self.dir, fn), dtype=_NP[t["wtype"]]).reshape(shape)
        if t["wtype"] == "int64":
            return a.astype(np.int64)
        if t["wtype"] == "uint8":
            a = (a.astype(np.float64) - t["zp"]) * np.float32(t["scale"]).astype(np.float64)
        a = a.astype(np.float64)
        return self._store(a)

    def _store(self, a):
        """Storage rounding: fp16 in fp16 mode, fp32 otherwise (values are carried as float64 between nodes)."""
        if a.dtype == np.int64:
            return a
        return a.astype(np.float16).astype(np.float64) if self.fp16 else a.astype(np.float32).astype(np.float64)

    def run(self, inputs: Dict[str, np.ndarray], extra_outputs=()) -> Dict[str, np.ndarray]:
        env: Dict[str, np.ndarray] = {}
        for k, v in inputs.items():
            env[k] = v.astype(np.int64) if v.dtype == np.int64 else self._store(v.astype(np.float64))
        refs: Dict[str, int] = {}
        for op in self.ops:
            for t in op["inputs"]:
                if t and t["wtype"] is None:
                    refs[t["name"]] = refs.get(t["name"], 0) + 1
        for n in extra_outputs:
            refs[n] = refs.get(n, 0) + 1
        for op in self.ops:
            xs = []
            for t in op["inputs"]:
                if t is None:
                    xs.append(None)
                elif t["wtype"] is not None:
                    xs.append(self._weight(t))
                else:
                    xs.append(env[t["name"]])
            ys = self._exec(op, xs)
            if not isinstance(ys, (list, tuple)):
                ys = [ys]
            for t, y in zip(op["outputs"], ys):
                assert tuple(y.shape) == tuple(t["shape"]), (op["type"], op["name"], y.shape, t["shape"])
                env[t["name"]] = self._store(y)
            for t in op["inputs"]:
                if t and t["wtype"] is None:
                    refs[t["name"]] -= 1
                    if refs[t["name"]] == 0:
                        del env[t["name"]]
        return {k: (v if v.dtype == np.int64 else v.astype(np.float32)) for k, v in env.items()}

    # -- per-op semantics --
    def _exec(self, op, x):
        t, a = op["type"], op["attrs"]
        if t == "Conv":  # src/onnxstream.cpp:4494-4707, 1292-1534 (padding re-symmetrised from the sums)
            inp, w = x[0], x[1]            # inp NCHW, w OHWI
            b = x[2] if len(x) > 2 and x[2] is not None else None
            pads = [int(v) for v in a["pads"].split(",")]
            s = int(a["strides"].split(",")[0])
            kh, kw = w.shape[1], w.shape[2]
            ph, pw = pads[0] + pads[2], pads[1] + pads[3]
            pt, pl = ph // 2, pw // 2
            _, c, h, wd = inp.shape
            ho, wo = (h + ph - kh) // s + 1, (wd + pw - kw) // s + 1
            xp = np.zeros((c, h + ph, wd + pw))
            xp[:, pt:pt + h, pl:pl + wd] = inp[0]
            cols = np.empty((ho * wo, kh * kw * c))
            k = 0
            for ky in range(kh):
                for kx in range(kw):
                    patch = xp[:, ky:ky + s * (ho - 1) + 1:s, kx:kx + s * (wo - 1) + 1:s]   # [c, ho, wo]
                    cols[:, k * c:(k + 1) * c] = patch.reshape(c, -1).T
                    k += 1
            y = cols @ w.reshape(w.shape[0], -1).T
            if b is not None:
                y = y + b
            return y.T.reshape(1, w.shape[0], ho, wo)
        if t == "MatMul":  # src/onnxstream.cpp:5669-5861
            p, q = x[0], x[1]
            if p.ndim == 4 and q.ndim == 4 and p.shape[1] != q.shape[1] and q.shape[1] and p.shape[1] % q.shape[1] == 0:
                # grouped KV heads: only reachable through the ScaledDotProductAttention rewrite (src/onnxstream.cpp:3643-3695,
                # 7767-7882), where query head h reads KV head h // (Hq / Hkv)
                q = np.repeat(q, p.shape[1] // q.shape[1], axis=1)
            return np.matmul(p, q)
        if t == "Gemm":  # src/onnxstream.cpp:4300-4375
            return x[0] @ x[1] + x[2]
        if t in ("Add", "Sub", "Mul", "Div"):  # src/onnxstream.cpp:5056-5175, 5394-5477, 3906-4000, 5605-5668
            p, q = x
            if p.dtype == np.int64 and q.dtype == np.int64:
                if t == "Add": return p + q
                if t == "Sub": return p - q
                if t == "Mul": return (p.astype(np.float32) * q.astype(np.float32)).astype(np.int64)
                return (p.astype(np.float32) / q.astype(np.float32)).astype(np.int64)
            p, q = p.astype(np.float64), q.astype(np.float64)
            if t == "Add": return p + q
            if t == "Sub": return p - q
            if t == "Mul": return p * q
            return p / q
        if t == "Sigmoid":  # src/onnxstream.cpp:4376-4493
            return 1.0 / (1.0 + np.exp(-x[0]))
        if t == "Erf": return _erf(x[0])
        if t == "Sqrt": return np.sqrt(x[0])
        if t == "Sin": return np.sin(x[0])
        if t == "Cos": return np.cos(x[0])
        if t == "Neg": return -x[0]
        if t == "Pow": return np.power(x[0], float(np.asarray(x[1]).reshape(-1)[0]))  # src/onnxstream.cpp:5478-5604
        if t == "Reshape":  # src/onnxstream.cpp:4708-4787
            shp = [int(v) for v in x[1]]
            shp = [x[0].shape[i] if v == 0 else v for i, v in enumerate(shp)]
            return x[0].reshape(shp)
        if t == "Unsqueeze":  # src/onnxstream.cpp:3859-3905
            y = x[0]
            rank = y.ndim + len(x[1])
            for ax in sorted(int(v) % rank for v in x[1]):
                y = np.expand_dims(y, ax)
            return y
        if t == "Squeeze":
            return np.squeeze(x[0], tuple(int(v) for v in x[1])) if len(x) > 1 and x[1] is not None else np.squeeze(x[0])
        if t == "Flatten":
            ax = int(a.get("axis", 1))
            return x[0].reshape(int(np.prod(x[0].shape[:ax])), -1)
        if t == "Transpose":  # src/onnxstream.cpp:5176-5236
            return np.transpose(x[0], [int(v) for v in a["perm"].split(",")])
        if t == "Concat":  # src/onnxstream.cpp:4140-4299
            return np.concatenate(x, axis=int(a["axis"]))
        if t == "Split":
            ax = int(a.get("axis", 0))
            sizes = [int(v) for v in x[1]] if len(x) > 1 and x[1] is not None else [x[0].shape[ax] // len(op["outputs"])] * len(op["outputs"])
            return np.split(x[0], np.cumsum(sizes)[:-1], axis=ax)
        if t == "Slice":  # src/onnxstream.cpp:6499-6695
            y = x[0]
            starts, ends = x[1], x[2]
            axes = x[3] if len(x) > 3 and x[3] is not None else np.arange(len(starts))
            for s, e, ax in zip(starts, ends, axes):
                idx = [slice(None)] * y.ndim
                idx[int(ax)] = slice(int(s), int(e))
                y = y[tuple(idx)]
            return y
        if t == "Resize":  # nearest / asymmetric / floor, src/onnxstream.cpp:6120-6315
            sc = x[2]
            sy, sx = float(sc[2]), float(sc[3])
            h, w = x[0].shape[2], x[0].shape[3]
            ho, wo = int(h * sy), int(w * sx)
            yi = np.minimum((np.arange(ho) / sy).astype(np.int64), h - 1)
            xi = np.minimum((np.arange(wo) / sx).astype(np.int64), w - 1)
            return x[0][:, :, yi][:, :, :, xi]
        if t == "Softmax":  # src/onnxstream.cpp:5862-5998
            ax = int(a.get("axis", -1))
            e = np.exp(x[0] - x[0].max(axis=ax, keepdims=True))
            return e / e.sum(axis=ax, keepdims=True)
        if t == "InstanceNormalization":  # src/onnxstream.cpp:4788-5055 (double statistics)
            eps = float(a.get("epsilon", 1e-5))
            v = x[0]
            mean = v.mean(axis=2, keepdims=True)
            var = ((v - mean) ** 2).mean(axis=2, keepdims=True)
            return x[1].reshape(1, -1, 1) * (v - mean) / np.sqrt(var + eps) + x[2].reshape(1, -1, 1)
        if t == "ReduceMean":  # src/onnxstream.cpp:5237-5393
            return x[0].mean(axis=int(a["axes"]), keepdims=bool(int(a.get("keepdims", "1"))))
        if t == "Gather":  # src/onnxstream.cpp:6316-6498
            return np.take(x[0], x[1].astype(np.int64), axis=int(a.get("axis", 0)))
        # ---- index / shape ops of the LLM graphs: int64 in, int64 out, bit-exact by construction ----
        if t == "Shape":  # src/onnxstream.cpp:7003-7033
            return np.asarray(x[0].shape, dtype=np.int64)
        if t == "Cast":  # src/onnxstream.cpp:7352-7424: ONNX type ids 1 = float, 10 = float16, 6/7/9 = int32/int64/bool (all int64 here)
            to = int(a["to"])
            if to in (1, 10):
                return x[0].astype(np.float64)
            if to in (6, 7, 9):
                return x[0].astype(np.float32).astype(np.int64) if x[0].dtype != np.int64 else x[0]
            raise NotImplementedError("Cast to %d" % to)
        if t == "ConstantOfShape":  # src/onnxstream.cpp:7543-7588: "value" without a decimal point -> int64 tensor
            v = a.get("value", "0")
            shape = tuple(int(d) for d in x[0])
            return np.full(shape, int(v), dtype=np.int64) if "." not in v else np.full(shape, float(v), dtype=np.float64)
        if t == "Range":  # src/onnxstream.cpp:7589-7636 (int64 only)
            return np.arange(int(x[0].reshape(-1)[0]), int(x[1].reshape(-1)[0]), int(x[2].reshape(-1)[0]), dtype=np.int64)
        if t in ("Less", "Greater", "Equal", "And"):  # src/onnxstream.cpp:7637-7766 (int64 operands, int64 0/1 result)
            p, q = x[0].astype(np.int64), x[1].astype(np.int64)
            r = {"Less": p < q, "Greater": p > q, "Equal": p == q, "And": (p != 0) & (q != 0)}[t]
            return r.astype(np.int64)
        if t == "Where":  # src/onnxstream.cpp:7034-7153
            r = np.where(x[0] != 0, x[1], x[2])
            return r.astype(np.int64) if x[1].dtype == np.int64 and x[2].dtype == np.int64 else r.astype(np.float64)
        if t == "Expand":  # src/onnxstream.cpp:7154-7351: numpy-style broadcast of the input against the target shape
            target = tuple(int(d) for d in x[1])
            return np.broadcast_to(x[0], np.broadcast_shapes(x[0].shape, target)).copy()
        if t == "Trilu":  # src/onnxstream.cpp:7883-7938: upper only, out[y][x] = x - k >= y ? in : 0
            assert a.get("upper", "1") == "1" and x[0].ndim == 2
            return np.triu(x[0], int(np.asarray(x[1]).reshape(-1)[0]))
        if t == "ScatterND":  # src/onnxstream.cpp:7939-8074: full-rank indices, element-wise scatter into a copy
            out = np.array(x[0], copy=True)
            idx = x[1].astype(np.int64).reshape(-1, x[0].ndim)
            out[tuple(idx[:, j] for j in range(x[0].ndim))] = x[2].reshape(-1)
            return out
        if t == "ArgMax":  # src/onnxstream.cpp:6930-7002: int64 (1, D), last axis, keepdims 0, first maximum
            assert x[0].dtype == np.int64 and x[0].ndim == 2 and x[0].shape[0] == 1 and int(a.get("keepdims", "1")) == 0
            return np.asarray([int(np.argmax(x[0][0]))], dtype=np.int64)
        if t == "MaxPool":  # src/onnxstream.cpp:8075-8143, 1537-1664: padding re-symmetrised, padded taps ignored
            kh, kw = (int(v) for v in a["kernel_shape"].split(","))
            pads = [int(v) for v in a["pads"].split(",")]
            st = int(a["strides"].split(",")[0])
            ph, pw = pads[0] + pads[2], pads[1] + pads[3]
            pt, pl = ph // 2, pw // 2
            _, c, h, wd = x[0].shape
            ho, wo = (h + ph - kh) // st + 1, (wd + pw - kw) // st + 1
            xp = np.full((c, h + ph, wd + pw), -np.inf)
            xp[:, pt:pt + h, pl:pl + wd] = x[0][0]
            y = np.full((c, ho, wo), -np.inf)
            for ky in range(kh):
                for kx in range(kw):
                    y = np.maximum(y, xp[:, ky:ky + st * (ho - 1) + 1:st, kx:kx + st * (wo - 1) + 1:st])
            return y[None]
        raise NotImplementedError(t)


# ---------------------------------------------------------------------------------------------------------------------
# uint8 path (m_use_uint8_qdq / m_use_uint8_arithmetic): restated from the reference and pinned BIT-EXACT against the reference run
# (tests/test_cpu.py::test_qu8_restatement_bit_exact) -- the XNNPACK kernels behind these calls are the real library in oracle/_ref.
# ---------------------------------------------------------------------------------------------------------------------
def qu8_percentiles(x, threads=4, from_left=0.001, from_right=0.001):
    """Model::get_percentiles (src/onnxstream.cpp:3104-3232) + FloatAsUInt::get_percentiles (2302-2386): the flat tensor is split over
    `threads` pool workers (get_start_and_end, 3091-3102), each span walked in 64 KiB chunks; per chunk the k-th smallest / largest
    finite value with k = (size_t)(n * 0.001f); min of the lows, max of the highs.  None when no chunk has a result."""
    flat = np.asarray(x).ravel()
    size = flat.size
    chunk = 16384 if flat.dtype == np.float32 else 32768
    n = size // threads or 1
    lo, hi, found = np.inf, -np.inf, False
    for i in range(threads):
        st, en = i * n, (size if i >= threads - 1 else (i + 1) * n)
        if st >= en or st >= size:
            continue
        for j in range(st, en, chunk):
            nn = min(en, j + chunk) - j
            c = np.sort(flat[j:j + nn].astype(np.float32))
            c = c[np.isfinite(c)]
            kl, kr = int(np.float32(nn) * np.float32(from_left)), int(np.float32(nn) * np.float32(from_right))
            if kl >= len(c) or kr >= len(c):
                continue
            lo, hi, found = min(lo, c[kl]), max(hi, c[len(c) - 1 - kr]), True
    return (np.float32(lo), np.float32(hi)) if found and lo < hi else None


def qu8_range_to_scale(lo, hi):
    """Model::range_to_scale (src/onnxstream.cpp:3234-3245)."""
    lo, hi = np.float32(lo), np.float32(hi)
    if lo > 0 and hi > 0:
        lo = np.float32(0)
    elif lo < 0 and hi < 0:
        hi = np.float32(0)
    scale = np.float32(np.float64(np.float32(hi - lo)) / 255.0)
    return scale, int(np.uint8(np.float32(abs(lo)) / scale))


def qu8_quantize(x, scale, zp):
    """xnn_run_convert_nc_f32_qu8: x * (1 / scale), clamp to [0 - zp, 255 - zp], round to nearest even, + zp."""
    q = np.asarray(x, np.float32) * np.float32(np.float32(1.0) / scale)
    q = np.clip(q, np.float32(0 - zp), np.float32(255 - zp))
    return (np.rint(q).astype(np.int32) + zp).astype(np.uint8)


def qu8_dequantize(q, scale, zp):
    return ((q.astype(np.int32) - zp).astype(np.float32) * np.float32(scale)).astype(np.float32)


def qu8_add(qa, sa, za, qb, sb, zb, so, zo):
    """XNNPACK qu8 vadd (fixed point): shift = 20 - exponent(max(|sa/so|, |sb/so|)); multipliers = lrintf(|s/so| * 2^shift)."""
    ao, bo = np.float32(sa / so), np.float32(sb / so)
    mx = np.float32(max(abs(ao), abs(bo)))
    shift = int(20 - ((int(mx.view(np.uint32)) >> 23) - 127))

    def mult(v):
        bits = np.uint32(int(np.float32(abs(v)).view(np.uint32)) + (shift << 23))
        m = int(np.rint(bits.view(np.float32)))
        return -m if v < 0 else m
    am, bm = mult(ao), mult(bo)
    acc = ((1 << (shift - 1)) - am * za - bm * zb) + qa.astype(np.int64) * am + qb.astype(np.int64) * bm
    return (np.clip(acc >> shift, 0 - zo, 255 - zo) + zo).astype(np.uint8)


def qu8_mul(qa, sa, za, qb, sb, zb, so, zo):
    """XNNPACK qu8 vmul (fp32 requantisation)."""
    scale = np.float32(np.float32(sa * sb) / so)
    f = ((qa.astype(np.int32) - za) * (qb.astype(np.int32) - zb)).astype(np.float32) * scale
    f = np.clip(f, np.float32(0 - zo), np.float32(255 - zo))
    return (np.rint(f).astype(np.int32) + zo).astype(np.uint8)


def qu8_gemm(qa, sa, za, qw, sw, zw, so, zo, bias_i32=None):
    """XNNPACK qu8 fully-connected / convolution arithmetic on an im2col'd problem: int32 accumulate, fp32 requantisation."""
    acc = (qa.astype(np.int64) - za) @ (qw.astype(np.int64) - zw)
    if bias_i32 is not None:
        acc = acc + bias_i32
    scale = np.float32(np.float32(np.float32(sa) * np.float32(sw)) / np.float32(so))
    f = acc.astype(np.int32).astype(np.float32) * scale
    f = np.clip(f, np.float32(0 - zo), np.float32(255 - zo))
    return (np.rint(f).astype(np.int32) + zo).astype(np.uint8)
