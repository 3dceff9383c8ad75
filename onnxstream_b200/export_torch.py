"""Torch-only exporter: `torch.nn.Module` -> OnnxStream model directory (`model.txt` + weight blobs), no `onnx` package needed.

SURVEY.md section 8 row f4.  The reference ships `onnx2txt/onnx2txt.ipynb`, which converts an ONNX file; this image has neither
`onnx` nor a torch ONNX exporter that works without it, so the module is traced with `torch.export` (ATen graph, concrete shapes)
and every ATen node is written as the ONNX-style op sequence the reference's `Model::run()` understands -- the same sequences a
diffusers / transformers ONNX export contains, so that the engine's fusion matchers (GroupNorm, LayerNorm, GELU/GEGLU, SiLU,
attention) see the patterns they were written for.  File format and naming rules follow the converter (cell 1):

* names: every character outside [A-Za-z0-9] becomes `_HEX_` (cell 1:65-74);
* Conv weights: listed as `X_nchw.bin(dtype:O,I,kh,kw)`, stored as `X_nhwc.bin` in OHWI (cell 1:142-149; src/onnxstream.cpp:2666-2692);
* `Linear` on a 2-D input with a bias is a `Gemm` whose B is stored pre-transposed `[K,N]` under a `_transposed` name (cell 1:129-141);
* constants (shapes, scalars) become blobs (cell 1:160-171); weight dtype float32 / float16 / uint8 percentile rule (cell 1:25-58).

Only inference graphs with static shapes and batch 1 image tensors are supported (what the reference runs).  Unsupported ATen ops
raise `NotImplementedError` naming the op.
"""
from __future__ import annotations

import math
import operator
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .emit import GraphBuilder, T


def mangle(name: str) -> str:
    """onnx2txt.ipynb cell 1:65-74."""
    return "".join(c if c.isalnum() and c.isascii() else "_%02X_" % ord(c) for c in name)


class TorchExporter:
    def __init__(self, out_dir: Optional[str], wdtype: str = "float32", keep_in_memory: bool = False):
        self.g = GraphBuilder(out_dir, wdtype, keep_in_memory=keep_in_memory)
        self.env: Dict[str, object] = {}     # fx node name -> T | list[T] | python scalar | np.ndarray (host constant)
        self.input_names: List[str] = []
        self.output_names: List[str] = []
        self.prov: Dict[str, tuple] = {}     # T.name -> how it was produced (used to canonicalise attention blocks)

    # ---------------------------------------------------------------------------------------------------------
    def export(self, module, example_args: Tuple, input_names: Optional[Sequence[str]] = None, output_names: Optional[Sequence[str]] = None) -> Dict:
        import torch
        module = module.eval()
        with torch.no_grad():
            ep = torch.export.export(module, tuple(example_args))
        sig = ep.graph_signature
        params = {}
        for spec in sig.input_specs:
            kind = spec.kind.name
            if kind in ("PARAMETER", "BUFFER"):
                params[spec.arg.name] = ep.state_dict[spec.target] if spec.target in ep.state_dict else ep.constants[spec.target]
            elif kind == "CONSTANT_TENSOR":
                params[spec.arg.name] = ep.constants[spec.target]
        user_inputs = [s.arg.name for s in sig.input_specs if s.kind.name == "USER_INPUT"]
        if input_names is not None:
            assert len(input_names) == len(user_inputs), "one name per positional tensor input"
        rename = dict(zip(user_inputs, input_names or user_inputs))
        nodes = list(ep.graph.nodes)
        # dead-code elimination: HF models trace pooling / bookkeeping branches whose results the wrapper does not return
        live, stack = set(), [n for n in nodes if n.op == "output"]
        while stack:
            n = stack.pop()
            if n in live:
                continue
            live.add(n)
            stack.extend(n.all_input_nodes)
        for n in nodes:
            if n not in live:
                continue
            if n.op == "placeholder":
                if n.name in params:
                    self.env[n.name] = params[n.name].detach().cpu()      # torch tensor: materialised as a blob where it is used
                else:
                    nm = mangle(rename[n.name])
                    self.input_names.append(nm)
                    self.env[n.name] = self.g.input(nm, tuple(n.meta["val"].shape))
            elif n.op == "call_function":
                if not self._fold(n):
                    self._node(n)
            elif n.op == "output":
                outs = n.args[0] if isinstance(n.args[0], (list, tuple)) else [n.args[0]]
                for i, o in enumerate(outs):
                    t = self._act(o)
                    nm = mangle(output_names[i]) if output_names else f"output{i}"
                    # an Identity-like rename keeps the producer line untouched: Mul by 1 would change numerics, so re-emit the
                    # last line of the producer with the requested output name instead
                    self._rename_output(t, nm)
                    self.output_names.append(nm)
        self.g.finish()
        return {"inputs": self.input_names, "outputs": self.output_names, "ops": len(self.g.lines), "weight_bytes": self.g.weight_bytes}

    def _fold(self, n) -> bool:
        """Constant folding: a node whose tensor inputs are all export-time constants (shape arithmetic, position ids, attention
        masks built from `arange`) is evaluated now with torch; its value becomes a blob wherever an emitted op consumes it."""
        import torch

        def const_only(v):
            if isinstance(v, torch.fx.Node):
                return const_only(self.env[v.name])
            if isinstance(v, T):
                return False
            if isinstance(v, (list, tuple)):
                return all(const_only(e) for e in v)
            return True

        def value(v):
            if isinstance(v, torch.fx.Node):
                return self.env[v.name]
            if isinstance(v, (list, tuple)):
                return type(v)(value(e) for e in v)
            return v

        if n.target is operator.getitem and not const_only(n.args[0]):
            return False
        if not (const_only(list(n.args)) and const_only(list(n.kwargs.values()))):
            return False
        kwargs = {k: value(v) for k, v in n.kwargs.items() if k != "device"}
        with torch.no_grad():
            self.env[n.name] = n.target(*value(list(n.args)), **kwargs)
        return True

    def _rename_output(self, t: T, new_name: str):
        old = t.text()
        new = T(new_name, t.shape).text()
        for i in range(len(self.g.lines) - 1, -1, -1):
            head, _, tail = self.g.lines[i].partition("*output:")
            if old in tail.split("*")[0].split(";"):
                self.g.lines[i] = self.g.lines[i].replace("*output:" + tail.split("*")[0], "*output:" + tail.split("*")[0].replace(old, new), 1)
                # later consumers (none for a graph output normally) keep working
                for j in range(i + 1, len(self.g.lines)):
                    self.g.lines[j] = self.g.lines[j].replace(old, new)
                return
        raise RuntimeError("graph output %s is not produced by any emitted op (a bare input or constant?)" % t.name)

    # ---------------------------------------------------------------------------------------------------------
    # value access
    def _val(self, a):
        import torch
        if isinstance(a, torch.fx.Node):
            return self.env[a.name]
        return a

    def _act(self, a) -> T:
        """An activation tensor reference; torch constants (parameters / buffers) are written as blobs."""
        import torch
        v = self._val(a)
        if isinstance(v, T):
            return v
        if isinstance(v, torch.Tensor):
            return self._const(v)
        raise TypeError("expected a tensor, got %r" % (type(v),))

    def _const(self, t, conv_weight: bool = False, quantizable: bool = True, name: Optional[str] = None) -> T:
        import torch
        if isinstance(t, torch.Tensor):
            arr = t.detach().cpu()
            arr = arr.to(torch.int64).numpy() if arr.dtype in (torch.int64, torch.int32, torch.bool) else arr.to(torch.float32).numpy()
        else:
            arr = np.asarray(t)
        return self.g.const(arr, name=name, conv_weight=conv_weight, quantizable=quantizable)

    def _canonical_heads(self, q: T, k: T, v: T):
        """If q, k, v are each Transpose(0,2,1,3)(Reshape[1,T,H,D](MatMul(x, W))) and those nine lines are exactly the tail of the
        emitted graph, replace them by the canonical head-split sequence and return (q [H,T,D], k^T [H,D,Tk], v [H,Tk,D])."""
        g = self.g
        pieces, lines = [], set()
        for t in (q, k, v):
            p1 = self.prov.get(t.name)
            if not p1 or p1[0] != "transpose" or p1[2] != (0, 2, 1, 3):
                return None
            p2 = self.prov.get(p1[1].name)
            if not p2 or p2[0] != "reshape" or len(p1[1].shape) != 4 or p1[1].shape[0] != 1:
                return None
            p3 = self.prov.get(p2[1].name)
            if not p3 or p3[0] != "linear" or len(p2[1].shape) != 3 or p2[1].shape[0] != 1:
                return None
            pieces.append((p3[1], p3[2], p1[1].shape))     # (x, W, [1, T, H, D])
            lines |= {p1[3], p2[2], p3[3]}
        n = len(g.lines)
        if lines != set(range(n - 9, n)):
            return None
        del g.lines[n - 9:]
        outs = []
        for idx, (x, w, (_, t, h, d)) in enumerate(pieces):
            y = g.node("MatMul", [x, w], [(1, t, h * d)])
            r = g.node("Reshape", [y, g.i64([1, t, h, d])], [(1, t, h, d)])
            p = g.node("Transpose", [r], [(1, h, t, d)], [("perm", "0,2,1,3")])
            r2 = g.node("Reshape", [p, g.i64([h, t, d])], [(h, t, d)])
            if idx == 1:
                r2 = g.node("Transpose", [r2], [(h, d, t)], [("perm", "0,2,1")])
            outs.append(r2)
        return tuple(outs)

    @staticmethod
    def _shape(n) -> Tuple[int, ...]:
        return tuple(int(d) for d in n.meta["val"].shape)

    def _reshape(self, x: T, shape: Sequence[int]) -> T:
        shape = tuple(int(d) for d in shape)
        if tuple(x.shape) == shape:
            return x
        return self.g.node("Reshape", [x, self.g.i64(list(shape))], [shape])

    # ---------------------------------------------------------------------------------------------------------
    def _node(self, n):
        import torch
        tgt = n.target
        name = tgt.__name__ if tgt is operator.getitem else str(tgt)
        a = n.args
        g = self.g
        out_shape = self._shape(n) if hasattr(n.meta.get("val", None), "shape") else None

        def binary(op):
            x, y = self._val(a[0]), self._val(a[1])
            if "alpha" in n.kwargs and n.kwargs["alpha"] != 1:
                raise NotImplementedError("alpha != 1 in " + name)
            xt = x if isinstance(x, T) else (self._const(x) if isinstance(x, torch.Tensor) else g.scalar(float(x)))
            yt = y if isinstance(y, T) else (self._const(y) if isinstance(y, torch.Tensor) else g.scalar(float(y)))
            return g.node(op, [xt, yt], [out_shape])

        if tgt is operator.getitem:
            self.env[n.name] = self._val(a[0])[a[1]]
        elif name in ("aten.conv2d.default", "aten.convolution.default"):
            x, w = self._act(a[0]), self._val(a[1])
            b = self._val(a[2]) if len(a) > 2 else None
            stride = list(a[3]) if len(a) > 3 else [1, 1]
            pad = list(a[4]) if len(a) > 4 else [0, 0]
            dil = list(a[5]) if len(a) > 5 else [1, 1]
            if name == "aten.convolution.default":
                if a[6]:
                    raise NotImplementedError("transposed convolution")
                groups = a[8]
            else:
                groups = a[6] if len(a) > 6 else 1
            if isinstance(pad, str) or groups != 1 or list(dil) != [1, 1] or stride[0] != stride[1] or w.dim() != 4:
                raise NotImplementedError("Conv: only 2-D, groups 1, dilation 1, square stride (src/onnxstream.cpp:4502-4560)")
            ins = [x, self._const(w, conv_weight=True)]
            if b is not None:
                ins.append(self._const(b, quantizable=False))
            kh, kw = int(w.shape[2]), int(w.shape[3])
            self.env[n.name] = g.node("Conv", ins, [out_shape], [("dilations", "1,1"), ("group", "1"), ("kernel_shape", f"{kh},{kw}"),
                                                                 ("pads", f"{pad[0]},{pad[1]},{pad[0]},{pad[1]}"), ("strides", f"{stride[0]},{stride[1]}")])
        elif name == "aten.linear.default":
            x, w = self._act(a[0]), self._val(a[1])
            b = self._val(a[2]) if len(a) > 2 and a[2] is not None else None
            wt = w.detach().t().contiguous()                       # [K, N], what MatMul / the folded Gemm read
            g.flops += 2 * int(np.prod(x.shape[:-1])) * int(wt.shape[0]) * int(wt.shape[1])
            if len(x.shape) == 2 and b is not None:
                wn = self._const(wt, name=g._uid("w") + "_transposed")
                self.env[n.name] = g.node("Gemm", [x, wn, self._const(b)], [out_shape])
            else:
                wT = self._const(wt)
                y = g.node("MatMul", [x, wT], [out_shape])
                if b is None:
                    self.prov[y.name] = ("linear", x, wT, len(g.lines) - 1)
                else:
                    y = g.node("Add", [self._const(b), y], [out_shape])
                self.env[n.name] = y
        elif name == "aten.group_norm.default":
            x = self._act(a[0])
            groups = int(a[1])
            w = self._val(a[2]) if len(a) > 2 else None
            b = self._val(a[3]) if len(a) > 3 else None
            eps = float(a[4]) if len(a) > 4 else 1e-5
            c = x.shape[1]
            r = g.node("Reshape", [x, g.i64([0, groups, -1])], [(1, groups, int(np.prod(x.shape)) // groups)])
            i = g.node("InstanceNormalization", [r, g.const(np.ones(groups, np.float32), quantizable=False), g.const(np.zeros(groups, np.float32), quantizable=False)],
                       [r.shape], [("epsilon", repr(eps))])
            y = g.node("Reshape", [i, g.i64(list(x.shape))], [x.shape])
            if w is not None:
                y = g.node("Mul", [y, self._const(w.reshape(c, *([1] * (len(x.shape) - 2))))], [x.shape])
            if b is not None:
                y = g.node("Add", [y, self._const(b.reshape(c, *([1] * (len(x.shape) - 2))))], [x.shape])
            self.env[n.name] = y
        elif name == "aten.layer_norm.default":
            x = self._act(a[0])
            if len(a[1]) != 1 or int(a[1][0]) != x.shape[-1]:
                raise NotImplementedError("LayerNorm over more than the last axis")
            w = self._val(a[2]) if len(a) > 2 else None
            b = self._val(a[3]) if len(a) > 3 else None
            eps = float(a[4]) if len(a) > 4 else 1e-5
            red = tuple(x.shape[:-1]) + (1,)
            mean = g.node("ReduceMean", [x], [red], [("axes", "-1"), ("keepdims", "1")])
            d = g.node("Sub", [x, mean], [x.shape])
            p = g.node("Pow", [d, g.scalar(2.0)], [x.shape])
            var = g.node("ReduceMean", [p], [red], [("axes", "-1"), ("keepdims", "1")])
            ve = g.node("Add", [var, g.scalar(eps)], [red])
            sd = g.node("Sqrt", [ve], [red])
            y = g.node("Div", [d, sd], [x.shape])
            if w is not None:
                y = g.node("Mul", [y, self._const(w)], [x.shape])
            if b is not None:
                y = g.node("Add", [y, self._const(b)], [x.shape])
            self.env[n.name] = y
        elif name == "aten.silu.default":
            x = self._act(a[0])
            self.env[n.name] = g.node("Mul", [x, g.node("Sigmoid", [x], [x.shape])], [x.shape])
        elif name == "aten.gelu.default":
            if n.kwargs.get("approximate", "none") != "none":
                raise NotImplementedError("tanh-approximated GELU (the reference has no Tanh)")
            x = self._act(a[0])
            d = g.node("Div", [x, g.scalar(math.sqrt(2.0))], [x.shape])
            e = g.node("Erf", [d], [x.shape])
            s1 = g.node("Add", [e, g.scalar(1.0)], [x.shape])
            m = g.node("Mul", [x, s1], [x.shape])
            self.env[n.name] = g.node("Mul", [m, g.scalar(0.5)], [x.shape])
        elif name in ("aten.sigmoid.default", "aten.erf.default", "aten.sqrt.default", "aten.sin.default", "aten.cos.default", "aten.neg.default"):
            op = {"sigmoid": "Sigmoid", "erf": "Erf", "sqrt": "Sqrt", "sin": "Sin", "cos": "Cos", "neg": "Neg"}[name.split(".")[1]]
            x = self._act(a[0])
            self.env[n.name] = g.node(op, [x], [x.shape])
        elif name in ("aten.add.Tensor", "aten.add.Scalar"):
            self.env[n.name] = binary("Add")
        elif name in ("aten.sub.Tensor", "aten.sub.Scalar"):
            self.env[n.name] = binary("Sub")
        elif name in ("aten.mul.Tensor", "aten.mul.Scalar"):
            self.env[n.name] = binary("Mul")
        elif name in ("aten.div.Tensor", "aten.div.Scalar"):
            self.env[n.name] = binary("Div")
        elif name == "aten.pow.Tensor_Scalar":
            x = self._act(a[0])
            self.env[n.name] = g.node("Pow", [x, g.scalar(float(a[1]))], [x.shape])
        elif name == "aten.rsqrt.default":
            x = self._act(a[0])
            self.env[n.name] = g.node("Div", [g.scalar(1.0), g.node("Sqrt", [x], [x.shape])], [x.shape])
        elif name == "aten.mean.dim":
            x = self._act(a[0])
            dims = [d % len(x.shape) for d in a[1]]
            keep = bool(a[2]) if len(a) > 2 else False
            if dims != [len(x.shape) - 1]:
                raise NotImplementedError("ReduceMean over an axis other than the last (src/onnxstream.cpp:5262-5268)")
            red = tuple(x.shape[:-1]) + (1,)
            y = g.node("ReduceMean", [x], [red], [("axes", "-1"), ("keepdims", "1")])
            self.env[n.name] = y if keep else self._reshape(y, out_shape)
        elif name in ("aten.softmax.int", "aten._softmax.default"):
            x = self._act(a[0])
            self.env[n.name] = g.node("Softmax", [x], [x.shape], [("axis", str(int(a[1])))])
        elif name in ("aten.matmul.default", "aten.bmm.default", "aten.mm.default"):
            x, y = self._act(a[0]), self._act(a[1])
            g.flops += 2 * int(np.prod(out_shape)) * int(x.shape[-1])
            self.env[n.name] = g.node("MatMul", [x, y], [out_shape])
        elif name == "aten.scaled_dot_product_attention.default":
            q, k, v = self._act(a[0]), self._act(a[1]), self._act(a[2])
            mask = self._val(a[3]) if len(a) > 3 else self._val(n.kwargs.get("attn_mask"))
            causal = bool(a[5]) if len(a) > 5 else bool(n.kwargs.get("is_causal", False))
            scale = n.kwargs.get("scale") or 1.0 / math.sqrt(q.shape[-1])
            if len(q.shape) != 4 or q.shape[0] != 1 or q.shape[1] != k.shape[1]:
                raise NotImplementedError("scaled_dot_product_attention: expected [1, H, T, D] with equal head counts")
            if causal:
                if mask is not None:
                    raise NotImplementedError("is_causal together with attn_mask")
                mask = torch.ones(q.shape[2], k.shape[2], dtype=torch.bool).tril()
            if isinstance(mask, torch.Tensor):
                # additive form; -1e4 is finite in fp16 and exp(-1e4) == 0 in every precision the engines use
                if mask.dtype == torch.bool:
                    mask = torch.zeros(mask.shape, dtype=torch.float32).masked_fill(~mask, -1.0e4)
                mask = mask.to(torch.float32).clamp(min=-1.0e4)
                while mask.dim() > 2 and mask.shape[0] == 1:
                    mask = mask[0]
                if mask.dim() == 3 and mask.shape[0] not in (1, q.shape[1]):
                    raise NotImplementedError("attention mask batch shape")
            # the diffusers-export form of attention on [H, T, D]: MatMul(q, k^T) -> Mul(scale) -> Softmax -> MatMul(p, v)
            h, tq, d = q.shape[1:]
            tk = k.shape[2]
            canon = self._canonical_heads(q, k, v) if mask is None else None
            if canon is not None:
                # q/k/v = transpose(view(linear(x))) and nothing else was emitted in between: re-emit the three projections in the order
                # of the diffusers ONNX export (projection, Reshape, Transpose, Reshape per operand; K transposed last), the 20-op
                # window the engine turns into one grouped projection launch + one flash-attention launch
                q3, kt, v3 = canon
            else:
                q3, k3, v3 = self._reshape(q, (h, tq, d)), self._reshape(k, (h, tk, d)), self._reshape(v, (h, tk, v.shape[3]))
                kt = g.node("Transpose", [k3], [(h, d, tk)], [("perm", "0,2,1")])
            s = g.node("MatMul", [q3, kt], [(h, tq, tk)])
            s = g.node("Mul", [s, g.scalar(float(scale))], [(h, tq, tk)])
            if mask is not None:
                mt = mask if isinstance(mask, T) else self._const(mask, quantizable=False)
                s = g.node("Add", [s, mt], [(h, tq, tk)])
            p = g.node("Softmax", [s], [(h, tq, tk)], [("axis", "-1")])
            o = g.node("MatMul", [p, v3], [(h, tq, v.shape[3])])
            g.flops += 2 * h * tq * tk * (d + v.shape[3])
            self.env[n.name] = self._reshape(o, out_shape)
        elif name in ("aten.view.default", "aten.reshape.default", "aten._unsafe_view.default", "aten.flatten.using_ints", "aten.unflatten.int",
                      "aten.unsqueeze.default", "aten.squeeze.dim", "aten.squeeze.default", "aten.squeeze.dims"):
            src = self._act(a[0])
            y = self._reshape(src, out_shape)
            if y is not src:
                self.prov[y.name] = ("reshape", src, len(g.lines) - 1)
            self.env[n.name] = y
        elif name in ("aten.permute.default", "aten.transpose.int", "aten.t.default"):
            x = self._act(a[0])
            r = len(x.shape)
            if name == "aten.permute.default":
                perm = [int(d) % r for d in a[1]]
            elif name == "aten.t.default":
                perm = [1, 0] if r == 2 else list(range(r))
            else:
                perm = list(range(r))
                d0, d1 = int(a[1]) % r, int(a[2]) % r
                perm[d0], perm[d1] = perm[d1], perm[d0]
            if perm == list(range(r)):
                self.env[n.name] = x
            else:
                y = g.node("Transpose", [x], [out_shape], [("perm", ",".join(map(str, perm)))])
                self.prov[y.name] = ("transpose", x, tuple(perm), len(g.lines) - 1)
                self.env[n.name] = y
        elif name == "aten.cat.default":
            parts = [self._act(p) for p in a[0]]
            axis = int(a[1]) if len(a) > 1 else 0
            self.env[n.name] = g.node("Concat", parts, [out_shape], [("axis", str(axis))])
        elif name in ("aten.chunk.default", "aten.split.Tensor", "aten.split_with_sizes.default"):
            x = self._act(a[0])
            dim = int(a[2] if len(a) > 2 else n.kwargs.get("dim", 0)) % len(x.shape)
            size = x.shape[dim]
            if name == "aten.chunk.default":
                per = -(-size // int(a[1]))
                sizes = [min(per, size - i) for i in range(0, size, per)]
            elif name == "aten.split.Tensor":
                per = int(a[1])
                sizes = [min(per, size - i) for i in range(0, size, per)]
            else:
                sizes = [int(s) for s in a[1]]
            outs, start = [], 0
            for s in sizes:
                shp = list(x.shape)
                shp[dim] = s
                outs.append(g.node("Slice", [x, g.i64([start]), g.i64([start + s]), g.i64([dim - len(x.shape) if dim == len(x.shape) - 1 else dim]), g.i64([1])], [tuple(shp)]))
                start += s
            self.env[n.name] = outs
        elif name == "aten.slice.Tensor":
            x = self._act(a[0])
            dim = int(a[1]) % len(x.shape) if len(a) > 1 else 0
            start = int(a[2]) if len(a) > 2 and a[2] is not None else 0
            end = int(a[3]) if len(a) > 3 and a[3] is not None else x.shape[dim]
            step = int(a[4]) if len(a) > 4 else 1
            end = min(end, x.shape[dim])
            if start < 0:
                start += x.shape[dim]
            if end < 0:
                end += x.shape[dim]
            if start == 0 and end == x.shape[dim] and step == 1:
                self.env[n.name] = x
            else:
                if step != 1:
                    raise NotImplementedError("Slice with step != 1 (src/onnxstream.cpp:6592)")
                self.env[n.name] = g.node("Slice", [x, g.i64([start]), g.i64([end]), g.i64([dim]), g.i64([1])], [out_shape])
        elif name == "aten.select.int":
            x = self._act(a[0])
            dim = int(a[1]) % len(x.shape)
            idx = int(a[2]) % x.shape[dim]
            shp = list(x.shape)
            shp[dim] = 1
            s = g.node("Slice", [x, g.i64([idx]), g.i64([idx + 1]), g.i64([dim]), g.i64([1])], [tuple(shp)])
            self.env[n.name] = self._reshape(s, out_shape)
        elif name in ("aten.upsample_nearest2d.vec", "aten.upsample_nearest2d.default"):
            x = self._act(a[0])
            sy, sx = out_shape[2] / x.shape[2], out_shape[3] / x.shape[3]
            if sy != int(sy) or sx != int(sx):
                raise NotImplementedError("non-integer nearest upsampling")
            scales = g.const(np.asarray([1, 1, sy, sx], np.float32), quantizable=False)
            self.env[n.name] = g.node("Resize", [x, None, scales], [out_shape], [("coordinate_transformation_mode", "asymmetric"), ("cubic_coeff_a", "-0.75"),
                                                                                 ("mode", "nearest"), ("nearest_mode", "floor")])
        elif name == "aten.embedding.default":
            w, idx = self._val(a[0]), self._act(a[1])
            self.env[n.name] = g.node("Gather", [self._const(w, quantizable=False), idx], [out_shape], [("axis", "0")])
        elif name == "aten.expand.default":
            x = self._act(a[0])
            self.env[n.name] = x if tuple(x.shape) == tuple(out_shape) else g.node("Expand", [x, g.i64(list(out_shape))], [out_shape])
        elif name in ("aten.to.dtype", "aten.to.dtype_layout", "aten._to_copy.default", "aten.contiguous.default", "aten.clone.default", "aten.alias.default",
                      "aten.detach.default", "aten.dropout.default", "aten.lift_fresh_copy.default", "aten.type_as.default"):
            self.env[n.name] = self._val(a[0])
        elif name == "aten.max_pool2d.default":
            x = self._act(a[0])
            k = list(a[1])
            st = list(a[2]) if len(a) > 2 and a[2] else k
            pd = list(a[3]) if len(a) > 3 else [0, 0]
            if (len(a) > 4 and list(a[4]) != [1, 1]) or (len(a) > 5 and a[5]):
                raise NotImplementedError("MaxPool with dilation / ceil_mode (src/onnxstream.cpp:8097-8100)")
            self.env[n.name] = g.node("MaxPool", [x], [out_shape], [("ceil_mode", "0"), ("dilations", "1,1"), ("kernel_shape", f"{k[0]},{k[1]}"),
                                                                    ("pads", f"{pd[0]},{pd[1]},{pd[0]},{pd[1]}"), ("strides", f"{st[0]},{st[1]}")])
        else:
            raise NotImplementedError("ATen op not supported by the OnnxStream exporter: " + name)


def export_module(module, example_args: Tuple, out_dir: str, wdtype: str = "float32", input_names: Optional[Sequence[str]] = None,
                  output_names: Optional[Sequence[str]] = None) -> Dict:
    """Write `out_dir/model.txt` + blobs for `module(*example_args)`; returns {"inputs", "outputs", "ops", "weight_bytes"}."""
    if not out_dir.endswith("/"):
        out_dir += "/"
    return TorchExporter(out_dir, wdtype).export(module, example_args, input_names, output_names)
