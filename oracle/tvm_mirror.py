"""ORACLE support (test infrastructure): an INDEPENDENT restatement of three diffusers building blocks that this
image happens to ship -- TVM's `relax.frontend.nn` mirrors of diffusers' `Timesteps` / `get_timestep_embedding`,
`TimestepEmbedding` and `Attention` (site-packages/tilelang/3rdparty/tvm/python/tvm/relax/frontend/nn/op.py:1741-1804,
modules.py:748-960; written by the TVM project to load Hugging Face diffusers checkpoints, parameter names included).

TVM here has no LLVM backend, so the modules cannot be compiled; they are exported to Relax IR (`export_tvm`) and the
dataflow block is evaluated op by op with numpy (fp64).  The arithmetic DEFINITION therefore comes from TVM's code,
only the evaluation of primitive ops (matmul, exp, softmax ...) from numpy.  tests/test_oracle_dit.py uses this to pin
oracle/dit.py's `timestep_sinusoid`, `_TimestepEmbedder` and `Attention` against a second source besides SURVEY.md
Appendix A.  Skipped when tilelang/tvm is not importable.
"""
from __future__ import annotations

import numpy as np


def available() -> bool:
    try:
        import tilelang  # noqa: F401  (puts its bundled tvm on sys.path)
        import tvm  # noqa: F401
        from tvm.relax.frontend import nn  # noqa: F401
        return True
    except Exception:
        return False


def _shape_of(expr):
    return [int(v) for v in expr.values]


def _const(expr):
    return expr.data.numpy().astype(np.float64)


def _eval_call(call, env):
    import tvm
    from tvm import relax
    op = call.op.name

    def val(a):
        if isinstance(a, relax.Constant):
            return _const(a)
        if isinstance(a, relax.ShapeExpr):
            return _shape_of(a)
        if isinstance(a, relax.PrimValue):
            return int(a.value)
        if isinstance(a, relax.Tuple):
            return [val(f) for f in a.fields]
        return env[a.name_hint] if not isinstance(a, relax.Var) else env[a]

    args = [val(a) for a in call.args]
    at = call.attrs
    if op == "relax.astype":
        return args[0]
    if op == "relax.expand_dims":
        out = args[0]
        for ax in sorted(int(a) for a in at.axis):
            out = np.expand_dims(out, ax)
        return out
    if op == "relax.arange":
        return np.arange(args[0], args[1], args[2], dtype=np.float64)
    if op == "relax.multiply":
        return args[0] * args[1]
    if op == "relax.divide":
        return args[0] / args[1]
    if op == "relax.add":
        return args[0] + args[1]
    if op == "relax.exp":
        return np.exp(args[0])
    if op == "relax.cos":
        return np.cos(args[0])
    if op == "relax.sin":
        return np.sin(args[0])
    if op == "relax.concat":
        return np.concatenate(args[0], axis=int(at.axis))
    if op == "relax.permute_dims":
        axes = None if at.axes is None else [int(a) for a in at.axes]
        return np.transpose(args[0], axes)
    if op == "relax.matmul":
        return np.matmul(args[0], args[1])
    if op == "relax.reshape":
        return np.reshape(args[0], args[1])
    if op == "relax.nn.silu":
        return args[0] / (1.0 + np.exp(-args[0]))
    if op == "relax.nn.attention":
        # relax.nn.attention: q [B, Sq, H, D], k/v [B, Sk, H, D] -> softmax(q k^T * scale) v as [B, Sq, H, D];
        # scale defaults to 1/sqrt(D) (tvm/relax/op/nn/nn.py "attention")
        q, k, v = args[:3]
        if len(args) > 3:
            raise NotImplementedError("attention bias")
        if at.causal_mask is not None:
            raise NotImplementedError("causal attention")
        scale = float(at.scale.value) if at.scale is not None else q.shape[-1] ** -0.5
        s = np.einsum("bqhd,bkhd->bhqk", q, k) * scale
        s = s - s.max(axis=-1, keepdims=True)
        p = np.exp(s)
        p = p / p.sum(axis=-1, keepdims=True)
        return np.einsum("bhqk,bkhd->bqhd", p, v)
    raise NotImplementedError(f"relax op {op}")


def run_module(module, spec_inputs: dict, inputs: dict, params: dict) -> np.ndarray:
    """Export `module.forward` to Relax and evaluate it with numpy.  spec_inputs: name -> shape; inputs: name -> array;
    params: the module's parameter names (dots) -> arrays."""
    from tvm import relax
    from tvm.relax.frontend.nn import spec
    mod, named = module.export_tvm(spec={"forward": {k: spec.Tensor(list(s), "float32") for k, s in spec_inputs.items()}})
    fn = mod["forward"]
    env = {}
    pnames = [n for n, _ in named]
    for var in fn.params:
        name = var.name_hint
        if name in inputs:
            env[var] = np.asarray(inputs[name], dtype=np.float64)
        else:
            match = [p for p in pnames if p == name or p.replace(".", "_") == name]
            assert len(match) == 1, (name, pnames)
            env[var] = np.asarray(params[match[0]], dtype=np.float64)
    body = fn.body
    for block in body.blocks:
        for b in block.bindings:
            v = b.value
            if isinstance(v, relax.Call):
                env[b.var] = _eval_call(v, env)
            elif isinstance(v, relax.Var):
                env[b.var] = env[v]
            else:
                raise NotImplementedError(type(v).__name__)
    return env[body.body]


def timesteps(t, dim=256, flip_sin_to_cos=True, downscale_freq_shift=0):
    from tvm.relax.frontend import nn
    m = nn.modules.Timesteps(dim, flip_sin_to_cos=flip_sin_to_cos, downscale_freq_shift=downscale_freq_shift)
    t = np.asarray(t, dtype=np.float64)
    return run_module(m, {"x": t.shape}, {"x": t}, {})


def timestep_embedding(sample, params):
    from tvm.relax.frontend import nn
    w1 = params["linear_1.weight"]
    m = nn.modules.TimestepEmbedding(w1.shape[1], w1.shape[0])
    return run_module(m, {"sample": sample.shape}, {"sample": sample}, params)


def attention(x, ctx, params, heads):
    from tvm.relax.frontend import nn
    d = x.shape[-1]
    m = nn.modules.Attention(d, cross_attention_dim=ctx.shape[-1], heads=heads, dim_head=d // heads, bias=True)
    return run_module(m, {"hidden_states": x.shape, "encoder_hidden_states": ctx.shape},
                      {"hidden_states": x, "encoder_hidden_states": ctx}, params)
