"""TEST INFRASTRUCTURE ONLY — run the reference's OWN entropy-model code on CPU.

Works only where ``/root/reference`` exists (the build container, never the GPU box).
Two mechanisms:

1. ``load_stanh_modules()`` imports ``/root/reference/src/{entropy_models,quantization}``
   UNMODIFIED under stub ``compressai`` / ``torchac`` modules (those third-party wheels
   are not installed; only ``LowerBound``, ``pmf_to_quantized_cdf`` and the coder
   registry are touched at import/forward time and they are restated here).
2. ``load_tcm_functions()`` extracts ``get_scale_table``, ``ste_round`` and the
   ``TCM._likelihood`` / ``TCM._standardized_cumulative`` methods (the reference's private
   copy of the CompressAI Gaussian likelihood, ``tcm.py:33-37,570-588``) from the source
   file with ``ast`` and executes exactly that source — ``tcm.py`` itself cannot be
   imported (needs compressai, timm).

Used by ``oracle/gen_golden.py`` to produce ``tests/golden`` and by the container-only
test ``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

import ast
import math
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("RESLIC_REFERENCE_ROOT", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "entropy_models"))


class _LowerBoundFn(torch.autograd.Function):
    """compressai.ops.LowerBound autograd rule."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        pass_through = (x >= bound) | (grad_output < 0)
        return pass_through * grad_output, None


class LowerBound(torch.nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)


def _install_stubs():
    from . import compressai_ref

    if "compressai" in sys.modules and getattr(sys.modules["compressai"], "_reslic_stub", False):
        return
    comp = types.ModuleType("compressai")
    comp._reslic_stub = True
    comp.available_entropy_coders = lambda: ["ans"]
    comp.get_entropy_coder = lambda: "ans"
    cxx = types.ModuleType("compressai._CXX")
    cxx.pmf_to_quantized_cdf = lambda pmf, precision: compressai_ref.pmf_to_quantized_cdf(pmf, precision)
    ops = types.ModuleType("compressai.ops")
    ops.LowerBound = LowerBound
    ans = types.ModuleType("compressai.ans")

    class _NoCoder:
        def __getattr__(self, name):
            raise RuntimeError("rANS coder is not part of the oracle")

    ans.RansEncoder = _NoCoder
    ans.RansDecoder = _NoCoder
    ans.BufferedRansEncoder = _NoCoder
    comp._CXX, comp.ops, comp.ans = cxx, ops, ans
    sys.modules.update({"compressai": comp, "compressai._CXX": cxx, "compressai.ops": ops,
                        "compressai.ans": ans, "torchac": types.ModuleType("torchac")})


def load_stanh_modules():
    """Returns the reference's ``entropy_models`` and ``quantization`` packages.

    The only patch: ``NonSymStanH/SymStanH.update_state`` and
    ``NonSymStanH.update_cumulative_weights`` default to ``device=cuda``
    (``activation.py:72,91,218``), which cannot run on a CUDA-less host; the defaults are
    rebound to CPU.  No function body is changed.
    """
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import quantization.activation as act

    if not getattr(act, "_reslic_cpu_patched", False):
        cpu = torch.device("cpu")
        act.NonSymStanH.update_state.__defaults__ = (cpu,)
        act.NonSymStanH.update_cumulative_weights.__defaults__ = (cpu,)
        act.SymStanH.update_state.__defaults__ = (cpu,)
        act._reslic_cpu_patched = True
    import quantization

    quantization.SymStanH = act.SymStanH
    quantization.NonSymStanH = act.NonSymStanH
    import entropy_models

    return entropy_models, quantization


def load_tcm_functions():
    """exec() the reference source of get_scale_table / ste_round / TCM._likelihood /
    TCM._standardized_cumulative (tcm.py:26-37, 570-588) and return them in a namespace."""
    path = os.path.join(REFERENCE_SRC, "models", "reference", "tcm.py")
    with open(path) as f:
        src = f.read()
    tree = ast.parse(src)
    wanted_funcs = {"get_scale_table", "ste_round"}
    wanted_methods = {"_likelihood", "_standardized_cumulative"}
    body = []
    for node in tree.body:
        if isinstance(node, ast.Assign) and all(
            isinstance(t, ast.Name) and t.id.startswith("SCALES_") for t in node.targets
        ):
            body.append(node)
        elif isinstance(node, ast.FunctionDef) and node.name in wanted_funcs:
            body.append(node)
        elif isinstance(node, ast.ClassDef) and node.name == "TCM":
            methods = [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name in wanted_methods]
            cls = ast.ClassDef(name="TCMLikelihood", bases=[], keywords=[], body=methods, decorator_list=[])
            if hasattr(cls, "type_params"):
                cls.type_params = []
            body.append(cls)
    mod = ast.Module(body=body, type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "math": math, "Tensor": torch.Tensor}
    exec(compile(mod, path, "exec"), ns)
    return types.SimpleNamespace(get_scale_table=ns["get_scale_table"], ste_round=ns["ste_round"],
                                 tcm=ns["TCMLikelihood"]())
