"""Tracing Python energies into CUDA (tsu_emulator_b200/trace.py): expression graph, analytic gradient, generated source.
CPU only: the gradient graph is evaluated on the host against central differences, and the generated kernel source is
compiled with NVRTC for sm_100a (no GPU needed to compile)."""
import os
import sys

import numpy as np
import pytest

from tsu_emulator_b200.trace import TraceError, trace_energy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num_grad(f, x, eps=1e-6):
    g = np.zeros_like(x)
    for i in range(len(x)):
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        g[i] = (f(xp) - f(xm)) / (2 * eps)
    return g


CENTERS = [np.array([-2.0, 0.0, 1.0]), np.array([2.0, 1.0, 0.0])]
CASES = {
    "quartic": lambda x: np.sum(x ** 4) - 2 * np.sum(x ** 2) + 0.3 * x[0] * x[1],
    "mixture": lambda x: -np.log(sum(w * np.exp(-0.5 * np.sum((x - c) ** 2)) for c, w in zip(CENTERS, (0.3, 0.7))) + 1e-10),
    "rosenbrock": lambda x: 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2 + np.cos(x[2]) * np.tanh(x[1]) + np.sqrt(1 + x[2] ** 2),
    "matmul_abs": (lambda A: (lambda x: 0.5 * x @ A @ x + np.sum(np.abs(x))))(np.array([[2, .5, 0], [.5, 1, .2], [0, .2, 3.]])),
    "mean_pow_div": lambda x: np.mean(np.power(x, 2)) + np.dot(x, x) / 3 + np.exp(-x[0]) / (1 + np.square(x[1])),
    "general_pow": lambda x: (1.5 + x[0] ** 2) ** (0.5 * x[1]) + np.sin(x[2]) ** 3,
    "norm": lambda x: np.linalg.norm(x) + 2.0 ** x[0],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_traced_gradient_matches_central_differences(name):
    f = CASES[name]
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(5, 3))
    tr, out, grads = trace_energy(f, 3, pts)
    for p in pts:
        assert tr.evaluate([out], p)[0] == pytest.approx(float(f(p)), rel=1e-12, abs=1e-12)
        assert np.allclose(tr.evaluate(grads, p), num_grad(f, p), rtol=2e-5, atol=2e-5)
    src = tr.cuda_source(grads)
    assert "tsu_user_grad" in src and src.count("g[") == 3


def test_untraceable_functions_say_why():
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(2, 3))
    for bad, word in ((lambda x: float(np.sum(x ** 2)), "float()"), (lambda x: x[0] if x[0] > 0 else -x[0], "compares"),
                      (lambda x: np.maximum(x, 0).sum(), "compares"), (lambda x: np.array([1.0, 2.0]) * x[0], "array of shape")):
        with pytest.raises(TraceError, match=word.replace("(", r"\(").replace(")", r"\)")):
            trace_energy(bad, 3, pts)
    state = {"n": 0}

    def sneaky(x):  # result does not follow from the recorded arithmetic
        state["n"] += 1
        return np.sum(x ** 2) * state["n"]

    with pytest.raises(TraceError, match="traced expression gives"):
        trace_energy(sneaky, 3, pts)


def test_generated_kernels_compile_with_nvrtc():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from nvrtc_check import nvrtc_compile

    tr, _, grads = trace_energy(CASES["rosenbrock"], 3, np.random.default_rng(0).normal(size=(2, 3)))
    for real in ("float", "double"):
        src = ('#include "langevin_body.cuh"\n' + tr.cuda_source(grads) +
               "struct TsuUserGrad { template <typename real> __device__ __forceinline__ void operator()(int, const real* x, "
               "real* g) const { tsu_user_grad<real>(x, g); } };\n"
               'extern "C" __global__ void __launch_bounds__(128) tsu_jit_langevin(tsu_langevin::LangevinParams P) {\n'
               "  tsu_langevin::langevin_chain<%s, 3>(P, TsuUserGrad());\n}\n" % real)
        cubin, log = nvrtc_compile(src, "tsu_jit_langevin.cu")
        if cubin is None and "not found" in log:
            pytest.skip("libnvrtc is not installed here")
        assert cubin is not None, log
