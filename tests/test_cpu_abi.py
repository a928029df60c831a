"""CPU tests of the drop-in boundary: libnfa.so loads, exports every symbol include/nfa.h declares, and the host
layer refuses to compute without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from nfempy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "nfa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(nfa_[a-z0-9_]+)\s*\(", text))


def test_library_loads_and_reports_version():
    lib = _lib.load()
    assert lib.nfa_version() >= 100
    assert isinstance(_lib.launch_count(), int)


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert declared <= exported, sorted(declared - exported)
    assert declared == set(_lib.SIGNATURES), sorted(declared ^ set(_lib.SIGNATURES))


def test_abi_has_no_torch_or_cxx_types_in_signatures():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    nfa = [ln.split()[-1] for ln in out.splitlines() if " T " in ln and "nfa_" in ln.split()[-1]]
    assert all(not s.startswith("_Z") for s in nfa)  # extern "C": unmangled
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    # library names only: the load addresses ldd prints are randomised and may contain "c10" as hex digits
    names = [ln.split()[0] for ln in ldd.splitlines() if ln.strip()]
    assert not any("torch" in n or "c10" in n for n in names), names


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_gpu():
    from nfempy_b200.assembly import GPUAssembler
    from tests.parity_helpers import elasticity_case

    case = elasticity_case(n=1, with_bc=False)
    space = case.space()
    asm = GPUAssembler(space, None, np.zeros(space.n_dof))
    asm.form_to_input_list = case.gpu_forms(space, np.zeros(space.n_dof))
    with pytest.raises(_lib.NfaError):
        asm.scatter_forms()


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may touch oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "nfempy_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_error_behaviour_of_the_assembler_front_end():
    """The reference's error behaviour at the boundary, checked without a GPU: unknown form
    types raise KeyError like its dispatch dictionary, known types without a kernel and foreign weak-form classes raise
    NotImplementedError (never a silent host path), an unset form list / unset ``functions`` raise ValueError."""
    from nfempy_b200.assembly import GPUAssembler
    from nfempy_b200.weak_forms import LEPrimalWeakForm, LEPrimalWeakFormBCDirichlet, WeakForm
    from tests.parity_helpers import elasticity_case

    case = elasticity_case(n=1, with_bc=True)
    space = case.space()
    asm = GPUAssembler(space, None, np.zeros(space.n_dof))
    with pytest.raises(ValueError):
        asm._entries()  # form_to_input_list not set
    form = LEPrimalWeakForm(space)
    form.functions = {"rhs": lambda x, y, z: np.zeros((3,) + np.shape(x)), "lambda": lambda x, y, z: 1.0, "mu": lambda x, y, z: 1.0}
    asm.form_to_input_list = {"v": ["no_such_form_type", range(6), form, None]}
    with pytest.raises(KeyError):
        asm._entries()
    asm.form_to_input_list = {"v": ["bc_interface_form", range(6), form, None]}  # the library has no form of this type
    with pytest.raises(NotImplementedError):
        asm._entries()
    asm.form_to_input_list = {"v": ["interface_form", range(6), form, None]}  # a volume form listed as an interface form
    with pytest.raises(ValueError):
        asm._entries()
    asm.form_to_input_list = {"v": ["bc_form", range(6), form, None]}  # a volume form listed as a boundary form
    with pytest.raises(ValueError):
        asm._entries()
    asm.form_to_input_list = {"v": ["form", [0, 99], form, None]}  # cell index out of range
    with pytest.raises(IndexError):
        asm._entries()

    class Foreign(WeakForm):  # a user form without a CUDA kernel
        def evaluate_form(self, element_index, alpha):
            return None

    asm.form_to_input_list = {"v": ["form", range(6), Foreign(space), None]}
    with pytest.raises(NotImplementedError):
        asm._entries()
    with pytest.raises(ValueError):  # functions unset (R/src/weak_forms/le_primal_weak_form.py:15-16)
        LEPrimalWeakFormBCDirichlet(space).evaluate_form(0, None)
    with pytest.raises(ValueError):
        GPUAssembler(space, pattern_policy="no_such_policy")


def test_chunked_residual_add_matches_numpy(monkeypatch):
    """``residual += rhs`` of ``scatter_forms()`` for long vectors: chunks on a few host threads, same result."""
    import numpy as np

    from nfempy_b200.assembly import gpu_assembler

    monkeypatch.setattr(gpu_assembler, "_CHUNK", 101)
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(1013), rng.standard_normal(1013)
    want = a + b
    gpu_assembler._add_into(a, b)
    assert np.array_equal(a, want)
    c = np.ones(5)
    gpu_assembler._add_into(c, np.arange(5.0))  # short vectors: one plain numpy call
    assert np.array_equal(c, 1.0 + np.arange(5.0))


def test_threaded_coefficient_evaluation_matches_one_call(monkeypatch):
    """Coefficient callables over many points are evaluated chunk by chunk on host threads: same arrays as one call, for
    scalar, constant, vector and time-dependent callables; callables that do not broadcast still work point by point."""
    import math

    import numpy as np

    from nfempy_b200 import coefficients

    rng = np.random.default_rng(5)
    x = rng.random((37, 4, 3))
    cases = [
        (lambda x, y, z, t: np.array([np.sin(x + t) + y]), 1, (0.3,)),
        (lambda x, y, z: 2.0 + x, None, ()),
        (lambda x, y, z: 1.5, None, ()),
        (lambda x, y, z: np.array([x * y, z + 1.0, np.cos(x)]), 3, ()),
        (lambda x, y, z: np.array([1.0, 2.0, 3.0]), 3, ()),
    ]
    for fn, nc, extra in cases:
        monkeypatch.setattr(coefficients, "_EVAL_CHUNK", 1 << 40)
        want = coefficients.evaluate(fn, x, nc, extra)
        monkeypatch.setattr(coefficients, "_EVAL_CHUNK", 11)
        got = coefficients.evaluate(fn, x, nc, extra)
        assert got.shape == want.shape and np.array_equal(got, want)
    got = coefficients.evaluate(lambda x, y, z: math.sin(x) + y, x)  # scalars only: the point-by-point convention
    assert np.allclose(got, np.sin(x[:, :, 0]) + x[:, :, 1])
