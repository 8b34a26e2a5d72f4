"""CPU checks of the drop-in boundary: liblmm.so loads, exports every symbol include/lmm.h
declares (and nothing is bound that the header does not declare), host-only entry points behave,
and compute entry points fail loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import lmm_b200 as lmm
from lmm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "lmm.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lmm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"liblmm.so does not export {n}"
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table covers the header exactly
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (lmm_[a-z0-9_]+)", out)))
    assert exported == names


def test_documented_options_match_the_library():
    """Every tunable include/lmm.h documents is accepted by lmm_ctx_set_option's dispatcher and vice versa (the header is
    the only place a caller of the C ABI learns about them)."""
    hdr = open(os.path.join(ROOT, "include", "lmm.h")).read()
    block = hdr[hdr.index("Tunables (key, value)"):hdr.index("int lmm_ctx_set_option")]
    documented = set(re.findall(r'^ \*   "([a-z_]+)"', block, flags=re.M))
    src = open(os.path.join(ROOT, "linearmixingmodels.jl_b200", "csrc", "api.cu")).read()
    accepted = set(re.findall(r'k == "([a-z_]+)"', src))
    assert documented == accepted, (sorted(documented - accepted), sorted(accepted - documented))


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_shim_ccalls_match_the_header():
    """The Julia shim cannot be executed here (no Julia), so its ccall signatures are checked mechanically against
    include/lmm.h: same symbol, same number of arguments, pointer / int / double / string in the same positions, and one
    value passed per declared argument."""
    hdr = open(os.path.join(ROOT, "include", "lmm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(lmm_[a-z0-9_]+)\s*\(([^;{]*)\)\s*;", hdr):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        kinds = []
        for a in args:
            if a in ("void", ""):
                continue
            if "*" in a:
                kinds.append("str" if re.match(r"const\s+char\s*\*", a) else "ptr")
            elif re.match(r"(const\s+)?double\b", a):
                kinds.append("double")
            else:
                kinds.append("int")
        protos[m.group(1)] = kinds
    jl = open(os.path.join(ROOT, "linearmixingmodels.jl_b200", "julia", "LinearMixingModelsB200.jl")).read()
    jl = re.sub(r"#=.*?=#", "", jl, flags=re.S)
    jl = "\n".join(line.split("#")[0] if not line.lstrip().startswith("#") else "" for line in jl.split("\n"))
    seen = set()
    for m in re.finditer(r"ccall\(\(:(lmm_[a-z0-9_]+), liblmm\),", jl):
        name = m.group(1)
        assert name in protos, f"{name} is not declared in lmm.h"
        # take the balanced argument list of this ccall
        i = m.start() + len("ccall(")
        depth, j = 1, i
        while depth:
            ch = jl[j]
            depth += ch in "({["
            depth -= ch in ")}]"
            j += 1
        parts = _split_top(jl[i:j - 1])
        types = _split_top(parts[2].strip()[1:-1]) if parts[2].strip() != "()" else []
        values = parts[3:]
        kinds = []
        for t in types:
            if t.startswith("Ptr{"):
                kinds.append("ptr")
            elif t == "Cstring":
                kinds.append("str")
            elif t == "Float64":
                kinds.append("double")
            elif t in ("Cint", "Int32"):
                kinds.append("int")
            else:
                raise AssertionError(f"{name}: unexpected ccall type {t}")
        assert kinds == protos[name], f"{name}: shim {kinds} vs header {protos[name]}"
        assert len(values) == len(kinds), f"{name}: {len(values)} values for {len(kinds)} declared arguments"
        seen.add(name)
    assert len(seen) >= 25


def test_struct_layout_matches_header():
    assert C.sizeof(_lib.GpDesc) == 64
    assert _lib.GpDesc.variance.offset == 8 and _lib.GpDesc.mean_const.offset == 24 and _lib.GpDesc.ard.offset == 32
    assert _lib.GpDesc.param.offset == 40 and _lib.GpDesc.n_extra.offset == 48 and _lib.GpDesc.extra.offset == 56
    assert C.sizeof(_lib.KernelTerm) == 40 and _lib.KernelTerm.param.offset == 24 and _lib.KernelTerm.ard.offset == 32


def test_version_and_host_only_entry_points():
    lib = _lib.load()
    assert b"sm_100a" in lib.lmm_version()
    U, _, _ = np.linalg.svd(np.random.default_rng(0).uniform(size=(5, 3)), full_matrices=False)
    Uf = np.asfortranarray(U)
    assert lib.lmm_orthogonal_validate(_lib.ptr(Uf), 5, 3) == 0
    bad = np.asfortranarray(np.random.default_rng(1).uniform(size=(5, 3)))
    assert lib.lmm_orthogonal_validate(_lib.ptr(bad), 5, 3) == _lib.LMM_E_NOT_ORTHOGONAL
    # known answers of test/independent_mogp.jl:86-98 (1-based there)
    out = np.zeros(6, dtype=np.int64)
    assert lib.lmm_reorder_indices(3, 2, 0, out.ctypes.data_as(C.c_void_p)) == 0
    np.testing.assert_array_equal(out + 1, [1, 4, 2, 5, 3, 6])
    assert lib.lmm_reorder_indices(3, 2, 1, out.ctypes.data_as(C.c_void_p)) == 0
    np.testing.assert_array_equal(out + 1, [1, 3, 5, 2, 4, 6])


def test_host_mirror_types_and_errors():
    """test/orthogonal_matrix.jl:1-15, test/ilmm.jl:55-72 (known answers on the host side)."""
    rng = np.random.default_rng(2)
    with pytest.raises(ValueError, match="not an orthogonal matrix"):
        lmm.Orthogonal(rng.uniform(size=(3, 2)), np.ones(2))
    U, S, _ = np.linalg.svd(rng.uniform(size=(3, 2)), full_matrices=False)
    H = lmm.Orthogonal(U, S)
    assert H.shape == (3, 2)
    np.testing.assert_allclose(np.asarray(H), U @ np.diag(np.sqrt(S)))
    fs = lmm.independent_mogp([lmm.GP(lmm.SEKernel()), lmm.GP(2.0, 0.5 * lmm.Matern32Kernel())])
    f = lmm.ILMM(fs, H)
    assert isinstance(f, lmm.OILMM) and lmm.get_latent_gp(f) is fs
    assert not isinstance(lmm.ILMM(fs, np.asarray(H)), lmm.OILMM)
    assert fs.fs[1].kernel.variance == 0.5 and fs.fs[1].mean_const == 2.0
    assert lmm.with_lengthscale(lmm.SEKernel(), 4.0).inv_lengthscale == 0.25
    x = lmm.MOInputIsotopicByOutputs(np.arange(4.0), 3)
    fx = f(x, 2.0)
    assert len(fx) == 12 and lmm.noise_var(fx) == 2.0  # noise_var(Diagonal(Fill(2, n))) == 2
    assert lmm.reshape_y(np.arange(16.0), 8).shape == (2, 8) and lmm.reshape_y(np.arange(16.0), 2).shape == (8, 2)
    lat, HH, s2, xx = lmm.unpack(fx)
    assert lat is fs and HH is H and s2 == 2.0 and xx is x.x
    with pytest.raises(RuntimeError, match="out dim of x != out dim of f."):
        lmm.unpack(f(lmm.MOInputIsotopicByOutputs(np.arange(4.0), 2), 0.1))
    with pytest.raises(TypeError):
        lmm.unpack(f(lmm.MOInputIsotopicByFeatures(np.arange(4.0), 3), 0.1))
    np.testing.assert_array_equal(lmm.indices_which_reorder_outputs_to_features(lmm.MOInputIsotopicByOutputs(np.arange(3.0), 2)) + 1,
                                  [1, 4, 2, 5, 3, 6])
    # prior marginals of an IndependentMOGP need no device
    M, V = lmm.mean_and_var(fs(lmm.MOInputIsotopicByOutputs(np.arange(3.0), 2), 0.1))
    np.testing.assert_allclose(M, [0, 0, 0, 2, 2, 2])
    np.testing.assert_allclose(V, [1.1, 1.1, 1.1, 0.6, 0.6, 0.6])


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is for CPU-only hosts")
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.lmm_ctx_create(0, C.byref(h)) == _lib.LMM_E_CUDA
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lmm.Context(0)
    buf = C.create_string_buffer(128)
    assert lib.lmm_comm_unique_id(C.cast(buf, C.c_void_p)) in (0, _lib.LMM_E_NCCL)


def test_product_path_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import, load or call it."""
    pkg = os.path.join(ROOT, "linearmixingmodels.jl_b200")
    pat = re.compile(r"(from|import)\s+oracle|lmm_oracle|oracle/")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".jl")):
                assert not pat.search(open(os.path.join(dirpath, fn)).read()), fn


C_CLIENT = os.path.join(ROOT, "examples", "c_client")


def _build_c_client():
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_client.c"),
                    "-L" + os.path.dirname(_lib.LIB_PATH), "-llmm", "-lm", "-Wl,-rpath,$ORIGIN/../linearmixingmodels.jl_b200", "-o", C_CLIENT],
                   check=True)


def test_header_is_valid_c_and_plain_c_client_links():
    """include/lmm.h compiles as C99 and a program that includes nothing but it links against liblmm.so; without a
    GPU the client reports the missing device (exit code 77): there is no CPU fallback to fall into."""
    _build_c_client()
    r = subprocess.run([C_CLIENT], capture_output=True, text=True)
    import torch

    if not torch.cuda.is_available():
        assert r.returncode == 77 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_plain_c_client_runs_on_gpu(tmp_path):
    """The drop-in boundary without Python in the call path: examples/c_client.c (BASELINE config 1 through the C ABI)."""
    _build_c_client()
    r = subprocess.run([C_CLIENT, str(tmp_path / "post.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_client ok" in r.stdout


@pytest.mark.parametrize("eps,ok", [(0.5e-8, True), (1.3e-8, True), (1.7e-8, False), (3e-8, False)])
def test_orthogonal_validate_uses_the_operator_norm(eps, ok):
    """Julia's `isapprox(U'U, I)` against a UniformScaling compares OPERATOR 2-norms with |I| = 1
    (src/orthogonal_matrix.jl:21-23): a rank-one deviation eps*v*v' of U'U must be rejected as soon as eps exceeds
    sqrt(eps(Float64)) = 1.49e-8 -- a Frobenius test with |I|_F = sqrt(m) would accept it up to sqrt(m) times that
    (ADVICE r01).  Library and oracle must agree on both sides of the threshold."""
    from oracle import lmm_oracle as o

    p, m = 24, 16
    rng = np.random.default_rng(3)
    Q, _ = np.linalg.qr(rng.standard_normal((p, m)))
    v = rng.standard_normal(m)
    v /= np.linalg.norm(v)
    U = np.asfortranarray(Q @ (np.eye(m) + 0.5 * eps * np.outer(v, v)))  # U'U = I + eps vv' + O(eps²)
    rc = _lib.load().lmm_orthogonal_validate(_lib.ptr(U), p, m)
    assert (rc == 0) == ok
    if ok:
        o.validate_orthogonal(U)
    else:
        with pytest.raises(ValueError):
            o.validate_orthogonal(U)


def test_julia_shim_describes_shape_parameter_and_ard():
    """the shim used to pass param = 1.0 and ard = C_NULL for every kernel (RationalQuadraticKernel's default
    α = 2 evaluated with α = 1, ARDTransform rejected).  Mechanical check of the source (Julia cannot run here): `describe`
    carries the shape parameter and the ARD vector, handles KernelSum / KernelProduct / PeriodicKernel, `gpdescs` stores the
    ARD / extra-term pointers and every ccall that takes the descriptors runs under `GC.@preserve keep`."""
    jl = open(os.path.join(ROOT, "linearmixingmodels.jl_b200", "julia", "LinearMixingModelsB200.jl")).read()
    assert "describe(k::TransformedKernel{<:Kernel,<:ARDTransform})" in jl
    assert "describe(k::KernelSum)" in jl and "describe(k::KernelProduct)" in jl and "kind(::PeriodicKernel)" in jl
    assert "(kind(k), 1.0, 1.0, shape(k), nothing)" in jl
    assert "pointer(a)" in jl and "pointer(extra)" in jl
    assert "GpDesc(t0[1], op, t0[2], t0[3], meanconst(f.mean), p, t0[4], length(extra), 0, px)" in jl
    assert "GpDesc.(" not in jl  # no call site builds descriptors without the keep-alive list any more
    # the Julia structs mirror the C layout field for field
    hdr = open(os.path.join(ROOT, "include", "lmm.h")).read()
    c_fields = re.findall(r"^\s+(?:const\s+)?(?:int32_t|double|lmm_kernel_term)\s*\*?\s*(\w+);", hdr[hdr.index("typedef struct lmm_gp_desc {"):hdr.index("} lmm_gp_desc;")], flags=re.M)
    jl_fields = re.findall(r"^\s+(\w+)::", jl[jl.index("struct GpDesc"):jl.index("const CTX")], flags=re.M)
    assert c_fields == jl_fields, (c_fields, jl_fields)
    # `keep` only exists where gpdescs built it: no stray `GC.@preserve keep` in another definition (per top-level definition)
    code = "\n".join(ln.split("#")[0] for ln in jl.split("\n"))  # comments may talk about it
    for defn in re.split(r"(?m)^(?=function |[a-z_]+\(.*\) = )", code):
        if "GC.@preserve keep" in defn:
            assert "= gpdescs(" in defn, defn[:100]
    for fn_src in re.split(r"(?m)^(?=function )", jl):
        if "= gpdescs(" in fn_src and not fn_src.startswith("function gpdescs"):
            n_desc_calls = len(re.findall(r"ccall\(\(:lmm_[a-z0-9_]+, liblmm\), Cint,\s*\(Ptr\{Cvoid\}, Ptr\{GpDesc\}", fn_src))
            assert n_desc_calls >= 1
            assert fn_src.count("GC.@preserve keep ccall(") >= n_desc_calls, fn_src[:80]
