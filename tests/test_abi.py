"""The C-ABI library loads without a GPU and exports every symbol include/multilinear_b200.h declares."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="multilinear_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ml_[a-z0-9_]+)\s*\(", text)))


def test_header_is_plain_c():
    src = "#include \"multilinear_b200.h\"\nint main(void){return ML_OK;}\n"
    p = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src.encode(), capture_output=True)
    assert p.returncode == 0, p.stderr.decode()


def test_library_exports_every_declared_symbol(ml):
    import multilinear_b200
    lib = ctypes.CDLL(multilinear_b200.lib_path())
    names = declared_symbols()
    assert len(names) > 90
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_instrumentation_is_outside_the_product_header(ml):
    """profiling / microbenchmark entry points are declared in multilinear_b200_instr.h only; the microbenchmark kernels are
    not linked into the product library"""
    import multilinear_b200
    product = set(declared_symbols())
    instr = set(declared_symbols("multilinear_b200_instr.h"))
    assert instr and not (product & instr)
    assert not [n for n in product if "profile" in n or "microbench" in n or "trace" in n]
    lib = ctypes.CDLL(multilinear_b200.lib_path())
    assert not hasattr(lib, "ml_microbench")
    ilib = ctypes.CDLL(os.path.join(os.path.dirname(multilinear_b200.lib_path()), "libmlb_instr.so"))
    assert all(hasattr(lib, n) or hasattr(ilib, n) for n in instr)


def test_no_torch_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "multilinear_b200.h")).read()
    assert "torch" not in text and "at::" not in text and "Tensor" not in text


def test_product_does_not_reference_oracle():
    # the oracle is test infrastructure; the product path must never import, link or call it
    pkg = os.path.join(ROOT, "multilinear_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                s = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in s.lower().replace("test infrastructure", ""), os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libmultilinear_b200.so")], capture_output=True).stdout.decode()
    assert "liboracle" not in out


def test_version_and_errors(ml):
    from multilinear_b200 import load
    L = load()
    assert b"sm_100a" in L.ml_version()
    out = (ctypes.c_uint8 * 16)()
    assert L.ml_pow2_generator(ctypes.c_uint64(41), out) == 3  # None
    assert b"None" in L.ml_last_error()
