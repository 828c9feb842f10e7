"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the header
declares, and the opcode enums are numerically identical to the reference's cunumeric_c.h."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cunumeric_b200.h")


def header_text():
    with open(HEADER) as f:
        return f.read()


def test_library_loads_and_exports_every_declared_symbol():
    import ctypes

    from cunumeric_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    text = re.sub(r"/\*.*?\*/", "", header_text(), flags=re.S)
    names = re.findall(r"\b((?:cnb|cunumeric)_[a-z0-9_]+)\s*\(", text)
    names = sorted(set(n for n in names if not n.endswith("_t")))
    assert len(names) >= 45, names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cunumeric_b200.h but not exported"
    # and the Python binding covers exactly the same surface
    assert sorted(_lib.PROTOTYPES) == names


def test_header_cites_reference_interfaces():
    text = header_text()
    for cite in ("binary_op.cu:85-88", "unary_op.cu:178-181", "where.cu:74-77", "convert.cu:71-74",
                 "scalar_unary_red.cu:26-29", "unary_red.cu:342-345", "cunumeric_c.h:337-339"):
        assert cite in text, cite


def _c_enum(text, prefix):
    body = re.search(r"typedef enum \w+ \{([^}]*" + prefix + r"[^}]*)\}", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    vals, cur = {}, -1
    for item in body.split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            name, v = (x.strip() for x in item.split("="))
            cur = int(v)
        else:
            name, cur = item, cur + 1
        vals[name] = cur
    return vals


def test_opcode_enums_match_python_and_reference_numbering():
    from cunumeric_b200 import config

    text = header_text()
    un = _c_enum(text, "CNB_UOP_")
    assert {k[len("CNB_UOP_"):]: v for k, v in un.items()} == {m.name: m.value for m in config.UnaryOpCode}
    bi = _c_enum(text, "CNB_BINOP_")
    assert {k[len("CNB_BINOP_"):]: v for k, v in bi.items()} == {m.name: m.value for m in config.BinaryOpCode}
    rd = _c_enum(text, "CNB_RED_")
    assert {k[len("CNB_RED_"):]: v for k, v in rd.items()} == {m.name: m.value for m in config.UnaryRedCode}
    # spot values from src/cunumeric/cunumeric_c.h (alphabetical from 1; task ids by position)
    assert config.UnaryOpCode.ABSOLUTE == 1 and config.UnaryOpCode.TRUNC == 47
    assert config.BinaryOpCode.ADD == 1 and config.BinaryOpCode.SUBTRACT == 35
    assert config.UnaryRedCode.ALL == 1 and config.UnaryRedCode.VARIANCE == 18
    assert config.ConvertCode.NOOP == 1 and config.ConvertCode.SUM == 3
    assert (config.CuNumericOpCode.BINARY_OP, config.CuNumericOpCode.CONVERT,
            config.CuNumericOpCode.SCALAR_UNARY_RED, config.CuNumericOpCode.UNARY_OP,
            config.CuNumericOpCode.UNARY_RED, config.CuNumericOpCode.WHERE) == (5, 11, 33, 43, 44, 49)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/cunumeric"),
                    reason="reference tree not mounted")
def test_enums_against_the_reference_header():
    """Where the reference is mounted, parse its cunumeric_c.h and compare every opcode."""
    from cunumeric_b200 import config

    with open("/root/reference/src/cunumeric/cunumeric_c.h") as f:
        text = f.read()

    def parse(enum_name):
        body = re.search(r"enum " + enum_name + r" \{(.*?)\};", text, flags=re.S).group(1)
        body = re.sub(r"//.*", "", body)
        vals, cur = {}, -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, v = (x.strip() for x in item.split("="))
                cur = int(v)
            else:
                name, cur = item, cur + 1
            vals[name] = cur
        return vals

    ops = parse("CuNumericOpCode")
    for m in config.CuNumericOpCode:
        assert ops["CUNUMERIC_" + m.name] == m.value
    for enum, prefix, py in (("CuNumericUnaryOpCode", "CUNUMERIC_UOP_", config.UnaryOpCode),
                             ("CuNumericUnaryRedCode", "CUNUMERIC_RED_", config.UnaryRedCode),
                             ("CuNumericBinaryOpCode", "CUNUMERIC_BINOP_", config.BinaryOpCode),
                             ("CuNumericConvertCode", "CUNUMERIC_CONVERT_NAN_", config.ConvertCode)):
        vals = parse(enum)
        assert {k[len(prefix):]: v for k, v in vals.items()} == {m.name: m.value for m in py}


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly (never route to the oracle / NumPy)."""
    from cunumeric_b200 import _lib

    lib = _lib.load()
    if lib.cnb_device_count() > 0:
        pytest.skip("a GPU is present")
    import cunumeric_b200 as cn

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cn.zeros((4,))
    d = _lib.cnb_store_t()
    assert lib.cnb_fill(d, None, None) == -3  # CNB_ERR_CUDA
    assert b"CUDA" in lib.cnb_last_error() or b"no CUDA" in lib.cnb_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cunumeric_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
                assert "libcunumeric_ref" not in src, fn


def test_argval_dtype_layout():
    from cunumeric_b200.config import argval_dtype, dtype_code

    for dt in (np.bool_, np.int8, np.float16, np.float32, np.float64, np.uint64):
        av = argval_dtype(dt)
        assert av.itemsize == 16 and av.fields["arg"][1] == 0 and av.fields["arg_value"][1] == 8
        assert dtype_code(av) == 32 + dtype_code(dt)
