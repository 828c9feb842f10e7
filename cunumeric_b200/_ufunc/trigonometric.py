"""Trigonometric / hyperbolic ufuncs (cunumeric/_ufunc/trigonometric.py:26-140)."""
from ..config import BinaryOpCode, UnaryOpCode
from .ufunc import create_binary_ufunc, create_unary_ufunc, float_and_complex, float_dtypes


def _fc(name: str, op: UnaryOpCode, summary: str):
    return create_unary_ufunc(summary, name, op, float_and_complex)


sin = _fc("sin", UnaryOpCode.SIN, "Trigonometric sine, element-wise.")
cos = _fc("cos", UnaryOpCode.COS, "Cosine element-wise.")
tan = _fc("tan", UnaryOpCode.TAN, "Compute tangent element-wise.")
arcsin = _fc("arcsin", UnaryOpCode.ARCSIN, "Inverse sine, element-wise.")
arccos = _fc("arccos", UnaryOpCode.ARCCOS, "Trigonometric inverse cosine, element-wise.")
arctan = _fc("arctan", UnaryOpCode.ARCTAN, "Trigonometric inverse tangent, element-wise.")
sinh = _fc("sinh", UnaryOpCode.SINH, "Hyperbolic sine, element-wise.")
cosh = _fc("cosh", UnaryOpCode.COSH, "Hyperbolic cosine, element-wise.")
tanh = _fc("tanh", UnaryOpCode.TANH, "Compute hyperbolic tangent element-wise.")
arcsinh = _fc("arcsinh", UnaryOpCode.ARCSINH, "Inverse hyperbolic sine element-wise.")
arccosh = _fc("arccosh", UnaryOpCode.ARCCOSH, "Inverse hyperbolic cosine, element-wise.")
arctanh = _fc("arctanh", UnaryOpCode.ARCTANH, "Inverse hyperbolic tangent element-wise.")
arctan2 = create_binary_ufunc("Element-wise arc tangent of x1/x2 choosing the quadrant "
                              "correctly.", "arctan2", BinaryOpCode.ARCTAN2, float_dtypes)
hypot = create_binary_ufunc("Given the legs of a right triangle, return its hypotenuse.",
                            "hypot", BinaryOpCode.HYPOT, float_dtypes)
deg2rad = create_unary_ufunc("Convert angles from degrees to radians.", "deg2rad",
                             UnaryOpCode.DEG2RAD, float_dtypes)
radians = deg2rad
rad2deg = create_unary_ufunc("Convert angles from radians to degrees.", "rad2deg",
                             UnaryOpCode.RAD2DEG, float_dtypes)
degrees = rad2deg
