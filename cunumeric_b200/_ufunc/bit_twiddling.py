"""Bit-twiddling ufuncs (cunumeric/_ufunc/bit_twiddling.py:21-65)."""
from ..config import BinaryOpCode, UnaryOpCode
from .ufunc import create_binary_ufunc, create_unary_ufunc, integer_dtypes

bitwise_and = create_binary_ufunc("Compute the bit-wise AND of two arrays element-wise.",
                                  "bitwise_and", BinaryOpCode.BITWISE_AND, ["?"] + integer_dtypes)
bitwise_or = create_binary_ufunc("Compute the bit-wise OR of two arrays element-wise.",
                                 "bitwise_or", BinaryOpCode.BITWISE_OR, ["?"] + integer_dtypes)
bitwise_xor = create_binary_ufunc("Compute the bit-wise XOR of two arrays element-wise.",
                                  "bitwise_xor", BinaryOpCode.BITWISE_XOR, ["?"] + integer_dtypes)
# bool is remapped to LOGICAL_NOT (bit_twiddling.py:42-48)
invert = create_unary_ufunc("Compute bit-wise inversion, or bit-wise NOT, element-wise.",
                            "invert", UnaryOpCode.INVERT, ["?"] + integer_dtypes,
                            overrides={"?": UnaryOpCode.LOGICAL_NOT})
bitwise_not = invert
left_shift = create_binary_ufunc("Shift the bits of an integer to the left.", "left_shift",
                                 BinaryOpCode.LEFT_SHIFT, integer_dtypes)
right_shift = create_binary_ufunc("Shift the bits of an integer to the right.", "right_shift",
                                  BinaryOpCode.RIGHT_SHIFT, integer_dtypes)
