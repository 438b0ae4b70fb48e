"""Import shim: `import quant_cuda` resolves to the adaqp_b200 codec
(reference: AdaQP/model/op_util.py:6 imports the pybind module of this name)."""
from adaqp_b200.quant import pack_single_precision, unpack_single_precision  # noqa: F401
