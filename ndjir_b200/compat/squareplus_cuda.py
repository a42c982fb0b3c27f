"""Stand-in for `squareplus_cuda` (csrc/activation/squareplus_cuda.cu:95-99)."""
from .._lib import call


def forward(size, output_ptr, input_ptr, b):
    call("ndjir_squareplus_forward", size, output_ptr, input_ptr, b, 0)


def backward(size, dinput_ptr, doutput_ptr, input_ptr, b, accum):
    call("ndjir_squareplus_backward", size, dinput_ptr, doutput_ptr, input_ptr, b, int(accum), 0)
