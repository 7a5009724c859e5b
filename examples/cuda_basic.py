"""The reference's examples/cuda_basic.py with the import switched (and torch tensors instead of
PyCUDA GPUArrays -- PyCUDA arrays work the same way through .gpudata):

    python examples/cuda_basic.py
"""
import os
import sys

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a source checkout
from pyfft_b200.cuda import Plan           # was: from pyfft.cuda import Plan

stream = torch.cuda.Stream()

# create plan (no Mako render / nvcc run: the kernels are compiled ahead of time)
plan = Plan((16, 16), stream=stream)

# prepare data
data = numpy.ones((16, 16), dtype=numpy.complex64)
gpu_data = torch.from_numpy(data).cuda()
print(gpu_data[:2, :2])

# forward transform, in place; with stream= given, execute() returns the stream instead of waiting
plan.execute(gpu_data).synchronize()
result = gpu_data.cpu().numpy()
print(result[:2, :2])                       # 256 at [0, 0], zeros elsewhere (doc/source/index.rst:61-99)

# inverse transform
plan.execute(gpu_data, inverse=True).synchronize()
result = gpu_data.cpu().numpy()
error = numpy.abs(numpy.sum(numpy.abs(data) - numpy.abs(result)) / data.size)
print(error < 1e-6)

# batched, split re/im layout, out of place: 256 transforms of 1024 x 1024 (BASELINE config 3 shape)
plan2 = Plan((1024, 1024), dtype=numpy.float32)
re, im = torch.randn(4, 1024, 1024, device="cuda"), torch.randn(4, 1024, 1024, device="cuda")
ore, oim = torch.empty_like(re), torch.empty_like(im)
plan2.execute(re, im, ore, oim, batch=4)
ref = torch.fft.fft2(torch.complex(re, im))
print(float((torch.complex(ore, oim) - ref).abs().max() / ref.abs().max()) < 1e-5)
