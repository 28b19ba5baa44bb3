// TEST INFRASTRUCTURE ONLY -- part of the oracle/_ref build recipe, never shipped.
//
// Same bridge as fully_connected_lattice_compat.cpp, for the reference's CUDA orchestration file:
// streamlined_fast_gpu.cpp:263 spells `.sum({0, 1})`, ambiguous under torch 2.x.  All headers are pulled
// in first (include guards keep them away from the macro), then the reference file is included from
// where it lies with that one call re-spelled.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <ATen/cuda/CUDAEvent.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>
#include "streamlined_fast_gpu.h"
#include "fully_connected_lattice.h"
#include "force_aligned_lattice.h"
#include "force_aligned_lattice_kernel.h"
#include "utils.h"

#define sum(...) sum(at::IntArrayRef(__VA_ARGS__))
#include "streamlined_fast_gpu.cpp"
#undef sum
