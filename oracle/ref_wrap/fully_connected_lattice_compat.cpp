// TEST INFRASTRUCTURE ONLY -- part of the oracle/_ref build recipe, never shipped.
//
// The reference's fully_connected_lattice.cpp calls `.sum({0, 1})` (line 60), which is
// ambiguous under torch 2.x (IntArrayRef vs DimnameList).  Rather than copying or editing
// the reference source, this translation unit pulls in every header first (so include
// guards keep them out of the macro's reach) and then re-spells the one ambiguous call
// through the preprocessor while including the reference file from where it lies.
#include <torch/extension.h>
#include "fully_connected_lattice.h"

#define sum(...) sum(at::IntArrayRef(__VA_ARGS__))
#include "fully_connected_lattice.cpp"
#undef sum
