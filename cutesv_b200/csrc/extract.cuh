// extract.cuh -- signature extraction kernels (CIGAR / SA walk); see extract_api.inl
#pragma once
#include "core.h"
namespace csv {
struct ExtractState { int unused = 0; };
inline void extract_release(ExtractState*) {}
}  // namespace csv
