// mt_kernels.cuh -- Marching Tetrahedra generate kernel (placeholder until the owner-rule kernel lands)
#pragma once
#include "iso_kernels.cuh"
#include "../../include/b200iso.h"
namespace iso {
inline int launch_mt_generate(const GenArgs&, const Grid&, const b200iso_params&, int, unsigned, cudaStream_t) { return -1; }
}  // namespace iso
