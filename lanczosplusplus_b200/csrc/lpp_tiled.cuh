// lpp_tiled.cuh -- the product-basis fast path: H = D + T_up (x) 1 + 1 (x) T_dn applied as two sweeps.
#pragma once
#include "lpp_kernels.cuh"

int lpp_tiled_create(const ModelDev& m, const double* hop_host, const HopTable& up, const HopTable& dn, uint64_t row0,
                     uint64_t nloc, cudaStream_t s, TiledPlan** out);
void lpp_tiled_destroy(TiledPlan* p);
const char* lpp_tiled_error();
int lpp_tiled_dot_blocks(const TiledPlan* p);
// returns the number of kernels launched, or <0 on error
int lpp_tiled_spmv(TiledPlan* p, const ModelDev& m, const HopTable& up, const HopTable& dn, const DiagTables& dt,
                   const SpmvArgs& a, cudaStream_t s);
