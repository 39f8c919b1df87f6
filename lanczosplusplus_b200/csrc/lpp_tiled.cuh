// lpp_tiled.cuh -- the product-basis fast path: H = D + T_up (x) 1 + 1 (x) T_dn applied as two sweeps.
#pragma once
#include "lpp_kernels.cuh"

int lpp_tiled_create(const ModelDev& m, const double* hop_host, const HopTable& up, const HopTable& dn, const DiagTables& dt,
                     uint64_t row0,
                     uint64_t nloc, cudaStream_t s, TiledPlan** out);
void lpp_tiled_destroy(TiledPlan* p);
const char* lpp_tiled_error();
int lpp_tiled_dot_blocks(const TiledPlan* p);
// returns the number of kernels launched, or <0 on error
int lpp_tiled_spmv(TiledPlan* p, const ModelDev& m, const HopTable& up, const HopTable& dn, const DiagTables& dt,
                   const SpmvArgs& a, cudaStream_t s);

// two-layout multi-GPU (see lpp_tiled.cu)
int lpp_tiled_two_layout_ok(const TiledPlan* p);
int lpp_tiled_up_rows_blocks(const TiledPlan* p, uint64_t nrows);
int lpp_tiled_sweep_up_rows(TiledPlan* p, const ModelDev& m, const SpmvArgs& a, uint64_t nrows, cudaStream_t s);
int lpp_tiled_down_cols_blocks(const TiledPlan* p, const ModelDev& m, uint64_t ncols);
int lpp_tiled_sweep_down_cols(TiledPlan* p, const ModelDev& m, const HopTable& dn, const DiagTables& dt, const SpmvArgs& a,
                              uint64_t u0, uint64_t ncols, cudaStream_t s);
