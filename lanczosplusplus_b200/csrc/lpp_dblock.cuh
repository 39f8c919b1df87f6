// lpp_dblock.cuh -- two-pass block down sweep (lpp_dblock_kernel.cuh) behind the interface of the other down sweeps:
//   x = beta x + alpha (D + 1 (x) T_dn) y   (the spin-down half of HubbardHelper.h:105-134 for HubbardOneBand)
#pragma once
#include "lpp_sweep_common.cuh"

struct DownBlockPlan;

// returns 0 = plan built, 1 = not applicable (caller keeps the streaming kernel), <0 = CUDA error
int lpp_dblock_create(const ModelDev& m, const HopTable& dn, const DiagTables& dt, cudaStream_t s, DownBlockPlan** out);
void lpp_dblock_destroy(DownBlockPlan* p);
const char* lpp_dblock_error();
// all down states must be local (d0 = 0, dcount = n2); 16-byte accesses need an even pitch and an even column count
int lpp_dblock_accepts(const DownBlockPlan* p, const ModelDev& m, const DiagTables& dt, uint64_t d0, uint64_t dcount, const ColView& cv);
// number of dot partial sums the sweep writes for a column view (one per pass-2 tile)
int lpp_dblock_partials(const DownBlockPlan* p, const ColView& cv);
int lpp_dblock_sweep(DownBlockPlan* p, const ModelDev& m, const DiagTables& dt, const SpmvArgs& a, const ColView& cv, cudaStream_t s);
void lpp_dblock_describe(const DownBlockPlan* p, char* buf, size_t n);
