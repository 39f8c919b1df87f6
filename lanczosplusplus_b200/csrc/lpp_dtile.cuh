// lpp_dtile.cuh -- down-spin sweep on shared-memory tiles:  x = beta x + alpha (D + 1 (x) T_dn) y
// (the spin-down half of HubbardHelper.h:105-134 / FeBasedSc.h:228-245 for product bases).
#pragma once
#include "lpp_sweep_common.cuh"

struct DownTilePlan;

// returns 0 = plan built, 1 = not applicable to this model/geometry (caller keeps its streaming kernel), <0 = CUDA error
// dv2_dev: DiagTables::dv2 (device), copied into the row records
int lpp_dtile_create(const ModelDev& m, const HopTable& dn, const double* dv2_dev, const MagTable& mt, cudaStream_t s,
                     DownTilePlan** out);
void lpp_dtile_destroy(DownTilePlan* p);
const char* lpp_dtile_error();
// does the kernel accept this column view (16-byte accesses need an even pitch and an even column count)?
int lpp_dtile_accepts(const DownTilePlan* p, const ColView& cv);
// number of CTAs (= number of dot partial sums) for a column view
int lpp_dtile_grid(const DownTilePlan* p, const ColView& cv);
// rows [d0, d0+dcount) of the down index are local (x holds only those); y is indexed by the global down index
int lpp_dtile_sweep(DownTilePlan* p, const ModelDev& m, const DiagTables& dt, const SpmvArgs& a, uint64_t d0, uint64_t dcount,
                    const ColView& cv, cudaStream_t s);
void lpp_dtile_describe(const DownTilePlan* p, char* buf, size_t n);

// ---------------------------------------------------------------------------------------------------------------
// Row-walking variant of the down sweep (k_sweep_down_rows): no staged tile; a CTA of 24 warps walks consecutive down
// states for one group of 16 columns, a quarter-warp per row, and gathers straight from global memory, so that the
// hop sources shared by neighbouring rows (colex neighbours differ in the low sites) are L1 hits instead of L2 reads.
struct DownRowsPlan;
int lpp_drows_create(const ModelDev& m, const HopTable& dn, const MagTable& mt, cudaStream_t s, DownRowsPlan** out);
void lpp_drows_destroy(DownRowsPlan* p);
int lpp_drows_accepts(const DownRowsPlan* p, const ColView& cv);
int lpp_drows_grid(const DownRowsPlan* p, const ColView& cv);
int lpp_drows_sweep(DownRowsPlan* p, const ModelDev& m, const DiagTables& dt, const SpmvArgs& a, uint64_t d0, uint64_t dcount,
                    const ColView& cv, cudaStream_t s);
