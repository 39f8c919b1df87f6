// lpp_dblock.cu -- engine side of the multi-pass block down sweep (kernel, plan builder and launcher: lpp_dblock_kernel.cuh).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "lpp_dblock.cuh"
#include "lpp_dblock_kernel.cuh"

static thread_local std::string g_dberr;
const char* lpp_dblock_error() { return g_dberr.c_str(); }

struct DownBlockPlan {
	DbHostPlan host;
	DbDevPlan dev;
	uint32_t* w1 = nullptr;        // 32-bit copy of the up words (one per column of the full matrix)
	uint64_t n1 = 0, n2 = 0;
	int nsm = 0;
	double U0 = 0;
};

int lpp_dblock_create(const ModelDev& m, const HopTable& dn, const DiagTables& dt, cudaStream_t s, DownBlockPlan** out)
{
	*out = nullptr;
	const char* env = getenv("LPP_DBLOCK");
	if (env && env[0] == '0') { g_dberr = "disabled by LPP_DBLOCK=0"; return 1; }
	if (m.model != LPP_MODEL_HUBBARD || !dt.uniformU) { g_dberr = "HubbardOneBand with one U only"; return 1; }
	if (m.nbits > 32 || dn.n >= (1u << 24) || dn.width < 1) { g_dberr = "basis out of range"; return 1; }
	const uint64_t n2 = dn.n;
	const int W = dn.width;
	std::vector<uint32_t> hidx((size_t)W * n2), hcnt(n2);
	std::vector<double> hval((size_t)W * n2), dv2(n2);
	std::vector<word_t> w2(n2), w1(m.n1);
	if (cudaStreamSynchronize(s) != cudaSuccess ||
	    cudaMemcpy(hidx.data(), dn.idx, hidx.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
	    cudaMemcpy(hval.data(), dn.val, hval.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
	    cudaMemcpy(hcnt.data(), dn.cnt, n2 * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
	    cudaMemcpy(dv2.data(), dt.dv2, n2 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
	    cudaMemcpy(w2.data(), m.b2, n2 * sizeof(word_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
	    cudaMemcpy(w1.data(), m.b1, m.n1 * sizeof(word_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
		g_dberr = "table download failed";
		return -1;
	}
	int dev = 0, maxblk = 0, maxsm = 0, nsm = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&maxblk, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
	cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
	cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
	DownBlockPlan* p = new DownBlockPlan();
	std::string err;
	const char* lay = getenv("LPP_DBLOCK_LAYOUT");                  // 1: one CTA of 1024 threads per SM only (A/B timing, tests)
	const char* psv = getenv("LPP_DBLOCK_PASSES");                  // 2 or 3: exactly that many passes (tests)
	const int passes = psv ? atoi(psv) : 0;
	if (!db_build_host_plan(w2.data(), n2, m.nbits, hidx.data(), hval.data(), hcnt.data(), W, dv2.data(), (size_t)maxblk, (size_t)maxsm,
	                        lay ? atoi(lay) : 0, (passes == 2 || passes == 3) ? passes : 0, &p->host, &err)) {
		g_dberr = err;
		delete p;
		return 1;
	}
	// small bases gain nothing: a tile must be worth a CTA
	if (p->host.max_pos < 64) { g_dberr = "blocks too small"; delete p; return 1; }
	if (!db_upload_plan(p->host, &p->dev, &err)) { g_dberr = err; db_free_plan(&p->dev); delete p; return -1; }
	std::vector<uint32_t> w32(m.n1);
	for (uint64_t i = 0; i < m.n1; i++) w32[i] = (uint32_t)w1[i];
	if (cudaMalloc(&p->w1, m.n1 * sizeof(uint32_t)) != cudaSuccess ||
	    cudaMemcpy(p->w1, w32.data(), m.n1 * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
		g_dberr = "up word upload failed";
		cudaFree(p->w1);
		db_free_plan(&p->dev);
		delete p;
		return -1;
	}
	const char* lag = getenv("LPP_DBLOCK_LAG");
	if (lag && atoi(lag) > 0) p->dev.lag = atoi(lag);
	p->n1 = m.n1;
	p->n2 = n2;
	p->nsm = nsm;
	p->U0 = dt.U0;
	*out = p;
	return 0;
}

void lpp_dblock_destroy(DownBlockPlan* p)
{
	if (!p) return;
	cudaFree(p->w1);
	db_free_plan(&p->dev);
	delete p;
}

int lpp_dblock_accepts(const DownBlockPlan* p, const ModelDev& m, const DiagTables& dt, uint64_t d0, uint64_t dcount, const ColView& cv)
{
	return p && d0 == 0 && dcount == p->n2 && m.n2 == p->n2 && dt.uniformU && cv.pitch % 2 == 0 && cv.ncols % 2 == 0 && cv.ncols >= 2 &&
	       cv.u0 + cv.ncols <= p->n1;
}

int lpp_dblock_partials(const DownBlockPlan* p, const ColView& cv)
{
	return (int)(((cv.ncols + DB_COLS - 1) / DB_COLS) * p->dev.pass[p->dev.npass - 1].nblocks);
}

int lpp_dblock_sweep(DownBlockPlan* p, const ModelDev& m, const DiagTables& dt, const SpmvArgs& a, const ColView& cv, cudaStream_t s)
{
	(void)m;
	DbArgs d;
	d.x = a.x;
	d.y = a.y;
	d.pitch = cv.pitch;
	d.ncols = cv.ncols;
	d.alpha = a.alpha.v;
	d.beta = a.beta.v;
	d.alpha_dev = a.alpha.p;
	d.beta_dev = a.beta.p;
	d.U0 = dt.U0;
	d.tmag = p->dev.tmag;
	d.w1 = p->w1 + cv.u0;
	d.dv1 = dt.dv1 + cv.u0;
	d.dot_partials = a.dot_partials;
	const int rc = db_launch(p->dev, d, p->nsm, s);
	if (rc != 0) { g_dberr = std::string("launch failed: ") + cudaGetErrorString(cudaGetLastError()); return -1; }
	return 0;
}

void lpp_dblock_describe(const DownBlockPlan* p, char* buf, size_t n)
{
	const DbHostPlan& h = p->host;
	int o = snprintf(buf, n, "%d passes, %d CTA(s) of %d threads per SM, max %u states, %zu bytes of shared memory, lag %d:", h.npass, h.ctas_per_sm,
	                 h.threads, h.max_pos, h.smem_bytes + (size_t)h.max_pos * 8, p->dev.lag);
	for (int k = 0; k < h.npass && o > 0 && (size_t)o < n; k++)
		o += snprintf(buf + o, n - (size_t)o, " F%d %#x (%u blocks, %.2f hops per state)", k + 1, h.fmask[k], p->dev.pass[k].nblocks, h.pass[k].mean_hops);
}
