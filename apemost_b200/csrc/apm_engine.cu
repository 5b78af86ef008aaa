// apm_engine.cu -- host side of the C ABI in include/apemost_gpu.h: device memory,
// launch plans, the per-step launch sequences and the NCCL hook.  No CPU compute path
// exists here: every likelihood, accept/reject, swap and calibration decision is taken
// by the kernels in apm_kernels.cuh.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/apemost_gpu.h"
#include "apm_kernels.cuh"

using namespace apm;

// ------------------------------------------------------------------ NCCL through dlopen
// (libnccl.so.2 is resolved at run time so the library also loads on boxes without NCCL;
// inside a torch process this binds to the copy torch already loaded)
typedef struct ncclComm * ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };
struct NcclApi {
	void * lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char * (*GetErrorString)(ncclResult_t) = nullptr;
	bool load() {
		if (lib)
			return true;
		lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!lib)
			lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!lib)
			return false;
		GetUniqueId = (decltype(GetUniqueId)) dlsym(lib, "ncclGetUniqueId");
		CommInitRank = (decltype(CommInitRank)) dlsym(lib, "ncclCommInitRank");
		AllReduce = (decltype(AllReduce)) dlsym(lib, "ncclAllReduce");
		CommDestroy = (decltype(CommDestroy)) dlsym(lib, "ncclCommDestroy");
		GetErrorString = (decltype(GetErrorString)) dlsym(lib, "ncclGetErrorString");
		Send = (decltype(Send)) dlsym(lib, "ncclSend");
		Recv = (decltype(Recv)) dlsym(lib, "ncclRecv");
		GroupStart = (decltype(GroupStart)) dlsym(lib, "ncclGroupStart");
		GroupEnd = (decltype(GroupEnd)) dlsym(lib, "ncclGroupEnd");
		return GetUniqueId && CommInitRank && AllReduce && CommDestroy && Send && Recv && GroupStart && GroupEnd;
	}
};
static NcclApi g_nccl;

// ------------------------------------------------------------------ handle
struct apm_gpu {
	apm_gpu_config cfg;
	int n_chains = 0;
	DevState S;
	cudaStream_t stream = nullptr;
	int sm_count = 0;
	int ll_grid = 0;
	// data
	double * d_data = nullptr;
	size_t data_cap = 0, tr_prob_cap = 0, tr_dl_cap = 0, tr_params_cap = 0;
	long long n_rows = 0;
	int n_cols = 0;
	int n_chunks = 0;
	bool have_data = false, have_bounds = false;
	// row-split plan for n_chains slots
	int plan_splits = 0, plan_cps = 0;
	double * d_grid_partials = nullptr, *d_grid_draws = nullptr; // grid path workspaces
	int * d_grid_active = nullptr;
	size_t grid_active_cap = 0;
	size_t grid_partials_cap = 0, grid_draws_cap = 0;
	size_t partial_cap = 0;
	unsigned long long * d_xabsmax = nullptr; // bits of max |x| over the table
	// trace (device)
	double * d_tr_prob = nullptr, *d_tr_dl = nullptr, *d_tr_params = nullptr;
	long long tr_prob_rows = 0, tr_param_rows = 0;
	int tr_dumped = 0;
	// calibration
	unsigned char * d_select = nullptr;
	int * d_sel_idx = nullptr;          // the selected chains, compacted (warp-group calibration kernel)
	std::vector<int> sel_idx;           // host copy, filled by apm_gpu_calibrate / apm_gpu_steps
	// timing / introspection
	long long launches = 0;
	std::vector<cudaEvent_t> ev;
	size_t ev_used = 0;
	double last_ll_ms = 0, last_total_ms = 0;
	long long last_ll_launches = 0;
	int last_path = APM_PATH_TILED;
	// nccl
	ncclComm_t comm = nullptr;
	int rank = 0, n_ranks = 1;
	double * d_shard_sum = nullptr;
	// ladder split: packs of the boundary chains, [n_ens][LADDER_PACK(n_par)] each
	bool ladder = false;
	double * d_pack_first = nullptr, *d_pack_last = nullptr, *d_pack_prev = nullptr, *d_pack_next = nullptr;
	std::vector<unsigned long long> host_draws; // per chain: uniforms handed out by apm_gpu_host_uniform
	// tiled path: one round (n_swap x {likelihood, control} launches) as an instantiated CUDA graph,
	// rebuilt only when something baked into the launches changes (the key = the bytes of the arguments)
	cudaGraphExec_t round_graph = nullptr;
	std::vector<unsigned char> round_graph_key;
	int per_launch_timing = 0;          // 1: per-launch CUDA events (and no graph): apm_gpu_set_timing
	std::string err;
};

static std::string g_create_error;

static int fail(apm_gpu * h, int code, const char * fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (h)
		h->err = buf;
	else
		g_create_error = buf;
	return code;
}

// internal: "this kernel path cannot be scheduled on the device as it is right now" (another context
// holds SMs, MPS partitioning, ...): APM_PATH_AUTO falls back to a path that always works
#define APM_ENOTAPPLICABLE (-100)

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(h, APM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

template<class T>
static cudaError_t dalloc(T ** p, size_t n) {
	cudaError_t e = cudaMalloc((void **) p, std::max<size_t>(n, 1) * sizeof(T));
	if (e == cudaSuccess)
		e = cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T));
	return e;
}

// ------------------------------------------------------------------ model dispatch
#ifdef APM_ONLY_MODEL
template<int ID> struct OnlyModel;
template<> struct OnlyModel<APM_MODEL_SIMPLESIN> { typedef ModelSimplesin type; };
template<> struct OnlyModel<APM_MODEL_SIMPLESIN5> { typedef ModelSimplesin5 type; };
template<> struct OnlyModel<APM_MODEL_SIMPLESIN2> { typedef ModelSimplesin2 type; };
template<> struct OnlyModel<APM_MODEL_NORMAL> { typedef ModelNormal type; };
template<> struct OnlyModel<APM_MODEL_PULSE_VROT> { typedef ModelPulseVrot type; };
template<> struct OnlyModel<APM_MODEL_PULSE> { typedef ModelPulse type; };
template<> struct OnlyModel<APM_MODEL_BERNOULLI> { typedef ModelBernoulli type; };
typedef OnlyModel<APM_ONLY_MODEL>::type APM_ONLY_MODEL_T;
#endif
#ifdef APM_USER_MODEL_HEADER
#define APM_USER_CASE(FN, ...) case APM_MODEL_USER: return FN<UserModel>(__VA_ARGS__);
#else
#define APM_USER_CASE(FN, ...)
#endif
// (-DAPM_ONLY_MODEL=<id>: a build with one model's kernels only -- seconds instead of minutes, for
// kernel experiments under build_variants/; the product library carries all of them)
#ifdef APM_ONLY_MODEL
#define APM_CASE(ID, MODEL, FN, ...) case ID: if (ID == APM_ONLY_MODEL) return FN<std::conditional_t<ID == APM_ONLY_MODEL, MODEL, APM_ONLY_MODEL_T>>(__VA_ARGS__); break;
#else
#define APM_CASE(ID, MODEL, FN, ...) case ID: return FN<MODEL>(__VA_ARGS__);
#endif
#define DISPATCH(model_id, FN, ...) \
	switch (model_id) { \
	APM_CASE(APM_MODEL_SIMPLESIN, ModelSimplesin, FN, __VA_ARGS__) \
	APM_CASE(APM_MODEL_SIMPLESIN5, ModelSimplesin5, FN, __VA_ARGS__) \
	APM_CASE(APM_MODEL_SIMPLESIN2, ModelSimplesin2, FN, __VA_ARGS__) \
	APM_CASE(APM_MODEL_NORMAL, ModelNormal, FN, __VA_ARGS__) \
	APM_CASE(APM_MODEL_PULSE_VROT, ModelPulseVrot, FN, __VA_ARGS__) \
	APM_CASE(APM_MODEL_PULSE, ModelPulse, FN, __VA_ARGS__) \
	APM_CASE(APM_MODEL_BERNOULLI, ModelBernoulli, FN, __VA_ARGS__) \
	APM_USER_CASE(FN, __VA_ARGS__) \
	default: break; } \
	return fail(h, APM_EINVAL, "unknown model id %d (or not in this build)", model_id);

template<class M> static int model_npar_t(apm_gpu *) { return M::NPAR; }
template<class M> static int model_ncols_t(apm_gpu *) { return M::HAS_DATA ? M::NCOLS : 0; } // -1: one per parameter
template<class M> static int model_row_w_t(apm_gpu *) { return M::ROW_W; }
template<class M> static int model_chunk_rows_t(apm_gpu *) { return ll_chunk<M>(); }
static int model_npar(apm_gpu * h, int id) { DISPATCH(id, model_npar_t, h) }
static int model_ncols(apm_gpu * h, int id) { DISPATCH(id, model_ncols_t, h) }
static int model_row_w(apm_gpu * h, int id) { DISPATCH(id, model_row_w_t, h) }
static int model_chunk_rows(apm_gpu * h, int id) { DISPATCH(id, model_chunk_rows_t, h) }

extern "C" int apm_gpu_model_n_par(int model_id) {
	int r = model_npar(nullptr, model_id);
	return r < 0 ? 0 : r;
}
extern "C" int apm_gpu_model_n_cols(int model_id) {
	if (model_id == APM_MODEL_BERNOULLI)
		return -1; // one column per parameter
	int r = model_ncols(nullptr, model_id);
	return r < 0 ? 0 : r;
}
extern "C" int apm_gpu_abi_version(void) { return APM_GPU_ABI_VERSION; }

extern "C" const char * apm_gpu_last_error(const apm_gpu * h) {
	return h ? h->err.c_str() : g_create_error.c_str();
}

// ------------------------------------------------------------------ lifecycle
template<class M>
static int configure_kernels(apm_gpu * h) {
	CU(cudaFuncSetAttribute(loglik_tiled_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			(int) LL_SMEM_BYTES));
	int occ = 0;
	CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, loglik_tiled_kernel<M>, LL_LAUNCH_THREADS,
			LL_SMEM_BYTES));
	if (occ < 1)
		return fail(h, APM_ECUDA, "likelihood kernel does not fit on an SM");
	h->ll_grid = h->sm_count * occ;
	const int fused_smem = (int) FUSED_SMEM_LIMIT;
	if constexpr (M::HAS_DATA)
		CU(cudaFuncSetAttribute(fused_run_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem));
	else
		CU(cudaFuncSetAttribute(free_run_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem));
	CU(cudaFuncSetAttribute(fused_calibrate_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem));
	CU(cudaFuncSetAttribute(group_calibrate_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem));
	CU(cudaFuncSetAttribute(cluster_run_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem));
	CU(cudaFuncSetAttribute(grid_run_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem));
	CU(cudaFuncSetAttribute(grid_calibrate_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem));
	return APM_OK;
}

extern "C" int apm_gpu_create(apm_gpu ** out, const apm_gpu_config * cfg) {
	apm_gpu * h = nullptr;
	struct timespec t_entry;
	clock_gettime(CLOCK_MONOTONIC, &t_entry);
	if (!out || !cfg)
		return fail(h, APM_EINVAL, "null argument");
	*out = nullptr;
	if (cfg->n_ensembles < 1 || cfg->n_beta < 1 || cfg->n_par < 1 || cfg->n_par > APM_MAX_PAR)
		return fail(h, APM_EINVAL, "need n_ensembles >= 1, n_beta >= 1, 1 <= n_par <= %d", APM_MAX_PAR);
	if ((long long) cfg->n_ensembles * cfg->n_beta > (1ll << 30))
		return fail(h, APM_EINVAL, "too many chains");
	int want = model_npar(nullptr, cfg->model_id);
	if (want < 0)
		return want;
	if (want > 0 && want != cfg->n_par)
		return fail(h, APM_EINVAL, "model %d has %d parameters, config says %d", cfg->model_id, want,
				cfg->n_par);
	if (cfg->model_id == APM_MODEL_PULSE && (cfg->n_par < 4 || (cfg->n_par - 2) % 2 != 0))
		return fail(h, APM_EINVAL, "pulse needs n_par = 2 + 2k");
	if (cfg->model_id == APM_MODEL_BERNOULLI && (cfg->n_par < 2 || cfg->n_par > 4))
		return fail(h, APM_EINVAL, "bernoulli: n_par = number of data columns, 2 to 4 on the device");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		return fail(h, APM_ENODEVICE, "no CUDA device available (this engine has no CPU fallback)");
	}
	if (cfg->device < 0 || cfg->device >= ndev)
		return fail(h, APM_ENODEVICE, "device %d out of range (%d devices)", cfg->device, ndev);
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess)
		return fail(h, APM_ECUDA, "cudaGetDeviceProperties failed");
	if (prop.major != 10)
		return fail(h, APM_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
				cfg->device, prop.major, prop.minor);
	h = new apm_gpu();
	h->cfg = *cfg;
	h->n_chains = cfg->n_ensembles * cfg->n_beta;
	h->sm_count = prop.multiProcessorCount;
	memset(&h->S, 0, sizeof(h->S));
	int rc = APM_OK;
	// APM_HOST_TIMING=1: where the start-up time goes (stderr)
	const bool timing = getenv("APM_HOST_TIMING") != nullptr;
	auto now = []() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; };
	double t_mark = t_entry.tv_sec + 1e-9 * t_entry.tv_nsec;
	auto mark = [&](const char * what) {
		if (timing) {
			const double t = now();
			fprintf(stderr, "[timing]   engine: %-26s %8.3f s\n", what, t - t_mark);
			t_mark = t;
		}
	};
	mark("driver + device query");
	do {
		if (cudaSetDevice(cfg->device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) {
			rc = fail(nullptr, APM_ECUDA, "cannot initialise device %d: %s", cfg->device,
					cudaGetErrorString(cudaGetLastError()));
			break;
		}
		mark("CUDA context");
		if (cudaStreamCreate(&h->stream) != cudaSuccess) {
			rc = fail(nullptr, APM_ECUDA, "cannot initialise device %d: %s", cfg->device,
					cudaGetErrorString(cudaGetLastError()));
			break;
		}
		DevState & S = h->S;
		const size_t n = h->n_chains, np = cfg->n_par, nv = n * np;
		S.n_chains = h->n_chains;
		S.n_beta = cfg->n_beta;
		S.n_ens = cfg->n_ensembles;
		S.n_par = cfg->n_par;
		S.seed = cfg->seed;
		S.proposal = cfg->proposal;
		S.circular_mask = cfg->circular_mask;
		S.quirks = cfg->quirks;
		S.chain_id_offset = cfg->chain_id_offset;
		S.ensemble_id_offset = cfg->ensemble_id_offset;
		S.n_beta_total = cfg->n_beta;
		S.id_stride = cfg->n_beta;
		S.k_offset = 0;
		for (int i = 0; i < 4; i++)
			S.model_const[i] = cfg->model_const[i];
		cudaError_t e = cudaSuccess;
#define A(ptr, count) if (e == cudaSuccess) e = dalloc(&ptr, count)
		A(S.params, nv); A(S.params_best, nv); A(S.steps, nv); A(S.prop, nv);
		A(S.prob, n); A(S.prior, n); A(S.prob_best, n); A(S.beta, n);
		A(S.accept, n); A(S.reject, n); A(S.pacc, nv); A(S.prej, nv); A(S.n_iter, n);
		A(S.swapcount, n); A(S.rng_ctr, n); A(S.swap_round, (size_t) cfg->n_ensembles);
		A(S.pmin, np); A(S.pmax, np); A(S.pend, n);
		A(S.stat_n, n); A(S.stat_sum_dl, n); A(S.stat_sum_p, nv); A(S.stat_sum_p2, nv);
		A(S.cal, n); A(S.progress_n, 1); A(S.n_active, 1);
		A(h->d_select, n); A(h->d_sel_idx, n); A(h->d_shard_sum, n); A(h->d_xabsmax, 1);
		A(S.act_idx, 2 * n); A(S.act_n, 2); A(S.run_ctr, 2);
#undef A
		if (e != cudaSuccess) {
			rc = fail(nullptr, APM_ENOMEM, "device allocation failed: %s", cudaGetErrorString(e));
			break;
		}
		// mcmc_init defaults (reference src/mcmc.c:37-78): prob = prob_best = -1e10, beta = 1
		std::vector<double> init(n, -1e10), ones(n, 1.0);
		std::vector<int> none(n, PEND_NONE);
		cudaMemcpy(S.prob, init.data(), n * sizeof(double), cudaMemcpyHostToDevice);
		cudaMemcpy(S.prob_best, init.data(), n * sizeof(double), cudaMemcpyHostToDevice);
		cudaMemcpy(S.beta, ones.data(), n * sizeof(double), cudaMemcpyHostToDevice);
		cudaMemcpy(S.pend, none.data(), n * sizeof(int), cudaMemcpyHostToDevice);
		S.n_splits = 1;
		mark("device state");
	} while (0);
	if (rc == APM_OK) {
		auto conf = [&]() -> int { DISPATCH(cfg->model_id, configure_kernels, h) };
		rc = conf();
		if (rc != APM_OK)
			g_create_error = h->err;
		mark("kernel module + attributes");
	}
	if (rc != APM_OK) {
		apm_gpu_destroy(h);
		return rc;
	}
	*out = h;
	return APM_OK;
}

extern "C" int apm_gpu_destroy(apm_gpu * h) {
	if (!h)
		return APM_OK;
	cudaSetDevice(h->cfg.device);
	if (h->stream)
		cudaStreamSynchronize(h->stream);
	DevState & S = h->S;
	void * ptrs[] = { S.params, S.params_best, S.steps, S.prop, S.prob, S.prior, S.prob_best, S.beta,
			S.accept, S.reject, S.pacc, S.prej, S.n_iter, S.swapcount, S.rng_ctr, S.swap_round, S.pmin,
			S.pmax, S.pend, S.partial, S.stat_n, S.stat_sum_dl, S.stat_sum_p, S.stat_sum_p2, S.cal,
			S.progress, S.progress_n, S.n_active, S.act_idx, S.act_n, S.run_ctr, h->d_select, h->d_sel_idx, h->d_shard_sum,
			h->d_pack_first, h->d_pack_last, h->d_pack_prev, h->d_pack_next, h->d_grid_partials, h->d_grid_draws, h->d_grid_active,
			h->d_xabsmax, h->d_data, h->d_tr_prob, h->d_tr_dl, h->d_tr_params,
			S.marg_counts, S.marg_bsum, S.marg_means, S.marg_n, S.marg_nb };
	for (void * p : ptrs)
		if (p)
			cudaFree(p);
	for (cudaEvent_t e : h->ev)
		cudaEventDestroy(e);
	if (h->round_graph)
		cudaGraphExecDestroy(h->round_graph);
	if (h->comm && g_nccl.CommDestroy)
		g_nccl.CommDestroy(h->comm);
	if (h->stream)
		cudaStreamDestroy(h->stream);
	delete h;
	return APM_OK;
}

// ------------------------------------------------------------------ launch plan
// How many row splits for n_slots chains in tiles of `tile`.  Work items = chain tiles x splits
// are dealt round-robin to the persistent grid, split-major, and a split has n_chunks / n_splits
// chunks (the first n_chunks % n_splits one more): a CTA's items come from all over the table, so
// its load is (its number of items) x (the mean split) but for the last, partly filled round,
// whose items are from the last -- the smaller -- splits.  Pick the count with the shortest
// makespan, charging a small per-item overhead (parameter load + fold) so that items do not get
// needlessly small.
static void make_plan(const apm_gpu * h, int n_slots, int tile, int & n_splits, int & cps) {
	const int n_chunks = std::max(h->n_chunks, 1);
	const long long grid = std::max(h->ll_grid, 1);
	const long long n_ctiles = std::max((n_slots + tile - 1) / tile, 1);
	if (const char * t = getenv("APM_SPLITS")) { // kernel-sweep override
		n_splits = std::max(1, std::min(atoi(t), n_chunks));
		cps = (n_chunks + n_splits - 1) / n_splits;
		return;
	}
	const double item_overhead = 0.15; // in units of one chunk's compute time
	double best = 1e300;
	n_splits = 1;
	cps = n_chunks;
	for (int s = 1; s <= std::min(n_chunks, 512); s++) {
		const int base = n_chunks / s;
		const double mean = (double) n_chunks / s;
		const long long items = n_ctiles * s;
		const long long full = items / grid, rest = items % grid;
		const double makespan = full * (mean + item_overhead) + (rest ? base + item_overhead : 0.0);
		if (makespan < best * (1 - 1e-9)) {
			best = makespan;
			n_splits = s;
			cps = (n_chunks + s - 1) / s;
		}
	}
}

// grow-only device buffers: cudaMalloc/cudaFree cost up to tens of ms, so repeated set_data /
// run calls reuse what is there and only the padding or the contents are rewritten
template<class T>
static cudaError_t ensure_cap(T ** p, size_t * cap, size_t count) {
	if (count <= *cap && *p != nullptr)
		return cudaSuccess;
	if (*p)
		cudaFree(*p);
	*p = nullptr;
	*cap = 0;
	cudaError_t e = cudaMalloc((void **) p, std::max<size_t>(count, 1) * sizeof(T));
	if (e == cudaSuccess)
		*cap = std::max<size_t>(count, 1);
	return e;
}

static int ensure_partial(apm_gpu * h, size_t count) {
	if (count <= h->partial_cap)
		return APM_OK;
	if (h->S.partial)
		cudaFree(h->S.partial);
	h->S.partial = nullptr;
	h->partial_cap = 0;
	CU(dalloc(&h->S.partial, count));
	h->partial_cap = count;
	return APM_OK;
}

// (re)plan the row splits for n_slots pending chains; the partial-sum buffer is laid out
// [n_chains][n_splits] whatever subset is evaluated
template<class M> static int model_tile_t(apm_gpu *) { return M::LL_C; }
static int model_tile(apm_gpu * h) { DISPATCH(h->cfg.model_id, model_tile_t, h) }

static int replan(apm_gpu * h, int n_slots) {
	const int tile = model_tile(h);
	if (tile < 0)
		return tile;
	make_plan(h, std::max(n_slots, 1), tile, h->plan_splits, h->plan_cps);
	// the likelihood kernel leaves LL_PARTS partial sums per (chain, row split)
	int rc = ensure_partial(h, (size_t) h->n_chains * h->plan_splits * LL_PARTS);
	if (rc != APM_OK)
		return rc;
	h->S.n_splits = h->plan_splits * LL_PARTS;
	return APM_OK;
}

// ------------------------------------------------------------------ inputs
extern "C" int apm_gpu_set_data(apm_gpu * h, const double * rowmajor, long long n_rows, int n_cols) {
	if (!h)
		return APM_EINVAL;
	CU(cudaSetDevice(h->cfg.device));
	int need = model_ncols(h, h->cfg.model_id);
	if (need < -1)
		return need;
	if (need == -1)
		need = h->cfg.n_par; // one column per parameter (bernoulli)
	const int row_w = model_row_w(h, h->cfg.model_id);       // doubles per row on the device
	const int chunk_rows = model_chunk_rows(h, h->cfg.model_id);
	if (row_w < 0 || chunk_rows < 0)
		return row_w < 0 ? row_w : chunk_rows;
	if (need > 0 && (rowmajor == nullptr || n_rows < 1))
		return fail(h, APM_EINVAL, "this model needs a data table");
	if (need > 0 && n_cols < need)
		return fail(h, APM_EINVAL, "model reads %d data columns, table has %d", need, n_cols);
	if (need > row_w)
		return fail(h, APM_EINVAL, "models reading more than %d columns are not supported", row_w);
	// (nothing of the old table survives a failed call: have_data is set again at the very end)
	h->have_data = false;
	h->n_rows = n_rows;
	h->n_cols = n_cols;
	h->n_chunks = 0;
	if (need > 0) {
		// device layout: [rows padded to a whole number of chunks][row_w]: for two columns the
		// gsl_matrix row-major layout (tda = 2); wider tables are narrowed to the columns the model
		// reads, narrower rows (3 columns in a 4-wide row) are padded with zeros
		h->n_chunks = (int) ((n_rows + chunk_rows - 1) / chunk_rows);
		const size_t padded = (size_t) h->n_chunks * chunk_rows;
		const int take = std::min(n_cols, row_w);
		CU(ensure_cap(&h->d_data, &h->data_cap, padded * row_w));
		if (take < row_w) // zero everything: the padding columns and the padding rows
			CU(cudaMemsetAsync(h->d_data, 0, padded * row_w * sizeof(double), h->stream));
		else if (padded > (size_t) n_rows) // zero the padding rows of the last chunk
			CU(cudaMemsetAsync(h->d_data + (size_t) n_rows * row_w, 0,
					(padded - (size_t) n_rows) * row_w * sizeof(double), h->stream));
		if (n_cols == row_w) {
			CU(cudaMemcpyAsync(h->d_data, rowmajor, (size_t) n_rows * row_w * sizeof(double),
					cudaMemcpyHostToDevice, h->stream));
		} else {
			CU(cudaMemcpy2DAsync(h->d_data, row_w * sizeof(double), rowmajor, (size_t) n_cols * sizeof(double),
					take * sizeof(double), (size_t) n_rows, cudaMemcpyHostToDevice, h->stream));
		}
		CU(cudaMemsetAsync(h->d_xabsmax, 0, sizeof(unsigned long long), h->stream));
		absmax_col0_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(h->d_data, n_rows, row_w, h->d_xabsmax);
		h->launches++;
		CU(cudaStreamSynchronize(h->stream));
	}
	int rc = replan(h, h->n_chains);
	if (rc != APM_OK)
		return rc;
	h->have_data = true;
	return APM_OK;
}

extern "C" int apm_gpu_set_bounds(apm_gpu * h, const double * pmin, const double * pmax) {
	if (!h || !pmin || !pmax)
		return APM_EINVAL;
	CU(cudaSetDevice(h->cfg.device));
	for (int i = 0; i < h->cfg.n_par; i++)
		if (!(pmin[i] <= pmax[i]))
			return fail(h, APM_EINVAL, "min(%f) > max(%f) for parameter %d", pmin[i], pmax[i], i);
	CU(cudaMemcpy(h->S.pmin, pmin, h->cfg.n_par * sizeof(double), cudaMemcpyHostToDevice));
	CU(cudaMemcpy(h->S.pmax, pmax, h->cfg.n_par * sizeof(double), cudaMemcpyHostToDevice));
	h->have_bounds = true;
	return APM_OK;
}

#define CHAIN_FIELDS(X) \
	X(beta, double, 1, S.beta) X(params, double, np, S.params) X(steps, double, np, S.steps) \
	X(prob, double, 1, S.prob) X(prior, double, 1, S.prior) X(prob_best, double, 1, S.prob_best) \
	X(params_best, double, np, S.params_best) X(accept, u64, 1, S.accept) X(reject, u64, 1, S.reject) \
	X(params_accepts, u64, np, S.pacc) X(params_rejects, u64, np, S.prej) X(n_iter, u64, 1, S.n_iter) \
	X(swapcount, u64, 1, S.swapcount) X(rng_counter, u64, 1, S.rng_ctr)

extern "C" int apm_gpu_set_chains(apm_gpu * h, int first, int count, const apm_gpu_chain_io * in) {
	if (!h || !in)
		return APM_EINVAL;
	if (first < 0 || count < 0 || first + count > h->n_chains)
		return fail(h, APM_EINVAL, "chain range [%d, %d) outside [0, %d)", first, first + count,
				h->n_chains);
	CU(cudaSetDevice(h->cfg.device));
	DevState & S = h->S;
	const size_t np = h->cfg.n_par;
	if (in->beta && !in->swapcount) // set_beta zeroes swapcount (src/parallel_tempering_beta.c:25-28)
		CU(cudaMemsetAsync(S.swapcount + first, 0, count * sizeof(u64), h->stream));
#define X(name, type, width, dev) \
	if (in->name) CU(cudaMemcpyAsync(dev + (size_t) first * (width), in->name, \
			(size_t) count * (width) * sizeof(type), cudaMemcpyHostToDevice, h->stream));
	CHAIN_FIELDS(X)
#undef X
	CU(cudaStreamSynchronize(h->stream));
	return APM_OK;
}

extern "C" int apm_gpu_get_chains(apm_gpu * h, int first, int count, apm_gpu_chain_io * out) {
	if (!h || !out)
		return APM_EINVAL;
	if (first < 0 || count < 0 || first + count > h->n_chains)
		return fail(h, APM_EINVAL, "chain range [%d, %d) outside [0, %d)", first, first + count,
				h->n_chains);
	CU(cudaSetDevice(h->cfg.device));
	DevState & S = h->S;
	const size_t np = h->cfg.n_par;
#define X(name, type, width, dev) \
	if (out->name) CU(cudaMemcpyAsync(out->name, dev + (size_t) first * (width), \
			(size_t) count * (width) * sizeof(type), cudaMemcpyDeviceToHost, h->stream));
	CHAIN_FIELDS(X)
#undef X
	CU(cudaStreamSynchronize(h->stream));
	return APM_OK;
}

// ------------------------------------------------------------------ launches
static cudaEvent_t next_event(apm_gpu * h) {
	if (h->ev_used == h->ev.size()) {
		cudaEvent_t e;
		cudaEventCreate(&e);
		h->ev.push_back(e);
	}
	return h->ev[h->ev_used++];
}

constexpr size_t MAX_TIMED_LAUNCHES = 8192;

template<class M>
static int launch_loglik(apm_gpu * h, const double * prop, const int * act_idx, const int * act_n,
		int n_slots, int n_upper, int n_splits, int cps, double * partial, bool timed) {
	// n_slots = rows of prop / partial; n_upper = upper bound of the number of slots evaluated
	// (= n_slots without an active list) -- it only sizes the grid
	if (!M::HAS_DATA)
		return APM_OK;
	LLArgs a;
	a.data = h->d_data;
	a.n_rows = h->n_rows;
	a.prop = prop;
	a.act_idx = act_idx;
	a.act_n = act_n;
	a.n_slots = n_slots;
	a.n_par = h->cfg.n_par;
	a.n_splits = n_splits;
	a.chunks_per_split = cps;
	a.n_chunks = h->n_chunks;
	a.xabsmax = reinterpret_cast<const double *>(h->d_xabsmax);
	a.partial = partial;
	for (int i = 0; i < 4; i++)
		a.model_const[i] = h->cfg.model_const[i];
	const long long n_items = (long long) ((n_upper + M::LL_C - 1) / M::LL_C) * n_splits;
	if ((long long) ((n_slots + M::LL_C - 1) / M::LL_C) * n_splits >= (1ll << 31))
		return fail(h, APM_EINVAL, "too many work items");
	const int grid = (int) std::max<long long>(1, std::min<long long>(h->ll_grid, n_items));
	timed = timed && h->ev_used + 2 <= 2 * MAX_TIMED_LAUNCHES;
	if (timed)
		cudaEventRecord(next_event(h), h->stream);
	loglik_tiled_kernel<M><<<grid, LL_LAUNCH_THREADS, LL_SMEM_BYTES, h->stream>>>(a);
	if (timed)
		cudaEventRecord(next_event(h), h->stream);
	h->launches++;
	return APM_OK;
}

// ---- data-sharded support: fold row splits per chain, all-reduce, finalise with 1 split
__global__ void fold_splits_kernel(const double * partial, int n_splits, int n, double * out) {
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n)
		return;
	double s = 0;
	for (int k = 0; k < n_splits; k++)
		s += partial[(size_t) g * n_splits + k];
	out[g] = s;
}

template<class M>
static int step_likelihood(apm_gpu * h, bool timed, int act_w = -1, int n_upper = -1) {
	// likelihood of every pending proposal -> S.partial (or, data-sharded, the all-reduced
	// per-chain sums in d_shard_sum).  act_w >= 0: only the chains of active list act_w.
	const int * idx = act_w >= 0 ? h->S.act_idx + (size_t) act_w * h->n_chains : nullptr;
	const int * cnt = act_w >= 0 ? h->S.act_n + act_w : nullptr;
	int rc = launch_loglik<M>(h, h->S.prop, idx, cnt, h->n_chains, n_upper >= 0 ? n_upper : h->n_chains,
			h->plan_splits, h->plan_cps, h->S.partial, timed);
	if (rc != APM_OK)
		return rc;
	if (h->comm && !h->ladder && M::HAS_DATA) {
		fold_splits_kernel<<<(h->n_chains + 255) / 256, 256, 0, h->stream>>>(h->S.partial, h->plan_splits * LL_PARTS,
				h->n_chains, h->d_shard_sum);
		h->launches++;
		ncclResult_t r = g_nccl.AllReduce(h->d_shard_sum, h->d_shard_sum, (size_t) h->n_chains,
				ncclFloat64, ncclSum, h->comm, h->stream);
		if (r != 0)
			return fail(h, APM_ENCCL, "ncclAllReduce failed: %s",
					g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
	}
	return APM_OK;
}

static DevState state_for_advance(const apm_gpu * h) {
	DevState S = h->S;
	if (h->comm && !h->ladder) {
		S.partial = h->d_shard_sum;
		S.n_splits = 1;
	}
	return S;
}

// ---- ladder split: every rank sends its first rung's pack down and its last rung's pack up
static int ladder_exchange(apm_gpu * h) {
	const size_t count = (size_t) h->cfg.n_ensembles * LADDER_PACK(h->cfg.n_par);
	ladder_pack_kernel<<<(h->cfg.n_ensembles + 127) / 128, 128, 0, h->stream>>>(h->S, h->d_pack_first, h->d_pack_last);
	h->launches++;
	ncclResult_t r = g_nccl.GroupStart();
	if (r == 0 && h->rank + 1 < h->n_ranks) {
		r = g_nccl.Send(h->d_pack_last, count, ncclFloat64, h->rank + 1, h->comm, h->stream);
		if (r == 0)
			r = g_nccl.Recv(h->d_pack_next, count, ncclFloat64, h->rank + 1, h->comm, h->stream);
	}
	if (r == 0 && h->rank > 0) {
		r = g_nccl.Send(h->d_pack_first, count, ncclFloat64, h->rank - 1, h->comm, h->stream);
		if (r == 0)
			r = g_nccl.Recv(h->d_pack_prev, count, ncclFloat64, h->rank - 1, h->comm, h->stream);
	}
	ncclResult_t e = g_nccl.GroupEnd();
	if (r == 0)
		r = e;
	if (r != 0)
		return fail(h, APM_ENCCL, "ladder exchange failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
	return APM_OK;
}

// ------------------------------------------------------------------ eval
template<class M>
static int eval_t(apm_gpu * h, int n, const double * params, const double * beta, double * prob_out,
		double * prior_out) {
	const int np = h->cfg.n_par;
	double * d_params = nullptr, *d_beta = nullptr, *d_prob = nullptr, *d_prior = nullptr, *d_partial = nullptr;
	int n_splits = 1, cps = 1;
	make_plan(h, n, M::LL_C, n_splits, cps);
	int rc = APM_OK;
	cudaError_t e = dalloc(&d_params, (size_t) n * np);
	if (e == cudaSuccess) e = dalloc(&d_beta, (size_t) n);
	if (e == cudaSuccess) e = dalloc(&d_prob, (size_t) n);
	if (e == cudaSuccess) e = dalloc(&d_prior, (size_t) n);
	if (e == cudaSuccess) e = dalloc(&d_partial, (size_t) n * n_splits * LL_PARTS);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_params, params, (size_t) n * np * sizeof(double),
			cudaMemcpyHostToDevice, h->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_beta, beta, (size_t) n * sizeof(double),
			cudaMemcpyHostToDevice, h->stream);
	if (e == cudaSuccess) {
		h->ev_used = 0;
		cudaEvent_t t0 = next_event(h), t1 = next_event(h);
		cudaEventRecord(t0, h->stream);
		rc = launch_loglik<M>(h, d_params, nullptr, nullptr, n, n, n_splits, cps, d_partial, true);
		const double * mc = h->cfg.model_const;
		eval_finish_kernel<M><<<(n + 127) / 128, 128, 0, h->stream>>>(n, np, d_params, d_beta, d_partial,
				n_splits * LL_PARTS, d_prob, d_prior, mc[0], mc[1], mc[2], mc[3]);
		h->launches++;
		cudaEventRecord(t1, h->stream);
		e = cudaMemcpyAsync(prob_out, d_prob, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
		if (e == cudaSuccess && prior_out)
			e = cudaMemcpyAsync(prior_out, d_prior, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost,
					h->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
		if (e == cudaSuccess) {
			// t0/t1 occupy ev[0], ev[1]; the loglik pair (if any) is ev[2], ev[3]
			float tot = 0;
			cudaEventElapsedTime(&tot, t0, t1);
			h->last_total_ms = tot;
			h->last_ll_ms = 0;
			h->last_ll_launches = 0;
			if (h->ev_used >= 4) {
				float ms = 0;
				cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
				h->last_ll_ms = ms;
				h->last_ll_launches = 1;
			}
		}
	}
	cudaFree(d_params); cudaFree(d_beta); cudaFree(d_prob); cudaFree(d_prior); cudaFree(d_partial);
	if (e != cudaSuccess)
		return fail(h, APM_ECUDA, "eval failed: %s", cudaGetErrorString(e));
	h->last_path = APM_PATH_TILED;
	return rc;
}

extern "C" int apm_gpu_eval(apm_gpu * h, int n, const double * params, const double * beta,
		double * prob_out, double * prior_out) {
	if (!h || n < 0 || !params || !beta || !prob_out)
		return APM_EINVAL;
	if (n == 0)
		return APM_OK;
	if (!h->have_data)
		return fail(h, APM_ESTATE, "apm_gpu_set_data must be called first");
	if (h->comm && !h->ladder)
		return fail(h, APM_ESTATE, "apm_gpu_eval is not available in data-sharded mode");
	CU(cudaSetDevice(h->cfg.device));
	DISPATCH(h->cfg.model_id, eval_t, h, n, params, beta, prob_out, prior_out)
}

// ------------------------------------------------------------------ path selection
// fused (one launch per run / calibration, table in shared memory) when the table fits and the
// likelihood is not sharded over GPUs; tiled otherwise.  cfg.path can force either.
template<class M> static int model_has_data_t(apm_gpu *) { return M::HAS_DATA ? 1 : 0; }
static int model_has_data(apm_gpu * h) { DISPATCH(h->cfg.model_id, model_has_data_t, h) }

// cluster geometry for a run: CTAs per ensemble (0 = the cluster path does not apply)
static int cluster_size_for(const apm_gpu * h, bool has_data) {
	if (!has_data || h->comm || h->cfg.n_beta < 2)
		return 0;
	int want = CLUSTER_MAX;
	if (const char * t = getenv("APM_CLUSTER")) // experiments: force a cluster size (0 = never)
		want = atoi(t);
	for (int cl = CLUSTER_MAX; cl >= 2; cl >>= 1) {
		if (cl > want || cl > h->cfg.n_beta)
			continue;
		if ((long long) h->cfg.n_ensembles * cl > h->sm_count)
			continue; // more CTAs than SMs: the fused kernel's one CTA per ensemble is the better deal
		const int gmax = (h->cfg.n_beta + cl - 1) / cl;
		if (gmax > FUSED_MAX_WARPS)
			continue;
		const int wc = gmax <= 8 ? FUSED_MAX_WARPS / gmax : 1; // named barriers 1..8 for the groups
		const long long n_slots = h->n_rows * (model_row_w(const_cast<apm_gpu *>(h), h->cfg.model_id) / 2);
		if (cluster_smem_bytes(n_slots, gmax, wc, h->cfg.n_par) > FUSED_SMEM_LIMIT)
			continue;
		return cl;
	}
	return 0;
}

// grid path: rows per CTA (0 = the path does not apply)
static int grid_slice_rows(const apm_gpu * h, bool has_data) {
	if (!has_data || h->comm || h->n_chains > GRID_MAX_CHAINS_PER_SM * h->sm_count || h->n_chains > GRID_MAX_CHAINS
			|| h->n_rows < 1)
		return 0;
	if (getenv("APM_NO_GRID"))
		return 0;
	const int row_w = model_row_w(const_cast<apm_gpu *>(h), h->cfg.model_id);
	const long long slice = (h->n_rows + h->sm_count - 1) / h->sm_count + 1;
	const size_t need = (((size_t) slice * row_w * 8 + 127) & ~(size_t) 127) + 64 + grid_state_bytes(h->n_chains, h->cfg.n_par);
	if (need > FUSED_SMEM_LIMIT)
		return 0;
	return (int) slice;
}

static int choose_path(apm_gpu * h, int * path, bool for_run = false) {
	const int has_data = model_has_data(h);
	if (has_data < 0)
		return has_data;
	const long long n_slots = has_data ? h->n_rows * (model_row_w(h, h->cfg.model_id) / 2) : 0; // 16-byte units
	const size_t need = fused_table_bytes(n_slots) + fused_state_bytes(h->cfg.n_beta, h->cfg.n_par)
			+ fused_draws_bytes(h->cfg.n_beta, has_data != 0);
	const bool fits = need <= FUSED_SMEM_LIMIT && (!has_data || h->n_rows < (1ll << 24));
	const bool rows_ok = !has_data || h->n_rows < (1ll << 24);
	const int cl = for_run && rows_ok ? cluster_size_for(h, has_data != 0) : 0;
	const int gslice = grid_slice_rows(h, has_data != 0);
	int want = h->cfg.path;
	if (want == APM_PATH_CLUSTER && !for_run) // calibration: chains do not interact, no cluster needed
		want = fits ? APM_PATH_FUSED : APM_PATH_AUTO;
	if (want == APM_PATH_GRID && !for_run && gslice == 0) // e.g. a data-free model
		want = APM_PATH_AUTO;
	if (want == APM_PATH_GRID) {
		if (gslice == 0)
			return fail(h, APM_EINVAL, "the grid path needs a data model, at most %d chains, a slice of the table "
					"(rows / %d SMs) plus the chains' proposals in shared memory and no multi-GPU sharding",
					GRID_MAX_CHAINS_PER_SM * h->sm_count, h->sm_count);
		*path = APM_PATH_GRID;
	} else if (want == APM_PATH_CLUSTER) {
		if (cl == 0)
			return fail(h, APM_EINVAL, "the cluster path needs a data model, 2 <= cluster size <= n_beta, "
					"n_ensembles x cluster size <= %d SMs, the table in shared memory and no multi-GPU sharding",
					h->sm_count);
		*path = APM_PATH_CLUSTER;
	} else if (want == APM_PATH_FUSED) {
		if (h->comm)
			return fail(h, APM_EINVAL, "the fused path cannot be used with a data-sharded likelihood");
		if (!fits)
			return fail(h, APM_EINVAL, "the fused path needs the table and one ensemble's state in shared memory: "
					"%lld rows x %d chains need %lld bytes > %lld", h->n_rows, h->cfg.n_beta, (long long) need,
					(long long) FUSED_SMEM_LIMIT);
		*path = APM_PATH_FUSED;
	} else if (want == APM_PATH_TILED) {
		*path = APM_PATH_TILED;
	} else {
		// a tiled step costs ~27 us of launch gaps and fixed latencies, a grid step two grid barriers
		// (~7 us); measured (tools/mid_bench.py) the grid path is ahead up to ~5e8 row evaluations
		// per step, where both are bound by the row evaluations themselves
		const bool grid_pays = gslice > 0 && (double) h->n_rows * h->n_chains <= 5e8;
		*path = cl > 0 ? APM_PATH_CLUSTER
				: ((fits && !h->comm) ? APM_PATH_FUSED : (grid_pays ? APM_PATH_GRID : APM_PATH_TILED));
	}
	return APM_OK;
}

// warp-group calibration (group_calibrate_kernel): chains a CTA works on at a time for n_sel selected
// chains, 0 = the kernel does not apply (data-free model, sharded likelihood, table + states do not
// fit in shared memory, or so many chains that a group would be a single warp: the fused kernel's case)
static int group_calibrate_ng(const apm_gpu * h, bool has_data, int n_sel) {
	if (!has_data || h->comm || n_sel < 1 || h->n_rows >= (1ll << 24))
		return 0;
	if (getenv("APM_NO_GROUP_CALIBRATE"))
		return 0;
	const int per_cta = (n_sel + h->sm_count - 1) / h->sm_count;
	if (per_cta > GROUP_MAX)
		return 0;
	int ng = 1;
	while (ng < per_cta)
		ng *= 2;
	const long long n_slots = h->n_rows * (model_row_w(const_cast<apm_gpu *>(h), h->cfg.model_id) / 2);
	if (fused_table_bytes(n_slots) + (size_t) ng * group_state_bytes(h->cfg.n_par) > FUSED_SMEM_LIMIT)
		return 0;
	return ng;
}

static void fused_geometry(const apm_gpu * h, bool has_data, int * threads, size_t * smem, FusedArgs * a) {
	if (has_data) { // a warp per chain, the ladder dealt evenly over as few passes as possible
		const int passes = (h->cfg.n_beta + FUSED_MAX_WARPS - 1) / FUSED_MAX_WARPS;
		*threads = 32 * ((h->cfg.n_beta + passes - 1) / passes);
	} else {        // free_run_kernel: groups of lanes per chain + warps that draw ahead; calibration: a thread per chain
		*threads = std::min(FUSED_MAX_WARPS * 32, 32 * ((h->cfg.n_beta + 31) / 32));
	}
	*smem = fused_table_bytes(has_data ? h->n_rows * (model_row_w(const_cast<apm_gpu *>(h), h->cfg.model_id) / 2) : 0)
			+ fused_state_bytes(h->cfg.n_beta, h->cfg.n_par) + fused_draws_bytes(h->cfg.n_beta, has_data);
	memset(a, 0, sizeof(*a));
	a->data = has_data ? h->d_data : nullptr;
	a->n_rows = has_data ? (int) h->n_rows : 0;
	a->xabsmax = reinterpret_cast<const double *>(h->d_xabsmax);
}

template<class M>
static int run_fused_t(apm_gpu * h, long long n_rounds, int n_swap) {
	int threads = 0;
	size_t smem = 0;
	FusedArgs a;
	fused_geometry(h, M::HAS_DATA, &threads, &smem, &a);
	a.n_rounds = n_rounds;
	a.n_swap = n_swap;
	h->ev_used = 0;
	cudaEvent_t t0 = next_event(h), t1 = next_event(h);
	CU(cudaEventRecord(t0, h->stream));
	if constexpr (M::HAS_DATA) {
		fused_run_kernel<M><<<h->cfg.n_ensembles, threads, smem, h->stream>>>(h->S, a);
	} else {
		// a cluster of two CTAs per ensemble: one plays the chains, one draws (free_run_kernel)
		cudaLaunchConfig_t lc;
		memset(&lc, 0, sizeof(lc));
		lc.gridDim = dim3((unsigned) (h->cfg.n_ensembles * FREE_CLUSTER));
		lc.blockDim = dim3((unsigned) FREE_THREADS);
		lc.dynamicSmemBytes = smem;
		lc.stream = h->stream;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeClusterDimension;
		attr[0].val.clusterDim.x = (unsigned) FREE_CLUSTER;
		attr[0].val.clusterDim.y = 1;
		attr[0].val.clusterDim.z = 1;
		lc.attrs = attr;
		lc.numAttrs = 1;
		int max_clusters = 0;
		CU(cudaOccupancyMaxActiveClusters(&max_clusters, free_run_kernel<M>, &lc));
		if (max_clusters < 1)
			return fail(h, APM_ENOTAPPLICABLE, "a cluster of %d CTAs x %d threads x %zu bytes cannot be scheduled",
					FREE_CLUSTER, FREE_THREADS, smem);
		CU(cudaLaunchKernelEx(&lc, free_run_kernel<M>, h->S, a));
	}
	h->launches++;
	CU(cudaEventRecord(t1, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	CU(cudaGetLastError());
	float tot = 0;
	cudaEventElapsedTime(&tot, t0, t1);
	h->last_total_ms = tot;
	h->last_ll_ms = 0;
	h->last_ll_launches = 0;
	h->last_path = APM_PATH_FUSED;
	return APM_OK;
}

template<class M>
static int run_cluster_t(apm_gpu * h, long long n_rounds, int n_swap) {
	if (!M::HAS_DATA)
		return fail(h, APM_EINVAL, "the cluster path is for models with data");
	ClusterArgs ca;
	ca.timing = nullptr;
#ifdef APM_CLUSTER_TIMING
	long long * d_timing = nullptr;
	CU(cudaMalloc((void **) &d_timing, 16 * 8 * sizeof(long long)));
	CU(cudaMemset(d_timing, 0, 16 * 8 * sizeof(long long)));
	ca.timing = d_timing;
#endif
	int threads = 0;
	size_t smem_unused = 0;
	fused_geometry(h, true, &threads, &smem_unused, &ca.f);
	ca.f.n_rounds = n_rounds;
	ca.f.n_swap = n_swap;
	ca.cl = cluster_size_for(h, true);
	if (ca.cl < 2)
		return fail(h, APM_ESTATE, "no cluster geometry");
	ca.gmax = (h->cfg.n_beta + ca.cl - 1) / ca.cl;
	ca.wc = ca.gmax <= 8 ? FUSED_MAX_WARPS / ca.gmax : 1;
	cudaLaunchConfig_t lc;
	memset(&lc, 0, sizeof(lc));
	lc.gridDim = dim3((unsigned) (h->cfg.n_ensembles * ca.cl));
	lc.blockDim = dim3((unsigned) (32 * ca.wc * ca.gmax));
	lc.dynamicSmemBytes = cluster_smem_bytes(h->n_rows * (M::ROW_W / 2), ca.gmax, ca.wc, h->cfg.n_par);
	lc.stream = h->stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned) ca.cl;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	lc.attrs = attr;
	lc.numAttrs = 1;
	int max_clusters = 0;
	CU(cudaOccupancyMaxActiveClusters(&max_clusters, cluster_run_kernel<M>, &lc));
	if (max_clusters < 1)
		return fail(h, APM_ENOTAPPLICABLE, "a cluster of %d CTAs x %u threads x %zu bytes cannot be scheduled", ca.cl,
				lc.blockDim.x, lc.dynamicSmemBytes);
	h->ev_used = 0;
	cudaEvent_t t0 = next_event(h), t1 = next_event(h);
	CU(cudaEventRecord(t0, h->stream));
	CU(cudaLaunchKernelEx(&lc, cluster_run_kernel<M>, h->S, ca));
	h->launches++;
	CU(cudaEventRecord(t1, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	CU(cudaGetLastError());
	float tot = 0;
	cudaEventElapsedTime(&tot, t0, t1);
	h->last_total_ms = tot;
	h->last_ll_ms = 0;
	h->last_ll_launches = 0;
	h->last_path = APM_PATH_CLUSTER;
#ifdef APM_CLUSTER_TIMING
	{
		long long t[16 * 8];
		CU(cudaMemcpy(t, d_timing, sizeof(t), cudaMemcpyDeviceToHost));
		cudaFree(d_timing);
		const double steps = (double) n_rounds * n_swap;
		fprintf(stderr, "cluster timing, CTA 0 (cl %d, gmax %d, wc %d), cycles per step: warp: loop + deferred barrier "
				"wait | own work (rows / bookkeeping + draws) | barrier | decision + next proposal\n", ca.cl, ca.gmax, ca.wc);
		for (int w = 0; w < ca.wc * ca.gmax; w++)
			fprintf(stderr, "  warp %2d (chain %d, %s): %7.0f %7.0f %7.0f %7.0f\n", w, w / ca.wc,
					ca.wc > 1 && w % ca.wc == ca.wc - 1 ? "service" : "rows   ",
					t[w * 8 + 0] / steps, t[w * 8 + 1] / steps, t[w * 8 + 2] / steps, t[w * 8 + 3] / steps);
	}
#endif
	return APM_OK;
}

template<class M>
static int run_grid_t(apm_gpu * h, long long n_rounds, int n_swap) {
	if (!M::HAS_DATA)
		return fail(h, APM_EINVAL, "the grid path is for models with data");
	const int slice = grid_slice_rows(h, true);
	if (slice == 0)
		return fail(h, APM_ESTATE, "no grid geometry");
	const int G = h->sm_count;
	CU(ensure_cap(&h->d_grid_partials, &h->grid_partials_cap, (size_t) h->n_chains * G));
	CU(ensure_cap(&h->d_grid_draws, &h->grid_draws_cap, (size_t) h->n_chains * h->cfg.n_par));
	GridArgs ga;
	ga.data = h->d_data;
	ga.n_rows = h->n_rows;
	ga.xabsmax = reinterpret_cast<const double *>(h->d_xabsmax);
	ga.n_rounds = n_rounds;
	ga.n_swap = n_swap;
	ga.partials = h->d_grid_partials;
	ga.props = h->d_grid_draws;
	ga.max_slice_rows = slice;
	const size_t smem = (((size_t) slice * sizeof(Row<M>) + 127) & ~(size_t) 127) + 64
			+ grid_state_bytes(h->n_chains, h->cfg.n_par);
	int per_sm = 0;
	CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grid_run_kernel<M>, GRID_THREADS, smem));
	if (per_sm < 1)
		return fail(h, APM_ENOTAPPLICABLE, "the grid kernel does not fit on an SM (%zu bytes of shared memory)", smem);
	DevState S = h->S;
	void * args[] = { (void *) &S, (void *) &ga };
	h->ev_used = 0;
	cudaEvent_t t0 = next_event(h), t1 = next_event(h);
	CU(cudaEventRecord(t0, h->stream));
	{
		cudaError_t le = cudaLaunchCooperativeKernel((const void *) grid_run_kernel<M>, dim3((unsigned) G),
				dim3(GRID_THREADS), args, smem, h->stream);
		if (le == cudaErrorCooperativeLaunchTooLarge) { // not all SMs are ours (MPS, another context): nothing was launched
			cudaGetLastError();
			return fail(h, APM_ENOTAPPLICABLE, "the device cannot co-schedule %d CTAs for the grid path", G);
		}
		CU(le);
	}
	h->launches++;
	CU(cudaEventRecord(t1, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	CU(cudaGetLastError());
	float tot = 0;
	cudaEventElapsedTime(&tot, t0, t1);
	h->last_total_ms = tot;
	h->last_ll_ms = 0;
	h->last_ll_launches = 0;
	h->last_path = APM_PATH_GRID;
	return APM_OK;
}

// ------------------------------------------------------------------ run (tiled path)
static int setup_trace(apm_gpu * h, long long n_steps, const apm_gpu_trace_cfg * tr) {
	h->tr_prob_rows = h->tr_param_rows = 0;
	h->tr_dumped = 0;
	DevState & S = h->S;
	S.tr_prob = S.tr_dl = S.tr_params = nullptr;
	S.tr_prob_every = tr ? tr->prob_every : 0;
	S.tr_params_chains = tr ? tr->params_chains : 0;
	if (S.tr_prob_every < 0 || S.tr_params_chains < 0 || S.tr_params_chains > 2)
		return fail(h, APM_EINVAL, "bad trace configuration");
	if (S.tr_prob_every > 0) {
		h->tr_prob_rows = (n_steps + S.tr_prob_every - 1) / S.tr_prob_every;
		CU(ensure_cap(&h->d_tr_prob, &h->tr_prob_cap, (size_t) h->tr_prob_rows * h->n_chains));
		CU(ensure_cap(&h->d_tr_dl, &h->tr_dl_cap, (size_t) h->tr_prob_rows * h->n_chains));
		S.tr_prob = h->d_tr_prob;
		S.tr_dl = h->d_tr_dl;
	}
	h->tr_dumped = S.tr_params_chains == 2 ? h->n_chains : (S.tr_params_chains == 1 ? h->cfg.n_ensembles : 0);
	if (h->tr_dumped > 0) {
		h->tr_param_rows = n_steps;
		CU(ensure_cap(&h->d_tr_params, &h->tr_params_cap, (size_t) n_steps * h->tr_dumped * h->cfg.n_par));
		S.tr_params = h->d_tr_params;
	}
	S.tr_dumped = h->tr_dumped;
	return APM_OK;
}

// one round of the tiled path on the engine's stream: n_swap x {likelihood of every pending proposal,
// control kernel: finalise + record (+ the ensemble's swap after the round's last step) + next proposals}
template<class M>
static int enqueue_round(apm_gpu * h, int n_swap, bool timed) {
	AdvArgs a;
	memset(&a, 0, sizeof(a));
	for (int sub = 0; sub < n_swap; sub++) {
		int rc = step_likelihood<M>(h, timed);
		if (rc != APM_OK)
			return rc;
		const bool swap_now = sub + 1 == n_swap;
		a.flags = ADV_FINALIZE | ADV_RECORD;
		if (swap_now && h->ladder) {
			// the pair may straddle two GPUs: finish the step, trade the boundary chains with the
			// neighbours, then swap + propose with their copies at hand
			advance_kernel<M><<<h->cfg.n_ensembles, ADV_THREADS, 0, h->stream>>>(state_for_advance(h), a);
			h->launches++;
			rc = ladder_exchange(h);
			if (rc != APM_OK)
				return rc;
			a.flags = 0;
			a.pack_prev = h->rank > 0 ? h->d_pack_prev : nullptr;
			a.pack_next = h->rank + 1 < h->n_ranks ? h->d_pack_next : nullptr;
		}
		if (swap_now)
			a.flags |= ADV_SWAP;
		// (the proposal drawn after a call's last step is drawn again, from the same counter, by the
		// next call's first launch: a round is the same sequence of launches wherever it stands)
		a.flags |= ADV_PROPOSE_RUN;
		advance_kernel<M><<<h->cfg.n_ensembles, ADV_THREADS, 0, h->stream>>>(state_for_advance(h), a);
		h->launches++;
	}
	return APM_OK;
}

template<class M>
static int run_tiled_t(apm_gpu * h, long long n_rounds, int n_swap) {
	AdvArgs a;
	memset(&a, 0, sizeof(a));
	// (ladder split: the row splits are planned for the WHOLE ladder's chain count, as the single-GPU
	// run of the same ensembles plans them -- a chain's sum then has the same summation order on
	// both, which is what makes the split run equal to it bit for bit)
	int rc0 = replan(h, h->ladder ? h->cfg.n_ensembles * h->S.n_beta_total : h->n_chains);
	if (rc0 != APM_OK)
		return rc0;
	h->ev_used = 0;
	cudaEvent_t t0 = next_event(h), t1 = next_event(h);
	// ev[0], ev[1] are the run brackets; pairs from ev[2] on are likelihood launches (per-launch timing)
	CU(cudaEventRecord(t0, h->stream));
	CU(cudaMemsetAsync(h->S.run_ctr, 0, 2 * sizeof(unsigned long long), h->stream));
	a.flags = ADV_PROPOSE_RUN;
	advance_kernel<M><<<h->cfg.n_ensembles, ADV_THREADS, 0, h->stream>>>(state_for_advance(h), a);
	h->launches++;
	// Without per-launch timing and without a communicator in the step, a round is replayed as a
	// CUDA graph: 2 n_swap launches for one cudaGraphLaunch, nothing for the host to do per step.
	const bool use_graph = !h->per_launch_timing && !h->comm && getenv("APM_NO_GRAPH") == nullptr;
	if (use_graph) {
		// everything the captured launches carry by value: the device state struct (pointers, trace
		// configuration, adapt settings), the likelihood plan and table geometry, the round's length
		std::vector<unsigned char> key(sizeof(DevState) + 6 * sizeof(long long));
		const DevState S_now = state_for_advance(h);
		const long long extra[6] = { (long long) n_swap, (long long) h->plan_splits, (long long) h->plan_cps,
				h->n_rows, (long long) h->n_chunks, (long long) (size_t) h->d_data };
		memcpy(key.data(), &S_now, sizeof(DevState));
		memcpy(key.data() + sizeof(DevState), extra, sizeof(extra));
		if (h->round_graph == nullptr || key != h->round_graph_key) {
			if (h->round_graph) {
				cudaGraphExecDestroy(h->round_graph);
				h->round_graph = nullptr;
			}
			cudaGraph_t graph = nullptr;
			const long long launches_before = h->launches;
			CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
			int rc = enqueue_round<M>(h, n_swap, false);
			cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
			h->launches = launches_before; // captured, not launched
			if (rc != APM_OK) {
				if (graph)
					cudaGraphDestroy(graph);
				return rc;
			}
			if (ce != cudaSuccess)
				return fail(h, APM_ECUDA, "capturing a round failed: %s", cudaGetErrorString(ce));
			ce = cudaGraphInstantiate(&h->round_graph, graph, 0);
			cudaGraphDestroy(graph);
			if (ce != cudaSuccess) {
				h->round_graph = nullptr;
				return fail(h, APM_ECUDA, "instantiating the round graph failed: %s", cudaGetErrorString(ce));
			}
			h->round_graph_key = key;
		}
		const long long per_round = (M::HAS_DATA ? 2ll : 1ll) * n_swap;
		for (long long round = 0; round < n_rounds; round++) {
			CU(cudaGraphLaunch(h->round_graph, h->stream));
			h->launches += per_round;
		}
	} else {
		for (long long round = 0; round < n_rounds; round++) {
			int rc = enqueue_round<M>(h, n_swap, h->per_launch_timing != 0);
			if (rc != APM_OK) {
				cudaStreamSynchronize(h->stream); // nothing of this call may still be in flight
				return rc;
			}
		}
	}
	CU(cudaEventRecord(t1, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	CU(cudaGetLastError());
	// timing: skip the bracket pair
	h->last_ll_ms = 0;
	h->last_ll_launches = 0;
	for (size_t i = 2; i + 1 < h->ev_used; i += 2) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) == cudaSuccess) {
			h->last_ll_ms += ms;
			h->last_ll_launches++;
		}
	}
	float tot = 0;
	cudaEventElapsedTime(&tot, t0, t1);
	h->last_total_ms = tot;
	h->last_path = APM_PATH_TILED;
	return APM_OK;
}

extern "C" int apm_gpu_run(apm_gpu * h, long long n_rounds, int n_swap, const apm_gpu_trace_cfg * trace) {
	if (!h || n_rounds < 0 || n_swap < 1)
		return APM_EINVAL;
	if (!h->have_data || !h->have_bounds)
		return fail(h, APM_ESTATE, "apm_gpu_set_data and apm_gpu_set_bounds must be called first");
	CU(cudaSetDevice(h->cfg.device));
	int rc = setup_trace(h, n_rounds * n_swap, trace);
	if (rc != APM_OK)
		return rc;
	if (n_rounds == 0)
		return APM_OK;
	int path = APM_PATH_TILED;
	rc = choose_path(h, &path, true);
	if (rc != APM_OK)
		return rc;
	auto go = [&](int which) -> int {
		if (which == APM_PATH_CLUSTER) {
			DISPATCH(h->cfg.model_id, run_cluster_t, h, n_rounds, n_swap)
		}
		if (which == APM_PATH_GRID) {
			DISPATCH(h->cfg.model_id, run_grid_t, h, n_rounds, n_swap)
		}
		if (which == APM_PATH_FUSED) {
			DISPATCH(h->cfg.model_id, run_fused_t, h, n_rounds, n_swap)
		}
		DISPATCH(h->cfg.model_id, run_tiled_t, h, n_rounds, n_swap)
	};
	rc = go(path);
	if (rc == APM_ENOTAPPLICABLE) {
		// the chosen path cannot be scheduled right now (nothing has been launched).  Asked for by
		// name: an error.  Chosen by AUTO: the tiled path always works.
		if (h->cfg.path != APM_PATH_AUTO)
			return APM_ECUDA;
		rc = go(APM_PATH_TILED);
	}
	return rc;
}

extern "C" int apm_gpu_read_trace(apm_gpu * h, double * prob, double * dl, double * params,
		long long * n_prob_rows, long long * n_param_rows) {
	if (!h)
		return APM_EINVAL;
	CU(cudaSetDevice(h->cfg.device));
	if (prob && h->d_tr_prob && h->tr_prob_rows > 0)
		CU(cudaMemcpy(prob, h->d_tr_prob, (size_t) h->tr_prob_rows * h->n_chains * sizeof(double),
				cudaMemcpyDeviceToHost));
	if (dl && h->d_tr_dl && h->tr_prob_rows > 0)
		CU(cudaMemcpy(dl, h->d_tr_dl, (size_t) h->tr_prob_rows * h->n_chains * sizeof(double),
				cudaMemcpyDeviceToHost));
	if (params && h->d_tr_params && h->tr_param_rows > 0)
		CU(cudaMemcpy(params, h->d_tr_params,
				(size_t) h->tr_param_rows * h->tr_dumped * h->cfg.n_par * sizeof(double),
				cudaMemcpyDeviceToHost));
	if (n_prob_rows)
		*n_prob_rows = h->tr_prob_rows;
	if (n_param_rows)
		*n_param_rows = h->tr_param_rows;
	return APM_OK;
}

// ------------------------------------------------------------------ calibrate
template<class M>
static int calibrate_t(apm_gpu * h, const apm_gpu_calib_cfg * cfg, int * status, int n_selected, int steps_kind,
		long long steps_n) {
	AdvArgs a;
	memset(&a, 0, sizeof(a));
	a.cal.steps_kind = steps_kind; // steps_n > 0: apm_gpu_steps, not a calibration
	a.cal.steps_n = steps_n;
	a.cal.burn_in_iterations = cfg->burn_in_iterations;
	a.cal.desired_acceptance_rate = cfg->desired_acceptance_rate;
	a.cal.max_ar_deviation = cfg->max_ar_deviation;
	a.cal.iter_limit = cfg->iter_limit;
	a.cal.mul = cfg->mul;
	a.cal.adjust_step = cfg->adjust_step;
	a.cal.skip_calibrate = cfg->skip_calibrate;
	a.cal.iter_readjust = cfg->iter_readjust;
	a.cal.no_rescaling_limit = cfg->no_rescaling_limit;
	a.select = h->d_select;
	h->ev_used = 0;
	cudaEvent_t t0 = next_event(h), t1 = next_event(h);
	CU(cudaEventRecord(t0, h->stream));
	int path = APM_PATH_TILED;
	int rcp = choose_path(h, &path);
	if (rcp != APM_OK)
		return rcp;
	// few selected chains and the table in shared memory: warp groups per chain, the chains dealt
	// out over the SMs (AUTO, or asked for as APM_PATH_CLUSTER); otherwise a warp per chain
	const int ng = (path == APM_PATH_FUSED && h->cfg.path != APM_PATH_FUSED) ? group_calibrate_ng(h, M::HAS_DATA, n_selected) : 0;
	if (ng > 0) {
		if constexpr (M::HAS_DATA) {
			int threads = 0;
			size_t smem = 0;
			GroupArgs ga;
			fused_geometry(h, true, &threads, &smem, &ga.f);
			ga.f.cal = a.cal;
			ga.f.select = h->d_select;
			ga.sel_idx = h->d_sel_idx;
			ga.n_sel = n_selected;
			ga.ng = ng;
			CU(cudaMemcpyAsync(h->d_sel_idx, h->sel_idx.data(), (size_t) n_selected * sizeof(int), cudaMemcpyHostToDevice,
					h->stream));
			const int grid = std::min(n_selected, h->sm_count);
			smem = fused_table_bytes(h->n_rows * (M::ROW_W / 2)) + (size_t) ng * group_state_bytes(h->cfg.n_par);
			group_calibrate_kernel<M><<<grid, FUSED_MAX_WARPS * 32, smem, h->stream>>>(h->S, ga);
			h->launches++;
		}
		n_selected = 0;
		path = APM_PATH_CLUSTER;
	} else if (path == APM_PATH_FUSED) {
		int threads = 0;
		size_t smem = 0;
		FusedArgs f;
		fused_geometry(h, M::HAS_DATA, &threads, &smem, &f);
		f.cal = a.cal;
		f.select = h->d_select;
		fused_calibrate_kernel<M><<<h->cfg.n_ensembles, threads, smem, h->stream>>>(h->S, f);
		h->launches++;
		n_selected = 0; // skips the tiled loop below
	}
	if (path == APM_PATH_GRID) {
		if constexpr (M::HAS_DATA) {
			const int slice = grid_slice_rows(h, true);
			const int G = h->sm_count;
			CU(ensure_cap(&h->d_grid_partials, &h->grid_partials_cap, (size_t) h->n_chains * G));
			CU(ensure_cap(&h->d_grid_draws, &h->grid_draws_cap, (size_t) h->n_chains * h->cfg.n_par));
			CU(ensure_cap(&h->d_grid_active, &h->grid_active_cap, (size_t) h->n_chains));
			GridArgs ga;
			memset(&ga, 0, sizeof(ga));
			ga.data = h->d_data;
			ga.n_rows = h->n_rows;
			ga.xabsmax = reinterpret_cast<const double *>(h->d_xabsmax);
			ga.partials = h->d_grid_partials;
			ga.props = h->d_grid_draws;
			ga.max_slice_rows = slice;
			const size_t smem = (((size_t) slice * sizeof(Row<M>) + 127) & ~(size_t) 127) + 64
					+ grid_state_bytes(h->n_chains, h->cfg.n_par);
			DevState S = h->S;
			CalibCfgDev cd = a.cal;
			const unsigned char * sel = h->d_select;
			int * act = h->d_grid_active;
			void * args[] = { (void *) &S, (void *) &ga, (void *) &cd, (void *) &sel, (void *) &act };
			cudaError_t le = cudaLaunchCooperativeKernel((const void *) grid_calibrate_kernel<M>, dim3((unsigned) G),
					dim3(GRID_THREADS), args, smem, h->stream);
			if (le == cudaErrorCooperativeLaunchTooLarge && h->cfg.path == APM_PATH_AUTO) {
				cudaGetLastError();
				path = APM_PATH_TILED; // not all SMs are ours right now: the tiled path always works
			} else {
				CU(le);
				h->launches++;
			}
		}
		if (path == APM_PATH_GRID)
			n_selected = 0;
	}
	// (tiled path) the likelihood kernel walks a compacted list of the chains still calibrating;
	// the control kernel of step s fills list (s + 1) & 1 and clears list s & 1
	CU(cudaMemsetAsync(h->S.act_n, 0, 2 * sizeof(int), h->stream));
	if (path == APM_PATH_TILED) {
		a.flags = ADV_CALIB_BEGIN;
		a.act_w = 0;
		advance_kernel<M><<<h->cfg.n_ensembles, ADV_THREADS, 0, h->stream>>>(state_for_advance(h), a);
		h->launches++;
	}
	int active = n_selected;
	const int block = steps_n > 0 ? (int) std::min<long long>(steps_n, 200) : 200;
	long long step = 0;
	while (active > 0) {
		int rc = replan(h, active); // the row splits follow the number of chains left
		if (rc != APM_OK) {
			cudaStreamSynchronize(h->stream);
			return rc;
		}
		for (int i = 0; i < block; i++, step++) {
			rc = step_likelihood<M>(h, true, (int) (step & 1), active);
			if (rc != APM_OK) {
				cudaStreamSynchronize(h->stream); // nothing of this call may still be in flight
				return rc;
			}
			a.flags = ADV_FINALIZE | ADV_CALIB;
			a.act_w = (int) ((step + 1) & 1);
			advance_kernel<M><<<h->cfg.n_ensembles, ADV_THREADS, 0, h->stream>>>(state_for_advance(h), a);
			h->launches++;
		}
		CU(cudaMemcpyAsync(&active, h->S.n_active, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
		CU(cudaStreamSynchronize(h->stream));
	}
	CU(cudaEventRecord(t1, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	CU(cudaGetLastError());
	h->last_ll_ms = 0;
	h->last_ll_launches = 0;
	for (size_t i = 2; i + 1 < h->ev_used; i += 2) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) == cudaSuccess) {
			h->last_ll_ms += ms;
			h->last_ll_launches++;
		}
	}
	float tot = 0;
	cudaEventElapsedTime(&tot, t0, t1);
	h->last_total_ms = tot;
	h->last_path = path;
	// per-chain status
	std::vector<CalState> cs(h->n_chains);
	CU(cudaMemcpy(cs.data(), h->S.cal, h->n_chains * sizeof(CalState), cudaMemcpyDeviceToHost));
	int any = 0;
	for (int g = 0; g < h->n_chains; g++) {
		if (status)
			status[g] = cs[g].status;
		if (cs[g].status > 0)
			any = 1;
	}
	if (any)
		return fail(h, APM_ECALIB, "calibration failed for at least one chain (see status[])");
	return APM_OK;
}

extern "C" int apm_gpu_calibrate(apm_gpu * h, const unsigned char * select, const apm_gpu_calib_cfg * cfg,
		int * status, apm_gpu_calib_progress * progress, long long progress_capacity, long long * n_progress) {
	if (!h || !cfg)
		return APM_EINVAL;
	if (!h->have_data || !h->have_bounds)
		return fail(h, APM_ESTATE, "apm_gpu_set_data and apm_gpu_set_bounds must be called first");
	CU(cudaSetDevice(h->cfg.device));
	std::vector<unsigned char> sel(h->n_chains, 1);
	if (select)
		memcpy(sel.data(), select, h->n_chains);
	CU(cudaMemcpy(h->d_select, sel.data(), h->n_chains, cudaMemcpyHostToDevice));
	CU(cudaMemset(h->S.n_active, 0, sizeof(int)));
	CU(cudaMemset(h->S.progress_n, 0, sizeof(unsigned long long)));
	if (h->S.progress)
		cudaFree(h->S.progress);
	h->S.progress = nullptr;
	h->S.progress_cap = 0;
	if (progress && progress_capacity > 0) {
		CU(dalloc(&h->S.progress, (size_t) progress_capacity));
		h->S.progress_cap = progress_capacity;
	}
	int n_selected = 0;
	h->sel_idx.clear();
	for (int g = 0; g < h->n_chains; g++)
		if (sel[g]) {
			n_selected++;
			h->sel_idx.push_back(g);
		}
	auto go = [&]() -> int { DISPATCH(h->cfg.model_id, calibrate_t, h, cfg, status, n_selected, 0, 0) };
	int rc = go();
	if (rc != APM_OK && rc != APM_ECALIB)
		return rc;
	unsigned long long np_rows = 0;
	CU(cudaMemcpy(&np_rows, h->S.progress_n, sizeof(np_rows), cudaMemcpyDeviceToHost));
	if (n_progress)
		*n_progress = (long long) np_rows;
	if (progress && progress_capacity > 0) {
		static_assert(sizeof(ProgressRow) == sizeof(apm_gpu_calib_progress), "progress row layout");
		size_t take = (size_t) std::min<unsigned long long>(np_rows, (unsigned long long) progress_capacity);
		CU(cudaMemcpy(progress, h->S.progress, take * sizeof(ProgressRow), cudaMemcpyDeviceToHost));
		// rows are appended by concurrently calibrating chains: hand them out ordered by
		// (chain, iter, param), the order the reference writes them per chain
		std::sort(progress, progress + take, [](const apm_gpu_calib_progress & x, const apm_gpu_calib_progress & y) {
			if (x.chain != y.chain) return x.chain < y.chain;
			if (x.iter != y.iter) return x.iter < y.iter;
			return x.param < y.param;
		});
	}
	return rc;
}

// ------------------------------------------------------------------ steps with an accept log
extern "C" int apm_gpu_steps(apm_gpu * h, const unsigned char * select, int kind, long long n_steps,
		unsigned char * accepted) {
	if (!h || n_steps < 0 || kind < 0 || kind > h->cfg.n_par)
		return APM_EINVAL;
	if (!h->have_data || !h->have_bounds)
		return fail(h, APM_ESTATE, "apm_gpu_set_data and apm_gpu_set_bounds must be called first");
	if (n_steps == 0)
		return APM_OK;
	CU(cudaSetDevice(h->cfg.device));
	std::vector<unsigned char> sel(h->n_chains, 1);
	if (select)
		memcpy(sel.data(), select, h->n_chains);
	CU(cudaMemcpy(h->d_select, sel.data(), h->n_chains, cudaMemcpyHostToDevice));
	CU(cudaMemset(h->S.n_active, 0, sizeof(int)));
	CU(cudaMemset(h->S.progress_n, 0, sizeof(unsigned long long)));
	if (h->S.progress)
		cudaFree(h->S.progress);
	h->S.progress = nullptr;
	h->S.progress_cap = 0;
	int n_selected = 0;
	h->sel_idx.clear();
	for (int g = 0; g < h->n_chains; g++)
		if (sel[g]) {
			n_selected++;
			h->sel_idx.push_back(g);
		}
	const size_t log_bytes = (size_t) n_steps * h->n_chains;
	unsigned char * d_log = nullptr;
	CU(cudaMalloc((void **) &d_log, log_bytes));
	cudaError_t e = cudaMemset(d_log, 0, log_bytes);
	if (e != cudaSuccess) {
		cudaFree(d_log);
		return fail(h, APM_ECUDA, "clearing the accept log failed: %s", cudaGetErrorString(e));
	}
	h->S.alog = d_log;
	apm_gpu_calib_cfg cfg;
	memset(&cfg, 0, sizeof(cfg));
	auto go = [&]() -> int { DISPATCH(h->cfg.model_id, calibrate_t, h, &cfg, nullptr, n_selected, kind, n_steps) };
	int rc = go();
	// whatever happened, no kernel that writes the log may still be in flight when it is freed
	cudaStreamSynchronize(h->stream);
	h->S.alog = nullptr;
	if (rc == APM_OK && accepted)
		e = cudaMemcpy(accepted, d_log, log_bytes, cudaMemcpyDeviceToHost);
	cudaFree(d_log);
	if (e != cudaSuccess)
		return fail(h, APM_ECUDA, "reading the accept log failed: %s", cudaGetErrorString(e));
	return rc;
}

// ------------------------------------------------------------------ host-side uniforms
extern "C" int apm_gpu_host_uniform(apm_gpu * h, int g, double * u) {
	if (!h || !u || g < 0 || g >= h->n_chains)
		return APM_EINVAL;
	if (h->host_draws.size() != (size_t) h->n_chains)
		h->host_draws.assign(h->n_chains, 0ull);
	const DevState & S = h->S;
	const uint32_t id = (uint32_t) (S.chain_id_offset + (g / S.n_beta) * S.id_stride + S.k_offset + g % S.n_beta);
	double u1;
	philox_uniforms(S.seed, id, h->host_draws[g]++, PURPOSE_HOST, 0, 0, *u, u1);
	return APM_OK;
}

// ------------------------------------------------------------------ -DADAPT
extern "C" int apm_gpu_set_adapt(apm_gpu * h, int enabled, double target_acceptance_rate) {
	if (!h)
		return APM_EINVAL;
	if (enabled && !(target_acceptance_rate > 0 && target_acceptance_rate < 1))
		return fail(h, APM_EINVAL, "adapt: target acceptance rate must lie in (0, 1)");
	h->S.adapt = enabled ? 1 : 0;
	h->S.adapt_target = target_acceptance_rate;
	return APM_OK;
}

// ------------------------------------------------------------------ accumulators
extern "C" int apm_gpu_reset_stats(apm_gpu * h) {
	if (!h)
		return APM_EINVAL;
	CU(cudaSetDevice(h->cfg.device));
	const size_t n = h->n_chains, nv = n * h->cfg.n_par;
	CU(cudaMemset(h->S.stat_n, 0, n * sizeof(u64)));
	CU(cudaMemset(h->S.stat_sum_dl, 0, n * sizeof(double)));
	CU(cudaMemset(h->S.stat_sum_p, 0, nv * sizeof(double)));
	CU(cudaMemset(h->S.stat_sum_p2, 0, nv * sizeof(double)));
	return APM_OK;
}

extern "C" int apm_gpu_get_stats(apm_gpu * h, unsigned long long * n, double * sum_dl, double * sum_p,
		double * sum_p2) {
	if (!h)
		return APM_EINVAL;
	CU(cudaSetDevice(h->cfg.device));
	const size_t nc = h->n_chains, nv = nc * h->cfg.n_par;
	if (n) CU(cudaMemcpy(n, h->S.stat_n, nc * sizeof(u64), cudaMemcpyDeviceToHost));
	if (sum_dl) CU(cudaMemcpy(sum_dl, h->S.stat_sum_dl, nc * sizeof(double), cudaMemcpyDeviceToHost));
	if (sum_p) CU(cudaMemcpy(sum_p, h->S.stat_sum_p, nv * sizeof(double), cudaMemcpyDeviceToHost));
	if (sum_p2) CU(cudaMemcpy(sum_p2, h->S.stat_sum_p2, nv * sizeof(double), cudaMemcpyDeviceToHost));
	return APM_OK;
}

// ------------------------------------------------------------------ marginal statistics
extern "C" int apm_gpu_set_marginals(apm_gpu * h, int which_chains, int n_bins, unsigned long long batch_size,
		int max_batches) {
	if (!h)
		return APM_EINVAL;
	if (which_chains < 0 || which_chains > 2 || (which_chains && (n_bins < 1 || max_batches < 0)))
		return fail(h, APM_EINVAL, "marginals: which_chains 0..2, n_bins >= 1, max_batches >= 0");
	CU(cudaSetDevice(h->cfg.device));
	CU(cudaStreamSynchronize(h->stream));
	DevState & S = h->S;
	void * old[] = { S.marg_counts, S.marg_bsum, S.marg_means, S.marg_n, S.marg_nb };
	for (void * p : old)
		if (p)
			cudaFree(p);
	S.marg_counts = nullptr;
	S.marg_bsum = nullptr;
	S.marg_means = nullptr;
	S.marg_n = S.marg_nb = nullptr;
	S.marg_mode = 0;
	if (which_chains == 0)
		return APM_OK;
	const size_t slots = which_chains == 2 ? (size_t) h->n_chains : (size_t) h->cfg.n_ensembles;
	const size_t np = h->cfg.n_par;
	cudaError_t e = dalloc(&S.marg_counts, slots * np * n_bins);
	if (e == cudaSuccess) e = dalloc(&S.marg_bsum, slots * np);
	if (e == cudaSuccess) e = dalloc(&S.marg_means, slots * np * std::max(max_batches, 1));
	if (e == cudaSuccess) e = dalloc(&S.marg_n, slots);
	if (e == cudaSuccess) e = dalloc(&S.marg_nb, slots);
	if (e != cudaSuccess)
		return fail(h, APM_ENOMEM, "marginals: device allocation failed: %s", cudaGetErrorString(e));
	S.marg_mode = which_chains;
	S.marg_bins = n_bins;
	S.marg_batch = batch_size;
	S.marg_cap = max_batches;
	return APM_OK;
}

extern "C" int apm_gpu_get_marginals(apm_gpu * h, unsigned long long * counts, double * batch_means,
		unsigned long long * n_values, unsigned long long * n_batches) {
	if (!h)
		return APM_EINVAL;
	const DevState & S = h->S;
	if (!S.marg_mode)
		return fail(h, APM_ESTATE, "apm_gpu_set_marginals has not been called");
	CU(cudaSetDevice(h->cfg.device));
	const size_t slots = S.marg_mode == 2 ? (size_t) h->n_chains : (size_t) h->cfg.n_ensembles;
	const size_t np = h->cfg.n_par;
	if (counts) CU(cudaMemcpy(counts, S.marg_counts, slots * np * S.marg_bins * sizeof(u64), cudaMemcpyDeviceToHost));
	if (batch_means && S.marg_cap > 0)
		CU(cudaMemcpy(batch_means, S.marg_means, slots * np * S.marg_cap * sizeof(double), cudaMemcpyDeviceToHost));
	if (n_values) CU(cudaMemcpy(n_values, S.marg_n, slots * sizeof(u64), cudaMemcpyDeviceToHost));
	if (n_batches) CU(cudaMemcpy(n_batches, S.marg_nb, slots * sizeof(u64), cudaMemcpyDeviceToHost));
	return APM_OK;
}

// ------------------------------------------------------------------ NCCL
extern "C" int apm_gpu_nccl_unique_id(unsigned char id_out[128]) {
	if (!g_nccl.load())
		return fail(nullptr, APM_ENCCL, "libnccl.so.2 not found");
	ncclUniqueId id;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
	ncclResult_t r = g_nccl.GetUniqueId(&id);
	if (r != 0)
		return fail(nullptr, APM_ENCCL, "ncclGetUniqueId failed (%d)", r);
	memcpy(id_out, &id, 128);
	return APM_OK;
}

extern "C" int apm_gpu_nccl_init(apm_gpu * h, const unsigned char id_in[128], int rank, int n_ranks) {
	if (!h || !id_in || n_ranks < 1 || rank < 0 || rank >= n_ranks)
		return APM_EINVAL;
	if (!g_nccl.load())
		return fail(h, APM_ENCCL, "libnccl.so.2 not found");
	CU(cudaSetDevice(h->cfg.device));
	ncclUniqueId id;
	memcpy(&id, id_in, 128);
	ncclResult_t r = g_nccl.CommInitRank(&h->comm, n_ranks, id, rank);
	if (r != 0) {
		h->comm = nullptr;
		return fail(h, APM_ENCCL, "ncclCommInitRank failed: %s",
				g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
	}
	h->rank = rank;
	h->n_ranks = n_ranks;
	return APM_OK;
}

extern "C" int apm_gpu_ladder_init(apm_gpu * h, const unsigned char id_in[128], int rank, int n_ranks,
		int n_beta_total) {
	if (!h || !id_in || n_ranks < 1 || rank < 0 || rank >= n_ranks || n_beta_total < n_ranks)
		return APM_EINVAL;
	if (h->comm)
		return fail(h, APM_ESTATE, "this handle already has a communicator");
	const int k0 = (int) ((long long) n_beta_total * rank / n_ranks);
	const int k1 = (int) ((long long) n_beta_total * (rank + 1) / n_ranks);
	if (k1 - k0 != h->cfg.n_beta)
		return fail(h, APM_EINVAL, "rank %d of %d holds rungs [%d, %d) of %d: the handle must be created with "
				"n_beta = %d, not %d", rank, n_ranks, k0, k1, n_beta_total, k1 - k0, h->cfg.n_beta);
	int rc = apm_gpu_nccl_init(h, id_in, rank, n_ranks);
	if (rc != APM_OK)
		return rc;
	const size_t count = (size_t) h->cfg.n_ensembles * LADDER_PACK(h->cfg.n_par);
	CU(dalloc(&h->d_pack_first, count));
	CU(dalloc(&h->d_pack_last, count));
	CU(dalloc(&h->d_pack_prev, count));
	CU(dalloc(&h->d_pack_next, count));
	h->ladder = true;
	h->S.n_beta_total = n_beta_total;
	h->S.id_stride = n_beta_total;
	h->S.k_offset = k0;
	return APM_OK;
}

// ------------------------------------------------------------------ introspection
extern "C" long long apm_gpu_launch_count(const apm_gpu * h) { return h ? h->launches : 0; }

extern "C" int apm_gpu_last_kernel_ms(const apm_gpu * h, double * loglik_ms, long long * loglik_launches,
		double * total_ms) {
	if (!h)
		return APM_EINVAL;
	if (loglik_ms) *loglik_ms = h->last_ll_ms;
	if (loglik_launches) *loglik_launches = h->last_ll_launches;
	if (total_ms) *total_ms = h->last_total_ms;
	return APM_OK;
}

extern "C" int apm_gpu_last_path(const apm_gpu * h) { return h ? h->last_path : 0; }

extern "C" int apm_gpu_set_timing(apm_gpu * h, int per_launch) {
	if (!h)
		return APM_EINVAL;
	h->per_launch_timing = per_launch ? 1 : 0;
	return APM_OK;
}

extern "C" int apm_gpu_measure_fp64_peak(int device, double seconds, double * instr_per_s) {
	apm_gpu * h = nullptr;
	if (!instr_per_s)
		return APM_EINVAL;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
		cudaGetLastError();
		return fail(h, APM_ENODEVICE, "no such CUDA device");
	}
	CU(cudaSetDevice(device));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	const int grid = prop.multiProcessorCount * 8, block = 256;
	double * out = nullptr;
	CU(dalloc(&out, (size_t) grid * block));
	cudaEvent_t t0, t1;
	CU(cudaEventCreate(&t0));
	CU(cudaEventCreate(&t1));
	int iters = 500;
	double best = 0;
	// warm up, then grow the launch until it lasts long enough to average over clock ramps
	for (int rep = 0; rep < 12; rep++) {
		CU(cudaEventRecord(t0));
		fp64_peak_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-9);
		CU(cudaEventRecord(t1));
		CU(cudaEventSynchronize(t1));
		float ms = 0;
		CU(cudaEventElapsedTime(&ms, t0, t1));
		double rate = (double) grid * block * (double) iters * 256.0 / (ms * 1e-3);
		if (rep >= 2 && rate > best)
			best = rate;
		if (ms * 1e-3 < seconds / 4 && iters < (1 << 24))
			iters *= 2;
	}
	cudaEventDestroy(t0);
	cudaEventDestroy(t1);
	cudaFree(out);
	*instr_per_s = best;
	return APM_OK;
}

extern "C" int apm_gpu_measure_fp64_per_clock(int device, double * lanes_per_sm_per_clock, int * n_sm) {
	apm_gpu * h = nullptr;
	if (!lanes_per_sm_per_clock)
		return APM_EINVAL;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
		cudaGetLastError();
		return fail(h, APM_ENODEVICE, "no such CUDA device");
	}
	CU(cudaSetDevice(device));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	// one 256-thread block per SM: 2 warps per scheduler x 8 independent chains, like the likelihood kernel
	const int grid = prop.multiProcessorCount, block = 256, iters = 2000;
	double * out = nullptr;
	long long * cyc = nullptr;
	CU(dalloc(&out, (size_t) grid * block));
	CU(dalloc(&cyc, (size_t) grid));
	std::vector<long long> host(grid);
	double best = 0;
	for (int rep = 0; rep < 6; rep++) {
		fp64_peak_clock_kernel<<<grid, block>>>(out, cyc, iters, 0.999999, 1e-9);
		CU(cudaDeviceSynchronize());
		CU(cudaMemcpy(host.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
		double mean = 0;
		for (long long c : host)
			mean += (double) c;
		mean /= grid;
		const double rate = (double) block * iters * 256.0 / mean;
		if (rep >= 1 && rate > best)
			best = rate;
	}
	cudaFree(out);
	cudaFree(cyc);
	*lanes_per_sm_per_clock = best;
	if (n_sm)
		*n_sm = prop.multiProcessorCount;
	return APM_OK;
}
