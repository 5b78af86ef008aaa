// apm_kernels.cuh -- the CUDA kernels (sm_100a).
//
//   loglik_tiled_kernel<M>   the hot kernel of the tiled path: a persistent grid walks
//                            (chain tile x row split) work items; the data rows stream
//                            through a shared-memory ring filled by the TMA engine
//                            (cp.async.bulk + full/empty mbarriers, no CTA barrier in the
//                            chunk loop); each thread keeps the tile's chain constants and
//                            accumulators in registers, so every LDS.128 of a row feeds all
//                            chains of the tile; per-chain sums are folded in a fixed order
//                            (fp64 warp shuffles).
//   advance_kernel<M>        the control kernel of the tiled path, one CTA per ensemble:
//                            finalise the pending step of every chain (accept/reject,
//                            counters, best, trace, accumulators), resolve the ensemble's
//                            swap, drive the calibration state machines, and draw the next
//                            proposals -- all reference semantics live in apm_chain.cuh.
//   fused_run_kernel<M>      the small-table path: one CTA per ensemble, one warp per chain,
//   fused_calibrate_kernel<M>  table and ensemble state resident in shared memory, a whole
//                            run (n_rounds x (n_swap steps + swap)) or calibration per launch.
//   cluster_run_kernel<M>    the fused path with an ensemble spread over a thread-block cluster.
//   grid_run_kernel<M>       mid-size tables: one cooperative launch per run, the table partitioned
//                            over the shared memories of all SMs, one grid barrier per step.
//   eval_finish_kernel<M>    turns running sums into (prob, prior) for apm_gpu_eval.
//   absmax_col0_kernel       max |x| of the table (range bound of the fast sine).
//   fp64_peak_kernel         DFMA issue-rate microbenchmark (the roofline denominator).
#pragma once

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <type_traits>
#include "apm_chain.cuh"

namespace apm {

// ------------------------------------------------------------------ mbarrier / TMA PTX
__device__ __forceinline__ uint32_t smem_u32(const void * p) {
	return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t * bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
			:: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
	asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"WAIT_%=:\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
			"@p bra DONE_%=;\n\t"
			"bra WAIT_%=;\n\t"
			"DONE_%=:\n\t"
			"}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared by the TMA engine (SASS: UBLKCP), completion on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void * dst_smem, const void * src_gmem, uint32_t bytes,
		uint64_t * bar) {
	asm volatile(
			"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// ------------------------------------------------------------------ table rows
// A table row on the device is M::ROW_W doubles: 2 (double2: the gsl_matrix layout of a two-column
// file) or 4 (double4: models reading three or four columns, padded with zeros).  Everything
// that stages rows is written for `Row<M>` and 16 * ROW_W / 2 bytes per row.
template<int W> struct RowT;
template<> struct RowT<2> { typedef double2 type; };
template<> struct RowT<4> { typedef double4 type; };
template<class M> using Row = typename RowT<M::ROW_W>::type;

// A thread's running sum over rows is a plain double unless the model declares its own
// accumulator (`typedef ... Acc; acc_zero(); acc_merge(a, b); acc_value(a)`): pulse and pulse_vrot
// carry the product of their quotients next to the sum, so that the rows' logarithms become one
// logarithm per thread (apm_models.cuh).  acc_merge folds two accumulators of a thread,
// acc_value gives the double that enters the cross-thread reduction.
template<class M, class = void>
struct ModelAcc {
	typedef double type;
	__device__ __forceinline__ static double zero() { return 0.0; }
	__device__ __forceinline__ static double merge(double a, double b) { return a + b; }
	__device__ __forceinline__ static double value(double a) { return a; }
};
template<class M>
struct ModelAcc<M, std::void_t<typename M::Acc>> {
	typedef typename M::Acc type;
	__device__ __forceinline__ static type zero() { return M::acc_zero(); }
	__device__ __forceinline__ static type merge(const type & a, const type & b) { return M::acc_merge(a, b); }
	__device__ __forceinline__ static double value(const type & a) { return M::acc_value(a); }
};
template<class M> using Acc = typename ModelAcc<M>::type;

template<class M>
__device__ __forceinline__ Acc<M> row_accum(const Acc<M> & acc, const typename M::Prep & q, const Row<M> & r) {
	if constexpr (M::ROW_W == 2)
		return M::accum(acc, q, r.x, r.y);
	else
		return M::accum(acc, q, r);
}
template<class M>
__device__ __forceinline__ Acc<M> row_accum_fast(const Acc<M> & acc, const typename M::Prep & q, const Row<M> & r) {
	if constexpr (M::ROW_W == 2)
		return M::accum_fast(acc, q, r.x, r.y);
	else
		return M::accum_fast(acc, q, r);
}

// ------------------------------------------------------------------ tiled likelihood
// tuning knobs (overridable at build time for kernel sweeps: tools/kernel_sweep.py)
#ifndef APM_LL_THREADS
#define APM_LL_THREADS 256
#endif
#ifndef APM_LL_RPT
#define APM_LL_RPT 16
#endif
#ifndef APM_LL_STAGES
#define APM_LL_STAGES 3
#endif
#ifndef APM_LL_UNROLL
#define APM_LL_UNROLL 4
#endif
constexpr int LL_THREADS = APM_LL_THREADS;
constexpr int LL_WARPS = LL_THREADS / 32;
constexpr int LL_RPT = APM_LL_RPT;              // rows per thread per chunk
constexpr int LL_CHUNK = LL_THREADS * LL_RPT;   // two-column rows per TMA chunk (4096 rows = 64 KB)
// rows per thread / per chunk for a model: a chunk is always LL_CHUNK * 16 bytes
template<class M> __host__ __device__ constexpr int ll_rpt() { return LL_RPT * 2 / M::ROW_W; }
template<class M> __host__ __device__ constexpr int ll_chunk() { return LL_THREADS * ll_rpt<M>(); }
constexpr int LL_STAGES = APM_LL_STAGES;
constexpr int LL_UNROLL = APM_LL_UNROLL;          // inner iterations unrolled together
constexpr int LL_MAX_C = 8;                     // upper bound of M::LL_C (chains per work item)

struct LLArgs {
	const double * data;    // [n_rows_padded][ROW_W], padded with zeros to a multiple of the chunk
	long long n_rows;
	const double * prop;    // [n_slots][n_par] parameter vectors to evaluate
	const int * act_idx;    // optional compacted list of the slots to evaluate (NULL = all n_slots)
	const int * act_n;      // device-side length of act_idx (read only when act_idx != NULL)
	int n_slots;
	int n_par;
	int n_splits;
	int chunks_per_split;
	int n_chunks;
	const double * xabsmax; // device scalar: max |x| over the table (bound for the fast sine)
	double * partial;       // [n_slots][n_splits]
	double model_const[4];
};

__device__ __forceinline__ void mbar_arrive(uint64_t * bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// Work item = (tile of C = M::LL_C chains) x (row split).  Each thread keeps the C chains'
// constants (M::Prep) and C x U fp64 accumulators in registers for the whole item and walks
// its rows of every staged chunk: one conflict-free LDS.128 per row feeds C row evaluations,
// so the inner loop is a steady stream of C x U independent dependency chains on the fp64
// pipe with no per-chain loop, branch, shared-memory accumulator or barrier inside it.
//
// The rows stream through an LL_STAGES-deep shared-memory ring filled by the TMA engine
// (cp.async.bulk, completion on a `full` mbarrier with expect_tx).  A stage is handed back
// through an `empty` mbarrier on which every warp arrives once it has read its rows, so no
// CTA-wide barrier exists in the chunk loop; the ring runs ahead across work-item boundaries
// (the producer walks the CTA's item list on its own), so the first chunk of the next item
// is already resident when the fold of the current one ends.
template<class M>
__global__ void __launch_bounds__(LL_THREADS, 1) loglik_tiled_kernel(const LLArgs a) {
	constexpr int C = M::LL_C, U = M::LL_U;
	constexpr int RPT = ll_rpt<M>(), CHUNK = ll_chunk<M>(); // rows per thread per chunk, rows per chunk
	static_assert(C <= LL_MAX_C && RPT % U == 0, "tile shape");
	extern __shared__ __align__(128) unsigned char ll_smem[];
	Row<M> * sdata = reinterpret_cast<Row<M> *>(ll_smem);                  // [STAGES][CHUNK]
	double * sacc = reinterpret_cast<double *>(sdata + LL_STAGES * CHUNK); // [MAX_C][THREADS]
	uint64_t * full = reinterpret_cast<uint64_t *>(sacc + LL_MAX_C * LL_THREADS); // [STAGES]
	uint64_t * empty = full + LL_STAGES;                                     // [STAGES]

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n_par = a.n_par;
	if (tid == 0) {
		for (int s = 0; s < LL_STAGES; s++) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], LL_WARPS);
		}
		mbar_fence_init();
	}
	__syncthreads();

	const int n_act = a.act_idx ? *a.act_n : a.n_slots;
	const int n_ctiles = (n_act + C - 1) / C;
	const int n_items = n_ctiles * a.n_splits; // < 2^31, checked by the host
	const double xub = *a.xabsmax;
	constexpr uint32_t CHUNK_BYTES = CHUNK * sizeof(Row<M>);

	// ---- producer (thread 0): walks this CTA's items chunk by chunk, LL_STAGES - 1 ahead
	int p_item = blockIdx.x;
	int p_k = 0;
	uint32_t p_n = 0;
	auto produce = [&]() {
		if (p_item >= n_items)
			return;
		const int k0 = (p_item / n_ctiles) * a.chunks_per_split;
		const int nk = min(a.chunks_per_split, a.n_chunks - k0);
		const uint32_t st = p_n % LL_STAGES;
		if (p_n >= LL_STAGES) // chunk p_n - LL_STAGES must have been read by every warp
			mbar_wait(&empty[st], ((p_n / LL_STAGES) + 1u) & 1u);
		mbar_arrive_expect_tx(&full[st], CHUNK_BYTES);
		tma_bulk_g2s(sdata + st * CHUNK, a.data + (size_t) (k0 + p_k) * CHUNK * M::ROW_W, CHUNK_BYTES,
				&full[st]);
		p_n++;
		if (++p_k == nk) {
			p_k = 0;
			p_item += gridDim.x;
		}
	};
	if (tid == 0)
		for (int s = 0; s < LL_STAGES - 1; s++)
			produce();

	uint32_t it = 0; // chunks consumed so far: stage = it % STAGES, parity = (it / STAGES) & 1
	for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
		const int split = item / n_ctiles;
		const int ctile = item - split * n_ctiles; // chain tile fastest: neighbours share rows in L2
		const int c0 = ctile * C;
		const int k0 = split * a.chunks_per_split;
		const int nk = min(a.chunks_per_split, a.n_chunks - k0);

		typename M::Prep q[C];
		int gid[C];
		Acc<M> acc[C][U];
		bool fast = true;
#pragma unroll
		for (int c = 0; c < C; c++) {
			const int slot = c0 + c;
			gid[c] = slot < n_act ? (a.act_idx ? a.act_idx[slot] : slot) : -1;
			// slots past the end of the list evaluate the tile's first chain again (never written)
			const int src = gid[c] >= 0 ? gid[c] : gid[0];
			M::prep(q[c], a.prop + (size_t) src * n_par, n_par, a.model_const);
			fast = fast && M::fast_ok(q[c], xub);
#pragma unroll
			for (int u = 0; u < U; u++)
				acc[c][u] = ModelAcc<M>::zero();
		}

		for (int k = 0; k < nk; k++, it++) {
			const uint32_t st = it % LL_STAGES;
			mbar_wait(&full[st], (it / LL_STAGES) & 1u);
			if (tid == 0)
				produce(); // chunk it + STAGES - 1 goes where chunk it - 1 was
			const Row<M> * srow = sdata + st * CHUNK + tid;
			const long long row0 = (long long) (k0 + k) * CHUNK;
			const int n_valid = (int) min((long long) CHUNK, a.n_rows - row0);
			if (fast && n_valid == CHUNK) {
				// branch-free path: every row of the chunk is real and inside the fast range
#pragma unroll LL_UNROLL
				for (int j = 0; j < RPT; j += U) {
					Row<M> r[U];
#pragma unroll
					for (int u = 0; u < U; u++)
						r[u] = srow[(j + u) * LL_THREADS];
#pragma unroll
					for (int c = 0; c < C; c++)
#pragma unroll
						for (int u = 0; u < U; u++)
							acc[c][u] = row_accum_fast<M>(acc[c][u], q[c], r[u]);
				}
			} else if (fast) {
				// ragged last chunk: padding rows masked out
#pragma unroll 1
				for (int j = 0; j < RPT; j++) {
					const Row<M> r = srow[j * LL_THREADS];
					if (j * LL_THREADS + tid < n_valid) {
#pragma unroll
						for (int c = 0; c < C; c++)
							acc[c][0] = row_accum_fast<M>(acc[c][0], q[c], r);
					}
				}
			} else {
				// some chain of the tile is outside the fast sine's range: exact path
#pragma unroll 1
				for (int j = 0; j < RPT; j++) {
					const Row<M> r = srow[j * LL_THREADS];
					if (j * LL_THREADS + tid < n_valid) {
#pragma unroll
						for (int c = 0; c < C; c++)
							acc[c][0] = row_accum<M>(acc[c][0], q[c], r);
					}
				}
			}
			__syncwarp();
			if (lane == 0)
				mbar_arrive(&empty[st]);
		}

		// fold the item in a fixed order: a thread's U accumulators in index order, then per
		// chain the 8 strided thread slots of a lane in index order, then a butterfly
#pragma unroll
		for (int c = 0; c < C; c++) {
			Acc<M> v = acc[c][0];
#pragma unroll
			for (int u = 1; u < U; u++)
				v = ModelAcc<M>::merge(v, acc[c][u]);
			sacc[c * LL_THREADS + tid] = ModelAcc<M>::value(v);
		}
		__syncthreads();
		for (int c = warp; c < C; c += LL_WARPS) {
			double v = 0.0;
#pragma unroll
			for (int w = 0; w < LL_WARPS; w++)
				v += sacc[c * LL_THREADS + w * 32 + lane];
			v = warp_sum(v);
			const int g = c0 + c < n_act ? (a.act_idx ? a.act_idx[c0 + c] : c0 + c) : -1;
			if (lane == 0 && g >= 0)
				a.partial[(size_t) g * a.n_splits + split] = v;
		}
		__syncthreads();
	}
}

constexpr size_t LL_SMEM_BYTES = sizeof(double2) * LL_STAGES * LL_CHUNK
		+ sizeof(double) * LL_MAX_C * LL_THREADS + sizeof(uint64_t) * 2 * LL_STAGES;
// partial sums one likelihood launch writes per (chain, row split).  (A barrier-free fold with one
// partial per warp was measured: 4.68 ms against 4.60 ms per C3 launch -- the CTA barrier at the
// end of an item re-aligns the warps on the ring, which is worth more than it costs.)
constexpr int LL_PARTS = 1;

// max |x| over the table (first column): bound for the per-item fast-sine range check.
// |double| ordering == unsigned ordering of the bit pattern; NaN compares above everything,
// which switches the fast path off.
__global__ void absmax_col0_kernel(const double * data, long long n_rows, int row_w, unsigned long long * out) {
	unsigned long long m = 0;
	for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n_rows;
			i += (long long) gridDim.x * blockDim.x) {
		unsigned long long b = (unsigned long long) __double_as_longlong(data[(size_t) row_w * i]) & 0x7fffffffffffffffull;
		m = b > m ? b : m;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
		m = t > m ? t : m;
	}
	if ((threadIdx.x & 31) == 0)
		atomicMax(out, m);
}

// ------------------------------------------------------------------ control kernel
enum {
	ADV_FINALIZE = 1, ADV_RECORD = 2, ADV_SWAP = 4, ADV_PROPOSE_RUN = 8, ADV_CALIB = 16, ADV_CALIB_BEGIN = 32
};

struct AdvArgs {
	int flags;
	long long step_index;
	CalibCfgDev cal;
	const unsigned char * select; // calibration selection (ADV_CALIB_BEGIN)
	int act_w;                    // calibration: which active-list buffer this launch fills
	const double * pack_prev;     // ladder split: neighbours' boundary chains (see ensemble_swap)
	const double * pack_next;
};

// calibration: append chain g to the list the next likelihood launch walks (the order of the
// list does not influence any result: a chain's sum does not depend on its tile)
APM_D void act_push(const DevState & S, int w, int g) {
	const int slot = atomicAdd(&S.act_n[w], 1);
	S.act_idx[(size_t) w * S.n_chains + slot] = g;
}

constexpr int ADV_THREADS = 128;

template<class M>
APM_D double chain_gather_sum(const DevState & S, int g) {
	// deterministic: the row splits are added in index order, after the model's initial value
	double sum = M::sum0(S.prop + (size_t) g * S.n_par);
	if (M::HAS_DATA)
		for (int s = 0; s < S.n_splits; s++)
			sum += S.partial[(size_t) g * S.n_splits + s];
	return sum;
}

template<class M>
__global__ void __launch_bounds__(ADV_THREADS) advance_kernel(const DevState S, const AdvArgs a) {
	const int ens = blockIdx.x;
	const int base = ens * S.n_beta;
	if (a.flags & ADV_CALIB_BEGIN) {
		for (int k = threadIdx.x; k < S.n_beta; k += blockDim.x) {
			const int g = base + k;
			S.pend[g] = PEND_NONE;
			if (a.select == nullptr || a.select[g]) {
				S.cal[g].phase = CAL_IDLE;
				atomicAdd(S.n_active, 1);
				cal_begin(S, g, a.cal);
				int kind = cal_next_kind(S, g);
				if (kind != PEND_NONE) {
					chain_propose(S, g, kind);
					act_push(S, a.act_w, g);
				}
			} else {
				S.cal[g].phase = CAL_IDLE;
				S.cal[g].status = -1;
			}
		}
		return;
	}
	if ((a.flags & ADV_CALIB) && blockIdx.x == 0 && threadIdx.x == 0)
		S.act_n[1 - a.act_w] = 0; // consumed by the likelihood launch before this one
	if (a.flags & ADV_FINALIZE) {
		for (int k = threadIdx.x; k < S.n_beta; k += blockDim.x) {
			const int g = base + k;
			if (S.pend[g] == PEND_NONE)
				continue;
			chain_finalize<M>(S, g, chain_gather_sum<M>(S, g));
			if (a.flags & ADV_RECORD)
				chain_record(S, g, a.step_index);
			if (a.flags & ADV_CALIB) {
				cal_after_step(S, g, a.cal);
				int kind = cal_next_kind(S, g);
				if (kind != PEND_NONE) {
					chain_propose(S, g, kind);
					act_push(S, a.act_w, g);
				}
			}
		}
	}
	if (a.flags & ADV_SWAP) {
		if (S.adapt) // adapt() precedes tempering_interaction (parallel_tempering.c:404-406)
			for (int k = threadIdx.x; k < S.n_beta; k += blockDim.x)
				chain_adapt(S, base + k);
		__syncthreads();
		if (threadIdx.x == 0)
			ensemble_swap(S, ens, a.pack_prev, a.pack_next);
	}
	if (a.flags & ADV_PROPOSE_RUN) {
		__syncthreads();
		for (int k = threadIdx.x; k < S.n_beta; k += blockDim.x)
			chain_propose(S, base + k, S.n_par);
	}
}

// ------------------------------------------------------------------ ladder split over GPUs
// swap-relevant state of every ensemble's first and last rung on this device, laid out for the
// neighbour exchange: [n_ens][LADDER_PACK(n_par)] each
__global__ void ladder_pack_kernel(const DevState S, double * first_pack, double * last_pack) {
	const int ens = blockIdx.x * blockDim.x + threadIdx.x;
	if (ens >= S.n_ens)
		return;
	const int n = S.n_par;
	for (int side = 0; side < 2; side++) {
		const int g = ens * S.n_beta + (side == 0 ? 0 : S.n_beta - 1);
		double * p = (side == 0 ? first_pack : last_pack) + (size_t) ens * LADDER_PACK(n);
		p[0] = S.prob[g];
		p[1] = S.beta[g];
		p[2] = S.prior[g];
		p[3] = S.prob_best[g];
		for (int i = 0; i < n; i++) {
			p[4 + i] = S.params[(size_t) g * n + i];
			p[4 + n + i] = S.params_best[(size_t) g * n + i];
		}
	}
}

// ------------------------------------------------------------------ fused small-table path
// One CTA per ensemble, the whole data table AND the ensemble's chain state resident in shared
// memory for the duration of the launch (one TMA bulk copy for the table, a cooperative copy
// in / copy out for the state).  A launch executes n_rounds x (n_swap Metropolis steps of every
// chain + the ensemble's swap), or complete calibrations, without returning to the host.
//   models with data: one warp per chain (warps loop when the ladder is longer than the CTA);
//     per step the warp evaluates its chain's likelihood over the table (lane-strided rows,
//     4 accumulators, fp64 butterfly), its lanes draw the coordinates of the next proposal in
//     parallel, and lane 0 runs the very same chain_* functions as the tiled path's control kernel;
//   data-free models (apps/normal.c): one thread per chain.
// The chain_* functions are used unchanged: they get a DevState whose pointers address the
// shared-memory copies, indexed by the chain's position in the ensemble; the id offsets carry the
// global numbering, so random streams, traces and progress rows are those of the tiled path.
constexpr int FUSED_MAX_WARPS = 16;
constexpr size_t FUSED_SMEM_LIMIT = 220 * 1024;

#define FUSED_ARRAYS(X) \
	X(params, double, n) X(params_best, double, n) X(steps, double, n) X(prop, double, n) \
	X(prob, double, 1) X(prior, double, 1) X(prob_best, double, 1) X(beta, double, 1) \
	X(accept, u64, 1) X(reject, u64, 1) X(n_iter, u64, 1) X(swapcount, u64, 1) X(rng_ctr, u64, 1) \
	X(pacc, u64, n) X(prej, u64, n) X(stat_n, u64, 1) X(stat_sum_dl, double, 1) \
	X(stat_sum_p, double, n) X(stat_sum_p2, double, n)

// bytes of shared memory for the state of one ensemble
__host__ __device__ inline size_t fused_state_bytes(int n_beta, int n_par) {
	const size_t per_chain = 8 * (size_t) (8 * n_par + 11) + sizeof(CalState) + sizeof(int);
	return (per_chain * n_beta + 8 /* swap_round */ + 15) & ~(size_t) 15;
}
// (n_slots = table rows x ROW_W / 2: 16-byte units)
// fused_run_kernel: per chain a batch of random draws (32 doubles) and its base counter
__host__ __device__ inline size_t fused_draws_bytes(int n_beta) {
	return (size_t) n_beta * (32 * sizeof(double) + sizeof(unsigned long long));
}
__host__ __device__ inline size_t fused_table_bytes(long long n_slots) {
	return (((size_t) n_slots * 16 + 127) & ~(size_t) 127) + 16 /* mbarrier */;
}

struct FusedArgs {
	const double * data;     // [n_rows][2]
	int n_rows;
	const double * xabsmax;
	long long n_rounds;      // run
	int n_swap;
	CalibCfgDev cal;         // calibrate
	const unsigned char * select;
};

template<class M>
__device__ __forceinline__ const Row<M> * fused_stage_table(const FusedArgs & a, unsigned char * smem) {
	Row<M> * sdata = reinterpret_cast<Row<M> *>(smem);
	uint64_t * bar = reinterpret_cast<uint64_t *>(smem + fused_table_bytes((long long) a.n_rows * (M::ROW_W / 2)) - 16);
	if (a.n_rows > 0) {
		if (threadIdx.x == 0) {
			mbar_init(bar, 1);
			mbar_fence_init();
			// bulk copies of at most 32 KB each, all completing on the one barrier
			const uint32_t total = (uint32_t) a.n_rows * sizeof(Row<M>);
			mbar_arrive_expect_tx(bar, total);
			for (uint32_t off = 0; off < total; off += 32768u) {
				const uint32_t n = min(32768u, total - off);
				tma_bulk_g2s(reinterpret_cast<unsigned char *>(sdata) + off,
						reinterpret_cast<const unsigned char *>(a.data) + off, n, bar);
			}
		}
		__syncthreads();
		mbar_wait(bar, 0);
	}
	return sdata;
}

// copy rungs [k0, k0 + nb) of ensemble `ens` into shared memory (the whole ensemble: k0 = 0, nb =
// S.n_beta); the returned DevState addresses the copies with chain index = position in the block
// and carries the ladder-split fields (n_beta_total, k_offset) that make chain ids, swap pairs
// and trace slots those of the whole ladder
__device__ inline DevState fused_localize_block(const DevState & S, int ens, int k0, int nb, unsigned char * mem) {
	DevState L = S;
	const int n = S.n_par;
	const size_t base = (size_t) ens * S.n_beta + k0;
	size_t off = 0;
#define X(name, type, width) { \
		type * p = reinterpret_cast<type *>(mem + off); \
		for (int i = threadIdx.x; i < nb * (width); i += blockDim.x) \
			p[i] = S.name[base * (width) + i]; \
		L.name = p; \
		off += sizeof(type) * (size_t) nb * (width); }
	FUSED_ARRAYS(X)
#undef X
	{
		constexpr int W = sizeof(CalState) / 8;
		u64 * p = reinterpret_cast<u64 *>(mem + off);
		const u64 * src = reinterpret_cast<const u64 *>(S.cal + base);
		for (int i = threadIdx.x; i < nb * W; i += blockDim.x)
			p[i] = src[i];
		L.cal = reinterpret_cast<CalState *>(p);
		off += sizeof(CalState) * (size_t) nb;
	}
	{
		u64 * p = reinterpret_cast<u64 *>(mem + off);
		if (threadIdx.x == 0)
			p[0] = S.swap_round[ens];
		L.swap_round = p;
		off += 8;
	}
	{
		int * p = reinterpret_cast<int *>(mem + off);
		for (int i = threadIdx.x; i < nb; i += blockDim.x)
			p[i] = S.pend[base + i];
		L.pend = p;
	}
	L.n_ens = 1;
	L.n_beta = nb;
	L.n_beta_total = S.n_beta_total;
	L.id_stride = S.id_stride;
	L.k_offset = S.k_offset + k0;
	L.g_base = S.g_base + (int) base;
	L.chain_id_offset = S.chain_id_offset + ens * S.id_stride;
	L.ensemble_id_offset = S.ensemble_id_offset + ens;
	if (S.tr_prob != nullptr) {
		L.tr_prob = S.tr_prob + base;
		L.tr_dl = S.tr_dl + base;
	}
	if (S.tr_params != nullptr)
		L.tr_params = S.tr_params + (S.tr_params_chains == 2 ? base : (size_t) ens) * n;
	__syncthreads();
	return L;
}

__device__ inline DevState fused_localize(const DevState & S, int ens, unsigned char * mem) {
	return fused_localize_block(S, ens, 0, S.n_beta, mem);
}

// L = what fused_localize_block returned for (ens, k0); copies its L.n_beta chains back
__device__ inline void fused_writeback_block(const DevState & S, const DevState & L, int ens, int k0,
		bool write_swap_round) {
	const int nb = L.n_beta, n = S.n_par;
	const size_t base = (size_t) ens * S.n_beta + k0;
	__syncthreads();
#define X(name, type, width) \
		for (int i = threadIdx.x; i < nb * (width); i += blockDim.x) \
			S.name[base * (width) + i] = L.name[i];
	FUSED_ARRAYS(X)
#undef X
	{
		constexpr int W = sizeof(CalState) / 8;
		u64 * dst = reinterpret_cast<u64 *>(S.cal + base);
		const u64 * src = reinterpret_cast<const u64 *>(L.cal);
		for (int i = threadIdx.x; i < nb * W; i += blockDim.x)
			dst[i] = src[i];
	}
	if (threadIdx.x == 0 && write_swap_round)
		S.swap_round[ens] = L.swap_round[0];
	for (int i = threadIdx.x; i < nb; i += blockDim.x)
		S.pend[base + i] = L.pend[i];
}

__device__ inline void fused_writeback(const DevState & S, const DevState & L, int ens) {
	fused_writeback_block(S, L, ens, 0, true);
}

// sum over the table of the model's row terms for chain g's pending proposal; every lane
// returns the same bits (butterfly of commutative adds)
template<class M>
__device__ __forceinline__ double fused_loglik(const DevState & S, int g, const Row<M> * sdata, int n_rows,
		double xub, int lane) {
	if (!M::HAS_DATA)
		return 0.0;
	typename M::Prep q;
	M::prep(q, S.prop + (size_t) g * S.n_par, S.n_par, S.model_const);
	Acc<M> a0 = ModelAcc<M>::zero(), a1 = a0, a2 = a0, a3 = a0;
	int i = lane;
	if (M::fast_ok(q, xub)) {
		for (; i + 96 < n_rows; i += 128) {
			const Row<M> r0 = sdata[i], r1 = sdata[i + 32], r2 = sdata[i + 64], r3 = sdata[i + 96];
			a0 = row_accum_fast<M>(a0, q, r0);
			a1 = row_accum_fast<M>(a1, q, r1);
			a2 = row_accum_fast<M>(a2, q, r2);
			a3 = row_accum_fast<M>(a3, q, r3);
		}
		for (; i < n_rows; i += 32) {
			const Row<M> r = sdata[i];
			a0 = row_accum_fast<M>(a0, q, r);
		}
	} else {
		for (; i < n_rows; i += 32) {
			const Row<M> r = sdata[i];
			a0 = row_accum<M>(a0, q, r);
		}
	}
	return warp_sum(ModelAcc<M>::value(ModelAcc<M>::merge(ModelAcc<M>::merge(a0, a1), ModelAcc<M>::merge(a2, a3))));
}

// chain_propose with the coordinates drawn by the lanes of a warp in parallel (every draw has
// its own counter, so who computes it is immaterial)
__device__ __forceinline__ void chain_propose_warp(const DevState & S, int g, int kind, int lane) {
	const int n = S.n_par;
	if (lane < n) {
		const double old = S.params[(size_t) g * n + lane];
		S.prop[(size_t) g * n + lane] = (kind == n || kind == lane)
				? propose_coordinate(S, g, S.rng_ctr[g], lane, old, S.steps[(size_t) g * n + lane]) : old;
	}
	if (lane == 0)
		S.pend[g] = kind;
	__syncwarp();
}

template<class M>
__global__ void __launch_bounds__(FUSED_MAX_WARPS * 32, 1) fused_run_kernel(const DevState S, const FusedArgs a) {
	extern __shared__ __align__(128) unsigned char fused_smem[];
	const Row<M> * sdata = fused_stage_table<M>(a, fused_smem);
	const int ens = blockIdx.x;
	const DevState L = fused_localize(S, ens, fused_smem + fused_table_bytes((long long) a.n_rows * (M::ROW_W / 2)));
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
	const int nb = L.n_beta, n = L.n_par;
	const double xub = M::HAS_DATA ? *a.xabsmax : 0.0;
	long long step = 0;
	if (M::HAS_DATA) {
		// per chain: a batch of K steps' random draws (lane-parallel, chain_draw_batch) and its base counter
		const int K = 32 / (n + 1);
		double * draws = reinterpret_cast<double *>(fused_smem + fused_table_bytes((long long) a.n_rows * (M::ROW_W / 2))
				+ fused_state_bytes(nb, n));
		u64 * draw_base = reinterpret_cast<u64 *>(draws + (size_t) nb * 32);
		for (long long round = 0; round < a.n_rounds; round++) {
			for (int k = warp; k < nb; k += n_warps) {
				u64 base;
				chain_first_proposal_warp(L, k, draws + (size_t) k * 32, base, K, lane);
				if (lane == 0)
					draw_base[k] = base;
			}
			__syncwarp();
			for (int sub = 0; sub < a.n_swap; sub++, step++) {
				const bool last_of_round = sub + 1 == a.n_swap;
				for (int k = warp; k < nb; k += n_warps) {
					const double sum = fused_loglik<M>(L, k, sdata, a.n_rows, xub, lane);
					u64 base = draw_base[k];
					chain_step_tail_warp<M>(L, k, M::sum0(L.prop + (size_t) k * n) + sum, draws + (size_t) k * 32, base, K,
							!last_of_round, lane);
					if (lane == 0)
						draw_base[k] = base;
					chain_record_warp(L, k, step, lane);
				}
			}
			// adapt (if compiled in), tempering_interaction for this ensemble
			__syncthreads();
			if (L.adapt)
				for (int k = threadIdx.x; k < nb; k += blockDim.x)
					chain_adapt(L, k);
			__syncthreads();
			if (threadIdx.x == 0)
				ensemble_swap(L, 0);
			__syncthreads();
		}
	} else {
		for (int k = threadIdx.x; k < nb; k += blockDim.x)
			chain_propose(L, k, n);
		for (long long round = 0; round < a.n_rounds; round++) {
			for (int sub = 0; sub < a.n_swap; sub++, step++) {
				for (int k = threadIdx.x; k < nb; k += blockDim.x) {
					chain_finalize<M>(L, k, M::sum0(L.prop + (size_t) k * n));
					chain_record(L, k, step);
					if (sub + 1 < a.n_swap)
						chain_propose(L, k, n);
				}
			}
			__syncthreads();
			if (L.adapt)
				for (int k = threadIdx.x; k < nb; k += blockDim.x)
					chain_adapt(L, k);
			__syncthreads();
			if (threadIdx.x == 0)
				ensemble_swap(L, 0);
			__syncthreads();
			if (round + 1 < a.n_rounds)
				for (int k = threadIdx.x; k < nb; k += blockDim.x)
					chain_propose(L, k, n);
		}
	}
	fused_writeback(S, L, ens);
}

// ------------------------------------------------------------------ cluster path
// The fused path with one ensemble spread over a thread-block CLUSTER of CL CTAs (CL SMs), for
// runs with fewer ensembles than SMs (config C1: ONE 20-rung ensemble): CTA r of the cluster
// holds rungs [nb r / CL, nb (r + 1) / CL) -- the ladder split of DESIGN.md section 6, inside
// one GPC instead of across GPUs -- and a group of WC warps works on each of its chains.
//   * the table is fetched from global memory ONCE per cluster: CTA 0 issues TMA bulk copies
//     with .multicast::cluster, which land at the same shared-memory offset of every CTA and
//     complete on every CTA's own mbarrier;
//   * a Metropolis step is latency-bound (rows -> sum -> accept -> next proposal -> rows), so the
//     group keeps everything that is not on that chain of dependencies off it: WC - 1 warps walk
//     the table (rows strided over their lanes, fp64 butterfly per warp, ONE named barrier per
//     step); then EVERY warp adds the warps' sums in index order and takes the accept decision
//     and forms the next proposal redundantly in registers (lane i holds coordinate i) -- the
//     same instructions on the same inputs, hence the same bits in every warp -- so nobody waits
//     for a leader; the group's last warp, which walks no rows, writes the outcome into the
//     chain state (counters, best, trace rows, accumulators) one step behind, and draws the
//     random numbers of the next K = 32 / (n_par + 1) steps in one lane-parallel batch (a
//     chain's draws depend on its id and step counter only);
//   * once per round the CTAs publish the swap-relevant state of their first and last rung
//     ("pack", LADDER_PACK doubles) in their shared memory, barrier.cluster, and read their
//     neighbours' packs through distributed shared memory: ensemble_swap decides a pair that
//     straddles two CTAs identically on both, each updating the chain it owns -- the very code
//     path of the multi-GPU ladder split, so results equal the other paths' chain for chain.
constexpr int CLUSTER_MAX = 8; // the portable cluster size (16 was measured: +10 % for one ensemble, -40 % for eight)

__host__ __device__ inline int cluster_block_lo(int nb, int cl, int r) { return (int) ((long long) nb * r / cl); }

struct ClusterArgs {
	long long * timing; // -DAPM_CLUSTER_TIMING builds: [16 warps][8] cycle sums of CTA 0 (tools/prof_small.py)
	FusedArgs f;
	int cl;      // CTAs per ensemble = cluster size
	int gmax;    // chains per CTA (upper bound): ceil(nb / cl)
	int wc;      // warps per chain
};

constexpr int CLUSTER_DRAW_RING = 4; // batches of draws in flight per chain
__host__ __device__ inline size_t cluster_smem_bytes(long long n_slots, int gmax, int wc, int n_par) {
	// packs (2) + per chain: warp sums [2][wc], draws [RING][32], per-warp proposal copies [wc][n_par]
	return fused_table_bytes(n_slots) + fused_state_bytes(gmax, n_par)
			+ sizeof(double) * (2 * LADDER_PACK(n_par)
					+ (size_t) gmax * (2 * wc + CLUSTER_DRAW_RING * 32 + (size_t) wc * n_par));
}

__device__ __forceinline__ void group_bar(int id, int n_threads) {
	asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n_threads) : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s_multicast(void * dst_smem, const void * src_gmem, uint32_t bytes,
		uint64_t * bar, uint16_t cta_mask) {
	asm volatile(
			"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
			:: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
// barrier.cluster.arrive.release + barrier.cluster.wait.acquire over all threads of the cluster
__device__ __forceinline__ void cluster_sync_all() {
	cooperative_groups::this_cluster().sync();
}
// generic (distributed shared memory) address of `p`, a shared-memory address of this CTA, in
// CTA `rank` of the cluster (mapa)
__device__ __forceinline__ const double * cluster_map(double * p, uint32_t rank) {
	return cooperative_groups::this_cluster().map_shared_rank(p, rank);
}

// the group's share of chain g's sum: lanes gl, gl + GL, ... of the table; every lane of a warp
// returns the warp's total
template<class M>
__device__ __forceinline__ double group_loglik(const DevState & S, const double * prop, const Row<M> * sdata,
		int n_rows, double xub, int gl, int GL) {
	typename M::Prep q;
	M::prep(q, prop, S.n_par, S.model_const);
	Acc<M> a0 = ModelAcc<M>::zero(), a1 = a0, a2 = a0, a3 = a0;
	int i = gl;
	if (M::fast_ok(q, xub)) {
		for (; i + 3 * GL < n_rows; i += 4 * GL) {
			const Row<M> r0 = sdata[i], r1 = sdata[i + GL], r2 = sdata[i + 2 * GL], r3 = sdata[i + 3 * GL];
			a0 = row_accum_fast<M>(a0, q, r0);
			a1 = row_accum_fast<M>(a1, q, r1);
			a2 = row_accum_fast<M>(a2, q, r2);
			a3 = row_accum_fast<M>(a3, q, r3);
		}
		// up to three rows left for this lane: evaluated side by side (a row's evaluation is one
		// long dependency chain), each into the accumulator it would have gone to above
		if (i < n_rows) {
			const bool v1 = i + GL < n_rows, v2 = i + 2 * GL < n_rows;
			const Row<M> r0 = sdata[i], r1 = sdata[v1 ? i + GL : i], r2 = sdata[v2 ? i + 2 * GL : i];
			const Acc<M> b0 = row_accum_fast<M>(a0, q, r0);
			const Acc<M> b1 = row_accum_fast<M>(a1, q, r1);
			const Acc<M> b2 = row_accum_fast<M>(a2, q, r2);
			a0 = b0;
			a1 = v1 ? b1 : a1;
			a2 = v2 ? b2 : a2;
		}
	} else {
		for (; i < n_rows; i += GL) {
			const Row<M> r = sdata[i];
			a0 = row_accum<M>(a0, q, r);
		}
	}
	return warp_sum(ModelAcc<M>::value(ModelAcc<M>::merge(ModelAcc<M>::merge(a0, a1), ModelAcc<M>::merge(a2, a3))));
}

#ifdef APM_CLUSTER_TIMING
#define APM_TICK(k) do { long long t_now; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_now) :: "memory"); \
		t_sum[k] += t_now - t_last; t_last = t_now; } while (0)
// the same after `val` has been computed (the clock read cannot be scheduled ahead of it)
#define APM_TICK_AFTER(k, val) do { asm volatile("" :: "d"(val) : "memory"); APM_TICK(k); } while (0)
#else
#define APM_TICK(k) do { } while (0)
#define APM_TICK_AFTER(k, val) do { } while (0)
#endif

template<class M>
__global__ void __launch_bounds__(FUSED_MAX_WARPS * 32, 1) cluster_run_kernel(const DevState S, const ClusterArgs ca) {
	extern __shared__ __align__(128) unsigned char fused_smem[];
#ifdef APM_CLUSTER_TIMING
	long long t_sum[4] = { 0, 0, 0, 0 }, t_last = clock64();
#endif
	const FusedArgs & a = ca.f;
	const int CL = ca.cl, WC = ca.wc;
	const uint32_t rank = cluster_ctarank();
	const int ens = blockIdx.x / CL;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nb_all = S.n_beta, n = S.n_par;
	const int k0 = cluster_block_lo(nb_all, CL, rank), nloc = cluster_block_lo(nb_all, CL, rank + 1) - k0;

	// ---- the table: one multicast fetch per cluster
	const long long n_slots = (long long) a.n_rows * (M::ROW_W / 2);
	Row<M> * sdata = reinterpret_cast<Row<M> *>(fused_smem);
	uint64_t * bar = reinterpret_cast<uint64_t *>(fused_smem + fused_table_bytes(n_slots) - 16);
	const uint32_t total = (uint32_t) a.n_rows * sizeof(Row<M>);
	if (tid == 0) {
		mbar_init(bar, 1);
		mbar_fence_init();
		mbar_arrive_expect_tx(bar, total);
	}
	cluster_sync_all(); // every CTA's barrier is armed before the copies that complete on it start
	if (rank == 0 && tid == 0) {
		for (uint32_t off = 0; off < total; off += 32768u)
			tma_bulk_g2s_multicast(reinterpret_cast<unsigned char *>(sdata) + off,
					reinterpret_cast<const unsigned char *>(a.data) + off, min(32768u, total - off), bar,
					(uint16_t) ((1u << CL) - 1u));
	}
	mbar_wait(bar, 0);

	// ---- this CTA's block of the ensemble, resident in shared memory
	// (the localized DevState -- some forty pointers -- lives in shared memory, not in every thread's
	// registers: the row loop needs the registers for independent row evaluations in flight)
	unsigned char * mem = fused_smem + fused_table_bytes(n_slots);
	__shared__ DevState L_shared;
	{
		const DevState L_tmp = fused_localize_block(S, ens, k0, nloc, mem);
		if (tid == 0)
			L_shared = L_tmp;
		__syncthreads();
	}
	const DevState & L = L_shared;
	double * pack_first = reinterpret_cast<double *>(mem + fused_state_bytes(ca.gmax, n));
	double * pack_last = pack_first + LADDER_PACK(n);
	double * red = pack_last + LADDER_PACK(n); // per-chain work areas, see cluster_smem_bytes
	const double * pack_prev = rank > 0 ? cluster_map(pack_last, rank - 1) : nullptr;
	const double * pack_next = rank + 1 < (uint32_t) CL ? cluster_map(pack_first, rank + 1) : nullptr;

	const int c = warp / WC, wi = warp - c * WC; // chain of this warp's group, warp within the group
	const bool active = c < nloc;
	double * sums = red + (size_t) c * 2 * WC;                                      // [2][WC]
	double * draws = red + (size_t) ca.gmax * 2 * WC + (size_t) c * CLUSTER_DRAW_RING * 32; // [RING][32]
	double * wprop = red + (size_t) ca.gmax * (2 * WC + CLUSTER_DRAW_RING * 32) + ((size_t) c * WC + wi) * n; // [n]
	const int LW = WC > 1 ? WC - 1 : 1;    // warps of the group that walk the table
	const bool service = WC > 1 && wi == WC - 1;
	const int GL = LW * 32, gl = wi * 32 + lane;
	const int GT = WC * 32;                // threads of the group (named barrier)
	const int K = 32 / (n + 1);            // steps per batch of draws (n <= 16: K >= 1)
	const double xub = *a.xabsmax;
	const double lo = lane < n ? L.pmin[lane] : 0.0, hi = lane < n ? L.pmax[lane] : 0.0;
	long long step = 0;
	if (active && wi == 0)
		chain_propose_warp(L, c, n, lane);
	__syncthreads();
	for (long long round = 0; round < a.n_rounds; round++) {
		if (active && WC == 1) {
			// one warp per chain: the fused path's step, on this CTA's block of the ladder
			for (int sub = 0; sub < a.n_swap; sub++, step++) {
				const double sum = group_loglik<M>(L, L.prop + (size_t) c * n, sdata, a.n_rows, xub, gl, GL);
				chain_finalize_warp<M>(L, c, M::sum0(L.prop + (size_t) c * n) + sum, nullptr, lane);
				chain_record_warp(L, c, step, lane);
				if (sub + 1 < a.n_swap)
					chain_propose_warp(L, c, n, lane);
			}
		} else if (active) {
			// the chain's state as of the round's start, replicated in the registers of every warp
			const u64 ctr0 = L.rng_ctr[c];
			const double beta = L.beta[c];
			const double stepw = lane < n ? L.steps[(size_t) c * n + lane] : 0.0;
			double cur = lane < n ? L.params[(size_t) c * n + lane] : 0.0;
			double prop = lane < n ? L.prop[(size_t) c * n + lane] : 0.0;
			double prob_cur = L.prob[c], prior_cur = L.prior[c];
			const double mc[4] = { L.model_const[0], L.model_const[1], L.model_const[2], L.model_const[3] };
			const unsigned quirks = L.quirks;
			int accepted = 0;
			double prob_new = 0.0, prior_new = 0.0;
			// position of step `sub`'s draws in the ring: batch sub / K, entry sub % K (kept incrementally)
			int ring_b = 0, ring_j = 0;

			// markov_chain_step's second half for step `sub` -- the same instructions in every warp of
			// the group -- and the next proposal (do_step): first attempt from the batch of draws,
			// the rare out-of-bounds rest as usual.  Leaves (accepted, prob_new, prior_new) of this
			// step and updates (cur, prob_cur, prior_cur, prop).
			auto decide = [&](int sub, double & prop_done) {
				const int par = sub & 1;
				double sum = sums[par * WC];
				for (int w = 1; w < LW; w++)
					sum += sums[par * WC + w];
				prior_new = prior_cur;
				if (M::HAS_PRIOR)
					prior_new = M::prior(wprop, n, mc);
				prob_new = M::finish(beta, M::sum0(wprop) + sum, prior_new, wprop, mc);
				const double * dr = draws + ring_b * 32 + ring_j * (n + 1);
				if (prob_new == prob_cur)
					accepted = 1;
				else if (prob_new > prob_cur)
					accepted = 1;
				else
					accepted = dr[n] < (prob_new - prob_cur) ? 1 : 0;
				prop_done = prop;
				if (accepted) {
					cur = prop;
					prob_cur = prob_new;
					prior_cur = prior_new;
				} else if (quirks & 2u) {
					prior_cur = prior_new;
				}
				if (++ring_j == K) {
					ring_j = 0;
					ring_b = (ring_b + 1) % CLUSTER_DRAW_RING;
				}
				if (sub + 1 < a.n_swap && lane < n) {
					double v = cur + draws[ring_b * 32 + ring_j * (n + 1) + lane];
					if (v > hi || v < lo)
						v = propose_coordinate(L, c, ctr0 + (u64) (sub + 1), lane, cur, stepw);
					prop = v;
				}
			};

			if (service) {
				// a private copy of the state's addresses: the bookkeeping below is chains of
				// read-modify-writes, which must not wait for the pointers to be re-read after every store
				const DevState Lr = L;
				int pend_acc = 0;
				double pend_prob = 0.0, pend_prior = 0.0, pend_prop = 0.0;
				int next_batch_at = K - 1, next_batch = 1; // batch b is drawn at the start of step b K - 1
				for (int sub = 0; sub < a.n_swap; sub++, step++) {
					APM_TICK(0);
					if (sub > 0) { // write down step sub - 1
						chain_apply_step_warp(Lr, c, pend_acc, pend_prob, pend_prior, pend_prop, lane);
						chain_record_warp(Lr, c, step - 1, lane);
					}
					if (sub == 0)
						chain_draw_batch(Lr, c, ctr0, K, lane, draws);
					if (sub == next_batch_at) {
						if (sub + 1 < a.n_swap)
							chain_draw_batch(Lr, c, ctr0 + (u64) next_batch * K, K, lane,
									draws + (next_batch % CLUSTER_DRAW_RING) * 32);
						next_batch_at += K;
						next_batch++;
					}
					if (lane < n)
						wprop[lane] = prop;
					__syncwarp();
					APM_TICK(1);
					group_bar(1 + c, GT);
					APM_TICK(2);
					decide(sub, pend_prop);
					pend_acc = accepted;
					pend_prob = prob_new;
					pend_prior = prior_new;
					APM_TICK_AFTER(3, prop);
				}
				// the round's last step
				chain_apply_step_warp(Lr, c, pend_acc, pend_prob, pend_prior, pend_prop, lane);
				chain_record_warp(Lr, c, step - 1, lane);
				if (lane == 0)
					Lr.pend[c] = PEND_NONE;
			} else {
				double unused;
				for (int sub = 0; sub < a.n_swap; sub++, step++) {
					APM_TICK(0);
					if (lane < n)
						wprop[lane] = prop;
					__syncwarp();
					const double part = group_loglik<M>(L, wprop, sdata, a.n_rows, xub, gl, GL);
					if (lane == 0)
						sums[(sub & 1) * WC + wi] = part;
					APM_TICK_AFTER(1, part);
					group_bar(1 + c, GT);
					APM_TICK(2);
					decide(sub, unused);
					APM_TICK_AFTER(3, prop);
				}
			}
		} else {
			step += a.n_swap;
		}
		// ---- round end: adapt, publish the boundary rungs, swap with the neighbours' packs at hand
		__syncthreads();
		if (L.adapt && tid < nloc)
			chain_adapt(L, tid);
		if (tid < 2) {
			const int g = tid == 0 ? 0 : nloc - 1;
			double * p = tid == 0 ? pack_first : pack_last;
			p[0] = L.prob[g];
			p[1] = L.beta[g];
			p[2] = L.prior[g];
			p[3] = L.prob_best[g];
			for (int i = 0; i < n; i++) {
				p[4 + i] = L.params[(size_t) g * n + i];
				p[4 + n + i] = L.params_best[(size_t) g * n + i];
			}
		}
		cluster_sync_all();
		if (tid == 0)
			ensemble_swap(L, 0, pack_prev, pack_next);
		cluster_sync_all(); // the neighbours have read this CTA's packs; the swap is visible to the CTA
		if (round + 1 < a.n_rounds && active && wi == 0)
			chain_propose_warp(L, c, n, lane);
		__syncthreads();
	}
#ifdef APM_CLUSTER_TIMING
	if (ca.timing != nullptr && blockIdx.x == 0 && lane == 0)
		for (int k = 0; k < 4; k++)
			ca.timing[warp * 8 + k] = t_sum[k];
#endif
	fused_writeback_block(S, L, ens, k0, rank == 0);
}

// ------------------------------------------------------------------ grid path
// Tables too large for one SM's shared memory but far too small to keep the GPU busy for long
// (10^4 .. 10^6 rows, a few ensembles): the tiled path spends ~28 us per Metropolis step there,
// almost all of it launch gaps and fixed kernel latencies.  The grid path is ONE cooperative
// launch per apm_gpu_run with the table PARTITIONED over the shared memories of all SMs:
//   * CTA b keeps rows [n_rows b / G, n_rows (b + 1) / G) resident for the whole run (one TMA
//     bulk copy at the start; 148 x ~200 KB holds ~1.8 M two-column rows);
//   * chain c is OWNED by CTA c mod G: its whole state lives in that CTA's shared memory for the
//     run (fused_localize_block, as on the fused path), and one of its warps plays the chain's
//     markov_chain_step;
//   * per step: every CTA evaluates ALL chains' proposals on its slice (a warp per chain, 4
//     independent row evaluations in flight per lane) and publishes one partial sum per chain;
//     grid barrier; the owner adds the G partials in a fixed order (lane-strided, then an fp64
//     butterfly), finalises the step, draws the next proposal (random draws in lane-parallel
//     batches of K steps) and publishes it; grid barrier; everybody fetches the new proposals
//     (one coalesced read) while the owner does the step's bookkeeping (best, trace rows,
//     accumulators) in its shared memory;
//   * once per round the owners write their chains back, CTA 0 runs adapt + ensemble_swap on the
//     global state, and the owners reload.
constexpr int GRID_THREADS = 256, GRID_WARPS = GRID_THREADS / 32;
constexpr int GRID_MAX_OWNED = 8;                // chains per CTA: one per warp
constexpr int GRID_MAX_CHAINS_PER_SM = GRID_MAX_OWNED;
constexpr int GRID_MAX_CHAINS = 2048;

struct GridArgs {
	const double * data;
	long long n_rows;
	const double * xabsmax;
	long long n_rounds;
	int n_swap;
	double * partials;  // [n_chains][G]
	double * props;     // [n_chains][n_par]: the pending proposals, published by the owners
	int max_slice_rows; // shared-memory room for the slice (host-computed)
};

// shared memory of one CTA after the slice: all chains' proposals, and per owned chain its
// localized state and a batch of draws
__host__ __device__ inline size_t grid_state_bytes(int n_chains, int n_par) {
	return (size_t) n_chains * n_par * sizeof(double)
			+ (size_t) GRID_MAX_OWNED * (fused_state_bytes(1, n_par) + 32 * sizeof(double)) + 64;
}

template<class M>
__global__ void __launch_bounds__(GRID_THREADS, 1) grid_run_kernel(const DevState S, const GridArgs a) {
	extern __shared__ __align__(128) unsigned char grid_smem[];
	namespace cg = cooperative_groups;
	cg::grid_group grid = cg::this_grid();
	const int G = gridDim.x, b = blockIdx.x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n = S.n_par, NC = S.n_chains;
	const int K = 32 / (n + 1);

	// ---- this CTA's slice of the table, resident for the whole run
	const long long r0 = a.n_rows * b / G, r1 = a.n_rows * (b + 1) / G;
	const int n_slice = (int) (r1 - r0);
	Row<M> * sdata = reinterpret_cast<Row<M> *>(grid_smem);
	const size_t slice_bytes = (((size_t) a.max_slice_rows * sizeof(Row<M>) + 127) & ~(size_t) 127);
	uint64_t * bar = reinterpret_cast<uint64_t *>(grid_smem + slice_bytes);
	if (n_slice > 0) {
		if (tid == 0) {
			mbar_init(bar, 1);
			mbar_fence_init();
			const uint32_t total = (uint32_t) n_slice * sizeof(Row<M>);
			mbar_arrive_expect_tx(bar, total);
			const unsigned char * src = reinterpret_cast<const unsigned char *>(a.data) + (size_t) r0 * sizeof(Row<M>);
			for (uint32_t off = 0; off < total; off += 32768u)
				tma_bulk_g2s(reinterpret_cast<unsigned char *>(sdata) + off, src + off, min(32768u, total - off), bar);
		}
		__syncthreads();
		mbar_wait(bar, 0);
	}
	double * props = reinterpret_cast<double *>(grid_smem + slice_bytes + 64); // [NC][n]
	unsigned char * own_mem = reinterpret_cast<unsigned char *>(props + (size_t) NC * n);
	const size_t own_stride = fused_state_bytes(1, n) + 32 * sizeof(double);
	__shared__ DevState L_own[GRID_MAX_OWNED];

	// the chain this warp plays (warp w of CTA b owns chain b + w G), if any
	const int my_chain = b + warp * G;
	const bool owner = my_chain < NC;
	unsigned char * my_mem = own_mem + (size_t) warp * own_stride;
	double * my_draws = reinterpret_cast<double *>(my_mem + fused_state_bytes(1, n)); // [K][n + 1]
	const double xub = *a.xabsmax;
	long long step = 0;

	for (long long round = 0; round < a.n_rounds; round++) {
		// ---- round start: the owners take their chains into shared memory and publish the round's
		// first proposals.  (fused_localize_block is written for a whole CTA; one chain is tiny,
		// so the warps go through it one after the other.)
		for (int w = 0; w < GRID_WARPS; w++) {
			const int c = b + w * G;
			if (c < NC) { // uniform over the CTA
				const DevState L_tmp = fused_localize_block(S, c / S.n_beta, c % S.n_beta, 1, own_mem + (size_t) w * own_stride);
				if (tid == 0)
					L_own[w] = L_tmp;
				__syncthreads();
			}
		}
		u64 draw_base = 0;
		if (owner) {
			const DevState & L = L_own[warp];
			chain_first_proposal_warp(L, 0, my_draws, draw_base, K, lane);
			if (lane < n)
				a.props[(size_t) my_chain * n + lane] = L.prop[lane];
		}
		__syncthreads();
		grid.sync();
		for (int i = tid; i < NC * n; i += GRID_THREADS)
			props[i] = __ldcg(a.props + i);
		__syncthreads();

		for (int sub = 0; sub < a.n_swap; sub++, step++) {
			// ---- every chain's proposal on this CTA's slice: a warp per chain
			for (int c = warp; c < NC; c += GRID_WARPS) {
				const double v = n_slice > 0 ? group_loglik<M>(S, props + (size_t) c * n, sdata, n_slice, xub, lane, 32) : 0.0;
				if (lane == 0)
					a.partials[(size_t) c * G + b] = v;
			}
			__syncthreads();
			grid.sync();
			// ---- the owners: markov_chain_step's second half and the next proposal
			if (owner) {
				const DevState & L = L_own[warp];
				// the G partial sums in a fixed order: lane l adds l, l + 32, ..., then a butterfly
				double sum = 0.0;
				for (int k = lane; k < G; k += 32)
					sum += __ldcg(a.partials + (size_t) my_chain * G + k);
				sum = warp_sum(sum);
				chain_step_tail_warp<M>(L, 0, M::sum0(L.prop) + sum, my_draws, draw_base, K, sub + 1 < a.n_swap, lane);
				if (sub + 1 < a.n_swap && lane < n)
					a.props[(size_t) my_chain * n + lane] = L.prop[lane];
			}
			__syncthreads();
			grid.sync();
			// ---- everybody fetches the proposals; the owners do the step's bookkeeping meanwhile
			if (sub + 1 < a.n_swap)
				for (int i = tid; i < NC * n; i += GRID_THREADS)
					props[i] = __ldcg(a.props + i);
			if (owner)
				chain_record_warp(L_own[warp], 0, step, lane);
			__syncthreads();
		}
		// ---- round end: write the chains back, adapt + swap on the global state
		for (int w = 0; w < GRID_WARPS; w++) {
			const int c = b + w * G;
			if (c < NC)
				fused_writeback_block(S, L_own[w], c / S.n_beta, c % S.n_beta, false);
		}
		__syncthreads();
		grid.sync();
		if (b == 0) {
			if (S.adapt)
				for (int c = tid; c < NC; c += GRID_THREADS)
					chain_adapt(S, c);
			__syncthreads();
			for (int e = tid; e < S.n_ens; e += GRID_THREADS)
				ensemble_swap(S, e);
		}
		__syncthreads();
		grid.sync();
	}
}

// markov_chain_calibrate (or apm_gpu_steps) of the selected chains on the grid path: the owners run
// the calibration state machine of apm_chain.cuh for their chains, chain by chain at its own pace;
// a chain that is done (or not selected) simply stops publishing proposals.  `active` is a global
// flag per chain, written by the owner together with the proposal.
template<class M>
__global__ void __launch_bounds__(GRID_THREADS, 1) grid_calibrate_kernel(const DevState S, const GridArgs a,
		const CalibCfgDev cal, const unsigned char * select, int * active) {
	extern __shared__ __align__(128) unsigned char grid_smem[];
	namespace cg = cooperative_groups;
	cg::grid_group grid = cg::this_grid();
	const int G = gridDim.x, b = blockIdx.x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n = S.n_par, NC = S.n_chains;

	const long long r0 = a.n_rows * b / G, r1 = a.n_rows * (b + 1) / G;
	const int n_slice = (int) (r1 - r0);
	Row<M> * sdata = reinterpret_cast<Row<M> *>(grid_smem);
	const size_t slice_bytes = (((size_t) a.max_slice_rows * sizeof(Row<M>) + 127) & ~(size_t) 127);
	uint64_t * bar = reinterpret_cast<uint64_t *>(grid_smem + slice_bytes);
	if (n_slice > 0) {
		if (tid == 0) {
			mbar_init(bar, 1);
			mbar_fence_init();
			const uint32_t total = (uint32_t) n_slice * sizeof(Row<M>);
			mbar_arrive_expect_tx(bar, total);
			const unsigned char * src = reinterpret_cast<const unsigned char *>(a.data) + (size_t) r0 * sizeof(Row<M>);
			for (uint32_t off = 0; off < total; off += 32768u)
				tma_bulk_g2s(reinterpret_cast<unsigned char *>(sdata) + off, src + off, min(32768u, total - off), bar);
		}
		__syncthreads();
		mbar_wait(bar, 0);
	}
	double * props = reinterpret_cast<double *>(grid_smem + slice_bytes + 64); // [NC][n]
	unsigned char * own_mem = reinterpret_cast<unsigned char *>(props + (size_t) NC * n);
	const size_t own_stride = fused_state_bytes(1, n) + 32 * sizeof(double);
	__shared__ DevState L_own[GRID_MAX_OWNED];
	__shared__ unsigned char s_active[GRID_MAX_CHAINS]; // flags of all chains
	__shared__ int s_any;

	const int my_chain = b + warp * G;
	const bool owner = my_chain < NC;
	const double xub = *a.xabsmax;
	for (int w = 0; w < GRID_WARPS; w++) {
		const int c = b + w * G;
		if (c < NC) {
			const DevState L_tmp = fused_localize_block(S, c / S.n_beta, c % S.n_beta, 1, own_mem + (size_t) w * own_stride);
			if (tid == 0)
				L_own[w] = L_tmp;
			__syncthreads();
		}
	}
	if (owner && lane == 0) {
		const DevState & L = L_own[warp];
		L.pend[0] = PEND_NONE;
		L.cal[0].phase = CAL_IDLE;
		if (select == nullptr || select[my_chain]) {
			atomicAdd(L.n_active, 1);
			cal_begin(L, 0, cal);
		} else {
			L.cal[0].status = -1;
		}
	}
	__syncthreads();
	while (true) {
		// ---- the owners publish what their chains need evaluated next (if anything)
		if (owner) {
			const DevState & L = L_own[warp];
			const int kind = cal_next_kind(L, 0); // same value in every lane (shared memory)
			if (kind != PEND_NONE) {
				chain_propose_warp(L, 0, kind, lane);
				if (lane < n)
					a.props[(size_t) my_chain * n + lane] = L.prop[lane];
			}
			if (lane == 0)
				active[my_chain] = kind != PEND_NONE;
		}
		__syncthreads();
		grid.sync();
		if (tid == 0)
			s_any = 0;
		__syncthreads();
		for (int c = tid; c < NC; c += GRID_THREADS) {
			const int f = __ldcg(active + c);
			s_active[c] = (unsigned char) (f != 0);
			if (f)
				s_any = 1; // benign race: everybody writes 1
		}
		for (int i = tid; i < NC * n; i += GRID_THREADS)
			props[i] = __ldcg(a.props + i);
		__syncthreads();
		if (!s_any)
			break; // the same decision in every CTA: all read the same flags
		// ---- the pending proposals on this CTA's slice
		for (int c = warp; c < NC; c += GRID_WARPS) {
			if (!s_active[c])
				continue;
			const double v = n_slice > 0 ? group_loglik<M>(S, props + (size_t) c * n, sdata, n_slice, xub, lane, 32) : 0.0;
			if (lane == 0)
				a.partials[(size_t) c * G + b] = v;
		}
		__syncthreads();
		grid.sync();
		// ---- the owners: finish the step, advance the chain's calibration
		if (owner && s_active[my_chain]) {
			const DevState & L = L_own[warp];
			double sum = 0.0;
			for (int k = lane; k < G; k += 32)
				sum += __ldcg(a.partials + (size_t) my_chain * G + k);
			sum = warp_sum(sum);
			chain_finalize_warp<M>(L, 0, M::sum0(L.prop) + sum, nullptr, lane);
			if (lane == 0)
				cal_after_step(L, 0, cal);
			__syncwarp();
		}
		__syncthreads();
	}
	for (int w = 0; w < GRID_WARPS; w++) {
		const int c = b + w * G;
		if (c < NC)
			fused_writeback_block(S, L_own[w], c / S.n_beta, c % S.n_beta, false);
	}
}

// markov_chain_calibrate of every selected chain, start to finish in one launch: the per-chain
// state machine of apm_chain.cuh needs nothing from other chains
template<class M>
__global__ void __launch_bounds__(FUSED_MAX_WARPS * 32, 1) fused_calibrate_kernel(const DevState S,
		const FusedArgs a) {
	extern __shared__ __align__(128) unsigned char fused_smem[];
	const Row<M> * sdata = fused_stage_table<M>(a, fused_smem);
	const int ens = blockIdx.x;
	const DevState L = fused_localize(S, ens, fused_smem + fused_table_bytes((long long) a.n_rows * (M::ROW_W / 2)));
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
	const int nb = L.n_beta, n = L.n_par;
	const size_t base = (size_t) ens * nb;
	const double xub = M::HAS_DATA ? *a.xabsmax : 0.0;
	// one chain per warp (with data) or per thread (data-free); `leader` runs the state machine
	const int first = M::HAS_DATA ? warp : (int) threadIdx.x, stride = M::HAS_DATA ? n_warps : (int) blockDim.x;
	const bool leader = !M::HAS_DATA || lane == 0;
	for (int k = first; k < nb; k += stride) {
		if (leader) {
			L.pend[k] = PEND_NONE;
			L.cal[k].phase = CAL_IDLE;
			if (a.select == nullptr || a.select[base + k]) {
				atomicAdd(L.n_active, 1);
				cal_begin(L, k, a.cal);
			} else {
				L.cal[k].status = -1;
			}
		}
		if (M::HAS_DATA)
			__syncwarp();
		int kind = cal_next_kind(L, k);
		while (kind != PEND_NONE) { // same value for the whole warp
			if (M::HAS_DATA)
				chain_propose_warp(L, k, kind, lane);
			else
				chain_propose(L, k, kind);
			const double sum = fused_loglik<M>(L, k, sdata, a.n_rows, xub, lane);
			if (leader) {
				chain_finalize<M>(L, k, M::sum0(L.prop + (size_t) k * n) + sum);
				cal_after_step(L, k, a.cal);
			}
			if (M::HAS_DATA)
				__syncwarp();
			kind = cal_next_kind(L, k);
		}
	}
	fused_writeback(S, L, ens);
}

// ------------------------------------------------------------------ eval
template<class M>
__global__ void eval_finish_kernel(int n, int n_par, const double * params, const double * beta,
		const double * partial, int n_splits, double * prob_out, double * prior_out, const double mc0,
		const double mc1, const double mc2, const double mc3) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n)
		return;
	const double mc[4] = { mc0, mc1, mc2, mc3 };
	const double * p = params + (size_t) k * n_par;
	double sum = M::sum0(p);
	if (M::HAS_DATA)
		for (int s = 0; s < n_splits; s++)
			sum += partial[(size_t) k * n_splits + s];
	double prior = M::HAS_PRIOR ? M::prior(p, n_par, mc) : 0.0;
	prob_out[k] = M::finish(beta[k], sum, prior, p, mc);
	prior_out[k] = prior;
}

// ------------------------------------------------------------------ FP64 peak
// 8 independent DFMA chains per thread, 256 DFMAs per loop trip (loop overhead ~1%); the
// result is stored so nothing is elided.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double * out, int iters, double a, double b) {
	double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
	double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 32; u++) {
			x0 = fma(x0, a, b);
			x1 = fma(x1, a, b);
			x2 = fma(x2, a, b);
			x3 = fma(x3, a, b);
			x4 = fma(x4, a, b);
			x5 = fma(x5, a, b);
			x6 = fma(x6, a, b);
			x7 = fma(x7, a, b);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

} // namespace apm
