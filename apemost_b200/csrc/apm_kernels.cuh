// apm_kernels.cuh -- the CUDA kernels (sm_100a).
//
//   loglik_tiled_kernel<M>   the hot kernel of the tiled path: a persistent grid walks
//                            (chain tile x row split) work items; the data rows stream
//                            through a 2-stage shared-memory ring filled by the TMA
//                            engine (cp.async.bulk + mbarrier), every staged row is used
//                            by all chains of the tile from registers, per-chain sums are
//                            reduced by fp64 warp shuffles in a fixed order.
//   advance_kernel<M>        the control kernel of the tiled path, one CTA per ensemble:
//                            finalise the pending step of every chain (accept/reject,
//                            counters, best, trace, accumulators), resolve the ensemble's
//                            swap, drive the calibration state machines, and draw the next
//                            proposals -- all reference semantics live in apm_chain.cuh.
//   fused_run_kernel<M>      the small-data path: one CTA per ensemble, one warp per
//                            chain, the whole data table resident in shared memory (one
//                            TMA bulk copy), n_rounds x (n_swap steps + swap) in a single
//                            launch.
//   eval_finish_kernel<M>    turns running sums into (prob, prior) for apm_gpu_eval.
//   fp64_peak_kernel         DFMA issue-rate microbenchmark (the roofline denominator).
#pragma once

#include <cuda_runtime.h>
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

// ------------------------------------------------------------------ tiled likelihood
// tuning knobs (overridable at build time for kernel sweeps: tools/kernel_sweep.py)
#ifndef APM_LL_THREADS
#define APM_LL_THREADS 256
#endif
#ifndef APM_LL_RPT
#define APM_LL_RPT 16
#endif
#ifndef APM_LL_MINBLOCKS
#define APM_LL_MINBLOCKS 1
#endif
#ifndef APM_LL_STAGES
#define APM_LL_STAGES 2
#endif
#ifndef APM_LL_CPV
#define APM_LL_CPV 1
#endif
constexpr int LL_THREADS = APM_LL_THREADS;
constexpr int LL_WARPS = LL_THREADS / 32;
constexpr int LL_RPT = APM_LL_RPT;                       // rows per thread per chunk (held in registers)
constexpr int LL_CHUNK = LL_THREADS * LL_RPT;   // 2048 rows = 32 KB per stage
constexpr int LL_STAGES = APM_LL_STAGES;
#ifndef APM_LL_MAX_TILE
#define APM_LL_MAX_TILE 16
#endif
constexpr int LL_MAX_TILE = APM_LL_MAX_TILE;    // chains per work item (upper bound)
constexpr int LL_CPV = APM_LL_CPV;              // chains evaluated together in one visit

struct LLArgs {
	const double * data;    // [n_rows_padded][2], padded with zeros to a multiple of LL_CHUNK
	long long n_rows;
	const double * prop;    // [n_slots][n_par] parameter vectors to evaluate
	const int * pend;       // optional [n_slots]: skip slots with pend < 0 (NULL = all active)
	int n_slots;
	int n_par;
	int tile;               // chains per work item
	int n_ctiles;
	int n_splits;
	int chunks_per_split;
	int n_chunks;
	double * partial;       // [n_slots][n_splits]
	double model_const[4];
};

// One visit: NC chains of the tile against the thread's LL_RPT register-held rows.  All
// NC x LL_RPT row evaluations are independent, so the compiler interleaves them and the fp64
// pipe sees long runs of back-to-back independent instructions; the per-visit costs (constant
// loads, accumulator update, loop control, pipeline drain at the branch) are paid once per
// NC x LL_RPT evaluations.  rows[] is only ever indexed by unrolled constants (registers).
template<class M, int NC>
__device__ __forceinline__ void ll_visit(const double2 (&rows)[LL_RPT], const double * sparams,
		double * sacc, const int c, const int n_par, const double * mc, const bool full_chunk,
		const int n_valid, const int tid, const double xub) {
	typename M::Prep q[NC];
	double acc[NC][2];
#pragma unroll
	for (int u = 0; u < NC; u++) {
		M::prep(q[u], sparams + (c + u) * APM_MAX_PAR, n_par, mc);
		acc[u][0] = 0.0;
		acc[u][1] = 0.0;
	}
	// branch-free fast path when the model says every row of this thread is inside the fast
	// range for these chains (one bound check per visit, none per row); otherwise -- and for
	// every thread of a ragged last chunk -- the exact, masked path
	bool fast = full_chunk;
#pragma unroll
	for (int u = 0; u < NC; u++)
		fast = fast && M::fast_ok(q[u], xub);
	if (fast) {
#pragma unroll
		for (int j = 0; j < LL_RPT; j += 2) {
#pragma unroll
			for (int u = 0; u < NC; u++) {
				acc[u][0] = M::accum_fast(acc[u][0], q[u], rows[j].x, rows[j].y);
				acc[u][1] = M::accum_fast(acc[u][1], q[u], rows[j + 1].x, rows[j + 1].y);
			}
		}
	} else {
#pragma unroll
		for (int u = 0; u < NC; u++) {
			acc[u][0] = 0.0;
			acc[u][1] = 0.0;
#pragma unroll
			for (int j = 0; j < LL_RPT; j++)
				if (j * LL_THREADS + tid < n_valid)
					acc[u][0] = M::accum(acc[u][0], q[u], rows[j].x, rows[j].y);
		}
	}
#pragma unroll
	for (int u = 0; u < NC; u++)
		sacc[(c + u) * LL_THREADS + tid] += acc[u][0] + acc[u][1];
}

template<class M>
__global__ void __launch_bounds__(LL_THREADS, APM_LL_MINBLOCKS) loglik_tiled_kernel(const LLArgs a) {
	extern __shared__ __align__(128) unsigned char ll_smem[];
	double2 * sdata = reinterpret_cast<double2 *>(ll_smem);                      // [STAGES][CHUNK]
	double * sparams = reinterpret_cast<double *>(sdata + LL_STAGES * LL_CHUNK); // [MAX_TILE][MAX_PAR]
	double * sacc = sparams + LL_MAX_TILE * APM_MAX_PAR;                         // [MAX_TILE][THREADS]
	int * sact = reinterpret_cast<int *>(sacc + LL_MAX_TILE * LL_THREADS);       // [MAX_TILE]
	uint64_t * full = reinterpret_cast<uint64_t *>(sact + LL_MAX_TILE);          // [STAGES]

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n_par = a.n_par;
	if (tid == 0) {
		for (int s = 0; s < LL_STAGES; s++)
			mbar_init(&full[s], 1);
		mbar_fence_init();
	}
	__syncthreads();

	const long long n_items = (long long) a.n_ctiles * a.n_splits;
	uint32_t it = 0; // chunks consumed so far by this CTA: stage = it % STAGES, parity = (it / STAGES) & 1
	constexpr uint32_t CHUNK_BYTES = LL_CHUNK * sizeof(double2);

	for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
		const int ctile = (int) (item % a.n_ctiles); // chain tile fastest: neighbours share rows in L2
		const int split = (int) (item / a.n_ctiles);
		const int c0 = ctile * a.tile;
		const int nT = min(a.tile, a.n_slots - c0);
		const int k0 = split * a.chunks_per_split;
		const int nk = min(a.chunks_per_split, a.n_chunks - k0);

		// stage the tile's parameter vectors, clear the accumulators.  Every thread owns one
		// fp64 accumulator per chain of the tile in shared memory (sacc[c][tid], conflict-free):
		// a chain visit costs one LDS + one STS instead of a 5-stage shuffle reduction, and the
		// cross-thread reduction happens once per work item, in a fixed order.
		for (int i = tid; i < nT * n_par; i += LL_THREADS)
			sparams[(i / n_par) * APM_MAX_PAR + (i % n_par)] = a.prop[(size_t) c0 * n_par + i];
		for (int c = 0; c < nT; c++)
			sacc[c * LL_THREADS + tid] = 0.0;
		if (tid < nT)
			sact[tid] = a.pend ? (a.pend[c0 + tid] >= 0) : 1;
		// prologue of the TMA ring
		if (tid == 0) {
			for (int s = 0; s < LL_STAGES && s < nk; s++) {
				const uint32_t st = (it + s) % LL_STAGES;
				mbar_arrive_expect_tx(&full[st], CHUNK_BYTES);
				tma_bulk_g2s(sdata + st * LL_CHUNK, a.data + (size_t) (k0 + s) * LL_CHUNK * 2, CHUNK_BYTES,
						&full[st]);
			}
		}
		__syncthreads();

		for (int k = 0; k < nk; k++, it++) {
			const uint32_t st = it % LL_STAGES;
			mbar_wait(&full[st], (it / LL_STAGES) & 1u);
			double2 rows[LL_RPT];
#pragma unroll
			for (int j = 0; j < LL_RPT; j++)
				rows[j] = sdata[st * LL_CHUNK + j * LL_THREADS + tid];
			__syncthreads(); // every thread holds its rows: the stage can be refilled
			if (tid == 0 && k + LL_STAGES < nk) {
				mbar_arrive_expect_tx(&full[st], CHUNK_BYTES);
				tma_bulk_g2s(sdata + st * LL_CHUNK, a.data + (size_t) (k0 + k + LL_STAGES) * LL_CHUNK * 2,
						CHUNK_BYTES, &full[st]);
			}
			const long long row0 = (long long) (k0 + k) * LL_CHUNK;
			const int n_valid = (int) min((long long) LL_CHUNK, a.n_rows - row0);
			const bool full_chunk = n_valid == LL_CHUNK;
			// upper bound of |x| over this thread's rows: max of the high words (monotonic for
			// |doubles|; Inf/NaN give the largest), rounded up to the next high word
			int xhi = 0;
#pragma unroll
			for (int j = 0; j < LL_RPT; j++)
				xhi = max(xhi, hi32(rows[j].x) & 0x7fffffff);
			const double xub = make_double(xhi + 1, 0);
			int c = 0;
			for (; c + LL_CPV <= nT; c += LL_CPV) {
				bool any = false;
#pragma unroll
				for (int u = 0; u < LL_CPV; u++)
					any |= sact[c + u] != 0;
				if (any)
					ll_visit<M, LL_CPV>(rows, sparams, sacc, c, n_par, a.model_const, full_chunk, n_valid, tid, xub);
			}
			for (; c < nT; c++)
				if (sact[c])
					ll_visit<M, 1>(rows, sparams, sacc, c, n_par, a.model_const, full_chunk, n_valid, tid, xub);
		}
		__syncthreads();
		// fold the tile: warp w reduces chains w, w + WARPS, ...; lane l first adds its 8 strided
		// thread slots in index order, then a butterfly -- the same order every run
		for (int c = warp; c < nT; c += LL_WARPS) {
			double v = 0.0;
#pragma unroll
			for (int k = 0; k < LL_THREADS / 32; k++)
				v += sacc[c * LL_THREADS + k * 32 + lane];
			v = warp_sum(v);
			if (lane == 0)
				a.partial[(size_t) (c0 + c) * a.n_splits + split] = v;
		}
		__syncthreads();
	}
}

constexpr size_t LL_SMEM_BYTES = sizeof(double2) * LL_STAGES * LL_CHUNK
		+ sizeof(double) * LL_MAX_TILE * APM_MAX_PAR + sizeof(double) * LL_MAX_TILE * LL_THREADS
		+ sizeof(int) * LL_MAX_TILE + sizeof(uint64_t) * LL_STAGES;

// ------------------------------------------------------------------ control kernel
enum {
	ADV_FINALIZE = 1, ADV_RECORD = 2, ADV_SWAP = 4, ADV_PROPOSE_RUN = 8, ADV_CALIB = 16, ADV_CALIB_BEGIN = 32
};

struct AdvArgs {
	int flags;
	long long step_index;
	CalibCfgDev cal;
	const unsigned char * select; // calibration selection (ADV_CALIB_BEGIN)
};

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
				if (kind != PEND_NONE)
					chain_propose(S, g, kind);
			} else {
				S.cal[g].phase = CAL_IDLE;
				S.cal[g].status = -1;
			}
		}
		return;
	}
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
				if (kind != PEND_NONE)
					chain_propose(S, g, kind);
			}
		}
	}
	if (a.flags & ADV_SWAP) {
		__syncthreads();
		if (threadIdx.x == 0)
			ensemble_swap(S, ens);
	}
	if (a.flags & ADV_PROPOSE_RUN) {
		__syncthreads();
		for (int k = threadIdx.x; k < S.n_beta; k += blockDim.x)
			chain_propose(S, base + k, S.n_par);
	}
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
