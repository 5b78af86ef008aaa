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
// the same for a thread that has nothing else to do (the likelihood kernel's producer lane): between
// polls it sleeps, so that it does not take issue slots from the compute warps of its scheduler
__device__ __forceinline__ void mbar_wait_idle(uint64_t * bar, uint32_t parity) {
	uint32_t done = 0;
	while (!done) {
		asm volatile(
				"{\n\t"
				".reg .pred p;\n\t"
				"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 2000;\n\t"
				"selp.u32 %0, 1, 0, p;\n\t"
				"}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if (!done)
			__nanosleep(128);
	}
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
	int n_splits;           // the table's chunks are dealt over the splits as evenly as they go: split s has
	int chunks_per_split;   // n_chunks / n_splits chunks, the first n_chunks % n_splits one more (this: the larger count)
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
// Warp-specialised: LL_THREADS compute threads (two warpgroups) + one producer warpgroup, of which a
// single lane talks to the TMA engine.  (With the copies issued from inside a compute warp that warp
// fell behind the other seven by the issue time of every chunk, and the seven waited for it at the
// barrier that ends every item.)  The kernel is launched with 384 threads x 168 registers; the
// producer warpgroup hands its registers back (setmaxnreg.dec) and the compute warpgroups take them
// (setmaxnreg.inc), so the row loop keeps its 8 x 2 register tile.
constexpr int LL_LAUNCH_THREADS = LL_THREADS + 128;
__device__ __forceinline__ void group_bar_compute() {
	asm volatile("bar.sync 1, %0;" :: "n"(LL_THREADS) : "memory");
}

template<class M>
__global__ void __launch_bounds__(LL_LAUNCH_THREADS, 1) loglik_tiled_kernel(const LLArgs a) {
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
	const int sp_base = a.n_chunks / a.n_splits, sp_rem = a.n_chunks - sp_base * a.n_splits; // chunks of a split

	// ---- producer warpgroup: gives its registers back; its first lane walks this CTA's items chunk by
	// chunk, as far ahead as the ring allows (a stage is free again once every compute warp has
	// arrived on its `empty` barrier); the rest of it is done at once
	if (warp >= LL_WARPS) {
		asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
		if (warp == LL_WARPS && lane == 0) {
			uint32_t p_st = 0, p_phase = 1, p_n = 0;
			for (int p_item = blockIdx.x; p_item < n_items; p_item += gridDim.x) {
				const int p_split = p_item / n_ctiles;
				const int k0 = p_split * sp_base + min(p_split, sp_rem);
				const int nk = sp_base + (p_split < sp_rem ? 1 : 0);
				const double * src = a.data + (size_t) k0 * CHUNK * M::ROW_W;
				for (int p_k = 0; p_k < nk; p_k++, p_n++, src += (size_t) CHUNK * M::ROW_W) {
					if (p_n >= LL_STAGES) // the chunk that was in this stage must have been read by every warp
						mbar_wait_idle(&empty[p_st], p_phase);
					mbar_arrive_expect_tx(&full[p_st], CHUNK_BYTES);
					tma_bulk_g2s(sdata + p_st * CHUNK, src, CHUNK_BYTES, &full[p_st]);
					if (++p_st == LL_STAGES) {
						p_st = 0;
						p_phase ^= 1u;
					}
				}
			}
		}
		return;
	}
	asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");

	uint32_t it = 0; // chunks consumed so far: stage = it % STAGES, parity = (it / STAGES) & 1
	for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
		const int split = item / n_ctiles;
		const int ctile = item - split * n_ctiles; // chain tile fastest: neighbours share rows in L2
		const int c0 = ctile * C;
		const int k0 = split * sp_base + min(split, sp_rem);
		const int nk = sp_base + (split < sp_rem ? 1 : 0);

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
		group_bar_compute();
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
		group_bar_compute();
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

constexpr int ADV_THREADS = 256;

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
	// the step of the run this launch records (the same for every block: the counter moves when
	// the LAST block of the launch is done, see the end of the kernel)
	const long long step_index = (a.flags & ADV_RECORD) ? (long long) *(volatile unsigned long long *) S.run_ctr : 0;
	if (a.flags & ADV_FINALIZE) {
		for (int k = threadIdx.x; k < S.n_beta; k += blockDim.x) {
			const int g = base + k;
			if (S.pend[g] == PEND_NONE)
				continue;
			chain_finalize<M>(S, g, chain_gather_sum<M>(S, g));
			if (a.flags & ADV_RECORD)
				chain_record(S, g, step_index);
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
		// do_step for every chain of the ensemble, a thread per (chain, coordinate): a coordinate's
		// draw (Philox + log + sqrt + cos, redraws) depends on nothing but the chain's counter
		__syncthreads();
		const int n = S.n_par;
		for (int idx = threadIdx.x; idx < S.n_beta * n; idx += blockDim.x) {
			const int g = base + idx / n, i = idx - (idx / n) * n;
			S.prop[(size_t) g * n + i] = propose_coordinate(S, g, S.rng_ctr[g], i, S.params[(size_t) g * n + i],
					S.steps[(size_t) g * n + i]);
			if (i == 0)
				S.pend[g] = n;
		}
	}
	if (a.flags & ADV_RECORD) {
		// every block has read the counter by the time it takes its ticket; the last one advances it
		__syncthreads();
		if (threadIdx.x == 0) {
			__threadfence();
			const unsigned long long t = atomicAdd(S.run_ctr + 1, 1ull);
			if (t == gridDim.x - 1) {
				S.run_ctr[1] = 0;
				S.run_ctr[0] = (unsigned long long) step_index + 1;
				__threadfence();
			}
		}
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
// (data-free models, free_run_kernel: two such batches, one being drawn while the other is used, and
// two batches of step outcomes, each of at most 64 doubles per chain)
#ifndef APM_FREE_ATT
#define APM_FREE_ATT 3
#endif
#ifndef APM_FREE_DRAW_SLOTS
#define APM_FREE_DRAW_SLOTS 64
#endif
constexpr int FREE_DRAW_SLOTS = APM_FREE_DRAW_SLOTS, FREE_RING_SLOTS = 64; // doubles per chain and batch
__host__ __device__ inline size_t fused_draws_bytes(int n_beta, bool has_data = true) {
	return (size_t) n_beta * ((has_data ? 32 : 2 * FREE_DRAW_SLOTS + 2 * FREE_RING_SLOTS) * sizeof(double) + sizeof(unsigned long long));
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
// (the copy is made by the `nthr` threads numbered tid = 0 .. nthr - 1 -- a CTA, a warp -- and is
// complete for them once they have synchronised among themselves)
__device__ inline DevState fused_localize_block_by(const DevState & S, int ens, int k0, int nb, unsigned char * mem,
		int tid, int nthr) {
	DevState L = S;
	const int n = S.n_par;
	const size_t base = (size_t) ens * S.n_beta + k0;
	size_t off = 0;
#define X(name, type, width) { \
		type * p = reinterpret_cast<type *>(mem + off); \
		for (int i = tid; i < nb * (width); i += nthr) \
			p[i] = S.name[base * (width) + i]; \
		L.name = p; \
		off += sizeof(type) * (size_t) nb * (width); }
	FUSED_ARRAYS(X)
#undef X
	{
		constexpr int W = sizeof(CalState) / 8;
		u64 * p = reinterpret_cast<u64 *>(mem + off);
		const u64 * src = reinterpret_cast<const u64 *>(S.cal + base);
		for (int i = tid; i < nb * W; i += nthr)
			p[i] = src[i];
		L.cal = reinterpret_cast<CalState *>(p);
		off += sizeof(CalState) * (size_t) nb;
	}
	{
		u64 * p = reinterpret_cast<u64 *>(mem + off);
		if (tid == 0)
			p[0] = S.swap_round[ens];
		L.swap_round = p;
		off += 8;
	}
	{
		int * p = reinterpret_cast<int *>(mem + off);
		for (int i = tid; i < nb; i += nthr)
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
	if (S.marg_mode) { // slot 0 of the local view = this block's first chain (mode 2) or its ensemble (mode 1)
		const size_t slot0 = S.marg_mode == 2 ? base : (size_t) ens;
		L.marg_counts = S.marg_counts + slot0 * n * S.marg_bins;
		L.marg_bsum = S.marg_bsum + slot0 * n;
		L.marg_means = S.marg_means + slot0 * n * S.marg_cap;
		L.marg_n = S.marg_n + slot0;
		L.marg_nb = S.marg_nb + slot0;
	}
	return L;
}

__device__ inline DevState fused_localize_block(const DevState & S, int ens, int k0, int nb, unsigned char * mem) {
	const DevState L = fused_localize_block_by(S, ens, k0, nb, mem, threadIdx.x, blockDim.x);
	__syncthreads();
	return L;
}

__device__ inline DevState fused_localize(const DevState & S, int ens, unsigned char * mem) {
	return fused_localize_block(S, ens, 0, S.n_beta, mem);
}

// L = what fused_localize_block returned for (ens, k0); copies its L.n_beta chains back
// (_by: the copying threads have synchronised among themselves before the call)
__device__ inline void fused_writeback_block_by(const DevState & S, const DevState & L, int ens, int k0,
		bool write_swap_round, int tid, int nthr) {
	const int nb = L.n_beta, n = S.n_par;
	const size_t base = (size_t) ens * S.n_beta + k0;
#define X(name, type, width) \
		for (int i = tid; i < nb * (width); i += nthr) \
			S.name[base * (width) + i] = L.name[i];
	FUSED_ARRAYS(X)
#undef X
	{
		constexpr int W = sizeof(CalState) / 8;
		u64 * dst = reinterpret_cast<u64 *>(S.cal + base);
		const u64 * src = reinterpret_cast<const u64 *>(L.cal);
		for (int i = tid; i < nb * W; i += nthr)
			dst[i] = src[i];
	}
	if (tid == 0 && write_swap_round)
		S.swap_round[ens] = L.swap_round[0];
	for (int i = tid; i < nb; i += nthr)
		S.pend[base + i] = L.pend[i];
}

__device__ inline void fused_writeback_block(const DevState & S, const DevState & L, int ens, int k0,
		bool write_swap_round) {
	__syncthreads();
	fused_writeback_block_by(S, L, ens, k0, write_swap_round, threadIdx.x, blockDim.x);
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
	{ // (data-free models run free_run_kernel below)
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
	}
	fused_writeback(S, L, ens);
}

// ------------------------------------------------------------------ data-free models
// apps/normal.c never touches the data table: a step is Philox + log + sqrt + cos for the jump, the
// model's closed form, Philox + log for the accept test, and the step's book-keeping -- one long
// chain of dependent instructions per Metropolis step and nothing to stream, so what counts is
// how little of it sits between one step's proposal and the next.  A cluster of two CTAs (two SMs)
// per ensemble, 512 threads each, in three roles that work on consecutive batches of K steps
// (16 for one parameter) at the same time, one cluster barrier per batch:
//   * DRAWERS (CTA 1) draw batch b + 1 straight into CTA 0's shared memory (distributed shared
//     memory): the draws depend on the chain's id and step counter only, two at a time per thread;
//   * DECIDERS (CTA 0) play batch b.  A model may split its evaluation into independent terms
//     (M::LANE_TERMS, term / reduce / finish_reduced: the ten bumps of apps/normal.c): a chain then
//     has a group of lanes instead of one thread (4 with the chain's whole state in every lane's
//     registers when the model has a few compile-time parameters, else 8), the terms are dealt out
//     over them and reduced with shuffles.  Every lane of the group forms the proposal, the new
//     prob and the accept decision redundantly (same instructions on the same inputs: same bits);
//     the group's first lane moves the chain and leaves the step's outcome in a ring;
//   * BOOK-KEEPERS (CTA 0, two warps) write batch b - 1 down from the ring (chain_book_batch:
//     counters, n_iter, trace rows, accumulators, marginal statistics), a thread per chain.
// Values and order of operations are those of chain_propose / chain_finalize / chain_record.
// -DAPM_FREE_CLOCKS prints where a warp of each role spends its clocks.
#ifndef APM_FREE_THREADS
#define APM_FREE_THREADS 512 /* (128 registers a thread: the deciders' loop keeps everything in registers) */
#endif
constexpr int FREE_THREADS = APM_FREE_THREADS, FREE_BOOK_WARPS = 2, FREE_CLUSTER = 2;
constexpr int FREE_MAX_DECIDERS = (FREE_THREADS - 32 * FREE_BOOK_WARPS - 32) / 128 * 128; static_assert(FREE_MAX_DECIDERS >= 128 && FREE_MAX_DECIDERS + 32 * FREE_BOOK_WARPS < FREE_THREADS, "a warp must be left for the swap's draws");

template<class M, class = void> struct ModelLaneTerms { static constexpr int value = 0; };
template<class M> struct ModelLaneTerms<M, std::void_t<decltype(M::LANE_TERMS)>> { static constexpr int value = M::LANE_TERMS; };

constexpr int FREE_REG_NPAR = 4; // up to so many compile-time parameters stay in every lane's registers
constexpr int FREE_ATT = APM_FREE_ATT; // attempts of every jump drawn ahead (the truncated proposal redraws until inside the bounds)

template<class M>
__global__ void __launch_bounds__(FREE_THREADS, 1) free_run_kernel(const DevState S, const FusedArgs a) {
	extern __shared__ __align__(128) unsigned char fused_smem[];
	constexpr int TERMS = ModelLaneTerms<M>::value;
#ifndef APM_FREE_LPC_SMALL
#define APM_FREE_LPC_SMALL 4
#endif
	constexpr bool SMALL = TERMS > 0 && M::NPAR > 0 && M::NPAR <= FREE_REG_NPAR;
	constexpr int LPC = TERMS > 0 ? (SMALL ? APM_FREE_LPC_SMALL : 8) : 1; // lanes per chain
	__shared__ DevState L_sh;
	__shared__ double swap_drawn[2];
	// a cluster of two CTAs (two SMs) per ensemble: CTA 0 plays the chains and keeps their books, CTA 1
	// only draws, straight into CTA 0's shared memory (distributed shared memory) -- the draws are
	// two thirds of a step's instructions and most of its fp64 work, and on one SM they competed
	// with the deciders for issue slots and the fp64 pipe
	const unsigned rank = cooperative_groups::this_cluster().block_rank();
	const int ens = blockIdx.x / FREE_CLUSTER, tid = threadIdx.x;
	{
		const DevState L_tmp = fused_localize(S, ens, fused_smem + fused_table_bytes(0));
		if (tid == 0)
			L_sh = L_tmp;
		__syncthreads();
	}
	const DevState & L = L_sh;
	// (a model with a fixed number of parameters -- apps/normal.c has one -- makes n a compile-time
	// constant: the strides and the per-coordinate loops below fold away)
	const int nb = L.n_beta, n = M::NPAR > 0 ? M::NPAR : L.n_par, RW = n + 3;
	const int NE = FREE_ATT * n + 1;                    // draws per step: FREE_ATT attempts per coordinate + the accept draw
	const int K = max(1, min(FREE_DRAW_SLOTS / NE, FREE_RING_SLOTS / RW));        // steps per batch
	// after the state: draws [2][K NE <= 64][nb], outcome ring [2][K RW <= 64][nb], counters [nb]
	double * dbuf = reinterpret_cast<double *>(fused_smem + fused_table_bytes(0) + fused_state_bytes(nb, n));
	double * ring = dbuf + (size_t) 2 * nb * FREE_DRAW_SLOTS;
	u64 * ctr0 = reinterpret_cast<u64 *>(ring + (size_t) 2 * nb * FREE_RING_SLOTS); // the chains' counters at the launch's start
	const int n_dec = min((nb * LPC + 31) / 32 * 32, FREE_MAX_DECIDERS);
	const int n_book = FREE_BOOK_WARPS * 32;
	double * dbuf_w = cooperative_groups::this_cluster().map_shared_rank(dbuf, 0); // CTA 0's draw buffers
	const int n_slots = n_dec / LPC, slot = tid / LPC, sl = tid % LPC;
	const unsigned gmask = LPC == 1 ? (1u << (tid & 31)) : (0xffu << ((tid & 31) & ~7));
	const int proposal = L.proposal;
	const unsigned circular = L.circular_mask, quirks = L.quirks;
	for (int k = tid; k < nb; k += blockDim.x)
		ctr0[k] = L.rng_ctr[k];
	__syncthreads();
	const long long total = a.n_rounds * a.n_swap;
	// batch: steps [s0, s0 + ns) of the launch, never across a round's end (the swap needs every chain
	// at the same step)
	// (pos = the batch's first step within its round, carried along: no 64-bit remainder per batch)
	auto batch_len = [&](long long s0, int pos) -> int {
		if (s0 >= total)
			return 0;
		const int left = a.n_swap - pos;
		return left < K ? left : K;
	};
	// draws [j][e][k], k fastest: e = d * FREE_ATT + attempt the unit jump of coordinate d, e = NE - 1
	// log(u) of the accept test.  The `count` drawing threads form a grid of (chains) x (draws of a
	// chain): a thread keeps its chain and walks (j, e) in strides, no division per draw.
	auto produce = [&](long long s0, int ns, double * buf, int t, int count) {
		const int cols = min(nb, count), rows = count / cols;
		const int pc = t % cols, pr = t / cols;
		if (pr >= rows)
			return;
		// a draw is one long chain of dependent instructions (Philox, logarithm, square root, cosine):
		// a thread takes them two at a time, side by side
		const int NJ = NE - 1, nj = ns * NJ; // jump draws of a chain in this batch: m = j NJ + e
		for (int k = pc; k < nb; k += cols) {
			const uint32_t id = chain_rng_id(L, k);
			const u64 c0 = ctr0[k] + (u64) s0;
			for (int m = pr; m < nj; m += 2 * rows) {
				const bool two = m + rows < nj;
				const int ma = m, mb = two ? m + rows : m;
				const int ja = ma / NJ, ea = ma - ja * NJ, jb = mb / NJ, eb = mb - jb * NJ;
				const int da = ea / FREE_ATT, db = eb / FREE_ATT; // (a constant divisor)
				double u0a, u1a, u0b, u1b, za, zb;
				philox_uniforms(L.seed, id, c0 + (u64) ja, PURPOSE_JUMP, (uint32_t) da, (uint32_t) (ea - da * FREE_ATT), u0a, u1a);
				philox_uniforms(L.seed, id, c0 + (u64) jb, PURPOSE_JUMP, (uint32_t) db, (uint32_t) (eb - db * FREE_ATT), u0b, u1b);
				jump_unit2(proposal, u0a, u1a, u0b, u1b, za, zb);
				buf[(size_t) (ja * NE + ea) * nb + k] = za;
				if (two)
					buf[(size_t) (jb * NE + eb) * nb + k] = zb;
			}
			for (int j = pr; j < ns; j += 2 * rows) {
				const bool two = j + rows < ns;
				const int jb = two ? j + rows : j;
				double u0a, u0b, dummy;
				philox_uniforms(L.seed, id, c0 + (u64) j, PURPOSE_ACCEPT, 0u, 0u, u0a, dummy);
				philox_uniforms(L.seed, id, c0 + (u64) jb, PURPOSE_ACCEPT, 0u, 0u, u0b, dummy);
				const double la = log(u0a), lb = log(u0b);
				buf[(size_t) (j * NE + NE - 1) * nb + k] = la;
				if (two)
					buf[(size_t) (jb * NE + NE - 1) * nb + k] = lb;
			}
		}
	};
	// do_step_for (reference src/markov_chain.c:226-270) with the first FREE_ATT attempts at hand:
	// z = the unit jumps of this coordinate, strided by nb
	auto propose_from = [&](int k, u64 ctr, int i, double x, double st, double lo, double hi, const double * z) -> double {
		double v = x + jump_apply(proposal, st, z[0]);
		if (v > hi || v < lo) {
			bool inside = false;
			if (!((circular >> i) & 1u)) {
#pragma unroll
				for (int t = 1; t < FREE_ATT; t++)
					if (!inside) {
						v = x + jump_apply(proposal, st, z[(size_t) t * nb]);
						inside = !(v > hi || v < lo);
					}
			}
			if (!inside) // the wrap of a circular parameter, or more redraws than were drawn ahead
				v = propose_coordinate(L, k, ctr, i, x, st, ((circular >> i) & 1u) ? 0u : (unsigned) FREE_ATT);
		}
		return v;
	};
	// the book-keepers: book-keeping thread (w, l) has chains l * BOOK_WARPS + w, + 32 * BOOK_WARPS, ...
	auto book = [&](long long s0, int ns, const double * rb, int t) {
		if (ns > 0)
			for (int k = (t & 31) * FREE_BOOK_WARPS + (t >> 5); k < nb; k += n_book)
				chain_book_batch(L, k, ns, rb + k, (size_t) RW * nb, (size_t) nb, s0);
	};
	long long s0 = 0, s_prev = 0;
	int pos = 0;
	int ns = batch_len(0, 0), ns_prev = 0;
	if (rank == 1)
		produce(0, ns, dbuf_w, tid, blockDim.x);
	cooperative_groups::this_cluster().sync();
	int cur = 0;
#ifdef APM_FREE_CLOCKS /* (timing experiments only: where a warp of each role spends its clocks) */
	long long clk_work = 0, clk_bar = 0, clk_end = 0, clk_t = clock64();
#define FREE_CLK(acc) { const long long t_ = clock64(); acc += t_ - clk_t; clk_t = t_; }
#else
#define FREE_CLK(acc)
#endif
	while (ns > 0) {
		const long long s_next = s0 + ns;
		const int pos_next = pos + ns == a.n_swap ? 0 : pos + ns;
		const int ns_next = batch_len(s_next, pos_next);
		const double * buf = dbuf + (size_t) cur * nb * FREE_DRAW_SLOTS;
		double * rb = ring + (size_t) cur * nb * FREE_RING_SLOTS;
		if (rank == 1) {
			produce(s_next, ns_next, dbuf_w + (size_t) (1 - cur) * nb * FREE_DRAW_SLOTS, tid, blockDim.x);
		} else if (tid >= n_dec + n_book) {
			// the rest of CTA 0 neither decides nor keeps books; one of its threads takes the draws of
			// the swap that follows a round's last batch, so that they are not waited for at the round's end
			if (tid == n_dec + n_book && pos_next == 0)
				ensemble_swap_draws(L, 0, swap_drawn);
		} else if (tid >= n_dec) {
			book(s_prev, ns_prev, ring + (size_t) (1 - cur) * nb * FREE_RING_SLOTS, tid - n_dec);
		} else if constexpr (TERMS > 0) {
			// 8 lanes per chain; the chain's point, prob, prior and best stay in the lanes' registers
			// for the batch (lane sl holds coordinate sl -- and sl + 8 when n_par > 8)
			auto decide = [&](auto cpl_tag) {
				constexpr int CPL = decltype(cpl_tag)::value;
				const size_t step_stride = (size_t) NE * nb, ring_stride = (size_t) RW * nb;
				for (int k = slot; k < nb; k += n_slots) {
					double * q = L.prop + (size_t) k * n;
					const double * mc = L.model_const;
					double x[CPL], st[CPL], lo[CPL], hi[CPL], v[CPL];
#pragma unroll
					for (int c = 0; c < CPL; c++) {
						const int i = sl + 8 * c;
						const bool mine = i < n;
						x[c] = mine ? L.params[(size_t) k * n + i] : 0.0;
						st[c] = mine ? L.steps[(size_t) k * n + i] : 0.0;
						lo[c] = mine ? L.pmin[i] : 0.0;
						hi[c] = mine ? L.pmax[i] : 0.0;
						v[c] = x[c];
					}
					double prob = L.prob[k], prior = L.prior[k], best = L.prob_best[k];
					const double beta = L.beta[k];
					const double * z = buf + k;   // this step's draws of chain k, strided by nb
					double * e = rb + k;          // this step's outcome
					u64 ctr = ctr0[k] + (u64) s0;
					for (int j = 0; j < ns; j++, z += step_stride, e += ring_stride, ctr++) {
#pragma unroll
						for (int c = 0; c < CPL; c++) {
							const int i = sl + 8 * c;
							if (i < n) {
								// do_step_for (reference src/markov_chain.c:226-270), the first FREE_ATT attempts at hand
								const double * zi = z + (size_t) (i * FREE_ATT) * nb;
								double w = x[c] + jump_apply(proposal, st[c], zi[0]);
								if (w > hi[c] || w < lo[c]) {
									bool inside = false;
									const bool wraps = (circular >> i) & 1u;
									if (!wraps) {
#pragma unroll
										for (int t = 1; t < FREE_ATT; t++)
											if (!inside) {
												w = x[c] + jump_apply(proposal, st[c], zi[(size_t) t * nb]);
												inside = !(w > hi[c] || w < lo[c]);
											}
									}
									if (!inside) // a circular parameter's wrap, or more redraws than were drawn ahead
										w = propose_coordinate(L, k, ctr, i, x[c], st[c], wraps ? 0u : (unsigned) FREE_ATT);
								}
								v[c] = w;
								q[i] = w;
							}
						}
						__syncwarp(gmask);
						double prior_new = prior;
						if (M::HAS_PRIOR)
							prior_new = M::prior(q, n, mc);
						// the terms: lane sl takes sl, sl + 8, ... side by side
						double r = M::reduce_init();
#pragma unroll
						for (int t0 = 0; t0 < TERMS; t0 += LPC) {
							const int t = t0 + sl;
							const double term = M::term(t < TERMS ? t : TERMS - 1, q, mc);
							r = t < TERMS ? M::reduce(r, term) : r;
						}
#pragma unroll
						for (int o = LPC / 2; o > 0; o >>= 1)
							r = M::reduce(r, __shfl_xor_sync(gmask, r, o));
						const double prob_new = M::finish_reduced(beta, r);
						// check_accept (reference src/markov_chain.c:282-311), as in chain_finalize_value
						bool accepted;
						if (prob_new == prob)
							accepted = true;
						else if (prob_new > prob)
							accepted = true;
						else
							accepted = z[(size_t) (NE - 1) * nb] < (prob_new - prob);
						if (accepted) {
#pragma unroll
							for (int c = 0; c < CPL; c++)
								x[c] = v[c];
							prob = prob_new;
							prior = prior_new;
						} else if (quirks & 2u) {
							prior = prior_new; // revert() restores prob only
						}
						const bool better = prob > best; // mcmc_check_best (chain_book_step sees it done)
						best = better ? prob : best;
#pragma unroll
						for (int c = 0; c < CPL; c++) {
							const int i = sl + 8 * c;
							if (i < n) {
								e[(size_t) (3 + i) * nb] = x[c];
								if (better)
									L.params_best[(size_t) k * n + i] = x[c];
							}
						}
						if (sl == 0) {
							e[0] = accepted ? 1.0 : 0.0;
							e[(size_t) nb] = prob;
							e[(size_t) 2 * nb] = prior;
						}
						__syncwarp(gmask);
					}
#pragma unroll
					for (int c = 0; c < CPL; c++) {
						const int i = sl + 8 * c;
						if (i < n)
							L.params[(size_t) k * n + i] = x[c];
					}
					if (sl == 0) {
						L.prob[k] = prob;
						L.prior[k] = prior;
						L.prob_best[k] = best;
					}
				}
			};
			// A model with a handful of parameters, their number known at compile time (apps/normal.c
			// has one): every lane of the chain's group forms the WHOLE proposal redundantly (same
			// instructions on the same inputs: same bits), so the point never leaves the registers --
			// no shared-memory round trip and no group barrier between proposal and terms -- and every
			// warp runs the same trips (a group without a chain shadows the last chain and writes
			// nothing), which makes the butterfly's shuffles full-mask ones.
			auto decide_small = [&]() {
				constexpr int NP = M::NPAR > 0 ? M::NPAR : 1;
				const size_t step_stride = (size_t) NE * nb, ring_stride = (size_t) RW * nb;
				const double * mc = L.model_const;
				// (the hot rungs redraw their proposals more often than the cold ones, and a warp waits for
				// whichever of its chains does: warp w takes chains w, w + n_warps, ... -- a mix of rungs)
				constexpr int SPW = 32 / LPC;
				const int pslot = (slot % SPW) * (n_dec / 32) + slot / SPW;
				for (int k0 = 0; k0 < nb; k0 += n_slots) {
					const bool writer = k0 + pslot < nb && sl == 0;
					const int k = k0 + pslot < nb ? k0 + pslot : nb - 1;
					double x[NP], st[NP], lo[NP], hi[NP], v[NP];
#pragma unroll
					for (int i = 0; i < NP; i++) {
						x[i] = L.params[(size_t) k * NP + i];
						st[i] = L.steps[(size_t) k * NP + i];
						lo[i] = L.pmin[i];
						hi[i] = L.pmax[i];
					}
					double prob = L.prob[k], prior = L.prior[k], best = L.prob_best[k];
					const double beta = L.beta[k];
					const double * z = buf + k;   // this step's draws of chain k, strided by nb
					double * e = rb + k;          // this step's outcome
					u64 ctr = ctr0[k] + (u64) s0;
					for (int j = 0; j < ns; j++, z += step_stride, e += ring_stride, ctr++) {
						const double logu = z[(size_t) (NE - 1) * nb];
#pragma unroll
						for (int i = 0; i < NP; i++) {
							// do_step_for (reference src/markov_chain.c:226-270), the first FREE_ATT attempts at hand
							const double * zi = z + (size_t) (i * FREE_ATT) * nb;
							double w = x[i] + jump_apply(proposal, st[i], zi[0]);
							if (w > hi[i] || w < lo[i]) {
								bool inside = false;
								const bool wraps = (circular >> i) & 1u;
								if (!wraps) {
#pragma unroll
									for (int t = 1; t < FREE_ATT; t++)
										if (!inside) {
											w = x[i] + jump_apply(proposal, st[i], zi[(size_t) t * nb]);
											inside = !(w > hi[i] || w < lo[i]);
										}
								}
								if (!inside) // a circular parameter's wrap, or more redraws than were drawn ahead
									w = propose_coordinate(L, k, ctr, i, x[i], st[i], wraps ? 0u : (unsigned) FREE_ATT);
							}
							v[i] = w;
						}
						double prior_new = prior;
						if (M::HAS_PRIOR)
							prior_new = M::prior(v, NP, mc);
						double r = M::reduce_init();
#pragma unroll
						for (int t0 = 0; t0 < TERMS; t0 += LPC) {
							const int t = t0 + sl;
							const double term = M::term(t < TERMS ? t : TERMS - 1, v, mc);
							r = t < TERMS ? M::reduce(r, term) : r;
						}
						__syncwarp();
#pragma unroll
						for (int o = LPC / 2; o > 0; o >>= 1)
							r = M::reduce(r, __shfl_xor_sync(0xffffffffu, r, o));
						const double prob_new = M::finish_reduced(beta, r);
						// check_accept (reference src/markov_chain.c:282-311), as in chain_finalize_value
						bool accepted;
						if (prob_new == prob)
							accepted = true;
						else if (prob_new > prob)
							accepted = true;
						else
							accepted = logu < (prob_new - prob);
						if (accepted) {
#pragma unroll
							for (int i = 0; i < NP; i++)
								x[i] = v[i];
							prob = prob_new;
							prior = prior_new;
						} else if (quirks & 2u) {
							prior = prior_new; // revert() restores prob only
						}
						const bool better = prob > best; // mcmc_check_best (chain_book_step sees it done)
						best = better ? prob : best;
						if (writer) {
							e[0] = accepted ? 1.0 : 0.0;
							e[(size_t) nb] = prob;
							e[(size_t) 2 * nb] = prior;
#pragma unroll
							for (int i = 0; i < NP; i++) {
								e[(size_t) (3 + i) * nb] = x[i];
								if (better)
									L.params_best[(size_t) k * NP + i] = x[i];
							}
						}
					}
					if (writer) {
#pragma unroll
						for (int i = 0; i < NP; i++)
							L.params[(size_t) k * NP + i] = x[i];
						L.prob[k] = prob;
						L.prior[k] = prior;
						L.prob_best[k] = best;
					}
				}
			};
			if constexpr (SMALL)
				decide_small();
			else if (n <= 8)
				decide(std::integral_constant<int, 1>());
			else
				decide(std::integral_constant<int, APM_MAX_PAR / 8>());
		} else {
			// a thread per chain
			for (int k = slot; k < nb; k += n_slots) {
				double * q = L.prop + (size_t) k * n;
				double * p = L.params + (size_t) k * n;
				for (int j = 0; j < ns; j++) {
					for (int i = 0; i < n; i++)
						q[i] = propose_from(k, ctr0[k] + (u64) (s0 + j), i, p[i], L.steps[(size_t) k * n + i], L.pmin[i],
								L.pmax[i], buf + (size_t) (j * NE + i * FREE_ATT) * nb + k);
					const double prob_old = L.prob[k], prior_old = L.prior[k];
					double prior_new = prior_old;
					if (M::HAS_PRIOR)
						prior_new = M::prior(q, n, L.model_const);
					const double prob_new = M::finish(L.beta[k], M::sum0(q), prior_new, q, L.model_const);
					bool accepted;
					if (prob_new == prob_old)
						accepted = true;
					else if (prob_new > prob_old)
						accepted = true;
					else
						accepted = buf[(size_t) (j * NE + NE - 1) * nb + k] < (prob_new - prob_old);
					double prob_after = prob_old, prior_after = prior_old;
					if (accepted) {
						for (int i = 0; i < n; i++)
							p[i] = q[i];
						prob_after = prob_new;
						prior_after = prior_new;
					} else if (quirks & 2u) {
						prior_after = prior_new;
					}
					L.prob[k] = prob_after;
					L.prior[k] = prior_after;
					if (prob_after > L.prob_best[k]) {
						L.prob_best[k] = prob_after;
						for (int i = 0; i < n; i++)
							L.params_best[(size_t) k * n + i] = p[i];
					}
					double * e = rb + (size_t) j * RW * nb + k;
					e[0] = accepted ? 1.0 : 0.0;
					e[(size_t) nb] = prob_after;
					e[(size_t) 2 * nb] = prior_after;
					for (int i = 0; i < n; i++)
						e[(size_t) (3 + i) * nb] = p[i];
				}
			}
		}
		FREE_CLK(clk_work)
		cooperative_groups::this_cluster().sync(); // batch b is decided, batch b + 1 is drawn (and visible in CTA 0), batch b - 1 is written down
		FREE_CLK(clk_bar)
		if (rank == 0 && pos_next == 0) {
			// round end: adapt (if compiled in: it reads the counters, so the books are brought up
			// to date first), tempering_interaction for this ensemble
			if (L.adapt) {
				if (tid >= n_dec && tid < n_dec + n_book)
					book(s0, ns, rb, tid - n_dec); // (the batch before was written down during this one)
				ns = 0; // nothing left to write down for this batch
				__syncthreads();
				for (int k = tid; k < nb; k += blockDim.x)
					chain_adapt(L, k);
				__syncthreads();
			}
			if (tid == 0)
				ensemble_swap(L, 0, nullptr, nullptr, swap_drawn);
			__syncthreads();
		}
		FREE_CLK(clk_end)
		s_prev = s0;
		ns_prev = ns;
		s0 = s_next;
		pos = pos_next;
		ns = ns_next;
		cur = 1 - cur;
	}
#ifdef APM_FREE_CLOCKS
	if (blockIdx.x < 2 && ((tid < n_dec && tid % 32 == 0) || tid == n_dec || tid == n_dec + n_book))
		printf("rank %u tid %d: work %lld barrier %lld round-end %lld clocks, %lld steps\n", rank, tid, clk_work, clk_bar, clk_end, total);
#endif
	if (rank != 0)
		return;
	if (tid >= n_dec && tid < n_dec + n_book)
		book(s_prev, ns_prev, ring + (size_t) (1 - cur) * nb * FREE_RING_SLOTS, tid - n_dec);
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

// ------------------------------------------------------------------ calibration, warp groups per chain
// markov_chain_calibrate (or apm_gpu_steps) when only a FEW chains are selected -- calibrate_first
// is ONE chain, calibrate_rest with SKIP_CALIBRATE_ALLCHAINS one more (reference
// src/parallel_tempering.c:78-207) -- and the table fits in shared memory.  The fused kernel above
// gives a chain one warp of one SM; here the selected chains are dealt out over the SMs (sel_idx:
// the host's compacted list) and a GROUP of WC = 16 / NG warps works on each (NG = chains a CTA
// holds at a time): all of them walk the table (rows strided over the group's lanes, fp64 butterfly
// per warp), the group's first warp adds the warps' sums in index order and plays the chain's
// state machine of apm_chain.cuh -- finalise, cal_after_step, next proposal, its coordinates drawn
// lane-parallel -- on a shared-memory copy of the chain's state.  Two named barriers per step.
// Chains do not interact during calibration, so no cluster is needed: every CTA stages the table
// itself (it is L2-resident after the first).
constexpr int GROUP_MAX = 8; // chains a CTA works on concurrently (named barriers 1 .. 8)

__host__ __device__ inline size_t group_state_bytes(int n_par) {
	// the chain's state, the warps' partial sums [16], the pending kind, the draws and the prior made ahead [APM_MAX_PAR + 2]
	return fused_state_bytes(1, n_par) + 16 * sizeof(double) + 16 + (APM_MAX_PAR + 2) * sizeof(double);
}

struct GroupArgs {
	FusedArgs f;
	const int * sel_idx; // [n_sel] chains to calibrate, ascending
	int n_sel;
	int ng;              // chains per CTA at a time (1, 2, 4 or 8); warps per chain = 16 / ng
};

template<class M>
__global__ void __launch_bounds__(FUSED_MAX_WARPS * 32, 1) group_calibrate_kernel(const DevState S, const GroupArgs ga) {
	extern __shared__ __align__(128) unsigned char fused_smem[];
	const FusedArgs & a = ga.f;
	const Row<M> * sdata = fused_stage_table<M>(a, fused_smem);
	__shared__ DevState L_grp[GROUP_MAX];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int G = gridDim.x, NG = ga.ng, WC = FUSED_MAX_WARPS / NG;
	const int n = S.n_par;
	const int q = warp / WC, wi = warp - q * WC;   // this warp's group, its place in the group
	// The group's last warp walks no rows: while the others do, it draws what the step's tail needs
	// and what does not depend on the step's outcome -- log(u) of this step's accept test and the
	// unit jumps of EVERY coordinate for the next step (its kind is not known yet; a chain's draws
	// depend on its id and step counter only) -- so that the first warp's serial tail is the sum,
	// the accept decision, the state machine and a multiply-add per coordinate.
	const int LW = WC - 1;                         // warps that walk the table (WC >= 2)
	const bool drawer = wi == WC - 1;
	const int GT = WC * 32;                        // threads of the group (named barrier)
	const int GL = LW * 32, gl = wi * 32 + lane;   // lanes walking the table
	const double xub = *a.xabsmax;
	unsigned char * mem = fused_smem + fused_table_bytes((long long) a.n_rows * (M::ROW_W / 2)) + (size_t) q * group_state_bytes(n);
	double * sums = reinterpret_cast<double *>(mem + fused_state_bytes(1, n));
	volatile int * kind_s = reinterpret_cast<volatile int *>(sums + 16);
	double * draws = sums + 18;                    // [0, n) unit jumps for counter + 1, [n] log(u) for counter, [n + 1] the proposal's prior
	// chains that were not selected: the same marks the other calibration kernels leave
	for (int g = blockIdx.x * blockDim.x + tid; g < S.n_chains; g += G * blockDim.x)
		if (a.select != nullptr && !a.select[g]) {
			S.pend[g] = PEND_NONE;
			S.cal[g].phase = CAL_IDLE;
			S.cal[g].status = -1;
		}
	// the next proposal from the unit jumps at hand (do_step_for, reference src/markov_chain.c:226-270);
	// a first attempt outside the bounds goes the regular way (wrap or redraws)
	auto propose_ahead = [&](const DevState & L, int kind) {
		if (lane < n) {
			const double x = L.params[lane];
			double v = x;
			if (kind == n || kind == lane) {
				const double st = L.steps[lane];
				v = x + jump_apply(L.proposal, st, draws[lane]);
				if (v > L.pmax[lane] || v < L.pmin[lane])
					v = propose_coordinate(L, 0, L.rng_ctr[0], lane, x, st);
			}
			L.prop[lane] = v;
		}
		if (lane == 0)
			L.pend[0] = kind;
		__syncwarp();
	};
	// list position p goes to CTA p mod G, group (p / G) mod NG, one after the other
	for (int p = blockIdx.x + q * G; p < ga.n_sel; p += G * NG) {
		const int g = ga.sel_idx[p];
		DevState & L = L_grp[q];
		if (wi == 0) {
			const DevState L_tmp = fused_localize_block_by(S, g / S.n_beta, g % S.n_beta, 1, mem, lane, 32);
			if (lane == 0)
				L = L_tmp;
			__syncwarp();
			if (lane == 0) {
				L.pend[0] = PEND_NONE;
				L.cal[0].phase = CAL_IDLE;
				atomicAdd(L.n_active, 1);
				cal_begin(L, 0, a.cal);
			}
			__syncwarp();
			const int kind = cal_next_kind(L, 0);
			if (kind != PEND_NONE)
				chain_propose_warp(L, 0, kind, lane);
			if (lane == 0)
				*kind_s = kind;
		}
		group_bar(1 + q, GT);
		while (*kind_s != PEND_NONE) {
			if (drawer) {
				if (lane <= n) {
					const u64 ctr = L.rng_ctr[0];
					const bool is_jump = lane < n;
					double u0, u1;
					philox_uniforms(L.seed, chain_rng_id(L, 0), is_jump ? ctr + 1 : ctr, is_jump ? PURPOSE_JUMP : PURPOSE_ACCEPT,
							is_jump ? (uint32_t) lane : 0u, 0, u0, u1);
					draws[lane] = is_jump ? jump_unit(L.proposal, u0, u1) : log(u0);
				} else if (M::HAS_PRIOR && lane == n + 1) {
					draws[n + 1] = M::prior(L.prop, n, L.model_const); // the pending proposal's prior
				}
			} else {
				const double part = group_loglik<M>(L, L.prop, sdata, a.n_rows, xub, gl, GL);
				if (lane == 0)
					sums[wi] = part;
			}
			group_bar(1 + q, GT);
			if (wi == 0) {
				double sum = sums[0];
				for (int w = 1; w < LW; w++)
					sum += sums[w];
				chain_finalize_warp<M>(L, 0, M::sum0(L.prop) + sum, draws + n, lane, M::HAS_PRIOR ? draws + n + 1 : nullptr);
				if (lane == 0)
					cal_after_step(L, 0, a.cal);
				__syncwarp();
				const int kind = cal_next_kind(L, 0);
				if (kind != PEND_NONE)
					propose_ahead(L, kind);
				if (lane == 0)
					*kind_s = kind;
			}
			group_bar(1 + q, GT);
		}
		if (wi == 0) {
			__syncwarp();
			fused_writeback_block_by(S, L, g / S.n_beta, g % S.n_beta, false, lane, 32);
		}
		group_bar(1 + q, GT); // the group's shared memory is free for its next chain
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

// the same per CLOCK: every block brackets its DFMA stream with clock64(), so the answer (FP64
// lane-operations per SM per clock; 64 by the architecture) does not depend on what the clocks do
// under a long pure-DFMA load (the per-second figure above sags with them: that load is the one
// thing on this GPU that runs into the power limit)
__global__ void __launch_bounds__(256) fp64_peak_clock_kernel(double * out, long long * cycles, int iters, double a, double b) {
	double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
	double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	__syncthreads();
	const long long t0 = clock64();
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
	const long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
	if (threadIdx.x == 0)
		cycles[blockIdx.x] = t1 - t0;
}

} // namespace apm
