// fp64_peak_probe.cu -- what the FP64 pipe of one SM can issue, measured per CLOCK (so the answer does
// not depend on what the clocks do under load): the denominator questions behind bench.py's roofline.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_variants/fp64_peak_probe tools/fp64_peak_probe.cu
//   build_variants/fp64_peak_probe            (on the GPU box; prints one JSON object)
//
// For every variant: FP64 lane-operations per SM per clock (the spec figure is 64), from clock64()
// inside the kernel (first start to last end over the blocks of SM-resident waves) and, for
// reference, per second from CUDA events.
//   dfma_c<C>_w<W>    C independent DFMA chains per thread, W warps per scheduler resident
//   mix_c8_w2         the likelihood kernel's own mix per row evaluation: 18 FP64 + 2 integer + 1/8 LDS.128
//   rcp64h_c8_w<W>    MUFU.RCP64H (rcp.approx.ftz.f64), the seed of the branch-free quotient
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template<int C>
__global__ void __launch_bounds__(256) dfma_kernel(double * out, long long * clk, int iters, double a, double b) {
	double x[C];
#pragma unroll
	for (int c = 0; c < C; c++)
		x[c] = threadIdx.x * 1e-3 + c;
	__syncthreads();
	const long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 16; u++)
#pragma unroll
			for (int c = 0; c < C; c++)
				x[c] = fma(x[c], a, b);
	}
	const long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int c = 0; c < C; c++)
		s += x[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) {
		clk[2 * blockIdx.x] = t0;
		clk[2 * blockIdx.x + 1] = t1;
	}
}

// 18 FP64 + 2 integer instructions per "row evaluation" and one 16-byte shared-memory load per 8 of them:
// the shape of loglik_tiled_kernel<simplesin5>'s inner loop (8 chains x 2 rows per thread), no TMA, no barriers
__global__ void __launch_bounds__(256) mix_kernel(double * out, long long * clk, int iters, double a, double b) {
	__shared__ double2 rows[256 * 2];
	rows[threadIdx.x] = make_double2(threadIdx.x * 1e-3, 0.5);
	rows[threadIdx.x + 256] = make_double2(threadIdx.x * 2e-3, 0.25);
	double acc[8][2];
#pragma unroll
	for (int c = 0; c < 8; c++)
		acc[c][0] = acc[c][1] = 0;
	__syncthreads();
	const long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
		double2 r[2];
		r[0] = rows[(threadIdx.x + i) & 255];
		r[1] = rows[256 + ((threadIdx.x + i) & 255)];
#pragma unroll
		for (int c = 0; c < 8; c++)
#pragma unroll
			for (int u = 0; u < 2; u++) {
				// 18 dependent-ish FP64 ops with the sine's structure: argument (2), reduction (4), polynomial (9), model (3)
				double arg = __dadd_rn(__dmul_rn(a + c, r[u].x), b);
				double t = fma(arg, 0.3183098861837907, 6755399441055744.0);
				int q = __double2loint(t);
				double qd = t - 6755399441055744.0;
				double rr = fma(qd, -3.141592653589793, arg);
				rr = fma(qd, -1.2246467991473532e-16, rr);
				rr = __hiloint2double(__double2hiint(rr) ^ (q << 31), __double2loint(rr));
				double s = rr * rr;
				double p = -0x1.9e96f0e4ab7e2p-41;
				p = fma(p, s, 0x1.60e23f9c870eep-33);
				p = fma(p, s, -0x1.ae6335183e8ccp-26);
				p = fma(p, s, 0x1.71de379039620p-19);
				p = fma(p, s, -0x1.a01a0198a4c74p-13);
				p = fma(p, s, 0x1.111111110723ap-7);
				p = fma(p, s, -0x1.5555555555421p-3);
				double sn = fma(rr * s, p, rr);
				double d = fma(a, sn, b) - r[u].y;
				acc[c][u] = fma(d, d, acc[c][u]);
			}
	}
	const long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int c = 0; c < 8; c++)
		s += acc[c][0] + acc[c][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) {
		clk[2 * blockIdx.x] = t0;
		clk[2 * blockIdx.x + 1] = t1;
	}
}

// 8 compute warps + a ninth warp that (mode 1) exits at once, (mode 2) sleeps until the others are done,
// (mode 3) spins on a shared flag: does a resident idle warp cost the FP64 stream anything?
__global__ void __launch_bounds__(288) ninth_warp_kernel(double * out, long long * clk, int iters, double a, double b, int mode) {
	__shared__ volatile int done;
	if (threadIdx.x == 0)
		done = 0;
	__syncthreads();
	if (threadIdx.x >= 256) {
		if (mode == 2)
			while (!done)
				__nanosleep(256);
		if (mode == 3)
			while (!done)
				;
		return;
	}
	double x[8];
#pragma unroll
	for (int c = 0; c < 8; c++)
		x[c] = threadIdx.x * 1e-3 + c;
	const long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 16; u++)
#pragma unroll
			for (int c = 0; c < 8; c++)
				x[c] = fma(x[c], a, b);
	}
	const long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int c = 0; c < 8; c++)
		s += x[c];
	out[blockIdx.x * 256 + threadIdx.x] = s;
	asm volatile("bar.sync 1, 256;");
	if (threadIdx.x == 0) {
		done = 1;
		clk[2 * blockIdx.x] = t0;
		clk[2 * blockIdx.x + 1] = t1;
	}
}

__global__ void __launch_bounds__(256) rcp_kernel(double * out, long long * clk, int iters) {
	double x[8];
#pragma unroll
	for (int c = 0; c < 8; c++)
		x[c] = 1.5 + threadIdx.x * 1e-3 + c;
	__syncthreads();
	const long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 8; u++)
#pragma unroll
			for (int c = 0; c < 8; c++)
				asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x[c]));
	}
	const long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int c = 0; c < 8; c++)
		s += x[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) {
		clk[2 * blockIdx.x] = t0;
		clk[2 * blockIdx.x + 1] = t1;
	}
}

struct Result {
	double per_sm_per_clock, per_second, ms;
};

template<class Launch>
static Result run(Launch launch, int sms, int blocks_per_sm, double lane_ops_per_thread, long long * d_clk) {
	const int grid = sms * blocks_per_sm;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	launch(grid); // warm-up
	CK(cudaDeviceSynchronize());
	Result best = { 0, 0, 0 };
	for (int rep = 0; rep < 5; rep++) {
		CK(cudaEventRecord(e0));
		launch(grid);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		long long * clk = (long long *) malloc(sizeof(long long) * 2 * grid);
		CK(cudaMemcpy(clk, d_clk, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
		// per block: its own cycle count (every block runs on one SM together with blocks_per_sm - 1 others)
		double cyc = 0;
		for (int b = 0; b < grid; b++)
			cyc += (double) (clk[2 * b + 1] - clk[2 * b]);
		cyc /= grid;
		free(clk);
		const double ops_per_sm = lane_ops_per_thread * 256.0 * blocks_per_sm;
		Result r = { ops_per_sm / cyc, lane_ops_per_thread * 256.0 * grid / (ms * 1e-3), ms };
		if (r.per_sm_per_clock > best.per_sm_per_clock)
			best = r;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return best;
}

int main() {
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	double * d_out;
	long long * d_clk;
	CK(cudaMalloc(&d_out, sizeof(double) * 256 * sms * 16));
	CK(cudaMalloc(&d_clk, sizeof(long long) * 2 * sms * 16));
	const int iters = 4000;
	printf("{\"device\": \"%s\", \"sms\": %d, \"spec_lanes_per_sm_per_clock\": 64", prop.name, sms);
#define DFMA(C, W) { Result r = run([&](int grid) { dfma_kernel<C><<<grid, 256>>>(d_out, d_clk, iters, 0.999999, 1e-9); }, sms, (W) / 2, \
			(double) iters * 16 * C, d_clk); \
		printf(",\n \"dfma_c%d_w%d\": {\"lanes_per_sm_per_clock\": %.3f, \"lane_ops_per_s\": %.4e, \"ms\": %.3f}", C, W, \
				r.per_sm_per_clock, r.per_second, r.ms); }
	// 256 threads = 8 warps = 2 per scheduler; W = warps per scheduler = 2 x blocks per SM
	DFMA(4, 2) DFMA(8, 2) DFMA(16, 2) DFMA(4, 4) DFMA(8, 4) DFMA(8, 8) DFMA(2, 16)
	{
		Result r = run([&](int grid) { mix_kernel<<<grid, 256>>>(d_out, d_clk, iters / 4, 1.000001, 1e-3); }, sms, 1,
				(double) (iters / 4) * 16 * 18, d_clk);
		printf(",\n \"mix_c8_w2\": {\"fp64_lanes_per_sm_per_clock\": %.3f, \"fp64_lane_ops_per_s\": %.4e, \"ms\": %.3f, "
				"\"note\": \"18 FP64 + 2 integer per row evaluation, one LDS.128 per 8\"}", r.per_sm_per_clock, r.per_second, r.ms);
	}
	for (int mode = 1; mode <= 3; mode++) {
		Result r = run([&](int grid) { ninth_warp_kernel<<<grid, 288>>>(d_out, d_clk, iters, 0.999999, 1e-9, mode); }, sms, 1,
				(double) iters * 16 * 8, d_clk);
		printf(",\n \"dfma_c8_w2_plus_ninth_warp_mode%d\": {\"lanes_per_sm_per_clock\": %.3f, \"lane_ops_per_s\": %.4e, \"ms\": %.3f}", mode,
				r.per_sm_per_clock, r.per_second, r.ms);
	}
	for (int w = 2; w <= 8; w *= 2) {
		Result r = run([&](int grid) { rcp_kernel<<<grid, 256>>>(d_out, d_clk, iters); }, sms, w / 2, (double) iters * 64, d_clk);
		printf(",\n \"rcp64h_c8_w%d\": {\"lanes_per_sm_per_clock\": %.3f, \"lane_ops_per_s\": %.4e, \"ms\": %.3f}", w,
				r.per_sm_per_clock, r.per_second, r.ms);
	}
	printf("\n}\n");
	return 0;
}
