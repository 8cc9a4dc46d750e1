// qpde_gmwb.cu — the 2-D (W, A) GMWB impulse-control HJBQVI sweep with the
// penalised scheme (BASELINE.json configs[4]).
//
// Reference call stack replaced:
//   HJBQVI::solve wiring                  Modules/HJBQVI/HJBQVI.hpp:590-811
//   ControlledOperator::A / b             :158-335   (5-point upwinded stencil)
//   HJBQVILinearBoundary / ZeroDiffusionRightBoundary   :1285-1340
//   MinPolicyIteration (stochastic, impulse)            Core/PolicyIteration.hpp:54-105
//   Impulse::A / b (bilinear rows)        Core/Impulse.hpp:87-216
//   MinPenaltyMethod (non-direct)         Core/PenaltyMethod.hpp:37-119
//   ReverseBDFOne                         Core/BDF.hpp:32-46
//   ToleranceIteration / relativeError    Core/IterativeMethod.hpp:1125-1254
//   BiCGSTABSolver (ILUT)                 Bindings/Eigen.hpp:170-204
//   model lambdas                         examples/hjbqvi/gmwb.cpp:92-176
//
// Per inner (policy) iteration:
//   k_gmwb_policy  one thread per node: argmin over the stochastic controls of
//                  (L^q x - f^q)_i, argmin over the impulse grid of
//                  ((I - M_z) x - g_z)_i, penalty flag, and the assembled row of
//                  (I + h L^q* + P (I - M*)) with its right-hand side;
//   k_gmwb_solve   the linear system.  Every entry outside a W-line points to an
//                  equal-or-lower A (upwinded drift -q <= 0, impulses with
//                  z > 0), so the matrix is block lower-triangular by A-line:
//                  block Gauss-Seidel over the lines in ascending A, each line a
//                  tridiagonal solve by the partition solver of
//                  qpde_partition.cuh; same-line impulse targets that are not
//                  adjacent use the previous sweep, hence a few sweeps to a 1e-15
//                  relative update (the reference iterates BiCGSTAB + ILUT to a
//                  machine-epsilon residual);
//   k_gmwb_relerr  relativeError(x_new, x_old) > tolerance.
#include <cooperative_groups.h>
#include <math_constants.h>

#include <algorithm>
#include <string>
#include <vector>

#include "qpde_kernels.cuh"
#include "qpde_partition.cuh"

namespace cg = cooperative_groups;

namespace qpde {

struct GmwbDev {
	int nw, na, nq, nz;
	const double *W, *A, *q, *z;
	double r, eta, sigma, kfee, c;
	double pen, tol;
};

// Assembled rows: W sub / diagonal / W super, coefficient of the node below in A,
// right-hand side, and up to four far impulse targets (index -1 = none).
struct GmwbRows {
	double *ra, *rb, *rc, *raa, *rhs;
	int *tgt; double *tw;       // [4][n]
};

namespace {

__device__ __forceinline__ double qmax(double x, double y) { return x > y ? x : y; }

// linearInterpolationData, Core/Interpolant.hpp:153-181
__device__ __forceinline__ void axis_interp(const double *x, int length, double ci, int &idx, double &w) {
	if(ci <= x[0]) { idx = 0; w = 1.; return; }
	if(ci >= x[length - 1]) { idx = length - 2; w = 0.; return; }
	int lo = 0, hi = length - 2, mid = 0;
	w = 0.;
	while(lo <= hi) {
		mid = (lo + hi) / 2;
		if(ci < x[mid]) hi = mid - 1;
		else if(ci >= x[mid + 1]) lo = mid + 1;
		else { w = (x[mid + 1] - ci) / (x[mid + 1] - x[mid]); break; }
	}
	idx = mid;
}

struct OpRow { double aw, cw, aa, ca, diag, b; };

// ControlledOperator::A / b for one node and one control value
__device__ __forceinline__ OpRow controlled_row(const GmwbDev &d, int iw, int ia, double q) {
	OpRow o; o.aw = o.cw = o.aa = o.ca = 0.;
	const double w = d.W[iw], a = d.A[ia];
	double total = 0.;
	{   // W direction: volatility sigma w, drift (r - eta) w - q 1{a > 0}
		const double mu = (d.r - d.eta) * w + ((0. < a) ? -q : 0.);
		if(iw == 0) {
		} else if(iw == d.nw - 1) {
			total += - mu / w;                         // HJBQVILinearBoundary
		} else {
			const double dxb = d.W[iw] - d.W[iw - 1], dxc = d.W[iw + 1] - d.W[iw - 1], dxf = d.W[iw + 1] - d.W[iw];
			const double v = d.sigma * w, vv = v * v;
			const double ac = vv / dxb / dxc, bc = vv / dxf / dxc;
			double alpha = ac - mu / dxc, beta = bc + mu / dxc;
			if(alpha < 0.) { alpha = ac; beta = bc + mu / dxf; }
			else if(beta < 0.) { alpha = ac - mu / dxb; beta = bc; }
			o.aw = -alpha; o.cw = -beta;
			total += alpha + beta;
		}
	}
	{   // A direction: no diffusion, drift -q 1{a > 0}
		const double mu = (0. < a) ? -q : 0.;
		if(ia == 0) {
		} else if(ia == d.na - 1) {
			const double alpha = - mu / (d.A[ia] - d.A[ia - 1]);   // ZeroDiffusionRightBoundary
			o.aa = -alpha;
			total += alpha;
		} else {
			const double dxb = d.A[ia] - d.A[ia - 1], dxc = d.A[ia + 1] - d.A[ia - 1], dxf = d.A[ia + 1] - d.A[ia];
			const double vv = 0. * 0.;
			const double ac = vv / dxb / dxc, bc = vv / dxf / dxc;
			double alpha = ac - mu / dxc, beta = bc + mu / dxc;
			if(alpha < 0.) { alpha = ac; beta = bc + mu / dxf; }
			else if(beta < 0.) { alpha = ac - mu / dxb; beta = bc; }
			o.aa = -alpha; o.ca = -beta;
			total += alpha + beta;
		}
	}
	o.diag = total + d.r;
	o.b = (0. < a) ? q : 0.;
	return o;
}

} // namespace

__global__ void k_gmwb_exit(GmwbDev d, double *x, int row_begin, int row_end) {
	const int row = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
	if(row >= row_end) return;
	const double w = d.W[row % d.nw], a = d.A[row / d.nw];
	x[row] = qmax(w, (1 - d.kfee) * a - d.c);                // gmwb.cpp:168
}

__global__ void __launch_bounds__(128)
k_gmwb_policy(GmwbDev d, GmwbRows R, const double *x, const double *v0, double h0, int row_begin, int row_end) {
	const int n = d.nw * d.na;
	const int row = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
	if(row >= row_end) return;
	const int nw = d.nw, iw = row % nw, ia = row / nw;

	// 1. stochastic control: argmin_q (L^q x - f^q)_row, strict <, grid order
	OpRow o; o.aw = o.cw = o.aa = o.ca = 0.; o.diag = 0.; o.b = 0.;
	{
		double best = CUDART_INF;
		for(int k = 0; k < d.nq; ++k) {
			const OpRow c = controlled_row(d, iw, ia, d.q[k]);
			double acc = 0.;
			if(ia > 0) acc += c.aa * x[row - nw];
			if(iw > 0 && iw < nw - 1) acc += c.aw * x[row - 1];
			acc += c.diag * x[row];
			if(iw > 0 && iw < nw - 1) acc += c.cw * x[row + 1];
			if(ia > 0 && ia < d.na - 1) acc += c.ca * x[row + nw];
			const double cand = acc - c.b;
			if(cand < best) { best = cand; o = c; }
		}
	}
	// 2. impulse control: argmin_z ((I - M_z) x - g_z)_row
	int t[4] = {0, 0, 0, 0}; double f[4] = {0., 0., 0., 0.}; double reward = 0.;
	double best = CUDART_INF; bool has = false;
	{
		const double w = d.W[iw], a = d.A[ia];
		for(int k = 0; k < d.nz; ++k) {
			const double z = d.z[k] * a;
			const double w_plus = qmax(w - z, 0.), a_plus = a - z;     // gmwb.cpp:139-148
			int i0, i1; double f0, f1;
			axis_interp(d.W, nw, w_plus, i0, f0);
			axis_interp(d.A, d.na, a_plus, i1, f1);
			const int tt[4] = { i0 + nw * i1, i0 + 1 + nw * i1, i0 + nw * (i1 + 1), i0 + 1 + nw * (i1 + 1) };
			const double ff[4] = { 1. * f0 * f1, 1. * (-f0 + 1.) * f1, 1. * f0 * (-f1 + 1.), 1. * (-f0 + 1.) * (-f1 + 1.) };
			const double g = (w - w_plus) + (a - a_plus) < DBL_EPSILON ? -CUDART_INF : (1 - d.kfee) * z - d.c;
			// row of I - M_z times x, accumulated in column order
			double acc = 0.; bool diag_done = false;
#pragma unroll
			for(int m = 0; m < 4; ++m) {
				if(!diag_done && row < tt[m]) { acc += 1. * x[row]; diag_done = true; }
				if(tt[m] == row) { acc += (1. - ff[m]) * x[row]; diag_done = true; }
				else acc += (0. - ff[m]) * x[tt[m]];
			}
			if(!diag_done) acc += 1. * x[row];
			const double cand = acc - g;
			if(cand < best) {
				best = cand; has = true; reward = g;
#pragma unroll
				for(int m = 0; m < 4; ++m) { t[m] = tt[m]; f[m] = ff[m]; }
			}
		}
	}
	// 3. penalty flag and the assembled row of I + h L^q* + P (I - M*)
	const bool active = has && (d.pen * best) < 0.;
	double a_ = o.aw * h0, b_ = 1. + o.diag * h0, c_ = o.cw * h0;
	double rhs = v0[row] + h0 * o.b;
	int tg[4] = {-1, -1, -1, -1}; double tw[4] = {0., 0., 0., 0.};
	if(active) {
		b_ += d.pen;
		rhs += d.pen * reward;
#pragma unroll
		for(int m = 0; m < 4; ++m) {
			const double fm = d.pen * f[m];
			if(t[m] == row) b_ -= fm;
			else if(t[m] == row - 1 && iw > 0 && iw < nw - 1) a_ -= fm;   // same line, adjacent; the last
			else if(t[m] == row + 1 && iw < nw - 2) c_ -= fm;             // W row stays decoupled in W
			else if(fm != 0.) { tg[m] = t[m]; tw[m] = fm; }
		}
	}
	R.ra[row] = a_; R.rb[row] = b_; R.rc[row] = c_; R.raa[row] = o.aa * h0; R.rhs[row] = rhs;
#pragma unroll
	for(int m = 0; m < 4; ++m) { R.tgt[m * n + row] = tg[m]; R.tw[m * n + row] = tw[m]; }
}

// ---- the line solves -------------------------------------------------------
// The system of one policy iteration is block lower-triangular by A-line: every
// entry outside a W-line points to an equal-or-lower A (drift in A is -q <= 0,
// impulses take z >= 0 out of a).  One ascending pass of line solves is therefore
// a direct solve; the only lagged entries are same-line impulse targets that are
// not W-neighbours (lines next to a = 0), and those lines are iterated in place
// until their update is below 1e-15 before the pass moves on.
//
// The lines are a serial chain, so the pass is latency-bound.  It runs on ONE
// thread-block cluster of G CTAs that take the lines round-robin: while line a - 1
// is being solved, the CTA that owns line a has already loaded its rows, folded in
// every target that lies G or more lines below (final, visible), and factorised
// its tridiagonal matrix; on the chain are only the hand-over, one fused
// multiply-add per row and the solve with the cached factors.  The hand-over is
// through distributed shared memory: the finishing CTA writes its line into the
// successor's `xin` (st.shared::cluster) and arrives on the successor's mbarrier
// (release.cluster); the global copy of the line (for far targets, later kernels
// and the other ranks) is stored before that arrive.
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_cluster(const uint64_t *bar, int cta) {
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
	asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
}

__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, int parity) {
	uint32_t done;
	do {
		asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
			: "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	} while(!done);
}

} // namespace

// Each W-line is a tridiagonal system of nw rows: M rows per thread; when
// nw == PT M + 1 the last row (decoupled in W: HJBQVILinearBoundary contributes to
// the diagonal only) is a scalar "tail" equation folded into row nw - 2.
template <int M, int PT>
__global__ void __launch_bounds__(PT, 1)
k_gmwb_lines(GmwbDev d, GmwbRows R, double *x, int ia_begin, int ia_end, int max_local, int *status) {
	extern __shared__ __align__(16) double xin[];      // [PT M + 2] the line below, written by the predecessor CTA
	__shared__ PartitionSmem sm;
	__shared__ __align__(8) uint64_t full;
	cg::cluster_group cluster = cg::this_cluster();
	const int G = (int)cluster.num_blocks(), me = (int)cluster.block_rank(), next = (me + 1) % G;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nw = d.nw, n = d.nw * d.na;
	const bool has_tail = nw == PT * M + 1;
	const int n_main = has_tail ? nw - 1 : nw;
	const bool tail_owner = has_tail && tid == PT - 1;
	if(tid == 0) mbar_init(&full, PT);
	cluster.sync();
	double *xnext = cluster.map_shared_rank(xin, next);
	PartitionSolver<M, PT> ps; ps.init();
	int received = 0;
	bool failed = false;

	for(int ia = ia_begin + me; ia < ia_end; ia += G) {
		const int base = ia * nw;
		const int far_end = (ia - G + 1) * nw;       // rows of lines <= ia - G: final, and visible through the chain
		const bool handed = ia > ia_begin;            // the line below arrives in xin (else it is final in x)
		// ---- off the chain: rows, far targets, factorisation
		double rr0[M], raa[M];
		unsigned near = 0, near_tail = 0;
		bool same_line = false;
		double rr0_t = 0., raa_t = 0., rb_t = 1., cc_t = 0.;
		{
			double aa[M], bb[M], cc[M];
#pragma unroll
			for(int j = 0; j < M; ++j) {
				const int iw = tid * M + j;
				if(iw < n_main) {
					const int row = base + iw;
					aa[j] = R.ra[row]; bb[j] = R.rb[row]; cc[j] = R.rc[row];
					raa[j] = ia > 0 ? R.raa[row] : 0.;
					double rr = R.rhs[row];
#pragma unroll
					for(int m = 0; m < 4; ++m) {
						const int tk = R.tgt[m * n + row];
						if(tk >= 0) {
							if(tk < far_end) rr += R.tw[m * n + row] * __ldcg(x + tk);
							else { near |= 1u << (4 * j + m); same_line |= tk >= base; }
						}
					}
					rr0[j] = rr;
				} else { aa[j] = 0.; bb[j] = 1.; cc[j] = 0.; rr0[j] = 0.; raa[j] = 0.; }
			}
			if(tail_owner) {
				const int row = base + nw - 1;
				raa_t = ia > 0 ? R.raa[row] : 0.; rb_t = R.rb[row];
				double rr = R.rhs[row];
#pragma unroll
				for(int m = 0; m < 4; ++m) {
					const int tk = R.tgt[m * n + row];
					if(tk >= 0) {
						if(tk < far_end) rr += R.tw[m * n + row] * __ldcg(x + tk);
						else { near_tail |= 1u << m; same_line |= tk >= base; }
					}
				}
				rr0_t = rr;
				cc_t = cc[M - 1]; cc[M - 1] = 0.;     // row nw - 2 couples to the tail through its super-diagonal
			}
			ps.factorize(aa, bb, cc, sm, lane, warp);
		}
		const bool iterate = __syncthreads_or(same_line ? 1 : 0) != 0;
		// ---- on the chain
		if(handed) { mbar_wait_cluster(&full, received & 1); ++received; }
		auto value_of = [&] (int tk) -> double {      // a near target: the line below, or x (own line / lines above far_end)
			return (handed && tk >= base - nw && tk < base) ? xin[tk - (base - nw)] : __ldcg(x + tk);
		};
		double Gs[M], x_tail = 0.;
		for(int it = 0;; ++it) {
			double r[M];
#pragma unroll
			for(int j = 0; j < M; ++j) {
				const int iw = tid * M + j;
				double rr = rr0[j];
				if(ia > 0 && iw < n_main) rr -= raa[j] * (handed ? xin[iw] : __ldcg(x + base - nw + iw));
				if((near >> (4 * j)) & 15u) {
#pragma unroll
					for(int m = 0; m < 4; ++m) if((near >> (4 * j + m)) & 1u) {
						const int row = base + iw;
						rr += R.tw[m * n + row] * value_of(R.tgt[m * n + row]);
					}
				}
				r[j] = rr;
			}
			double x_tail_old = x_tail;
			if(tail_owner) {
				const int row = base + nw - 1;
				double rr = rr0_t;
				if(ia > 0) rr -= raa_t * (handed ? xin[nw - 1] : __ldcg(x + row - nw));
#pragma unroll
				for(int m = 0; m < 4; ++m) if((near_tail >> m) & 1u) rr += R.tw[m * n + row] * value_of(R.tgt[m * n + row]);
				if(it == 0 && iterate) x_tail_old = __ldcg(x + row);
				x_tail = rr / rb_t;
				r[M - 1] = fma(-cc_t, x_tail, r[M - 1]);
			}
			double Gn[M], x_next;
			ps.solve(r, Gn, x_next, sm, lane, warp);
			bool moved = false;
			if(iterate) {                               // rare: compare with the previous iterate of this line
#pragma unroll
				for(int j = 0; j < M; ++j) {
					const int iw = tid * M + j;
					if(iw < n_main) {
						const double old = it == 0 ? __ldcg(x + base + iw) : Gs[j];
						const double den = fabs(Gn[j]) > 1. ? fabs(Gn[j]) : 1.;
						moved |= fabs(Gn[j] - old) > 1e-15 * den;
					}
				}
				if(tail_owner) {
					const double den = fabs(x_tail) > 1. ? fabs(x_tail) : 1.;
					moved |= fabs(x_tail - x_tail_old) > 1e-15 * den;
				}
			}
#pragma unroll
			for(int j = 0; j < M; ++j) {
				const int iw = tid * M + j;
				Gs[j] = Gn[j];
				if(iw < n_main) x[base + iw] = Gn[j];
			}
			if(tail_owner) x[base + nw - 1] = x_tail;
			if(!iterate) break;
			__threadfence_block();
			if(!__syncthreads_or(moved ? 1 : 0)) break;
			if(it + 1 >= max_local) { failed = true; break; }
		}
		// ---- hand the line to the CTA that owns the next one
		if(ia + 1 < ia_end) {
			double2 *dst = reinterpret_cast<double2 *>(xnext + tid * M);
#pragma unroll
			for(int j = 0; j < M; j += 2) dst[j / 2] = make_double2(Gs[j], Gs[j + 1]);
			if(tail_owner) xnext[nw - 1] = x_tail;
			mbar_arrive_cluster(&full, next);
		}
	}
	if(failed && tid == 0 && status) atomicOr(status, 2);
	cluster.sync();                                     // nobody leaves while a peer may still write into its xin
}

__global__ void k_gmwb_relerr(int row_begin, int row_end, const double *x, const double *xprev, double scale, double tol, int *notdone) {
	const int i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
	bool nd = false;
	if(i < row_end) {
		const double fa = fabs(x[i]), fb = fabs(xprev[i]);
		double den = fa < fb ? fb : fa;
		den = scale < den ? den : scale;
		nd = fabs(x[i] - xprev[i]) / den > tol;                 // IterativeMethod.hpp:1131-1135
	}
	if(__syncthreads_or(nd ? 1 : 0) && threadIdx.x == 0) atomicOr(notdone, 1);
}

// One rank's share of the problem: the A-lines [a_begin, a_end);
// the single-GPU solve is the shard [0, n_a).
struct GmwbShard {
	cudaStream_t stream;
	long long *launches;
	GmwbDev d;
	GmwbRows R;
	int a_begin, a_end, PT, max_local;
	int *flags;                 // [0] not converged, [1] solve status
	std::vector<void *> allocs;
};

static const int GMWB_CLUSTER = 8;       // CTAs in the cluster (portable maximum): depth of the line pipeline

template <int PT>
static cudaError_t launch_lines_t(GmwbShard &sh, double *x, int ia0, int ia1, int max_local, int *status) {
	const size_t smem = sizeof(double) * (PT * 8 + 2);
	cudaError_t e = cudaFuncSetAttribute(k_gmwb_lines<8, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if(e != cudaSuccess) return e;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(GMWB_CLUSTER); cfg.blockDim = dim3(PT); cfg.dynamicSmemBytes = smem; cfg.stream = sh.stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = GMWB_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, k_gmwb_lines<8, PT>, sh.d, sh.R, x, ia0, ia1, max_local, status);
}

static cudaError_t launch_lines(GmwbShard &sh, double *x, int ia0, int ia1, int *status) {
	++*sh.launches;
	switch(sh.PT) {
	case 32:  return launch_lines_t<32>(sh, x, ia0, ia1, sh.max_local, status);
	case 64:  return launch_lines_t<64>(sh, x, ia0, ia1, sh.max_local, status);
	case 128: return launch_lines_t<128>(sh, x, ia0, ia1, sh.max_local, status);
	case 256: return launch_lines_t<256>(sh, x, ia0, ia1, sh.max_local, status);
	default:  return launch_lines_t<512>(sh, x, ia0, ia1, sh.max_local, status);
	}
}

void gmwb_shard_destroy(GmwbShard *sh) {
	if(!sh) return;
	for(void *ptr : sh->allocs) cudaFree(ptr);
	delete sh;
}

int gmwb_shard_create(const qpde_gmwb2d *p, const double *W, int nw, const double *A, int na, const double *z, int nz,
		int a_begin, int a_end, cudaStream_t stream, long long *launches, GmwbShard **out, std::string *error) {
	const int n = nw * na;
	int PT = 32;
	while(PT * 8 + 1 < nw && PT < 512) PT *= 2;
	if(PT * 8 + 1 < nw) { *error = "gmwb: W axis longer than 4097 nodes is not supported"; return QPDE_ERR_UNSUPPORTED; }
	if(a_begin < 0 || a_end > na || a_begin > a_end) { *error = "gmwb: bad A-line range"; return QPDE_ERR_INVALID; }
	// the block-triangular structure the line pass relies on
	for(int k = 0; k < p->n_controls; ++k) if(!(p->controls[k] >= 0.)) { *error = "gmwb: withdrawal rates must be >= 0"; return QPDE_ERR_INVALID; }
	for(int k = 0; k < nz; ++k) if(!(z[k] >= 0. && z[k] <= 1.)) { *error = "gmwb: impulse fractions must lie in [0, 1]"; return QPDE_ERR_INVALID; }
	GmwbShard *sh = new GmwbShard();
	sh->stream = stream; sh->launches = launches; sh->a_begin = a_begin; sh->a_end = a_end; sh->PT = PT; sh->max_local = p->max_sweeps > 0 ? p->max_sweeps : 500;
	bool ok = true;
	auto dalloc = [&] (size_t bytes) -> void * {
		void *ptr = nullptr;
		if(cudaMalloc(&ptr, bytes ? bytes : 8) != cudaSuccess) { ok = false; return nullptr; }
		sh->allocs.push_back(ptr);
		return ptr;
	};
	double *dW = (double *)dalloc(sizeof(double) * nw), *dA = (double *)dalloc(sizeof(double) * na);
	double *dq = (double *)dalloc(sizeof(double) * p->n_controls), *dz = (double *)dalloc(sizeof(double) * nz);
	GmwbRows &R = sh->R;
	R.ra = (double *)dalloc(sizeof(double) * n); R.rb = (double *)dalloc(sizeof(double) * n); R.rc = (double *)dalloc(sizeof(double) * n);
	R.raa = (double *)dalloc(sizeof(double) * n); R.rhs = (double *)dalloc(sizeof(double) * n);
	R.tgt = (int *)dalloc(sizeof(int) * 4 * (size_t)n); R.tw = (double *)dalloc(sizeof(double) * 4 * (size_t)n);
	sh->flags = (int *)dalloc(sizeof(int) * 4);
	if(!ok) { gmwb_shard_destroy(sh); *error = "gmwb: out of device memory"; return QPDE_ERR_CUDA; }
	cudaMemcpyAsync(dW, W, sizeof(double) * nw, cudaMemcpyHostToDevice, stream);
	cudaMemcpyAsync(dA, A, sizeof(double) * na, cudaMemcpyHostToDevice, stream);
	cudaMemcpyAsync(dq, p->controls, sizeof(double) * p->n_controls, cudaMemcpyHostToDevice, stream);
	cudaMemcpyAsync(dz, z, sizeof(double) * nz, cudaMemcpyHostToDevice, stream);
	cudaMemsetAsync(sh->flags, 0, sizeof(int) * 4, stream);
	cudaStreamSynchronize(stream);      // the host arrays may go away
	GmwbDev &d = sh->d;
	d.nw = nw; d.na = na; d.nq = p->n_controls; d.nz = nz;
	d.W = dW; d.A = dA; d.q = dq; d.z = dz;
	d.r = p->r; d.eta = p->eta; d.sigma = p->sigma; d.kfee = p->kappa; d.c = p->c;
	d.pen = 1. / (p->scaling_factor * (p->T / p->timesteps));   // HJBQVI.hpp:635, PenaltyMethod.hpp:85
	d.tol = p->tolerance;
	*out = sh;
	return QPDE_OK;
}

// exit values V(T) on the shard's lines, written into the full-grid vector x
void gmwb_shard_exit(GmwbShard *sh, double *x) {
	const int r0 = sh->a_begin * sh->d.nw, r1 = sh->a_end * sh->d.nw;
	if(r1 > r0) { k_gmwb_exit<<<(r1 - r0 + 255) / 256, 256, 0, sh->stream>>>(sh->d, x, r0, r1); ++*sh->launches; }
}

// control searches + row assembly for the shard's lines, from the full previous iterate
void gmwb_shard_policy(GmwbShard *sh, const double *x, const double *v0, double h0) {
	const int r0 = sh->a_begin * sh->d.nw, r1 = sh->a_end * sh->d.nw;
	if(r1 > r0) { k_gmwb_policy<<<(r1 - r0 + 127) / 128, 128, 0, sh->stream>>>(sh->d, sh->R, x, v0, h0, r0, r1); ++*sh->launches; }
}

// the shard's line solves in ascending A (exact given final lines below a_begin);
// *status_dev |= 2 if a line's in-place iteration did not settle
cudaError_t gmwb_shard_sweep(GmwbShard *sh, double *x, int *status_dev) {
	if(sh->a_end > sh->a_begin) return launch_lines(*sh, x, sh->a_begin, sh->a_end, status_dev);
	return cudaSuccess;
}

// relativeError(x, xprev) > tolerance on the shard's lines; *notdone_dev |= 1
void gmwb_shard_relerr(GmwbShard *sh, const double *x, const double *xprev, int *notdone_dev) {
	const int r0 = sh->a_begin * sh->d.nw, r1 = sh->a_end * sh->d.nw;
	if(r1 > r0) {
		k_gmwb_relerr<<<(r1 - r0 + 255) / 256, 256, 0, sh->stream>>>(r0, r1, x, xprev, 1., sh->d.tol, notdone_dev);
		++*sh->launches;
	}
}

// Host driver of the single-GPU solve: the reference's time loop and
// ToleranceIteration around the kernels, on the shard [0, n_a).
int gmwb_run(const qpde_gmwb2d *p, const double *W, int nw, const double *A, int na, const double *z, int nz,
		double *V_host, qpde_gmwb2d_stats *stats, cudaStream_t stream, long long *launches, std::string *error) {
	const int n = nw * na;
	GmwbShard *sh = nullptr;
	int rc = gmwb_shard_create(p, W, nw, A, na, z, nz, 0, na, stream, launches, &sh, error);
	if(rc != QPDE_OK) return rc;
	auto dalloc = [&] (size_t bytes) -> double * {
		void *ptr = nullptr;
		if(cudaMalloc(&ptr, bytes) != cudaSuccess) return nullptr;
		sh->allocs.push_back(ptr);
		return (double *)ptr;
	};
	double *x = dalloc(sizeof(double) * n), *xprev = dalloc(sizeof(double) * n), *v0 = dalloc(sizeof(double) * n);
	if(!x || !xprev || !v0) { gmwb_shard_destroy(sh); *error = "gmwb: out of device memory"; return QPDE_ERR_CUDA; }
	int *flags = sh->flags;
	gmwb_shard_exit(sh, x);

	Replay rp; rp.start(p->T, p->T / p->timesteps, 0);
	double time1 = p->T;
	long inner_sum = 0; int n_steps = 0, status = 0;
	bool more = true;
	while(more) {
		qpde_step st;
		more = rp.next(st);
		const double h0 = time1 - st.t;
		cudaMemcpyAsync(v0, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream);
		int its = 0, notdone = 0;
		do {
			cudaMemcpyAsync(xprev, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream);
			gmwb_shard_policy(sh, xprev, v0, h0);
			cudaError_t le = launch_lines(*sh, x, 0, na, flags + 1);
			cudaMemsetAsync(flags, 0, sizeof(int), stream);
			gmwb_shard_relerr(sh, x, xprev, flags);
			int host_flags[2] = {0, 0};
			cudaError_t e = le != cudaSuccess ? le : cudaMemcpyAsync(host_flags, flags, sizeof(int) * 2, cudaMemcpyDeviceToHost, stream);
			if(e == cudaSuccess) e = cudaStreamSynchronize(stream);
			if(e != cudaSuccess) {
				gmwb_shard_destroy(sh);
				*error = std::string("gmwb iteration: ") + cudaGetErrorString(e);
				return QPDE_ERR_CUDA;
			}
			notdone = host_flags[0];
			if(host_flags[1]) status = 2;
			++its;
			if(notdone && its >= p->max_inner) { if(!status) status = 1; break; }
		} while(notdone);
		if(stats && stats->inner && n_steps < stats->inner_capacity) stats->inner[n_steps] = its;
		inner_sum += its;
		++n_steps;
		time1 = st.t;
	}
	cudaError_t e = cudaMemcpyAsync(V_host, x, sizeof(double) * n, cudaMemcpyDeviceToHost, stream);
	if(e == cudaSuccess) e = cudaStreamSynchronize(stream);
	gmwb_shard_destroy(sh);
	if(e != cudaSuccess) { *error = std::string("gmwb fetch: ") + cudaGetErrorString(e); return QPDE_ERR_CUDA; }
	if(stats) { stats->n_steps = n_steps; stats->mean_inner = (double)inner_sum / n_steps; stats->status = status; stats->nodes_w = nw; stats->nodes_a = na; }
	return QPDE_OK;
}

} // namespace qpde
