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
#include <cuda.h>
#include <math_constants.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
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
	// bucket tables for the interpolation searches of the impulse step: lutW[b] = largest node
	// index i with W[i] <= b / invhW (same for A); GMWB_LUT buckets over [0, axis end]
	const int *lutW, *lutA;
	double invhW, invhA;
};
constexpr int GMWB_LUT = 1 << 16;

// Assembled rows: W sub / diagonal / W super, coefficient of the node below in A,
// right-hand side, and up to four far impulse targets (index -1 = none).
// The right-hand side is kept in two terms, rhs = h b(q*) and rhs2 = penalty x reward, and the line
// pass forms (V^n + rhs) + rhs2 — the reference's order — so that the control search of an
// iteration does not depend on which V^n its line pass will run with (gmwb_run overlaps the two).
struct GmwbRows {
	double *ra, *rb, *rc, *raa, *rhs, *rhs2;
	int *tgt; double *tw;       // [4][n]
};

namespace {

__device__ __forceinline__ double qmax(double x, double y) { return x > y ? x : y; }

// linearInterpolationData, Core/Interpolant.hpp:153-181
// The same through a bucket table: a start index not above the bracket, then a walk of a few
// nodes instead of a bisection (same idx, same weight).  Needs x[0] == 0 and lut from make_lut().
__device__ __forceinline__ void axis_interp_lut(const double *x, int length, const int *lut, double invh,
		double ci, int &idx, double &w) {
	if(ci <= x[0]) { idx = 0; w = 1.; return; }
	if(ci >= x[length - 1]) { idx = length - 2; w = 0.; return; }
	int bkt = (int)(ci * invh);
	bkt = bkt < GMWB_LUT - 1 ? bkt : GMWB_LUT - 1;
	int i = lut[bkt];
	while(x[i + 1] <= ci) ++i;            // x[i] <= ci < x[i + 1]; terminates: ci < x[length - 1]
	idx = i;
	w = (x[i + 1] - ci) / (x[i + 1] - x[i]);
}

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

// MODE 0: control search + row assembly (one GPU).  Sharded over ranks the two halves run apart:
// MODE 1 searches and leaves one word per node — the index of the stochastic control (bits 0 .. 7, 255 =
// none) and, where the penalty is active, the index of the impulse candidate + 1 (bits 8 ..) — which is
// all a rank has to send; MODE 2 assembles the rows of any node from such a word with the expressions
// of the search, hence to the same bits.
enum { POLICY_FUSED = 0, POLICY_SEARCH = 1, POLICY_ASSEMBLE = 2 };

template <int MODE>
__global__ void __launch_bounds__(128)
k_gmwb_policy(GmwbDev d, GmwbRows R, const double *x, double h0, int row_begin, int row_end, int *decision) {
	const int n = d.nw * d.na;
	const int row = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
	const int nw = d.nw;
	// The A-direction half of every impulse candidate — z = z_k a, a - z and its bracket on the A
	// axis — depends on (line, k) only: formed once per CTA for the lines its rows touch.
	extern __shared__ __align__(16) unsigned char pol_smem[];
	const int lines = blockDim.x / nw + 2;
	double *const tab_z = reinterpret_cast<double *>(pol_smem);        // [lines][nz]
	double *const tab_f1 = tab_z + lines * d.nz;
	int *const tab_i1 = reinterpret_cast<int *>(tab_f1 + lines * d.nz);
	const int ia_first = (row_begin + blockIdx.x * blockDim.x) / nw;
	for(int e = threadIdx.x; e < lines * d.nz; e += blockDim.x) {
		const int ial = ia_first + e / d.nz, k = e % d.nz;
		if(ial < d.na) {
			const double a = d.A[ial], z = d.z[k] * a;
			int i1; double f1;
			axis_interp_lut(d.A, d.na, d.lutA, d.invhA, a - z, i1, f1);
			tab_z[e] = z; tab_f1[e] = f1; tab_i1[e] = i1;
		}
	}
	__syncthreads();
	if(row >= row_end) return;
	const int iw = row % nw, ia = row / nw;
	const int tab0 = (ia - ia_first) * d.nz;

	// 1. stochastic control: argmin_q (L^q x - f^q)_row, strict <, grid order
	OpRow o; o.aw = o.cw = o.aa = o.ca = 0.; o.diag = 0.; o.b = 0.;
	const int word = MODE == POLICY_ASSEMBLE ? decision[row] : 0;
	int kq_sel = 255, kz_sel = -1;
	if(MODE == POLICY_ASSEMBLE) {
		kq_sel = word & 255;
		if(kq_sel != 255) o = controlled_row(d, iw, ia, d.q[kq_sel]);
	} else {
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
			if(cand < best) { best = cand; o = c; kq_sel = k; }
		}
	}
	// 2. impulse control: argmin_z ((I - M_z) x - g_z)_row
	int t[4] = {0, 0, 0, 0}; double f[4] = {0., 0., 0., 0.}; double reward = 0.;
	double best = CUDART_INF; bool has = false;
	{
		const double w = d.W[iw], a = d.A[ia];
		// the search walks every candidate; the assembly re-forms the chosen one (none where the penalty is off)
		const int chosen = (word >> 8) - 1;                       // -1: the penalty is off at this node
		const int k_begin = MODE == POLICY_ASSEMBLE && chosen >= 0 ? chosen : 0;
		const int k_end = MODE == POLICY_ASSEMBLE ? (chosen >= 0 ? chosen + 1 : 0) : d.nz;
		for(int k = k_begin; k < k_end; ++k) {
			const double z = tab_z[tab0 + k];
			const double w_plus = qmax(w - z, 0.), a_plus = a - z;     // gmwb.cpp:139-148
			int i0; double f0;
			const int i1 = tab_i1[tab0 + k]; const double f1 = tab_f1[tab0 + k];
			axis_interp_lut(d.W, nw, d.lutW, d.invhW, w_plus, i0, f0);
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
			if(MODE == POLICY_ASSEMBLE || cand < best) {
				best = cand; has = true; reward = g; kz_sel = k;
#pragma unroll
				for(int m = 0; m < 4; ++m) { t[m] = tt[m]; f[m] = ff[m]; }
			}
		}
	}
	// 3. penalty flag and the assembled row of I + h L^q* + P (I - M*)
	const bool active = MODE == POLICY_ASSEMBLE ? has : has && (d.pen * best) < 0.;
	if(MODE == POLICY_SEARCH) {
		decision[row] = kq_sel | (active ? (kz_sel + 1) << 8 : 0);
		return;
	}
	double a_ = o.aw * h0, b_ = 1. + o.diag * h0, c_ = o.cw * h0;
	const double rhs = h0 * o.b;
	double rhs2 = 0.;
	int tg[4] = {-1, -1, -1, -1}; double tw[4] = {0., 0., 0., 0.};
	if(active) {
		b_ += d.pen;
		rhs2 = d.pen * reward;
#pragma unroll
		for(int m = 0; m < 4; ++m) {
			const double fm = d.pen * f[m];
			if(t[m] == row) b_ -= fm;
			else if(t[m] == row - 1 && iw > 0 && iw < nw - 1) a_ -= fm;   // same line, adjacent; the last
			else if(t[m] == row + 1 && iw < nw - 2) c_ -= fm;             // W row stays decoupled in W
			else if(fm != 0.) { tg[m] = t[m]; tw[m] = fm; }
		}
	}
	R.ra[row] = a_; R.rb[row] = b_; R.rc[row] = c_; R.raa[row] = o.aa * h0; R.rhs[row] = rhs; R.rhs2[row] = rhs2;
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
// thread-block cluster of G CTAs that take the lines round-robin, and everything
// that does not need the two lines below is done before they are finished:
//   stage 1  rows of line a, targets on lines <= a - G - 3 (visible since this
//            CTA's previous line), factorisation of the tridiagonal matrix;
//   stage 2  after the `pre` barrier (lines <= a - 3 are in x and visible): the
//            remaining targets on those lines;
//   chain    wait for lines a - 1 and a - 2 in shared memory (`full_a`, `full_b`),
//            one multiply-add per row (plus the rare targets on those two lines),
//            the solve with the cached factors, hand-over.
// Hand-over through distributed shared memory: the finishing CTA stores its line
// into `xa` of the next CTA and `xb` of the one after it (st.shared::cluster) and
// arrives once per warp on their barriers (release.cluster) — no global memory on
// the chain.  Afterwards, off the chain, the line is written to x and released to
// the CTA three lines up (`pre`); thread 0 of every CTA forwards its own `pre` to
// its successor, so that by induction `pre`(a) covers every line <= a - 3.
//
// Two things ptxas / the hardware punish in this kernel, both measured:
//  * a warp that reaches a shuffle while diverged (an `if(tid == ...)` block in
//    front of the solver is enough) runs every shuffle of the solver on the
//    WARPSYNC slow path — 10x on the factorisation.  All single-thread work here
//    is predicated instead of branched, and the tail row is formed by every thread.
//  * a predicate that lives across a gather load makes ptxas issue the loads one
//    after the other (seven predicate registers).  Targets that are not wanted
//    read x[0] with weight 0 instead.
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ uint32_t cluster_addr(const void *p, int cta) {
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(p)), "r"(cta));
	return remote;
}

// `elected` is a predicate, not a branch (see above)
__device__ __forceinline__ void mbar_arrive_cluster(const uint64_t *bar, int cta, bool elected = true) {
	const uint32_t remote = cluster_addr(bar, cta);
	asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n @p mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n}"
		:: "r"(remote), "r"((unsigned)elected) : "memory");
}

__device__ __forceinline__ void mbar_arrive_local(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, int parity) {
	uint32_t done;
	do {
		asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
			: "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	} while(!done);
}

// store into a peer's shared memory
__device__ __forceinline__ void st_cluster_f64(uint32_t remote_addr, double v) {
	asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(remote_addr), "d"(v) : "memory");
}

} // namespace

__host__ __device__ constexpr int gmwb_near_k(int PT) { return PT > 256 ? 4 : 8; }   // listed targets per thread (shared-memory budget)

// Each W-line is a tridiagonal system of nw rows: M rows per thread; when
// nw == PT M + 1 the last row (decoupled in W: HJBQVILinearBoundary contributes to
// the diagonal only) is a scalar "tail" equation folded into row nw - 2.
template <int M, int PT>
__global__ void __launch_bounds__(PT, 1)
k_gmwb_lines(GmwbDev d, GmwbRows R, double *x, const double *v0, const double *xguess, int ia_begin, int ia_end, int max_local, int *status) {
	// dynamic shared memory, each line buffer [M PT + 2]: row iw = tid M + j at j PT + tid, the
	// tail row at M PT.  xa / xb: lines ia - 1 / ia - 2, written by the CTAs that solved them;
	// rhs: the right-hand side under construction; then this thread's list of targets that
	// stage 1 could not fold in
	constexpr int LINE = M * PT + 2, NEAR_K = gmwb_near_k(PT), TK_BITS = 27, TK_MASK = (1 << TK_BITS) - 1;
	extern __shared__ __align__(16) unsigned char dyn_smem[];
	double *xa = reinterpret_cast<double *>(dyn_smem), *xb = xa + LINE, *rhs = xb + LINE;
	double *nl_tw = rhs + LINE;                                         // [NEAR_K][PT] weight
	int *nl_tk = reinterpret_cast<int *>(nl_tw + NEAR_K * PT);        // [NEAR_K][PT] target row | row slot j << 27
	__shared__ PartitionSmem sm;
	__shared__ __align__(8) uint64_t full_a, full_b, pre;
	cg::cluster_group cluster = cg::this_cluster();
	const int G = (int)cluster.num_blocks(), me = (int)cluster.block_rank();
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nw = d.nw, n = d.nw * d.na;
	const bool has_tail = nw == PT * M + 1;
	const int n_main = has_tail ? nw - 1 : nw;
	const bool tail_owner = has_tail && tid == PT - 1;
	if(tid == 0) { mbar_init(&full_a, PT / 32); mbar_init(&full_b, PT / 32); mbar_init(&pre, PT + 1); }
	cluster.sync();
	const uint32_t xa_next = cluster_addr(xa, (me + 1) % G), xb_next2 = cluster_addr(xb, (me + 2) % G);
	PartitionSolver<M, PT> ps; ps.init();
	int n_full_a = 0, n_full_b = 0, n_pre = 0;
	bool failed = false;
	auto slot = [&] (int iw) -> int { return iw == M * PT ? M * PT : (iw % M) * PT + iw / M; };

	for(int ia = ia_begin + me; ia < ia_end; ia += G) {
		const int base = ia * nw;
		const int far_end = (ia - G - 2) * nw;       // rows of lines <= ia - G - 3: visible since this CTA's previous line
		const int lower_end = base - 2 * nw;          // rows of lines <= ia - 3: visible after `pre`
		const bool has_a = ia - 1 >= ia_begin;        // line ia - 1 arrives in xa (else it is final in x)
		const bool has_b = ia - 2 >= ia_begin;        // line ia - 2 arrives in xb (else it is final in x)
		// ---- stage 1: rows, far targets, factorisation
		double raa[M];
		unsigned later = 0, later_tail = 0;           // bit 4 j + m: target m of row j is not folded in yet
		double raa_t = 0., rb_t = 1., rinv_t = 1., cc_t = 0.;
		{
			double aa[M], bb[M], cc[M];
#pragma unroll
			for(int j0 = 0; j0 < M; j0 += 4) {
				int tk[16]; double tw[16], xv[16], rr[4];
#pragma unroll
				for(int jj = 0; jj < 4; ++jj) {
					const int j = j0 + jj, iw = tid * M + j, row = base + (iw < n_main ? iw : 0);
					aa[j] = R.ra[row]; bb[j] = R.rb[row]; cc[j] = R.rc[row]; raa[j] = R.raa[row]; rr[jj] = (v0[row] + R.rhs[row]) + R.rhs2[row];
#pragma unroll
					for(int m = 0; m < 4; ++m) { tk[4 * jj + m] = R.tgt[m * n + row]; tw[4 * jj + m] = R.tw[m * n + row]; }
				}
#pragma unroll
				for(int e = 0; e < 16; ++e) {
					const bool real = tid * M + j0 + e / 4 < n_main;
					const bool far = tk[e] >= 0 && tk[e] < far_end;
					later |= real && tk[e] >= far_end && tk[e] >= 0 ? 1u << (4 * j0 + e) : 0u;
					tk[e] = far ? tk[e] : 0;
					tw[e] = far ? tw[e] : 0.;
				}
				// published before the acquire at this CTA's previous `pre`: ordinary loads (ptxas
				// never overlaps ld.cg loads)
#pragma unroll
				for(int e = 0; e < 16; ++e) xv[e] = x[tk[e]];
#pragma unroll
				for(int jj = 0; jj < 4; ++jj) {
					const int j = j0 + jj;
					const bool real = tid * M + j < n_main;
#pragma unroll
					for(int m = 0; m < 4; ++m) rr[jj] += tw[4 * jj + m] * xv[4 * jj + m];
					aa[j] = real ? aa[j] : 0.; bb[j] = real ? bb[j] : 1.; cc[j] = real ? cc[j] : 0.;
					raa[j] = real && ia > 0 ? raa[j] : 0.;
					rhs[j * PT + tid] = real ? rr[jj] : 0.;
				}
			}
			if(has_tail) {       // uniform: every thread forms the tail row (broadcast loads), its owner uses it
				const int row = base + nw - 1;
				raa_t = ia > 0 ? R.raa[row] : 0.; rb_t = R.rb[row]; rinv_t = 1. / rb_t;
				double rr = (v0[row] + R.rhs[row]) + R.rhs2[row];
				int tk[4]; double tw[4], xv[4];
#pragma unroll
				for(int m = 0; m < 4; ++m) { tk[m] = R.tgt[m * n + row]; tw[m] = R.tw[m * n + row]; }
#pragma unroll
				for(int m = 0; m < 4; ++m) {
					const bool far = tk[m] >= 0 && tk[m] < far_end;
					later_tail |= tail_owner && tk[m] >= far_end && tk[m] >= 0 ? 1u << m : 0u;
					tk[m] = far ? tk[m] : 0; tw[m] = far ? tw[m] : 0.;
				}
#pragma unroll
				for(int m = 0; m < 4; ++m) xv[m] = x[tk[m]];
#pragma unroll
				for(int m = 0; m < 4; ++m) rr += tw[m] * xv[m];
				rhs[M * PT] = rr;                              // every thread writes the same value
				cc_t = tail_owner ? cc[M - 1] : 0.;
				cc[M - 1] = tail_owner ? 0. : cc[M - 1];     // row nw - 2 couples to the tail through its super-diagonal
			}
			ps.factorize(aa, bb, cc, sm, lane, warp);
		}
		// the marked targets go to this thread's list: (row slot in rhs, target, weight), in (j, m) order
		int n_list = 0; unsigned over = 0, over_tail = 0; bool same_line = false;
#pragma unroll 1
		for(unsigned bits = later; bits; bits &= bits - 1) {
			const int b = __ffs(bits) - 1, row = base + tid * M + (b >> 2);
			const int tk = R.tgt[(b & 3) * n + row];
			same_line |= tk >= base;
			if(n_list < NEAR_K) { nl_tk[n_list * PT + tid] = tk | ((b >> 2) << TK_BITS); nl_tw[n_list * PT + tid] = R.tw[(b & 3) * n + row]; ++n_list; }
			else over |= 1u << b;
		}
#pragma unroll 1
		for(unsigned bits = later_tail; bits; bits &= bits - 1) {
			const int b = __ffs(bits) - 1, row = base + nw - 1;
			const int tk = R.tgt[b * n + row];
			same_line |= tk >= base;
			if(n_list < NEAR_K) { nl_tk[n_list * PT + tid] = tk | (M << TK_BITS); nl_tw[n_list * PT + tid] = R.tw[b * n + row]; ++n_list; }
			else over_tail |= 1u << b;
		}
		const bool iterate = __syncthreads_or(same_line ? 1 : 0) != 0;
		if(iterate && xguess != x) {
			// a line that is iterated in place starts from the previous iterate of this line
#pragma unroll
			for(int j = 0; j < M; ++j) if(tid * M + j < n_main) x[base + tid * M + j] = xguess[base + tid * M + j];
			if(tail_owner) x[base + nw - 1] = xguess[base + nw - 1];
			__threadfence_block();
			__syncthreads();
		}
		// ---- stage 2: lines <= ia - 3 are published
		if(has_a) {
			if(ia - 3 < ia_begin) mbar_arrive_local(&pre);         // nobody owns line ia - 3 in this pass
			mbar_wait_cluster(&pre, n_pre & 1); ++n_pre;
		}
		mbar_arrive_cluster(&pre, (me + 1) % G, tid == 0 && ia + 1 < ia_end);   // forwards what this CTA has seen
		{
			// fold the listed targets on lines <= ia - 3 into rhs; keep the rest, in order
			int kept = 0;
#pragma unroll 1
			for(int e0 = 0; e0 < n_list; e0 += 4) {
				int tk[4], t[4]; double p[4], xv[4];
#pragma unroll
				for(int k = 0; k < 4; ++k) {
					const bool live = e0 + k < n_list;
					tk[k] = live ? nl_tk[(e0 + k) * PT + tid] : base;
					p[k] = live ? nl_tw[(e0 + k) * PT + tid] : 0.;
					t[k] = tk[k] & TK_MASK;
				}
#pragma unroll
				for(int k = 0; k < 4; ++k) xv[k] = x[t[k] < lower_end ? t[k] : 0];
#pragma unroll
				for(int k = 0; k < 4; ++k) if(e0 + k < n_list) {
					const int j = tk[k] >> TK_BITS;
					if(t[k] < lower_end) rhs[(j == M ? M * PT : j * PT + tid)] += p[k] * xv[k];
					else { nl_tk[kept * PT + tid] = tk[k]; nl_tw[kept * PT + tid] = p[k]; ++kept; }
				}
			}
			n_list = kept;
		}
		// ---- on the chain
		if(has_b) { mbar_wait_cluster(&full_b, n_full_b & 1); ++n_full_b; }
		if(has_a) { mbar_wait_cluster(&full_a, n_full_a & 1); ++n_full_a; }
		auto value_of = [&] (int t) -> double {       // a target on line ia - 2, ia - 1 or ia
			if(t >= base) return __ldcg(x + t);
			if(t >= base - nw) return has_a ? xa[slot(t - (base - nw))] : __ldcg(x + t);
			return has_b ? xb[slot(t - lower_end)] : __ldcg(x + t);
		};
		double Gs[M], x_tail = 0.;
		for(int it = 0;; ++it) {
			double r[M];
#pragma unroll
			for(int j = 0; j < M; ++j) {
				const int iw = tid * M + j;
				const double below = has_a ? xa[j * PT + tid] : __ldcg(x + (ia > 0 && iw < n_main ? base - nw + iw : 0));
				r[j] = rhs[j * PT + tid] - raa[j] * below;
			}
			double r_t = rhs[M * PT] - raa_t * (has_a ? xa[M * PT] : __ldcg(x + (ia > 0 ? base - 1 : 0)));   // uniform; used by the tail owner
			if(__any_sync(0xffffffffu, (n_list > 0) | (over != 0) | (over_tail != 0))) {
				double add[M + 1];
#pragma unroll
				for(int j = 0; j <= M; ++j) add[j] = 0.;
#pragma unroll 1
				for(int e = 0; e < n_list; ++e) {
					const int t = nl_tk[e * PT + tid] & TK_MASK, je = nl_tk[e * PT + tid] >> TK_BITS;
					const double p = nl_tw[e * PT + tid] * value_of(t);
#pragma unroll
					for(int j = 0; j <= M; ++j) add[j] += je == j ? p : 0.;
				}
#pragma unroll 1
				for(unsigned bits = over; bits; bits &= bits - 1) {      // beyond the list: re-read the row
					const int b = __ffs(bits) - 1, row = base + tid * M + (b >> 2);
					const int t = R.tgt[(b & 3) * n + row];
					const double p = R.tw[(b & 3) * n + row] * (t < lower_end ? __ldcg(x + t) : value_of(t));
#pragma unroll
					for(int j = 0; j < M; ++j) add[j] += (b >> 2) == j ? p : 0.;
				}
#pragma unroll 1
				for(unsigned bits = over_tail; bits; bits &= bits - 1) {
					const int b = __ffs(bits) - 1, row = base + nw - 1;
					const int t = R.tgt[b * n + row];
					add[M] += R.tw[b * n + row] * (t < lower_end ? __ldcg(x + t) : value_of(t));
				}
#pragma unroll
				for(int j = 0; j < M; ++j) r[j] += add[j];
				r_t += add[M];
				__syncwarp();
			}
			double x_tail_old = x_tail;
			if(it == 0 && iterate && tail_owner) x_tail_old = __ldcg(x + base + nw - 1);
			x_tail = r_t * rinv_t;
			x_tail = fma(fma(-rb_t, x_tail, r_t), rinv_t, x_tail);       // r_t / rb_t to the last bit or so, without the division
			r[M - 1] = tail_owner ? fma(-cc_t, x_tail, r[M - 1]) : r[M - 1];
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
			for(int j = 0; j < M; ++j) Gs[j] = Gn[j];
			if(!iterate) break;
#pragma unroll
			for(int j = 0; j < M; ++j) if(tid * M + j < n_main) x[base + tid * M + j] = Gs[j];
			if(tail_owner) x[base + nw - 1] = x_tail;
			__threadfence_block();
			if(!__syncthreads_or(moved ? 1 : 0)) break;
			if(it + 1 >= max_local) { failed = true; break; }
		}
		// ---- hand the line to the CTA that owns the next one, then to the one after it
		if(ia + 1 < ia_end) {
#pragma unroll
			for(int j = 0; j < M; ++j) st_cluster_f64(xa_next + 8u * (j * PT + tid), Gs[j]);
			st_cluster_f64(xa_next + 8u * (M * PT + (tail_owner ? 0 : 1)), x_tail);   // slot M PT + 1 is scratch
		}
		__syncwarp();
		mbar_arrive_cluster(&full_a, (me + 1) % G, lane == 0 && ia + 1 < ia_end);
		if(ia + 2 < ia_end) {
#pragma unroll
			for(int j = 0; j < M; ++j) st_cluster_f64(xb_next2 + 8u * (j * PT + tid), Gs[j]);
			st_cluster_f64(xb_next2 + 8u * (M * PT + (tail_owner ? 0 : 1)), x_tail);
		}
		__syncwarp();
		mbar_arrive_cluster(&full_b, (me + 2) % G, lane == 0 && ia + 2 < ia_end);
		// ---- off the chain again: publish the line in x, release it to the CTA three lines up
#pragma unroll
		for(int j = 0; j < M; ++j) if(tid * M + j < n_main) x[base + tid * M + j] = Gs[j];
		if(tail_owner) x[base + nw - 1] = x_tail;
		mbar_arrive_cluster(&pre, (me + 3) % G, ia + 3 < ia_end);
	}
	if(failed && tid == 0 && status) atomicOr(status, 2);
	cluster.sync();                                     // nobody leaves while a peer may still write into its buffers
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

static int gmwb_cluster() {                // CTAs in the cluster: depth of the line pipeline
	static int g = 0;
	if(!g) { const char *e = getenv("QPDE_GMWB_CLUSTER"); g = e ? atoi(e) : 8; if(g < 4 || g > 16) g = 8; }
	return g;
}

template <int PT>
static cudaError_t launch_lines_t(GmwbShard &sh, GmwbRows R, double *x, const double *v0, const double *xguess, int ia0, int ia1, int max_local, int *status, cudaStream_t stream) {
	const size_t smem = sizeof(double) * 3 * (PT * 8 + 2) + (sizeof(double) + sizeof(int)) * gmwb_near_k(PT) * PT;
	cudaError_t e = cudaSuccess;
	const int G = gmwb_cluster();
	// The function attributes are per device and are set once per device: cudaFuncSetAttribute is not
	// stream-ordered — called per launch it held the host back until the line kernels in flight had
	// finished, so a pass meant to trail another one was issued only when that one was over (seen with
	// QPDE_GMWB_TIMELINE).
	{
		static std::mutex lock;
		static bool configured[64] = {};
		int device = 0; cudaGetDevice(&device);
		std::lock_guard<std::mutex> guard(lock);
		if(device < 0 || device >= 64 || !configured[device]) {
			e = cudaFuncSetAttribute(k_gmwb_lines<8, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if(e != cudaSuccess) return e;
			if(G > 8) { e = cudaFuncSetAttribute(k_gmwb_lines<8, PT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); if(e != cudaSuccess) return e; }
			if(device >= 0 && device < 64) configured[device] = true;
		}
	}
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(G); cfg.blockDim = dim3(PT); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, k_gmwb_lines<8, PT>, sh.d, R, x, v0, xguess, ia0, ia1, max_local, status);
}

static cudaError_t launch_lines(GmwbShard &sh, const GmwbRows &R, double *x, const double *v0, const double *xguess,
		int ia0, int ia1, int *status, cudaStream_t stream) {
	++*sh.launches;
	switch(sh.PT) {
	case 32:  return launch_lines_t<32>(sh, R, x, v0, xguess, ia0, ia1, sh.max_local, status, stream);
	case 64:  return launch_lines_t<64>(sh, R, x, v0, xguess, ia0, ia1, sh.max_local, status, stream);
	case 128: return launch_lines_t<128>(sh, R, x, v0, xguess, ia0, ia1, sh.max_local, status, stream);
	case 256: return launch_lines_t<256>(sh, R, x, v0, xguess, ia0, ia1, sh.max_local, status, stream);
	default:  return launch_lines_t<512>(sh, R, x, v0, xguess, ia0, ia1, sh.max_local, status, stream);
	}
}

// lut[b] = largest i with x[i] <= b h, h = x[n - 1] / GMWB_LUT; returns 1 / h.  A value ci in
// bucket b = floor(ci / h) has its bracket at or above lut[b] whatever the rounding of ci / h does
// to the bucket (a bucket too low only lengthens the walk; too high cannot happen by more than
// one, which is guarded by taking the entry of the bucket below).
static double make_lut(const double *x, int n, std::vector<int> &lut) {
	const double h = x[n - 1] / GMWB_LUT, invh = 1. / h;
	lut.assign(GMWB_LUT, 0);
	int i = 0;
	for(int bkt = 0; bkt < GMWB_LUT; ++bkt) {
		const double edge = (bkt > 0 ? bkt - 1 : 0) * h;      // one bucket of slack against rounding
		while(i + 1 < n - 1 && x[i + 1] <= edge) ++i;
		lut[bkt] = i;
	}
	return invh;
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
	R.raa = (double *)dalloc(sizeof(double) * n); R.rhs = (double *)dalloc(sizeof(double) * n); R.rhs2 = (double *)dalloc(sizeof(double) * n);
	R.tgt = (int *)dalloc(sizeof(int) * 4 * (size_t)n); R.tw = (double *)dalloc(sizeof(double) * 4 * (size_t)n);
	sh->flags = (int *)dalloc(sizeof(int) * 4);
	int *dlutW = (int *)dalloc(sizeof(int) * GMWB_LUT), *dlutA = (int *)dalloc(sizeof(int) * GMWB_LUT);
	std::vector<int> lutW, lutA;
	const double invhW = make_lut(W, nw, lutW), invhA = make_lut(A, na, lutA);
	if(!ok) { gmwb_shard_destroy(sh); *error = "gmwb: out of device memory"; return QPDE_ERR_CUDA; }
	cudaMemcpyAsync(dW, W, sizeof(double) * nw, cudaMemcpyHostToDevice, stream);
	cudaMemcpyAsync(dA, A, sizeof(double) * na, cudaMemcpyHostToDevice, stream);
	cudaMemcpyAsync(dq, p->controls, sizeof(double) * p->n_controls, cudaMemcpyHostToDevice, stream);
	cudaMemcpyAsync(dz, z, sizeof(double) * nz, cudaMemcpyHostToDevice, stream);
	cudaMemcpyAsync(dlutW, lutW.data(), sizeof(int) * GMWB_LUT, cudaMemcpyHostToDevice, stream);
	cudaMemcpyAsync(dlutA, lutA.data(), sizeof(int) * GMWB_LUT, cudaMemcpyHostToDevice, stream);
	cudaMemsetAsync(sh->flags, 0, sizeof(int) * 4, stream);
	cudaStreamSynchronize(stream);      // the host arrays may go away
	GmwbDev &d = sh->d;
	d.nw = nw; d.na = na; d.nq = p->n_controls; d.nz = nz;
	d.W = dW; d.A = dA; d.q = dq; d.z = dz;
	d.lutW = dlutW; d.lutA = dlutA; d.invhW = invhW; d.invhA = invhA;
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
static void launch_policy(GmwbShard *sh, const GmwbRows &R, const double *x, double h0, int a0, int a1, cudaStream_t stream,
		int mode = POLICY_FUSED, int *decision = nullptr) {
	const int r0 = a0 * sh->d.nw, r1 = a1 * sh->d.nw;
	if(r1 > r0) {
		const size_t smem = (size_t)(128 / sh->d.nw + 2) * sh->d.nz * (2 * sizeof(double) + sizeof(int));
		const int grid = (r1 - r0 + 127) / 128;
		if(mode == POLICY_SEARCH) k_gmwb_policy<POLICY_SEARCH><<<grid, 128, smem, stream>>>(sh->d, R, x, h0, r0, r1, decision);
		else if(mode == POLICY_ASSEMBLE) k_gmwb_policy<POLICY_ASSEMBLE><<<grid, 128, smem, stream>>>(sh->d, R, x, h0, r0, r1, decision);
		else k_gmwb_policy<POLICY_FUSED><<<grid, 128, smem, stream>>>(sh->d, R, x, h0, r0, r1, decision);
		++*sh->launches;
	}
}

void gmwb_shard_policy(GmwbShard *sh, const double *x, double h0) {
	launch_policy(sh, sh->R, x, h0, sh->a_begin, sh->a_end, sh->stream);
}

// the shard's line solves in ascending A (exact given final lines below a_begin);
// *status_dev |= 2 if a line's in-place iteration did not settle
cudaError_t gmwb_shard_sweep(GmwbShard *sh, double *x, const double *v0, int *status_dev) {
	if(sh->a_end > sh->a_begin) return launch_lines(*sh, sh->R, x, v0, x, sh->a_begin, sh->a_end, status_dev, sh->stream);
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

// Host driver of the single-GPU solve: the reference's time loop and ToleranceIteration around the
// kernels, on the shard [0, n_a).
//
// One policy iteration is a control search over every node (k_gmwb_policy: all SMs, FP64-bound) and
// a line pass (k_gmwb_lines: one cluster of 8 CTAs, a serial chain of n_a line solves).  Run one
// after the other they leave 140 SMs idle for half of the time and 8 busy for the other half, so
// the driver overlaps them along the A-axis, which is cut into a few chunks of lines:
//   * the control search of iteration k+1 on chunk j needs iterate k on the lines of chunks <= j
//     only (impulses and the upwinded drift point to equal-or-lower A), so it starts as soon as the
//     line pass of iteration k has finished chunk j — behind the chain, on the idle SMs;
//   * the line pass of iteration k+1 on chunk j needs that search and its own chunks < j, so it
//     trails the pass of iteration k by about one chunk on a second cluster.  What it cannot know
//     yet is whether iteration k converged: if it did, iteration k+1 belongs to the next time step
//     and runs with V^n = iterate k, otherwise with the old V^n.  The rows do not depend on V^n (it
//     enters in the line pass), so only the trailing pass is speculative: it is launched on the
//     likelier outcome ("this step takes as many iterations as the last") into a buffer of its own and simply
//     discarded — the right one starts at once from the same rows — when the verdict differs.
// Every iterate lives in a buffer of its own (a ring), so the passes never overwrite what another
// one reads; results are bit-identical to running the iterations one after the other
// (QPDE_GMWB_DEPTH=0).  Everything is ordered by events; the host waits for one 8-byte flag copy
// per iteration.
namespace {

// Spatial partition of the SMs (CUDA green contexts, driver API through cudaGetDriverEntryPoint: no link-time
// libcuda).  A line kernel is a cluster of G CTAs that each need a whole SM (512 threads x 128 registers);
// next to a control search whose small CTAs refill every SM the moment one of them retires, a second cluster
// never finds G empty SMs at once and the "concurrent" trailing pass runs only in the gaps (measured with
// QPDE_GMWB_TIMELINE: it started when the leading pass ended).  So the two line streams get G SMs each for
// themselves and the search, the convergence test and the exchange get the rest.
struct SmPartition {
	bool on = false;
	CUgreenCtx ctx[3] = { nullptr, nullptr, nullptr };      // line stream 0, line stream 1, the rest
	CUresult (*destroy)(CUgreenCtx) = nullptr;
	int line_sms = 0, rest_sms = 0;
};

template <typename F>
bool driver_entry(const char *name, F *fn) {
	void *ptr = nullptr;
	cudaDriverEntryPointQueryResult q;
	if(cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !ptr) return false;
	*fn = (F)ptr;
	return true;
}

// streams[0..1]: the line streams, streams[2..3]: search and convergence test.  false: the caller uses ordinary streams.
bool sm_partition_create(int device, int cluster, int prio_hi, int prio_lo, SmPartition *part, cudaStream_t streams[4]) {
	CUresult (*getRes)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
	CUresult (*split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
	CUresult (*genDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
	CUresult (*create)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
	CUresult (*streamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
	if(!driver_entry("cuDeviceGetDevResource", &getRes) || !driver_entry("cuDevSmResourceSplitByCount", &split)
			|| !driver_entry("cuDevResourceGenerateDesc", &genDesc) || !driver_entry("cuGreenCtxCreate", &create)
			|| !driver_entry("cuGreenCtxStreamCreate", &streamCreate) || !driver_entry("cuGreenCtxDestroy", &part->destroy)) return false;
	CUdevResource all, groups[2], rest;
	if(getRes((CUdevice)device, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
	unsigned int nb = 2;
	const unsigned int want = (unsigned int)((cluster + 7) / 8 * 8);          // granularity on sm_90+: multiples of 8
	if(split(groups, &nb, &all, &rest, 0, want) != CUDA_SUCCESS || nb != 2 || rest.sm.smCount < 64) return false;
	CUdevResource *res[3] = { &groups[0], &groups[1], &rest };
	for(int k = 0; k < 3; ++k) {
		CUdevResourceDesc desc;
		if(genDesc(&desc, res[k], 1) != CUDA_SUCCESS || create(&part->ctx[k], desc, (CUdevice)device, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
			for(int q = 0; q < k; ++q) part->destroy(part->ctx[q]);
			return false;
		}
	}
	bool ok = streamCreate(&streams[0], part->ctx[0], CU_STREAM_NON_BLOCKING, prio_hi) == CUDA_SUCCESS
		&& streamCreate(&streams[1], part->ctx[1], CU_STREAM_NON_BLOCKING, prio_hi) == CUDA_SUCCESS
		&& streamCreate(&streams[2], part->ctx[2], CU_STREAM_NON_BLOCKING, prio_lo) == CUDA_SUCCESS
		&& streamCreate(&streams[3], part->ctx[2], CU_STREAM_NON_BLOCKING, prio_hi) == CUDA_SUCCESS;
	if(!ok) { for(int q = 0; q < 3; ++q) part->destroy(part->ctx[q]); return false; }
	part->on = true; part->line_sms = (int)groups[0].sm.smCount; part->rest_sms = (int)rest.sm.smCount;
	return true;
}

struct GmwbPass {
	int cur = -1, prev = -1, v0 = -1, rset = 0, slot = 0;
	double h0 = 0.;
	std::vector<cudaEvent_t> lines_done;
	cudaEvent_t flag_ready = nullptr;
};

struct GmwbSearch {         // the control search that last filled a row set
	int prev = -1; double h0 = 0.; bool valid = false;
	int n_ch = 1; long gen = -1;          // generation of the search: its events are ring[(gen % slots) n_ch + chunk]
	std::vector<cudaEvent_t> ring;
	std::vector<cudaEvent_t> readers;     // line passes that read the row set since
	cudaEvent_t done(int chunk) const { const long slots = (long)ring.size() / n_ch; return ring[(size_t)((gen % slots) * n_ch + chunk)]; }
};

} // namespace

int gmwb_run(const qpde_gmwb2d *p, const double *W, int nw, const double *A, int na, const double *z, int nz,
		double *V_host, qpde_gmwb2d_stats *stats, cudaStream_t stream, long long *launches, std::string *error,
		Comm *comm) {
	const int n = nw * na;
	// Sharded by the A-axis: what shards is the control search — the part of an iteration
	// that is throughput-bound (every node x every impulse candidate).  Rank r searches its block of the
	// lines of every chunk, the decisions of the blocks (4 bytes per node) are exchanged with NCCL broadcasts
	// (the "all-gather before the impulse step") and every rank assembles all rows from them; the line pass — a latency-bound chain on one cluster of 8 SMs —
	// is run by every rank on the full grid, which costs nothing and saves the exchange of the iterate and
	// of the stopping flag: all ranks hold the same rows, hence bit-identical iterates and verdicts, and
	// take the same (deterministic) decisions, speculative ones included.
	const int world = comm_world(comm), rank = comm_rank(comm);
	bool comm_ok = true;
	// QPDE_GMWB_SPLIT=1: the sharded code path (search -> decisions -> assembly) on one rank, for testing
	const bool split = world > 1 || (getenv("QPDE_GMWB_SPLIT") && atoi(getenv("QPDE_GMWB_SPLIT")) != 0);
	GmwbShard *sh = nullptr;
	int rc = gmwb_shard_create(p, W, nw, A, na, z, nz, 0, na, stream, launches, &sh, error);
	if(rc != QPDE_OK) return rc;
	bool ok = true;
	auto dalloc = [&] (size_t bytes) -> void * {
		void *ptr = nullptr;
		if(cudaMalloc(&ptr, bytes) != cudaSuccess) { ok = false; return nullptr; }
		sh->allocs.push_back(ptr);
		return ptr;
	};
	const char *env_depth = getenv("QPDE_GMWB_DEPTH"), *env_chunks = getenv("QPDE_GMWB_CHUNKS");
	// Measured on B200 at 4097 x 1025 nodes (64 steps): 3.49 s one after the other, 2.27 s with the
	// control search overlapped (depth 1).  The speculative trailing pass (depth 2) pays only where the
	// control search is much shorter than the line pass — here they are level, and a second pass in
	// flight has nothing left to hide behind — so it is opt-in.
	// With the search cut into `world` pieces the line chain dominates and the trailing pass pays: depth 2.
	int depth = env_depth ? atoi(env_depth) : (world > 1 ? 2 : 1);
	if(depth < 0 || depth > 2) depth = 1;
	// chunks of at least 128 lines: a chunk is one cluster launch and one pipeline fill of the chain
	int n_ch = env_chunks ? atoi(env_chunks) : 8;
	if(n_ch > na / 128) n_ch = na / 128;
	if(n_ch < 1 || depth == 0) n_ch = 1;
	std::vector<int> a0((size_t)n_ch + 1);
	for(int j = 0; j <= n_ch; ++j) a0[j] = (int)((long long)na * j / n_ch);

	// Events are re-recorded when their pass slot / search generation comes round again, and a wait binds to
	// the latest record: a reference kept in buf_events[] must not outlive its slot.  A buffer is reused
	// after at most NB passes, so rings of 3 NB pass slots and search generations keep every kept
	// reference on the record it was made for (with 6 slots a recycled buffer made each new pass wait for
	// the END of the pass before it through a stale flag_ready: no two passes ever overlapped).
	constexpr int NB = 8, NPASS = 3 * NB, NSEARCH = 3 * NB;
	double *buf[NB];
	for(int k = 0; k < NB; ++k) buf[k] = (double *)dalloc(sizeof(double) * n);
	GmwbRows R[2];
	R[0] = sh->R;
	R[1].ra = (double *)dalloc(sizeof(double) * n); R[1].rb = (double *)dalloc(sizeof(double) * n); R[1].rc = (double *)dalloc(sizeof(double) * n);
	R[1].raa = (double *)dalloc(sizeof(double) * n); R[1].rhs = (double *)dalloc(sizeof(double) * n); R[1].rhs2 = (double *)dalloc(sizeof(double) * n);
	R[1].tgt = (int *)dalloc(sizeof(int) * 4 * (size_t)n); R[1].tw = (double *)dalloc(sizeof(double) * 4 * (size_t)n);
	int *flags_dev = (int *)dalloc(sizeof(int) * 2 * NPASS);
	int *decision[2] = { nullptr, nullptr };      // sharded: what the control search of a row set decided, per node
	if(split) for(int k = 0; k < 2; ++k) decision[k] = (int *)dalloc(sizeof(int) * n);
	int *hflags = nullptr;
	if(!ok || cudaHostAlloc((void **)&hflags, sizeof(int) * 2 * NPASS, cudaHostAllocDefault) != cudaSuccess) {
		gmwb_shard_destroy(sh); *error = "gmwb: out of memory"; return QPDE_ERR_CUDA;
	}
	cudaStream_t sL[2], sP, sR, sC = nullptr;
	// the line passes are the critical path: their clusters go first when SMs free up
	int prio_lo = 0, prio_hi = 0;
	cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
	// with two passes in flight the line streams own their SMs (SmPartition above); QPDE_GMWB_PARTITION=0 / 1 overrides
	SmPartition part;
	{
		const char *env_part = getenv("QPDE_GMWB_PARTITION");
		const bool want = env_part ? atoi(env_part) != 0 : depth >= 2;
		int device = 0; cudaGetDevice(&device);
		cudaStream_t st4[4];
		if(want && sm_partition_create(device, gmwb_cluster(), prio_hi, prio_lo, &part, st4)) {
			sL[0] = st4[0]; sL[1] = st4[1]; sP = st4[2]; sR = st4[3];
		} else {
			cudaStreamCreateWithPriority(&sL[0], cudaStreamNonBlocking, prio_hi); cudaStreamCreateWithPriority(&sL[1], cudaStreamNonBlocking, prio_hi);
			cudaStreamCreateWithPriority(&sP, cudaStreamNonBlocking, prio_lo); cudaStreamCreateWithPriority(&sR, cudaStreamNonBlocking, prio_hi);
		}
	}
	// the exchange of the rows runs behind the search of the next chunk, on a stream of its own
	if(split) cudaStreamCreateWithPriority(&sC, cudaStreamNonBlocking, prio_hi);
	cudaEvent_t searched = nullptr;
	std::vector<cudaEvent_t> all_events;
	// QPDE_GMWB_TIMELINE=n: events carry times, and the start / end of every launch of the last n passes is printed
	const int timeline = getenv("QPDE_GMWB_TIMELINE") ? atoi(getenv("QPDE_GMWB_TIMELINE")) : 0;
	auto new_event = [&] () { cudaEvent_t e; cudaEventCreateWithFlags(&e, timeline > 0 ? cudaEventDefault : cudaEventDisableTiming); all_events.push_back(e); return e; };
	struct Mark { char what; long pass; int chunk; cudaEvent_t begin, end; double host_ms; };
	std::vector<Mark> marks;
	long pass_counter = 0;
	auto mark_begin = [&] (char what, long pass, int chunk, cudaStream_t s) -> size_t {
		if(timeline <= 0) return 0;
		const double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
		Mark m{what, pass, chunk, new_event(), new_event(), now};
		cudaEventRecord(m.begin, s);
		marks.push_back(m);
		return marks.size() - 1;
	};
	auto mark_end = [&] (size_t at, cudaStream_t s) { if(timeline > 0) cudaEventRecord(marks[at].end, s); };
	GmwbPass passes[NPASS];
	for(int k = 0; k < NPASS; ++k) {
		passes[k].slot = k;
		for(int j = 0; j < n_ch; ++j) passes[k].lines_done.push_back(new_event());
		passes[k].flag_ready = new_event();
	}
	GmwbSearch search[2];
	for(int k = 0; k < 2; ++k) { search[k].n_ch = n_ch; for(int j = 0; j < n_ch * NSEARCH; ++j) search[k].ring.push_back(new_event()); }
	std::vector<cudaEvent_t> buf_events[NB];       // what must have finished before buffer b is overwritten
	cudaEvent_t start_ready = new_event();

	auto cleanup = [&] () {
		cudaStreamSynchronize(sL[0]); cudaStreamSynchronize(sL[1]); cudaStreamSynchronize(sP); cudaStreamSynchronize(sR);
		cudaStreamSynchronize(stream);
		for(cudaEvent_t e : all_events) cudaEventDestroy(e);
		if(sC) { cudaStreamSynchronize(sC); cudaStreamDestroy(sC); }
		cudaStreamDestroy(sL[0]); cudaStreamDestroy(sL[1]); cudaStreamDestroy(sP); cudaStreamDestroy(sR);
		if(part.on) for(int q = 0; q < 3; ++q) part.destroy(part.ctx[q]);
		cudaFreeHost(hflags);
		gmwb_shard_destroy(sh);
	};

	// the control search of one iteration: reads buf[prev] chunk by chunk behind its producer
	auto enqueue_search = [&] (int rset, int prev, double h0, const GmwbPass *producer) {
		GmwbSearch &S = search[rset];
		++S.gen;                               // a fresh set of events for this search
		for(int j = 0; j < n_ch; ++j) {
			if(producer) cudaStreamWaitEvent(sP, producer->lines_done[j], 0);
			else if(j == 0) cudaStreamWaitEvent(sP, start_ready, 0);
			if(j == 0) for(cudaEvent_t e : S.readers) cudaStreamWaitEvent(sP, e, 0);      // the passes that still read these rows
			const size_t mk = mark_begin('S', pass_counter, j, sP);
			if(!split) {
				launch_policy(sh, R[rset], buf[prev], h0, a0[j], a0[j + 1], sP);
			} else {
				// this rank searches its block of the chunk; the decisions (one word per node) of every block reach
				// every rank, and every rank assembles the rows of the whole chunk from them
				auto cut = [&] (int r) { return a0[j] + (int)((long long)(a0[j + 1] - a0[j]) * r / world); };
				launch_policy(sh, R[rset], buf[prev], h0, cut(rank), cut(rank + 1), sP, POLICY_SEARCH, decision[rset]);
				if(!searched) searched = new_event();
				// one event serves every chunk: a wait binds to the record that precedes it
				cudaEventRecord(searched, sP);
				cudaStreamWaitEvent(sC, searched, 0);
				if(j == 0) for(cudaEvent_t e : S.readers) cudaStreamWaitEvent(sC, e, 0);
				if(world > 1) {
					comm_ok = comm_ok && comm_group_start(comm);
					for(int r = 0; r < world; ++r) {
						const size_t r0 = (size_t)cut(r) * nw, cnt = (size_t)(cut(r + 1) - cut(r)) * nw;
						if(cnt) comm_ok = comm_ok && comm_broadcast(comm, decision[rset] + r0, cnt * sizeof(int), r, sC);
					}
					comm_ok = comm_ok && comm_group_end(comm);
				}
				launch_policy(sh, R[rset], buf[prev], h0, a0[j], a0[j + 1], sC, POLICY_ASSEMBLE, decision[rset]);
			}
			mark_end(mk, split ? sC : sP);
			cudaEventRecord(S.done(j), split ? sC : sP);
		}
		S.readers.clear();
		S.prev = prev; S.h0 = h0; S.valid = true;
		buf_events[prev].push_back(S.done(n_ch - 1));
	};
	int next_pass = 0, next_buf = 1;
	auto enqueue_pass = [&] (int prev, int v0, double h0, int rset, int line_stream, const int *keep, int n_keep) -> GmwbPass * {
		GmwbPass &P = passes[next_pass];
		next_pass = (next_pass + 1) % NPASS;
		// a buffer nobody of interest holds
		for(;; next_buf = (next_buf + 1) % NB) {
			bool held = next_buf == prev || next_buf == v0;
			for(int q = 0; q < n_keep; ++q) held = held || next_buf == keep[q];
			if(!held) break;
		}
		P.cur = next_buf; next_buf = (next_buf + 1) % NB;
		P.prev = prev; P.v0 = v0; P.h0 = h0; P.rset = rset;
		cudaStream_t sl = sL[line_stream];
		for(cudaEvent_t e : buf_events[P.cur]) cudaStreamWaitEvent(sl, e, 0);
		buf_events[P.cur].clear();
		cudaMemsetAsync(flags_dev + 2 * P.slot, 0, sizeof(int) * 2, sl);
		for(int j = 0; j < n_ch; ++j) {
			cudaStreamWaitEvent(sl, search[rset].done(j), 0);
			const size_t mk = mark_begin('L', pass_counter, j, sl);
			launch_lines(*sh, R[rset], buf[P.cur], buf[v0], buf[prev], a0[j], a0[j + 1], flags_dev + 2 * P.slot + 1, sl);
			mark_end(mk, sl);
			cudaEventRecord(P.lines_done[j], sl);
			cudaStreamWaitEvent(sR, P.lines_done[j], 0);
			const int r0 = a0[j] * nw, r1 = a0[j + 1] * nw;
			k_gmwb_relerr<<<(r1 - r0 + 255) / 256, 256, 0, sR>>>(r0, r1, buf[P.cur], buf[prev], 1., sh->d.tol, flags_dev + 2 * P.slot);
			++*launches;
		}
		cudaMemcpyAsync(hflags + 2 * P.slot, flags_dev + 2 * P.slot, sizeof(int) * 2, cudaMemcpyDeviceToHost, sR);
		cudaEventRecord(P.flag_ready, sR);
		++pass_counter;
		search[rset].readers.push_back(P.lines_done[n_ch - 1]);
		buf_events[P.cur].push_back(P.flag_ready);
		buf_events[prev].push_back(P.flag_ready);
		buf_events[v0].push_back(P.lines_done[n_ch - 1]);
		return &P;
	};

	gmwb_shard_exit(sh, buf[0]);
	cudaEventRecord(start_ready, stream);

	Replay rp; rp.start(p->T, p->T / p->timesteps, 0);
	double time1 = p->T;
	long inner_sum = 0; int n_steps = 0, status = 0;
	int accepted = 0;                     // buffer of the latest accepted iterate
	const GmwbPass *accepted_pass = nullptr;
	GmwbPass *spec = nullptr;             // the trailing pass launched on a prediction
	long spec_used = 0, spec_wasted = 0;  // trailing passes that were kept / discarded (QPDE_GMWB_TRACE)
	long k = 0;                           // iteration counter: row set k % 2, line stream k % 2
	int last_its = 2;                     // iterations the previous step took: the prediction for this one
	cudaError_t err = cudaSuccess;
	bool more = true;
	while(more && err == cudaSuccess) {
		qpde_step st;
		more = rp.next(st);
		const double h0 = time1 - st.t;
		// the step after this one, for the look-ahead
		double h0_next = 0.; bool has_next = more;
		if(more) { Replay peek = rp; qpde_step st2; peek.next(st2); h0_next = st.t - st2.t; }
		const int v0 = accepted;
		int its = 0, notdone = 0;
		do {
			GmwbPass *P = nullptr;
			if(spec && spec->prev == accepted && spec->v0 == v0 && spec->h0 == h0) P = spec;
			if(spec) { if(P) ++spec_used; else ++spec_wasted; }
			spec = nullptr;
			const int rset = (int)(k % 2);
			if(!P) {
				if(!(search[rset].valid && search[rset].prev == accepted && search[rset].h0 == h0)) enqueue_search(rset, accepted, h0, accepted_pass);
				P = enqueue_pass(accepted, v0, h0, rset, rset, nullptr, 0);
			}
			// ---- look ahead to iteration k + 1
			const bool predict_converged = its + 1 >= last_its;
			if(depth >= 1 && (!predict_converged || has_next)) {
				const double h0_la = predict_converged ? h0_next : h0;
				const int rs1 = (int)((k + 1) % 2);
				enqueue_search(rs1, P->cur, h0_la, P);
				if(depth >= 2 && its + 1 < p->max_inner) {
					const int keep[3] = { P->cur, v0, accepted };
					spec = enqueue_pass(P->cur, predict_converged ? P->cur : v0, h0_la, rs1, rs1, keep, 3);
				}
			}
			// ---- the verdict of iteration k
			err = cudaEventSynchronize(P->flag_ready);
			if(err != cudaSuccess) break;
			notdone = hflags[2 * P->slot];
			if(hflags[2 * P->slot + 1]) status = 2;
			accepted = P->cur; accepted_pass = P;
			++its; ++k;
			if(notdone && its >= p->max_inner) { if(!status) status = 1; break; }
		} while(notdone);
		if(stats && stats->inner && n_steps < stats->inner_capacity) stats->inner[n_steps] = its;
		inner_sum += its;
		last_its = its;
		++n_steps;
		time1 = st.t;
	}
	if(err == cudaSuccess) err = cudaGetLastError();
	if(err == cudaSuccess) {
		// the accepted iterate is complete (its flag was read); copy it out on the handle's stream
		err = cudaMemcpyAsync(V_host, buf[accepted], sizeof(double) * n, cudaMemcpyDeviceToHost, stream);
		if(err == cudaSuccess) err = cudaStreamSynchronize(stream);
	}
	if(timeline > 0 && !marks.empty() && rank == 0) {
		cudaDeviceSynchronize();
		const long first = pass_counter - timeline;
		cudaEvent_t base = nullptr; double base_host = 0.;
		for(const Mark &m : marks) if(m.pass >= first && (m.what == 'L' || m.what == 'S')) { base = m.begin; base_host = m.host_ms; break; }
		for(const Mark &m : marks) {
			if(m.pass < first || !base) continue;
			float b = 0.f, e = 0.f;
			cudaEventElapsedTime(&b, base, m.begin); cudaEventElapsedTime(&e, base, m.end);
			fprintf(stderr, "timeline %c pass %ld chunk %d: %8.3f .. %8.3f ms (issued by the host at %8.3f)\n", m.what, m.pass, m.chunk, b, e, m.host_ms - base_host);
		}
	}
	cleanup();
	if(!comm_ok) { *error = "gmwb solve: an NCCL call failed"; return QPDE_ERR_CUDA; }
	if(getenv("QPDE_GMWB_TRACE")) {
		fprintf(stderr, "gmwb: rank %d of %d, SM partition %d + %d + %d, depth %d, %d chunks, %ld iterations in %d steps, trailing passes kept %ld, discarded %ld\n",
			rank, world, part.on ? part.line_sms : 0, part.on ? part.line_sms : 0, part.on ? part.rest_sms : 0, depth, n_ch, k, n_steps, spec_used, spec_wasted);
	}
	if(err != cudaSuccess) { *error = std::string("gmwb solve: ") + cudaGetErrorString(err); return QPDE_ERR_CUDA; }
	if(stats) { stats->n_steps = n_steps; stats->mean_inner = (double)inner_sum / n_steps; stats->status = status; stats->nodes_w = nw; stats->nodes_a = na; }
	return QPDE_OK;
}

} // namespace qpde
