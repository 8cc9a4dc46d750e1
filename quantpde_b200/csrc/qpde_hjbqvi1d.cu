// qpde_hjbqvi1d.cu — 1-D impulse-control HJBQVI, penalised scheme: one CTA per problem, a batch of
// problems per launch, the whole backward sweep (time loop, policy iteration, control searches,
// linear solves) in one kernel with the state on chip.
//
// Reference call stack replaced, per problem (examples/hjbqvi/exchange_rate.cpp:29-175):
//   HJBQVI<1,1,1>::solve(refinement)          Modules/HJBQVI/HJBQVI.hpp:590-811
//     ReverseConstantStepper / TimeIteration  Core/IterativeMethod.hpp:1259-1317
//     ToleranceIteration + relativeError      :1125-1137,1211-1254
//     MinPolicyIteration (stochastic control) Core/PolicyIteration.hpp:54-105 over
//       ControlledOperator::A / b             HJBQVI.hpp:158-335 (no boundary routines: rows 0, n-1 = [rho])
//     MinPolicyIteration (impulse control)    over Impulse::A / b, Core/Impulse.hpp:87-216
//     MinPenaltyMethod (non-direct)           Core/PenaltyMethod.hpp:37-119,127-139
//     ReverseBDFOne                           Core/BDF.hpp:32-46
//     BiCGSTABSolver                          Bindings/Eigen.hpp:170-204
//
// The matrix of an inner iteration is tridiagonal plus, in every penalised row, the two
// interpolation entries of the impulse target.  Entries that fall on the three diagonals are folded
// into them; the others are lagged and the tridiagonal solve (PartitionSolver, cached
// factorisation) is repeated until the update is below 1e-15 relative — the reference drives
// BiCGSTAB + ILUT to a machine-epsilon residual, so both reach the solution of the same system.
//
// The coefficient functions are tagged forms (qpde_hjbqvi1d_batch::model); model 1 is the exchange-rate
// intervention of Mundaca & Oksendal with per-problem parameters (rho, sigma, a, b, m, lambda, c).
#include <math_constants.h>

#include "qpde_kernels.cuh"
#include "qpde_partition.cuh"

namespace qpde {

namespace {

constexpr int HP = 256;   // threads per problem; M rows per thread (the smallest of 2, 3, 5, 8, 9 with M * HP >= nodes: rows are
                          // contiguous per thread, so a larger M would leave most threads without rows), MP = M * HP rows of a system

// model 1: examples/hjbqvi/exchange_rate.cpp:118-152
struct ExchangeRate {
	double rho, sigma, a, b, m, lambda, c;
	__device__ __forceinline__ void load(const double *p) { rho = p[0]; sigma = p[1]; a = p[2]; b = p[3]; m = p[4]; lambda = p[5]; c = p[6]; }
	__device__ __forceinline__ double discount(double) const { return rho; }
	__device__ __forceinline__ double volatility(double) const { return sigma; }
	__device__ __forceinline__ double drift(double, double q) const { return -a * q; }
	__device__ __forceinline__ double flow(double x, double q) const { const double tmp = x - m; return - tmp * tmp - q * q * b; }
	// the impulse control is the new state itself: targets do not depend on the node
	__device__ __forceinline__ double transition(double, double z) const { return z; }
	__device__ __forceinline__ double impulse_flow(double x, double z) const {
		if(fabs(z - m) >= fabs(x - m)) return -CUDART_INF;      // controls that do not move towards the parity are pruned
		return - lambda * fabs(z - x) - c;
	}
	__device__ __forceinline__ double exit_value(double) const { return 0.; }
};

// node i of a node-addressable shared-memory array: the rows of one thread are HP doubles apart
template <int M>
__device__ __forceinline__ int slot(int i) { return (i % M) * HP + (i / M); }

struct Hjbqvi1dSmem {
	PartitionSmem ps;
	int n_steps, status; long long inner_sum, sweep_sum;
};

// shared memory: 7 row arrays + the control grids and the per-candidate tables (doubles), then two int arrays
// 8 row arrays of doubles (grid, iterate, flow, reward and weight of the chosen controls, the hand-over of the
// assembled rows), the control grids and per-candidate tables, two int arrays, two u16 arrays (the chosen
// controls as grid indices), one byte array (flags)
__host__ __device__ inline size_t hjbqvi1d_smem_bytes(int M, int nq, int nz) {
	return ((sizeof(Hjbqvi1dSmem) + 15) / 16) * 16 + sizeof(double) * ((size_t)8 * M * HP + nq + 5 * (size_t)nz)
		+ sizeof(int) * ((size_t)M * HP + nz) + sizeof(unsigned short) * 2 * (size_t)M * HP + (size_t)M * HP;
}

// linearInterpolationData (Core/Interpolant.hpp:153-181) over a node-addressable grid
template <int M>
__device__ __forceinline__ void interp_nodes(const double *sX, int n, double ci, int &idx, double &w) {
	if(ci <= sX[slot<M>(0)]) { idx = 0; w = 1.; return; }
	if(ci >= sX[slot<M>(n - 1)]) { idx = n - 2; w = 0.; return; }
	int lo = 0, hi = n - 2, mid = 0;
	w = 0.;
	while(lo <= hi) {
		mid = (lo + hi) / 2;
		if(ci < sX[slot<M>(mid)]) hi = mid - 1;
		else if(ci >= sX[slot<M>(mid + 1)]) lo = mid + 1;
		else { w = (sX[slot<M>(mid + 1)] - ci) / (sX[slot<M>(mid + 1)] - sX[slot<M>(mid)]); break; }
	}
	idx = mid;
}

// ((I - M) x)_row for the impulse row M = { (idx, w), (idx + 1, -w + 1) }, accumulated in column order
// (the product of a row-major sparse matrix with a vector).  A = (0 - w) x[idx], B = (0 - (1 - w)) x[idx + 1].
__device__ __forceinline__ double impulse_lhs(int row, int idx, double w, double xr, double A, double B) {
	if(row < idx) return ((0. + 1. * xr) + A) + B;
	if(row == idx) return (0. + (1. - w) * xr) + B;
	if(row == idx + 1) return (0. + A) + (1. - (-w + 1.)) * xr;
	return ((0. + A) + B) + 1. * xr;
}

} // namespace

template <typename Model, int M>
__global__ void __launch_bounds__(HP, M <= 3 ? 2 : 1)
k_hjbqvi1d(Hjbqvi1dBatch b, Hjbqvi1dResult out) {
	constexpr int HM = M, HMP = M * HP;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Hjbqvi1dSmem &sm = *reinterpret_cast<Hjbqvi1dSmem *>(smem_raw);
	double *const sX = reinterpret_cast<double *>(smem_raw + ((sizeof(Hjbqvi1dSmem) + 15) / 16) * 16);   // grid, node-addressable (slot<M>())
	double *const sx = sX + HMP;          // current iterate, node-addressable
	double *const r_b = sx + HMP;         // [j][tid]: the flow b^q* of the chosen stochastic control
	double *const i_rew = r_b + HMP;      // reward and interpolation weight of the chosen impulse control
	double *const i_w = i_rew + HMP;
	// The searches run node-strided (thread t takes the nodes t, t + 256, ...), so that every warp gets near and
	// far nodes alike — the admissible impulse candidates of a node grow with its distance from the parity, and with
	// the solver's contiguous rows per thread the warps at the ends of the grid searched while the central ones sat
	// at the solver's barrier (ncu: 32 % of the stall samples).  The rows then reach their owners through shared memory.
	double *const h_a = i_w + HMP;        // hand-over of the assembled rows
	double *const h_b = h_a + HMP;
	double *const h_c = h_b + HMP;
	double *const sq = h_c + HMP;         // stochastic control grid [nq]
	double *const sz = sq + b.nq;         // impulse control grid [nz], its distance to the parity, weight of the target,
	double *const sdz = sz + b.nz;        // A_k, B_k (impulse_lhs)
	double *const z_w = sdz + b.nz;
	double *const sA = z_w + b.nz;
	double *const sB = sA + b.nz;
	int *const i_k = reinterpret_cast<int *>(sB + b.nz);   // [j][tid]: lower interpolation node of the chosen impulse target
	int *const z_idx = i_k + HMP;         // [nz]: lower interpolation node of candidate k's target
	unsigned short *const q_k = reinterpret_cast<unsigned short *>(z_idx + b.nz);   // [j][tid]: chosen stochastic control (index into sq)
	unsigned short *const z_k = q_k + HMP;                                          // chosen impulse control: index into sz + 1, 0 = none
	unsigned char *const h_f = reinterpret_cast<unsigned char *>(z_k + HMP);        // bit 0 active, bits 1-2 far entries

	const int pidx = blockIdx.x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int N = b.N, nq = b.nq, nz = b.nz;
	const int i0 = tid * HM;
	Model mdl; mdl.load(b.params + (size_t)pidx * b.params_stride);
	const double *const xaxis = b.x_axis + (size_t)pidx * b.x_axis_stride;
	const double *const qaxis = b.q_axis + (size_t)pidx * b.q_axis_stride;
	const double *const zaxis = b.z_axis + (size_t)pidx * b.z_axis_stride;

	// ---- set-up: refined grids (RectilinearGrid::refined, Core/Domain.hpp:1092-1151), targets
	for(int i = tid; i < HMP; i += HP) sX[slot<M>(i)] = i < N ? refined_node(xaxis, b.refine, i) : 0.;
	for(int k = tid; k < nq; k += HP) sq[k] = refined_node(qaxis, b.refine_q, k);
	for(int k = tid; k < nz; k += HP) sz[k] = refined_node(zaxis, b.refine_z, k);
	if(tid == 0) { sm.n_steps = 0; sm.status = 0; sm.inner_sum = 0; sm.sweep_sum = 0; }
	__syncthreads();
	for(int k = tid; k < nz; k += HP) {
		int idx; double w;
		interp_nodes<M>(sX, N, mdl.transition(0., sz[k]), idx, w);      // node-independent transition (Model::transition ignores x)
		z_idx[k] = idx; z_w[k] = w;
		sdz[k] = fabs(sz[k] - mdl.m);
	}
	double x[HM], v0[HM], xprev[HM];
#pragma unroll
	for(int j = 0; j < HM; ++j) x[j] = i0 + j < N ? mdl.exit_value(sX[j * HP + tid]) : 0.;
	const double pen = 1. / (b.scaling_factor * (b.T / (double)b.timesteps));   // HJBQVI.hpp:634-635, PenaltyMethod.hpp:85
	PartitionSolver<HM, HP> ps; ps.init();
	unsigned active = 0;       // bit j: row j is penalised
	unsigned far = 0;          // bit 2j / 2j+1: entry idx / idx + 1 of row j is off the three diagonals
	__syncthreads();

	Replay rp; rp.start(b.T, b.T / (double)b.timesteps, 0, 0);
	double time1 = rp.initial();
	bool more = true;
	while(more) {
		qpde_step st;
		more = rp.next(st);
		const double t = st.t, h0 = rp.lapse(time1, t);
#pragma unroll
		for(int j = 0; j < HM; ++j) v0[j] = x[j];
		int its = 0;
		bool notdone;
		do {
			// ---- iterand(0) for the searches
#pragma unroll
			for(int j = 0; j < HM; ++j) { xprev[j] = x[j]; sx[j * HP + tid] = x[j]; }
			__syncthreads();
			for(int k = tid; k < nz; k += HP) {
				const int idx = z_idx[k]; const double w = z_w[k];
				sA[k] = (0. - w) * sx[slot<M>(idx)];
				sB[k] = (0. - (-w + 1.)) * sx[slot<M>(idx + 1)];
			}
			__syncthreads();
			active = 0; far = 0;
			double aa[HM], bb[HM], cc[HM];
#pragma unroll
			for(int j = 0; j < HM; ++j) {
				const int i = j * HP + tid;                 // the node this thread searches
				const int at = slot<M>(i);                  // its place in the [row][thread] arrays
				if(i >= N) { h_a[at] = 0.; h_b[at] = 1.; h_c[at] = 0.; h_f[at] = 0; continue; }
				const double xi = sX[at], vi = sx[at];
				const bool interior = i > 0 && i < N - 1;
				// 1. stochastic policy: argmin_q (A(q) x - b(q))_i, strict < (PolicyIteration.hpp:88-96)
				double dxb = 1., dxc = 1., dxf = 1., alpha_common = 0., beta_common = 0., vm = 0., vp = 0.;
				if(interior) {
					const double xm = sX[slot<M>(i - 1)], xp = sX[slot<M>(i + 1)];
					dxb = xi - xm; dxc = xp - xm; dxf = xp - xi;
					const double v = mdl.volatility(xi), vv = v * v;
					alpha_common = vv / dxb / dxc; beta_common = vv / dxf / dxc;
					vm = sx[slot<M>(i - 1)]; vp = sx[slot<M>(i + 1)];
				}
				const double rho = mdl.discount(xi);
				// The search divides the drift of every candidate by a spacing of the node: the quotient comes from a
				// reciprocal formed once per node plus one correction (div_by: the correctly rounded quotient but for
				// rare last-bit cases) instead of the ~25-instruction division sequence; the row of the CHOSEN control
				// is then formed again with the reference's plain divisions, so that the matrix is the reference's.
				const double rdxc = 1. / dxc, rdxf = 1. / dxf, rdxb = 1. / dxb;
				auto row_of = [&] (double q, bool exact, double &lo, double &up, double &di) {
					lo = 0.; up = 0.;
					double total = 0.;
					if(interior) {
						const double mu = 0. + mdl.drift(xi, q);
						const double mc = exact ? mu / dxc : div_by(mu, dxc, rdxc);
						double alpha = alpha_common - mc, beta = beta_common + mc;
						if(alpha < 0.) { alpha = alpha_common; beta = beta_common + (exact ? mu / dxf : div_by(mu, dxf, rdxf)); }
						else if(beta < 0.) { alpha = alpha_common - (exact ? mu / dxb : div_by(mu, dxb, rdxb)); beta = beta_common; }
						lo = -alpha; up = -beta; total += alpha + beta;
					}
					di = total + rho;
				};
				double best = CUDART_INF, bfl = 0., bq = 0.;
				int bqk = 0;
				for(int k = 0; k < nq; ++k) {
					const double q = sq[k];
					double lo, up, di;
					row_of(q, false, lo, up, di);
					const double fl = mdl.flow(xi, q);
					double acc = 0.;
					if(interior) acc += lo * vm;
					acc += di * vi;
					if(interior) acc += up * vp;
					const double cand = acc - fl;
					if(cand < best) { best = cand; bfl = fl; bq = q; bqk = k; }
				}
				double blo, bup, bdi;
				row_of(bq, true, blo, bup, bdi);
				r_b[at] = bfl; q_k[at] = (unsigned short)bqk;
				// 2. impulse policy: argmin_z ((I - M_z) x - reward_z)_i, strict <; pruned candidates are +inf
				best = CUDART_INF;
				int bk = -1;
				const double dxi = fabs(xi - mdl.m);
				for(int k = 0; k < nz; ++k) {
					if(sdz[k] >= dxi) continue;
					const double rew = - mdl.lambda * fabs(sz[k] - xi) - mdl.c;
					const double cand = impulse_lhs(i, z_idx[k], z_w[k], vi, sA[k], sB[k]) - rew;
					if(cand < best) { best = cand; bk = k; }
				}
				// a node without an admissible candidate keeps the control of a zero-filled vector
				// (PolicyIteration.hpp:64-66 hands setInputs a fresh domain->vector())
				double zsel, rew, wsel; int isel;
				if(bk >= 0) { zsel = sz[bk]; isel = z_idx[bk]; wsel = z_w[bk]; }
				else { zsel = 0.; interp_nodes<M>(sX, N, mdl.transition(xi, 0.), isel, wsel); }
				rew = mdl.impulse_flow(xi, zsel);
				const double lhs = impulse_lhs(i, isel, wsel, vi, (0. - wsel) * sx[slot<M>(isel)], (0. - (-wsel + 1.)) * sx[slot<M>(isel + 1)]);
				// 3. penalty: active iff scale * ((I - M*) x - reward*) < 0 (PenaltyMethod.hpp:45-68)
				const bool act = (pen * (lhs - rew)) < 0.;
				i_k[at] = isel; z_k[at] = (unsigned short)(bk + 1); i_rew[at] = rew; i_w[at] = wsel;
				// 4. row of (I + h0 L^q + P (I - M*)): entries on the three diagonals folded in
				double a_ = blo * h0, b_ = 1. + bdi * h0, c_ = bup * h0;
				unsigned fl = 0;                // bit 0: penalised; bits 1, 2: entry idx / idx + 1 is off the three diagonals
				if(act) {
					fl = 1u;
					b_ += pen;
					const double f0 = pen * wsel, f1 = pen * (-wsel + 1.);
					if(isel == i) b_ -= f0; else if(isel == i - 1) a_ -= f0; else if(isel == i + 1) c_ -= f0; else fl |= 2u;
					if(isel + 1 == i) b_ -= f1; else if(isel + 1 == i - 1) a_ -= f1; else if(isel + 1 == i + 1) c_ -= f1; else fl |= 4u;
				}
				h_a[at] = a_; h_b[at] = b_; h_c[at] = c_; h_f[at] = (unsigned char)fl;
			}
			__syncthreads();
#pragma unroll
			for(int j = 0; j < HM; ++j) {
				const int at = j * HP + tid;
				const unsigned fl = h_f[at];
				aa[j] = h_a[at]; bb[j] = h_b[at]; cc[j] = h_c[at];
				active |= (fl & 1u) << j; far |= (fl >> 1) << (2 * j);
			}
			ps.factorize(aa, bb, cc, sm.ps, lane, warp);
			// ---- linear solve: lagged far entries, repeated to a 1e-15 relative update
			int sweep = 0;
			bool unsettled;
			do {
				double r[HM], y[HM], x_next;
#pragma unroll
				for(int j = 0; j < HM; ++j) {
					const int at = j * HP + tid;
					double r_ = 0.;
					if(i0 + j < N) {
						r_ = v0[j] + h0 * r_b[at];
						if(active >> j & 1u) {
							r_ += pen * i_rew[at];
							const int isel = i_k[at]; const double wsel = i_w[at];
							if(far >> (2 * j) & 1u) r_ += (pen * wsel) * sx[slot<M>(isel)];
							if(far >> (2 * j + 1) & 1u) r_ += (pen * (-wsel + 1.)) * sx[slot<M>(isel + 1)];
						}
					}
					r[j] = r_;
				}
				ps.solve(r, y, x_next, sm.ps, lane, warp);
				bool big = false;
#pragma unroll
				for(int j = 0; j < HM; ++j) {
					const double d = fabs(y[j] - x[j]);
					const double den = fabs(y[j]) > 1. ? fabs(y[j]) : 1.;
					big = big || d / den > 1e-15;
					x[j] = y[j];
				}
				// (the barrier inside solve() separates this sweep's gathers from these stores)
#pragma unroll
				for(int j = 0; j < HM; ++j) sx[j * HP + tid] = x[j];
				unsettled = __syncthreads_or(big ? 1 : 0) != 0;
				++sweep;
			} while(unsettled && sweep < b.max_sweeps);
			if(unsettled && tid == 0) sm.status = 2;
			if(tid == 0) sm.sweep_sum += sweep;
			++its;
			// relativeError(iterand(0), iterand(1)) > tolerance (IterativeMethod.hpp:1125-1137, scale 1)
			bool nd = false;
#pragma unroll
			for(int j = 0; j < HM; ++j) {
				const double fa = fabs(x[j]), fb = fabs(xprev[j]);
				double den = fa < fb ? fb : fa;
				den = 1. < den ? den : 1.;
				nd = nd || fabs(x[j] - xprev[j]) / den > b.tolerance;
			}
			notdone = __syncthreads_or(nd ? 1 : 0) != 0;
			if(notdone && its >= b.max_inner) { if(tid == 0 && sm.status == 0) sm.status = 1; notdone = false; }
		} while(notdone);
		if(tid == 0) { sm.inner_sum += its; sm.n_steps += 1; }
		time1 = t;
	}

	// ---- outputs: the controls of the last iteration, PenaltyMethod::constraintMask on the final
	// iterand (PenaltyMethod.hpp:127-139), NaN where the other control acts (HJBQVI.hpp:977-990)
	__syncthreads();
#pragma unroll
	for(int j = 0; j < HM; ++j) {
		const int i = i0 + j, at = j * HP + tid;
		if(i >= N) continue;
		const size_t o = (size_t)pidx * N + i;
		const int isel = i_k[at]; const double wsel = i_w[at];
		const double lhs = impulse_lhs(i, isel, wsel, x[j], (0. - wsel) * sx[slot<M>(isel)], (0. - (-wsel + 1.)) * sx[slot<M>(isel + 1)]);
		const bool act = lhs - i_rew[at] < 0.;
		if(out.V) out.V[o] = x[j];
		if(out.X) out.X[o] = sX[at];
		if(out.mask) out.mask[o] = act ? 1 : 0;
		const double qnan = __longlong_as_double(0x7ff8000000000000LL);      // std::nan(""): the positive quiet NaN
		if(out.Q) out.Q[o] = act ? qnan : sq[q_k[at]];
		if(out.Z) out.Z[o] = act ? (z_k[at] ? sz[z_k[at] - 1] : 0.) : qnan;
	}
	if(tid == 0) {
		if(out.n_steps) out.n_steps[pidx] = sm.n_steps;
		if(out.mean_inner) out.mean_inner[pidx] = (double)sm.inner_sum / (double)sm.n_steps;
		if(out.mean_sweeps) out.mean_sweeps[pidx] = (double)sm.sweep_sum / (double)sm.inner_sum;
		if(out.status) out.status[pidx] = sm.status;
	}
}

static int rows_per_thread(int N) {
	const int choices[5] = { 2, 3, 5, 8, 9 };
	for(int m : choices) if(N <= m * HP) return m;
	return 0;
}

size_t hjbqvi1d_shared_bytes(int N, int nq, int nz) { return hjbqvi1d_smem_bytes(rows_per_thread(N) ? rows_per_thread(N) : 9, nq, nz); }

int launch_hjbqvi1d(const Hjbqvi1dBatch &b, const Hjbqvi1dResult &out, cudaStream_t stream, long long *launches) {
	if(b.N < 3 || b.N > 9 * HP || b.model != QPDE_HJBQVI1D_EXCHANGE_RATE) return QPDE_ERR_UNSUPPORTED;
	const size_t bytes = hjbqvi1d_shared_bytes(b.N, b.nq, b.nz);
	if(bytes > 227 * 1024) return QPDE_ERR_UNSUPPORTED;
	void (*kern)(Hjbqvi1dBatch, Hjbqvi1dResult) = nullptr;
	switch(rows_per_thread(b.N)) {
	case 2: kern = k_hjbqvi1d<ExchangeRate, 2>; break;
	case 3: kern = k_hjbqvi1d<ExchangeRate, 3>; break;
	case 5: kern = k_hjbqvi1d<ExchangeRate, 5>; break;
	case 8: kern = k_hjbqvi1d<ExchangeRate, 8>; break;
	default: kern = k_hjbqvi1d<ExchangeRate, 9>; break;
	}
	if(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return QPDE_ERR_CUDA;
	kern<<<b.n_problems, HP, bytes, stream>>>(b, out);
	*launches += 1;
	return cudaGetLastError() == cudaSuccess ? QPDE_OK : QPDE_ERR_CUDA;
}

} // namespace qpde
